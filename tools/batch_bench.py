"""Config-4 slice through dxo_encode_batch: Mvertices/s and meshes/s (wall clock, host buffers in, streams out).
usage: python tools/batch_bench.py [num_meshes=1024] [reps=3] [num_gpus=1]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import draco_oxide_b200 as dxo
from draco_oxide_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
gpus = int(sys.argv[3]) if len(sys.argv) > 3 else 1
counts = synth.batch_vertex_counts()[:n]
t = time.perf_counter()
ms = synth.batch_meshes(counts)
verts = sum(m.num_points() for m in ms)
print(f"{n} meshes, {verts} vertices generated in {time.perf_counter() - t:.1f} s", flush=True)
batch = dxo.Batch(ms)
for r in range(reps):
    t = time.perf_counter()
    out = dxo.encode_batch(batch, first_gpu=0, num_gpus=gpus)
    dt = time.perf_counter() - t
    print(f"rep {r}: {dt * 1e3:.1f} ms  {verts / dt / 1e6:.1f} Mvertices/s  {n / dt:.0f} meshes/s  ({sum(len(o) for o in out)} bytes)", flush=True)
