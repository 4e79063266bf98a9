"""End-to-end throughput of concurrent dxo_encode() calls (host buffers in pinned memory) against the number of caller
threads: python tools/e2e_threads.py [workload] [threads ...]"""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import draco_oxide_b200 as dxo
import bench

workload = sys.argv[1] if len(sys.argv) > 1 else "config2"
counts = [int(a) for a in sys.argv[2:]] or [1, 2, 4, 6, 8, 12, 16]
mesh, desc = bench.make_mesh(workload)
pmesh = bench.pinned_copy(mesh)
cfg = dxo.Config(device=0)
V = mesh.num_points()
out = bytearray(); dxo.encode(pmesh, out, cfg)
ref = bytes(out)
steps = 4
for T in counts:
    ready, go = threading.Barrier(T + 1), threading.Barrier(T + 1)
    bad = []
    def worker():
        o = bytearray(); dxo.encode(pmesh, o, cfg)   # this thread's streams and staging buffers
        ready.wait(); go.wait()
        for _ in range(steps):
            o = bytearray(); dxo.encode(pmesh, o, cfg)
            if bytes(o) != ref: bad.append(1)
    th = [threading.Thread(target=worker) for _ in range(T)]
    for t in th: t.start()
    ready.wait(); torch.cuda.synchronize()
    t0 = time.perf_counter(); go.wait()
    for t in th: t.join()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"threads={T:2d}  {1e3 * dt / steps:7.1f} ms per step of {T} meshes  {V * T * steps / dt / 1e6:7.1f} Mvertices/s  mismatches={len(bad)}", flush=True)
