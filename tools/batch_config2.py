"""e2e of config 2 through the batch entry: N copies of the 1M-vertex mesh in one dxo_encode_batch call (host buffers in)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import draco_oxide_b200 as dxo
from draco_oxide_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
m = synth.config2_mesh()
ms = [m] * n
ref = bytearray(); dxo.encode(m, ref)
for r in range(4):
    t = time.perf_counter()
    out = dxo.encode_batch(ms)
    dt = time.perf_counter() - t
    assert all(o == bytes(ref) for o in out)
    print(f"rep {r}: {n} meshes {dt*1e3:.1f} ms  {n * m.num_points() / dt / 1e6:.1f} Mvertices/s", flush=True)
