"""Device step time of config 2 plus the speculative-rANS counters (DXO_RANS_DEBUG=1) for the current
DXO_RANS_CHUNK / DXO_RANS_WARMUP / DXO_RANS_SUB (unset: the adaptive chunk size)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import draco_oxide_b200 as dxo
from draco_oxide_b200 import synth
m = synth.config2_mesh()
s = dxo.Session(m)
for _ in range(2): s.run(want_bytes=False)
dxo.set_profiling(True)
s.run(want_bytes=False)
t = dxo.last_timing()
dxo.set_profiling(False)
k10 = [k for k in t["kernels"] if k["name"].startswith("K10")]
print("K10 ms per attribute:", [round(k["ms"], 3) for k in k10])
ms, launches = s.run_steps(5)
print(f"step {ms / 5:.3f} ms")
