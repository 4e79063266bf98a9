import os, sys
sys.path.insert(0, '/root/repo')
import draco_oxide_b200 as dxo
from draco_oxide_b200 import synth
s = dxo.Session(synth.config2_mesh())
for _ in range(3): s.run(want_bytes=False)
for mode in (1, 2, 2):
    dxo.set_profiling(mode)
    s.run(want_bytes=False)
    t = dxo.last_timing()
    print("mode", mode, [(k["name"][:3], round(k["ms"] * 1e3, 1)) for k in t["kernels"]])
