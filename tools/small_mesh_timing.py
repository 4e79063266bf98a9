import os, sys, time
sys.path.insert(0, '/root/repo')
os.environ["DXO_TIMING"] = "1"
import draco_oxide_b200 as dxo
from draco_oxide_b200 import synth
m = synth.grid_mesh(150, 145, 5)
for r in range(4):
    out = bytearray(); t0 = time.perf_counter(); dxo.encode(m, out); dt = time.perf_counter() - t0
    t = dxo.last_timing()
    print(f"---- wall {dt*1e3:.2f} ms total {t['total_ms']:.2f} host {t['host_connectivity_ms']:.2f} h2d {t['h2d_ms']:.2f} device {t['device_ms']:.2f} d2h {t['d2h_ms']:.2f} launches {t['num_launches']}", file=sys.stderr)
