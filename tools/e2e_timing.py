"""One-shot dxo.encode() timing with the DXO_TIMING breakdown: python tools/e2e_timing.py [config2|config3|torus_quarter] [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DXO_TIMING"] = "1"
os.environ["DXO_RANS_DEBUG"] = "1"
import draco_oxide_b200 as dxo
from draco_oxide_b200 import synth
which = sys.argv[1] if len(sys.argv) > 1 else "config2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
m = {"config2": synth.config2_mesh, "config3": synth.config3_mesh, "torus_quarter": lambda: synth.torus_mesh(1000, 1250, 3)}[which]()
print(which, "points", m.num_points(), "faces", m.faces.shape[0], file=sys.stderr)
for r in range(reps):
    print(f"---- encode {r} ----", file=sys.stderr)
    out = bytearray()
    t0 = time.perf_counter()
    dxo.encode(m, out)
    dt = time.perf_counter() - t0
    t = dxo.last_timing()
    print(f"wall {dt * 1e3:.1f} ms  total {t['total_ms']:.1f}  host {t['host_connectivity_ms']:.1f}  h2d {t['h2d_ms']:.1f}  device {t['device_ms']:.2f}  d2h+side {t['d2h_ms']:.1f}  "
          f"-> {m.num_points() / dt / 1e6:.2f} Mpoints/s, {len(out)} bytes", file=sys.stderr)
