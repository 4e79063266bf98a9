"""Two resident-session steps of config 2 — the command profiled under ncu (profiles/)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import draco_oxide_b200 as dxo
from draco_oxide_b200 import synth
s = dxo.Session(synth.config2_mesh())
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    s.run(want_bytes=False)
s.close()
