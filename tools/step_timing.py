"""Prints the host/device split of one resident-session step on config 2 (DXO_TIMING=1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DXO_TIMING"] = "1"
import draco_oxide_b200 as dxo
from draco_oxide_b200 import synth
m = synth.config2_mesh()
s = dxo.Session(m)
for _ in range(3): s.run(want_bytes=False)
print("---- steady state ----", file=sys.stderr)
s.run(want_bytes=False)
print("---- one-shot encode ----", file=sys.stderr)
out = bytearray(); dxo.encode(m, out)
out = bytearray(); dxo.encode(m, out)
print(dxo.last_timing())
