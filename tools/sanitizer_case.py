import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
import draco_oxide_b200 as dxo
from draco_oxide_b200 import synth
rng = np.random.default_rng(1)
sym = np.minimum(rng.geometric(0.08, 300_000) - 1, 4000).astype(np.uint32)
a = dxo.encode_symbols(sym)
m = synth.grid_mesh(120, 130, 77)
out = bytearray(); dxo.encode(m, out)
print("ok", len(a), len(out))
