"""Micro-benchmark of the entropy stage (K8 histogram, K9 table, K10 rANS) through
dxo_encode_symbols on a position-residual-like symbol stream."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import draco_oxide_b200 as dxo

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rng = np.random.default_rng(1)
sym = np.minimum(rng.geometric(0.08, n) - 1, 4000).astype(np.uint32)
for r in range(reps):
    data, t = dxo.encode_symbols(sym, timing=True)
    print(f"n={n} bytes={len(data)} hist={t['histogram_ms']:.3f} ms table={t['table_ms']:.3f} ms rans={t['rans_ms']:.3f} ms "
          f"-> {t['rans_ms'] * 1e6 / n:.2f} ns/symbol")
