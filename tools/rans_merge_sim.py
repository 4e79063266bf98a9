"""CPU simulation of rANS state merging on the real symbol streams of config 2 (test infrastructure: uses the oracle
trace). For chunk starts every `stride` steps it runs the true trajectory and G guess trajectories for W warm-up steps
and reports, per warm-up length, the fraction of chunks whose true state is NOT matched by (a) guess 0, (b) any guess."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import orc
from draco_oxide_b200 import synth

orc.build()
m = synth.config2_mesh(int(sys.argv[1]) if len(sys.argv) > 1 else 1000)
_, tr = orc.encode(m, trace=True)
G = 32
for att in range(3):
    sym = tr.get(f"att{att}.symbols", np.uint32)[::-1].astype(np.int64)  # coding order: last symbol first
    dist = tr.get(f"att{att}.distribution", np.uint64).astype(np.int64)
    P = int(tr.get(f"att{att}.bit_length", np.uint32)[1])
    cum = np.concatenate([[0], np.cumsum(dist)[:-1]])
    f_all, c_all = dist[sym], cum[sym]
    n = sym.size
    l_base = 4 << P
    # true trajectory, sampled where chunks start
    W = 2048
    stride = 1024
    starts = np.arange(W + stride, n - 1, stride)[:2500]
    # true states at (start - W): run the real coder once (vectorised over nothing: plain loop in numpy scalars is slow -> chunked C-like loop)
    x = l_base
    want = set((starts - W).tolist())
    true_at = {}
    last = int(starts[-1])
    fa, ca = f_all.tolist(), c_all.tolist()
    for e in range(last):
        if e in want:
            true_at[e] = x
        f = fa[e]
        thr = f << 10
        while x >= thr:
            x >>= 8
        x = (x // f << P) + x % f + ca[e]
    K = len(starts)
    X = np.empty((K, G + 1), np.int64)
    X[:, 0] = [true_at[int(s - W)] for s in starts]
    # guesses: spread geometrically over [l_base, 256 l_base)
    g = (l_base * (256.0 ** ((np.arange(G) + 0.5) / G))).astype(np.int64)
    g[0] = l_base
    X[:, 1:] = g[None, :]
    report_at = [64, 128, 256, 512, 1024, 2048]
    idx = (starts - W)[:, None]
    for t in range(W):
        f = f_all[idx + t]
        c = c_all[idx + t]
        thr = f << 10
        for _ in range(3):
            big = X >= thr
            X = np.where(big, X >> 8, X)
        X = ((X // f) << P) + X % f + c
        if t + 1 in report_at:
            miss0 = np.mean(X[:, 1] != X[:, 0])
            miss_any = np.mean(~(X[:, 1:] == X[:, :1]).any(axis=1))
            clusters = np.mean([len(set(r[1:].tolist())) for r in X])
            print(f"att{att} P={P} warm-up {t + 1:6d}: guess0 misses {miss0:6.3f}   all {G} guesses miss {miss_any:6.3f}   distinct guess states {clusters:5.2f}")
