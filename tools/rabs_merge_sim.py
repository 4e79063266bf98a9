"""CPU simulation: do binary rANS (rABS, precision 8) trajectories started from G spread states merge with the true one?
Uses the real flip / orientation bit streams of config 2 (oracle trace; test infrastructure)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import orc
from draco_oxide_b200 import synth

orc.build()
m = synth.config2_mesh(int(sys.argv[1]) if len(sys.argv) > 1 else 1000)
G = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
_, tr = orc.encode(m, trace=True)
for att in (1, 2):
    bits = tr.get(f"att{att}.side_bits", np.uint8).astype(np.int64)
    n = bits.size
    zeros = int((bits == 0).sum())
    p0 = int(np.clip(np.float32(np.float32(zeros) / np.float32(n)) * np.float32(256.0) + np.float32(0.5), 1, 255))
    f = np.array([p0, 256 - p0], np.int64)
    cum = np.array([256 - p0, 0], np.int64)
    print(f"att{att}: {n} bits, zero_prob {p0}/256")
    W = 4096
    stride = 4096
    starts = np.arange(W + stride, n - 1, stride)[:200]
    # true trajectory
    x = 4096
    want = set((starts - W).tolist()); true_at = {}
    bl = bits.tolist()
    for e in range(int(starts[-1])):
        if e in want: true_at[e] = x
        b = bl[e]; fb = int(f[b])
        if x >= fb << 12: x >>= 8
        x = (x // fb << 8) + x % fb + int(cum[b])
    K = len(starts)
    X = np.empty((K, G + 1), np.int64)
    X[:, 0] = [true_at[int(s - W)] for s in starts]
    X[:, 1:] = (4096 * (256.0 ** ((np.arange(G) + 0.5) / G))).astype(np.int64)[None, :]
    idx = (starts - W)
    for t in range(W):
        b = bits[idx + t]
        fb = f[b][:, None]; cb = cum[b][:, None]
        X = np.where(X >= (fb << 12), X >> 8, X)
        X = ((X // fb) << 8) + X % fb + cb
        if t + 1 in (64, 256, 1024, 2048, 4096):
            miss = np.mean(~(X[:, 1:] == X[:, :1]).any(axis=1))
            clusters = np.mean([len(set(r[1:].tolist())) for r in X])
            print(f"   warm-up {t + 1:5d}: all {G} guesses miss {miss:6.3f}   distinct guess states {clusters:7.1f}")
