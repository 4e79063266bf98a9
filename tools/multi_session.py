"""S concurrent resident sessions of config 2: per-session CUDA-event time of K steps (diagnostics)."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import draco_oxide_b200 as dxo
from draco_oxide_b200 import synth
S = int(sys.argv[1]) if len(sys.argv) > 1 else 4
K = int(sys.argv[2]) if len(sys.argv) > 2 else 30
mesh = synth.config2_mesh()
res = [None] * S
ready, go = threading.Barrier(S + 1), threading.Barrier(S + 1)
def worker(i):
    s = dxo.Session(mesh)
    s.run_steps(3)
    ready.wait(); go.wait()
    t0 = time.perf_counter()
    ms, _ = s.run_steps(K)
    res[i] = (ms / K, (time.perf_counter() - t0) * 1e3 / K)
    s.close()
th = [threading.Thread(target=worker, args=(i,)) for i in range(S)]
[t.start() for t in th]
ready.wait(); go.wait()
[t.join() for t in th]
print("S", S, "per-session ms/step (events, wall):", [(round(a, 2), round(b, 2)) for a, b in res])
