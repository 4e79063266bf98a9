"""Workload for compute-sanitizer (memcheck / racecheck), logs under profiles/:
  * a 74-chunk rANS stream through every K10 kernel,
  * a 31k-face mesh through the whole per-mesh path,
  * the batch entry: a group of 12 meshes (pos-only, seams, point maps) through the segmented launches,
  * with `sessions N`: N resident sessions stepping concurrently from N threads (the process-wide K10 role counters,
    g_rans_consumers / g_rans_lock, shared by all streams)."""
import os, sys, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import draco_oxide_b200 as dxo
from draco_oxide_b200 import synth

mode = sys.argv[1] if len(sys.argv) > 1 else "single"
if mode == "single":
    rng = np.random.default_rng(1)
    sym = np.minimum(rng.geometric(0.05, 300_000) - 1, 2000).astype(np.uint32)
    print("symbols", len(dxo.encode_symbols(sym)))
    out = bytearray(); dxo.encode(synth.torus_mesh(125, 125, 3), out); print("mesh", len(out))
    counts = synth.batch_vertex_counts(12, 200, 6000, seed=3)
    ms = [synth.batch_mesh(k, int(c)) for k, c in enumerate(counts)] + [synth.grid_mesh(20, 20, 1, with_normals=False, with_uvs=False)]
    print("batch", sum(len(b) for b in dxo.encode_batch(ms)))
else:
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    mesh = synth.grid_mesh(150, 150, 7)
    def work(i):
        s = dxo.Session(mesh)
        ref = s.run()
        s.run_steps(3)
        assert s.run() == ref
        s.close()
    th = [threading.Thread(target=work, args=(i,)) for i in range(n)]
    [t.start() for t in th]; [t.join() for t in th]
    print("sessions", n, "ok")
