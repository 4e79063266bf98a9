"""Throughput of dxo_encode_batch on a slice of BASELINE config 4 (primitives with log-uniform vertex counts in
[1k, 100k], alternating grid patches and tori): python tools/batch_bench.py [num_meshes] [num_gpus]
Prints Mvertices/s for several DXO_WORKERS_PER_GPU values and checks every stream against per-mesh dxo.encode."""
import os, sys, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
gpus = int(sys.argv[2]) if len(sys.argv) > 2 else 1
if len(sys.argv) > 3:  # child: one measurement with the current environment
    import draco_oxide_b200 as dxo
    from draco_oxide_b200 import synth
    counts = synth.batch_vertex_counts()[:n]
    meshes = [synth.batch_mesh(k, int(v)) for k, v in enumerate(counts)]
    total = sum(m.num_points() for m in meshes)
    dxo.encode_batch(meshes[: min(n, 16)], num_gpus=gpus)  # warm-up (context, pools)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        outs = dxo.encode_batch(meshes, num_gpus=gpus)
        best = min(best, time.perf_counter() - t0)
    if sys.argv[3] == "check":
        for k in range(0, n, max(1, n // 16)):
            ref = bytearray(); dxo.encode(meshes[k], ref)
            assert bytes(ref) == outs[k], k
    print(f"workers/gpu={os.environ.get('DXO_WORKERS_PER_GPU', 'default'):>7}  meshes={n} vertices={total} gpus={gpus}  {best * 1e3:8.1f} ms  {total / best / 1e6:7.2f} Mvertices/s")
    sys.exit(0)
for w in ["default", "2", "4", "8", "12", "16"]:
    env = dict(os.environ)
    if w != "default":
        env["DXO_WORKERS_PER_GPU"] = w
    subprocess.run([sys.executable, __file__, str(n), str(gpus), "check" if w == "default" else "run"], env=env, check=True)
