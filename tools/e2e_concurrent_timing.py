"""Stage breakdown of concurrent dxo.encode() calls: T caller threads (pinned host buffers) with DXO_TIMING=1, the library's
per-stage laps (wall clock inside each call) summed per label and divided by the number of calls.
usage: python tools/e2e_concurrent_timing.py [threads=16] [steps=3] [workload=config2]"""
import collections, os, re, subprocess, sys, threading, time
here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, here)
    import torch
    import draco_oxide_b200 as dxo
    import bench
    T, steps, workload = int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    mesh, _ = bench.make_mesh(workload)
    pmesh = bench.pinned_copy(mesh)
    cfg = dxo.Config(device=0)
    ready, go = threading.Barrier(T + 1), threading.Barrier(T + 1)
    tms, walls = [], []
    def worker():
        for _ in range(2):
            o = bytearray(); dxo.encode(pmesh, o, cfg)
        ready.wait(); go.wait()
        for _ in range(steps):
            t = time.perf_counter()
            o = bytearray(); dxo.encode(pmesh, o, cfg)
            walls.append(time.perf_counter() - t)
            tms.append(dxo.last_timing())
    th = [threading.Thread(target=worker) for _ in range(T)]
    for t in th: t.start()
    ready.wait(); torch.cuda.synchronize()
    print("[mark] timed region begins", file=sys.stderr, flush=True)
    t0 = time.perf_counter(); go.wait()
    for t in th: t.join()
    dt = time.perf_counter() - t0
    keys = [k for k in tms[0] if k.endswith("_ms")]
    print("[result] per call, from dxo_last_timing: " + "  ".join(f"{k} {sum(t[k] for t in tms) / len(tms):.2f}" for k in keys) + f"  python wall {1e3 * sum(walls) / len(walls):.2f}", file=sys.stderr)
    print(f"[result] {T} callers x {steps} calls: {dt * 1e3:.1f} ms, {mesh.num_points() * T * steps / dt / 1e6:.1f} Mvertices/s, {dt * 1e3 / steps:.1f} ms per call", file=sys.stderr)
    sys.exit(0)
T = int(sys.argv[1]) if len(sys.argv) > 1 else 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
workload = sys.argv[3] if len(sys.argv) > 3 else "config2"
env = dict(os.environ, DXO_TIMING="1")
p = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", str(T), str(steps), workload], env=env, stderr=subprocess.PIPE, text=True)
lines = p.stderr.splitlines()
start = max(i for i, l in enumerate(lines) if l.startswith("[mark]")) if any(l.startswith("[mark]") for l in lines) else 0
agg = collections.OrderedDict()
for l in lines[start:]:
    m = re.match(r"\[dxo\] (.*?)\s+([0-9.]+) ms$", l)
    if m and not m.group(1).lstrip().startswith(("side stream", "thread ")):
        k = m.group(1).rstrip()
        a = agg.setdefault(k, [0.0, 0])
        a[0] += float(m.group(2)); a[1] += 1
    elif l.startswith("[result]"):
        print(l)
calls = T * steps
for k, (s, n) in agg.items():
    print(f"{k:44s} {s / calls:8.2f} ms per call   ({n} laps)")
