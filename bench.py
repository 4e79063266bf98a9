#!/usr/bin/env python
"""bench.py — Mvertices/s of the attribute-encoding hot path on BASELINE.json's configs (default: config 2).

  python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload config2|config3|config4|config5]

Workloads: config2 = the headline (1M-vertex grid; also what the driver runs), config3 = 10M-triangle torus
with uv seams, config4 = 4096 primitives (1k-100k vertices) through the batch entry dxo_encode_batch, sharded over the
ranks (strong scaling, meshes/s), config5 = qp sweep on config 2's mesh. The default run also appends a `config4` block
(the batch entry on this job's GPUs) to the config-2 line, so every driver run measures the sharded batch as well.

A "step" = one pass of the hot path over one synthetic mesh (config 2: 1000x1000 grid,
1 000 000 vertices, 1 996 002 triangles, positions + normals + texcoords, qp/qn/qt
11/8/10). One process per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE); meshes are
independent units, so ranks share nothing on the data path (weak scaling, no collective).

value  : vertices/s with every input of the attribute kernels resident in HBM
         (dxo_session_run_steps: quantize -> predict -> symbolize -> histogram -> table
         -> rANS -> D2H of the results -> stream assembly), CUDA-event timed.
e2e    : the same metric through dxo_encode() from pinned HOST buffers: host
         connectivity (corner tables, Edgebreaker, sequencer), H2D, kernels, D2H, assembly.
--impl reference : the CPU oracle (C++ restatement of the reference encoder; the Rust
         reference cannot be built in this image) on the host cores, same metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mvertices/s encoded"
PEAKS_FALLBACK_GBS = 6650.0  # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent
# dram__bytes_read.sum + dram__bytes_write.sum per launch on config 2, from the `ncu --set full` capture summarised in
# profiles/r1_ncu_full_final_summary.md (bytes)
NCU_TRAFFIC_CONFIG2 = {"K4_predict_parallelogram": 110.7e6, "K5_predict_normal": 145.9e6, "K6_predict_texcoord": 102.7e6}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config2", choices=["config1", "config2", "config2_small", "config3", "config4", "config5"])
    ap.add_argument("--batch-meshes", type=int, default=4096, help="primitives of config 4 used by the config4 workload / block (4096 = all)")
    ap.add_argument("--no-config4", action="store_true", help="skip the config4 block of the default run")
    ap.add_argument("--sessions", type=int, default=0,
                    help="concurrent resident sessions per GPU for `value`; 0 = one per three host threads available to this GPU, "
                         "at most 6 (each session keeps a main thread and two side-stream coders busy)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph-replay", default="auto", choices=["auto", "on", "off"],
                    help="resident sessions replay their step as one CUDA graph (auto: when host threads are scarce)")
    ap.add_argument("--e2e-callers", type=int, default=0,
                    help="host threads calling dxo_encode() concurrently in the end-to-end arm (0 = the host threads this GPU can count on)")
    return ap.parse_args()


def host_threads():
    """Host threads this process may use (the affinity mask, so that `taskset` runs emulate a box with fewer threads)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def default_sessions(world):
    """(sessions per GPU, replay the step as a CUDA graph?) from the host threads this GPU can count on. A session is ONE host
    thread: it replays its step as one CUDA graph, spins in its waits and codes both side streams itself in one interleaved
    loop (DXO_SIDE_INLINE) — measured on one B200 (tools/session_modes_sweep.sh, Mvertices/s):
      16 threads: helpers + no graph 8 / 12 / 16 / 24 sessions 1981 / 2053 / 2021 / 2101; inline + graph 2186 / 2255 / 2356 / 2489
       8 threads: inline + graph, spinning 6 / 12 / 24 sessions 2354 / 2370 / 2363; sleeping waits 1912 / 2349 / 2315
       4 threads: inline + graph, spinning 1607 / 1564 / 1599; sleeping waits 1317 / 1305 / 1622
    so: one and a half sessions per thread (6 … 24), spinning waits unless DXO_BLOCKING_WAIT is set by the caller."""
    per_gpu = max(1, host_threads() // max(1, world))
    return max(6, min(24, per_gpu * 3 // 2)), True


def make_mesh(workload):
    from draco_oxide_b200 import synth
    if workload == "config1":
        return synth.config1_mesh(), "config1: 101x251 grid, 25 351 vertices, 50 000 triangles, pos+normal+uv, 11/8/10 bits"
    if workload == "config2_small":
        return synth.config2_mesh(300), "config2_small: 300x300 grid (debug)"
    if workload == "config3":
        return synth.config3_mesh(), "config3: 2000x2500-quad torus, 5 000 000 position vertices (5 004 501 points), 10 000 000 triangles, uv seams on both cuts, pos+normal+uv, 11/8/10 bits"
    return synth.config2_mesh(), "config2: 1000x1000 grid, 1 000 000 vertices, 1 996 002 triangles, pos+normal+uv, Edgebreaker, qp/qn/qt=11/8/10"


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return PEAKS_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this workload
    (profiles/ncu_traffic.json: {workload: {kernel: bytes}}, written by tools/ncu_traffic.py); the round-1 constants otherwise."""
    try:
        return {k: float(v) for k, v in json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(workload, {}).items()}
    except Exception:
        return dict(NCU_TRAFFIC_CONFIG2) if workload == "config2" else {}


def oracle_stage_split(orc, mesh):
    """One single-thread oracle encode with ORC_TIMING=1: wall clock of the host-only stages (corner tables, Edgebreaker,
    sequencer) and of the attribute stages (quantize, predict, transform, entropy, side streams) — BASELINE.md §3."""
    import re
    import tempfile
    os.environ["ORC_TIMING"] = "1"
    sys.stderr.flush()
    saved = os.dup(2)
    with tempfile.TemporaryFile(mode="w+b") as tmp:
        os.dup2(tmp.fileno(), 2)
        try:
            orc.encode(mesh)
        finally:
            os.dup2(saved, 2)
            os.close(saved)
            os.environ.pop("ORC_TIMING", None)
        tmp.seek(0)
        text = tmp.read().decode(errors="replace")
    host = attr = 0.0
    for name, ms in re.findall(r"\[orc\]\s+(.+?)\s+([0-9.]+) ms", text):
        name = name.strip()
        if name in ("corner table", "attribute corner tables", "edgebreaker", "sequencer"):
            host += float(ms)
        elif name in ("portabilize", "predict", "transform", "entropy (hist+table+rANS)", "metadata (rABS)"):
            attr += float(ms)
    return {"host_ms": host, "attribute_ms": attr, "note": "one oracle encode on one thread: host = corner tables + Edgebreaker + sequencers, attribute = quantize + predict + transform + entropy + side streams"}


class Config4Shard:
    """This rank's primitives of config 4: the deterministic longest-processing-time-first assignment of
    draco_oxide_b200.sharding over the ranks, on the vertex counts (known before the meshes exist).
    The numpy worker processes are forked when the object is made — before this process holds a CUDA context — and stay
    idle until generate() is called, so that neither their memory traffic nor the ~5 GB of generated meshes sit next to the
    measurements that come first (measured: the config-2 e2e arm loses a quarter with them resident)."""

    def __init__(self, n_meshes, rank, world):
        import multiprocessing as mp
        import numpy as np
        from draco_oxide_b200 import sharding, synth
        counts = synth.batch_vertex_counts()[:n_meshes]
        mine = np.asarray(sharding.my_shard(counts, rank, world), dtype=np.int64)
        mine = mine[np.argsort(-counts[mine], kind="stable")]  # longest first inside the shard as well
        # a shard is not a consecutive range of primitives: generate by index
        self.jobs = [(int(k), int(counts[k])) for k in mine]
        self.total_vertices, self.total_meshes = int(counts.sum()), int(counts.size)
        procs = max(1, min(16, host_threads() // max(1, world)))
        self.pool = mp.get_context("fork").Pool(procs) if procs > 1 and len(self.jobs) >= 8 else None

    def generate(self):
        from draco_oxide_b200 import synth
        if self.pool is not None:
            meshes = self.pool.map(synth._batch_mesh_job, self.jobs, chunksize=4)
            self.pool.close()
            self.pool.join()
            self.pool = None
        else:
            meshes = [synth._batch_mesh_job(j) for j in self.jobs]
        return meshes, self.total_vertices, self.total_meshes


def config4_shard(n_meshes, rank, world):
    return Config4Shard(n_meshes, rank, world).generate()


def measure_config4(meshes, local_rank, reps, dist, torch, dxo):
    """The batch entry on this rank's shard: wall clock of dxo_encode_batch (host buffers in, streams out), max over ranks."""
    sampler = ClockSampler(local_rank)
    sampler.start()
    batch = dxo.Batch(meshes)  # the dxo_mesh[] array the C ABI takes (plain pointers into the host buffers), built once
    out = dxo.encode_batch(batch, first_gpu=local_rank, num_gpus=1)  # warm-up: slabs, pinned blocks, contexts
    out = dxo.encode_batch(batch, first_gpu=local_rank, num_gpus=1)
    times = []
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.mark_begin()
    for _ in range(reps):
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        out = dxo.encode_batch(batch, first_gpu=local_rank, num_gpus=1)
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t.item()))
    clocks = sampler.stop()
    in_bytes = sum(m.faces.nbytes + sum(a.values.nbytes + (a.point_to_value.nbytes if a.point_to_value is not None else 0) for a in m.attributes) for m in meshes)
    out_bytes = sum(len(o) for o in out)
    tot = torch.tensor([float(in_bytes), float(out_bytes), float(sum(m.num_points() for m in meshes)), float(len(meshes))], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    return times, [float(x) for x in tot.tolist()], clocks, out


def config4_cpu_baseline(orc, meshes, sample_every=16):
    """The oracle on every host thread over a bounded sample of the shard (every 16th primitive of the size-sorted list)."""
    from concurrent.futures import ThreadPoolExecutor
    sample_every = max(1, min(sample_every, len(meshes) // 256))  # about 256 primitives, whatever the shard's size
    sample = meshes[::sample_every]
    cores = max(1, min(64, host_threads()))
    verts = sum(m.num_points() for m in sample)
    with ThreadPoolExecutor(cores) as ex:
        list(ex.map(orc.encode, sample[: cores]))  # warm-up
        t0 = time.perf_counter()
        list(ex.map(orc.encode, sample))
        dt = time.perf_counter() - t0
    return {"value": verts / dt / 1e6, "unit": "Mvertices/s", "meshes_per_s": len(sample) / dt, "cores": cores, "kind": "port",
            "sample": f"{len(sample)} of the {len(meshes)} primitives (every {sample_every}th of the size-sorted list, {verts} vertices), one oracle encode() per primitive, {cores} threads pulling from one queue"}


def config4_block(args, shard, rank, world, local_rank, dist, torch, dxo, with_cpu=True):
    meshes, total_vertices, total_meshes = shard
    reps = max(1, min(args.steps, 3))
    times, tot, clocks, out = measure_config4(meshes, local_rank, reps, dist, torch, dxo)
    best = min(times)
    mean = sum(times) / len(times)
    block = {
        "workload": f"config4: {int(tot[3])} synthetic glTF primitives (vertex counts log-uniform in [1k, 100k], alternating grid patches / tori with uv seams), "
                    f"{int(tot[2])} vertices, through dxo_encode_batch; primitives assigned longest-processing-time-first to {world} rank(s) (draco_oxide_b200.sharding), one GPU each",
        "scaling": "strong", "n_gpus": world, "reps": reps,
        "e2e": {"value": tot[2] / mean / 1e6, "unit": "Mvertices/s", "best": tot[2] / best / 1e6, "meshes_per_s": tot[3] / mean, "ms_per_batch": 1e3 * mean,
                "h2d_bytes_per_batch": int(tot[0]), "d2h_bytes_per_batch": int(tot[1]), "times_s": times,
                "note": "wall clock of the call (max over ranks): host buffers in, Draco streams out; connectivity walks on the host workers, everything else in one segmented launch set per group of meshes"},
        "host_threads": host_threads(), "clocks": clocks,
    }
    if with_cpu and rank == 0:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import orc
        orc.build()
        block["cpu_baseline"] = config4_cpu_baseline(orc, meshes)
    return block


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def mark_begin(self):
        """Start of the timed region: nvidia-smi was launched earlier so that it is already sampling."""
        self.t_begin = time.perf_counter()

    def stop(self):
        t_end = time.perf_counter()
        t_begin = getattr(self, "t_begin", 0.0)
        inside = [r for t, r in self.rows if t_begin <= t <= t_end + 0.05]
        if not inside:  # region shorter than the sampling period: the samples that bracket it
            inside = [r for t, r in self.rows if t_begin - 0.15 <= t <= t_end + 0.15]
        self.rows = inside
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def pinned_copy(mesh):
    """Copies the mesh arrays into pinned host memory (torch) so H2D runs as DMA."""
    import numpy as np
    import torch
    import draco_oxide_b200 as dxo

    def pin(a):
        t = torch.empty(max(a.nbytes, 1), dtype=torch.uint8, pin_memory=True)
        v = t.numpy()[: a.nbytes].view(a.dtype).reshape(a.shape)
        v[...] = a
        return v, t
    keep = []
    faces, t = pin(mesh.faces); keep.append(t)
    atts = []
    for a in mesh.attributes:
        vals, t = pin(a.values); keep.append(t)
        pm = None
        if a.point_to_value is not None:
            pm, t = pin(a.point_to_value); keep.append(t)
        atts.append(dxo.Attribute(vals, a.att_type, a.domain, a.parents, pm, a.unique_id))
    m = dxo.Mesh(faces, atts)
    m._pinned = keep
    return m


def run_reference(args, rank, world):
    """CPU arm: the oracle restatement of the reference encoder on the host cores."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    orc.build()
    mesh, desc = make_mesh(args.workload)
    V = mesh.num_points()
    threads = max(1, min(64, host_threads()))  # every host thread (at most 64: ~250 MB each) encodes the workload mesh; the reference is single-threaded per mesh
    for _ in range(max(0, min(args.warmup, 1))):
        orc.encode_timed(mesh, 1, threads)
    t = 0.0
    for _ in range(args.steps):
        t += orc.encode_timed(mesh, 1, threads)
    value = V * threads * args.steps / t / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mvertices/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32->i32/u32 (bit-exact integer pipeline)", "data": "synthetic",
        "config": {"workload": desc},
        "cpu_baseline": {"value": value, "unit": "Mvertices/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} full encode() calls of the workload mesh per thread, {threads} thread(s); the reference encoder is single-threaded per mesh"},
        "e2e": {"value": value, "unit": "Mvertices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "C++ oracle (oracle/): the Rust reference cannot be compiled in this image; the oracle replaces its O(V^2) membership scans by rank lookups, so it is faster than the reference would be",
    }
    print(json.dumps(line), file=RESULT_OUT or sys.stdout, flush=True)


def run_config4(args, shard, rank, world, local_rank, dist, torch, dxo):
    """--workload config4: the sharded batch as the line itself (strong scaling). The batch entry has no resident mode —
    host connectivity of a group overlaps with the device work of the others — so value is the end-to-end figure."""
    block = config4_block(args, shard, rank, world, local_rank, dist, torch, dxo, with_cpu=not args.no_cpu_baseline)
    if rank != 0:
        return
    e = block["e2e"]
    line = {"metric": METRIC, "value": e["value"], "unit": "Mvertices/s", "n_gpus": world, "steps": block["reps"], "warmup": 2,
            "ms_per_step": e["ms_per_batch"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32->i32/u32 (bit-exact integer pipeline)", "data": "synthetic", "config": {"workload": block["workload"], "host_threads": host_threads()},
            "meshes_per_s": e["meshes_per_s"], "gpu_launches": None,
            "e2e": {"value": e["value"], "unit": "Mvertices/s", "h2d_bytes_per_step": e["h2d_bytes_per_batch"], "d2h_bytes_per_step": e["d2h_bytes_per_batch"],
                    "meshes_per_s": e["meshes_per_s"], "note": e["note"]},
            "value_scope": "the batch entry has no resident mode: value == e2e (host buffers in, streams out)",
            "roofline": {"bound": "hbm", "kernel": None, "achieved": None, "peak": hbm_peak()[0], "unit": "GB/s", "frac": None, "traffic": None,
                         "note": "per-kernel figures are taken on config 2 (same kernel bodies); the batch is bound by the host connectivity walks"},
            "clocks": block["clocks"], "cpu_baseline": block.get("cpu_baseline")}
    print(json.dumps(line), file=RESULT_OUT or sys.stdout, flush=True)


def run_config5(args, rank, world, local_rank, dist, torch, dxo):
    """--workload config5: qp in {8,10,11,12,14,16} on config 2's mesh (Edgebreaker; the reference has no sequential-connectivity
    attribute path). Per qp: a lone dxo_encode call from pinned host buffers, and the oracle on one thread."""
    from draco_oxide_b200 import synth
    mesh = synth.config2_mesh()
    V = mesh.num_points()
    pmesh = pinned_copy(mesh)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    orc.build()
    rows = []
    for qp in (8, 10, 11, 12, 14, 16):
        cfg = dxo.Config(position_bits=qp, device=local_rank)
        for _ in range(2):
            out = bytearray(); dxo.encode(pmesh, out, cfg)
        t0 = time.perf_counter()
        n = max(1, min(args.steps, 5))
        for _ in range(n):
            out = bytearray(); dxo.encode(pmesh, out, cfg)
        ms = 1e3 * (time.perf_counter() - t0) / n
        tm = dxo.last_timing()
        row = {"qp": qp, "stream_bytes": len(out), "e2e_single_call_ms": ms, "e2e_mvertices_per_s": V / ms / 1e3, "device_ms": tm["device_ms"],
               "host_connectivity_ms": tm["host_connectivity_ms"]}
        if not args.no_cpu_baseline and rank == 0:
            secs = orc.encode_timed(mesh, 1, 1, cfg)
            row["oracle_single_thread_mvertices_per_s"] = V / secs / 1e6
        rows.append(row)
    if rank != 0:
        return
    main_row = [r for r in rows if r["qp"] == 11][0]
    line = {"metric": METRIC, "value": main_row["e2e_mvertices_per_s"], "unit": "Mvertices/s", "n_gpus": world, "steps": max(1, min(args.steps, 5)), "warmup": 2,
            "ms_per_step": main_row["e2e_single_call_ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32->i32/u32 (bit-exact integer pipeline)", "data": "synthetic",
            "config": {"workload": "config5: qp sweep {8,10,11,12,14,16} on config 2's 1M-vertex mesh, Edgebreaker; 'sequential' is not applicable (unimplemented in the reference)"},
            "value_scope": "lone dxo_encode calls from pinned host buffers (latency figure); value is the qp = 11 row", "sweep": rows,
            "e2e": {"value": main_row["e2e_mvertices_per_s"], "unit": "Mvertices/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None}}
    print(json.dumps(line), file=RESULT_OUT or sys.stdout, flush=True)


RESULT_OUT = None  # the process's real stdout, see main()


def main():
    global RESULT_OUT
    args = parse_args()
    # The JSON line is the only thing that may reach stdout: libraries that write to fd 1 on their own (NCCL prints its
    # version banner there on the first collective) are sent to stderr for the rest of the run.
    sys.stdout.flush()
    RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    os.environ.setdefault("DXO_SIDE_INLINE", "1")  # read by the library when a session runs; DXO_SIDE_INLINE=0 gives the helper threads back
    # one process per GPU: each rank's batch entry gets its share of the host threads, not all of them
    os.environ.setdefault("DXO_BATCH_WORKERS", str(max(2, host_threads() // max(1, world))))
    # config 4's primitives are generated by forked numpy workers: before this process holds a CUDA context
    c4_shard = None
    if args.workload == "config4" or (args.workload == "config2" and not args.no_config4):
        c4_shard = Config4Shard(args.batch_meshes, rank, world)
    import numpy as np
    import torch
    import draco_oxide_b200 as dxo
    if not os.path.exists(dxo._capi.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    if not torch.cuda.is_available() or dxo.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the attribute path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    if args.workload == "config4":
        run_config4(args, c4_shard.generate(), rank, world, local_rank, dist, torch, dxo)
        if dist is not None:
            dist.destroy_process_group()
        return
    if args.workload == "config5":
        run_config5(args, rank, world, local_rank, dist, torch, dxo)
        if dist is not None:
            dist.destroy_process_group()
        return
    mesh, desc = make_mesh(args.workload)
    V = mesh.num_points()
    cfg = dxo.Config(device=local_rank)
    graph = args.graph_replay == "on" or (args.graph_replay == "auto" and default_sessions(world)[1])
    session_cfg = dxo.Config(device=local_rank, flags=dxo.Config.GRAPH_REPLAY if graph else 0)
    if args.workload == "config3" and args.sessions <= 0:
        args.sessions = 2       # ~0.9 GB resident per session
    if args.workload == "config3" and args.e2e_callers <= 0:
        args.e2e_callers = 4

    # ---- device-resident arm ------------------------------------------------------------
    # `value`: S sessions (one host thread each, own CUDA streams) encode the resident mesh concurrently, the way the
    # batch entry keeps a GPU busy: the serial rANS chains and the host-coded side streams of one mesh overlap with
    # the kernels of the others. A step = one pass of every session; every pass produces the complete Draco stream.
    S = args.sessions if args.sessions > 0 else default_sessions(world)[0]
    sessions, results, errors = [None] * S, [None] * S, []
    ready, go = threading.Barrier(S + 1), threading.Barrier(S + 1)

    def worker(i):
        try:
            torch.cuda.set_device(local_rank)
            sessions[i] = dxo.Session(mesh, session_cfg)
            if i == 0:
                results[i] = {"timing": dxo.last_timing(), "bytes": sessions[i].run()}
            sessions[i].run_steps(max(args.warmup, 3))
        except Exception as e:  # noqa: BLE001
            errors.append(e)
        ready.wait()
        go.wait()
        try:
            if not errors:
                ms, n = sessions[i].run_steps(args.steps)
                results[i] = dict(results[i] or {}, ms=ms, launches=n)
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    sampler = ClockSampler(local_rank)
    sampler.start()  # before the warm-up: nvidia-smi needs a moment to start, the timed region may be short
    threads = [threading.Thread(target=worker, args=(i,)) for i in range(S)]
    for th in threads:
        th.start()
    ready.wait()
    if errors:
        go.wait()
        raise errors[0]
    barrier()
    sampler.mark_begin()
    go.wait()
    for th in threads:
        th.join()
    barrier()
    if errors:
        raise errors[0]
    clocks_a = sampler.stop()
    input_bytes = results[0]["timing"]["h2d_bytes"]
    ref_bytes = results[0]["bytes"]
    ms_total = max(r["ms"] for r in results)        # the sessions start together; the slowest one ends the step loop
    launches = sum(r["launches"] for r in results)
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    n_launch = torch.tensor([launches], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(n_launch, op=dist.ReduceOp.SUM)
    ms_total_max = float(t.item())
    launches = int(n_launch.item())
    for extra in sessions[1:]:
        extra.close()
    sess = sessions[0]
    # one session alone (the latency of one mesh), for reference; this thread's streams and staging buffers are
    # created by the warm-up call
    sess.run_steps(2)
    ms_single, _ = sess.run_steps(args.steps)

    # per-kernel timing: separate profiled steps, CUDA events around every launch on its own stream, the three
    # attributes run one after the other (profiling mode 2) so that each kernel is timed alone
    dxo.set_profiling(2)
    agg = {}
    prof_steps = 3
    for _ in range(prof_steps):
        sess.run(want_bytes=False)
        for k in dxo.last_timing()["kernels"]:
            a = agg.setdefault(k["name"], {"ms": 0.0, "bytes": 0, "launches": 0})
            a["ms"] += k["ms"]; a["bytes"] += k["algorithmic_bytes"]; a["launches"] += 1
    dxo.set_profiling(False)
    peak, peak_src = hbm_peak()
    kernels = []
    for name, a in agg.items():
        ms = a["ms"] / a["launches"]
        by = a["bytes"] / a["launches"]
        kernels.append({"name": name, "launches_per_step": a["launches"] // prof_steps, "ms_per_launch": ms, "algorithmic_bytes_per_launch": by,
                        "gbs": by / ms / 1e6 if ms > 0 else None})
    tot_ms = sum(k["ms_per_launch"] * k["launches_per_step"] for k in kernels)
    for k in kernels:
        k["share_of_kernel_time"] = k["ms_per_launch"] * k["launches_per_step"] / tot_ms if tot_ms else None
    hbm_kernels = [k for k in kernels if not k["name"].startswith(("K10", "K9"))]
    # the longest single launch: small kernels' event intervals include launch latency in profiling mode, so summing
    # them over their launches would overstate them
    dom = max(hbm_kernels, key=lambda k: k["ms_per_launch"])
    attr_bytes = sum(k["algorithmic_bytes_per_launch"] * k["launches_per_step"] for k in hbm_kernels)
    attr_ms = sum(k["ms_per_launch"] * k["launches_per_step"] for k in hbm_kernels)
    rans = [k for k in kernels if k["name"].startswith("K10")]
    # The same profiled steps with the per-step forms of K4-K6 (what one-shot calls and the batch entry run: they chase the
    # operand indices through the tables in every step instead of reading a session's records): a second session under
    # DXO_NO_RINGS, read by the library when the session is uploaded.
    one_shot = None
    if args.workload == "config2":
        os.environ["DXO_NO_RINGS"] = "1"
        try:
            plain = dxo.Session(mesh, session_cfg)
        finally:
            del os.environ["DXO_NO_RINGS"]
        plain.run(want_bytes=False)
        assert plain.run() == ref_bytes
        dxo.set_profiling(2)
        agg2 = {}
        for _ in range(prof_steps):
            plain.run(want_bytes=False)
            for k in dxo.last_timing()["kernels"]:
                a = agg2.setdefault(k["name"], {"ms": 0.0, "bytes": 0, "launches": 0})
                a["ms"] += k["ms"]; a["bytes"] += k["algorithmic_bytes"]; a["launches"] += 1
        dxo.set_profiling(False)
        plain.close()
        hb = {n: a for n, a in agg2.items() if not n.startswith(("K10", "K9"))}
        ob, om = sum(a["bytes"] for a in hb.values()) / prof_steps, sum(a["ms"] for a in hb.values()) / prof_steps
        one_shot = {"ms_per_step": om, "gbs": ob / om / 1e6 if om else None, "frac_of_peak": ob / om / 1e6 / peak if om else None,
                    "kernels_ms_per_launch": {n: a["ms"] / a["launches"] for n, a in hb.items()},
                    "note": "per-step forms of K4-K6 (one-shot calls, batch entry); the line's kernels[] / roofline are a resident session's record-based forms"}
    n_sym = sum(len(a) * (2 if a.att_type == dxo.AttributeType.Normal else a.get_num_components()) for a in mesh.attributes)

    # ---- end-to-end arm: host buffers in pinned memory -> dxo_encode ---------------------
    # The reference-facing call is thread-safe and handle-free; its connectivity walks are serial per mesh, so the job
    # keeps the GPU fed by calling it from T host threads at once (what the batch entry does for a file set), every call
    # with its own H2D / D2H inside. A step = one call per caller thread. `single_call_ms` is the latency of a lone call.
    pmesh = pinned_copy(mesh)
    for _ in range(2):
        out = bytearray(); dxo.encode(pmesh, out, cfg)
    assert bytes(out) == ref_bytes
    t0 = time.perf_counter()
    lone_calls = 3
    host_ms = 0.0
    for _ in range(lone_calls):
        out = bytearray(); dxo.encode(pmesh, out, cfg)
        host_ms += dxo.last_timing()["host_connectivity_ms"]
    single_call_ms = 1e3 * (time.perf_counter() - t0) / lone_calls
    tm = dxo.last_timing()
    h2d, d2h = tm["h2d_bytes"], tm["d2h_bytes"]
    T = args.e2e_callers if args.e2e_callers > 0 else max(1, min(32, host_threads() // max(1, world)))  # ~250 MB of pinned tables per caller
    e2e_steps = max(1, min(args.steps, 10))
    e_ready, e_go, e_done = threading.Barrier(T + 1), threading.Barrier(T + 1), threading.Barrier(T + 1)
    e_errors = []
    e2e_reps = 3  # the arm is host-bound and shares the box's memory system with other tenants: three passes, the median is reported

    def caller():
        try:
            torch.cuda.set_device(local_rank)
            for _ in range(2):  # this thread's streams and staging buffers; the host block pools reach their working size
                o = bytearray(); dxo.encode(pmesh, o, cfg)
        except Exception as e:  # noqa: BLE001
            e_errors.append(e)
        for _ in range(e2e_reps):
            e_ready.wait()
            e_go.wait()
            try:
                for _ in range(e2e_steps):
                    o = bytearray(); dxo.encode(pmesh, o, cfg)
                if bytes(o) != ref_bytes:
                    e_errors.append(AssertionError("end-to-end stream differs from the resident session's"))
            except Exception as e:  # noqa: BLE001
                e_errors.append(e)
            e_done.wait()

    callers = [threading.Thread(target=caller) for _ in range(T)]
    for th in callers:
        th.start()
    sampler2 = ClockSampler(local_rank)
    e2e_times = []
    for rep in range(e2e_reps):
        e_ready.wait()
        if rep == 0:
            sampler2.start()
        barrier()
        if rep == 0:
            sampler2.mark_begin()
        t0 = time.perf_counter()
        e_go.wait()
        e_done.wait()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_times.append(float(t.item()))
    for th in callers:
        th.join()
    if e_errors:
        raise e_errors[0]
    e2e_s = sorted(e2e_times)[len(e2e_times) // 2]
    clocks_e2e = sampler2.stop()
    barrier()
    sess.close()
    # the sharded batch (config 4) on the same GPUs, before rank 0 spends its host threads on the CPU baseline
    c4 = None
    if args.workload == "config2" and not args.no_config4:
        c4 = config4_block(args, c4_shard.generate(), rank, world, local_rank, dist, torch, dxo, with_cpu=not args.no_cpu_baseline)
        c4_shard = None

    if rank == 0:
        value = V * world * S * args.steps / (ms_total_max * 1e-3) / 1e6
        line = {
            "metric": METRIC, "value": value, "unit": "Mvertices/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32->i32/u32 (bit-exact integer pipeline)", "data": "synthetic",
            "config": {"workload": desc, "units_per_step_per_gpu": f"{S} meshes ({S} concurrent resident sessions of the workload mesh)",
                       "host_threads": host_threads(), "host_waits": "blocking" if os.environ.get("DXO_BLOCKING_WAIT") else "spinning", "graph_replay": bool(graph), "side_streams": "inline, interleaved" if os.environ.get("DXO_SIDE_INLINE", "0") != "0" else "two helper threads per session",
                       "parallelism": f"{world} independent replicas, no collective",
                       "l2": (f"no flush needed: between two steps of a session the other {S - 1} sessions run theirs — {S} x "
                              f"{sum(k['algorithmic_bytes_per_launch'] * k['launches_per_step'] for k in kernels) / 1e6:.0f} MB of algorithmic kernel traffic per step "
                              f"({input_bytes / 1e6:.0f} MB of inputs per mesh) cycle through the 126 MB L2"),
                       "stream_bytes": len(ref_bytes)},
            "gpu_launches": int(launches),
            "single_session": {"value": V * args.steps / (ms_single * 1e-3) / 1e6, "unit": "Mvertices/s", "ms_per_step": ms_single / args.steps,
                               "note": "one mesh at a time on rank 0: bounded by the serial rANS chains and the host-coded side streams"},
            "clocks": clocks_a,
            "e2e": {"value": V * world * T * e2e_steps / e2e_s / 1e6, "unit": "Mvertices/s", "h2d_bytes_per_step": int(h2d) * T, "d2h_bytes_per_step": int(d2h) * T,
                    "ms_per_step": 1e3 * e2e_s / e2e_steps, "meshes_per_step_per_gpu": T, "caller_threads_per_gpu": T,
                    "single_call_ms": single_call_ms, "single_call_mvertices_per_s": V / single_call_ms / 1e3,
                    "host_connectivity_ms_per_call": host_ms / lone_calls, "steps": e2e_steps, "passes_s": e2e_times, "reported": "median pass", "clocks": clocks_e2e},
            "roofline": {"bound": "hbm", "kernel": dom["name"], "achieved": dom["gbs"], "peak": peak, "unit": "GB/s",
                         "frac": dom["gbs"] / peak if dom["gbs"] else None,
                         "traffic": ncu_traffic(args.workload).get(dom["name"]), "peak_source": peak_src,
                         "note": "dominant HBM-bound attribute kernel; K10 (rANS) is a serial latency-bound loop and is reported under 'rans' (SURVEY.md §8d)"},
            "attribute_kernels": {"algorithmic_bytes_per_step": attr_bytes, "ms_per_step": attr_ms, "gbs": attr_bytes / attr_ms / 1e6 if attr_ms else None,
                                  "frac_of_peak": attr_bytes / attr_ms / 1e6 / peak if attr_ms else None,
                                  "per_step_forms": one_shot},
            "rans": {"symbols_per_step": n_sym, "streams_in_flight": len(mesh.attributes),
                     "ms_longest_stream": max((k["ms_per_launch"] for k in rans), default=None),
                     "msymbols_per_s_per_stream": (max(len(a) * (2 if a.att_type == dxo.AttributeType.Normal else a.get_num_components()) for a in mesh.attributes)
                                                   / max((k["ms_per_launch"] for k in rans), default=1) / 1e3) if rans else None},
            "kernels": kernels,
        }
        if not args.no_cpu_baseline:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import orc
            orc.build()
            reps = 1 if args.workload == "config3" else 3
            secs = orc.encode_timed(mesh, reps, 1)
            cores = max(1, min(64, host_threads()))
            if args.workload == "config3":
                cores = min(cores, 16)  # ~1.3 GB per oracle thread
            secs_all = orc.encode_timed(mesh, 2, cores)
            line["cpu_baseline"] = {"value": V * 2 * cores / secs_all / 1e6, "unit": "Mvertices/s", "cores": cores, "kind": "port",
                                    "sample": f"2 full encode() calls of the same mesh per thread by the C++ oracle on {cores} threads at once (the reference "
                                              "encoder is single-threaded per mesh; compare with e2e.value); one thread alone: single_thread_value "
                                              "(compare with e2e.single_call_mvertices_per_s)",
                                    "single_thread_value": V * reps / secs / 1e6,
                                    "split": oracle_stage_split(orc, mesh)}
            line["cpu_baseline"]["split"]["repo_host_connectivity_ms"] = host_ms / lone_calls
            line["cpu_baseline"]["split"]["repo_device_ms"] = tm["device_ms"]
        line["value_scope"] = ("value = resident device step (K1-K10, result D2H, side streams, assembly); corner tables, Edgebreaker, sequencer and H2D are "
                               "outside it, so compare the reference arm with e2e, not with value")
        if c4 is not None:
            line["config4"] = c4
            line["meshes_per_s"] = c4["e2e"]["meshes_per_s"]
        print(json.dumps(line), file=RESULT_OUT or sys.stdout, flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
