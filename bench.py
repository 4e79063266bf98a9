#!/usr/bin/env python
"""bench.py — Mvertices/s of the attribute-encoding hot path on BASELINE.json's config 2.

  python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload config2]

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
    ap.add_argument("--workload", default="config2", choices=["config1", "config2", "config2_small"])
    ap.add_argument("--sessions", type=int, default=0,
                    help="concurrent resident sessions per GPU for `value`; 0 = one per three host threads available to this GPU, "
                         "at most 6 (each session keeps a main thread and two side-stream coders busy)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-callers", type=int, default=0,
                    help="host threads calling dxo_encode() concurrently in the end-to-end arm (0 = the host threads this GPU can count on)")
    return ap.parse_args()


def default_sessions(world):
    """(sessions per GPU, blocking waits?) from the host threads this GPU can count on. A session keeps two side-stream
    coders busy; its main thread spins in the waits (fastest) when there are three threads per session to spare and
    sleeps in them (DXO_BLOCKING_WAIT) when host threads are scarce."""
    per_gpu = max(1, (os.cpu_count() or 2) // max(1, world))
    if per_gpu >= 15:
        return max(1, min(8, per_gpu // 2)), False  # measured on 16 threads: 5 sessions 1290, 8 sessions 1420, 16 sessions 1460 Mvertices/s
    # few host threads (e.g. 32 threads for 8 GPUs): sleeping waits, and more sessions than threads still help because a
    # session's host work comes in bursts (measured with 4 threads: 2 / 4 / 6 sessions give 617 / 751 / 811 Mvertices/s)
    return max(2, min(8, per_gpu * 3 // 2)), True


def make_mesh(workload):
    from draco_oxide_b200 import synth
    if workload == "config1":
        return synth.config1_mesh(), "config1: 101x251 grid, 25 351 vertices, 50 000 triangles, pos+normal+uv, 11/8/10 bits"
    if workload == "config2_small":
        return synth.config2_mesh(300), "config2_small: 300x300 grid (debug)"
    return synth.config2_mesh(), "config2: 1000x1000 grid, 1 000 000 vertices, 1 996 002 triangles, pos+normal+uv, Edgebreaker, qp/qn/qt=11/8/10"


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return PEAKS_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


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
    threads = max(1, min(64, os.cpu_count() or 1))  # every host thread (at most 64: ~250 MB each) encodes the workload mesh; the reference is single-threaded per mesh
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

    if args.sessions <= 0 and default_sessions(world)[1]:
        os.environ.setdefault("DXO_BLOCKING_WAIT", "1")  # read when a thread's device context is created
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

    mesh, desc = make_mesh(args.workload)
    V = mesh.num_points()
    cfg = dxo.Config(device=local_rank)

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
            sessions[i] = dxo.Session(mesh, cfg)
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
    T = args.e2e_callers if args.e2e_callers > 0 else max(1, min(32, (os.cpu_count() or 1) // max(1, world)))  # ~250 MB of pinned tables per caller
    e2e_steps = max(1, min(args.steps, 10))
    e_ready, e_go = threading.Barrier(T + 1), threading.Barrier(T + 1)
    e_errors = []

    def caller():
        try:
            torch.cuda.set_device(local_rank)
            for _ in range(2):  # this thread's streams and staging buffers; the host block pools reach their working size
                o = bytearray(); dxo.encode(pmesh, o, cfg)
        except Exception as e:  # noqa: BLE001
            e_errors.append(e)
        e_ready.wait()
        e_go.wait()
        try:
            for _ in range(e2e_steps):
                o = bytearray(); dxo.encode(pmesh, o, cfg)
            if bytes(o) != ref_bytes:
                e_errors.append(AssertionError("end-to-end stream differs from the resident session's"))
        except Exception as e:  # noqa: BLE001
            e_errors.append(e)

    callers = [threading.Thread(target=caller) for _ in range(T)]
    for th in callers:
        th.start()
    e_ready.wait()
    sampler2 = ClockSampler(local_rank)
    sampler2.start()
    barrier()
    sampler2.mark_begin()
    t0 = time.perf_counter()
    e_go.wait()
    for th in callers:
        th.join()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if e_errors:
        raise e_errors[0]
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    clocks_e2e = sampler2.stop()
    barrier()

    if rank == 0:
        value = V * world * S * args.steps / (ms_total_max * 1e-3) / 1e6
        line = {
            "metric": METRIC, "value": value, "unit": "Mvertices/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32->i32/u32 (bit-exact integer pipeline)", "data": "synthetic",
            "config": {"workload": desc, "units_per_step_per_gpu": f"{S} meshes ({S} concurrent resident sessions of the workload mesh)",
                       "host_threads": os.cpu_count(), "host_waits": "blocking" if os.environ.get("DXO_BLOCKING_WAIT") else "spinning",
                       "parallelism": f"{world} independent replicas, no collective",
                       "l2": f"inputs larger than L2: {input_bytes / 1e6:.0f} MB resident per mesh vs 126 MB L2",
                       "stream_bytes": len(ref_bytes)},
            "gpu_launches": int(launches),
            "single_session": {"value": V * args.steps / (ms_single * 1e-3) / 1e6, "unit": "Mvertices/s", "ms_per_step": ms_single / args.steps,
                               "note": "one mesh at a time on rank 0: bounded by the serial rANS chains and the host-coded side streams"},
            "clocks": clocks_a,
            "e2e": {"value": V * world * T * e2e_steps / e2e_s / 1e6, "unit": "Mvertices/s", "h2d_bytes_per_step": int(h2d) * T, "d2h_bytes_per_step": int(d2h) * T,
                    "ms_per_step": 1e3 * e2e_s / e2e_steps, "meshes_per_step_per_gpu": T, "caller_threads_per_gpu": T,
                    "single_call_ms": single_call_ms, "single_call_mvertices_per_s": V / single_call_ms / 1e3,
                    "host_connectivity_ms_per_call": host_ms / lone_calls, "steps": e2e_steps, "clocks": clocks_e2e},
            "roofline": {"bound": "hbm", "kernel": dom["name"], "achieved": dom["gbs"], "peak": peak, "unit": "GB/s",
                         "frac": dom["gbs"] / peak if dom["gbs"] else None,
                         "traffic": NCU_TRAFFIC_CONFIG2.get(dom["name"]) if args.workload == "config2" else None, "peak_source": peak_src,
                         "note": "dominant HBM-bound attribute kernel; K10 (rANS) is a serial latency-bound loop and is reported under 'rans' (SURVEY.md §8d)"},
            "attribute_kernels": {"algorithmic_bytes_per_step": attr_bytes, "ms_per_step": attr_ms, "gbs": attr_bytes / attr_ms / 1e6 if attr_ms else None,
                                  "frac_of_peak": attr_bytes / attr_ms / 1e6 / peak if attr_ms else None},
            "rans": {"symbols_per_step": n_sym, "streams_in_flight": len(mesh.attributes),
                     "ms_longest_stream": max((k["ms_per_launch"] for k in rans), default=None),
                     "msymbols_per_s_per_stream": (max(len(a) * (2 if a.att_type == dxo.AttributeType.Normal else a.get_num_components()) for a in mesh.attributes)
                                                   / max((k["ms_per_launch"] for k in rans), default=1) / 1e3) if rans else None},
            "kernels": kernels,
        }
        if not args.no_cpu_baseline and world == 1:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import orc
            reps = 3
            secs = orc.encode_timed(mesh, reps, 1)
            cores = max(1, min(64, os.cpu_count() or 1))
            secs_all = orc.encode_timed(mesh, 2, cores)
            line["cpu_baseline"] = {"value": V * 2 * cores / secs_all / 1e6, "unit": "Mvertices/s", "cores": cores, "kind": "port",
                                    "sample": f"2 full encode() calls of the same mesh per thread by the C++ oracle on {cores} threads at once (the reference "
                                              "encoder is single-threaded per mesh; compare with e2e.value); one thread alone: single_thread_value "
                                              "(compare with e2e.single_call_mvertices_per_s)",
                                    "single_thread_value": V * reps / secs / 1e6}
        print(json.dumps(line), file=RESULT_OUT or sys.stdout, flush=True)
    sess.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
