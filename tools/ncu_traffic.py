"""Reads `ncu -i <rep> --page raw --csv` output (one or more files) and writes / updates profiles/ncu_traffic.json:
{workload: {kernel label: dram__bytes_read.sum + dram__bytes_write.sum per launch}} plus a markdown table on stdout.
usage: python tools/ncu_traffic.py <workload> <raw.csv> [...]   (labels: K1_minmax ... as bench.py names them)"""
import csv, json, os, re, sys

LABELS = [("minmax", "K1_minmax"), ("oct_quantize", "K3_oct_quantize"), ("quantize", "K2_quantize"), ("seq_prepare", "seq_prepare"),
          ("predict_parallelogram", "K4_predict_parallelogram"), ("predict_normal", "K5_predict_normal"), ("predict_texcoord", "K6_predict_texcoord"),
          ("predict_delta", "K7_predict_delta"), ("histogram", "K8_histogram"), ("build_table", "K9_build_table")]


def label(kernel):
    for key, lab in LABELS:
        if key in kernel:
            return lab
    return re.sub(r"\(.*", "", kernel).replace("dxo::gpu::", "")


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


def main():
    workload, files = sys.argv[1], sys.argv[2:]
    rows = {}
    for f in files:
        lines = [l for l in open(f) if not l.startswith("==")]
        rd = csv.reader(lines)
        header = next(rd)
        units = next(rd)
        idx = {h: i for i, h in enumerate(header)}
        for r in rd:
            if len(r) < len(header):
                continue
            k = label(r[idx["Kernel Name"]])
            def get(name, scale_to=None):
                if name not in idx:
                    return None
                v = num(r[idx[name]])
                if v is None:
                    return None
                u = units[idx[name]]
                mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}.get(u, 1)
                return v * mult
            rec = rows.setdefault(k, {"n": 0, "read": 0.0, "write": 0.0, "us": 0.0, "l2hit": 0.0, "dram_pct": 0.0, "regs": 0, "warps_active": 0.0})
            rec["n"] += 1
            rec["read"] += get("dram__bytes_read.sum") or 0
            rec["write"] += get("dram__bytes_write.sum") or 0
            rec["us"] += get("gpu__time_duration.sum") or 0
            rec["l2hit"] += get("lts__t_sector_hit_rate.pct") or 0
            rec["dram_pct"] += get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed") or 0
            rec["regs"] = int(get("launch__registers_per_thread") or 0)
            rec["warps_active"] += get("sm__warps_active.avg.pct_of_peak_sustained_active") or 0
    out_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
    data = json.load(open(out_path)) if os.path.exists(out_path) else {}
    data.setdefault(workload, {})
    print("| kernel | launches | time us | DRAM read MB | DRAM write MB | DRAM % peak | L2 hit % | warps active % | regs |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|---:|")
    for k, r in rows.items():
        n = r["n"]
        data[workload][k] = (r["read"] + r["write"]) / n
        print(f"| {k} | {n} | {r['us'] / n:.1f} | {r['read'] / n / 1e6:.1f} | {r['write'] / n / 1e6:.1f} | {r['dram_pct'] / n:.1f} | {r['l2hit'] / n:.1f} | {r['warps_active'] / n:.1f} | {r['regs']} |")
    json.dump(data, open(out_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
