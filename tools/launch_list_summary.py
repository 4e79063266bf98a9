"""Per-kernel summary of an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X.csv`):
python tools/launch_list_summary.py X.csv  ->  markdown table (launches, avg / max us, share)."""
import collections, csv, io, re, sys
rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
agg = collections.OrderedDict()
total = 0.0
for r in csv.DictReader(io.StringIO("".join(rows))):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^void ", "", name)[:72]
    us = float(r["Metric Value"].replace(",", "")) / 1e3
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1; a[1] += us; a[2] = max(a[2], us)
    total += us
n = sum(a[0] for a in agg.values())
print(f"total {total:.1f} us over {n} launches\n")
print("| kernel | launches | avg us | max us | share |\n|---|---:|---:|---:|---:|")
for k, (c, s, m) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {k} | {c} | {s / c:.1f} | {m:.1f} | {100 * s / total:.1f}% |")
