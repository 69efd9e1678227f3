"""Markdown summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, summed time and share per kernel.
usage: python tools/launch_list_summary.py launches.csv "title" "command" > profiles/rNN_bench_launches.md"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
ix = {k: i for i, k in enumerate(rows[h])}
agg, n = {}, 0
for r in rows[h + 1:]:
    if len(r) < len(ix) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
    unit, val = r[ix["Metric Unit"]], float(r[ix["Metric Value"]].replace(",", ""))
    ms = val * {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0}.get(unit, 1e-6)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ms
    n += 1
tot = sum(v[1] for v in agg.values())
print(f"# {sys.argv[2]}\n\ncommand: `{sys.argv[3]}`\n(per-launch times are serialised and cold-cache: only the SHARES are comparable with the CUDA-event numbers of `bench.py`)\n")
print(f"{n} launches, {tot:.2f} ms summed\n\n| kernel | launches | summed ms | avg us | share |\n|---|---|---|---|---|")
for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {c} | {ms:.3f} | {ms * 1e3 / c:.2f} | {100 * ms / tot:.1f} % |")
