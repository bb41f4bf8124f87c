"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into the per-kernel table kept under profiles/.
usage: python tools/launch_summary.py launches.csv "title" "note" > profiles/xxx.md"""
import collections, csv, re, sys

rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
rd = csv.DictReader(rows)
tot = collections.OrderedDict()
for r in rd:
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])[:72]
    v = float(r["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r["Metric Unit"], 1e-3)
    c = tot.setdefault(name, [0, 0.0])
    c[0] += 1; c[1] += v
total = sum(v for _, v in tot.values()); n = sum(c for c, _ in tot.values())
print(f"# {sys.argv[2]}\n\n{sys.argv[3]}\n\ntotal device time {total / 1e3:.2f} ms over {n} launches\n")
print("| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|")
for k, (c, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"| {k} | {c} | {v:.1f} | {100 * v / total:.1f}% | {v / c:.2f} |")
