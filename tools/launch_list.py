"""Turn `ncu --metrics gpu__time_duration.sum --csv --log-file x.csv` into the per-kernel launch list kept under profiles/.
usage: python tools/launch_list.py x.csv "title" > profiles/launches_xxx.md"""
import csv, re, sys
from collections import OrderedDict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(u, 1.0)
        rows.append((re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", ""), v))
g = OrderedDict()
for k, v in rows:
    g.setdefault(k, []).append(v)
tot = sum(v for _, v in rows)
print(f"# {sys.argv[2]}\n")
print(f"total device time {tot / 1e3:.2f} ms over {len(rows)} launches (per-launch times under ncu are cold-cache and serialised: compare SHARES)\n")
print("| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|")
for k, v in sorted(g.items(), key=lambda kv: -sum(kv[1])):
    print(f"| {k} | {len(v)} | {sum(v):.1f} | {100 * sum(v) / tot:.1f}% | {sum(v) / len(v):.2f} |")
