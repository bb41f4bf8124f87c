"""Per-kernel HBM traffic table from
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none --csv --log-file x.csv <cmd>
usage: python tools/dram_list.py x.csv "title" [hbm_peak_GBps] > profiles/xxx.md
Times under ncu are cold-cache and serialised; the GB/s column is bytes / that time, so it is a LOWER bound of what the kernel
reaches inside a warm evaluation."""
import csv, re, sys
from collections import OrderedDict

UNIT = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6,
        "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
peak = float(sys.argv[3]) if len(sys.argv) > 3 else 6564.5
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
per = OrderedDict()                      # (launch id) -> {metric: value}
names = {}
for r in csv.DictReader(lines):
    i = r["ID"]
    names[i] = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
    v = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
    per.setdefault(i, {})[r["Metric Name"]] = v
g = OrderedDict()
for i, m in per.items():
    c = g.setdefault(names[i], [0, 0.0, 0.0, 0.0, 0.0])
    c[0] += 1
    c[1] += m.get("gpu__time_duration.sum", 0.0)
    c[2] += m.get("dram__bytes_read.sum", 0.0)
    c[3] += m.get("dram__bytes_write.sum", 0.0)
    c[4] += m.get("lts__t_bytes.sum", 0.0)
print(f"# {sys.argv[2]}\n")
print(f"HBM peak used for the last column: {peak:.1f} GB/s (MEASURED_PEAKS.json, burst copy).  Per-launch averages.\n")
print("| kernel | launches | avg us | DRAM read MB | DRAM write MB | L2 traffic MB | DRAM GB/s | % of HBM peak | L2 GB/s |\n|---|---|---|---|---|---|---|---|---|")
for k, (n, t, rd, wr, l2) in sorted(g.items(), key=lambda kv: -kv[1][1]):
    t_us = t / n
    gbs = (rd + wr) / n / (t_us * 1e-6) / 1e9 if t_us > 0 else 0.0
    l2g = l2 / n / (t_us * 1e-6) / 1e9 if t_us > 0 else 0.0
    print(f"| {k} | {n} | {t_us:.1f} | {rd / n / 1e6:.3f} | {wr / n / 1e6:.3f} | {l2 / n / 1e6:.2f} | {gbs:.0f} | {100 * gbs / peak:.1f} | {l2g:.0f} |")
