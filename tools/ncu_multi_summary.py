"""Per-kernel summary of an .ncu-rep holding many launches (ncu --set full): launches, device time, DRAM bytes and achieved GB/s,
L2 / SM throughput, achieved occupancy, registers.  usage: python tools/ncu_multi_summary.py rep.ncu-rep "title" > profiles/x.md"""
import csv, io, re, subprocess, sys
from collections import OrderedDict

UNITS = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}


def num(v):
    try:
        return float(v.replace(",", ""))
    except Exception:
        return float("nan")


def main():
    rep, title = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    names, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(names)}

    def get(r, key):
        i = col.get(key)
        if i is None:
            return float("nan")
        return num(r[i]) * UNITS.get(units[i], 1.0)

    groups = OrderedDict()
    for r in rows[2:]:
        kn = re.sub(r"\(.*", "", r[col["Kernel Name"]])
        kn = re.sub(r"^void ", "", kn)
        groups.setdefault(kn, []).append(r)
    print(f"# {title}\n")
    print("ncu --set full --clock-control none; per-launch times under ncu are cold-cache and serialised (compare shares, not absolutes).  "
          "DRAM GB/s = (dram__bytes_read + dram__bytes_write) / gpu__time_duration of the same launch; peak (MEASURED_PEAKS.json) 6564.5 GB/s.\n")
    print("| kernel | launches | avg us | DRAM MB / launch | DRAM GB/s | % of HBM peak | L2 throughput % | SM throughput % | achieved occupancy % | regs | grid x block |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    for kn, rs in groups.items():
        dur = [get(r, "gpu__time_duration.sum") for r in rs]
        by = [get(r, "dram__bytes_read.sum") + get(r, "dram__bytes_write.sum") for r in rs]
        gbs = [b / t / 1e9 if t > 0 else float("nan") for b, t in zip(by, dur)]
        l2 = [num(r[col["lts__throughput.avg.pct_of_peak_sustained_elapsed"]]) if "lts__throughput.avg.pct_of_peak_sustained_elapsed" in col else float("nan") for r in rs]
        smt = [num(r[col["sm__throughput.avg.pct_of_peak_sustained_elapsed"]]) for r in rs]
        occ = [num(r[col["sm__warps_active.avg.pct_of_peak_sustained_active"]]) for r in rs]
        regs = rs[0][col["launch__registers_per_thread"]]
        grid = rs[0][col["launch__grid_size"]] + " x " + rs[0][col["launch__block_size"]]
        n = len(rs)
        avg = lambda v: sum(v) / len(v)
        print(f"| {kn} | {n} | {avg(dur) * 1e6:.1f} | {avg(by) / 1e6:.2f} | {avg(gbs):.0f} | {avg(gbs) / 6564.5 * 100:.1f} | {avg(l2):.1f} | {avg(smt):.1f} | {avg(occ):.1f} | {regs} | {grid} |")


main()
