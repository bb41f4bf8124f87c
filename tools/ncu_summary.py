"""Turn an .ncu-rep (ncu --set full) into the markdown summary kept under profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep "title" > profiles/xxx.md   (needs `ncu` on PATH; no GPU)"""
import csv, io, json, subprocess, sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
    "smsp__sass_inst_executed_op_tmem_ldt.sum", "smsp__sass_inst_executed_op_tmem_stt.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
]


def main():
    rep, title = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    names, units, vals = rows[0], rows[1], rows[2]
    m = {n: (u, v) for n, u, v in zip(names, units, vals)}
    print(f"# {title}\n")
    print("Numbers under ncu are cold-cache/serialised; the bench value comes from CUDA events (bench.py), not from here.\n")
    print("| metric | unit | value |\n|---|---|---|")
    print(f"| Kernel Name |  | {m.get('Kernel Name', ('', '?'))[1]} |")
    for w in WANT:
        if w in m:
            print(f"| {w} | {m[w][0]} | {m[w][1]} |")
    if len(sys.argv) > 3:
        rd = float(m["dram__bytes_read.sum"][1].replace(",", "")); wr = float(m["dram__bytes_write.sum"][1].replace(",", ""))
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot = rd * mult[m["dram__bytes_read.sum"][0]] + wr * mult[m["dram__bytes_write.sum"][0]]
        json.dump({"kernel": "gp::kmv_tc_kernel", "dram_bytes_per_launch": tot, "source": rep.split("/")[-1]}, open(sys.argv[3], "w"))


main()
