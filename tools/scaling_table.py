"""profiles/scaling_rNN.md from the JSON lines of `bench.py --gpus N` (one file per N).
usage: python tools/scaling_table.py out.md n1.json n2.json n4.json n8.json"""
import json, sys

rows = []
for f in sys.argv[2:]:
    try:
        line = [l for l in open(f) if l.startswith("{")][-1]
        rows.append(json.loads(line))
    except Exception as e:  # noqa: BLE001
        print(f"skip {f}: {e}", file=sys.stderr)
rows.sort(key=lambda r: r["n_gpus"])
base = {r["n_gpus"]: r for r in rows}.get(1)
out = ["# Strong scaling of one MLL evaluation (bench.py --gpus N, one process per GPU, NCCL)\n",
       "Times are the max over ranks of CUDA-event timings between barriers (the bench contract); total work is fixed.\n",
       "| GPUs | C2 evals/s | C2 ms/eval | speed-up | e2e evals/s | C3 evals/s | C3 ms/eval | C3 speed-up | clocks MHz | throttle reasons |",
       "|---|---|---|---|---|---|---|---|---|---|"]
for r in rows:
    c3 = r.get("c3") or {}
    s2 = (base["ms_per_step"] / r["ms_per_step"]) if base else float("nan")
    s3 = (base["c3"]["ms_per_step"] / c3["ms_per_step"]) if base and base.get("c3") and c3 else float("nan")
    out.append(f"| {r['n_gpus']} | {r['value']:.2f} | {r['ms_per_step']:.2f} | {s2:.2f}x | {r['e2e']['value']:.2f} | "
               f"{c3.get('value', float('nan')):.3f} | {c3.get('ms_per_step', float('nan')):.1f} | {s3:.2f}x | "
               f"{r['clocks'].get('sm_mhz')} | {','.join(r['clocks'].get('reasons', [])) or '-'} |")
out.append("")
for r in rows:
    cfg = r["config"]
    out.append(f"* N={r['n_gpus']}: parallelism `{cfg.get('parallelism')}`, nsplit {cfg.get('nsplit')}, cg_iters {cfg.get('cg_iters')}, "
               f"mll {cfg.get('mll'):.6f}, gpu_launches {r.get('gpu_launches')}")
open(sys.argv[1], "w").write("\n".join(out) + "\n")
print("\n".join(out))
