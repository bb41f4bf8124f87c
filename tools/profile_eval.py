#!/usr/bin/env python
"""One MLL evaluation at a BASELINE shape (+ a short Lanczos run) for ncu captures:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv python tools/profile_eval.py      (launch list)
    ncu --set full --clock-control none -k regex:'cg_|pc_|gram|chol_small|wsolve|slq|lz_|probes' -c 60 -o small python tools/profile_eval.py
The first evaluation warms up (allocation, attribute calls); `GP_PROFILE_WARM=0` profiles from the first launch on."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gpytorch_b200.engine import Plan
from oracle import mll as om
dev = torch.device("cuda:0")
kind, n, d, ls = (sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])) if len(sys.argv) > 4 else ("rbf", 50000, 10, 1.0)
x, y = om.synthetic_problem(n, d, 0, torch.float32)
pn = om.make_probe_noise(n, 100, 10, 1)
p = Plan(x.to(dev), backend="tcgen05").set_hypers(kind, ls, 1.0, 0.1)
args = (y.to(dev), pn[0].to(dev), pn[1].to(dev), pn[2].to(dev), 10, 100, 2000)
evals = int(os.environ.get("GP_EVALS", 2))
for _ in range(evals):
    res, _ = p.mll(*args)
torch.cuda.synchronize()
if os.environ.get("GP_LANCZOS", "1") == "1":
    q, t = p.lanczos(torch.randn(n, device=dev), 12)
    torch.cuda.synchronize()
print("ok", res.cg_iters, res.inv_quad, res.logdet, p.launches())
