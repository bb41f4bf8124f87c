#!/usr/bin/env python
"""Workload for ncu: a few fused K.V products (and optionally one MLL eval) at the BASELINE C2 shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gpytorch_b200.engine import Plan
from oracle import mll as om

n = int(os.environ.get("GP_N", 50000)); d = int(os.environ.get("GP_D", 10)); t = 11
kind = os.environ.get("GP_KIND", "rbf"); backend = os.environ.get("GP_BACKEND", "tcgen05")
reps = int(os.environ.get("GP_REPS", 3)); do_mll = int(os.environ.get("GP_MLL", 0))
dev = torch.device("cuda:0")
x, y = om.synthetic_problem(n, d, 0, torch.float32)
p = Plan(x.to(dev), backend=backend).set_hypers(kind, float(os.environ.get("GP_LS", 1.0)), 1.0, 0.1)
v = torch.randn(n, t, device=dev)
for _ in range(reps):
    out = p.kmv(v, True)
torch.cuda.synchronize()
if do_mll:
    pn = om.make_probe_noise(n, 100, 10, 1)
    res, _ = p.mll(y.to(dev), pn[0].to(dev), pn[1].to(dev), pn[2].to(dev), 10, 100, 2000)
    print("mll", res.mll, res.cg_iters)
print("done", float(out.abs().sum()))
