#!/usr/bin/env python
"""A few launches of the fused K.V kernel at a BASELINE shape, for `ncu -k regex:kmv_tc -s 1 -c 1` captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gpytorch_b200.engine import Plan
from oracle import mll as om
dev = torch.device("cuda:0")
kind, n, d, ls = (sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])) if len(sys.argv) > 4 else ("rbf", 50000, 10, 1.0)
x, y = om.synthetic_problem(n, d, 0, torch.float32)
p = Plan(x.to(dev), backend="tcgen05").set_hypers(kind, ls, 1.0, 0.1)
v = torch.randn(n, 11, device=dev)
for _ in range(3):
    out = p.kmv(v)
torch.cuda.synchronize()
print("ok", p.info(), float(out.abs().sum()))
