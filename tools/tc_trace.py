#!/usr/bin/env python
"""Dump the pipeline event trace of CTA (0,0) of the tcgen05 K.V kernel (clock64 stamps per tile)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gpytorch_b200.engine import Plan
from oracle import mll as om
dev = torch.device("cuda:0")
n, d = 50000, 10
x, y = om.synthetic_problem(n, d, 0, torch.float32)
p = Plan(x.to(dev), backend="tcgen05").set_hypers("rbf", 1.0, 1.0, 0.1)
v = torch.randn(n, 11, device=dev)
p.kmv(v); torch.cuda.synchronize()
tr = torch.zeros(256 * 8, dtype=torch.int64, device=dev)
p.lib.gp_plan_set_trace(p._h, C.c_void_p(tr.data_ptr()))
p.kmv(v); torch.cuda.synchronize()
p.lib.gp_plan_set_trace(p._h, None)
t = tr.cpu().view(256, 8)
t0 = int(t[0, 0])
names = ["g1_issue", "g2_issue", "sfull_wait", "sfull_done", "ld_fold", "math_done", "tile_end"]
print("tile " + " ".join(f"{n_:>10s}" for n_ in names))
for u in range(0, 60):
    print(f"{u:4d} " + " ".join(f"{int(t[u, e]) - t0 if int(t[u, e]) else -1:10d}" for e in range(7)))
