#!/usr/bin/env python
"""Pipeline event trace of CTA (0,0) of the second-generation fused K.V kernel (kmv_tc2.cu): clock64 stamps per column tile.
Columns: G1(0,u) issued | P(0,u) seen by the issuer (GEMM2(0,u) issue) | P(1,u) seen | issuer done with tile u |
         warpgroup 0 starts tile u | warpgroup 1 starts tile u | warpgroup 0 released P(u) | warpgroup 1 released P(u)"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gpytorch_b200.engine import Plan
from oracle import mll as om
dev = torch.device("cuda:0")
kind, n, d, ls = (sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])) if len(sys.argv) > 4 else ("rbf", 50000, 10, 1.0)
x, y = om.synthetic_problem(n, d, 0, torch.float32)
p = Plan(x.to(dev), backend="tcgen05").set_hypers(kind, ls, 1.0, 0.1)
v = torch.randn(n, 11, device=dev)
p.kmv(v); torch.cuda.synchronize()
tr = torch.zeros(256 * 16, dtype=torch.int64, device=dev)
p.lib.gp_plan_set_trace(p._h, C.c_void_p(tr.data_ptr()))
p.kmv(v); torch.cuda.synchronize()
p.lib.gp_plan_set_trace(p._h, None)
t = tr.cpu().view(256, 16)
t0 = int(t[4, 2])
order = [2, 4, 1, 9, 10, 7, 8, 0, 3, 5, 6]
names = ["wg0_start", "wg0_rel", "iss0_seesP", "iss0_G2iss", "iss0_bfull", "iss0_done", "O(0,u)done", "S(0,u)done", "wg1_start", "wg1_rel", "iss1_seesP"]
print(f"# {kind} N={n} d={d} npoly={os.environ.get('GP_NPOLY', 'default')} info={p.info()}")
print("tile " + " ".join(f"{n_:>10s}" for n_ in names))
for u in range(0, 48):
    print(f"{u:4d} " + " ".join(f"{int(t[u, e]) - t0 if int(t[u, e]) else -1:10d}" for e in order))
