#!/usr/bin/env python
"""Time the fused K.V kernel alone (CUDA events inside libgpbbmm, back-to-back launches) and one MLL evaluation at C2/C3.
GPBBMM_LIB=<path> selects another build of the library (A/B comparisons inside one gpurun call: boxes differ by ~3 %)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gpytorch_b200.engine import Plan
from oracle import mll as om
dev = torch.device("cuda:0")
cases = [("rbf", 50000, 10)] + ([("matern52", 50000, 20)] if os.environ.get("GP_C3") else [])
for kind, n, d in cases:
    x, y = om.synthetic_problem(n, d, 0, torch.float32)
    p = Plan(x.to(dev), backend="tcgen05").set_hypers(kind, 1.0, 1.0, 0.1)
    v = torch.randn(n, 11, device=dev)
    ts = [p.time_kmv_kernel(v, 5, 100) for _ in range(3)]
    pn = om.make_probe_noise(n, 100, 10, 1)
    yd = y.to(dev); a, b, c = (q.to(dev) for q in pn)
    res, _ = p.mll(yd, a, b, c, 10, 100, 2000)
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(5): res, _ = p.mll(yd, a, b, c, 10, 100, 2000)
    torch.cuda.synchronize(); dt = (time.time() - t0) / 5
    print(f"{os.environ.get('GPBBMM_LIB', 'default')}: {kind} N={n} d={d}: K.V kernel {min(ts):.4f} ms (runs {['%.4f' % t for t in ts]}), "
          f"MLL eval {dt * 1e3:.2f} ms, iq={res.inv_quad:.3f} ld={res.logdet:.2f} it={res.cg_iters}", flush=True)
    p.close()
