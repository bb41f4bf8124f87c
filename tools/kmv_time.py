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
    if os.environ.get("GP_GRAD"):
        l = torch.randn(n, 11, device=dev); r = torch.randn(n, 11, device=dev)
        p.bilinear_grad(l, r); torch.cuda.synchronize(); t0 = time.time()
        for _ in range(3): g = p.bilinear_grad(l, r)
        torch.cuda.synchronize()
        print(f"   bilinear_grad (s=11): {(time.time() - t0) / 3 * 1e3:.2f} ms  -> {g}", flush=True)
    p.close()

# full training step through the public API (forward MLL + backward: hyper-parameter gradients via gp_bilinear_grad)
if os.environ.get("GP_TRAIN"):
    import gpytorch_b200 as gp
    n, d = 50000, 10
    x, y = om.synthetic_problem(n, d, 0, torch.float32)
    x, y = x.to(dev), y.to(dev)

    class M(gp.models.ExactGP):
        def __init__(self, tx, ty, lik):
            super().__init__(tx, ty, lik)
            self.mean_module = gp.means.ConstantMean()
            self.covar_module = gp.kernels.ScaleKernel(gp.kernels.RBFKernel())

        def forward(self, xx):
            return gp.distributions.MultivariateNormal(self.mean_module(xx), self.covar_module(xx))

    lik = gp.likelihoods.GaussianLikelihood().to(dev)
    model = M(x, y, lik).to(dev)
    mll = gp.mlls.ExactMarginalLogLikelihood(lik, model)
    model.train(); lik.train()
    params = [p_ for p_ in model.parameters() if p_.requires_grad]
    with gp.settings.max_preconditioner_size(100), gp.settings.probe_seed(1):
        def step():
            for p_ in params: p_.grad = None
            loss = -mll(model(x), y)
            loss.backward()
            return loss
        for _ in range(2): step()
        torch.cuda.synchronize(); t0 = time.time()
        for _ in range(5): loss = step()
        torch.cuda.synchronize(); dt = (time.time() - t0) / 5
        t0 = time.time()
        with torch.no_grad():
            for _ in range(5): l2 = -mll(model(x), y)
        torch.cuda.synchronize(); df = (time.time() - t0) / 5
        if os.environ.get("GP_TORCH_PROF"):
            from torch.profiler import profile, ProfilerActivity
            with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
                step(); torch.cuda.synchronize()
            print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
    print(f"train step (fwd+bwd, API, N={n}): {dt * 1e3:.2f} ms; forward only {df * 1e3:.2f} ms; loss={float(loss):.5f}", flush=True)
