#!/usr/bin/env python
"""Bring-up diagnostics on a B200 (run through gpurun).  Prints, never asserts: each section is isolated."""
import math, os, subprocess, sys, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gpytorch_b200.engine import Plan
from oracle import kernels as ok, linalg as ol, mll as om

dev = torch.device("cuda:0")
SECTIONS = sys.argv[1:] or ["probe", "simt", "tc", "rows", "pivchol", "mbcg", "mll", "lanczos", "grad", "time"]

def section(name):
    def deco(fn):
        if name not in SECTIONS: return fn
        print(f"\n===== {name} =====", flush=True)
        try:
            fn()
        except Exception:
            traceback.print_exc()
        sys.stdout.flush()
        return fn
    return deco

def rel(a, b):
    a = a.double().cpu(); b = b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item(), (a - b).abs().max().item()

@section("probe")
def _():
    exe = os.path.join(os.path.dirname(__file__), "..", "gpytorch_b200", "lib", "umma_probe")
    for v in ("0", "1"):
        r = subprocess.run(["timeout", "60", exe, v], capture_output=True, text=True)
        print(r.stdout.strip(), r.stderr.strip()[-300:], "rc=", r.returncode)

def kmv_case(backend, kind, n1, n2, d, t, ls, same, seed=0, ard=False):
    g = torch.Generator().manual_seed(seed)
    x1 = torch.rand(n1, d, generator=g, dtype=torch.float64)
    x2 = x1 if same else torch.rand(n2, d, generator=g, dtype=torch.float64)
    v = torch.randn(x2.size(0), t, generator=g, dtype=torch.float64)
    lsv = torch.linspace(0.7, 1.3, d, dtype=torch.float64) * ls if ard else ls
    K = ok.kernel_matrix(kind, x1, x2, lsv, 1.7, same)
    ref = K @ v + (0.3 * v if same else 0)
    p = Plan(x1.float().to(dev), None if same else x2.float().to(dev), backend=backend)
    p.set_hypers(kind, lsv.tolist() if ard else ls, 1.7, 0.3)
    out = p.kmv(v.float().to(dev), add_noise=same)
    torch.cuda.synchronize()
    r, m = rel(out, ref)
    print(f"{backend:8s} {kind:9s} n1={n1} n2={x2.size(0)} d={d} t={t} same={same} ard={ard} info={p.info()} rel={r:.2e} maxabs={m:.2e}", flush=True)
    p.close()
    return r

@section("simt")
def _():
    for kind in ("rbf", "matern12", "matern32", "matern52"):
        kmv_case("simt", kind, 1000, 1000, 3, 11, 0.5, True)
    kmv_case("simt", "rbf", 777, 1300, 10, 5, 1.2, False)
    kmv_case("simt", "rbf", 1000, 1000, 10, 20, 1.2, True, ard=True)
    kmv_case("simt", "matern52", 2500, 2500, 20, 11, 2.0, True)

@section("tc")
def _():
    for kind in ("rbf", "matern12", "matern32", "matern52"):
        kmv_case("tcgen05", kind, 1000, 1000, 3, 11, 0.5, True)
    kmv_case("tcgen05", "rbf", 777, 1300, 10, 5, 1.2, False)
    kmv_case("tcgen05", "rbf", 1000, 1000, 10, 20, 1.2, True, ard=True)
    kmv_case("tcgen05", "matern52", 2500, 2500, 20, 11, 2.0, True)
    kmv_case("tcgen05", "rbf", 4096, 4096, 10, 11, 1.0, True)
    kmv_case("tcgen05", "rbf", 128, 96, 10, 16, 1.0, False)

@section("rows")
def _():
    g = torch.Generator().manual_seed(3)
    x = torch.rand(900, 5, generator=g, dtype=torch.float64)
    p = Plan(x.float().to(dev), backend="simt").set_hypers("matern32", 0.8, 2.0, 0.1)
    idx = torch.tensor([0, 5, 899, 17])
    K = ok.kernel_matrix("matern32", x, x, 0.8, 2.0, True)
    print("rows", rel(p.rows(idx), K[idx]), "diag", rel(p.diag(), K.diagonal()))

@section("pivchol")
def _():
    for (n, d, kind, ls, rank) in [(1000, 3, "rbf", 0.5, 15), (3000, 10, "rbf", 1.0, 100), (2000, 4, "matern52", 0.7, 50)]:
        x, y = om.synthetic_problem(n, d, 0, torch.float64)
        K = ok.kernel_matrix(kind, x, x, ls, 1.0, True)
        Lo, pivo = ol.pivoted_cholesky(torch.ones(n, dtype=torch.float64), lambda i: K[i], rank, 1e-3)
        p = Plan(x.float().to(dev), backend="simt").set_hypers(kind, ls, 1.0, 0.1)
        lt, piv, st = p.pivoted_cholesky(rank, 1e-3)
        same_piv = piv.cpu().tolist() == pivo.tolist()
        m = min(lt.size(0), Lo.size(1))
        print(f"n={n} d={d} {kind} rank oracle={Lo.size(1)} gpu={lt.size(0)} pivots_equal={same_piv} first={piv[:6].cpu().tolist()} / {pivo[:6].tolist()}",
              "L rel/max", rel(lt[:m].t(), Lo[:, :m]))
        pre = ol.build_preconditioner(Lo, 0.1, pivo)
        w, ld, st = p.precond_build(lt)
        v = torch.randn(n, 4, dtype=torch.float64)
        pv = (v - (w.double().cpu() @ (w.double().cpu().t() @ v))) / 0.1
        print("   logdetP gpu/oracle", ld, pre.logdet, " apply rel", rel(pv, pre.apply(v)))
        eps1, eps2, rad = om.make_probe_noise(n, lt.size(0), 10, 1)
        z = p.precond_probes(lt, eps1.to(dev), eps2.to(dev))
        print("   probes rel", rel(z, pre.probes(eps1.double()[: Lo.size(1)], eps2.double())))
        p.close()

@section("mbcg")
def _():
    for backend in ("simt", "tcgen05"):
        for (n, d, precond) in [(1000, 3, False), (3000, 10, True)]:
            x, y = om.synthetic_problem(n, d, 0, torch.float64)
            ls = 0.5 if d == 3 else 1.0
            K = ok.kernel_matrix("rbf", x, x, ls, 1.0, True)
            A = K + 0.1 * torch.eye(n, dtype=torch.float64)
            g = torch.Generator().manual_seed(5)
            rhs = torch.randn(n, 11, generator=g, dtype=torch.float64)
            p = Plan(x.float().to(dev), backend=backend).set_hypers("rbf", ls, 1.0, 0.1)
            W = None; pre = None
            if precond:
                lt, piv, _ = p.pivoted_cholesky(50, 1e-3)
                W, ld, _ = p.precond_build(lt)
                pre = ol.build_preconditioner(lt.double().cpu().t().contiguous(), 0.1)
            for tol, mi in ((1.0, 1000), (1e-4, 1000)):
                so, to, io = ol.linear_cg(lambda v: A @ v, rhs, n_tridiag=10, tolerance=tol, max_iter=mi,
                                          preconditioner=(pre.apply if pre else None), return_info=True)
                sg, tg, ig = p.mbcg(rhs.float().to(dev), 10, tol, mi, 20, W)
                exact = torch.linalg.solve(A, rhs)
                print(f"{backend} n={n} precond={precond} tol={tol}: iters gpu/oracle {ig.iters}/{io.iters} J {ig.tridiag_size}/{to.size(-1)}",
                      "solve vs oracle", rel(sg, so), "vs exact", rel(sg, exact)[0], "oracle vs exact", rel(so, exact)[0],
                      "tmat", rel(tg, to) if tg.shape == to.shape else (tg.shape, to.shape),
                      "slq gpu/oracle", p.slq_logdet(tg, n), ol.slq_logdet(to, n))
            p.close()

@section("mll")
def _():
    for backend in ("simt", "tcgen05"):
        for (n, d, kind, ls, rank, minp) in [(1000, 3, "rbf", 0.5, 15, 2000), (3000, 10, "rbf", 1.0, 100, 2000), (2500, 6, "matern52", 1.0, 30, 2000)]:
            x, y = om.synthetic_problem(n, d, 0, torch.float32)
            pn = om.make_probe_noise(n, rank, 10, 1)
            xo, yo = x.double(), y.double()
            ch = om.mll_cholesky(kind, xo, yo, 0.0, ls, 1.0, 0.1)
            ro = om.mll_bbmm(kind, xo, yo, 0.0, ls, 1.0, 0.1, tuple(a.double() for a in pn), precond_size=rank, min_precond_size=minp)
            p = Plan(x.to(dev), backend=backend).set_hypers(kind, ls, 1.0, 0.1)
            res, _ = p.mll(y.to(dev), pn[0].to(dev), pn[1].to(dev), pn[2].to(dev), 10, rank, minp)
            print(f"{backend} n={n} {kind}: gpu iq={res.inv_quad:.5f} ld={res.logdet:.4f} mll={res.mll:.6f} it={res.cg_iters} J={res.tridiag_size} k={res.precond_rank} |"
                  f" oracle iq={ro.inv_quad:.5f} ld={ro.logdet:.4f} mll={ro.mll:.6f} it={ro.iters} k={0 if ro.precond is None else ro.precond.L.size(1)} |"
                  f" chol iq={ch.inv_quad:.5f} ld={ch.logdet:.4f} mll={ch.mll:.6f}", flush=True)
            p.close()

@section("lanczos")
def _():
    n, d = 1500, 4
    x, y = om.synthetic_problem(n, d, 0, torch.float64)
    A = ok.kernel_matrix("rbf", x, x, 0.6, 1.0, True) + 0.1 * torch.eye(n, dtype=torch.float64)
    init = torch.randn(n, 1, dtype=torch.float64)
    Qo, To = ol.lanczos_tridiag(lambda v: A @ v, 30, init)
    p = Plan(x.float().to(dev), backend="simt").set_hypers("rbf", 0.6, 1.0, 0.1)
    Q, T = p.lanczos(init[:, 0].float().to(dev), 30)
    Qd = Q.double().cpu()
    print("J", T.shape, To.shape, "T rel", rel(T, To[0]) if T.shape == To[0].shape else None,
          "orth", (Qd.t() @ Qd - torch.eye(Qd.size(1), dtype=torch.float64)).abs().max().item(),
          "QtAQ-T", (Qd.t() @ A @ Qd - T.double().cpu()).abs().max().item())

@section("grad")
def _():
    n, d, s = 800, 5, 7
    g = torch.Generator().manual_seed(2)
    x = torch.rand(n, d, generator=g, dtype=torch.float64)
    Lf = torch.randn(n, s, generator=g, dtype=torch.float64); Rt = torch.randn(n, s, generator=g, dtype=torch.float64)
    for kind in ("rbf", "matern12", "matern32", "matern52"):
        for ard in (False, True):
            ls = (torch.linspace(0.6, 1.1, d, dtype=torch.float64) if ard else torch.tensor(0.8, dtype=torch.float64)).requires_grad_(True)
            os_ = torch.tensor(1.3, dtype=torch.float64, requires_grad=True)
            K = ok.kernel_matrix(kind, x, x, ls, os_, True)
            (Lf * (K @ Rt)).sum().backward()
            p = Plan(x.float().to(dev), backend="simt").set_hypers(kind, ls.detach().tolist() if ard else float(ls), 1.3, 0.1)
            gl, go = p.bilinear_grad(Lf.float().to(dev), Rt.float().to(dev))
            print(kind, "ard" if ard else "iso", "dls gpu", [round(v, 4) for v in gl][:3], "ref", [round(v, 4) for v in ls.grad.reshape(-1).tolist()][:3],
                  "dos gpu/ref", round(go, 4), round(os_.grad.item(), 4))
            p.close()

@section("time")
def _():
    n, d, t = 50000, 10, 11
    x, y = om.synthetic_problem(n, d, 0, torch.float32)
    v = torch.randn(n, t)
    outs = {}
    for backend in ("simt", "tcgen05"):
        try:
            p = Plan(x.to(dev), backend=backend).set_hypers("rbf", 1.0, 1.0, 0.1)
            vd = v.to(dev)
            for _ in range(3): out = p.kmv(vd, True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            for _ in range(10): out = p.kmv(vd, True)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            outs[backend] = out
            print(f"{backend}: K.V N={n} d={d} t={t}: {ms:.3f} ms  -> {2*n*n*(d+t)/ms/1e9:.1f} TF/s algorithmic, {n*n/ms/1e6:.1f} Gpair/s, info={p.info()}", flush=True)
            pn = om.make_probe_noise(n, 100, 10, 1)
            yd = y.to(dev); a, b, c = (q.to(dev) for q in pn)
            res, _ = p.mll(yd, a, b, c, 10, 100, 2000)
            torch.cuda.synchronize(); t0 = time.time()
            for _ in range(3): res, _ = p.mll(yd, a, b, c, 10, 100, 2000)
            torch.cuda.synchronize(); dt = (time.time() - t0) / 3
            print(f"{backend}: MLL eval {dt*1e3:.1f} ms ({1/dt:.2f} evals/s) iq={res.inv_quad:.3f} ld={res.logdet:.2f} it={res.cg_iters} k={res.precond_rank}", flush=True)
            p.close()
        except Exception:
            traceback.print_exc()
    if len(outs) == 2:
        print("tc vs simt", rel(outs["tcgen05"], outs["simt"]))
