#!/usr/bin/env python
"""Bring-up / A-B of the second-generation fused K.V kernel (kmv_tc2.cu) on a B200 (run through gpurun).  Prints, never asserts.
  parity : K.V vs the fp64 oracle for every kind, polynomial share (GP_NPOLY = 0, 2, 4), TS (KP <= 48) and SS (KP = 64) modes,
           ragged / cross / sharded-like shapes
  time   : C2 and C3-like kernel time for the round-1 kernel (GP_KMV_V1=1) and the new one at each polynomial share
"""
import os
import sys
import time
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gpytorch_b200.engine import Plan
from oracle import kernels as ok, mll as om

dev = torch.device("cuda:0")
SECTIONS = sys.argv[1:] or ["parity", "time", "mll"]


def rel(a, b):
    a = a.double().cpu(); b = b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()


def setenv(v1, npoly):
    if v1:
        os.environ["GP_KMV_V1"] = "1"
    else:
        os.environ.pop("GP_KMV_V1", None)
    os.environ["GP_NPOLY"] = str(npoly)


def case(kind, n1, n2, d, t, ls, same, v1=False, npoly=2, seed=0):
    setenv(v1, npoly)
    g = torch.Generator().manual_seed(seed)
    x1 = torch.rand(n1, d, generator=g, dtype=torch.float64)
    x2 = x1 if same else torch.rand(n2, d, generator=g, dtype=torch.float64)
    v = torch.randn(x2.size(0), t, generator=g, dtype=torch.float64)
    K = ok.kernel_matrix(kind, x1, x2, ls, 1.7, same)
    ref = K @ v + (0.3 * v if same else 0)
    try:
        p = Plan(x1.float().to(dev), None if same else x2.float().to(dev), backend="tcgen05")
        p.set_hypers(kind, ls, 1.7, 0.3)
        out = p.kmv(v.float().to(dev), add_noise=same)
        torch.cuda.synchronize()
        r = rel(out, ref)
        print(f"{'v1' if v1 else 'v2'} npoly={npoly} {kind:9s} n1={n1} n2={x2.size(0)} d={d} t={t} same={same} {p.info()} rel={r:.2e}"
              f"{'   <<<<<< BAD' if not (r < 5e-6) else ''}", flush=True)
        p.close()
    except Exception:
        traceback.print_exc()


if "parity" in SECTIONS:
    print("===== parity =====", flush=True)
    case("rbf", 300, 300, 3, 11, 0.5, True, npoly=0)          # one CTA, T = 5 tiles
    case("rbf", 64, 64, 3, 1, 0.5, True, npoly=0)            # T = 1
    case("rbf", 129, 200, 3, 4, 0.5, False, npoly=0)         # cross, 2 column tiles... T = 4
    for npoly in (0, 2, 4):
        for kind in ("rbf", "matern12", "matern32", "matern52"):
            case(kind, 1500, 1500, 7, 11, 0.9, True, npoly=npoly)
    case("rbf", 4096, 4096, 10, 16, 1.0, True, npoly=2)
    case("rbf", 777, 1300, 10, 5, 1.2, False, npoly=2)
    case("matern52", 300, 40, 20, 33, 1.3, False, npoly=2)   # KP = 64 -> SS mode, t > 16
    case("matern52", 5000, 5000, 20, 11, 2.0, True, npoly=2)  # SS mode multi tile
    case("matern32", 3000, 3000, 14, 11, 1.5, True, npoly=4)  # KP = 48: widest TS
    case("rbf", 20000, 20000, 10, 11, 1.0, True, npoly=2)     # several splits, ring wrap
    case("rbf", 20000, 20000, 10, 11, 1.0, True, v1=True)

if "time" in SECTIONS:
    print("===== time =====", flush=True)
    for kind, n, d, ls in (("rbf", 50000, 10, 1.0), ("matern52", 50000, 20, 2.0)):
        x, y = om.synthetic_problem(n, d, 0, torch.float32)
        xd = x.to(dev)
        v = torch.randn(n, 11, device=dev)
        ref = None
        for v1, npoly in ((True, 0), (False, 0), (False, 2), (False, 4)):
            setenv(v1, npoly)
            try:
                p = Plan(xd, backend="tcgen05").set_hypers(kind, ls, 1.0, 0.1)
                out = p.kmv(v)
                ts = [p.time_kmv_kernel(v, 5, 50) for _ in range(3)]
                if ref is None:
                    ref = out
                print(f"{'v1' if v1 else 'v2'} npoly={npoly} {kind} N={n} d={d} {p.info()}: K.V kernel {min(ts):.4f} ms {['%.4f' % t for t in ts]}"
                      f" rel vs v1 {rel(out, ref):.2e}", flush=True)
                p.close()
            except Exception:
                traceback.print_exc()

if "mll" in SECTIONS:
    print("===== mll =====", flush=True)
    n, d = 50000, 10
    x, y = om.synthetic_problem(n, d, 0, torch.float32)
    pn = om.make_probe_noise(n, 100, 10, 1)
    yd = y.to(dev); a, b, c = (q.to(dev) for q in pn)
    for v1, npoly in ((True, 0), (False, 2)):
        setenv(v1, npoly)
        p = Plan(x.to(dev), backend="tcgen05").set_hypers("rbf", 1.0, 1.0, 0.1)
        res, _ = p.mll(yd, a, b, c, 10, 100, 2000)
        torch.cuda.synchronize(); t0 = time.time()
        for _ in range(5):
            res, _ = p.mll(yd, a, b, c, 10, 100, 2000)
        torch.cuda.synchronize(); dt = (time.time() - t0) / 5
        print(f"{'v1' if v1 else 'v2'} npoly={npoly}: MLL eval {dt * 1e3:.2f} ms, iq={res.inv_quad:.3f} ld={res.logdet:.3f} it={res.cg_iters} "
              f"(oracle fp32: iq=186933.8 ld=-112077.41)", flush=True)
        p.close()
