#!/usr/bin/env python
"""Row-sharded multi-GPU check: `python -m torch.distributed.run --nproc-per-node N tools/mgpu_check.py [--assert]`.

The sharded K.V / mBCG / MLL / Lanczos (csrc/comm.cu: NCCL all-gather of the direction block + all-reduce of the packed dot
products) must reproduce the single-GPU result on the same inputs and probes, and both must match the CPU oracle.  With
--assert every comparison is an assertion and the process exits non-zero on failure (tests/test_gpu_multi.py runs it that way);
without it the script also prints timings at larger sizes."""
import os
import sys
import time
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from gpytorch_b200.distributed import Comm, init_from_env, shard_rows
from gpytorch_b200.engine import Plan
from oracle import kernels as ok, mll as om

ASSERT = "--assert" in sys.argv
rank, world, local = init_from_env()
dev = torch.device("cuda", local)
comm = Comm(rank, world) if world > 1 else None


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()


def check(cond, msg):
    if ASSERT:
        assert cond, msg
    elif not cond:
        print("MISMATCH:", msg, flush=True)


cases = [(4096, 5, "rbf", 0.8, 30), (6144, 20, "matern52", 2.0, 50)]
if not ASSERT:
    cases += [(int(os.environ.get("GP_N", 48000)), 10, "rbf", 1.0, 100), (16000, 20, "matern52", 2.0, 50)]
for (n, d, kind, ls, krank) in cases:
    n = (n // (world * 128)) * world * 128 if n > 8192 else (n // world) * world
    x, y = om.synthetic_problem(n, d, 0, torch.float32)
    eps1, eps2, rad = om.make_probe_noise(n, krank, 10, 1)
    rb, rc, _ = shard_rows(n, world, rank)
    xd = x.to(dev)
    p = Plan(xd, row_begin=rb, row_count=rc if world > 1 else 0, comm=comm).set_hypers(kind, ls, 1.0, 0.1)
    args = (y[rb:rb + rc].contiguous().to(dev), eps1.to(dev), eps2[rb:rb + rc].contiguous().to(dev), rad[rb:rb + rc].contiguous().to(dev))
    res, sol = p.mll(*args, 10, krank, 2000, want_solve=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.time()
    for _ in range(3):
        res, sol = p.mll(*args, 10, krank, 2000, want_solve=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = (time.time() - t0) / 3
    # the kernel seam on the shard: local rows of K @ v
    g = torch.Generator().manual_seed(9)
    v = torch.randn(n, 7, generator=g)
    kv = p.kmv(v.to(dev), add_noise=True)
    # every rank checks its own shard against rank-independent ground truth
    p1 = Plan(xd).set_hypers(kind, ls, 1.0, 0.1)
    full = (y.to(dev), eps1.to(dev), eps2.to(dev), rad.to(dev))
    r1, s1 = p1.mll(*full, 10, krank, 2000, want_solve=True)
    kv1 = p1.kmv(v.to(dev), add_noise=True)
    check(res.cg_iters == r1.cg_iters and res.precond_rank == r1.precond_rank, f"iters {res.cg_iters} vs {r1.cg_iters}")
    check(abs(res.inv_quad - r1.inv_quad) <= 2e-5 * abs(r1.inv_quad), f"inv_quad sharded {res.inv_quad} single {r1.inv_quad}")
    check(abs(res.logdet - r1.logdet) <= 2e-5 * abs(r1.logdet), f"logdet sharded {res.logdet} single {r1.logdet}")
    serr = rel(sol, s1[rb:rb + rc])
    check(serr < 5e-4, f"solve shard rel diff {serr}")
    check(rel(kv, kv1[rb:rb + rc]) < 2e-6, f"K.V shard vs single-GPU rows {rel(kv, kv1[rb:rb + rc])}")
    # Lanczos (LOVE root decomposition, exact_prediction_strategies.py:268-272) on the shard: same tridiagonal, same basis rows
    g2 = torch.Generator().manual_seed(5)
    init = torch.randn(n, generator=g2)
    ql, tl = p.lanczos(init[rb:rb + rc].contiguous().to(dev), 25)
    q1, t1 = p1.lanczos(init.to(dev), 25)
    check(tuple(tl.shape) == tuple(t1.shape), f"Lanczos size {tuple(tl.shape)} vs {tuple(t1.shape)}")
    if tuple(tl.shape) == tuple(t1.shape):
        check(rel(tl, t1) < 2e-4, f"Lanczos T sharded vs single {rel(tl, t1)}")
        check(rel(ql[:, :8], q1[rb:rb + rc, :8]) < 2e-3, f"Lanczos Q rows {rel(ql[:, :8], q1[rb:rb + rc, :8])}")
    if n <= 8192:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            K = ok.kernel_matrix(kind, x.double(), x.double(), ls, 1.0, True)
            o64 = om.mll_bbmm(kind, x.double(), y.double(), 0.0, ls, 1.0, 0.1, (eps1.double(), eps2.double(), rad.double()), precond_size=krank, K=K)
            o32 = om.mll_bbmm(kind, x, y, 0.0, ls, 1.0, 0.1, (eps1, eps2, rad), precond_size=krank)
        check(rel(kv, (K @ v.double() + 0.1 * v.double())[rb:rb + rc]) < 5e-6, "K.V shard vs oracle")
        check(res.cg_iters == o64.iters, f"iters {res.cg_iters} vs oracle {o64.iters}")
        check(abs(res.inv_quad - o64.inv_quad) <= max(1e-4 * abs(o64.inv_quad), 3 * abs(o32.inv_quad - o64.inv_quad)), f"inv_quad {res.inv_quad} oracle {o64.inv_quad}")
        check(abs(res.logdet - o64.logdet) <= max(1e-4 * abs(o64.logdet), 3 * abs(o32.logdet - o64.logdet)), f"logdet {res.logdet} oracle {o64.logdet}")
        check(rel(sol, o64.solves[rb:rb + rc, -1]) <= max(5e-4, 3 * rel(o32.solves[:, -1], o64.solves[:, -1])), "solve shard vs oracle")
    if rank == 0 and not ASSERT:
        torch.cuda.synchronize(); t0 = time.time()
        for _ in range(3):
            r1, s1 = p1.mll(*full, 10, krank, 2000, want_solve=True)
        torch.cuda.synchronize(); dt1 = (time.time() - t0) / 3
        print(f"world={world} n={n} {kind}: sharded iq={res.inv_quad:.4f} ld={res.logdet:.3f} it={res.cg_iters} k={res.precond_rank} {dt*1e3:.1f} ms | "
              f"single iq={r1.inv_quad:.4f} ld={r1.logdet:.3f} it={r1.cg_iters} {dt1*1e3:.1f} ms | solve rel diff {serr:.2e} | speedup {dt1/dt:.2f}x", flush=True)
    p1.close()
    p.close()
if world > 1:
    dist.barrier()
    comm.close()
    dist.destroy_process_group()
if rank == 0:
    print(f"mgpu_check ok (world={world}, assert={ASSERT})", flush=True)
