#!/usr/bin/env python
"""Row-sharded multi-GPU check (torchrun --nproc-per-node N tools/mgpu_check.py): the sharded MLL / mBCG must
reproduce the single-GPU result on the same inputs and probes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from gpytorch_b200.distributed import Comm, init_from_env, shard_rows
from gpytorch_b200.engine import Plan
from oracle import mll as om

rank, world, local = init_from_env()
dev = torch.device("cuda", local)
comm = Comm(rank, world) if world > 1 else None
for (n, d, kind, ls, krank) in [(4096, 5, "rbf", 0.8, 30), (int(os.environ.get("GP_N", 48000)), 10, "rbf", 1.0, 100), (16000, 20, "matern52", 2.0, 50)]:
    n = (n // (world * 128)) * world * 128 if n > 8192 else (n // world) * world
    x, y = om.synthetic_problem(n, d, 0, torch.float32)
    eps1, eps2, rad = om.make_probe_noise(n, krank, 10, 1)
    rb, rc, _ = shard_rows(n, world, rank)
    xd = x.to(dev)
    p = Plan(xd, row_begin=rb, row_count=rc if world > 1 else 0, comm=comm).set_hypers(kind, ls, 1.0, 0.1)
    args = (y[rb:rb+rc].contiguous().to(dev), eps1.to(dev), eps2[rb:rb+rc].contiguous().to(dev), rad[rb:rb+rc].contiguous().to(dev))
    res, sol = p.mll(*args, 10, krank, 2000, want_solve=True)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t0 = time.time()
    for _ in range(3): res, sol = p.mll(*args, 10, krank, 2000, want_solve=True)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    dt = (time.time() - t0) / 3
    if rank == 0:
        p1 = Plan(xd).set_hypers(kind, ls, 1.0, 0.1)
        r1, s1 = p1.mll(y.to(dev), eps1.to(dev), eps2.to(dev), rad.to(dev), 10, krank, 2000, want_solve=True)
        torch.cuda.synchronize(); t0 = time.time()
        for _ in range(3): r1, s1 = p1.mll(y.to(dev), eps1.to(dev), eps2.to(dev), rad.to(dev), 10, krank, 2000, want_solve=True)
        torch.cuda.synchronize(); dt1 = (time.time() - t0) / 3
        serr = ((sol - s1[rb:rb+rc]).norm() / s1[rb:rb+rc].norm()).item()
        print(f"world={world} n={n} {kind}: sharded iq={res.inv_quad:.4f} ld={res.logdet:.3f} it={res.cg_iters} k={res.precond_rank} {dt*1e3:.1f} ms | "
              f"single iq={r1.inv_quad:.4f} ld={r1.logdet:.3f} it={r1.cg_iters} {dt1*1e3:.1f} ms | solve rel diff {serr:.2e} | speedup {dt1/dt:.2f}x", flush=True)
        p1.close()
    p.close()
if world > 1:
    dist.barrier(); comm.close(); dist.destroy_process_group()
