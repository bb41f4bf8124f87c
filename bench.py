#!/usr/bin/env python
"""bench.py -- ExactGP MLL evaluations/s (and fused K.V GF/s) at the BASELINE C2 workload.

A "step" = one evaluation of the exact-GP marginal log likelihood through the BBMM path: pivoted-Cholesky
preconditioner (rank 100) -> N(0,P) probes -> mBCG (t = 10 probes + y, J = 21 iterations) -> SLQ log-det ->
log_prob, on synthetic data: X ~ U[0,1]^{N x d}, y = sin(3 sum x) + 0.1 eps, RBF.

    python bench.py --gpus 1 --steps 10 --warmup 3            # our engine (libgpbbmm, sm_100a)
    python bench.py --impl reference --steps 2 --warmup 1     # the reference algorithm on the host CPU cores
    torchrun ... bench.py --gpus N ...                        # rows of K sharded over N GPUs (strong scaling)

Rank 0 prints ONE JSON line (see the task contract): value = whole-job MLL evals/s with inputs resident in HBM,
e2e = same metric through the public gpytorch-style API with HOST inputs (H2D of X, y and D2H of the result
inside the timed region), roofline = the fused K.V kernel alone (CUDA events on its own stream), cpu_baseline =
the oracle port on the host cores over a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
if not os.environ.get("GP_KEEP_NCCL_DEBUG"):
    os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the single JSON line (NCCL prints its version banner there)

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on
    "c2": dict(name="ExactGP RBF N=50000 d=10, mBCG t=11 probes + rank-100 pivoted-Cholesky precond", n=50000, d=10,
               kind="rbf", lengthscale=1.0, outputscale=1.0, noise=0.1, probes=10, rank=100),
    # configs[2]: the 8-GPU strong-scaling case
    "c3": dict(name="ExactGP Matern-5/2 N=200000 d=20, row-sharded K.V + NCCL CG dots", n=200000, d=20,
               kind="matern52", lengthscale=2.0, outputscale=1.0, noise=0.1, probes=10, rank=100),
    # small case for quick checks
    "c1": dict(name="ExactGP RBF N=4000 d=3", n=4000, d=3, kind="rbf", lengthscale=0.5, outputscale=1.0, noise=0.1,
               probes=10, rank=15),
}
METRIC = "exactgp_mll_evals_per_sec"
UNIT = "evals/s"


def host_cores() -> int:
    """Usable host cores: scheduler affinity capped by the cgroup CPU quota (os.cpu_count() over-reports inside a
    container and 128 torch threads on a few real cores run ~80x slower than 8)."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(int(txt[0]) / int(txt[1]))))
            else:
                q = int(txt[0])
                per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                if q > 0:
                    n = min(n, max(1, q // per))
        except Exception:
            pass
    env = os.environ.get("GP_CPU_THREADS")
    return int(env) if env else n


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port (torch on the host cores) on a bounded sample
# ------------------------------------------------------------------------------------------------------------
def cpu_reference_eval(w, n_sample, seed=0):
    """One full MLL evaluation of the reference algorithm (dense K once + mBCG with dense K @ V) at n_sample rows."""
    import torch
    from oracle import mll as om

    x, y = om.synthetic_problem(n_sample, w["d"], seed, torch.float32)
    pn = om.make_probe_noise(n_sample, w["rank"], w["probes"], 1)
    t0 = time.perf_counter()
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r = om.mll_bbmm(w["kind"], x, y, 0.0, w["lengthscale"], w["outputscale"], w["noise"], pn, precond_size=w["rank"])
    return time.perf_counter() - t0, r


def tune_threads(w):
    """Pick the torch thread count that runs the reference path fastest on this host (<= usable cores)."""
    import torch

    cores = host_cores()
    cands = sorted({c for c in (cores, 64, 32, 16, 8) if c <= cores}, reverse=True)
    best, best_t = cores, None
    for c in cands:
        torch.set_num_threads(c)
        cpu_reference_eval(w, 1500)
        t, _ = cpu_reference_eval(w, 3000)
        if best_t is None or t < best_t:
            best, best_t = c, t
    torch.set_num_threads(best)
    return best


def cpu_baseline(w, budget_rows=12500):
    import torch

    cores = tune_threads(w)
    n = w["n"]
    ns = min(n, budget_rows)
    scale = (n / ns) ** 2  # the work is O(n^2): pairs scale quadratically
    cpu_reference_eval(w, min(ns, 2000))  # warm the thread pool / allocator
    dt, r = cpu_reference_eval(w, ns)
    return {
        "value": 1.0 / (dt * scale), "unit": UNIT, "cores": cores, "kind": "port",
        "sample": f"oracle port (torch CPU, dense K once + 21 dense K@V) on the first {ns} rows: {dt:.2f} s per eval, "
                  f"scaled by (N/{ns})^2 = {scale:.2f} to N={n}",
        "sample_seconds": dt, "sample_rows": ns, "cg_iters": r.iters,
    }


def run_reference(args, w):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import torch

    cores = tune_threads(w)
    ns = min(w["n"], args.ref_rows)
    scale = (w["n"] / ns) ** 2
    for _ in range(max(args.warmup, 1)):
        cpu_reference_eval(w, min(ns, 4000))
    times = []
    for _ in range(args.steps):
        dt, r = cpu_reference_eval(w, ns)
        times.append(dt)
    ms = 1e3 * sum(times) / len(times) * scale
    val = 1e3 / ms
    sample = (f"oracle port of the reference path (gpytorch kernels + linear_operator mBCG restated, torch CPU, {cores} threads) "
              f"on {ns} rows per step, time scaled by (N/{ns})^2 = {scale:.2f}")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["name"], "parallelism": "host cpu", "timing": "perf_counter, bounded sample"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------
def run_ours(args, w):
    import torch
    import torch.distributed as dist

    import gpytorch_b200 as gp
    from gpytorch_b200 import settings
    from gpytorch_b200.distributed import Comm, init_from_env, shard_rows
    from gpytorch_b200.engine import Plan
    from oracle import mll as om  # synthetic data + probe base samples only (inputs, not compute)

    rank, world, local = init_from_env()
    dev = torch.device("cuda", local)
    n, d = w["n"], w["d"]
    if n % world:
        raise SystemExit(f"N={n} must be divisible by the number of GPUs ({world})")
    comm = Comm(rank, world) if world > 1 else None
    rb, rc, _ = shard_rows(n, world, rank)

    x, y = om.synthetic_problem(n, d, 0, torch.float32)
    eps1, eps2, rad = om.make_probe_noise(n, w["rank"], w["probes"], 1)
    xd, yd = x.to(dev), y.to(dev)
    e1d, e2d, radd = eps1.to(dev), eps2[rb : rb + rc].contiguous().to(dev), rad[rb : rb + rc].contiguous().to(dev)
    y_loc = yd[rb : rb + rc].contiguous()

    plan = Plan(xd, backend=args.backend, row_begin=rb, row_count=rc if world > 1 else 0, comm=comm)
    plan.set_hypers(w["kind"], w["lengthscale"], w["outputscale"], w["noise"])
    info = plan.info()

    l2_flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step():
        l2_flush.zero_()  # evict the L2 between steps (timing rule); ~60 us of the ~30 ms step
        res, _ = plan.mll(y_loc, e1d, e2d, radd, w["probes"], w["rank"], 2000, 1e-3, 1.0, 1000, 20, warn=False)
        return res

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        res = step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = plan.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        res = step()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = plan.launches() - l0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = 1e3 / ms_step

    # ---- the fused K.V kernel alone: roofline ----
    v = torch.randn(n, w["probes"] + 1, device=dev)
    kms = plan.time_kmv_kernel(v, warmup=3, reps=20)
    tt = w["probes"] + 1
    flops = 2.0 * rc * n * (d + tt)  # algorithmic flops of this rank's row block
    peaks, peak_src = measured_peaks()
    ach = flops / (kms * 1e-3) / 1e12
    peak = float(peaks["bf16_tflops"])
    trans = (2 if w["kind"] != "rbf" else 1) * rc * n  # MUFU ops per launch (ex2, + sqrt for Matern)
    mufu_peak = 16.0 * info["n_sm"] * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6  # 16 MUFU/clk/SM (unmeasured doc figure)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "kmv_tc_dram_bytes.json")
    if os.path.exists(tpath) and world == 1 and args.workload == "c2" and info["backend"] == "tcgen05":
        with open(tpath) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    roofline = {
        "bound": "tensor", "kernel": "gp::kmv_tc_kernel" if info["backend"] == "tcgen05" else "gp::kmv_simt_kernel",
        "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
        "peak_source": f"bf16 dense burst, {peak_src}; the kernel runs kind::tf32 (nominal half of bf16) with a 3xTF32 split",
        "ms_per_launch": kms, "algorithmic_flops_per_launch": flops,
        "gpairs_per_s": rc * n / (kms * 1e-3) / 1e9,
        "mufu_bound": {"transcendentals_per_launch": trans, "achieved_per_s": trans / (kms * 1e-3),
                       "peak_per_s_at_max_clock": mufu_peak, "frac": trans / (kms * 1e-3) / mufu_peak},
        "algorithmic_bytes_per_launch": 4.0 * (n * d + 2 * n * tt),
    }

    # ---- e2e: the public API with HOST inputs (pinned), H2D + D2H inside the timed region ----
    e2e = None
    if world == 1:
        xh, yh = x.pin_memory(), y.pin_memory()
        xdev = torch.empty_like(xd); ydev = torch.empty_like(yd)
        lik = gp.likelihoods.GaussianLikelihood().to(dev)
        lik.noise = w["noise"]
        base = gp.kernels.RBFKernel() if w["kind"] == "rbf" else gp.kernels.MaternKernel(nu={"matern12": 0.5, "matern32": 1.5, "matern52": 2.5}[w["kind"]])
        base.lengthscale = w["lengthscale"]
        cov = gp.kernels.ScaleKernel(base).to(dev)
        cov.outputscale = w["outputscale"]
        mean = gp.means.ZeroMean()

        class Model(gp.models.ExactGP):
            def __init__(self):
                super().__init__(xdev, ydev, lik)
                self.mean_module, self.covar_module = mean, cov

            def forward(self, xx):
                return gp.distributions.MultivariateNormal(self.mean_module(xx), self.covar_module(xx))

        model = Model().to(dev)
        mll = gp.mlls.ExactMarginalLogLikelihood(lik, model)
        model.train(); lik.train()

        def e2e_step():
            l2_flush.zero_()
            xdev.copy_(xh, non_blocking=True)
            ydev.copy_(yh, non_blocking=True)
            with torch.no_grad(), settings.max_preconditioner_size(w["rank"]), settings.num_trace_samples(w["probes"]), \
                    settings.backend(args.backend), settings.probe_seed(1):
                out = mll(model(xdev), ydev)
            return float(out.item())  # D2H read of the result

        for _ in range(max(args.warmup, 1)):
            e2e_step()
        torch.cuda.synchronize(dev)
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.steps):
            last = e2e_step()
        f1.record()
        torch.cuda.synchronize(dev)
        ems = f0.elapsed_time(f1) / args.steps
        e2e = {"value": 1e3 / ems, "unit": UNIT, "h2d_bytes_per_step": int(x.numel() * 4 + y.numel() * 4),
               "d2h_bytes_per_step": 4, "ms_per_step": ems, "api": "gpytorch_b200.mlls.ExactMarginalLogLikelihood(model(x), y)",
               "mll": last}
    else:
        # multi-GPU: same call through the engine API with host inputs on every rank
        xh, yh = x.pin_memory(), y[rb : rb + rc].contiguous().pin_memory()
        xdev = torch.empty_like(xd); ydev = torch.empty_like(y_loc)
        plan2 = Plan(xdev, backend=args.backend, row_begin=rb, row_count=rc, comm=comm)

        def e2e_step():
            l2_flush.zero_()
            xdev.copy_(xh, non_blocking=True); ydev.copy_(yh, non_blocking=True)
            plan2.set_hypers(w["kind"], w["lengthscale"], w["outputscale"], w["noise"])
            r, _ = plan2.mll(ydev, e1d, e2d, radd, w["probes"], w["rank"], 2000, 1e-3, 1.0, 1000, 20, warn=False)
            return r.mll

        for _ in range(max(args.warmup, 1)):
            e2e_step()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.steps):
            last = e2e_step()
        f1.record()
        barrier()
        t2 = torch.tensor([f0.elapsed_time(f1)], device=dev, dtype=torch.float64)
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        ems = float(t2.item()) / args.steps
        e2e = {"value": 1e3 / ems, "unit": UNIT, "h2d_bytes_per_step": int(x.numel() * 4 + rc * 4) * world, "d2h_bytes_per_step": 8 * world,
               "ms_per_step": ems, "api": "gpytorch_b200.Plan.mll (row-sharded)", "mll": last}

    if rank == 0:
        cpu = cpu_baseline(w, args.ref_rows) if not args.no_cpu else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {
                "workload": w["name"], "parallelism": f"row-shard x{world}" if world > 1 else "single GPU",
                "backend": info["backend"], "nsplit": info["nsplit"], "kpad": info["kpad"],
                "l2_policy": "L2 flushed between timed steps by a 192 MiB memset inside the timed region",
                "cg_iters": res.cg_iters, "precond_rank": res.precond_rank, "tridiag_size": res.tridiag_size,
                "mll": res.mll, "inv_quad": res.inv_quad, "logdet": res.logdet,
                "kv_gflops_algorithmic": flops / (kms * 1e-3) / 1e9 * world,
            },
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        comm.close()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("GP_WORKLOAD", "c2"), choices=sorted(WORKLOADS))
    ap.add_argument("--backend", default="auto", choices=["auto", "tcgen05", "simt"])
    ap.add_argument("--ref-rows", type=int, default=12500, help="rows of the bounded CPU sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
