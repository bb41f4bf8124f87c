#!/bin/bash
# compute-sanitizer passes over a small instance of every kernel on the hot path (run through gpurun).
# usage: tools/sanitize.sh [memcheck|racecheck|synccheck]
set -u
TOOL=${1:-memcheck}
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool "$TOOL" --print-limit 20 python - <<'PY' > gpurun_out/sanitizer_$TOOL.log 2>&1
import os, sys, torch
sys.path.insert(0, os.getcwd())
from gpytorch_b200.engine import Plan
from oracle import mll as om
dev = torch.device("cuda:0")
for backend, n, d, kind, rank in (("tcgen05", 2304, 10, "rbf", 20), ("tcgen05", 2100, 20, "matern52", 20), ("simt", 2100, 6, "matern32", 10)):
    x, y = om.synthetic_problem(n, d, 0, torch.float32)
    pn = om.make_probe_noise(n, rank, 10, 1)
    p = Plan(x.to(dev), backend=backend).set_hypers(kind, 1.0, 1.0, 0.1)
    res, _ = p.mll(y.to(dev), pn[0].to(dev), pn[1].to(dev), pn[2].to(dev), 10, rank, 2000)
    out = p.kmv(torch.randn(n, 11, device=dev), True)
    q, t = p.lanczos(torch.randn(n, device=dev), 8)
    gl, go = p.bilinear_grad(torch.randn(n, 3, device=dev), torch.randn(n, 3, device=dev))
    torch.cuda.synchronize()
    print(backend, kind, "mll", res.mll, "iters", res.cg_iters)
    p.close()
print("sanitizer workload done")
PY
echo "rc=$?"; grep -E "ERROR SUMMARY|done|mll" gpurun_out/sanitizer_$TOOL.log | tail -8
