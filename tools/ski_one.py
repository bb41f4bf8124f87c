#!/usr/bin/env python
"""A few K_ski.V products at the BASELINE C5 shape (N = 1e6, d = 3, grid 100^3), for ncu launch lists / captures of csrc/ski.cu:
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none --csv \
        --log-file ski.csv python tools/ski_one.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gpytorch_b200.engine import Plan

dev = torch.device("cuda:0")
n, d, G = int(os.environ.get("GP_SKI_N", 1000000)), 3, 100
g = torch.Generator().manual_seed(0)
x = torch.rand(n, d, generator=g)
axes = [torch.linspace(0.0 - 1.0 / (G - 2), 1.0 + 1.0 / (G - 2), G) for _ in range(d)]
p = Plan(x.to(dev)).set_ski([G] * d, [float(a[0]) for a in axes], [float(a[1] - a[0]) for a in axes]).set_hypers("rbf", 0.2, 1.0, 0.1)
v = torch.randn(n, 11, device=dev)
for _ in range(3):
    out = p.kmv(v)
torch.cuda.synchronize()
ms = p.time_kmv_kernel(v, 2, 10)
print("ok", p.info(), float(out.abs().sum()), f"product {ms:.3f} ms")
