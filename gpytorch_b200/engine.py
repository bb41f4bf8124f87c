"""Thin torch-facing wrapper over a gp_plan.  PyTorch is used only for device memory and streams."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import _lib
from ._lib import BACKEND, KIND, MllOpts, MllResult, check


def _ptr(t: torch.Tensor | None):
    return C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr())


def _require_cuda_f32(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must live on a CUDA device: gpytorch_b200 has no CPU path")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32 (got {t.dtype}); the sm_100a engine computes in fp32/3xTF32")


@dataclass
class MbcgInfo:
    iters: int
    tridiag_size: int
    residual_norms: list
    status: int


class Plan:
    """Owns a gp_plan: inputs, hyper-parameters and workspaces for one covariance operator K(X1, X2)."""

    def __init__(self, x1: torch.Tensor, x2: torch.Tensor | None = None, backend: str = "auto",
                 row_begin: int = 0, row_count: int = 0, comm=None):
        self.lib = _lib.load()
        _require_cuda_f32(x1, "x1")
        if x1.dim() != 2:
            raise RuntimeError("x1 must be [n, d]")
        self.x1 = x1.contiguous()
        self.same = x2 is None or x2 is x1
        if not self.same:
            _require_cuda_f32(x2, "x2")
            if x2.dim() != 2 or x2.size(1) != x1.size(1):
                raise RuntimeError("x1 and x2 must have the same feature dimension")  # kernels/kernel.py:506-507
            self.x2 = x2.contiguous()
        else:
            self.x2 = self.x1
        self.device = x1.device
        self.n1, self.d = self.x1.shape
        self.n2 = self.x2.size(0)
        self.row_begin = row_begin
        self.row_count = row_count if row_count > 0 else self.n1
        self._h = C.c_void_p()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        check(self.lib.gp_plan_create(C.byref(self._h), self.device.index or 0, C.c_void_p(stream)))
        check(self.lib.gp_plan_set_backend(self._h, BACKEND[backend]))
        if comm is not None:
            check(self.lib.gp_plan_set_comm(self._h, comm.handle))
        self.comm = comm
        self.refresh_data()
        self.noise = 0.0
        self.outputscale = 1.0

    def refresh_data(self):
        """(Re-)register the input buffers: the engine re-packs its tiles (centre / scale / 3xTF32 split) from the
        CURRENT contents of x1 / x2.  Call after an in-place update of the inputs (operators._get_plan does, keyed on
        the tensors' version counters)."""
        with torch.cuda.device(self.device):
            check(self.lib.gp_plan_set_data(
                self._h, _ptr(self.x1), self.n1, self.x1.stride(0),
                _ptr(None if self.same else self.x2), self.n2, self.x2.stride(0), self.d,
                self.row_begin, self.row_count if self.row_count != self.n1 else 0))
        return self

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.gp_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- hyper-parameters --------------------------------------------------------------------
    def set_hypers(self, kind: str, lengthscale, outputscale: float = 1.0, noise: float = 0.0):
        ls = [float(v) for v in (lengthscale if hasattr(lengthscale, "__len__") else [lengthscale])]
        arr = (C.c_float * len(ls))(*ls)
        self.kind, self.lengthscale, self.outputscale, self.noise = kind, ls, float(outputscale), float(noise)
        with torch.cuda.device(self.device):
            check(self.lib.gp_plan_set_hypers(self._h, KIND[kind], arr, len(ls), float(outputscale), float(noise)))
        return self

    def set_ski(self, grid_sizes, grid_lo, grid_step):
        """SKI / KISS-GP: the operator becomes W (T_0 x ... x T_{d-1}) W^T on a regular grid (first node grid_lo[i], spacing
        grid_step[i], grid_sizes[i] nodes per dimension).  Call before set_hypers."""
        d = len(grid_sizes)
        gs = (C.c_int * d)(*[int(v) for v in grid_sizes])
        lo = (C.c_float * d)(*[float(v) for v in grid_lo])
        st = (C.c_float * d)(*[float(v) for v in grid_step])
        with torch.cuda.device(self.device):
            check(self.lib.gp_plan_set_ski(self._h, gs, lo, st, d))
        return self

    def set_sum(self, terms):
        """Kernel sum (AdditiveKernel): this plan's operator becomes sum_t K_t (+ its own noise).  `terms` are ready plans over
        the same rows (their own kind / lengthscales / outputscale / active dimensions); they must outlive this plan.  Call after
        set_hypers (only the noise of this plan is used)."""
        terms = list(terms)
        arr = (C.c_void_p * len(terms))(*[t._h.value for t in terms])
        self._terms = terms                       # keep them alive
        with torch.cuda.device(self.device):
            check(self.lib.gp_plan_set_sum(self._h, arr, len(terms)))
        return self

    def set_noise_diag(self, diag: torch.Tensor | None):
        """Per-row noise variances (FixedNoiseGaussianLikelihood): K_hat = K + diag(d).  None restores the scalar noise."""
        if diag is None:
            self._noise_diag = None
            check(self.lib.gp_plan_set_noise_diag(self._h, _ptr(None), 0))
            return self
        _require_cuda_f32(diag, "noise diagonal")
        self._noise_diag = diag.contiguous()      # keep it alive: the engine holds the raw pointer
        check(self.lib.gp_plan_set_noise_diag(self._h, _ptr(self._noise_diag), self._noise_diag.numel()))
        return self

    def info(self):
        b, s, k, m = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        check(self.lib.gp_plan_info(self._h, C.byref(b), C.byref(s), C.byref(k), C.byref(m)))
        return {"backend": {1: "tcgen05", 2: "simt", 3: "ski", 4: "sum"}.get(b.value, "?"), "nsplit": s.value, "kpad": k.value, "n_sm": m.value}

    def time_kmv_kernel(self, v: torch.Tensor, warmup: int = 3, reps: int = 20) -> float:
        """Average device time (ms) of ONE launch of the fused K.V kernel alone (CUDA events on the plan stream)."""
        v = v.contiguous()
        ms = C.c_float()
        check(self.lib.gp_time_kmv_kernel(self._h, _ptr(v), v.stride(0), v.size(1), warmup, reps, C.byref(ms)))
        return ms.value

    def launches(self) -> int:
        return int(self.lib.gp_kernel_launches(self._h))

    # ---- kernel seam -------------------------------------------------------------------------
    def kmv(self, v: torch.Tensor, add_noise: bool = False) -> torch.Tensor:
        """K(X1,X2) @ v (+ noise*v).  v [n2] or [n2, t]."""
        _require_cuda_f32(v, "rhs")
        vec = v.dim() == 1
        v2 = (v.unsqueeze(-1) if vec else v).contiguous()
        if v2.size(0) != self.n2:
            raise RuntimeError(f"Size mismatch: operator has {self.n2} columns, rhs has {v2.size(0)} rows")
        out = torch.empty(self.row_count, v2.size(1), device=self.device, dtype=torch.float32)
        check(self.lib.gp_kmv(self._h, _ptr(v2), v2.stride(0), v2.size(1), _ptr(out), out.stride(0), int(add_noise)))
        return out.squeeze(-1) if vec else out

    def rows(self, idx: torch.Tensor) -> torch.Tensor:
        idx = idx.to(device=self.device, dtype=torch.int64).contiguous()
        out = torch.empty(idx.numel(), self.n2, device=self.device, dtype=torch.float32)
        check(self.lib.gp_krows(self._h, _ptr(idx), idx.numel(), _ptr(out), out.stride(0)))
        return out

    def diag(self) -> torch.Tensor:
        out = torch.empty(self.row_count, device=self.device, dtype=torch.float32)
        check(self.lib.gp_kdiag(self._h, _ptr(out)))
        return out

    def bilinear_grad(self, left: torch.Tensor, right: torch.Tensor):
        """(d/d lengthscale[*], d/d outputscale) of sum(left * (K @ right))."""
        left = left.contiguous(); right = right.contiguous()
        s = left.size(1)
        nls = len(self.lengthscale)
        gl = (C.c_double * nls)()
        go = C.c_double()
        check(self.lib.gp_bilinear_grad(self._h, _ptr(left), left.stride(0), _ptr(right), right.stride(0), s, gl, C.byref(go)))
        return [gl[i] for i in range(nls)], go.value

    # ---- solver seam -------------------------------------------------------------------------
    def pivoted_cholesky(self, rank: int, error_tol: float = 1e-3):
        """Returns (Lt [m, n] (= L^T), pivots [m], status)."""
        rank = min(rank, self.n2)
        lt = torch.empty(rank, self.n2, device=self.device, dtype=torch.float32)
        piv = torch.empty(rank, device=self.device, dtype=torch.int64)
        r = C.c_int()
        st = check(self.lib.gp_pivoted_cholesky(self._h, rank, float(error_tol), _ptr(lt), _ptr(piv), C.byref(r)))
        return lt[: r.value], piv[: r.value], st

    def precond_build(self, lt: torch.Tensor):
        """W [n_local, k] with P^-1 v = (v - W W^T v)/noise, and log det P."""
        lt = lt.contiguous()
        k = lt.size(0)
        w = torch.empty(self.row_count, k, device=self.device, dtype=torch.float32)
        ld = C.c_double()
        st = check(self.lib.gp_precond_build(self._h, _ptr(lt), k, _ptr(w), C.byref(ld)))
        return w, ld.value, st

    def precond_probes(self, lt, eps1, eps2):
        lt = lt.contiguous(); eps1 = eps1.contiguous(); eps2 = eps2.contiguous()
        k, tp = lt.size(0), eps2.size(1)
        z = torch.empty(self.row_count, tp, device=self.device, dtype=torch.float32)
        check(self.lib.gp_precond_probes(self._h, _ptr(lt), k, _ptr(eps1), _ptr(eps2), tp, _ptr(z)))
        return z

    def mbcg(self, rhs: torch.Tensor, n_tridiag: int = 0, tolerance: float = 1.0, max_iter: int = 1000,
             max_tridiag_iter: int = 20, precond_w: torch.Tensor | None = None, warn: bool = True):
        """linear_cg on K + noise I.  rhs [n, t], t <= 16.  Returns (solves, t_mat | None, MbcgInfo)."""
        _require_cuda_f32(rhs, "rhs")
        rhs = rhs.contiguous()
        n, t = rhs.shape
        solves = torch.empty_like(rhs)
        mti = int(max_tridiag_iter)  # > max_iter is rejected by the engine like the reference does
        tmat = torch.zeros(max(n_tridiag, 1), min(mti, 4096), min(mti, 4096), device=self.device, dtype=torch.float32)
        it, js = C.c_int(), C.c_int()
        resid = (C.c_float * 16)()
        w = None if precond_w is None else precond_w.contiguous()
        st = self.lib.gp_mbcg(self._h, _ptr(rhs), rhs.stride(0), t, n_tridiag, float(tolerance), int(max_iter), int(mti),
                              _ptr(w), 0 if w is None else w.size(1), _ptr(solves), solves.stride(0), _ptr(tmat),
                              C.byref(it), C.byref(js), resid)
        check(st, warn=warn)
        info = MbcgInfo(it.value, js.value, [resid[i] for i in range(t)], st)
        tm = tmat[:n_tridiag, : js.value, : js.value] if n_tridiag else None
        return solves, tm, info

    def slq_logdet(self, tmat: torch.Tensor, n: int | None = None) -> float:
        tmat = tmat.contiguous()
        tp, j, _ = tmat.shape
        out = C.c_double()
        check(self.lib.gp_slq_logdet(self._h, _ptr(tmat), tp, j, j, int(n if n is not None else self.n2), C.byref(out)))
        return out.value

    def lanczos(self, init: torch.Tensor, max_iter: int, tol: float = 1e-5):
        """Returns (Q [n_local, J], T [J, J]); on a row-sharded plan init / Q hold this rank's rows."""
        init = init.contiguous()
        if init.numel() != self.row_count:
            raise RuntimeError(f"Lanczos start vector has {init.numel()} entries, the plan owns {self.row_count} rows")
        qt = torch.zeros(max_iter, self.row_count, device=self.device, dtype=torch.float32)
        tm = torch.zeros(max_iter, max_iter, device=self.device, dtype=torch.float32)
        j = C.c_int()
        check(self.lib.gp_lanczos(self._h, _ptr(init), int(max_iter), float(tol), _ptr(qt), _ptr(tm), C.byref(j)))
        return qt[: j.value].t(), tm[: j.value, : j.value]

    def mll(self, y_minus_mean, eps1, eps2, rademacher, num_probes=10, precond_rank=15, min_precond_size=2000,
            precond_tol=1e-3, cg_tol=1.0, max_cg_iter=1000, max_tridiag_iter=20, want_solve=False, warn=True):
        opts = MllOpts(num_probes, precond_rank, min_precond_size, precond_tol, cg_tol, max_cg_iter, max_tridiag_iter)
        res = MllResult()
        solve = torch.empty(self.row_count, device=self.device, dtype=torch.float32) if want_solve else None
        st = self.lib.gp_mll(self._h, _ptr(y_minus_mean.contiguous()), _ptr(eps1), _ptr(eps2), _ptr(rademacher),
                             C.byref(opts), _ptr(solve), C.byref(res))
        check(st, warn=warn)
        if res.status_flags & 6 and warn:   # bit 1: CG not converged, bit 2: SLQ eigen-solver not converged
            import warnings
            warnings.warn(_lib.last_error(), _lib.NumericalWarning)
        return res, solve
