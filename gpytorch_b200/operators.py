"""Engine-backed covariance operators: the LinearOperator duck-type subset the exact-GP path calls.

Mirrors (reference paths under /root/reference/gpytorch):
  * KernelLinearOperator returned by a kernel's forward, the KeOps plug-in pattern
    (kernels/keops/rbf_kernel.py:44-55, keops/matern_kernel.py:68-80) -- `evaluate_kernel()` hands the
    operator itself to the solver (lazy/lazy_evaluated_kernel_tensor.py:345-373);
  * LazyEvaluatedKernelTensor hooks: _matmul (:245-276), _getitem (:136-243), _diagonal (:107-133),
    _size (:277-318), _bilinear_derivative (:69-105);
  * AddedDiagLinearOperator / InvQuadLogdet / solve of linear_operator,
    reached from distributions/multivariate_normal.py:248-249 and
    models/exact_prediction_strategies.py:286,444.

All arithmetic runs in libgpbbmm (CUDA); torch supplies memory, streams and the autograd graph.
"""
from __future__ import annotations

import torch

from . import settings
from .engine import Plan


def _as_list(ls: torch.Tensor):
    return [float(v) for v in ls.detach().reshape(-1).tolist()]


# Plans own device workspaces (packed tiles, CG vectors); re-use them across operators over the same buffers so a
# training loop does not re-allocate every step.  Hyper-parameters / data are re-packed whenever a new operator
# first touches a cached plan.
_PLAN_CACHE: "dict[tuple, Plan]" = {}
_PLAN_CACHE_MAX = 64          # a batch of 16 independent operators (+ their cross-covariances) must fit
_PLAN_LOCK = __import__("threading").RLock()   # BatchLinearOperator drives the cache from worker threads


def _get_plan(x1, x2, backend, row_begin, row_count, comm, slot=0) -> Plan:
    if not x1.is_cuda:
        raise RuntimeError("x1 must live on a CUDA device: gpytorch_b200 has no CPU path")
    with _PLAN_LOCK:
        return _get_plan_locked(x1, x2, backend, row_begin, row_count, comm, slot)


def _get_plan_locked(x1, x2, backend, row_begin, row_count, comm, slot=0) -> Plan:
    # a plan enqueues on the stream that was current when it was created: the stream is part of the identity; so is the slot:
    # the terms of a kernel sum over the SAME inputs need one plan each (each holds its own packed lengthscales)
    key = (x1.data_ptr(), tuple(x1.shape), x1.stride(0), None if x2 is None else (x2.data_ptr(), tuple(x2.shape), x2.stride(0)),
           backend, str(x1.device), row_begin, row_count, id(comm), torch.cuda.current_stream(x1.device).cuda_stream, slot)
    plan = _PLAN_CACHE.pop(key, None)
    src_versions = (x1._version, None if x2 is None else x2._version)
    if plan is not None and plan._src_versions != src_versions:
        # same buffers, new contents (x.copy_(new), an optimiser step on the inputs, ...): the packed tiles are stale.
        # torch bumps a tensor's version counter on every in-place write, so this is exact, not a heuristic.
        plan.x1 = x1.contiguous()
        plan.x2 = plan.x1 if x2 is None else x2.contiguous()
        plan.refresh_data()
    if plan is None:
        plan = Plan(x1, x2, backend="auto" if backend.startswith("ski:") else backend, row_begin=row_begin, row_count=row_count, comm=comm)
        plan._hyp_key = None
    plan._src_versions = src_versions
    _PLAN_CACHE[key] = plan  # most recently used last
    while len(_PLAN_CACHE) > _PLAN_CACHE_MAX:
        _PLAN_CACHE.pop(next(iter(_PLAN_CACHE))).close()
    return plan


def clear_plan_cache():
    with _PLAN_LOCK:
        while _SUM_PLANS:
            _SUM_PLANS.popitem()[1].close()
        while _PLAN_CACHE:
            _PLAN_CACHE.popitem()[1].close()


class ConstantDiagLinearOperator:
    """sigma^2 I (likelihoods/noise_models.py:57-92 returns this for homoskedastic noise)."""

    def __init__(self, diag_value: torch.Tensor, diag_shape: int):
        self.diag_value = diag_value
        self.n = int(diag_shape)

    def __add__(self, other):
        if isinstance(other, ConstantDiagLinearOperator):
            return ConstantDiagLinearOperator(self.diag_value + other.diag_value, self.n)
        if isinstance(other, DiagLinearOperator):
            return other + self
        return NotImplemented

    @property
    def shape(self):
        return torch.Size([self.n, self.n])

    def to_dense(self):
        return self.diag_value.reshape(()) * torch.eye(self.n, device=self.diag_value.device, dtype=self.diag_value.dtype)


class DiagLinearOperator:
    """diag(d) with a per-row vector d (FixedGaussianNoise, likelihoods/noise_models.py:150-190)."""

    def __init__(self, diag: torch.Tensor):
        self.diag_vec = diag
        self.n = int(diag.shape[-1])

    @property
    def shape(self):
        return torch.Size([self.n, self.n])

    def to_dense(self):
        return torch.diag_embed(self.diag_vec)

    def __add__(self, other):
        if isinstance(other, ConstantDiagLinearOperator):
            return DiagLinearOperator(self.diag_vec + other.diag_value.reshape(-1)[:1])
        if isinstance(other, DiagLinearOperator):
            return DiagLinearOperator(self.diag_vec + other.diag_vec)
        return NotImplemented


class KernelLinearOperator:
    """K(x1, x2) (outputscale folded in) that never materialises: every product is the fused CUDA kernel."""

    def __init__(self, x1, x2, kind, lengthscale, outputscale=None, plan: Plan | None = None, comm=None,
                 row_begin=0, row_count=0):
        self.x1, self.x2 = x1, x2
        self.kind = kind
        self.lengthscale = lengthscale          # tensor (scalar or [d]); may require grad
        self.outputscale = outputscale if outputscale is not None else torch.ones((), device=x1.device)
        self._plan = plan
        self._comm = comm
        self._row_begin, self._row_count = row_begin, row_count
        self.same = x2 is None or x2 is x1 or (x1.shape == x2.shape and x1.data_ptr() == x2.data_ptr())

    # -- plumbing --
    def _host_hypers(self, noise_t=None):
        """(lengthscale list, outputscale, noise) as host floats with ONE device->host read per operator (cached: the
        parameter tensors of an operator never change; a new forward builds a new operator)."""
        cached = getattr(self, "_hyp_host", None)
        if cached is None or (noise_t is not None and cached[3] is not noise_t):
            parts = [self.lengthscale.detach().reshape(-1).float(), self.outputscale.detach().reshape(1).float()]
            if noise_t is not None:
                parts.append(noise_t.detach().reshape(-1)[:1].float())
            vals = torch.cat(parts).tolist()
            nl = self.lengthscale.numel()
            cached = (vals[:nl], vals[nl], vals[nl + 1] if noise_t is not None else None, noise_t)
            self._hyp_host = cached
        return cached

    def plan(self, noise=0.0) -> Plan:
        """noise: a host float, or the device tensor of the likelihood (read together with the other hyper-parameters)."""
        fresh = False
        if self._plan is None:
            self._plan = _get_plan(self.x1, None if self.same else self.x2, settings.backend.value(),
                                   self._row_begin, self._row_count, self._comm, getattr(self, "_plan_slot", 0))
            fresh = True
        if torch.is_tensor(noise):
            ls, os_, nz, _ = self._host_hypers(noise)
        else:
            ls, os_, _, _ = self._host_hypers(None)
            nz = float(noise)
        key = (self.kind, tuple(ls), os_, nz)
        if fresh or getattr(self._plan, "_hyp_key", None) != key:
            self._plan.set_hypers(self.kind, ls, os_, nz)
            self._plan._hyp_key = key
        return self._plan

    @property
    def shape(self):
        n2 = self.x1.size(0) if self.same else self.x2.size(0)
        return torch.Size([self.x1.size(0), n2])

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    @property
    def dtype(self):
        return self.x1.dtype

    @property
    def device(self):
        return self.x1.device

    @property
    def batch_shape(self):
        return torch.Size([])

    @property
    def requires_grad(self):
        return bool(self.lengthscale.requires_grad or self.outputscale.requires_grad)

    def evaluate_kernel(self):
        return self  # "meta LinearOperator" branch, lazy_evaluated_kernel_tensor.py:345-348

    def representation(self):
        return (self.x1, self.x2, self.lengthscale, self.outputscale)

    def hyper_tensors(self):
        """The differentiable hyper-parameters of this operator, in the order _bilinear_derivative_list returns their
        gradients (the autograd functions below take them as explicit inputs)."""
        return [self.lengthscale, self.outputscale]

    def _bilinear_derivative_list(self, left, right):
        gl, go = self._bilinear_derivative(left, right)
        return [gl.reshape(self.lengthscale.shape), go.reshape(self.outputscale.shape)]

    # -- LinearOperator protocol: shape helpers / transpose (lazy_evaluated_kernel_tensor.py:277-341) --
    def _size(self):
        return self.shape

    @property
    def matrix_shape(self):
        return self.shape

    def dim(self):
        return 2

    ndimension = dim

    def numel(self):
        return self.shape[0] * self.shape[1]

    def _transpose_nonbatch(self):
        """K(x1, x2)^T = K(x2, x1) for every stationary kernel on this path (no data is moved)."""
        if self.same:
            return self
        return KernelLinearOperator(self.x2, self.x1, self.kind, self.lengthscale, self.outputscale)

    def transpose(self, dim1, dim2):
        return self if dim1 % 2 == dim2 % 2 else self._transpose_nonbatch()

    def t(self):
        return self._transpose_nonbatch()

    @property
    def mT(self):
        return self._transpose_nonbatch()

    def detach(self):
        return KernelLinearOperator(self.x1, self.x2, self.kind, self.lengthscale.detach(), self.outputscale.detach())

    # -- products --
    def matmul(self, rhs):
        return _KernelMatmul.apply(self, rhs, *self.hyper_tensors())

    __matmul__ = matmul
    _matmul = matmul

    def to_dense(self):
        p = self.plan()
        idx = torch.arange(p.row_count, device=self.device)   # local rows of a row-sharded plan
        return p.rows(idx)

    def diagonal(self, dim1=-2, dim2=-1):
        """kernel(x1, x2, diag=True) (lazy_evaluated_kernel_tensor.py:107-133): the constant outputscale for x2 == x1,
        k(x1_i, x2_i) for a cross-covariance of equal sizes."""
        if not self.same and self.x1.size(0) != self.x2.size(0):
            raise RuntimeError(f"diagonal of a non-square operator {tuple(self.shape)} is undefined")
        return self.plan().diag()

    _diagonal = diagonal

    def _getitem(self, row_index, col_index, *batch_indices):
        return self[row_index, col_index]

    def __getitem__(self, index):
        """Row / column slicing by re-indexing x1 / x2 (lazy_evaluated_kernel_tensor.py:136-243)."""
        if not isinstance(index, tuple):
            index = (index, slice(None))
        ri, ci = index
        if isinstance(ri, int):
            return self.plan().rows(torch.tensor([ri], device=self.device))[0][ci]
        x1 = self.x1[ri]
        x2 = (self.x1 if self.same else self.x2)[ci]
        return KernelLinearOperator(x1, x2, self.kind, self.lengthscale, self.outputscale)

    def __add__(self, other):
        if isinstance(other, (ConstantDiagLinearOperator, DiagLinearOperator)):
            return AddedDiagLinearOperator(self, other)
        if isinstance(other, KernelLinearOperator) and not isinstance(other, SKIKernelLinearOperator) \
                and not isinstance(self, SKIKernelLinearOperator):
            return SumKernelLinearOperator([self, other])     # K_1 + K_2 stays lazy: one engine operator (csrc/sum.cu)
        raise NotImplementedError("a kernel operator adds a (Constant)DiagLinearOperator or another kernel operator")

    def add_jitter(self, jitter_val=1e-3):
        return AddedDiagLinearOperator(self, ConstantDiagLinearOperator(torch.tensor(jitter_val, device=self.device), self.shape[0]))

    def _bilinear_derivative(self, left, right):
        """(d/d lengthscale, d/d outputscale) of sum(left * (K @ right)); lazy_evaluated_kernel_tensor.py:69-105."""
        gl, go = self.plan(getattr(self, "_last_noise", 0.0)).bilinear_grad(left, right)
        return torch.tensor(gl, device=self.device, dtype=self.dtype), torch.tensor(go, device=self.device, dtype=self.dtype)


class SumKernelLinearOperator(KernelLinearOperator):
    """K_1 + ... + K_m over the same inputs (AdditiveKernel, kernels/kernel.py:592-621).  The reference evaluates every term
    densely and adds the matrices; here the sum is ONE engine operator (gp_plan_set_sum): a product launches the fused kernel of
    every term into disjoint partial slots, the solves / preconditioner / SLQ run on the sum, and the hyper-parameter gradients
    of each term come from that term's own bilinear derivative with the shared left / right factors."""

    def __init__(self, ops):
        ops = list(ops)
        flat = []
        for o in ops:
            flat.extend(o.ops if isinstance(o, SumKernelLinearOperator) else [o])
        if not 1 <= len(flat) <= 4:
            raise RuntimeError(f"a kernel sum takes 1 to 4 terms (got {len(flat)})")
        first = flat[0]
        for o in flat[1:]:
            if o.shape != first.shape or o.same != first.same:
                raise RuntimeError(f"cannot add kernels of shapes {tuple(first.shape)} and {tuple(o.shape)}")
        # one engine plan per term even when the terms see the same inputs: re-wrap (the caller's operators stay usable on their
        # own) and give every term its own slot of the plan cache
        self.ops = []
        for i, o in enumerate(flat):
            t = KernelLinearOperator(o.x1, o.x2, o.kind, o.lengthscale, o.outputscale, comm=o._comm, row_begin=o._row_begin, row_count=o._row_count)
            t._plan_slot = 1 + i
            self.ops.append(t)
        self.x1, self.x2, self.same = first.x1, first.x2, first.same
        self.kind = "sum"
        self.lengthscale, self.outputscale = first.lengthscale, first.outputscale   # representative only (device / dtype)
        self._comm, self._row_begin, self._row_count = first._comm, first._row_begin, first._row_count
        self._plan = None

    def hyper_tensors(self):
        return [t for o in self.ops for t in o.hyper_tensors()]

    def _bilinear_derivative_list(self, left, right):
        out = []
        for o in self.ops:
            o._last_noise = 0.0
            out.extend(o._bilinear_derivative_list(left, right))
        return out

    @property
    def requires_grad(self):
        return any(o.requires_grad for o in self.ops)

    def representation(self):
        return tuple(t for o in self.ops for t in o.representation())

    def plan(self, noise=0.0) -> Plan:
        terms = [o.plan(0.0) for o in self.ops]
        nz = float(noise.detach().reshape(-1)[0]) if torch.is_tensor(noise) else float(noise)
        with _PLAN_LOCK:
            key = ("sum", tuple(id(t) for t in terms))
            parent = _SUM_PLANS.pop(key, None)
            if parent is None:
                parent = Plan(self.x1, None if self.same else self.x2, backend="auto", row_begin=self._row_begin,
                              row_count=self._row_count, comm=self._comm)
                parent._hyp_key = None
            _SUM_PLANS[key] = parent
            while len(_SUM_PLANS) > 16:
                _SUM_PLANS.pop(next(iter(_SUM_PLANS))).close()
        # the terms may have been re-created / re-packed since the last use: (re-)attach them every time (validation + a
        # 64-float upload, no allocation), then set the noise of the sum
        parent.set_sum(terms)
        if parent._hyp_key != nz:
            parent.set_hypers("rbf", [1.0], 1.0, nz)
            parent._hyp_key = nz
        self._plan = parent
        return parent

    def _transpose_nonbatch(self):
        return self if self.same else SumKernelLinearOperator([o._transpose_nonbatch() for o in self.ops])

    def detach(self):
        return SumKernelLinearOperator([o.detach() for o in self.ops])

    def to_dense(self):
        out = self.ops[0].to_dense()
        for o in self.ops[1:]:
            out = out + o.to_dense()
        return out

    def diagonal(self, dim1=-2, dim2=-1):
        out = self.ops[0].diagonal()
        for o in self.ops[1:]:
            out = out + o.diagonal()
        return out

    _diagonal = diagonal

    def __getitem__(self, index):
        parts = [o[index] for o in self.ops]
        if torch.is_tensor(parts[0]):
            return sum(parts[1:], parts[0])
        return SumKernelLinearOperator(parts)

    def _bilinear_derivative(self, left, right):
        raise NotImplementedError("a kernel sum has one (lengthscale, outputscale) pair per term: use _bilinear_derivative_list")


_SUM_PLANS: "dict[tuple, Plan]" = {}


class SKIKernelLinearOperator(KernelLinearOperator):
    """K_ski = s W (T_0 x ... x T_{d-1}) W^T (kernels/grid_interpolation_kernel.py:132-213): the engine's SKI backend -- cubic
    interpolation weights kept in compact per-dimension form, Kronecker / Toeplitz grid covariance applied mode by mode."""

    def __init__(self, x1, kind, lengthscale, outputscale, grid_sizes, grid_lo, grid_step):
        super().__init__(x1, None, kind, lengthscale, outputscale)
        self.grid_sizes, self.grid_lo, self.grid_step = tuple(grid_sizes), tuple(grid_lo), tuple(grid_step)

    def plan(self, noise=0.0) -> Plan:
        if self._plan is None:
            key = "ski:" + repr((self.grid_sizes, self.grid_lo, self.grid_step))
            self._plan = _get_plan(self.x1, None, key, 0, 0, None)
            if getattr(self._plan, "_ski_key", None) != key:
                self._plan.set_ski(self.grid_sizes, self.grid_lo, self.grid_step)
                self._plan._ski_key = key
                self._plan._hyp_key = None
        if torch.is_tensor(noise):
            ls, os_, nz, _ = self._host_hypers(noise)
        else:
            ls, os_, _, _ = self._host_hypers(None)
            nz = float(noise)
        hk = (self.kind, tuple(ls), os_, nz)
        if getattr(self._plan, "_hyp_key", None) != hk:
            self._plan.set_hypers(self.kind, ls, os_, nz)
            self._plan._hyp_key = hk
        return self._plan

    def detach(self):
        return SKIKernelLinearOperator(self.x1, self.kind, self.lengthscale.detach(), self.outputscale.detach(), self.grid_sizes,
                                       self.grid_lo, self.grid_step)

    def to_dense(self):
        n = self.x1.size(0)
        eye = torch.eye(n, device=self.device, dtype=torch.float32)
        return torch.cat([self.plan().kmv(eye[:, c0:c0 + 16].contiguous()) for c0 in range(0, n, 16)], -1)

    def diagonal(self, dim1=-2, dim2=-1):
        raise NotImplementedError("diagonal of the SKI operator")

    def __getitem__(self, index):
        """Rows / columns of the interpolated operator: K[r, c] = W[r] K_uu W[c]^T (the reference slices the interpolation
        indices / values of InterpolatedLinearOperator).  Used by the prediction strategy on the joint train + test operator."""
        if not isinstance(index, tuple):
            index = (index, slice(None))
        ri, ci = index
        n = self.x1.size(0)
        ar = torch.arange(n, device=self.device)
        rows = ar[ri].reshape(-1)
        cols = ar[ci].reshape(-1)
        sub = _SKISliceOperator(self, rows, cols)
        return sub.to_dense()[0] if isinstance(ri, int) else sub


class _SKISliceOperator:
    """K_ski[rows, cols] of a square SKI operator, never materialised: a product zero-pads the right-hand side to the full point
    set, runs the parent's scatter / mode products / gather, and keeps the requested rows (evaluation mode: no autograd)."""

    def __init__(self, parent, rows, cols):
        self.parent, self.rows, self.cols = parent, rows, cols

    @property
    def shape(self):
        return torch.Size([self.rows.numel(), self.cols.numel()])

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    @property
    def device(self):
        return self.parent.device

    @property
    def dtype(self):
        return self.parent.dtype

    def evaluate_kernel(self):
        return self

    def _transpose_nonbatch(self):
        return _SKISliceOperator(self.parent, self.cols, self.rows)      # the parent is symmetric

    t = _transpose_nonbatch

    def transpose(self, dim1, dim2):
        return self if dim1 % 2 == dim2 % 2 else self._transpose_nonbatch()

    def matmul(self, rhs):
        vec = rhs.dim() == 1
        r2 = (rhs.unsqueeze(-1) if vec else rhs).detach().float()
        if r2.size(0) != self.cols.numel():
            raise RuntimeError(f"LinearOperator (size={tuple(self.shape)}) cannot be multiplied with right-hand-side Tensor "
                               f"(size={tuple(rhs.shape)})")
        plan = self.parent.plan(getattr(self.parent, "_last_noise", 0.0))
        outs = []
        for c0 in range(0, r2.size(1), 16):
            full = torch.zeros(self.parent.x1.size(0), min(16, r2.size(1) - c0), device=self.device)
            full[self.cols] = r2[:, c0:c0 + 16]
            outs.append(plan.kmv(full)[self.rows])
        out = outs[0] if len(outs) == 1 else torch.cat(outs, -1)
        return out.squeeze(-1) if vec else out

    __matmul__ = matmul
    _matmul = matmul

    def to_dense(self):
        return self.matmul(torch.eye(self.cols.numel(), device=self.device))


class _KernelMatmul(torch.autograd.Function):
    @staticmethod
    def forward(ctx, op, rhs, *hypers):
        ctx.op = op
        out = op.plan(getattr(op, "_last_noise", 0.0)).kmv(rhs.detach())
        ctx.save_for_backward(rhs.detach())
        return out

    @staticmethod
    def backward(ctx, grad_out):
        op = ctx.op
        (rhs,) = ctx.saved_tensors
        g = grad_out.contiguous()
        vec = g.dim() == 1
        g2 = g.unsqueeze(-1) if vec else g
        r2 = rhs.unsqueeze(-1) if vec else rhs
        grad_rhs = None
        nh = len(ctx.needs_input_grad) - 2
        grads = [None] * nh
        if ctx.needs_input_grad[1]:
            # K^T g: for x1 == x2 the operator is symmetric
            if not op.same:
                grad_rhs = op._transpose_nonbatch().plan().kmv(g)
            else:
                grad_rhs = op.plan(getattr(op, "_last_noise", 0.0)).kmv(g)
        if any(ctx.needs_input_grad[2:]):
            gs = op._bilinear_derivative_list(g2, r2)
            grads = [gi if need else None for gi, need in zip(gs, ctx.needs_input_grad[2:])]
        return (None, grad_rhs, *grads)


class AddedDiagLinearOperator:
    """K + D with the BBMM solves (linear_operator AddedDiagLinearOperator): D = sigma^2 I (constant-diagonal branch, Appendix
    A.4) or a per-row diagonal (FixedNoiseGaussianLikelihood; the non-constant-diagonal branch of the preconditioner)."""

    def __init__(self, kernel_op: KernelLinearOperator, diag):
        if kernel_op.shape[0] != kernel_op.shape[1]:
            raise RuntimeError("AddedDiagLinearOperator needs a square operator")
        self.kernel_op = kernel_op
        self.diag = diag
        self.per_row = isinstance(diag, DiagLinearOperator)
        self._precond_cache = None

    @property
    def noise(self) -> torch.Tensor:
        """sigma^2 (0-d) for the constant diagonal, the vector d [n] for a per-row diagonal."""
        return self.diag.diag_vec if self.per_row else self.diag.diag_value.reshape(())

    @property
    def _noise_param(self) -> torch.Tensor:
        return self.diag.diag_vec if self.per_row else self.diag.diag_value

    def _noise_col(self):
        return self.diag.diag_vec.unsqueeze(-1) if self.per_row else self.noise

    @property
    def shape(self):
        return self.kernel_op.shape

    def size(self, dim=None):
        return self.kernel_op.size(dim)

    @property
    def dtype(self):
        return self.kernel_op.dtype

    @property
    def device(self):
        return self.kernel_op.device

    @property
    def batch_shape(self):
        return torch.Size([])

    def evaluate_kernel(self):
        return self

    def _plan(self) -> Plan:
        if self.per_row:
            p = self.kernel_op.plan(0.0)
            d = self.diag.diag_vec.detach().float().contiguous()
            if getattr(p, "_noise_diag", None) is None or p._noise_diag.data_ptr() != d.data_ptr() or p._noise_diag_version != d._version:
                p.set_noise_diag(d)
                p._noise_diag_version = d._version
            self.kernel_op._last_noise = 0.0
            return p
        p = self.kernel_op.plan(self.diag.diag_value)
        if getattr(p, "_noise_diag", None) is not None:   # a cached plan last used with a per-row diagonal
            p.set_noise_diag(None)
        self.kernel_op._last_noise = p.noise
        return p

    def matmul(self, rhs):
        d = self._noise_col() if (self.per_row and rhs.dim() > 1) else self.noise
        return self.kernel_op.matmul(rhs) + d * rhs

    __matmul__ = matmul
    _matmul = matmul

    # -- LinearOperator protocol subset the callers use (SURVEY.md section 8b "Operator seam") --
    def _size(self):
        return self.shape

    @property
    def matrix_shape(self):
        return self.shape

    def dim(self):
        return 2

    ndimension = dim

    def numel(self):
        return self.shape[0] * self.shape[1]

    @property
    def requires_grad(self):
        return bool(self.kernel_op.requires_grad or self._noise_param.requires_grad)

    def representation(self):
        return self.kernel_op.representation() + (self._noise_param,)

    def transpose(self, dim1, dim2):
        return self          # K + sigma^2 I is symmetric

    def t(self):
        return self

    _transpose_nonbatch = t

    @property
    def mT(self):
        return self

    def detach(self):
        d = DiagLinearOperator(self.diag.diag_vec.detach()) if self.per_row else ConstantDiagLinearOperator(self.diag.diag_value.detach(), self.shape[0])
        return AddedDiagLinearOperator(self.kernel_op.detach(), d)

    def __add__(self, other):
        if isinstance(other, (ConstantDiagLinearOperator, DiagLinearOperator)):   # (K + D1) + D2
            return AddedDiagLinearOperator(self.kernel_op, self.diag + other)
        raise NotImplementedError("AddedDiagLinearOperator only adds a (Constant)DiagLinearOperator")

    def logdet(self):
        return self.inv_quad_logdet(None, logdet=True)[1]

    def inv_quad(self, inv_quad_rhs, reduce_inv_quad=True):
        return self.inv_quad_logdet(inv_quad_rhs, logdet=False, reduce_inv_quad=reduce_inv_quad)[0]

    def to_dense(self):
        return self.kernel_op.to_dense() + self.diag.to_dense()

    def diagonal(self, dim1=-2, dim2=-1):
        return self.kernel_op.diagonal() + self.noise

    _diagonal = diagonal

    def add_jitter(self, jitter_val=1e-3):
        jit = ConstantDiagLinearOperator(torch.as_tensor(jitter_val, device=self.device, dtype=self.dtype), self.shape[0])
        return AddedDiagLinearOperator(self.kernel_op, self.diag + jit)

    # -- preconditioner (AddedDiagLinearOperator._preconditioner, Appendix A.4) --
    def _preconditioner(self):
        """Returns (W [n,k] | None, Lt [k,n] | None, logdet_P)."""
        n = self.shape[0]
        if settings.max_preconditioner_size.value() == 0 or n < settings.min_preconditioning_size.value():
            return None, None, 0.0
        if isinstance(self.kernel_op, SKIKernelLinearOperator):
            return None, None, 0.0      # the reference's interpolated operator has no pivoted-Cholesky preconditioner either
        if self._precond_cache is None:
            p = self._plan()
            lt, piv, st = p.pivoted_cholesky(settings.max_preconditioner_size.value(), settings.preconditioner_tolerance.value())
            if st != 0 or lt.size(0) == 0:
                self._precond_cache = (None, None, 0.0)
            else:
                w, logdet, st2 = p.precond_build(lt)
                self._precond_cache = (None, None, 0.0) if st2 != 0 else (w, lt, logdet)
        return self._precond_cache

    def _probes(self, lt, tp):
        n = self.shape[0]
        seed = settings.probe_seed.value()
        if seed is None and settings.deterministic_probes.on():
            seed = settings.deterministic_probes.seed       # one set of base samples for every estimate while the flag is on
        gen = None
        if seed is not None:
            gen = torch.Generator(device=self.device).manual_seed(int(seed))
        if lt is None:
            z = torch.randint(0, 2, (n, tp), device=self.device, generator=gen).to(torch.float32) * 2 - 1
            return z
        eps1 = torch.randn(lt.size(0), tp, device=self.device, generator=gen)
        eps2 = torch.randn(n, tp, device=self.device, generator=gen)
        return self._plan().precond_probes(lt, eps1, eps2)

    def _dense_cholesky(self):
        return torch.linalg.cholesky(self.to_dense())

    def solve(self, rhs, lhs=None):
        """K_hat^{-1} rhs by preconditioned CG (LinearOperator.solve -> linear_cg, n_tridiag = 0)."""
        out = _Solve.apply(self, rhs, self._noise_param, *self.kernel_op.hyper_tensors())
        return out if lhs is None else lhs @ out

    def inv_quad_logdet(self, inv_quad_rhs=None, logdet=False, reduce_inv_quad=True):
        """(rhs^T K_hat^{-1} rhs, log det K_hat): distributions/multivariate_normal.py:249."""
        if inv_quad_rhs is not None and inv_quad_rhs.size(0) != self.shape[0]:
            raise RuntimeError(
                f"LinearOperator (size={tuple(self.shape)}) cannot be multiplied with right-hand-side Tensor "
                f"(size={tuple(inv_quad_rhs.shape)})")
        rhs = inv_quad_rhs
        if rhs is not None and rhs.dim() == 1:
            rhs = rhs.unsqueeze(-1)
        iq, ld = _InvQuadLogdet.apply(self, rhs, bool(logdet), self._noise_param, *self.kernel_op.hyper_tensors())
        if rhs is None:
            iq = torch.empty(0, device=self.device)
        elif reduce_inv_quad:
            iq = iq.sum(-1)
        return iq, (ld if logdet else None)

    def root_inv_decomposition(self, initial_vectors=None):
        """Lanczos root of K_hat^{-1}: R with R R^T ~= K_hat^{-1} (exact_prediction_strategies.py:268-272)."""
        n = self.shape[0]
        if settings.fast_computations.covar_root_decomposition.off():
            # exact root through the dense Cholesky factor: K_hat = L L^T  =>  K_hat^{-1} = L^{-T} L^{-1}
            chol = self._dense_cholesky()
            return torch.linalg.solve_triangular(chol.transpose(-1, -2), torch.eye(n, device=self.device, dtype=chol.dtype), upper=True)
        p = self._plan()
        init = initial_vectors if initial_vectors is not None else torch.randn(n, device=self.device)
        if init.dim() == 2:
            init = init[:, 0]
        q, t = p.lanczos(init.float(), settings.max_root_decomposition_size.value())
        evals, evecs = torch.linalg.eigh(t.double())      # J x J, J <= 100: plumbing-sized
        mask = evals >= 0                                 # lanczos_tridiag_to_diag masks negative Ritz values
        evecs = evecs * mask
        evals = evals.masked_fill(~mask, 1.0)
        return (q.double() @ (evecs / evals.sqrt())).float()


class BatchLinearOperator:
    """One leading batch dimension of independent operators (BASELINE config 4: batch = 16, independent hyper-parameters per
    element; the reference broadcasts every LinearOperator op over leading dims, kernels/kernel.py:119-121,
    distributions/multivariate_normal.py:236-245).  Each element owns an engine plan on its own CUDA stream; solver calls
    (inv_quad_logdet / solve) of all elements run concurrently from a thread pool (the C ABI releases the GIL), so the small
    per-element grids and launch chains overlap on the device."""

    _pool = None

    def __init__(self, ops):
        self.ops = list(ops)
        first = self.ops[0]
        self._mshape = first.shape

    @property
    def batch_shape(self):
        return torch.Size([len(self.ops)])

    @property
    def shape(self):
        return torch.Size([len(self.ops), *self._mshape])

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    @property
    def device(self):
        return self.ops[0].device

    @property
    def dtype(self):
        return self.ops[0].dtype

    def evaluate_kernel(self):
        return self

    def __getitem__(self, b):
        return self.ops[b]

    def __add__(self, other):
        # batched homoskedastic noise: ConstantDiagLinearOperator with diag_value [B, 1]
        if isinstance(other, ConstantDiagLinearOperator):
            dv = other.diag_value
            outs = []
            for b, op in enumerate(self.ops):
                d_b = dv[b] if dv.dim() >= 2 or (dv.dim() == 1 and dv.numel() == len(self.ops) and dv.numel() > 1) else dv
                outs.append(op + ConstantDiagLinearOperator(d_b.reshape(-1)[:1], other.n))
            return BatchLinearOperator(outs)
        if isinstance(other, DiagLinearOperator):
            dv = other.diag_vec
            return BatchLinearOperator([op + DiagLinearOperator(dv[b] if dv.dim() == 2 else dv) for b, op in enumerate(self.ops)])
        raise NotImplementedError("BatchLinearOperator only adds a (Constant)DiagLinearOperator")

    def _map(self, fn):
        """Run fn(b, op) for every element concurrently, element b on its own stream."""
        import concurrent.futures as cf

        B = len(self.ops)
        dev = self.device
        main = torch.cuda.current_stream(dev)
        streams = _batch_streams(dev, B)
        grad = torch.is_grad_enabled()
        ctxs = settings.snapshot()

        def work(b):
            with torch.cuda.device(dev), torch.cuda.stream(streams[b]), torch.set_grad_enabled(grad), settings.restore(ctxs):
                return fn(b, self.ops[b])

        for st in streams:
            st.wait_stream(main)
        if BatchLinearOperator._pool is None:
            BatchLinearOperator._pool = cf.ThreadPoolExecutor(max_workers=16, thread_name_prefix="gpbatch")
        outs = list(BatchLinearOperator._pool.map(work, range(B)))
        for st in streams:
            main.wait_stream(st)
        return outs

    def matmul(self, rhs):
        return torch.stack(self._map(lambda b, op: op.matmul(rhs[b] if rhs.dim() == 3 else rhs)))

    __matmul__ = matmul

    def diagonal(self, dim1=-2, dim2=-1):
        return torch.stack([op.diagonal() for op in self.ops])

    def to_dense(self):
        return torch.stack([op.to_dense() for op in self.ops])

    def solve(self, rhs, lhs=None):
        return torch.stack(self._map(lambda b, op: op.solve(rhs[b] if rhs.dim() >= 2 and rhs.size(0) == len(self.ops) else rhs)))

    def inv_quad_logdet(self, inv_quad_rhs=None, logdet=False, reduce_inv_quad=True):
        def one(b, op):
            r = None if inv_quad_rhs is None else inv_quad_rhs[b]
            return op.inv_quad_logdet(r, logdet=logdet, reduce_inv_quad=reduce_inv_quad)

        outs = self._map(one)
        iq = torch.stack([o[0] for o in outs])
        ld = torch.stack([o[1] for o in outs]) if logdet else None
        return iq, ld


_BATCH_STREAMS = {}


def _batch_streams(dev, n):
    key = str(dev)
    pool = _BATCH_STREAMS.setdefault(key, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(dev))
    return pool[:n]


def _cg_tolerance():
    return settings.eval_cg_tolerance.value() if settings._use_eval_tolerance.on() else settings.cg_tolerance.value()


def _run_cg(op: AddedDiagLinearOperator, rhs, n_tridiag, w):
    """linear_cg over <= 16 columns per call."""
    p = op._plan()
    outs, tmat, iters = [], None, 0
    max_iter = settings.max_cg_iterations.value()
    if settings.terminate_cg_by_size.on():
        max_iter = min(max_iter, op.shape[0])
    for c0 in range(0, rhs.size(1), 16):
        blk = rhs[:, c0 : c0 + 16].contiguous()
        nt = n_tridiag if c0 == 0 else 0
        s, tm, info = p.mbcg(blk, nt, _cg_tolerance(), max_iter,
                             settings.max_lanczos_quadrature_iterations.value(), w)
        outs.append(s)
        iters = max(iters, info.iters)
        if nt:
            tmat = tm
    if settings.verbose_linalg.on():
        settings.verbose_linalg.logger.debug(
            f"Running CG on a {tuple(rhs.shape)} RHS for {iters} iterations (tol={_cg_tolerance()}). Output: {tuple(rhs.shape)}.")
    return (outs[0] if len(outs) == 1 else torch.cat(outs, -1)), tmat, iters


def _dense_branch(n: int, fast_flag) -> bool:
    """The reference's Cholesky branch: small systems (max_cholesky_size) or the Krylov path switched off (fast_computations)."""
    dense = n <= settings.max_cholesky_size.value() or fast_flag.off()
    if dense and settings.verbose_linalg.on():
        settings.verbose_linalg.logger.debug(f"Running Cholesky on a matrix of size {(n, n)}.")
    return dense


class _Solve(torch.autograd.Function):
    @staticmethod
    def forward(ctx, op, rhs, noise, *hypers):
        vec = rhs.dim() == 1
        r2 = (rhs.unsqueeze(-1) if vec else rhs).detach().float().contiguous()
        n = op.shape[0]
        ctx.dense = _dense_branch(n, settings.fast_computations.solves)
        if ctx.dense:
            chol = op._dense_cholesky()
            sol = torch.cholesky_solve(r2, chol)
        else:
            w, _, _ = op._preconditioner()
            sol, _, _ = _run_cg(op, r2, 0, w)
        ctx.op, ctx.vec = op, vec
        ctx.save_for_backward(sol)
        return sol.squeeze(-1) if vec else sol

    @staticmethod
    def backward(ctx, grad_out):
        op = ctx.op
        (sol,) = ctx.saved_tensors
        g = (grad_out.unsqueeze(-1) if ctx.vec else grad_out).contiguous()
        if ctx.dense:
            gsol = torch.cholesky_solve(g, op._dense_cholesky())
        else:
            w, _, _ = op._preconditioner()
            gsol, _, _ = _run_cg(op, g, 0, w)
        grad_rhs = (gsol.squeeze(-1) if ctx.vec else gsol) if ctx.needs_input_grad[1] else None
        gn = None
        grads = [None] * (len(ctx.needs_input_grad) - 3)
        if any(ctx.needs_input_grad[2:]):
            if any(ctx.needs_input_grad[3:]):
                gs = op.kernel_op._bilinear_derivative_list(-gsol, sol)
                grads = [gi if need else None for gi, need in zip(gs, ctx.needs_input_grad[3:])]
            if ctx.needs_input_grad[2]:
                gn = (-(gsol * sol).sum(-1)) if op.per_row else (-(gsol * sol).sum()).reshape(op.diag.diag_value.shape)
        return (None, grad_rhs, gn, *grads)


class _InvQuadLogdet(torch.autograd.Function):
    """linear_operator.functions._inv_quad_logdet.InvQuadLogdet."""

    @staticmethod
    def forward(ctx, op, rhs, want_logdet, noise, *hypers):
        n = op.shape[0]
        dev = op.device
        ctx.op, ctx.want_logdet, ctx.has_rhs = op, want_logdet, rhs is not None
        nr = 0 if rhs is None else rhs.size(1)
        r = None if rhs is None else rhs.detach().float().contiguous()
        if _dense_branch(n, settings.fast_computations.log_prob):  # the reference's dense branch (not the accelerated path)
            chol = op._dense_cholesky()
            sol = torch.cholesky_solve(r, chol) if r is not None else None
            iq = (sol * r).sum(-2) if r is not None else torch.zeros(0, device=dev)
            ld = 2 * chol.diagonal().log().sum() if want_logdet else torch.zeros((), device=dev)
            ctx.mode = "chol"
            ctx.save_for_backward(chol, sol if sol is not None else torch.zeros(0, device=dev))
            return iq, ld
        ctx.mode = "cg"
        w, lt, logdet_p = op._preconditioner()
        tp = settings.num_trace_samples.value() if (want_logdet or op.kernel_op.requires_grad or noise.requires_grad) else 0
        cols = []
        probes = None
        if tp:
            probes = op._probes(lt, tp)
            norms = probes.norm(2, dim=-2, keepdim=True)
            cols.append(probes / norms)
        if r is not None:
            cols.append(r)
        full = torch.cat(cols, -1)
        solves, tmat, iters = _run_cg(op, full, tp, w)
        ld = torch.zeros((), device=dev)
        if want_logdet and settings.skip_logdet_forward.off():
            if torch.isnan(tmat).any():
                ld = torch.tensor(float("nan"), device=dev)
            else:
                ld = torch.tensor(op._plan().slq_logdet(tmat, n) + logdet_p, device=dev, dtype=torch.float32)
        iq = (solves[:, tp:] * r).sum(-2) if r is not None else torch.zeros(0, device=dev)
        ctx.tp, ctx.iters = tp, iters
        ctx.save_for_backward(solves, probes if probes is not None else torch.zeros(0, device=dev),
                              w if w is not None else torch.zeros(0, device=dev))
        op.last_cg_iters = iters
        return iq, ld

    @staticmethod
    def backward(ctx, grad_iq, grad_ld):
        op = ctx.op
        gn = grad_rhs = None
        grads = [None] * (len(ctx.needs_input_grad) - 4)
        need_k = any(ctx.needs_input_grad[3:])
        if ctx.mode == "chol":
            chol, sol = ctx.saved_tensors
            n = op.shape[0]
            # dense: d iq = -sol sol^T : dK ; d logdet = K^-1 : dK
            left_cols, right_cols = [], []
            if ctx.has_rhs:
                left_cols.append(-sol * grad_iq.reshape(1, -1)); right_cols.append(sol)
                if ctx.needs_input_grad[1]:
                    grad_rhs = 2 * sol * grad_iq.reshape(1, -1)
            if need_k:
                if ctx.want_logdet:
                    kinv = torch.cholesky_inverse(chol)
                    left_cols.append(kinv * grad_ld); right_cols.append(torch.eye(n, device=op.device))
                left = torch.cat(left_cols, -1).contiguous(); right = torch.cat(right_cols, -1).contiguous()
        else:
            solves, probes, w = ctx.saved_tensors
            tp = ctx.tp
            left_cols, right_cols = [], []
            if ctx.want_logdet and tp and need_k:
                coef = 1.0 / tp
                norms = probes.norm(2, dim=-2, keepdim=True)
                pv_solves = solves[:, :tp] * coef * norms * grad_ld   # (1/tp) K^-1 z_i
                pz = probes
                if w.numel():
                    if op.per_row:   # P^-1 z = z / d - W (W^T z), W pre-scaled by D^-1
                        pz = probes / op.diag.diag_vec.detach().unsqueeze(-1) - w @ (w.t() @ probes)
                    else:
                        pz = (probes - w @ (w.t() @ probes)) / op.noise.detach()  # P^-1 z_i
                left_cols.append(pv_solves); right_cols.append(pz)
            if ctx.has_rhs:
                iq_solves = solves[:, tp:]
                neg = -iq_solves * grad_iq.reshape(1, -1)
                left_cols.append(neg); right_cols.append(iq_solves)
                if ctx.needs_input_grad[1]:
                    grad_rhs = -2 * neg
            if need_k and left_cols:
                left = torch.cat(left_cols, -1).contiguous(); right = torch.cat(right_cols, -1).contiguous()
        if need_k and left_cols:
            if any(ctx.needs_input_grad[4:]):
                gs = op.kernel_op._bilinear_derivative_list(left, right)
                grads = [gi if need else None for gi, need in zip(gs, ctx.needs_input_grad[4:])]
            if ctx.needs_input_grad[3]:
                gn = (left * right).sum(-1) if op.per_row else (left * right).sum().reshape(op.diag.diag_value.shape)
        return (None, grad_rhs, None, gn, *grads)
