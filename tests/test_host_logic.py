"""Host-side mirror of the reference interface: settings, constraints, kernel modules, error behaviour (CPU only)."""
import math

import pytest
import torch

import gpytorch_b200 as gp
from gpytorch_b200 import settings
from gpytorch_b200.constraints import GreaterThan, Positive, inv_softplus
from gpytorch_b200.distributed import padded_size, shard_rows
from gpytorch_b200.kernels import MaternKernel, RBFKernel, ScaleKernel


def test_settings_defaults_match_reference():
    # SURVEY.md Appendix A.1 (gpytorch/settings.py:6-31, :173-180)
    assert settings.cg_tolerance.value() == 1.0
    assert settings.eval_cg_tolerance.value() == 0.01
    assert settings.max_cg_iterations.value() == 1000
    assert settings.max_cholesky_size.value() == 800
    assert settings.max_lanczos_quadrature_iterations.value() == 20
    assert settings.max_preconditioner_size.value() == 15
    assert settings.min_preconditioning_size.value() == 2000
    assert settings.num_trace_samples.value() == 10
    assert settings.preconditioner_tolerance.value() == 1e-3
    assert settings.max_root_decomposition_size.value() == 100


def test_settings_context_managers_nest_and_restore():
    with settings.cg_tolerance(1e-4):
        assert settings.cg_tolerance.value() == 1e-4
        with settings.cg_tolerance(0.5), settings.max_preconditioner_size(100):
            assert settings.cg_tolerance.value() == 0.5
            assert settings.max_preconditioner_size.value() == 100
        assert settings.cg_tolerance.value() == 1e-4
    assert settings.cg_tolerance.value() == 1.0
    assert settings.skip_logdet_forward.off()
    with settings.skip_logdet_forward(True):
        assert settings.skip_logdet_forward.on()
    assert settings.skip_logdet_forward.off()


def test_settings_remaining_reference_knobs():
    """The knobs the reference re-exports from linear_operator beyond the solver sizes (gpytorch/settings.py:6-31)."""
    assert settings.terminate_cg_by_size.off() and settings.verbose_linalg.off() and settings.deterministic_probes.off()
    assert settings.tridiagonal_jitter.value() == 1e-6
    fc = settings.fast_computations
    assert fc.covar_root_decomposition.on() and fc.log_prob.on() and fc.solves.on()
    with fc(log_prob=False):
        assert fc.log_prob.off() and fc.solves.on() and fc.covar_root_decomposition.on()
        with fc(solves=False, log_prob=True):
            assert fc.log_prob.on() and fc.solves.off()
        assert fc.log_prob.off() and fc.solves.on()
    assert fc.log_prob.on()
    with settings.deterministic_probes(True):
        seed = settings.deterministic_probes.seed
        assert settings.deterministic_probes.on() and isinstance(seed, int)
        with settings.deterministic_probes(True):
            assert settings.deterministic_probes.seed == seed          # one set of probes for the whole region
    assert settings.deterministic_probes.off() and settings.deterministic_probes.seed is None
    # hand-over to worker threads: the captured overrides win inside restore(), the thread's own state comes back afterwards
    with settings.cg_tolerance(0.25), settings.fast_pred_var(True):
        snap = settings.snapshot()
    assert settings.cg_tolerance.value() == 1.0 and settings.fast_pred_var.off()
    with settings.restore(snap):
        assert settings.cg_tolerance.value() == 0.25 and settings.fast_pred_var.on()
    assert settings.cg_tolerance.value() == 1.0 and settings.fast_pred_var.off()
    assert settings.verbose_linalg.logger.name == "LinAlg (Verbose)"


def test_constraints_roundtrip():
    v = torch.tensor([0.3, 1.0, 7.5])
    assert torch.allclose(Positive().transform(Positive().inverse_transform(v)), v, atol=1e-6)
    g = GreaterThan(1e-4)  # noise floor, likelihoods/noise_models.py:29-30
    assert torch.all(g.transform(torch.tensor([-50.0, 0.0, 3.0])) >= 0.9999e-4)
    assert torch.allclose(inv_softplus(torch.nn.functional.softplus(v)), v, atol=1e-6)


def test_kernel_hyperparameters_and_initialize():
    k = RBFKernel().initialize(lengthscale=2.0)
    assert k.lengthscale.shape == (1, 1) and k.lengthscale.item() == pytest.approx(2.0, rel=1e-6)
    k.lengthscale = 0.5
    assert k.lengthscale.item() == pytest.approx(0.5, rel=1e-6)
    ka = RBFKernel(ard_num_dims=3)
    ka.lengthscale = torch.tensor([1.0, 2.0, 3.0])
    assert torch.allclose(ka.lengthscale, torch.tensor([[1.0, 2.0, 3.0]]), atol=1e-6)
    s = ScaleKernel(MaternKernel(nu=1.5)).initialize(outputscale=3.0)
    assert s.outputscale.item() == pytest.approx(3.0, rel=1e-6)
    assert s.base_kernel.kind == "matern32"
    assert len(list(s.parameters())) == 2


def test_matern_rejects_bad_nu():
    with pytest.raises(RuntimeError, match="nu expected to be 0.5, 1.5, or 2.5"):  # matern_kernel.py:80-81
        MaternKernel(nu=2.0)


def test_kernel_call_contract_errors():
    k = RBFKernel(ard_num_dims=3)
    with pytest.raises(RuntimeError, match="ard_num_dims"):
        k(torch.rand(5, 2))
    k2 = RBFKernel()
    with pytest.raises(RuntimeError, match="same number of dimensions"):  # kernels/kernel.py:506-507
        k2(torch.rand(5, 2), torch.rand(4, 3))


def test_kernel_forward_returns_lazy_operator_without_compute():
    x = torch.rand(50, 4)
    op = ScaleKernel(RBFKernel())(x)
    assert tuple(op.shape) == (50, 50) and op.same and op.kind == "rbf"
    assert op.evaluate_kernel() is op
    op2 = RBFKernel()(x, torch.rand(7, 4))
    assert tuple(op2.shape) == (50, 7) and not op2.same
    sub = op[3:10, :]
    assert tuple(sub.shape) == (7, 50)


def test_ops_fail_loudly_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    op = RBFKernel()(torch.rand(20, 2))
    with pytest.raises(RuntimeError, match="CUDA"):
        op.matmul(torch.rand(20, 1))
    with pytest.raises(RuntimeError, match="CUDA"):
        gp.Plan(torch.rand(5, 2))


def test_mll_module_type_checks():
    lik = gp.likelihoods.GaussianLikelihood()
    assert lik.noise.item() > 1e-4
    lik.noise = 0.1
    assert lik.noise.item() == pytest.approx(0.1, rel=1e-5)
    with pytest.raises(RuntimeError, match="Gaussian"):
        gp.mlls.ExactMarginalLogLikelihood(object(), None)


def test_dense_log_prob_known_answer():
    # /root/reference/test/distributions/test_multivariate_normal.py:23-43
    mvn = gp.distributions.MultivariateNormal(torch.tensor([0.0, 1, 2]), torch.diag(torch.tensor([1.0, 0.75, 1.5])))
    assert mvn.log_prob(torch.zeros(3)).item() == pytest.approx(-4.8157, abs=1e-4)


def test_exact_gp_train_mode_guard():
    class M(gp.models.ExactGP):
        def __init__(self, x, y, lik):
            super().__init__(x, y, lik)
            self.mean_module = gp.means.ConstantMean()
            self.covar_module = ScaleKernel(RBFKernel())

        def forward(self, x):
            return gp.distributions.MultivariateNormal(self.mean_module(x), self.covar_module(x))

    x, y = torch.rand(10, 2), torch.rand(10)
    m = M(x, y, gp.likelihoods.GaussianLikelihood())
    m.train()
    out = m(x)
    assert tuple(out.lazy_covariance_matrix.shape) == (10, 10)
    with pytest.raises(RuntimeError, match="You must train on the training inputs!"):  # exact_gp.py:276-280
        m(torch.rand(10, 2))


def test_shard_rows_cover_exactly():
    for n, w in ((50000, 8), (200000, 8), (1004, 4), (64, 8)):
        tot, prev_end = 0, 0
        for r in range(w):
            b, c, per = shard_rows(n, w, r)
            assert b == prev_end and c == per == n // w
            prev_end = b + c
            tot += c
        assert tot == n
    # unequal shards would hand ncclAllGather different counts per rank (hang / corruption): refused up front
    for n, w in ((1001, 4), (7, 8)):
        assert padded_size(n, w) % w == 0 and padded_size(n, w) >= n
        with pytest.raises(ValueError, match="not divisible"):
            shard_rows(n, w, 0)
        shard_rows(padded_size(n, w), w, w - 1)


def test_bench_generators_match_the_oracle_generators_bitwise():
    """bench.py's product arm imports nothing from oracle/; its input generators must still be the oracle's."""
    import importlib.util
    import os

    import torch
    from oracle import mll as om

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    x, y = bench.synthetic_problem(777, 5, 3)
    xo, yo = om.synthetic_problem(777, 5, 3, torch.float32)
    assert torch.equal(x, xo) and torch.equal(y, yo)
    for a, b in zip(bench.make_probe_noise(500, 30, 10, 1), om.make_probe_noise(500, 30, 10, 1)):
        assert torch.equal(a, b)
    src = open(bench.__file__).read()
    ours = src[src.index("def measure_workload"):src.index("def main()")]
    assert "oracle" not in ours.replace("oracle port", "").replace("(oracle", "")   # the GPU arm never touches oracle/


def test_state_dict_layout_matches_the_reference():
    """Parameter / buffer names of a reference checkpoint (means/constant_mean.py: raw_constant; likelihoods/noise_models.py:
    noise_covar.raw_noise; constraints/constraints.py:44-45: *_constraint.lower_bound / upper_bound)."""
    import torch
    import gpytorch_b200 as gp

    class M(gp.models.ExactGP):
        def __init__(self, x, y, lik):
            super().__init__(x, y, lik)
            self.mean_module = gp.means.ConstantMean()
            self.covar_module = gp.kernels.ScaleKernel(gp.kernels.RBFKernel(ard_num_dims=2))

        def forward(self, x):
            return gp.distributions.MultivariateNormal(self.mean_module(x), self.covar_module(x))

    lik = gp.likelihoods.GaussianLikelihood()
    m = M(torch.rand(10, 2), torch.rand(10), lik)
    keys = set(m.state_dict().keys())
    want = {
        "likelihood.noise_covar.raw_noise", "likelihood.noise_covar.raw_noise_constraint.lower_bound",
        "likelihood.noise_covar.raw_noise_constraint.upper_bound", "mean_module.raw_constant",
        "covar_module.raw_outputscale", "covar_module.base_kernel.raw_lengthscale",
        "covar_module.base_kernel.raw_lengthscale_constraint.lower_bound",
        "covar_module.base_kernel.raw_lengthscale_constraint.upper_bound",
        "covar_module.raw_outputscale_constraint.lower_bound", "covar_module.raw_outputscale_constraint.upper_bound",
    }
    assert want <= keys, want - keys
    assert tuple(m.state_dict()["covar_module.base_kernel.raw_lengthscale"].shape) == (1, 2)     # kernel.py:213-219
    assert tuple(m.state_dict()["likelihood.noise_covar.raw_noise"].shape) == (1,)
    assert tuple(m.state_dict()["mean_module.raw_constant"].shape) == ()
    # a reference-shaped checkpoint loads strictly and lands in the right parameters
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    sd["mean_module.raw_constant"] = torch.tensor(0.7)
    sd["likelihood.noise_covar.raw_noise"] = torch.tensor([1.5])
    m2 = M(torch.rand(10, 2), torch.rand(10), gp.likelihoods.GaussianLikelihood())
    m2.load_state_dict(sd, strict=True)
    assert float(m2.mean_module.constant) == pytest.approx(0.7)
    assert float(m2.likelihood.noise) == pytest.approx(float(torch.nn.functional.softplus(torch.tensor(1.5)) + 1e-4))
    # old layouts are renamed on load (constant_mean.py:18-31 does the same for `constant`)
    sd_old = {k: v for k, v in sd.items() if k not in ("mean_module.raw_constant", "likelihood.noise_covar.raw_noise")}
    sd_old["mean_module.constant"] = torch.tensor([0.25])
    sd_old["likelihood.raw_noise"] = torch.tensor([0.5])
    m2.load_state_dict(sd_old, strict=True)
    assert float(m2.mean_module.constant) == pytest.approx(0.25)


def test_oracle_love_root_reproduces_inverse_on_small_system():
    """oracle.linalg.root_inv_decomposition: with J = N Lanczos steps R R^T is the exact inverse, so the LOVE covariance
    equals the exact predictive covariance (exact_prediction_strategies.py:268-272, 464-478)."""
    import torch
    from oracle import kernels as ok, linalg as ol

    torch.manual_seed(0)
    n, m = 60, 7
    x, xs = torch.rand(n, 2, dtype=torch.float64), torch.rand(m, 2, dtype=torch.float64)
    K = ok.kernel_matrix("rbf", x, x, 0.5, 1.0, True) + 0.3 * torch.eye(n, dtype=torch.float64)
    ksx = ok.kernel_matrix("rbf", xs, x, 0.5, 1.0, False)
    kss = ok.kernel_matrix("rbf", xs, xs, 0.5, 1.0, True)
    r = ol.root_inv_decomposition(lambda v: K @ v, n, torch.randn(n, dtype=torch.float64))
    exact = kss - ksx @ torch.linalg.solve(K, ksx.T)
    love = ol.love_predictive_covar(kss, ksx, r)
    assert (love - exact).abs().max().item() < 1e-6
    # and a truncated decomposition is a PSD under-estimate of the correction: predictive variances only grow
    r20 = ol.root_inv_decomposition(lambda v: K @ v, 20, torch.randn(n, dtype=torch.float64))
    love20 = ol.love_predictive_covar(kss, ksx, r20)
    assert torch.all(love20.diagonal() >= exact.diagonal() - 1e-9)


def test_operator_protocol_without_compute():
    """Shape / transpose / lazy-slicing protocol of the covariance operators needs no device work
    (lazy_evaluated_kernel_tensor.py:136-243, 277-341): it must behave on CPU tensors, and only products may fail."""
    import torch
    import gpytorch_b200 as gp
    from gpytorch_b200.operators import AddedDiagLinearOperator, ConstantDiagLinearOperator, KernelLinearOperator

    x1, x2 = torch.rand(30, 4), torch.rand(12, 4)
    ls = torch.tensor(0.5)
    sq = KernelLinearOperator(x1, None, "rbf", ls)
    cr = KernelLinearOperator(x1, x2, "matern52", ls, torch.tensor(2.0))
    assert sq.shape == sq._size() == sq.matrix_shape == torch.Size([30, 30]) and sq.dim() == 2 and sq.numel() == 900
    assert cr.shape == torch.Size([30, 12]) and cr.t().shape == cr.mT.shape == cr.transpose(-2, -1).shape == torch.Size([12, 30])
    assert cr.transpose(-1, -1) is cr and sq.t() is sq and cr.t().t().shape == cr.shape
    assert cr.t().x1 is x2 and cr.t().x2 is x1 and cr.t().kind == "matern52"
    sub = sq[3:9, 10:20]
    assert isinstance(sub, KernelLinearOperator) and sub.shape == torch.Size([6, 10]) and not sub.same
    assert sq._getitem(slice(0, 5), slice(None)).shape == torch.Size([5, 30])
    assert sq.batch_shape == torch.Size([]) and sq.dtype == torch.float32 and not sq.requires_grad
    assert KernelLinearOperator(x1, None, "rbf", ls.clone().requires_grad_(True)).requires_grad
    assert sq.detach().shape == sq.shape and len(sq.representation()) == 4 and sq.evaluate_kernel() is sq
    khat = sq + ConstantDiagLinearOperator(torch.tensor(0.1), 30)
    assert isinstance(khat, AddedDiagLinearOperator) and khat.shape == sq.shape and khat.t() is khat and khat.mT is khat
    assert float(khat.add_jitter(0.2).noise) == pytest.approx(0.3)
    assert float((khat + ConstantDiagLinearOperator(torch.tensor(0.4), 30)).noise) == pytest.approx(0.5)
    assert len(khat.representation()) == 5 and not khat.requires_grad and khat.dim() == 2
    with pytest.raises(RuntimeError, match="square"):
        AddedDiagLinearOperator(cr, ConstantDiagLinearOperator(torch.tensor(0.1), 30))
    with pytest.raises(NotImplementedError):
        sq + 1.0
    with pytest.raises(RuntimeError):      # products need the CUDA engine: no CPU fallback
        sq.matmul(torch.rand(30, 2))
    assert gp.settings.fast_pred_var.off() and gp.settings.skip_posterior_variances.off()
    with gp.settings.fast_pred_var(True), gp.settings.skip_posterior_variances(True):
        assert gp.settings.fast_pred_var.on() and gp.settings.skip_posterior_variances.on()
    assert gp.settings.fast_pred_var.off()


def test_additive_kernel_protocol_without_compute():
    """k1 + k2 -> AdditiveKernel with nested sums flattened (kernels/kernel.py:541-545, :592-621); calling it gives ONE lazy sum
    operator whose terms keep their own active dimensions; slicing / transposing / adding the noise stay lazy."""
    import gpytorch_b200 as gp
    from gpytorch_b200.operators import AddedDiagLinearOperator, ConstantDiagLinearOperator, SumKernelLinearOperator

    k = gp.kernels.ScaleKernel(gp.kernels.RBFKernel(active_dims=[0, 2])) + gp.kernels.ScaleKernel(gp.kernels.MaternKernel(nu=1.5))
    assert isinstance(k, gp.kernels.AdditiveKernel) and len(k.kernels) == 2
    assert len((k + gp.kernels.RBFKernel()).kernels) == 3 and len((gp.kernels.RBFKernel() + k).kernels) == 3
    names = sorted(n for n, _ in k.named_parameters())
    assert names == ["kernels.0.base_kernel.raw_lengthscale", "kernels.0.raw_outputscale",
                     "kernels.1.base_kernel.raw_lengthscale", "kernels.1.raw_outputscale"]
    x = torch.rand(40, 3)
    op = k(x)
    assert isinstance(op, SumKernelLinearOperator) and [o.kind for o in op.ops] == ["rbf", "matern32"]
    assert op.ops[0].x1.shape == (40, 2) and op.ops[1].x1.shape == (40, 3) and op.shape == torch.Size([40, 40]) and op.same
    assert len(op.hyper_tensors()) == 4 and op.requires_grad and not op.detach().requires_grad
    sub = op[5:, :7]
    assert isinstance(sub, SumKernelLinearOperator) and sub.shape == torch.Size([35, 7]) and sub.t().shape == torch.Size([7, 35])
    khat = op + ConstantDiagLinearOperator(torch.tensor(0.1), 40)
    assert isinstance(khat, AddedDiagLinearOperator) and khat.kernel_op is op
    # operator-level addition gives the same lazy sum, nested sums are flattened, five terms are refused
    a = op.ops[0]
    assert len((a + a).ops) == 2 and len(((a + a) + (a + a)).ops) == 4
    with pytest.raises(RuntimeError, match="1 to 4 terms"):
        (a + a) + (a + a) + a
    with pytest.raises(RuntimeError, match="cannot add kernels of shapes"):
        a + a[:10]
    with pytest.raises(RuntimeError, match="must be kernels"):
        gp.kernels.AdditiveKernel(gp.kernels.RBFKernel(), 3.0)


def test_ski_operator_slicing_protocol_without_compute():
    """Rows / columns of the interpolated operator stay lazy (prediction slices the joint train + test operator):
    shapes, transposes and the size check need no device work; products need the CUDA engine."""
    import gpytorch_b200 as gp
    from gpytorch_b200.operators import SKIKernelLinearOperator

    x = torch.rand(50, 2)
    k = gp.kernels.ScaleKernel(gp.kernels.GridInterpolationKernel(gp.kernels.RBFKernel(), grid_size=12, num_dims=2, grid_bounds=[(0.0, 1.0)] * 2))
    op = k(x)
    assert isinstance(op, SKIKernelLinearOperator) and op.shape == torch.Size([50, 50])
    ks = op[40:, :40]
    assert ks.shape == torch.Size([10, 40]) and ks.t().shape == torch.Size([40, 10]) and ks.transpose(-1, -2).shape == torch.Size([40, 10])
    assert ks.transpose(-1, -1) is ks and ks.evaluate_kernel() is ks and ks.device == x.device
    assert op[torch.tensor([1, 3, 5]), :].shape == torch.Size([3, 50])
    with pytest.raises(RuntimeError, match="cannot be multiplied"):
        ks.matmul(torch.rand(7, 2))
    with pytest.raises(RuntimeError):          # the product itself: CUDA only
        ks.matmul(torch.rand(40, 2))
    with pytest.raises(NotImplementedError):
        op + op                                  # interpolated operators are not summed on the accelerated path
