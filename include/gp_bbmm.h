/*
 * gp_bbmm.h -- C ABI of the B200-native BBMM exact-GP engine (libgpbbmm.so).
 *
 * Drop-in boundary for ONE hot path of cornellius-gp/gpytorch: the mBCG / SLQ evaluation of
 * the exact-GP marginal log likelihood.  Plain pointers and sizes only; no torch types.
 * Every entry point names the reference interface it replaces (paths relative to
 * /root/reference/gpytorch unless noted; "linear_operator" = the pinned third-party
 * dependency linear_operator>=0.6.1, setup.py:44, whose source is not vendored).
 *
 * Conventions
 *  - all device buffers are caller-owned (torch caching allocator in the Python host),
 *    row-major fp32 unless stated, and must stay alive until the stream work finishes;
 *  - every call enqueues on the cudaStream_t passed at plan creation (as void*) and
 *    returns an int status (GP_OK == 0); nothing throws, nothing calls exit();
 *  - hyper-parameters arrive as already-constrained host scalars (module.py /
 *    constraints/constraints.py stay in PyTorch);
 *  - there is NO CPU fallback: every function needs a CUDA device.
 */
#ifndef GP_BBMM_H
#define GP_BBMM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes (SURVEY.md section 8b "Errors"): mapped by the binding to RuntimeError /
 * NumericalWarning (utils/warnings.py:5) / NanError (utils/errors.py:8-22). */
enum {
  GP_OK = 0,
  GP_E_SHAPE = 1,         /* bad sizes / unsupported configuration            */
  GP_E_CUDA = 2,          /* a CUDA runtime call failed (gp_last_error())     */
  GP_E_NAN_MVM = 3,       /* "NaNs encountered when trying to perform matrix-vector multiplication" */
  GP_W_NOT_CONVERGED = 4, /* CG hit max_iter above tolerance (NumericalWarning) */
  GP_W_PIVCHOL_NAN = 5,   /* NaN in pivoted Cholesky -> preconditioner dropped  */
  GP_E_NCCL = 6,
  GP_E_STATE = 7,         /* call order violated (data / hypers not set)      */
  GP_W_EIG_NOT_CONVERGED = 8 /* tridiagonal QL iteration hit its sweep limit: the SLQ log-det is unreliable (NumericalWarning) */
};

/* covariance function kinds: kernels/rbf_kernel.py:68-85, kernels/matern_kernel.py:85-110 */
enum { GP_RBF = 0, GP_MATERN12 = 1, GP_MATERN32 = 2, GP_MATERN52 = 3 };

/* which fused K.V kernel runs: GP_BACKEND_TCGEN05 = tcgen05/TMEM/bulk-TMA 3xTF32 kernel,
 * GP_BACKEND_SIMT = fp32 CUDA-core kernel (bring-up / cross-check / d > 41). */
enum { GP_BACKEND_AUTO = 0, GP_BACKEND_TCGEN05 = 1, GP_BACKEND_SIMT = 2, GP_BACKEND_SKI = 3 /* set by gp_plan_set_ski */,
       GP_BACKEND_SUM = 4 /* set by gp_plan_set_sum */ };

typedef struct gp_plan gp_plan;   /* opaque: repacked X, workspaces, stream, comm */
typedef struct gp_comm gp_comm;   /* opaque: NCCL communicator for row-sharded runs */

const char* gp_version(void);
const char* gp_last_error(void);          /* thread-local text of the last failure */
const char* gp_status_string(int status);

/* ---- plan ------------------------------------------------------------------------- */

/* Create a plan on `device` working on `stream` (a cudaStream_t, NULL = legacy default).
 * Replaces nothing 1:1; it is the "optional opaque handle" of SURVEY.md section 8b Ownership. */
int gp_plan_create(gp_plan** out, int device, void* stream);
int gp_plan_destroy(gp_plan* plan);
int gp_plan_set_backend(gp_plan* plan, int backend);

/* Training / test inputs.  X1 [n1, d] (ld1 floats per row), X2 [n2, d] or NULL (X2 == X1:
 * the x1_eq_x2 branch of sq_dist, kernels/kernel.py:26-49, incl. its exact diagonal).
 * Row-sharded runs pass row_begin/row_count: this rank owns output rows
 * [row_begin, row_begin+row_count) of K (multi_device_kernel.py:38,56-62); row_count<=0 = all. */
int gp_plan_set_data(gp_plan* plan, const float* X1, int64_t n1, int64_t ld1,
                     const float* X2, int64_t n2, int64_t ld2, int d,
                     int64_t row_begin, int64_t row_count);

/* Hyper-parameters: kind, lengthscale (host array, n_ls == 1 or d [ARD]), outputscale
 * (scale_kernel.py:108-118), noise sigma^2 (noise_models.py:57-92; added iff X2 == X1 and
 * the call asks for it).  Re-packs X on the device; call again whenever they change. */
int gp_plan_set_hypers(gp_plan* plan, int kind, const float* lengthscale, int n_ls,
                       float outputscale, float noise);

/* Per-row noise diagonal D (FixedNoiseGaussianLikelihood, likelihoods/gaussian_likelihood.py:245-363; FixedGaussianNoise,
 * noise_models.py:150-190): K_hat = K + diag(d).  `diag` is a device pointer to n2 floats (caller-owned, must outlive the plan's
 * use of it) that replaces the scalar noise in every product / solve / preconditioner / probe; NULL restores the scalar. */
int gp_plan_set_noise_diag(gp_plan* plan, const float* diag, int64_t n);

/* SKI / KISS-GP (kernels/grid_interpolation_kernel.py:132-213, kernels/grid_kernel.py:107-177, utils/interpolation.py:15-167):
 * the plan's operator becomes K_ski = W (T_0 x ... x T_{d-1}) W^T with cubic interpolation onto a regular grid; grid_lo[i] /
 * grid_step[i] are the first node and the spacing of dimension i (utils/grid.py:142-180), grid_sizes[i] in [4, 128], d <= 4.
 * Call after gp_plan_set_data (square operator); products, mBCG, SLQ, gp_mll (without preconditioner) and gp_lanczos then run
 * on the interpolated operator.  Out-of-bounds inputs fail like the reference ("Received data that was out of bounds ..."). */
int gp_plan_set_ski(gp_plan* plan, const int* grid_sizes, const float* grid_lo, const float* grid_step, int d);

/* Kernel sums (AdditiveKernel, kernels/kernel.py:592-621: k = k_1 + ... + k_m, each term with its own covariance function,
 * lengthscale(s), outputscale and active dimensions): `plan` becomes the operator  sum_t K_t  (+ its own noise), where every
 * K_t is a ready plan over the same rows (same n1 / n2 / row shard / stream; the data pointers may differ: active_dims).
 * Call after gp_plan_set_data on `plan`; gp_plan_set_hypers on `plan` (before or after) supplies the sum's noise, its kind /
 * lengthscale / outputscale are ignored.  The terms stay owned by the caller and must outlive `plan`; re-packing a term (gp_plan_set_hypers on it) is
 * picked up by the next call on `plan`.  One K.V of the sum = one fused kernel launch per term into disjoint partial slots,
 * reduced by the same finish kernels; gp_pivoted_cholesky evaluates rows of the sum.  n_terms in [1, 4].
 * Hyper-parameter gradients: gp_bilinear_grad on each term plan. */
int gp_plan_set_sum(gp_plan* plan, gp_plan* const* terms, int n_terms);

/* ---- kernel seam (LazyEvaluatedKernelTensor, lazy/lazy_evaluated_kernel_tensor.py) --- */

/* OUT[n1_local, t] = K(X1,X2) V [+ noise * V when add_noise and X2 == X1].
 * Replaces LazyEvaluatedKernelTensor._matmul (:245-276) / KernelLinearOperator._matmul
 * (kernels/keops/rbf_kernel.py:44-55); K is never written to HBM.  V [n2, t] ldv, OUT ldo. */
int gp_kmv(gp_plan* plan, const float* V, int64_t ldv, int t, float* OUT, int64_t ldo, int add_noise);

/* OUT[m, n2] = K(X1[idx], X2): row extraction, LazyEvaluatedKernelTensor._getitem (:136-243);
 * idx is a DEVICE int64 array. */
int gp_krows(gp_plan* plan, const int64_t* idx, int64_t m, float* OUT, int64_t ldo);

/* OUT[n1] = diag K(X1,X1): LazyEvaluatedKernelTensor._diagonal (:107-133). */
int gp_kdiag(gp_plan* plan, float* OUT);

/* d/d(theta) sum_ij sum_s Lf[i,s] K_theta(x_i,x_j) Rt[j,s]: LazyEvaluatedKernelTensor.
 * _bilinear_derivative (:69-105) + RBFCovariance/MaternCovariance.backward
 * (functions/rbf_covariance.py:26-29, matern_covariance.py:52-56).
 * grad_ls: host double[n_ls]; grad_os: host double (d/d outputscale).  Lf [n1,s], Rt [n2,s]. */
int gp_bilinear_grad(gp_plan* plan, const float* Lf, int64_t ldl, const float* Rt, int64_t ldr,
                     int s, double* grad_ls, double* grad_os);

/* ---- solver seam (linear_operator) --------------------------------------------------- */

/* Greedy pivoted partial Cholesky of K(X1,X1) (outputscale included, no noise):
 * linear_operator.functions._pivoted_cholesky, surfaced at gpytorch/__init__.py:146-173.
 * Lt [rank, n] row-major (= L^T), piv int64[rank] (device), *rank_out <= rank. */
int gp_pivoted_cholesky(gp_plan* plan, int rank, float error_tol, float* Lt, int64_t* piv,
                        int* rank_out);

/* Preconditioner for K + noise I from Lt [k, n]: W [n, k] with P^{-1} v = (v - W W^T v)/noise,
 * log det P.  AddedDiagLinearOperator._preconditioner / _init_cache_for_constant_diag
 * (linear_operator); W spans the same space as the reference's Q[:n] (W W^T == Q Q^T). */
int gp_precond_build(gp_plan* plan, const float* Lt, int k, float* W, double* logdet_out);

/* z = L eps1 + sqrt(noise) eps2 ~ N(0, P): the probe draw of InvQuadLogdet.forward
 * (linear_operator) with the base samples supplied.  eps1 [k, tp], eps2 [n, tp], Z [n, tp]. */
int gp_precond_probes(gp_plan* plan, const float* Lt, int k, const float* eps1, const float* eps2,
                      int tp, float* Z);

/* modified batched preconditioned CG on (K + noise I) with K applied by the fused kernel:
 * linear_operator.utils.linear_cg (signature attested at
 * variational/ciq_variational_strategy.py:56-64).  RHS/SOLVES [n, t] (t <= 16 per call),
 * W [n,k] or NULL.  TMAT fp32 [n_tridiag, max_tridiag_iter, max_tridiag_iter] (zero-filled by
 * the call; the leading J x J block is valid, J -> *tridiag_size).  resid_out host float[t]. */
int gp_mbcg(gp_plan* plan, const float* RHS, int64_t ldr, int t, int n_tridiag, float tolerance,
            int max_iter, int max_tridiag_iter, const float* W, int k, float* SOLVES, int64_t lds,
            float* TMAT, int* iters_out, int* tridiag_size, float* resid_out);

/* log det estimate from the mBCG tridiagonals: lanczos_tridiag_to_diag + StochasticLQ.to_dense
 * (linear_operator.utils.lanczos / stochastic_lq).  TMAT [n_tridiag, ldt, ldt] device fp32,
 * leading J x J blocks used; result (n / n_tridiag) sum_i sum_j (V_i[0,j])^2 log lambda_ij. */
int gp_slq_logdet(gp_plan* plan, const float* TMAT, int n_tridiag, int ldt, int J, int64_t n,
                  double* logdet_out);

/* Lanczos tridiagonalisation with full re-orthogonalisation of (K + noise I):
 * linear_operator.utils.lanczos.lanczos_tridiag (root_inv_decomposition,
 * models/exact_prediction_strategies.py:268-272).  INIT [n], Q [max_iter, n] row-major
 * (= Q^T), T [max_iter, max_iter]; *J_out = steps run. */
int gp_lanczos(gp_plan* plan, const float* INIT, int max_iter, float tol, float* Qt, float* T,
               int* J_out);

/* one-shot: MultivariateNormal.log_prob (distributions/multivariate_normal.py:221-252) through
 * inv_quad_logdet (:249), i.e. pivoted Cholesky -> preconditioner -> probes -> mBCG -> SLQ. */
typedef struct gp_mll_opts {
  int num_probes;          /* settings.num_trace_samples (10)                    */
  int precond_rank;        /* settings.max_preconditioner_size (15; C2 uses 100) */
  int min_precond_size;    /* settings.min_preconditioning_size (2000)           */
  float precond_tol;       /* settings.preconditioner_tolerance (1e-3)           */
  float cg_tol;            /* settings.cg_tolerance (1.0)                        */
  int max_cg_iter;         /* settings.max_cg_iterations (1000)                  */
  int max_tridiag_iter;    /* settings.max_lanczos_quadrature_iterations (20)    */
} gp_mll_opts;

typedef struct gp_mll_result {
  double inv_quad, logdet, logdet_precond, log_prob, mll;
  int cg_iters, tridiag_size, precond_rank, status_flags;
  float resid[16];
} gp_mll_result;

/* y_minus_mean [n]; eps1 [precond_rank, tp], eps2 [n, tp] N(0,1) base samples, rademacher [n, tp]
 * (used when no preconditioner applies); solve_out [n] = K_hat^{-1}(y - mu) or NULL. */
int gp_mll(gp_plan* plan, const float* y_minus_mean, const float* eps1, const float* eps2,
           const float* rademacher, const gp_mll_opts* opts, float* solve_out, gp_mll_result* res);

/* ---- multi-GPU (one process per GPU; replaces MultiDeviceKernel, multi_device_kernel.py:14-95) */
int gp_comm_unique_id(uint8_t out[128]);                 /* rank 0, then broadcast by the host */
int gp_comm_init(gp_comm** out, const uint8_t id[128], int rank, int world);
int gp_comm_destroy(gp_comm* comm);
int gp_plan_set_comm(gp_plan* plan, gp_comm* comm);      /* NULL = single GPU */

/* ---- introspection for bench.py --------------------------------------------------- */
int64_t gp_kernel_launches(gp_plan* plan);               /* kernels launched by this plan so far */
int gp_plan_info(gp_plan* plan, int* backend, int* nsplit, int* kpad, int* n_sm);
/* Times `reps` back-to-back launches of the fused K.V kernel ALONE (after `warmup` untimed ones) with CUDA
 * events on the plan's stream; V [n2, t].  *ms_per_launch is the average device time of one launch. */
/* Debug: device buffer of 256*8 int64 receiving clock64() stamps of the tcgen05 pipeline events of CTA (0,0)
 * (NULL disables).  Used by tools/tc_trace.py to study pipeline bubbles. */
int gp_plan_set_trace(gp_plan* plan, long long* trace);
int gp_time_kmv_kernel(gp_plan* plan, const float* V, int64_t ldv, int t, int warmup, int reps, float* ms_per_launch);

#ifdef __cplusplus
}
#endif
#endif /* GP_BBMM_H */
