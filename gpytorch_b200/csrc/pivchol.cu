// pivchol.cu -- pivoted-Cholesky preconditioner: greedy partial Cholesky, Woodbury factor, N(0,P) probes.
//
// gp_pivoted_cholesky restates linear_operator.functions._pivoted_cholesky (SURVEY.md Appendix A.3;
// surfaced at /root/reference/gpytorch/__init__.py:146-173): k sequential steps, each picks the largest
// remaining diagonal entry (ties -> earliest position in the running permutation, as torch.max over
// permuted_diags does), evaluates ONE kernel row on the fly, and applies the rank-m update.
// gp_precond_build restates AddedDiagLinearOperator._init_cache_for_constant_diag (Appendix A.4)
// through the k x k Cholesky of (L^T L + sigma^2 I) in fp64 instead of a QR of [L; sigma I]:
// W = L C^{-T} spans the same column space as Q[:n] with W W^T == Q Q^T, and
// log det P = log det(L^T L + sigma^2 I) + (n - k) log sigma^2.
#include "gp_common.cuh"

namespace gp {

struct PcState {
  int done;        // stop rule fired (error <= tol, or NaN) -> later launches are no-ops
  int rank;        // steps completed
  int pivot;       // pivot (global row) selected for the NEXT step to run
  int nan_flag;
  float dpiv;      // sqrt(max diag) = L[m][pivot]
  float orig_err;  // max of the initial diagonal
  float err;
  unsigned int counter;  // last-block-done ticket (step-wise path)
  unsigned int bar;      // monotonic grid-barrier counter (persistent path)
};

constexpr int PC_THREADS = 128;

// One launch per step m, fused: row update with the pivot chosen by the previous launch,
// then -- in the same pass -- the per-CTA (max, earliest permutation position, sum |diag|) over the remaining
// entries; the last CTA to finish reduces the partials, applies the stop rule and selects / swaps the next pivot.
//   L[m][j] = (K[pi, j] - sum_{q<m} L[q][pi] L[q][j]) / L[m][pi] ; diag[j] -= L[m][j]^2
template <int KIND>
__global__ void __launch_bounds__(PC_THREADS)
pc_step_kernel(const float* __restrict__ Z, int DP, float os, float* __restrict__ Lt, int64_t n, int m, int max_rank,
               float tol, float* __restrict__ diag, int* __restrict__ perm, int* __restrict__ pos, PcState* __restrict__ st,
               int64_t* __restrict__ piv_out, float* __restrict__ pval, int* __restrict__ ppos, double* __restrict__ psum) {
  if (st->done) return;
  extern __shared__ float sh[];
  float* zp = sh;           // [DP]
  float* lp = sh + DP;      // [m]   L[q][pivot]
  __shared__ float s_val[PC_THREADS];
  __shared__ int s_pos[PC_THREADS];
  __shared__ double s_sum[PC_THREADS];
  __shared__ bool s_last;
  const int tid = threadIdx.x;
  const int pi = st->pivot;
  const float dpiv = st->dpiv;
  for (int c = tid; c < DP; c += PC_THREADS) zp[c] = Z[(int64_t)pi * DP + c];
  for (int q = tid; q < m; q += PC_THREADS) lp[q] = Lt[(int64_t)q * n + pi];
  __syncthreads();
  const int64_t j = (int64_t)blockIdx.x * PC_THREADS + tid;
  float best = -INFINITY;
  int best_pos = 0x7fffffff;
  double asum = 0.0;
  if (j < n) {
    float* Lm = Lt + (int64_t)m * n;
    const int pj = pos[j];
    if (pj < m) {
      Lm[j] = 0.f;               // earlier pivots stay zero in this row
    } else if (pj == m) {
      Lm[j] = dpiv;              // the pivot itself
    } else {
      float s = 0.f;
      for (int c = 0; c < DP; ++c) {
        float df = zp[c] - Z[j * DP + c];
        s = fmaf(df, df, s);
      }
      float v = os * cov_from_arg<KIND>(-0.5f * s);
      {
        // independent partial sums keep 8 L2 loads in flight per thread (the step is latency bound, not bandwidth bound)
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int q = 0;
        for (; q + 4 <= m; q += 4) {
          s0 = fmaf(lp[q], Lt[(int64_t)q * n + j], s0);
          s1 = fmaf(lp[q + 1], Lt[(int64_t)(q + 1) * n + j], s1);
          s2 = fmaf(lp[q + 2], Lt[(int64_t)(q + 2) * n + j], s2);
          s3 = fmaf(lp[q + 3], Lt[(int64_t)(q + 3) * n + j], s3);
        }
        for (; q < m; ++q) s0 = fmaf(lp[q], Lt[(int64_t)q * n + j], s0);
        v -= (s0 + s1) + (s2 + s3);
      }
      v /= dpiv;
      Lm[j] = v;
      float dn = diag[j] - v * v;
      diag[j] = dn;
      if (dn != dn) { best = INFINITY; best_pos = -1; }  // NaN poisons the selection
      else { best = dn; best_pos = pj; }
      asum = fabs((double)dn);
    }
  }
  s_val[tid] = best; s_pos[tid] = best_pos; s_sum[tid] = asum;
  __syncthreads();
  for (int s = PC_THREADS / 2; s > 0; s >>= 1) {
    if (tid < s) {
      float v2 = s_val[tid + s]; int p2 = s_pos[tid + s];
      if (v2 > s_val[tid] || (v2 == s_val[tid] && p2 < s_pos[tid])) { s_val[tid] = v2; s_pos[tid] = p2; }
      s_sum[tid] += s_sum[tid + s];
    }
    __syncthreads();
  }
  if (tid == 0) {
    pval[blockIdx.x] = s_val[0]; ppos[blockIdx.x] = s_pos[0]; psum[blockIdx.x] = s_sum[0];
    __threadfence();
    unsigned int ticket = atomicAdd(&st->counter, 1u);
    s_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // ---- last CTA: fixed-order reduction of the partials, stop rule, next pivot ----
  best = -INFINITY; best_pos = 0x7fffffff; asum = 0.0;
  for (int b = tid; b < (int)gridDim.x; b += PC_THREADS) {   // fixed assignment + fixed tree => deterministic
    float v2 = ((volatile float*)pval)[b]; int p2 = ((volatile int*)ppos)[b];
    if (v2 > best || (v2 == best && p2 < best_pos)) { best = v2; best_pos = p2; }
    asum += ((volatile double*)psum)[b];
  }
  s_val[tid] = best; s_pos[tid] = best_pos; s_sum[tid] = asum;
  __syncthreads();
  for (int s = PC_THREADS / 2; s > 0; s >>= 1) {
    if (tid < s) {
      float v2 = s_val[tid + s]; int p2 = s_pos[tid + s];
      if (v2 > s_val[tid] || (v2 == s_val[tid] && p2 < s_pos[tid])) { s_val[tid] = v2; s_pos[tid] = p2; }
      s_sum[tid] += s_sum[tid + s];
    }
    __syncthreads();
  }
  if (tid == 0) {
    const double tot = s_sum[0];
    st->counter = 0;
    st->rank = m + 1;
    const float err = (float)(tot / (double)st->orig_err);
    st->err = err;
    const float mx = s_val[0];
    const int pp = s_pos[0];
    // while (m == 0) or (m < max_iter and max(errors) > error_tol): will step m+1 run?
    if (m + 1 >= max_rank || (int64_t)(m + 1) >= n || !(err > tol)) {
      st->done = 1;
    } else if (pp < 0 || !(mx > 0.f)) {  // NaN / non-positive pivot: the reference ends up with NaNs in L
      st->nan_flag = 1;
      st->done = 1;
    } else {
      const int pi_new = perm[pp];
      const int pi_old = perm[m + 1];
      perm[m + 1] = pi_new; perm[pp] = pi_old;
      pos[pi_new] = m + 1; pos[pi_old] = pp;
      st->pivot = pi_new;
      st->dpiv = sqrtf(mx);
      piv_out[m + 1] = (int64_t)pi_new;
    }
  }
}

// ---- persistent variant: ALL steps in one cooperative launch -----------------------------------------------------------
// The step-wise path above pays a kernel launch + a last-CTA hand-off (~20 us) per step for ~2 us of work.  Here the
// grid stays resident (cudaLaunchCooperativeKernel guarantees co-residency), every thread owns the same rows in every
// step, and the steps are separated by one grid barrier (pc_persistent1_kernel below).  Arithmetic per entry is identical to
// pc_step_kernel.
__device__ __forceinline__ void pc_grid_barrier(unsigned int* bar, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    while (*((volatile unsigned int*)bar) < target) { }
    __threadfence();
  }
  __syncthreads();
}

constexpr int PCP_THREADS = 384;  // persistent kernel: one CTA per SM keeps the grid barrier small (148 arrivals)
constexpr int PCP_RED = 512;

// Persistent kernel with ONE grid barrier per step.  Every CTA reduces the per-CTA partials itself (same loads, same tree => the
// same pivot, error and stop decision everywhere), so no second barrier ("state published"; the round-1 form had one) is
// needed: nothing global is read back except the partials, which are double-buffered by step parity (a CTA can only
// overwrite buffer m & 1 at step m + 2, i.e. after barrier m + 1, which every CTA reaches after it has read the step-m
// partials).  The permutation is not materialised: a row's position is private to the thread that owns the row (pos[j],
// patched by the owner when the row is swapped), the winning partial carries its row index, and the row that sits at
// position m + 1 (the one the swap moves to the winner's old position) announces itself through one more partial.
// Arithmetic and tie-breaking (earliest position) per entry are those of pc_step_kernel: bit-identical pivots.
// kernel sums (GP_BACKEND_SUM): K[pivot, j] = sum_t os_t k_t(|z_t,pivot - z_t,j|^2), every term with its own packed inputs
constexpr int PC_KIND_SUM = 64;
struct PcTerms {
  int n;
  int kind[4], DP[4];
  float os[4];
  const float* Z[4];
};
__device__ __forceinline__ float pc_cov_rt(int kind, float a) {
  switch (kind) {
    case GP_RBF: return cov_from_arg<GP_RBF>(a);
    case GP_MATERN12: return cov_from_arg<GP_MATERN12>(a);
    case GP_MATERN32: return cov_from_arg<GP_MATERN32>(a);
    default: return cov_from_arg<GP_MATERN52>(a);
  }
}

struct PcPart {
  double sum;
  float val;
  int pos, idx, old;
};

template <int KIND>
__global__ void __launch_bounds__(PCP_THREADS)
pc_persistent1_kernel(const float* __restrict__ Z, int DP, float os, float* Lt, int64_t n, int max_rank, float tol,
                      float* diag, int* pos, PcState* st, int64_t* piv_out, PcPart* part, const PcTerms tt) {
  extern __shared__ float sh[];
  float* zp = sh;
  float* lp = sh + DP;
  float* col = lp + max_rank;
  __shared__ float s_val[PCP_RED];
  __shared__ int s_pos[PCP_RED];
  __shared__ int s_idx[PCP_RED];
  __shared__ double s_sum[PCP_RED];
  __shared__ int s_old;
  const int tid = threadIdx.x;
  const int G = (int)gridDim.x;
  if (tid < PCP_RED - PCP_THREADS) {
    s_val[PCP_THREADS + tid] = -INFINITY; s_pos[PCP_THREADS + tid] = 0x7fffffff; s_idx[PCP_THREADS + tid] = -1; s_sum[PCP_THREADS + tid] = 0.0;
  }
  const int64_t stride = (int64_t)G * PCP_THREADS;
  const int64_t jfirst = (int64_t)blockIdx.x * PCP_THREADS + tid;
  // CTA-uniform running state (identical in every CTA)
  int pi = st->pivot;                 // pivot row of the step about to run (position m)
  float dpiv = st->dpiv;
  const float orig_err = st->orig_err;
  int fx_new = -1, fx_old = -1, fx_pp = 0;   // pending position patches from the previous selection
  auto better = [](float v2, int p2, float v1, int p1) { return v2 > v1 || (v2 == v1 && p2 < p1); };
  for (int m = 0; m < max_rank; ++m) {
    if (tid == 0) s_old = -1;
    if (KIND == PC_KIND_SUM) {   // DP = sum of the terms' widths: the pivot rows of all terms back to back
      int off = 0;
      for (int t = 0; t < tt.n; ++t) {
        for (int c = tid; c < tt.DP[t]; c += PCP_THREADS) zp[off + c] = tt.Z[t][(int64_t)pi * tt.DP[t] + c];
        off += tt.DP[t];
      }
    } else {
      for (int c = tid; c < DP; c += PCP_THREADS) zp[c] = Z[(int64_t)pi * DP + c];
    }
    for (int q = tid; q < m; q += PCP_THREADS) lp[q] = __ldcg(Lt + (int64_t)q * n + pi);   // written by another SM, before the last barrier
    __syncthreads();
    float best = -INFINITY;
    int best_pos = 0x7fffffff, best_idx = -1;
    double asum = 0.0;
    float* Lm = Lt + (int64_t)m * n;
    for (int64_t j = jfirst; j < n; j += stride) {
      int pj = pos[j];                                   // owner-private
      if ((int)j == fx_new) { pj = m; pos[j] = m; }      // pos[pi_new] = m (this step's pivot)
      else if ((int)j == fx_old) { pj = fx_pp; pos[j] = fx_pp; }
      if (pj == m + 1) s_old = (int)j;                   // exactly one thread of the grid
      const bool cached = (j == jfirst);
      if (pj < m) {
        Lm[j] = 0.f;
        if (cached) col[m * PCP_THREADS + tid] = 0.f;
      } else if (pj == m) {
        Lm[j] = dpiv;
        if (cached) col[m * PCP_THREADS + tid] = dpiv;
      } else {
        float v;
        if (KIND == PC_KIND_SUM) {
          v = 0.f;
          int off = 0;
          for (int t = 0; t < tt.n; ++t) {
            const float* zj = tt.Z[t] + j * tt.DP[t];
            float s = 0.f;
            for (int c = 0; c < tt.DP[t]; ++c) {
              float df = zp[off + c] - zj[c];
              s = fmaf(df, df, s);
            }
            v = fmaf(tt.os[t], pc_cov_rt(tt.kind[t], -0.5f * s), v);
            off += tt.DP[t];
          }
        } else {
          float s = 0.f;
          for (int c = 0; c < DP; ++c) {
            float df = zp[c] - Z[j * DP + c];
            s = fmaf(df, df, s);
          }
          v = os * cov_from_arg<(KIND == PC_KIND_SUM ? GP_RBF : KIND)>(-0.5f * s);
        }
        {
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
          int q = 0;
          if (cached) {
            const float* cj = col + tid;
            for (; q + 4 <= m; q += 4) {
              s0 = fmaf(lp[q], cj[q * PCP_THREADS], s0);
              s1 = fmaf(lp[q + 1], cj[(q + 1) * PCP_THREADS], s1);
              s2 = fmaf(lp[q + 2], cj[(q + 2) * PCP_THREADS], s2);
              s3 = fmaf(lp[q + 3], cj[(q + 3) * PCP_THREADS], s3);
            }
            for (; q < m; ++q) s0 = fmaf(lp[q], cj[q * PCP_THREADS], s0);
          } else {
            for (; q + 4 <= m; q += 4) {
              s0 = fmaf(lp[q], Lt[(int64_t)q * n + j], s0);
              s1 = fmaf(lp[q + 1], Lt[(int64_t)(q + 1) * n + j], s1);
              s2 = fmaf(lp[q + 2], Lt[(int64_t)(q + 2) * n + j], s2);
              s3 = fmaf(lp[q + 3], Lt[(int64_t)(q + 3) * n + j], s3);
            }
            for (; q < m; ++q) s0 = fmaf(lp[q], Lt[(int64_t)q * n + j], s0);
          }
          v -= (s0 + s1) + (s2 + s3);
        }
        v /= dpiv;
        Lm[j] = v;
        if (cached) col[m * PCP_THREADS + tid] = v;
        const float dn = diag[j] - v * v;
        diag[j] = dn;
        float cv; int cp;
        if (dn != dn) { cv = INFINITY; cp = -1; }
        else { cv = dn; cp = pj; }
        if (better(cv, cp, best, best_pos)) { best = cv; best_pos = cp; best_idx = (int)j; }
        asum += fabs((double)dn);
      }
    }
    s_val[tid] = best; s_pos[tid] = best_pos; s_idx[tid] = best_idx; s_sum[tid] = asum;
    __syncthreads();
    for (int s = PCP_RED / 2; s > 0; s >>= 1) {
      if (tid < s && tid + s < PCP_RED) {
        if (better(s_val[tid + s], s_pos[tid + s], s_val[tid], s_pos[tid])) { s_val[tid] = s_val[tid + s]; s_pos[tid] = s_pos[tid + s]; s_idx[tid] = s_idx[tid + s]; }
        s_sum[tid] += s_sum[tid + s];
      }
      __syncthreads();
    }
    PcPart* mine = part + (size_t)(m & 1) * G;
    if (tid == 0) {
      PcPart pp; pp.sum = s_sum[0]; pp.val = s_val[0]; pp.pos = s_pos[0]; pp.idx = s_idx[0]; pp.old = s_old;
      mine[blockIdx.x] = pp;
    }
    pc_grid_barrier(&st->bar, (unsigned)(m + 1) * gridDim.x);   // all partials of step m are visible
    best = -INFINITY; best_pos = 0x7fffffff; best_idx = -1; asum = 0.0;
    int old = -1;
    for (int b = tid; b < G; b += PCP_THREADS) {   // fixed assignment + fixed tree => deterministic, identical in every CTA
      const PcPart* q = mine + b;
      const float v2 = __ldcg(&q->val); const int p2 = __ldcg(&q->pos);
      if (better(v2, p2, best, best_pos)) { best = v2; best_pos = p2; best_idx = __ldcg(&q->idx); }
      asum += __ldcg(&q->sum);
      old = max(old, __ldcg(&q->old));
    }
    s_val[tid] = best; s_pos[tid] = best_pos; s_idx[tid] = best_idx; s_sum[tid] = asum;
    if (old >= 0) s_old = old;     // at most one thread of the CTA (after the loop barrier below: s_old is re-read only then)
    __syncthreads();
    for (int s = PCP_RED / 2; s > 0; s >>= 1) {
      if (tid < s) {
        if (better(s_val[tid + s], s_pos[tid + s], s_val[tid], s_pos[tid])) { s_val[tid] = s_val[tid + s]; s_pos[tid] = s_pos[tid + s]; s_idx[tid] = s_idx[tid + s]; }
        s_sum[tid] += s_sum[tid + s];
      }
      __syncthreads();
    }
    const double tot = s_sum[0];
    const float err = (float)(tot / (double)orig_err);
    const float mx = s_val[0];
    const int ppw = s_pos[0], inew = s_idx[0], iold = s_old;
    bool stop = false, nan = false;
    if (m + 1 >= max_rank || (int64_t)(m + 1) >= n || !(err > tol)) stop = true;
    else if (ppw < 0 || !(mx > 0.f)) { stop = true; nan = true; }
    if (blockIdx.x == 0 && tid == 0) {
      st->rank = m + 1;
      st->err = err;
      if (stop) st->done = 1;
      if (nan) st->nan_flag = 1;
      if (!stop) { st->pivot = inew; st->dpiv = sqrtf(mx); piv_out[m + 1] = (int64_t)inew; }
    }
    if (stop) break;
    pi = inew; dpiv = sqrtf(mx);
    fx_new = inew; fx_old = (iold == inew) ? -1 : iold; fx_pp = ppw;
    __syncthreads();   // s_val / s_old are rewritten at the top of the next step
  }
}

__global__ void pc_init_kernel(float* __restrict__ diag, int* __restrict__ perm, int* __restrict__ pos, int64_t n, float os,
                               PcState* __restrict__ st, int64_t* __restrict__ piv_out) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j == 0) {
    // the initial diagonal of a stationary kernel is constant: torch.max returns the first entry -> pivot 0
    st->done = 0; st->rank = 0; st->pivot = 0; st->nan_flag = (os > 0.f) ? 0 : 1; st->dpiv = sqrtf(os);
    st->orig_err = os; st->err = 0.f; st->counter = 0u; st->bar = 0u;
    piv_out[0] = 0;
  }
  if (j >= n) return;
  diag[j] = os;  // _approx_diagonal of a stationary kernel
  perm[j] = (int)j;
  pos[j] = (int)j;
}

// ---- preconditioner factor ---------------------------------------------------------------------
// Gpart[z][a][b] = sum_{j in slice z} L[a][j] L[b][j] s_j     (s_j = 1, or 1 / d_j for a per-row noise diagonal)
// 64 x 64 output tiles, 4 x 4 register tiles per thread.  Products and sums run in fp32 over chunks of 64 columns (<= 64 terms of
// magnitude <= outputscale: absolute error ~1e-6 per chunk), every chunk is then added to fp64 accumulators: the error of G
// stays orders of magnitude below the noise floor it is compared with in  log det P = log det(L^T L + sigma^2 I) + ...  (an
// error E in G shifts log det P by ~tr(E) / sigma^2), at a fraction of the cost of the all-fp64 product of round 1 (0.38 ms at C2).
constexpr int GT = 64;
__global__ void __launch_bounds__(256)
gram_kernel(const float* __restrict__ Lt, int k, int64_t n, int64_t jslice, const float* __restrict__ dvec, double* __restrict__ Gpart) {
  if (blockIdx.x > blockIdx.y) return;   // lower triangle of tiles only; the mirror image is written below
  __shared__ __align__(16) float As[32][GT + 4];   // [column][row of L^T = index a]
  __shared__ __align__(16) float Bs[32][GT + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int a0 = blockIdx.y * GT, b0 = blockIdx.x * GT;
  const int64_t j_begin = (int64_t)blockIdx.z * jslice, j_end = min(n, j_begin + jslice);
  double acc64[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc64[i][j] = 0.0;
  for (int64_t j0 = j_begin; j0 < j_end; j0 += 64) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int hf = 0; hf < 2; ++hf) {
      const int64_t jb = j0 + hf * 32;
      __syncthreads();
      for (int e = tid; e < GT * 32; e += 256) {
        const int r = e >> 5, cc = e & 31;
        const int64_t j = jb + cc;
        const bool ok = j < j_end;
        const float sc = (ok && dvec) ? 1.f / dvec[j] : 1.f;
        As[cc][r] = (ok && a0 + r < k) ? Lt[(int64_t)(a0 + r) * n + j] * sc : 0.f;
        Bs[cc][r] = (ok && b0 + r < k) ? Lt[(int64_t)(b0 + r) * n + j] : 0.f;
      }
      __syncthreads();
#pragma unroll 8
      for (int cc = 0; cc < 32; ++cc) {
        const float4 av = *reinterpret_cast<const float4*>(&As[cc][ty * 4]);
        const float4 bv = *reinterpret_cast<const float4*>(&Bs[cc][tx * 4]);
        const float a4[4] = {av.x, av.y, av.z, av.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc64[i][j] += (double)acc[i][j];
  }
  double* G = Gpart + (int64_t)blockIdx.z * k * k;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int a = a0 + ty * 4 + i, b = b0 + tx * 4 + j;
      if (a < k && b < k) {
        G[(int64_t)a * k + b] = acc64[i][j];
        if (blockIdx.x != blockIdx.y) G[(int64_t)b * k + a] = acc64[i][j];
      }
    }
}

// single CTA: G = sum_z Gpart + noise I ; in-place lower Cholesky C ; logdet
// diag_add = sigma^2 and logdet_tail = (n - k) log sigma^2 for the constant diagonal; 1 and sum_j log d_j for a per-row diagonal
// (then G = L^T D^-1 L and log det P = log det(I + L^T D^-1 L) + sum log d)
__global__ void chol_small_kernel(const double* __restrict__ Gpart, int nz, int k, double diag_add, const double* __restrict__ logdet_tail,
                                  double* __restrict__ C, double* __restrict__ logdet_out, int* __restrict__ fail) {
  extern __shared__ double G[];  // [k][k]
  const int tid = threadIdx.x;
  for (int e = tid; e < k * k; e += blockDim.x) {
    double s = 0.0;
    for (int z = 0; z < nz; ++z) s += Gpart[(int64_t)z * k * k + e];
    if (e / k == e % k) s += diag_add;
    G[e] = s;
  }
  __syncthreads();
  // Right-looking factorisation with ONE barrier per step: the scaled column j goes straight to the output (nobody reads it
  // back), the trailing update uses the unscaled column and 1 / d_jj:  G[i][l] -= G[i][j] G[l][j] / d_jj.
  double ld = 0.0;
  for (int j = 0; j < k; ++j) {
    double d = G[j * k + j];
    if (!(d > 0.0)) { if (tid == 0) *fail = 1; d = 1e-300; }
    const double rinv = 1.0 / d;
    if (tid == 0) ld += log(d);
    if (tid >= 256) {   // upper half of the CTA also writes column j of C (rows j..k-1) and the zeros above the diagonal
      const double rs = 1.0 / sqrt(d);
      for (int i = tid - 256; i < k; i += (int)blockDim.x - 256) C[i * k + j] = (i < j) ? 0.0 : G[i * k + j] * rs;
    }
    // (2-D thread mapping, no integer division: every (i, l), j < l <= i, is updated exactly once)
    for (int i = j + 1 + (tid >> 5); i < k; i += (int)(blockDim.x >> 5)) {
      const double gij = G[i * k + j] * rinv;
      for (int l = j + 1 + (tid & 31); l <= i; l += 32) G[i * k + l] -= gij * G[l * k + j];
    }
    __syncthreads();
  }
  if (tid == 0) *logdet_out = ld + *logdet_tail;
}

// Cinv = C^{-1} (lower triangular, fp64): one WARP (= one CTA) per column j, forward substitution with the column held in
// registers (entry b on lane b & 31) and the inner product of step i split over the lanes + a shuffle tree: ~100 dependent
// steps of ~150 cycles instead of 5000 dependent FMAs of one thread per column.  k <= 128.
__global__ void __launch_bounds__(32) cinv_kernel(const double* __restrict__ C, int k, double* __restrict__ Cinv) {
  const int j = blockIdx.x, lane = threadIdx.x;
  double x[4] = {0.0, 0.0, 0.0, 0.0};   // x[t] = Cinv[lane + 32 t][j]
  for (int i = 0; i < k; ++i) {
    double xi = 0.0;
    if (i >= j) {
      const double* ci = C + (size_t)i * k;
      double s = 0.0;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int b = lane + 32 * t;
        if (b >= j && b < i) s = fma(ci[b], x[t], s);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      xi = (((i == j) ? 1.0 : 0.0) - s) / ci[i];
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (lane + 32 * t == i) x[t] = xi;
    }
    if (lane == 0) Cinv[(size_t)i * k + j] = xi;
  }
}

constexpr int WS_BLOCKS = 4;
// W[r][a] = sum_{b<=a} Cinv[a][b] L[b][r]   (W = L C^{-T}); 32 rows x 4 interleaved a-groups per pass, fp64 accumulate
__global__ void __launch_bounds__(128)
wsolve_kernel(const float* __restrict__ Lt, int k, int64_t n_total, int64_t row_begin, int64_t n_local,
              const double* __restrict__ Cinv, const float* __restrict__ dvec, float* __restrict__ W) {
  extern __shared__ double shw[];
  double* Ci = shw;                        // [k][k]
  double* Ls = shw + (size_t)k * k;        // [k][32], fp64 copy of the L rows (exact conversion, once per element)
  for (int e = threadIdx.x; e < k * k; e += 128) Ci[e] = Cinv[e];
  const int rl = threadIdx.x & 31, ag = threadIdx.x >> 5;
  for (int blk = 0; blk < WS_BLOCKS; ++blk) {          // C^{-1} (80 KB at k = 100) is loaded once per WS_BLOCKS * 32 rows
    const int64_t r0 = ((int64_t)blockIdx.x * WS_BLOCKS + blk) * 32;
    if (r0 >= n_local) break;
    __syncthreads();
    for (int e = threadIdx.x; e < k * 32; e += 128) {
      int b = e >> 5, rr = e & 31;
      Ls[e] = (r0 + rr < n_local) ? (double)Lt[(int64_t)b * n_total + row_begin + r0 + rr] : 0.0;
    }
    __syncthreads();
    const int64_t r = r0 + rl;
    for (int a = ag; a < k; a += 4) {
      double s = 0.0;
      for (int b = 0; b <= a; ++b) s = fma(Ci[a * k + b], Ls[b * 32 + rl], s);
      if (r < n_local) W[r * k + a] = (float)(dvec ? s / (double)dvec[row_begin + r] : s);   // per-row noise: W = D^-1 L C^-T
    }
  }
}

// sum_j log d_j (fp64, one CTA; the per-row-noise tail of log det P)
__global__ void logsum_kernel(const float* __restrict__ d, int64_t n, double* __restrict__ out) {
  __shared__ double sh[256];
  double s = 0.0;
  for (int64_t j = threadIdx.x; j < n; j += 256) s += log((double)d[j]);
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st) sh[threadIdx.x] += sh[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sh[0];
}

// Z[r][c] = sum_a L[a][r] eps1[a][c] + sigma eps2[r][c]
__global__ void probes_kernel(const float* __restrict__ Lt, int k, int64_t n_total, int64_t row_begin, int64_t n_local,
                              const float* __restrict__ eps1, const float* __restrict__ eps2, int tp, float sigma,
                              const float* __restrict__ dvec, float* __restrict__ Z) {
  extern __shared__ float e1[];  // [k][tp]
  for (int e = threadIdx.x; e < k * tp; e += blockDim.x) e1[e] = eps1[e];
  __syncthreads();
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_local * tp) return;
  int64_t r = idx / tp;
  int c = (int)(idx % tp);
  float s = (dvec ? sqrtf(dvec[row_begin + r]) : sigma) * eps2[r * tp + c];   // z = L eps1 + D^1/2 eps2
  for (int a = 0; a < k; ++a) s = fmaf(Lt[(int64_t)a * n_total + row_begin + r], e1[a * tp + c], s);
  Z[r * tp + c] = s;
}

}  // namespace gp

using namespace gp;

extern "C" int gp_pivoted_cholesky(gp_plan* p, int rank, float error_tol, float* Lt, int64_t* piv, int* rank_out) {
  GP_REQUIRE(p && p->data_set && p->hypers_set, GP_E_STATE, "plan not ready");
  GP_REQUIRE(p->same, GP_E_SHAPE, "pivoted Cholesky needs a square operator");
  GP_REQUIRE(p->backend != GP_BACKEND_SKI, GP_E_SHAPE, "pivoted Cholesky is not available for the SKI backend");
  GP_CHECK(sum_prepare(p));
  const int64_t n = p->n2;
  GP_REQUIRE(n < (int64_t)1 << 31, GP_E_SHAPE, "n too large");
  rank = (int)std::min<int64_t>(rank, n);
  GP_REQUIRE(rank >= 1, GP_E_SHAPE, "rank must be >= 1");
  cudaStream_t st = p->stream;
  const unsigned gb = (unsigned)cdiv(n, PC_THREADS);
  GP_CHECK(p->pcdiag.ensure(sizeof(float) * n));
  GP_CHECK(p->pcperm.ensure(sizeof(int) * n));
  GP_CHECK(p->pcpos.ensure(sizeof(int) * n));
  GP_CHECK(p->pcstate.ensure(sizeof(PcState) + 256 + (sizeof(float) + sizeof(int) + sizeof(double)) * gb + 64));
  float* diag = p->pcdiag.as<float>();
  int* perm = p->pcperm.as<int>();
  int* pos = p->pcpos.as<int>();
  PcState* S = p->pcstate.as<PcState>();
  double* psum = reinterpret_cast<double*>(reinterpret_cast<char*>(S) + 256);
  float* pval = reinterpret_cast<float*>(psum + gb);
  int* ppos = reinterpret_cast<int*>(pval + gb);
  GP_CUDA(cudaMemsetAsync(Lt, 0, sizeof(float) * (size_t)rank * n, st));
  GP_CUDA(cudaMemsetAsync(piv, 0, sizeof(int64_t) * rank, st));
  const bool sum = p->backend == GP_BACKEND_SUM;
  PcTerms tt;
  memset(&tt, 0, sizeof(tt));
  float os_total = p->outputscale;
  int dp_total = p->DP;
  if (sum) {
    os_total = 0.f;
    dp_total = 0;
    tt.n = (int)p->terms.size();
    for (int t = 0; t < tt.n; ++t) {
      const gp_plan* q = p->terms[t];
      tt.kind[t] = q->kind; tt.DP[t] = q->DP; tt.os[t] = q->outputscale; tt.Z[t] = q->Z2.as<float>();
      os_total += q->outputscale;
      dp_total += q->DP;
    }
  }
  pc_init_kernel<<<gb, PC_THREADS, 0, st>>>(diag, perm, pos, n, os_total, S, piv);   // stationary terms: constant initial diagonal
  p->launches += 1;
  const float* Z = sum ? nullptr : p->Z2.as<float>();
  const bool stepwise = getenv("GP_PC_STEPWISE") != nullptr;   // debugging / A-B switch: one launch per step
  int coop = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, p->device);
  GP_REQUIRE(!sum || (coop && !stepwise), GP_E_STATE, "pivoted Cholesky of a kernel sum needs the cooperative kernel");
  if (coop && !stepwise) {
    const size_t sh = sizeof(float) * (dp_total + rank + (size_t)rank * PCP_THREADS);
    const void* fn1;
    switch (sum ? PC_KIND_SUM : p->kind) {
      case GP_RBF: fn1 = (const void*)pc_persistent1_kernel<GP_RBF>; break;
      case GP_MATERN12: fn1 = (const void*)pc_persistent1_kernel<GP_MATERN12>; break;
      case GP_MATERN32: fn1 = (const void*)pc_persistent1_kernel<GP_MATERN32>; break;
      case PC_KIND_SUM: fn1 = (const void*)pc_persistent1_kernel<PC_KIND_SUM>; break;
      default: fn1 = (const void*)pc_persistent1_kernel<GP_MATERN52>; break;
    }
    GP_CUDA(cudaFuncSetAttribute(fn1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
    int per_sm1 = 0;
    GP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm1, fn1, PCP_THREADS, sh));
    GP_REQUIRE(per_sm1 >= 1, GP_E_CUDA, "pivoted Cholesky kernel does not fit on an SM");
    const unsigned grid1 = (unsigned)std::min<int64_t>(cdiv(n, PCP_THREADS), (int64_t)per_sm1 * p->n_sm);
    GP_CHECK(p->pcpart.ensure(sizeof(PcPart) * 2 * (size_t)grid1));
    PcPart* part = p->pcpart.as<PcPart>();
    int DPv = dp_total, rk = rank;
    float osv = os_total, tolv = error_tol;
    int64_t nn = n;
    void* args[] = {(void*)&Z, &DPv, &osv, &Lt, &nn, &rk, &tolv, &diag, &pos, &S, &piv, &part, &tt};
    GP_CUDA(cudaLaunchCooperativeKernel(fn1, dim3(grid1), dim3(PCP_THREADS), args, sh, st));
    p->launches += 1;
  } else {
  for (int m = 0; m < rank; ++m) {
    size_t sh = sizeof(float) * (p->DP + m);
#define GP_PC_LAUNCH(KK) pc_step_kernel<KK><<<gb, PC_THREADS, sh, st>>>(Z, p->DP, p->outputscale, Lt, n, m, rank, error_tol, diag, perm, pos, S, piv, pval, ppos, psum)
    switch (p->kind) {
      case GP_RBF: GP_PC_LAUNCH(GP_RBF); break;
      case GP_MATERN12: GP_PC_LAUNCH(GP_MATERN12); break;
      case GP_MATERN32: GP_PC_LAUNCH(GP_MATERN32); break;
      default: GP_PC_LAUNCH(GP_MATERN52); break;
    }
#undef GP_PC_LAUNCH
    p->launches += 1;
  }
  }
  GP_CUDA(cudaGetLastError());
  PcState* hs = reinterpret_cast<PcState*>(reinterpret_cast<char*>(p->pinned) + 2048);
  GP_CUDA(cudaMemcpyAsync(hs, S, sizeof(PcState), cudaMemcpyDeviceToHost, st));
  GP_CUDA(cudaStreamSynchronize(st));
  if (rank_out) *rank_out = hs->rank;
  if (hs->nan_flag) {
    set_error("NaNs encountered in preconditioner computation. Attempting to continue without preconditioning.");
    return GP_W_PIVCHOL_NAN;
  }
  return GP_OK;
}

extern "C" int gp_precond_build(gp_plan* p, const float* Lt, int k, float* W, double* logdet_out) {
  GP_REQUIRE(p && p->data_set && p->hypers_set, GP_E_STATE, "plan not ready");
  GP_REQUIRE(k >= 1 && k <= 128, GP_E_SHAPE, "preconditioner rank %d not in [1,128]", k);
  const float* dvec = p->noise_diag;
  GP_REQUIRE(dvec != nullptr || p->noise > 0.f, GP_E_SHAPE, "preconditioner needs noise > 0");
  cudaStream_t st = p->stream;
  const int64_t n = p->n2;
  const int ntile = (int)cdiv(k, GT);
  const int nz = (int)std::min<int64_t>(64, std::max<int64_t>(1, std::min<int64_t>(n / 1024, (2 * p->n_sm) / (ntile * (ntile + 1) / 2))));
  const int64_t jslice = cdiv(cdiv(n, nz), 64) * 64;
  GP_CHECK(p->gram.ensure(sizeof(double) * (size_t)nz * k * k));
  GP_CHECK(p->cholC.ensure(sizeof(double) * ((size_t)2 * k * k + 4) + 64));
  double* C = p->cholC.as<double>();
  double* Cinv = C + (size_t)k * k;
  double* d_logdet = Cinv + (size_t)k * k;
  double* d_tail = d_logdet + 1;
  int* d_fail = reinterpret_cast<int*>(d_tail + 1);
  GP_CUDA(cudaMemsetAsync(d_fail, 0, sizeof(int), st));
  dim3 gg((unsigned)ntile, (unsigned)ntile, (unsigned)nz);
  gram_kernel<<<gg, 256, 0, st>>>(Lt, k, n, jslice, dvec, p->gram.as<double>());
  if (dvec) {
    logsum_kernel<<<1, 256, 0, st>>>(dvec, n, d_tail);
    p->launches++;
  } else {
    const double tail = (double)(n - k) * log((double)p->noise);
    GP_CUDA(cudaMemcpyAsync(d_tail, &tail, sizeof(double), cudaMemcpyHostToDevice, st));   // pageable source: copied before return
  }
  const size_t shc = sizeof(double) * (size_t)k * k;
  static bool attr_done[64] = {};
  if (!attr_done[p->device & 63]) {
    GP_CUDA(cudaFuncSetAttribute(chol_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));   // k <= 128: 128 KB
    GP_CUDA(cudaFuncSetAttribute(wsolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 168 * 1024));       // 128 KB + 32 KB
    attr_done[p->device & 63] = true;
  }
  chol_small_kernel<<<1, 512, shc, st>>>(p->gram.as<double>(), nz, k, dvec ? 1.0 : (double)p->noise, d_tail, C, d_logdet, d_fail);
  // W = L C^-T through the explicit inverse (a per-row forward substitution against C in shared memory was tried: 0.52 ms at C2
  // against 0.16 + 0.19 ms for these two kernels -- one 128-thread CTA per SM is latency bound on 5000 dependent steps per row)
  cinv_kernel<<<k, 32, 0, st>>>(C, k, Cinv);
  wsolve_kernel<<<(unsigned)cdiv(p->row_count, 32 * WS_BLOCKS), 128, shc + sizeof(double) * (size_t)k * 32, st>>>(Lt, k, n, p->row_begin,
                                                                                                      p->row_count, Cinv, dvec, W);
  p->launches += 4;
  GP_CUDA(cudaGetLastError());
  double* h = reinterpret_cast<double*>(reinterpret_cast<char*>(p->pinned) + 3072);
  GP_CUDA(cudaMemcpyAsync(h, d_logdet, sizeof(double) * 2 + sizeof(int), cudaMemcpyDeviceToHost, st));
  GP_CUDA(cudaStreamSynchronize(st));
  if (logdet_out) *logdet_out = h[0];
  int fail = *reinterpret_cast<int*>(h + 2);
  GP_REQUIRE(!fail, GP_W_PIVCHOL_NAN, "preconditioner Gram matrix is not positive definite");
  return GP_OK;
}

extern "C" int gp_precond_probes(gp_plan* p, const float* Lt, int k, const float* eps1, const float* eps2, int tp, float* Z) {
  GP_REQUIRE(p && p->data_set && p->hypers_set, GP_E_STATE, "plan not ready");
  GP_REQUIRE(k >= 1 && tp >= 1 && (size_t)k * tp * 4 <= 40 * 1024, GP_E_SHAPE, "bad probe shape k=%d tp=%d", k, tp);
  int64_t tot = p->row_count * tp;
  probes_kernel<<<(unsigned)cdiv(tot, 256), 256, sizeof(float) * k * tp, p->stream>>>(Lt, k, p->n2, p->row_begin, p->row_count,
                                                                                  eps1, eps2, tp, sqrtf(p->noise), p->noise_diag, Z);
  p->launches++;
  GP_CUDA(cudaGetLastError());
  return GP_OK;
}
