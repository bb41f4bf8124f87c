// kmv_simt.cu -- fp32 CUDA-core fused kernel-matmul, row extraction, diagonal and the
// bilinear hyper-parameter gradient.
//
// The SIMT K.V kernel is the bring-up / cross-check path for the tcgen05 kernel (kmv_tc.cu) and
// the backend for d > 41.  It evaluates a_ij = -0.5 |z_i - z_j|^2 by direct differences (no
// cancellation), so it is also the more accurate of the two.
// Reference semantics: LazyEvaluatedKernelTensor._matmul (lazy/lazy_evaluated_kernel_tensor.py:245-276),
// _getitem (:136-243), _diagonal (:107-133), _bilinear_derivative (:69-105).
#include "gp_common.cuh"

namespace gp {

// grid (row blocks, nsplit); 128 threads, one output row each; partial[split][row][16]
template <int KIND, int DP>
__global__ void __launch_bounds__(SIMT_TI)
kmv_simt_kernel(const float* __restrict__ Z1, const float* __restrict__ Z2, const float* __restrict__ V16,
                float* __restrict__ partial, int64_t n1, int64_t n2, int64_t rows_pad, int64_t cols_per_split,
                int same, int64_t row_begin, const int* __restrict__ done_flag) {
  if (done_flag && *done_flag) return;
  __shared__ __align__(16) float zj[SIMT_TJ][DP];
  __shared__ __align__(16) float vj[SIMT_TJ][TP];
  const int tid = threadIdx.x;
  const int64_t i = (int64_t)blockIdx.x * SIMT_TI + tid;
  const int split = blockIdx.y;
  const int64_t j_begin = (int64_t)split * cols_per_split;
  const int64_t j_end = min(n2, j_begin + cols_per_split);
  float zi[DP];
  const bool rv = i < n1;
#pragma unroll
  for (int c = 0; c < DP; ++c) zi[c] = rv ? Z1[i * DP + c] : 0.f;
  float acc[TP];
#pragma unroll
  for (int c = 0; c < TP; ++c) acc[c] = 0.f;
  const int64_t gi = i + row_begin;

  for (int64_t j0 = j_begin; j0 < j_end; j0 += SIMT_TJ) {
    const int nj = (int)min((int64_t)SIMT_TJ, j_end - j0);
    __syncthreads();
    for (int e = tid; e < SIMT_TJ * DP; e += SIMT_TI) {
      int jj = e / DP;
      (&zj[0][0])[e] = (jj < nj) ? Z2[j0 * DP + e] : 0.f;
    }
    for (int e = tid; e < SIMT_TJ * TP; e += SIMT_TI) {
      int jj = e / TP;
      (&vj[0][0])[e] = (jj < nj) ? V16[j0 * TP + e] : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int jj = 0; jj < SIMT_TJ; ++jj) {
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < DP; ++c) {
        float df = zi[c] - zj[jj][c];
        s = fmaf(df, df, s);
      }
      float a = -0.5f * s;
      if (same && (j0 + jj) == gi) a = 0.f;  // exact diagonal (kernel.py:44-45)
      float k = cov_from_arg<KIND>(a);
#pragma unroll
      for (int c = 0; c < TP; ++c) acc[c] = fmaf(k, vj[jj][c], acc[c]);
    }
  }
  if (i < rows_pad) {
    float4* dst = reinterpret_cast<float4*>(partial + ((int64_t)split * rows_pad + i) * TP);
#pragma unroll
    for (int q = 0; q < 4; ++q) dst[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
  }
}

template <int KIND>
static int launch_simt_kind(gp_plan* p, const float* V16, const int* done_flag) {
  const float* Z1 = p->same ? p->Z2.as<float>() + p->row_begin * p->DP : p->Z1.as<float>();
  const float* Z2 = p->Z2.as<float>();
  int64_t rows_pad = p->rows_pad;
  dim3 grid((unsigned)cdiv(p->row_count, SIMT_TI), (unsigned)p->nsplit);
  int64_t cps = p->tiles_per_split * SIMT_TJ;
#define GP_SIMT_CASE(D)                                                                                          \
  case D:                                                                                                        \
    kmv_simt_kernel<KIND, D><<<grid, SIMT_TI, 0, p->stream>>>(Z1, Z2, V16, partial_ptr(p), p->row_count, \
                                                              p->n2, rows_pad, cps, p->same ? 1 : 0,             \
                                                              p->row_begin, done_flag);                          \
    break;
  switch (p->DP) {
    GP_SIMT_CASE(4) GP_SIMT_CASE(8) GP_SIMT_CASE(12) GP_SIMT_CASE(16) GP_SIMT_CASE(24) GP_SIMT_CASE(32)
    GP_SIMT_CASE(48) GP_SIMT_CASE(64) GP_SIMT_CASE(96) GP_SIMT_CASE(128)
    default:
      set_error("unsupported DP=%d", p->DP);
      return GP_E_SHAPE;
  }
#undef GP_SIMT_CASE
  p->launches++;
  GP_CUDA(cudaGetLastError());
  return GP_OK;
}

int kmv_simt_launch(gp_plan* p, const float* V16, const int* done_flag) {
  switch (p->kind) {
    case GP_RBF: return launch_simt_kind<GP_RBF>(p, V16, done_flag);
    case GP_MATERN12: return launch_simt_kind<GP_MATERN12>(p, V16, done_flag);
    case GP_MATERN32: return launch_simt_kind<GP_MATERN32>(p, V16, done_flag);
    case GP_MATERN52: return launch_simt_kind<GP_MATERN52>(p, V16, done_flag);
  }
  set_error("bad kernel kind %d", p->kind);
  return GP_E_SHAPE;
}

int kmv_partials(gp_plan* p, const float* V16, const int* done_flag) {
  if (p->backend == GP_BACKEND_SKI) return ski_kmv_partials(p, V16, done_flag);
  if (p->backend == GP_BACKEND_SUM) {
    if (p->sum_any_tc) GP_CHECK(pack_v_tiles(p, V16));
    return sum_kmv_launch(p, V16, done_flag);
  }
  if (p->backend == GP_BACKEND_TCGEN05) {
    GP_CHECK(pack_v_tiles(p, V16));
    return kmv_tc_launch(p, done_flag);
  }
  return kmv_simt_launch(p, V16, done_flag);
}

// OUT[r, c] = os * sum_s partial[s][r][c] + noise * V16[row_begin + r][c]
__global__ void kmv_finish_user_kernel(const float* __restrict__ partial, int nsplit, int64_t rows, int64_t rows_pad,
                                       float os, const float* __restrict__ pscale, float noise_add, const float* __restrict__ dvec, const float* __restrict__ V16,
                                       int64_t row_begin, float* __restrict__ OUT, int64_t ldo, int t, const int* __restrict__ xbad) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * TP) return;
  int64_t r = idx / TP;
  int c = (int)(idx % TP);
  if (c >= t) return;
  float s = 0.f;
  if (pscale) {   // kernel sum: slot sp belongs to the term with outputscale pscale[sp]
    for (int sp = 0; sp < nsplit; ++sp) s = fmaf(pscale[sp], partial[((int64_t)sp * rows_pad + r) * TP + c], s);
  } else {
    for (int sp = 0; sp < nsplit; ++sp) s += partial[((int64_t)sp * rows_pad + r) * TP + c];
    s *= os;
  }
  float o = s;
  if (dvec) o = fmaf(dvec[row_begin + r], V16[(row_begin + r) * TP + c], o);
  else if (noise_add != 0.f) o = fmaf(noise_add, V16[(row_begin + r) * TP + c], o);
  if (*xbad) o = __int_as_float(0x7fc00000);
  OUT[r * ldo + c] = o;
}

int kmv_finish_user(gp_plan* p, const float* V16, float* OUT, int64_t ldo, int t, int add_noise) {
  int64_t rows_pad = p->rows_pad;
  int64_t tot = p->row_count * TP;
  float na = (add_noise && p->same) ? p->noise : 0.f;
  const float* dv = (add_noise && p->same) ? p->noise_diag : nullptr;
  kmv_finish_user_kernel<<<(unsigned)cdiv(tot, 256), 256, 0, p->stream>>>(p->partial.as<float>(), p->nparts, p->row_count,
                                                                          rows_pad, p->outputscale, part_scale_ptr(p), na, dv, V16, p->row_begin,
                                                                          OUT, ldo, t, p->xbad);
  p->launches++;
  GP_CUDA(cudaGetLastError());
  return GP_OK;
}

// ---- row extraction: OUT[m][n2] = os * k(x1[idx[r]], x2[j]) -----------------------------------
template <int KIND>
__global__ void krows_kernel(const float* __restrict__ Z1, const float* __restrict__ Z2, int DP,
                             const int64_t* __restrict__ idx, int64_t n1_local, int64_t n2, float os, int same,
                             int64_t row_begin, float* __restrict__ OUT, int64_t ldo) {
  extern __shared__ float zi[];
  const int64_t r = blockIdx.y;
  const int64_t i = idx[r];
  if (i < 0 || i >= n1_local) {   // out-of-range row index (CTA-uniform): NaN row instead of an out-of-bounds read
    int64_t jj = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (jj < n2) OUT[r * ldo + jj] = __int_as_float(0x7fc00000);
    return;
  }
  for (int c = threadIdx.x; c < DP; c += blockDim.x) zi[c] = Z1[i * DP + c];
  __syncthreads();
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n2) return;
  float s = 0.f;
  for (int c = 0; c < DP; ++c) {
    float df = zi[c] - Z2[j * DP + c];
    s = fmaf(df, df, s);
  }
  float a = -0.5f * s;
  if (same && (i + row_begin) == j) a = 0.f;
  OUT[r * ldo + j] = os * cov_from_arg<KIND>(a);
}

// diagonal of a cross-covariance K(x1, x2) (n1 == n2): OUT[i] = os * k(x1_i, x2_i)   (kernel(x1, x2, diag=True),
// lazy_evaluated_kernel_tensor.py:107-133 / kernels/kernel.py:307-352 with diag=True)
template <int KIND>
__global__ void kdiag_cross_kernel(const float* __restrict__ Z1, const float* __restrict__ Z2, int DP, int64_t n, float os,
                                   float* __restrict__ OUT) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int c = 0; c < DP; ++c) {
    float df = Z1[i * DP + c] - Z2[i * DP + c];
    s = fmaf(df, df, s);
  }
  OUT[i] = os * cov_from_arg<KIND>(-0.5f * s);
}

__global__ void fill_kernel(float* __restrict__ p, int64_t n, float v) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ---- bilinear derivative: sum_ij (Lf_i . Rt_j) dk_ij/dtheta ---------------------------------
// grid (row blocks, col splits); per-thread row, staged columns; outputs per block partial sums
//   gout[block][0]      = sum_ij w_ij k_ij                 (-> d/d outputscale)
//   gout[block][1 + c]  = sum_ij w_ij g_ij (z_ic-z_jc)^2   (ARD)  or  gout[block][1] = sum w_ij g_ij (scalar l)
template <int KIND, int DP, bool ARD>
__global__ void __launch_bounds__(SIMT_TI)
bilinear_kernel(const float* __restrict__ Z1, const float* __restrict__ Z2, const float* __restrict__ L16,
                const float* __restrict__ R16, int64_t n1, int64_t n2, int64_t cols_per_split, int same,
                int64_t row_begin, int d, double* __restrict__ gout, int gstride) {
  __shared__ __align__(16) float zj[SIMT_TJ][DP];
  __shared__ __align__(16) float rj[SIMT_TJ][TP];
  const int tid = threadIdx.x;
  const int64_t i = (int64_t)blockIdx.x * SIMT_TI + tid;
  const int64_t j_begin = (int64_t)blockIdx.y * cols_per_split;
  const int64_t j_end = min(n2, j_begin + cols_per_split);
  const bool rv = i < n1;
  float zi[DP], li[TP];
#pragma unroll
  for (int c = 0; c < DP; ++c) zi[c] = rv ? Z1[i * DP + c] : 0.f;
#pragma unroll
  for (int c = 0; c < TP; ++c) li[c] = rv ? L16[i * TP + c] : 0.f;
  constexpr int NG = ARD ? DP : 1;
  float gk = 0.f, gl[NG];
#pragma unroll
  for (int c = 0; c < NG; ++c) gl[c] = 0.f;
  const int64_t gi = i + row_begin;
  for (int64_t j0 = j_begin; j0 < j_end; j0 += SIMT_TJ) {
    const int nj = (int)min((int64_t)SIMT_TJ, j_end - j0);
    __syncthreads();
    for (int e = tid; e < SIMT_TJ * DP; e += SIMT_TI) (&zj[0][0])[e] = (e / DP < nj) ? Z2[j0 * DP + e] : 0.f;
    for (int e = tid; e < SIMT_TJ * TP; e += SIMT_TI) (&rj[0][0])[e] = (e / TP < nj) ? R16[j0 * TP + e] : 0.f;
    __syncthreads();
    for (int jj = 0; jj < nj; ++jj) {
      float w = 0.f;
#pragma unroll
      for (int c = 0; c < TP; ++c) w = fmaf(li[c], rj[jj][c], w);
      float s = 0.f;
      float df2[DP];
#pragma unroll
      for (int c = 0; c < DP; ++c) {
        float df = zi[c] - zj[jj][c];
        df2[c] = df * df;
        s += df2[c];
      }
      float a = -0.5f * s;
      if (same && (j0 + jj) == gi) a = 0.f;
      float k;
      float g = dcov_from_arg<KIND>(a, &k);
      gk = fmaf(w, k, gk);
      if (ARD) {
        // dk/dl_c = g * (z_ic - z_jc)^2 / (s * l_c)   (g/l is the scalar-lengthscale derivative)
        float gs = (s > 0.f) ? w * g / s : 0.f;
#pragma unroll
        for (int c = 0; c < NG; ++c) gl[c] = fmaf(gs, df2[c], gl[c]);
      } else {
        gl[0] = fmaf(w, g, gl[0]);
      }
    }
  }
  // block reduction in double
  __shared__ double red[SIMT_TI];
  const int nout = 1 + (ARD ? d : 1);
  const int64_t blk = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
  for (int o = 0; o < nout; ++o) {
    float v = (o == 0) ? gk : 0.f;
    if (o > 0) {
#pragma unroll
      for (int c = 0; c < NG; ++c)
        if (c == o - 1) v = gl[c];
    }
    __syncthreads();
    red[tid] = (double)v;
    __syncthreads();
    for (int sft = SIMT_TI / 2; sft > 0; sft >>= 1) {
      if (tid < sft) red[tid] += red[tid + sft];
      __syncthreads();
    }
    if (tid == 0) gout[blk * gstride + o] = red[0];
  }
}

__global__ void sum_partials_double_kernel(const double* __restrict__ in, int64_t nblk, int stride, int nout,
                                           double* __restrict__ out) {
  int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= nout) return;
  double s = 0.0;
  for (int64_t b = 0; b < nblk; ++b) s += in[b * stride + o];
  out[o] = s;
}

// tcgen05 path of the bilinear derivative: gout[block][o] = sum_{r,c} L16[r][c] * sum_split partial[split][r][c]
__global__ void bilin_dot_kernel(const float* __restrict__ partial, int nsplit, int64_t rows, int64_t rows_pad,
                                 const float* __restrict__ L16, double* __restrict__ gout, int gstride, int o) {
  __shared__ double red[256];
  double acc = 0.0;
  for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < rows * TP; e += (int64_t)gridDim.x * 256) {
    float s = 0.f;
    for (int sp = 0; sp < nsplit; ++sp) s += partial[(int64_t)sp * rows_pad * TP + e];
    acc += (double)L16[e] * (double)s;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int sft = 128; sft > 0; sft >>= 1) {
    if (threadIdx.x < sft) red[threadIdx.x] += red[threadIdx.x + sft];
    __syncthreads();
  }
  if (threadIdx.x == 0) gout[(int64_t)blockIdx.x * gstride + o] = red[0];
}

template <int KIND, bool ARD>
static int launch_bilinear(gp_plan* p, const float* L16, const float* R16, double* gout, int gstride, dim3 grid, int64_t cps) {
  const float* Z1 = p->same ? p->Z2.as<float>() + p->row_begin * p->DP : p->Z1.as<float>();
  const float* Z2 = p->Z2.as<float>();
#define GP_BL_CASE(D)                                                                                              \
  case D:                                                                                                          \
    bilinear_kernel<KIND, D, ARD><<<grid, SIMT_TI, 0, p->stream>>>(Z1, Z2, L16, R16, p->row_count, p->n2, cps,     \
                                                                   p->same ? 1 : 0, p->row_begin, p->d, gout, gstride); \
    break;
  switch (p->DP) {
    GP_BL_CASE(4) GP_BL_CASE(8) GP_BL_CASE(12) GP_BL_CASE(16) GP_BL_CASE(24) GP_BL_CASE(32) GP_BL_CASE(48) GP_BL_CASE(64)
    default:
      set_error("bilinear gradient supports d <= 64 (DP=%d)", p->DP);
      return GP_E_SHAPE;
  }
#undef GP_BL_CASE
  p->launches++;
  GP_CUDA(cudaGetLastError());
  return GP_OK;
}

}  // namespace gp

using namespace gp;

extern "C" int gp_krows(gp_plan* p, const int64_t* idx, int64_t m, float* OUT, int64_t ldo) {
  GP_REQUIRE(p && p->data_set && p->hypers_set, GP_E_STATE, "plan not ready");
  GP_REQUIRE(p->backend != GP_BACKEND_SKI, GP_E_SHAPE, "row extraction is not available for the SKI backend");
  GP_REQUIRE(p->backend != GP_BACKEND_SUM, GP_E_SHAPE, "row extraction of a kernel sum: call gp_krows on every term and add");
  GP_REQUIRE(m >= 0 && ldo >= p->n2, GP_E_SHAPE, "bad krows shape");
  if (m == 0) return GP_OK;
  const float* Z1 = p->same ? p->Z2.as<float>() + p->row_begin * p->DP : p->Z1.as<float>();
  dim3 grid((unsigned)cdiv(p->n2, 256), (unsigned)m);
  size_t sh = sizeof(float) * p->DP;
  switch (p->kind) {
    case GP_RBF: krows_kernel<GP_RBF><<<grid, 256, sh, p->stream>>>(Z1, p->Z2.as<float>(), p->DP, idx, p->row_count, p->n2, p->outputscale, p->same, p->row_begin, OUT, ldo); break;
    case GP_MATERN12: krows_kernel<GP_MATERN12><<<grid, 256, sh, p->stream>>>(Z1, p->Z2.as<float>(), p->DP, idx, p->row_count, p->n2, p->outputscale, p->same, p->row_begin, OUT, ldo); break;
    case GP_MATERN32: krows_kernel<GP_MATERN32><<<grid, 256, sh, p->stream>>>(Z1, p->Z2.as<float>(), p->DP, idx, p->row_count, p->n2, p->outputscale, p->same, p->row_begin, OUT, ldo); break;
    default: krows_kernel<GP_MATERN52><<<grid, 256, sh, p->stream>>>(Z1, p->Z2.as<float>(), p->DP, idx, p->row_count, p->n2, p->outputscale, p->same, p->row_begin, OUT, ldo); break;
  }
  p->launches++;
  GP_CUDA(cudaGetLastError());
  return GP_OK;
}

extern "C" int gp_kdiag(gp_plan* p, float* OUT) {
  GP_REQUIRE(p && p->data_set && p->hypers_set, GP_E_STATE, "plan not ready");
  GP_REQUIRE(p->backend != GP_BACKEND_SKI, GP_E_SHAPE, "the diagonal is not available for the SKI backend");
  GP_REQUIRE(p->backend != GP_BACKEND_SUM, GP_E_SHAPE, "diagonal of a kernel sum: call gp_kdiag on every term and add");
  if (p->same) {
    // stationary kernels: k(x,x) = outputscale (lazy_evaluated_kernel_tensor.py:107-133 evaluates kernel(diag=True))
    fill_kernel<<<(unsigned)cdiv(p->row_count, 256), 256, 0, p->stream>>>(OUT, p->row_count, p->outputscale);
  } else {
    GP_REQUIRE(p->n1 == p->n2, GP_E_SHAPE, "diagonal of a %lld x %lld cross-covariance is undefined (kernel(x1, x2, diag=True) needs equal sizes)",
               (long long)p->n1, (long long)p->n2);
    const unsigned g = (unsigned)cdiv(p->n1, 256);
    const float* Z1 = p->Z1.as<float>();
    const float* Z2 = p->Z2.as<float>();
    switch (p->kind) {
      case GP_RBF: kdiag_cross_kernel<GP_RBF><<<g, 256, 0, p->stream>>>(Z1, Z2, p->DP, p->n1, p->outputscale, OUT); break;
      case GP_MATERN12: kdiag_cross_kernel<GP_MATERN12><<<g, 256, 0, p->stream>>>(Z1, Z2, p->DP, p->n1, p->outputscale, OUT); break;
      case GP_MATERN32: kdiag_cross_kernel<GP_MATERN32><<<g, 256, 0, p->stream>>>(Z1, Z2, p->DP, p->n1, p->outputscale, OUT); break;
      default: kdiag_cross_kernel<GP_MATERN52><<<g, 256, 0, p->stream>>>(Z1, Z2, p->DP, p->n1, p->outputscale, OUT); break;
    }
  }
  p->launches++;
  GP_CUDA(cudaGetLastError());
  return GP_OK;
}

extern "C" int gp_bilinear_grad(gp_plan* p, const float* Lf, int64_t ldl, const float* Rt, int64_t ldr, int s,
                                double* grad_ls, double* grad_os) {
  GP_REQUIRE(p && p->data_set && p->hypers_set, GP_E_STATE, "plan not ready");
  GP_REQUIRE(p->backend != GP_BACKEND_SUM, GP_E_SHAPE, "gradients of a kernel sum: call gp_bilinear_grad on every term");
  GP_REQUIRE(s >= 1, GP_E_SHAPE, "s must be >= 1");
  const bool ard = p->ls.size() > 1;
  if (p->backend == GP_BACKEND_SKI) {
    // interpolated operator: everything happens on the grid (ski.cu); one sweep per 16 columns
    std::vector<double> tot(1 + p->d, 0.0);
    GP_CHECK(p->misc2.ensure(sizeof(float) * p->row_count * TP));
    GP_CHECK(p->misc3.ensure(sizeof(float) * p->n2 * TP));
    for (int c0 = 0; c0 < s; c0 += TP) {
      const int tc = std::min(TP, s - c0);
      GP_CHECK(to_v16(p, Lf + c0, ldl, tc, p->row_count, p->misc2.as<float>()));
      GP_CHECK(to_v16(p, Rt + c0, ldr, tc, p->n2, p->misc3.as<float>()));
      GP_CHECK(ski_bilinear(p, p->misc2.as<float>(), p->misc3.as<float>(), tot.data()));
    }
    *grad_os = tot[0];
    if (ard) {
      for (int c = 0; c < p->d; ++c) grad_ls[c] = p->outputscale * tot[1 + c] / (double)p->ls[c];
    } else {
      double sum = 0.0;
      for (int c = 0; c < p->d; ++c) sum += tot[1 + c];
      grad_ls[0] = p->outputscale * sum / (double)p->ls[0];
    }
    return GP_OK;
  }
  const int nout = 1 + (ard ? p->d : 1);
  std::vector<double> total(nout, 0.0);
  int64_t ntj = cdiv(p->n2, SIMT_TJ);
  int nsp = (int)std::min<int64_t>(ntj, std::max<int64_t>(1, (2 * p->n_sm) / std::max<int64_t>(1, cdiv(p->row_count, SIMT_TI))));
  int64_t cps = cdiv(ntj, nsp) * SIMT_TJ;
  nsp = (int)cdiv(p->n2, cps);
  dim3 grid((unsigned)cdiv(p->row_count, SIMT_TI), (unsigned)nsp);
  int64_t nblk = (int64_t)grid.x * grid.y;
  GP_CHECK(p->misc.ensure(sizeof(double) * (nblk * nout + nout)));
  GP_CHECK(p->misc2.ensure(sizeof(float) * p->row_count * TP));
  GP_CHECK(p->misc3.ensure(sizeof(float) * p->n2 * TP));
  double* gout = p->misc.as<double>();
  double* gsum = gout + nblk * nout;
  // scalar lengthscale on the tensor-core backend: sum_ij (L_i . R_j) f_ij = sum_i L_i . (F R)_i, i.e. two launches of
  // the fused K.V kernel (f = k, then f = g = l dk/dl through the derivative kinds) + a dot product with L.  ARD needs d
  // weighted sums per pair and stays on the SIMT kernel.
  const bool use_tc = !ard && p->backend == GP_BACKEND_TCGEN05;
  const int64_t rows_pad = p->rows_pad;
  const int dot_blocks = (int)std::min<int64_t>(nblk, 2 * p->n_sm);
  for (int c0 = 0; c0 < s; c0 += TP) {
    int tc = std::min(TP, s - c0);
    GP_CHECK(to_v16(p, Lf + c0, ldl, tc, p->row_count, p->misc2.as<float>()));
    GP_CHECK(to_v16(p, Rt + c0, ldr, tc, p->n2, p->misc3.as<float>()));
    if (use_tc) {
      GP_CHECK(pack_v_tiles(p, p->misc3.as<float>()));
      for (int pass = 0; pass < 2; ++pass) {
        GP_CHECK(kmv_tc_launch_kind(p, pass == 0 ? p->kind : GP_DERIV + p->kind, nullptr));
        bilin_dot_kernel<<<dot_blocks, 256, 0, p->stream>>>(p->partial.as<float>(), p->nparts, p->row_count, rows_pad,
                                                            p->misc2.as<float>(), gout, nout, pass);
        p->launches++;
      }
      GP_CUDA(cudaGetLastError());
      sum_partials_double_kernel<<<(unsigned)cdiv(nout, 64), 64, 0, p->stream>>>(gout, dot_blocks, nout, nout, gsum);
      p->launches++;
      std::vector<double> h(nout);
      GP_CUDA(cudaMemcpyAsync(h.data(), gsum, sizeof(double) * nout, cudaMemcpyDeviceToHost, p->stream));
      GP_CUDA(cudaStreamSynchronize(p->stream));
      for (int o = 0; o < nout; ++o) total[o] += h[o];
      continue;
    }
    int st;
#define GP_BL_KIND(KK)                                                                                         \
  st = ard ? launch_bilinear<KK, true>(p, p->misc2.as<float>(), p->misc3.as<float>(), gout, nout, grid, cps)   \
           : launch_bilinear<KK, false>(p, p->misc2.as<float>(), p->misc3.as<float>(), gout, nout, grid, cps);
    switch (p->kind) {
      case GP_RBF: GP_BL_KIND(GP_RBF) break;
      case GP_MATERN12: GP_BL_KIND(GP_MATERN12) break;
      case GP_MATERN32: GP_BL_KIND(GP_MATERN32) break;
      default: GP_BL_KIND(GP_MATERN52) break;
    }
#undef GP_BL_KIND
    GP_CHECK(st);
    sum_partials_double_kernel<<<(unsigned)cdiv(nout, 64), 64, 0, p->stream>>>(gout, nblk, nout, nout, gsum);
    p->launches++;
    std::vector<double> h(nout);
    GP_CUDA(cudaMemcpyAsync(h.data(), gsum, sizeof(double) * nout, cudaMemcpyDeviceToHost, p->stream));
    GP_CUDA(cudaStreamSynchronize(p->stream));
    for (int o = 0; o < nout; ++o) total[o] += h[o];
  }
  // d/d outputscale of os*k = k ; d/dl: scalar -> sum w g / l ; ARD -> sum w g dz_c^2/s / l_c ; both times os
  *grad_os = total[0];
  if (ard)
    for (int c = 0; c < p->d; ++c) grad_ls[c] = p->outputscale * total[1 + c] / (double)p->ls[c];
  else
    grad_ls[0] = p->outputscale * total[1] / (double)p->ls[0];
  return GP_OK;
}
