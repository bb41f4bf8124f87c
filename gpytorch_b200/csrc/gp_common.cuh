// gp_common.cuh -- shared declarations for libgpbbmm (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/gp_bbmm.h"

#ifndef __CUDA_ARCH_FEAT_SM100_ALL
#if defined(__CUDA_ARCH__)
#error "libgpbbmm is written for sm_100a only: compile with -gencode arch=compute_100a,code=sm_100a"
#endif
#endif

namespace gp {

// ---- compile-time geometry ------------------------------------------------------------
constexpr int TP = 16;         // padded column count of every [N, t] block (t <= 16)
constexpr int TILE_I = 128;    // rows of K per CTA tile (UMMA M)
constexpr int TILE_J = 64;     // columns of K per pipeline step (UMMA N of GEMM1, K of GEMM2)
constexpr int KP_MAX = 128;    // max padded augmented feature width of the tcgen05 path (3d+4 <= 128)
constexpr int SIMT_TI = 128;   // rows per CTA in the SIMT kernel
constexpr int SIMT_TJ = 64;    // staged columns per step in the SIMT kernel
constexpr float LOG2E = 1.4426950408889634f;

// ---- error plumbing -----------------------------------------------------------------------
void set_error(const char* fmt, ...);
#define GP_CUDA(call)                                                                     \
  do {                                                                                    \
    cudaError_t e__ = (call);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      gp::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return GP_E_CUDA;                                                                   \
    }                                                                                     \
  } while (0)
#define GP_CHECK(st)              \
  do {                            \
    int s__ = (st);               \
    if (s__ != GP_OK) return s__; \
  } while (0)
#define GP_REQUIRE(cond, code, ...) \
  do {                              \
    if (!(cond)) {                  \
      gp::set_error(__VA_ARGS__);   \
      return (code);                \
    }                               \
  } while (0)

// grow-only device buffer owned by a plan
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return GP_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
      return GP_E_CUDA;
    }
    cap = want;
    return GP_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

// on-device scalar state of one mBCG run
struct CgState {
  double gamma[2][TP];    // residual_inner_prod, double-buffered by iteration parity
  float alpha[TP];
  float beta[TP];
  float rhs_norm[TP];
  float rnorm[TP];
  int conv[TP];           // has_converged
  int rhs_zero[TP];
  float prev_ar[TP];      // prev_alpha_reciprocal
  float prev_beta[TP];
  int update_tridiag;
  int last_tridiag_iter;
  int done;               // set once the stop rule fires; later launches become no-ops
  int iters;              // iterations executed when done was set
  int tol_reached;
  int nan_flag;
};

}  // namespace gp

// SKI / KISS-GP backend state (ski.cu): grid geometry, compact interpolation data, grid work blocks, Toeplitz factors
struct gp_ski_state {
  int G[4] = {0, 0, 0, 0};
  float lo[4] = {0, 0, 0, 0}, step[4] = {0, 0, 0, 0};
  int64_t M = 0;
  int tile_edge[4] = {0, 0, 0, 0}, tile_num[4] = {0, 0, 0, 0}, ntiles = 0;   // spatial buckets of the points (ski.cu)
  gp::DevBuf first, wts, gridA, gridB, gridC, gridD, T, dT, flag;
  gp::DevBuf perm, tile_off, tile_cnt, first_s, wts_s;                        // points sorted by tile + the permutation back
};

struct gp_comm {
  void* nccl_comm = nullptr;
  int rank = 0, world = 1;
};

struct gp_plan {
  int device = 0;
  cudaStream_t stream = nullptr;
  int n_sm = 148;
  int backend_req = GP_BACKEND_AUTO;
  int backend = GP_BACKEND_SIMT;
  int64_t launches = 0;
  // data
  const float* X1 = nullptr;
  const float* X2 = nullptr;
  int64_t n1 = 0, n2 = 0, ld1 = 0, ld2 = 0;
  int d = 0;
  bool same = false;
  int64_t row_begin = 0, row_count = 0;  // local rows of X1 (row sharding)
  bool data_set = false, hypers_set = false;
  // hypers
  int kind = GP_RBF;
  std::vector<float> ls;
  float outputscale = 1.f, noise = 0.f;
  const float* noise_diag = nullptr;  // optional per-row diagonal D [n2] (FixedNoiseGaussianLikelihood); replaces the scalar noise
  // derived geometry
  int DP = 0;      // padded feature width of the SIMT arrays
  int KP = 0;      // padded augmented width (3d+4 -> multiple of 8) of the tcgen05 tiles
  int nsplit = 1;  // column splits of the K.V work (load balance over 148 SMs)
  int nparts = 1;  // partial-sum slots written by the K.V kernel (nsplit, x2 for the tcgen05 kernel)
  int64_t ntile_i = 0, ntile_j = 0, tiles_per_split = 0;
  int64_t rows_pad = 0;  // local rows padded to the 256-row CTA block of the tcgen05 kernel: pitch of `partial`, rows of XA
  bool tc2 = false;      // second-generation tcgen05 kernel (kmv_tc2.cu); the round-1 kernel (kmv_tc.cu) serves KP > 64
  int npoly = 2;         // of every 8 ex2 evaluations, how many run as a polynomial on the FMA pipe (0, 2, 4)
  // device buffers
  int* xbad = nullptr;  // device flag: non-finite value in the packed inputs (lives behind mean[])
  gp::DevBuf mean, scale, Z1, Z2, XA, XB, V16, Vtiles, partial, out16;
  gp::DevBuf cgU, cgR, cgZ, cgP, cgV, cgPfull, red, sums, qtr, state, tmat_tmp, misc, misc2, misc3;
  gp::DevBuf pcdiag, pcperm, pcpos, pcstate, pcpart, gram, cholC;
  gp_comm* comm = nullptr;
  gp_ski_state* ski = nullptr;   // non-null: backend == GP_BACKEND_SKI
  // kernel sums (GP_BACKEND_SUM): the terms (caller-owned plans over the same rows); while the parent launches a term's K.V
  // kernel the term writes into the parent's partial slots and reads the parent's packed V tiles
  std::vector<gp_plan*> terms;
  bool sum_tc = false;            // every term runs the tcgen05 kernel (the direction block is packed once for all of them)
  bool sum_any_tc = false;        // at least one term reads the packed V tiles
  float* partial_ext = nullptr;   // set on a TERM for the duration of one launch by its parent
  float* vtiles_ext = nullptr;
  gp::DevBuf part_scale;          // [nparts] outputscale of the term that owns each partial slot
  std::vector<float> part_scale_host;
  void* pinned = nullptr;  // small pinned host scratch
  long long* tc_trace = nullptr;  // optional device buffer [256][8] for the pipeline event trace of CTA (0,0)
};

namespace gp {

// ---- launches implemented across the .cu files -------------------------------------------
int pack_inputs(gp_plan* p);                                            // pack.cu
int to_v16(gp_plan* p, const float* V, int64_t ldv, int t, int64_t n, float* V16);
int pack_v_tiles(gp_plan* p, const float* V16);                         // pack.cu (tcgen05 B operand of GEMM2)
int kmv_partials(gp_plan* p, const float* V16, const int* done_flag);   // dispatch simt / tcgen05
int kmv_tc_launch_kind(gp_plan* p, int kind, const int* done_flag);     // kind may be GP_DERIV + kind
int kmv_tc2_launch_kind(gp_plan* p, int kind, const int* done_flag);    // kmv_tc2.cu
int kmv_simt_launch(gp_plan* p, const float* V16, const int* done_flag);
int kmv_tc_launch(gp_plan* p, const int* done_flag);
int kmv_finish_user(gp_plan* p, const float* V16, float* OUT, int64_t ldo, int t, int add_noise);
int choose_geometry(gp_plan* p);
int ski_pack(gp_plan* p);                                               // ski.cu
int ski_kmv_partials(gp_plan* p, const float* V16, const int* done_flag);
int ski_bilinear(gp_plan* p, const float* L16, const float* R16, double* total);   // total[0] += <A, K_uu B>, total[1 + i] += <A, (l_i dK_uu/dl_i) B>
int sum_pack(gp_plan* p);                                               // sum.cu: geometry / buffers of a kernel-sum plan
int sum_prepare(gp_plan* p);                                            // refresh the per-slot outputscales (no-op for other plans)
int sum_kmv_launch(gp_plan* p, const float* V16, const int* done_flag); // one launch per term into the parent's partial slots
inline float* partial_ptr(gp_plan* p) { return p->partial_ext ? p->partial_ext : p->partial.as<float>(); }
inline float* vtiles_ptr(gp_plan* p) { return p->vtiles_ext ? p->vtiles_ext : p->Vtiles.as<float>(); }
// per-slot scales of the finish kernels: nullptr = one outputscale for all slots
inline const float* part_scale_ptr(gp_plan* p) { return p->backend == GP_BACKEND_SUM ? p->part_scale.as<float>() : nullptr; }
inline bool plan_is_tc(const gp_plan* p) { return p->backend == GP_BACKEND_TCGEN05 || (p->backend == GP_BACKEND_SUM && p->sum_tc); }

__host__ __device__ inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- device helpers -----------------------------------------------------------------------
#if defined(__CUDACC__)
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// round-to-nearest (ties away) to tf32: the tensor core truncates fp32 containers to their top 19 bits,
// so operands are pre-rounded and the 2-term split x = hi + lo is exact to ~2^-24 |x|.
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }

// covariance from a = -0.5 |z_i - z_j|^2 in the pre-scaled units of pack.cu:
//   RBF     z = (x - mean) sqrt(log2 e) / l      k = 2^a                (rbf_covariance.py:19)
//   Matern  z = (x - mean) sqrt(4 nu) / l        rho^2 = -a = 2 nu r^2  (matern_covariance.py:21-47)
// internal "derivative kinds": the same tile loop evaluates g = l dk/dl (scalar lengthscale) instead of k, so that
// the bilinear derivative  sum_ij (L_i . R_j) g_ij = sum_i L_i . (G R)_i  is one more fused K.V launch
// (lazy_evaluated_kernel_tensor.py:69-105, functions/rbf_covariance.py:20-29, functions/matern_covariance.py:27-56)
constexpr int GP_DERIV = 4;   // GP_DERIV + kind

template <int KIND>
__device__ __forceinline__ float dcov_from_arg(float a, float* kout);

template <int KIND>
__device__ __forceinline__ float cov_from_arg(float a) {
  if (KIND >= GP_DERIV) {
    float k;
    return dcov_from_arg<KIND - GP_DERIV>(a, &k);
  } else if (KIND == GP_RBF) {
    return ex2_approx(fminf(a, 0.f));
  } else {
    float rho = sqrt_approx(fmaxf(-a, 0.f));
    float e = ex2_approx(-LOG2E * rho);
    if (KIND == GP_MATERN12) return e;
    if (KIND == GP_MATERN32) return fmaf(rho, e, e);
    return fmaf(fmaf(rho, 0.33333334f, 1.f), rho, 1.f) * e;  // 1 + rho + rho^2/3
  }
}
// derivative factor: dk/d(log-ish) pieces for the bilinear gradient.  Returns g with
// dk/dl = g / l (scalar lengthscale):  RBF: sq*k ; M12: rho e ; M32: rho^2 e ; M52: (1+rho) rho^2/3 e
template <int KIND>
__device__ __forceinline__ float dcov_from_arg(float a, float* kout) {
  if (KIND == GP_RBF) {
    float am = fminf(a, 0.f);
    float k = ex2_approx(am);
    *kout = k;
    return (-2.f / LOG2E) * am * k;  // |dx/l|^2 = -2 a / log2e
  } else {
    float m = fmaxf(-a, 0.f);
    float rho = sqrt_approx(m);
    float e = ex2_approx(-LOG2E * rho);
    if (KIND == GP_MATERN12) { *kout = e; return rho * e; }
    if (KIND == GP_MATERN32) { *kout = fmaf(rho, e, e); return m * e; }
    *kout = fmaf(fmaf(rho, 0.33333334f, 1.f), rho, 1.f) * e;
    return (rho + 1.f) * m * 0.33333334f * e;
  }
}
#endif

}  // namespace gp
