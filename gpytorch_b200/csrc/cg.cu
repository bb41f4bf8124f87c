// cg.cu -- modified batched preconditioned conjugate gradients (mBCG) on the device.
//
// Restates linear_operator.utils.linear_cg (signature attested at
// /root/reference/gpytorch/variational/ciq_variational_strategy.py:56-64) with the matmul closure
// fixed to  v -> K(X,X) v + noise v  evaluated by the fused kernels.  All vectors are [n_local][16]
// fp32 in HBM (L2 resident at the BASELINE sizes); alpha/beta/gamma, the convergence flags and the
// Lanczos tridiagonals live in a CgState struct on the device; the stop rule is evaluated on the
// device and later launches become no-ops, so the host never synchronises inside the loop.
// Column-wise dots are two-stage (per-CTA fp32 partials -> fixed-order fp64 sum): deterministic,
// and the sum is the message of the NCCL all-reduce when rows are sharded across GPUs.
#include <algorithm>
#include <dlfcn.h>

#include "gp_common.cuh"

namespace gp {

constexpr int CG_THREADS = 256;
constexpr int CG_ROWS = 64;   // rows per pass of a CTA (4 float4 column groups x 64 row lanes)
constexpr int KMAX = 128;     // max preconditioner rank handled by the fused apply

int nccl_allreduce_double(gp_comm* c, double* buf, size_t count, cudaStream_t st);   // comm.cu
int nccl_allgather_float(gp_comm* c, float* buf, size_t count_per_rank, cudaStream_t st);

__device__ __forceinline__ void block_reduce_cols(float4 acc, float* red /*[CG_ROWS][TP]*/, float* out /*[TP] global*/) {
  // thread layout: cg = tid & 3 (float4 column group), rl = tid >> 2 (row lane)
  const int tid = threadIdx.x, cg = tid & 3, rl = tid >> 2;
  reinterpret_cast<float4*>(red)[rl * 4 + cg] = acc;
  __syncthreads();
  for (int s = CG_ROWS / 2; s > 0; s >>= 1) {
    if (rl < s) {
      float4 a = reinterpret_cast<float4*>(red)[rl * 4 + cg];
      float4 b = reinterpret_cast<float4*>(red)[(rl + s) * 4 + cg];
      reinterpret_cast<float4*>(red)[rl * 4 + cg] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
    __syncthreads();
  }
  if (tid < TP) out[tid] = red[tid];
}

// sum G partial vectors of length L (fp32) into fp64, fixed order.  Block = 32 outputs x 8 partial groups.
__global__ void cg_sum_kernel(const float* __restrict__ in, int G, int L, double* __restrict__ out, const int* done) {
  if (done && *done) return;
  __shared__ double sh[8][33];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int o = blockIdx.x * 32 + lane;
  double s = 0.0;
  if (o < L)
    for (int b = grp; b < G; b += 8) s += (double)in[(size_t)b * L + o];
  sh[grp][lane] = s;
  __syncthreads();
  if (grp == 0 && o < L) {
    double t = 0.0;
#pragma unroll
    for (int g2 = 0; g2 < 8; ++g2) t += sh[g2][lane];
    out[o] = t;
  }
}

__global__ void cg_rhs_sq_kernel(const float* __restrict__ RHS, int64_t ldr, int t, int64_t n, float* __restrict__ part) {
  __shared__ __align__(16) float red[CG_ROWS * TP];
  const int tid = threadIdx.x, cg = tid & 3, rl = tid >> 2;
  float4 acc = make_float4(0, 0, 0, 0);
  for (int64_t r = (int64_t)blockIdx.x * CG_ROWS + rl; r < n; r += (int64_t)gridDim.x * CG_ROWS) {
    float v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int c = cg * 4 + q;
      v[q] = (c < t) ? RHS[r * ldr + c] : 0.f;
    }
    acc.x = fmaf(v[0], v[0], acc.x); acc.y = fmaf(v[1], v[1], acc.y);
    acc.z = fmaf(v[2], v[2], acc.z); acc.w = fmaf(v[3], v[3], acc.w);
  }
  block_reduce_cols(acc, red, part + (size_t)blockIdx.x * TP);
}

// R = rhs / |rhs| ; U = 0 ; state init ; partial rr = sum R^2
__global__ void cg_init_kernel(const float* __restrict__ RHS, int64_t ldr, int t, int64_t n, const double* __restrict__ sums,
                               float eps, float* __restrict__ U, float* __restrict__ R, CgState* __restrict__ st,
                               float* __restrict__ part) {
  __shared__ __align__(16) float red[CG_ROWS * TP];
  __shared__ float inv_norm[TP];
  const int tid = threadIdx.x, cg = tid & 3, rl = tid >> 2;
  if (tid < TP) {
    float nrm = (float)sqrt(sums[tid]);
    int zero = nrm < eps;
    if (zero) nrm = 1.f;
    inv_norm[tid] = 1.f / nrm;
    if (blockIdx.x == 0) {
      st->rhs_norm[tid] = nrm;
      st->rhs_zero[tid] = zero;
      st->conv[tid] = zero;  // |R_0| = 1 for non-zero columns, < stop_updating_after for zero ones
      st->rnorm[tid] = zero ? 0.f : 1.f;
      st->alpha[tid] = 0.f; st->beta[tid] = 0.f; st->prev_ar[tid] = 0.f; st->prev_beta[tid] = 0.f;
      st->gamma[0][tid] = 0.0; st->gamma[1][tid] = 0.0;
      if (tid == 0) { st->update_tridiag = 1; st->last_tridiag_iter = 0; st->done = 0; st->iters = 0; st->tol_reached = 0; st->nan_flag = 0; }
    }
  }
  __syncthreads();
  float4 acc = make_float4(0, 0, 0, 0);
  for (int64_t r = (int64_t)blockIdx.x * CG_ROWS + rl; r < n; r += (int64_t)gridDim.x * CG_ROWS) {
    float v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int c = cg * 4 + q;
      v[q] = (c < t) ? RHS[r * ldr + c] * inv_norm[c] : 0.f;
    }
    reinterpret_cast<float4*>(R)[r * 4 + cg] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(U)[r * 4 + cg] = make_float4(0, 0, 0, 0);
    acc.x = fmaf(v[0], v[0], acc.x); acc.y = fmaf(v[1], v[1], acc.y);
    acc.z = fmaf(v[2], v[2], acc.z); acc.w = fmaf(v[3], v[3], acc.w);
  }
  block_reduce_cols(acc, red, part + (size_t)blockIdx.x * TP);
}

// partial QtR[blk][kk*16 + c] = sum_rows W[r][kk] R[r][c]     (skinny GEMM, [k x rows] . [rows x 16])
// Register tile of 4 (kk) x 4 (c) per thread; thread = (half, kg, cg): 2 row halves x 32 kk-groups x 4 column groups.
__global__ void __launch_bounds__(CG_THREADS)
cg_qtr_kernel(const float* __restrict__ W, int k, const float* __restrict__ R, int64_t n,
              float* __restrict__ part, int L, int off, const int* done) {
  if (done && *done) return;
  extern __shared__ __align__(16) float sh[];
  const int kp = (k + 3) & ~3;   // row pitch of the staged W chunk (multiple of 4 for float4 reads)
  float* Ws = sh;                // [32][kp]
  float* Rs = sh + 32 * kp;      // [32][16]
  const int tid = threadIdx.x, cg = tid & 3, kg = (tid >> 2) & 31, half = tid >> 7;
  const bool act = kg * 4 < k;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  for (int64_t r0 = (int64_t)blockIdx.x * 32; r0 < n; r0 += (int64_t)gridDim.x * 32) {
    const int nr = (int)min((int64_t)32, n - r0);
    __syncthreads();
    for (int e = tid; e < 32 * kp; e += CG_THREADS) {
      int rr = e / kp, kk = e - rr * kp;
      Ws[e] = (rr < nr && kk < k) ? W[(r0 + rr) * k + kk] : 0.f;
    }
    for (int e = tid; e < 32 * TP; e += CG_THREADS) Rs[e] = (e / TP < nr) ? R[r0 * TP + e] : 0.f;
    __syncthreads();
    if (act) {
#pragma unroll 4
      for (int rr = half; rr < 32; rr += 2) {
        const float4 wv = *reinterpret_cast<const float4*>(&Ws[rr * kp + kg * 4]);
        const float4 rv = *reinterpret_cast<const float4*>(&Rs[rr * TP + cg * 4]);
        const float w4[4] = {wv.x, wv.y, wv.z, wv.w}, r4[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(w4[a], r4[b], acc[a][b]);
      }
    }
  }
  // combine the two row halves through shared memory (fixed order), then write the CTA partial
  __syncthreads();
  float* red = sh;  // reuse: [128][16]
  if (half == 1 && act) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) red[((kg * 4 + a) * 4 + cg) * 4 + b] = acc[a][b];
  }
  __syncthreads();
  if (half == 0 && act) {
    float* o = part + (size_t)blockIdx.x * L + off;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int kk = kg * 4 + a;
      if (kk < k) {
#pragma unroll
        for (int b = 0; b < 4; ++b) o[kk * TP + cg * 4 + b] = acc[a][b] + red[((kg * 4 + a) * 4 + cg) * 4 + b];
      }
    }
  }
}

// Z = (R - W w) / noise ; partial zr = sum Z.R      (w = all-reduced QtR, fp64 [k][16])
// thread = (row lane, 4 columns); W rows are streamed from global/L2 as float4, w sits in shared memory.
__global__ void __launch_bounds__(CG_THREADS)
cg_precond_kernel(const float* __restrict__ W, int k, const double* __restrict__ w, float inv_noise,
                  const float* __restrict__ R, float* __restrict__ Z, int64_t n, float* __restrict__ part,
                  const int* done) {
  if (done && *done) return;
  extern __shared__ __align__(16) float sh[];
  float* ws = sh;                                 // [k][16]
  float* red = sh + ((k * TP + 3) & ~3);          // [CG_ROWS][16]
  const int tid = threadIdx.x, cg = tid & 3, rl = tid >> 2;
  for (int e = tid; e < k * TP; e += CG_THREADS) ws[e] = (float)w[e];
  __syncthreads();
  const bool vec = (k & 3) == 0;
  float4 acc = make_float4(0, 0, 0, 0);
  for (int64_t r = (int64_t)blockIdx.x * CG_ROWS + rl; r < n; r += (int64_t)gridDim.x * CG_ROWS) {
    const float* wr = W + r * k;
    float4 s = make_float4(0, 0, 0, 0);
    int kk = 0;
    if (vec) {
      for (; kk < k; kk += 4) {
        const float4 wv = *reinterpret_cast<const float4*>(wr + kk);
        const float w4[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const float4 c4 = *reinterpret_cast<const float4*>(&ws[(kk + a) * TP + cg * 4]);
          s.x = fmaf(w4[a], c4.x, s.x); s.y = fmaf(w4[a], c4.y, s.y); s.z = fmaf(w4[a], c4.z, s.z); s.w = fmaf(w4[a], c4.w, s.w);
        }
      }
    } else {
      for (; kk < k; ++kk) {
        const float wv = wr[kk];
        const float4 c4 = *reinterpret_cast<const float4*>(&ws[kk * TP + cg * 4]);
        s.x = fmaf(wv, c4.x, s.x); s.y = fmaf(wv, c4.y, s.y); s.z = fmaf(wv, c4.z, s.z); s.w = fmaf(wv, c4.w, s.w);
      }
    }
    const float4 rv = reinterpret_cast<const float4*>(R)[r * 4 + cg];
    float4 z = make_float4((rv.x - s.x) * inv_noise, (rv.y - s.y) * inv_noise, (rv.z - s.z) * inv_noise, (rv.w - s.w) * inv_noise);
    reinterpret_cast<float4*>(Z)[r * 4 + cg] = z;
    acc.x = fmaf(z.x, rv.x, acc.x); acc.y = fmaf(z.y, rv.y, acc.y); acc.z = fmaf(z.z, rv.z, acc.z); acc.w = fmaf(z.w, rv.w, acc.w);
  }
  block_reduce_cols(acc, red, part + (size_t)blockIdx.x * TP);
}

// partial = sum A.B per column (A, B [n][16])
__global__ void cg_dot_kernel(const float* __restrict__ A, const float* __restrict__ B, int64_t n, float* __restrict__ part,
                              const int* done) {
  if (done && *done) return;
  __shared__ __align__(16) float red[CG_ROWS * TP];
  const int tid = threadIdx.x, cg = tid & 3, rl = tid >> 2;
  float4 acc = make_float4(0, 0, 0, 0);
  for (int64_t r = (int64_t)blockIdx.x * CG_ROWS + rl; r < n; r += (int64_t)gridDim.x * CG_ROWS) {
    float4 a = reinterpret_cast<const float4*>(A)[r * 4 + cg];
    float4 b = reinterpret_cast<const float4*>(B)[r * 4 + cg];
    acc.x = fmaf(a.x, b.x, acc.x); acc.y = fmaf(a.y, b.y, acc.y);
    acc.z = fmaf(a.z, b.z, acc.z); acc.w = fmaf(a.w, b.w, acc.w);
  }
  block_reduce_cols(acc, red, part + (size_t)blockIdx.x * TP);
}

// initial direction: P = Z ; gamma[0] = sum Z.R
__global__ void cg_initdir_kernel(const float* __restrict__ Z, float* __restrict__ P, int64_t n,
                                  const double* __restrict__ sums_zr, CgState* __restrict__ st) {
  if (blockIdx.x == 0 && threadIdx.x < TP) st->gamma[0][threadIdx.x] = sums_zr[threadIdx.x];
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n * 4; e += (int64_t)gridDim.x * blockDim.x)
    reinterpret_cast<float4*>(P)[e] = reinterpret_cast<const float4*>(Z)[e];
}

// V = os * sum_s partial + noise * P ; partial pv = sum P.V
__global__ void cg_finishv_kernel(const float* __restrict__ kpart, int nsplit, int64_t rows_pad, float os, float noise,
                                  const float* __restrict__ P, float* __restrict__ V, int64_t n, float* __restrict__ part,
                                  const int* done, const int* __restrict__ xbad) {
  if (done && *done) return;
  __shared__ __align__(16) float red[CG_ROWS * TP];
  const int tid = threadIdx.x, cg = tid & 3, rl = tid >> 2;
  const float poison = *xbad ? __int_as_float(0x7fc00000) : 0.f;  // non-finite inputs: K.V is NaN in the reference
  float4 acc = make_float4(0, 0, 0, 0);
  for (int64_t r = (int64_t)blockIdx.x * CG_ROWS + rl; r < n; r += (int64_t)gridDim.x * CG_ROWS) {
    float4 s = make_float4(poison, poison, poison, poison);
    for (int sp = 0; sp < nsplit; ++sp) {
      float4 a = reinterpret_cast<const float4*>(kpart)[((int64_t)sp * rows_pad + r) * 4 + cg];
      s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
    }
    float4 p = reinterpret_cast<const float4*>(P)[r * 4 + cg];
    float4 v = make_float4(fmaf(noise, p.x, os * s.x), fmaf(noise, p.y, os * s.y), fmaf(noise, p.z, os * s.z),
                           fmaf(noise, p.w, os * s.w));
    reinterpret_cast<float4*>(V)[r * 4 + cg] = v;
    acc.x = fmaf(p.x, v.x, acc.x); acc.y = fmaf(p.y, v.y, acc.y);
    acc.z = fmaf(p.z, v.z, acc.z); acc.w = fmaf(p.w, v.w, acc.w);
  }
  block_reduce_cols(acc, red, part + (size_t)blockIdx.x * TP);
}

// alpha = gamma / pv (guards) ; U += alpha P ; R -= alpha V ; partial rr = sum R^2
__global__ void cg_update_kernel(const double* __restrict__ sums_pv, int iter, float eps, const float* __restrict__ P,
                                 const float* __restrict__ V, float* __restrict__ U, float* __restrict__ R, int64_t n,
                                 CgState* __restrict__ st, float* __restrict__ part, int L) {
  if (st->done) return;
  __shared__ __align__(16) float red[CG_ROWS * TP];
  __shared__ float alpha_s[TP];
  const int tid = threadIdx.x, cg = tid & 3, rl = tid >> 2;
  if (tid < TP) {
    float pv = (float)sums_pv[tid];
    float gam = (float)st->gamma[iter & 1][tid];
    bool zero = pv < eps;
    float a = zero ? 0.f : gam / pv;
    if (st->conv[tid]) a = 0.f;
    alpha_s[tid] = a;
    if (blockIdx.x == 0) {
      st->alpha[tid] = a;
      if (iter == 0 && !(pv == pv)) st->nan_flag = 1;
    }
  }
  __syncthreads();
  const float4 al = reinterpret_cast<float4*>(alpha_s)[cg];
  float4 acc = make_float4(0, 0, 0, 0);
  for (int64_t r = (int64_t)blockIdx.x * CG_ROWS + rl; r < n; r += (int64_t)gridDim.x * CG_ROWS) {
    float4 p = reinterpret_cast<const float4*>(P)[r * 4 + cg];
    float4 v = reinterpret_cast<const float4*>(V)[r * 4 + cg];
    float4 u = reinterpret_cast<float4*>(U)[r * 4 + cg];
    float4 rr = reinterpret_cast<float4*>(R)[r * 4 + cg];
    u.x = fmaf(al.x, p.x, u.x); u.y = fmaf(al.y, p.y, u.y); u.z = fmaf(al.z, p.z, u.z); u.w = fmaf(al.w, p.w, u.w);
    rr.x = fmaf(-al.x, v.x, rr.x); rr.y = fmaf(-al.y, v.y, rr.y); rr.z = fmaf(-al.z, v.z, rr.z); rr.w = fmaf(-al.w, v.w, rr.w);
    reinterpret_cast<float4*>(U)[r * 4 + cg] = u;
    reinterpret_cast<float4*>(R)[r * 4 + cg] = rr;
    acc.x = fmaf(rr.x, rr.x, acc.x); acc.y = fmaf(rr.y, rr.y, acc.y);
    acc.z = fmaf(rr.z, rr.z, acc.z); acc.w = fmaf(rr.w, rr.w, acc.w);
  }
  block_reduce_cols(acc, red, part + (size_t)blockIdx.x * L);
}

// beta = gamma'/gamma ; P = Z + beta P ; convergence bookkeeping + stop rule + tridiagonal update (block 0)
__global__ void cg_dir_kernel(const double* __restrict__ sums_rr, const double* __restrict__ sums_zr, int iter, float eps,
                              float stop_after, float tol, int t, int n_tridiag, int n_tridiag_iter, int max_iter,
                              const float* __restrict__ Z, float* __restrict__ P, int64_t n, CgState* __restrict__ st,
                              float* __restrict__ TMAT, int ldt) {
  if (st->done) return;
  __shared__ float beta_s[TP];
  __shared__ float rn_s[TP];
  const int tid = threadIdx.x;
  if (tid < TP) {
    float gold = (float)st->gamma[iter & 1][tid];
    float gnew = (float)sums_zr[tid];
    bool zero = gold < eps;
    float b = zero ? 0.f : gnew / gold;
    beta_s[tid] = b;
    float rn = sqrtf((float)sums_rr[tid]);
    if (st->rhs_zero[tid]) rn = 0.f;
    rn_s[tid] = rn;
  }
  __syncthreads();
  // NOTE: block 0 mutates st->done / gamma[(iter+1)&1] / conv below; other CTAs of THIS launch only read
  // gamma[iter&1] and the pre-launch value of done, and P rows written after a stop are never read again.
  const float4 be = reinterpret_cast<float4*>(beta_s)[tid & 3];
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + tid; e < n * 4; e += (int64_t)gridDim.x * blockDim.x) {
    float4 z = reinterpret_cast<const float4*>(Z)[e];
    float4 p = reinterpret_cast<float4*>(P)[e];
    // e & 3 == tid & 3 because blockDim and gridDim*blockDim are multiples of 4
    p.x = fmaf(be.x, p.x, z.x); p.y = fmaf(be.y, p.y, z.y); p.z = fmaf(be.z, p.z, z.z); p.w = fmaf(be.w, p.w, z.w);
    reinterpret_cast<float4*>(P)[e] = p;
  }
  if (blockIdx.x == 0 && tid < 32) {
    const int c = tid;
    float rn = (c < TP) ? rn_s[c] : 0.f;
    if (c < TP) {
      st->gamma[(iter + 1) & 1][c] = sums_zr[c];
      st->beta[c] = beta_s[c];
      st->rnorm[c] = rn;
    }
    float msum = (c < t) ? rn : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) msum += __shfl_xor_sync(0xffffffffu, msum, o);
    const float mean_rn = msum / (float)t;
    const int kmin = min(10, max_iter - 1);
    const bool stop = (iter >= kmin) && (mean_rn < tol) && !(n_tridiag && iter < min(n_tridiag_iter, max_iter - 1));
    // tridiagonal update happens only when the loop does not break (linear_cg order)
    float off = 0.f;
    const bool do_tri = !stop && n_tridiag && iter < n_tridiag_iter && st->update_tridiag;
    if (do_tri && c < n_tridiag) {
      float a = st->alpha[c];
      float ar = 1.f / (a == 0.f ? 1.f : a);
      float* T = TMAT + (size_t)c * ldt * ldt;
      if (iter == 0) {
        T[0] = ar;
      } else {
        float pb = st->prev_beta[c], par = st->prev_ar[c];
        T[(size_t)iter * ldt + iter] = fmaf(pb, par, ar);
        off = sqrtf(pb) * par;
        T[(size_t)iter * ldt + iter - 1] = off;
        T[(size_t)(iter - 1) * ldt + iter] = off;
      }
      st->prev_ar[c] = ar;
      st->prev_beta[c] = beta_s[c];
    }
    float mx = (do_tri && c < n_tridiag) ? off : -1e30f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    __syncwarp();
    if (c < TP) st->conv[c] = (rn < stop_after) ? 1 : 0;
    if (c == 0) {
      if (do_tri) {
        if (iter > 0 && mx < 1e-6f) st->update_tridiag = 0;
        st->last_tridiag_iter = iter;
      }
      if (stop) {
        st->tol_reached = 1;
        st->iters = iter + 1;
        __threadfence();
        st->done = 1;
      } else if (st->nan_flag) {  // NaN in the first MVM: freeze (the host reports GP_E_NAN_MVM)
        st->iters = iter + 1;
        __threadfence();
        st->done = 1;
      } else if (iter == max_iter - 1) {
        st->iters = max_iter;
      }
    }
  }
}

// SOLVES[r][c] = U[r][c] * rhs_norm[c]
__global__ void cg_finalize_kernel(const float* __restrict__ U, const CgState* __restrict__ st, int64_t n, int t,
                                   float* __restrict__ S, int64_t lds) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * TP) return;
  int64_t r = idx / TP;
  int c = (int)(idx % TP);
  if (c < t) S[r * lds + c] = U[idx] * st->rhs_norm[c];
}

static int allreduce(gp_plan* p, double* buf, size_t count) {
  if (p->comm && p->comm->world > 1) return nccl_allreduce_double(p->comm, buf, count, p->stream);
  return GP_OK;
}

int mbcg_run(gp_plan* p, const float* RHS, int64_t ldr, int t, int n_tridiag, float tol, int max_iter,
             int max_tridiag_iter, const float* W, int k, float* SOLVES, int64_t lds, float* TMAT, int* iters_out,
             int* tridiag_size, float* resid_out) {
  GP_REQUIRE(p->data_set && p->hypers_set, GP_E_STATE, "plan not ready");
  GP_REQUIRE(p->same, GP_E_SHAPE, "mBCG needs a square operator (X2 == X1)");
  GP_REQUIRE(t >= 1 && t <= TP, GP_E_SHAPE, "mBCG handles 1..%d right-hand sides per call (t=%d)", TP, t);
  GP_REQUIRE(n_tridiag >= 0 && n_tridiag <= t, GP_E_SHAPE, "n_tridiag=%d out of range", n_tridiag);
  GP_REQUIRE(max_tridiag_iter <= max_iter, GP_E_SHAPE,
             "Getting a tridiagonalization larger than the number of CG iterations run is not possible!");
  GP_REQUIRE(W == nullptr || (k >= 1 && k <= KMAX), GP_E_SHAPE, "preconditioner rank %d not in [1,%d]", k, KMAX);
  cudaStream_t st = p->stream;
  const int64_t n = p->row_count;       // local rows
  const int64_t N = p->n2;              // global size
  if (p->comm && p->comm->world > 1) {
    // the all-gather of the direction blocks needs equal, rank-ordered shards (ncclAllGather has one count for all ranks)
    GP_REQUIRE(n * p->comm->world == N && p->row_begin == (int64_t)p->comm->rank * n, GP_E_SHAPE,
               "row-sharded mBCG needs equal contiguous shards: rank %d/%d owns [%lld,+%lld) of %lld rows (pad N to a multiple of the world size)",
               p->comm->rank, p->comm->world, (long long)p->row_begin, (long long)n, (long long)N);
  } else {
    GP_REQUIRE(n == N && p->row_begin == 0, GP_E_SHAPE, "a row shard [%lld,+%lld) of %lld rows needs a communicator (gp_plan_set_comm)",
               (long long)p->row_begin, (long long)n, (long long)N);
  }
  const float eps = 1e-10f, stop_after = 1e-10f;
  const int n_tridiag_iter = (int)std::min<int64_t>(max_tridiag_iter, N);
  const bool precond = W != nullptr;
  const int G = (int)std::min<int64_t>(cdiv(n, CG_ROWS), 2 * p->n_sm);
  const int L2 = TP + (precond ? k * TP : 0);  // length of the second reduction message (rr | QtR)

  GP_CHECK(p->cgU.ensure(sizeof(float) * n * TP));
  GP_CHECK(p->cgR.ensure(sizeof(float) * n * TP));
  GP_CHECK(p->cgV.ensure(sizeof(float) * n * TP));
  if (precond) GP_CHECK(p->cgZ.ensure(sizeof(float) * n * TP));
  const int64_t n_full = (p->comm && p->comm->world > 1) ? (int64_t)p->comm->world * n : N;
  GP_CHECK(p->cgPfull.ensure(sizeof(float) * n_full * TP));
  GP_CHECK(p->red.ensure(sizeof(float) * (size_t)G * L2));
  GP_CHECK(p->sums.ensure(sizeof(double) * (size_t)(3 * TP + L2)));
  GP_CHECK(p->state.ensure(sizeof(CgState)));
  float* U = p->cgU.as<float>();
  float* R = p->cgR.as<float>();
  float* V = p->cgV.as<float>();
  float* Z = precond ? p->cgZ.as<float>() : R;
  float* Pfull = p->cgPfull.as<float>();
  float* P = Pfull + p->row_begin * TP;
  float* red = p->red.as<float>();
  double* sums_a = p->sums.as<double>();         // [16]   rhs^2, then pv
  double* sums_zr = sums_a + TP;                 // [16]
  double* sums_b = sums_zr + TP;                 // [16 + k*16]  rr | QtR
  CgState* S = p->state.as<CgState>();
  const int* done = &S->done;
  const float inv_noise = 1.f / p->noise;
  const size_t sh_qtr = sizeof(float) * std::max<size_t>(32 * (size_t)((k + 3) & ~3) + 32 * TP, 128 * TP);
  const size_t sh_pre = sizeof(float) * ((((size_t)k * TP + 3) & ~(size_t)3) + CG_ROWS * TP);
  if (precond && sh_pre > 48 * 1024) {
    GP_CUDA(cudaFuncSetAttribute(cg_precond_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh_pre));
  }
  if (n_tridiag > 0) GP_CUDA(cudaMemsetAsync(TMAT, 0, sizeof(float) * (size_t)n_tridiag * max_tridiag_iter * max_tridiag_iter, st));
  if (n_full > N) GP_CUDA(cudaMemsetAsync(Pfull, 0, sizeof(float) * n_full * TP, st));

  // ---- init: normalise rhs, R, U, Z = M^-1 R, P = Z, gamma ----
  cg_rhs_sq_kernel<<<G, CG_THREADS, 0, st>>>(RHS, ldr, t, n, red);
  cg_sum_kernel<<<1, 256, 0, st>>>(red, G, TP, sums_a, nullptr);
  GP_CHECK(allreduce(p, sums_a, TP));
  cg_init_kernel<<<G, CG_THREADS, 0, st>>>(RHS, ldr, t, n, sums_a, eps, U, R, S, red);
  p->launches += 3;
  if (precond) {
    cg_qtr_kernel<<<G, CG_THREADS, sh_qtr, st>>>(W, k, R, n, red, L2, TP, nullptr);
    cg_sum_kernel<<<(unsigned)cdiv(L2, 32), 256, 0, st>>>(red, G, L2, sums_b, nullptr);
    GP_CHECK(allreduce(p, sums_b, L2));
    cg_precond_kernel<<<G, CG_THREADS, sh_pre, st>>>(W, k, sums_b + TP, inv_noise, R, Z, n, red, nullptr);
    cg_sum_kernel<<<1, 256, 0, st>>>(red, G, TP, sums_zr, nullptr);
    GP_CHECK(allreduce(p, sums_zr, TP));
    p->launches += 4;
  } else {
    cg_sum_kernel<<<1, 256, 0, st>>>(red, G, TP, sums_zr, nullptr);  // Z = R: gamma = sum R^2 (partials of cg_init)
    GP_CHECK(allreduce(p, sums_zr, TP));
    p->launches += 1;
  }
  cg_initdir_kernel<<<G, CG_THREADS, 0, st>>>(Z, P, n, sums_zr, S);
  p->launches += 1;
  if (p->comm && p->comm->world > 1) GP_CHECK(nccl_allgather_float(p->comm, Pfull, (size_t)n * TP, st));
  GP_CUDA(cudaGetLastError());

  // ---- iterations ----
  int* h_done = reinterpret_cast<int*>(p->pinned);  // [0..3] ring of done flags
  cudaEvent_t ev[2];
  GP_CUDA(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
  GP_CUDA(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
  const int first_stop = std::max(std::min(10, max_iter - 1), n_tridiag ? std::min(n_tridiag_iter, max_iter - 1) : 0);
  const int64_t rows_pad = p->rows_pad;
  int status = GP_OK;
  int kk = 0;
  bool finished = false;
  for (kk = 0; kk < max_iter && !finished; ++kk) {
    status = kmv_partials(p, Pfull, done);
    if (status != GP_OK) break;
    cg_finishv_kernel<<<G, CG_THREADS, 0, st>>>(p->partial.as<float>(), p->nparts, rows_pad, p->outputscale, p->noise, P, V, n, red, done, p->xbad);
    cg_sum_kernel<<<1, 256, 0, st>>>(red, G, TP, sums_a, done);
    if ((status = allreduce(p, sums_a, TP)) != GP_OK) break;
    cg_update_kernel<<<G, CG_THREADS, 0, st>>>(sums_a, kk, eps, P, V, U, R, n, S, red, L2);
    p->launches += 3;
    if (precond) {
      cg_qtr_kernel<<<G, CG_THREADS, sh_qtr, st>>>(W, k, R, n, red, L2, TP, done);
      cg_sum_kernel<<<(unsigned)cdiv(L2, 32), 256, 0, st>>>(red, G, L2, sums_b, done);
      if ((status = allreduce(p, sums_b, L2)) != GP_OK) break;
      cg_precond_kernel<<<G, CG_THREADS, sh_pre, st>>>(W, k, sums_b + TP, inv_noise, R, Z, n, red, done);
      cg_sum_kernel<<<1, 256, 0, st>>>(red, G, TP, sums_zr, done);
      if ((status = allreduce(p, sums_zr, TP)) != GP_OK) break;
      p->launches += 4;
    } else {
      cg_sum_kernel<<<1, 256, 0, st>>>(red, G, L2, sums_b, done);
      if ((status = allreduce(p, sums_b, TP)) != GP_OK) break;
      p->launches += 1;
    }
    cg_dir_kernel<<<G, CG_THREADS, 0, st>>>(sums_b, precond ? sums_zr : sums_b, kk, eps, stop_after, tol, t, n_tridiag,
                                           n_tridiag_iter, max_iter, Z, P, n, S, TMAT, max_tridiag_iter);
    p->launches += 1;
    if (p->comm && p->comm->world > 1) {
      if ((status = nccl_allgather_float(p->comm, Pfull, (size_t)n * TP, st)) != GP_OK) break;
    }
    if (kk >= first_stop) {
      // look-ahead stop check: read the flag of iteration kk after iteration kk+1 has been enqueued
      cudaMemcpyAsync(&h_done[kk & 1], &S->done, sizeof(int), cudaMemcpyDeviceToHost, st);
      cudaEventRecord(ev[kk & 1], st);
      if (kk > first_stop) {
        cudaEventSynchronize(ev[(kk - 1) & 1]);
        if (h_done[(kk - 1) & 1]) finished = true;
      }
    }
  }
  cudaError_t le = cudaGetLastError();
  if (status == GP_OK && le != cudaSuccess) {
    set_error("mBCG launch failed: %s", cudaGetErrorString(le));
    status = GP_E_CUDA;
  }
  if (status == GP_OK) {
    cg_finalize_kernel<<<(unsigned)cdiv(n * TP, 256), 256, 0, st>>>(U, S, n, t, SOLVES, lds);
    p->launches += 1;
    CgState* hs = reinterpret_cast<CgState*>(reinterpret_cast<char*>(p->pinned) + 64);
    cudaMemcpyAsync(hs, S, sizeof(CgState), cudaMemcpyDeviceToHost, st);
    cudaError_t se = cudaStreamSynchronize(st);
    if (se != cudaSuccess) {
      set_error("mBCG execution failed: %s", cudaGetErrorString(se));
      status = GP_E_CUDA;
    } else {
      if (iters_out) *iters_out = hs->done ? hs->iters : max_iter;
      if (tridiag_size) *tridiag_size = n_tridiag ? hs->last_tridiag_iter + 1 : 0;
      if (resid_out)
        for (int c = 0; c < t; ++c) resid_out[c] = hs->rnorm[c];
      if (hs->nan_flag) {
        set_error("NaNs encountered when trying to perform matrix-vector multiplication");
        status = GP_E_NAN_MVM;
      } else if (!hs->tol_reached) {
        float m = 0.f;
        for (int c = 0; c < t; ++c) m += hs->rnorm[c];
        set_error("CG terminated in %d iterations with average residual norm %g which is larger than the tolerance of %g",
                  max_iter, m / t, tol);
        status = GP_W_NOT_CONVERGED;
      }
    }
  }
  cudaEventDestroy(ev[0]);
  cudaEventDestroy(ev[1]);
  return status;
}

}  // namespace gp

extern "C" int gp_mbcg(gp_plan* plan, const float* RHS, int64_t ldr, int t, int n_tridiag, float tolerance, int max_iter,
                       int max_tridiag_iter, const float* W, int k, float* SOLVES, int64_t lds, float* TMAT,
                       int* iters_out, int* tridiag_size, float* resid_out) {
  GP_REQUIRE(plan != nullptr, GP_E_STATE, "null plan");
  return gp::mbcg_run(plan, RHS, ldr, t, n_tridiag, tolerance, max_iter, max_tridiag_iter, W, k, SOLVES, lds, TMAT,
                      iters_out, tridiag_size, resid_out);
}
