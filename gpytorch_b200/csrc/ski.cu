// ski.cu -- SKI / KISS-GP kernel-matmul  out = W (T_0 x T_1 x ... x T_{d-1}) W^T V  as a backend of gp_plan (SURVEY.md section 8f row 3,
// BASELINE configs[4]: N = 1e6, d = 3, grid 100^3).
//
// Replaces (reference, paths under /root/reference/gpytorch):
//   Interpolation.interpolate                       utils/interpolation.py:15-167   (Keys cubic convolution, 4 nodes per dimension,
//                                                   one-hot snapping in the first / last grid cell)
//   GridInterpolationKernel.forward / _compute_grid kernels/grid_interpolation_kernel.py:132-213
//   GridKernel.forward (Toeplitz / Kronecker K_uu)  kernels/grid_kernel.py:107-177
//   InterpolatedLinearOperator._matmul, KroneckerProductLinearOperator / ToeplitzLinearOperator._matmul   (linear_operator, absent)
// The interpolation matrix W (4^d non-zeros per row) is never stored expanded: per row and dimension the plan keeps the first
// node index and the 4 one-dimensional weights (16 d + 4 d bytes per row instead of 4^d (8 + 4)); the 4^d products are re-formed
// in registers by the scatter and gather kernels.  A product is three passes:
//   scatter  U  = W^T V           red.global.add.v4.f32 into the [M][16] grid block (M = prod G_i; L2-resident at 100^3: 64 MB)
//   modes    U' = (T_0 x ... x T_{d-1}) U   one dense [G x G] product per dimension (the Toeplitz structure saves nothing at
//                                 G = 100: an FFT of length 2G-2 costs as many flops as the direct product)
//   gather   out = W U'           64 float4 reads per (row, column group), written as the K.V partial block the mBCG finish
//                                 kernels read (outputscale and noise are applied there)
// HBM/L2-bound: algorithmic bytes per product = N 4^d (4 + 8) B as the reference stores W explicitly.
#include <math.h>

#include <algorithm>

#include "gp_common.cuh"

namespace gp {

constexpr int SKI_MAXD = 4;

struct SkiGeom {
  int d;
  int G[SKI_MAXD];
  int64_t stride[SKI_MAXD];   // flat index stride of dimension i (dimension 0 slowest, interpolation.py:157-163)
  float lo[SKI_MAXD], step[SKI_MAXD];
  int64_t M;
};

// Keys (1981) cubic convolution kernel, a = -1/2, in the reference's Horner order (utils/interpolation.py:33-43)
__device__ __forceinline__ float cubic_w(float s) {
  const float u = fabsf(s);
  const float nearv = ((1.5f * u - 2.5f) * u) * u + 1.f;
  const float farv = ((-0.5f * u + 2.5f) * u - 4.f) * u + 2.f;
  return (u < 1.f) ? nearv : farv;     // u in [0, 2]: 1 - clamp(floor(u), 0, 1) selects the branch
}

// per row and dimension: first node index and the 4 weights
__global__ void ski_interp_kernel(const float* __restrict__ X, int64_t n, int64_t ldx, SkiGeom g, int* __restrict__ first,
                                  float* __restrict__ wts, int* __restrict__ oob) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  for (int i = 0; i < g.d; ++i) {
    const float x = X[r * ldx + i];
    const float hi = g.lo[i] + g.step[i] * (float)(g.G[i] - 1);
    if (!(x - g.lo[i] >= -1e-7f) || !(x - hi <= 1e-7f)) *oob = 1;   // "Received data that was out of bounds for the specified grid."
    const float t = (x - g.lo[i]) / fmaxf(g.step[i], 1e-10f);
    const float cell = floorf(t);
    const float frac = t - cell;
    int f = (int)cell - 1;                       // left-most of the 4 nodes
    float w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) w[j] = cubic_w(frac + (float)(1 - j));   // distances f+1, f, f-1, f-2
    if (f < 0 || f > g.G[i] - 4) {
      // first / last cell: the nearest of the first / last 4 nodes gets weight 1 (interpolation.py:84-131)
      const int base = (f < 0) ? 0 : g.G[i] - 4;
      int best = 0;
      float bd = 3.4e38f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float dj = fabsf(g.lo[i] + g.step[i] * (float)(base + j) - x);
        if (dj < bd) { bd = dj; best = j; }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = (j == best) ? 1.f : 0.f;
      f = base;
    }
    first[r * g.d + i] = f;
#pragma unroll
    for (int j = 0; j < 4; ++j) wts[(r * g.d + i) * 4 + j] = w[j];
  }
}

__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// decode non-zero p (0 .. 4^d - 1; dimension 0 the most significant base-4 digit) of row r: flat grid index and weight
template <int D>
__device__ __forceinline__ void ski_nnz(const int* __restrict__ fr, const float* __restrict__ wr, const SkiGeom& g, int p, int64_t& idx, float& w) {
  idx = 0;
  w = 1.f;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    const int c = (p >> (2 * (D - 1 - i))) & 3;
    idx += (int64_t)(fr[i] + c) * g.stride[i];
    w *= wr[i * 4 + c];
  }
}

// U[idx][c] += w V[r][c]: thread = (row, non-zero); 4^D threads per row
template <int D>
__global__ void ski_scatter_kernel(const int* __restrict__ first, const float* __restrict__ wts, SkiGeom g, const float* __restrict__ V16,
                                   int64_t n, int tq /*float4 groups with data: ceil(t / 4)*/, float* __restrict__ U) {
  constexpr int NNZ = 1 << (2 * D);
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t r = e / NNZ;
  if (r >= n) return;
  const int pnz = (int)(e % NNZ);
  int64_t idx;
  float w;
  ski_nnz<D>(first + r * D, wts + r * D * 4, g, pnz, idx, w);
  if (w == 0.f) return;
  const float4* v = reinterpret_cast<const float4*>(V16 + r * TP);
  float* u = U + idx * TP;
  for (int q = 0; q < tq; ++q) {
    const float4 x = v[q];
    red_add_v4(u + 4 * q, make_float4(w * x.x, w * x.y, w * x.z, w * x.w));
  }
}

// out[r][c4] = sum_p w_p U[idx_p][c4]: thread = (row, float4 column group)
template <int D>
__global__ void ski_gather_kernel(const int* __restrict__ first, const float* __restrict__ wts, SkiGeom g, const float* __restrict__ U,
                                  int64_t n, float* __restrict__ out) {
  constexpr int NNZ = 1 << (2 * D);
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t r = e >> 2;
  if (r >= n) return;
  const int cg = (int)(e & 3);
  int fr[D];
  float wr[D * 4];
#pragma unroll
  for (int i = 0; i < D; ++i) fr[i] = first[r * D + i];
#pragma unroll
  for (int i = 0; i < D * 4; ++i) wr[i] = wts[r * D * 4 + i];
  float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll 4
  for (int pnz = 0; pnz < NNZ; ++pnz) {
    int64_t idx;
    float w;
    ski_nnz<D>(fr, wr, g, pnz, idx, w);
    const float4 u = __ldg(reinterpret_cast<const float4*>(U + idx * TP) + cg);
    acc.x = fmaf(w, u.x, acc.x); acc.y = fmaf(w, u.y, acc.y); acc.z = fmaf(w, u.z, acc.z); acc.w = fmaf(w, u.w, acc.w);
  }
  reinterpret_cast<float4*>(out)[r * 4 + cg] = acc;
}

// mode product: tensor viewed as [outer][G][inner] (inner includes the 16 columns): out[o][i][x] = sum_k T[i][k] in[o][k][x].
// One CTA = one 64-wide slab of the (outer, inner) index space; thread tile 8 rows x 4 columns (G <= 128).
constexpr int SKI_MT = 64;
__global__ void __launch_bounds__(256)
ski_mode_kernel(const float* __restrict__ T, int G, const float* __restrict__ in, float* __restrict__ out, int64_t inner, int64_t total,
                int64_t nslab) {
  extern __shared__ __align__(16) float smm[];
  float* Ts = smm;                         // [G][G + 1]
  float* Bs = smm + (size_t)G * (G + 1);   // [G][64], moved up to the next 16-byte boundary
  Bs = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(Bs) + 15) & ~(uintptr_t)15);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  for (int e = tid; e < G * G; e += 256) Ts[(e / G) * (G + 1) + (e % G)] = T[e];
  for (int64_t slab = blockIdx.x; slab < nslab; slab += gridDim.x) {
    // slab -> 64 consecutive positions q = o * inner + x of the flattened (outer, inner) space, q < total = outer * inner
    const int64_t q0 = slab * SKI_MT;
    __syncthreads();
    for (int e = tid; e < G * (SKI_MT / 4); e += 256) {
      const int k = e / (SKI_MT / 4), j4 = e % (SKI_MT / 4);
      const int64_t q = q0 + j4 * 4;       // inner is a multiple of 16, so 4 consecutive positions share o
      float4 v = make_float4(0, 0, 0, 0);
      if (q < total) {
        const int64_t o = q / inner, x = q % inner;
        v = *reinterpret_cast<const float4*>(in + (o * G + k) * inner + x);
      }
      *reinterpret_cast<float4*>(&Bs[k * SKI_MT + j4 * 4]) = v;
    }
    __syncthreads();
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int row0 = ty * 8;
    for (int k = 0; k < G; ++k) {
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k * SKI_MT + tx * 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float a = (row0 + i < G) ? Ts[(row0 + i) * (G + 1) + k] : 0.f;
        acc[i][0] = fmaf(a, b.x, acc[i][0]); acc[i][1] = fmaf(a, b.y, acc[i][1]);
        acc[i][2] = fmaf(a, b.z, acc[i][2]); acc[i][3] = fmaf(a, b.w, acc[i][3]);
      }
    }
    const int64_t q = q0 + tx * 4;
    if (q < total) {
      const int64_t o = q / inner, x = q % inner;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (row0 + i < G)
          *reinterpret_cast<float4*>(out + (o * G + row0 + i) * inner + x) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
  }
}

// T_i[a][b] = k_1d(|a - b| step_i / l_i): per-dimension factor of the grid covariance (grid_kernel.py:138-157 evaluates the base
// kernel on every dimension separately, last_dim_is_batch=True)
__global__ void ski_toeplitz_kernel(float* __restrict__ T, int G, float step, float inv_ls, int kind) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= G * G) return;
  const int a = e / G, b = e % G;
  const float r = fabsf((float)(a - b)) * step * inv_ls;   // |dx| / l
  float v;
  if (kind == GP_RBF) v = expf(-0.5f * r * r);
  else {
    const float nu2 = (kind == GP_MATERN12) ? 1.f : (kind == GP_MATERN32 ? 3.f : 5.f);
    const float rho = sqrtf(nu2) * r;
    const float ex = expf(-rho);
    v = (kind == GP_MATERN12) ? ex : (kind == GP_MATERN32 ? (1.f + rho) * ex : (1.f + rho + rho * rho * (1.f / 3.f)) * ex);
  }
  T[e] = v;
}

int ski_pack(gp_plan* p) {
  gp_ski_state* s = p->ski;
  GP_REQUIRE(s != nullptr, GP_E_STATE, "SKI grid not set");
  GP_REQUIRE(p->same && p->row_begin == 0 && p->row_count == p->n1, GP_E_SHAPE, "the SKI backend needs a square, unsharded operator");
  cudaStream_t st = p->stream;
  const int d = p->d;
  const int64_t n = p->n1;
  SkiGeom g;
  g.d = d;
  g.M = 1;
  for (int i = d - 1; i >= 0; --i) { g.G[i] = s->G[i]; g.lo[i] = s->lo[i]; g.step[i] = s->step[i]; g.stride[i] = g.M; g.M *= s->G[i]; }
  s->M = g.M;
  GP_CHECK(s->first.ensure(sizeof(int) * n * d));
  GP_CHECK(s->wts.ensure(sizeof(float) * n * d * 4));
  GP_CHECK(s->gridA.ensure(sizeof(float) * g.M * TP));
  GP_CHECK(s->gridB.ensure(sizeof(float) * g.M * TP));
  GP_CHECK(s->flag.ensure(64));
  GP_CHECK(p->mean.ensure(sizeof(float) * (d + 4)));
  p->xbad = reinterpret_cast<int*>(p->mean.as<float>() + d);
  GP_CUDA(cudaMemsetAsync(p->xbad, 0, sizeof(int), st));
  GP_CUDA(cudaMemsetAsync(s->flag.p, 0, sizeof(int), st));
  ski_interp_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(p->X1, n, p->ld1, g, s->first.as<int>(), s->wts.as<float>(), s->flag.as<int>());
  p->launches++;
  size_t toff = 0;
  for (int i = 0; i < d; ++i) toff += (size_t)s->G[i] * s->G[i];
  GP_CHECK(s->T.ensure(sizeof(float) * toff));
  toff = 0;
  for (int i = 0; i < d; ++i) {
    const float l = (p->ls.size() == 1) ? p->ls[0] : p->ls[i];
    ski_toeplitz_kernel<<<(unsigned)cdiv((int64_t)s->G[i] * s->G[i], 256), 256, 0, st>>>(s->T.as<float>() + toff, s->G[i], s->step[i], 1.f / l, p->kind);
    p->launches++;
    toff += (size_t)s->G[i] * s->G[i];
  }
  int h_oob = 0;
  GP_CUDA(cudaMemcpyAsync(&h_oob, s->flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  GP_CUDA(cudaStreamSynchronize(st));
  GP_REQUIRE(!h_oob, GP_E_SHAPE, "Received data that was out of bounds for the specified grid.");
  // geometry of the K.V partial block: one "split", rows padded like the dense backends
  p->nsplit = 1;
  p->nparts = 1;
  p->rows_pad = cdiv(p->row_count, 2 * TILE_I) * 2 * TILE_I;
  GP_CHECK(p->partial.ensure(sizeof(float) * (size_t)p->rows_pad * TP));
  GP_CUDA(cudaGetLastError());
  return GP_OK;
}

template <int D>
static int ski_matmul_d(gp_plan* p, const float* V16, int t, float* OUT16) {
  gp_ski_state* s = p->ski;
  cudaStream_t st = p->stream;
  const int64_t n = p->n1;
  SkiGeom g;
  g.d = D;
  g.M = 1;
  for (int i = D - 1; i >= 0; --i) { g.G[i] = s->G[i]; g.lo[i] = s->lo[i]; g.step[i] = s->step[i]; g.stride[i] = g.M; g.M *= s->G[i]; }
  constexpr int NNZ = 1 << (2 * D);
  float* A = s->gridA.as<float>();
  float* B = s->gridB.as<float>();
  GP_CUDA(cudaMemsetAsync(A, 0, sizeof(float) * g.M * TP, st));
  ski_scatter_kernel<D><<<(unsigned)cdiv(n * NNZ, 256), 256, 0, st>>>(s->first.as<int>(), s->wts.as<float>(), g, V16, n, (t + 3) / 4, A);
  size_t toff = 0;
  float* cur = A;
  float* nxt = B;
  for (int i = 0; i < D; ++i) {
    const int G = s->G[i];
    const int64_t inner = g.stride[i] * TP;                // elements after mode i (incl. the 16 columns)
    const int64_t total = g.M / G * TP;                    // positions of the flattened (outer, inner) space
    const int64_t nslab = cdiv(total, SKI_MT);
    const size_t sh = sizeof(float) * ((size_t)G * (G + 1) + 4 + (size_t)G * SKI_MT);
    ski_mode_kernel<<<(unsigned)std::min<int64_t>(nslab, 8 * p->n_sm), 256, sh, st>>>(s->T.as<float>() + toff, G, cur, nxt, inner, total, nslab);
    toff += (size_t)G * G;
    std::swap(cur, nxt);
  }
  ski_gather_kernel<D><<<(unsigned)cdiv(n * 4, 256), 256, 0, st>>>(s->first.as<int>(), s->wts.as<float>(), g, cur, n, OUT16);
  p->launches += 2 + D;
  GP_CUDA(cudaGetLastError());
  return GP_OK;
}

// partial[0][r][:] = (W K_uu W^T V16)[r][:]   (outputscale / noise are applied by the finish kernels)
int ski_kmv_partials(gp_plan* p, const float* V16, const int* done_flag) {
  (void)done_flag;   // the products of a finished mBCG are cheap no-ops for the dense kernels; here they simply run
  static bool attr_done[64] = {};
  if (!attr_done[p->device & 63]) {
    GP_CUDA(cudaFuncSetAttribute(ski_mode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
    attr_done[p->device & 63] = true;
  }
  switch (p->d) {
    case 1: return ski_matmul_d<1>(p, V16, TP, p->partial.as<float>());
    case 2: return ski_matmul_d<2>(p, V16, TP, p->partial.as<float>());
    case 3: return ski_matmul_d<3>(p, V16, TP, p->partial.as<float>());
    case 4: return ski_matmul_d<4>(p, V16, TP, p->partial.as<float>());
  }
  set_error("SKI backend supports 1 <= d <= 4 (d=%d)", p->d);
  return GP_E_SHAPE;
}

}  // namespace gp

using namespace gp;

// GridInterpolationKernel(base_kernel, grid_size, num_dims, grid_bounds): grid_lo = first node, grid_step = node spacing
// (utils/grid.py:142-180 create_grid: linspace(lo - step, hi + step, size) per dimension)
extern "C" int gp_plan_set_ski(gp_plan* p, const int* grid_sizes, const float* grid_lo, const float* grid_step, int d) {
  GP_REQUIRE(p != nullptr && p->data_set, GP_E_STATE, "set_data must precede set_ski");
  GP_REQUIRE(d == p->d && d >= 1 && d <= SKI_MAXD, GP_E_SHAPE, "SKI: grid dimension %d does not match the data (d=%d, max %d)", d, p->d, SKI_MAXD);
  for (int i = 0; i < d; ++i)
    GP_REQUIRE(grid_sizes[i] >= 4 && grid_sizes[i] <= 128 && grid_step[i] > 0.f, GP_E_SHAPE, "SKI: grid size %d (dim %d) must be in [4, 128]", grid_sizes[i], i);
  if (!p->ski) p->ski = new gp_ski_state();
  for (int i = 0; i < d; ++i) { p->ski->G[i] = grid_sizes[i]; p->ski->lo[i] = grid_lo[i]; p->ski->step[i] = grid_step[i]; }
  p->backend_req = GP_BACKEND_SKI;
  p->backend = GP_BACKEND_SKI;
  if (p->hypers_set) return pack_inputs(p);
  return GP_OK;
}
