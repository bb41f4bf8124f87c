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
// in registers by the scatter and gather kernels.  The points are bucketed once per data update by the tile of their first
// node, so that a CTA works on the (E + 3)^d grid nodes of one tile in shared memory.  A product is three passes:
//   scatter  U  = W^T V           per tile: shared-memory accumulation, then one red.global.add.v4.f32 per touched node into the
//                                 [M][16] grid block (M = prod G_i; 64 MB at 100^3)
//   modes    U' = (T_0 x ... x T_{d-1}) U   one dense [G x G] product per dimension (the Toeplitz structure saves nothing at
//                                 G = 100: an FFT of length 2G-2 costs as many flops as the direct product)
//   gather   out = W U'           per tile: node block staged in shared memory, 4^d reads per (row, column group) from there,
//                                 written as the K.V partial block the mBCG finish kernels read (outputscale / noise applied there)
// and the bilinear derivative (hyper-parameter gradients) is d + 1 sweeps of the mode products between two scatters and a dot.
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

// ---- spatial tiling of the points -------------------------------------------------------------------------------------------------
// W has 4^d non-zeros per row and neighbouring points share most of their grid nodes.  Touching the [M][16] grid block once per
// (point, node) costs 4 GB of L2 atomics / reads per product at BASELINE C5 (1.3 + 0.8 ms of a 2.3 ms product).  Instead the points
// are bucketed ONCE per data update by the tile of their first node (edge E cells per dimension); a CTA then owns a tile, keeps
// the (E + 3)^d nodes the tile's points can touch in shared memory (<= 85 KB), accumulates / reads them there, and exchanges
// each node with the global grid block once per tile: ~25x less L2 traffic.
struct SkiTiles {
  int E[SKI_MAXD];    // tile edge in cells
  int nt[SKI_MAXD];   // tiles per dimension
  int ntiles;
};

template <int D>
__device__ __forceinline__ int ski_tile_of(const int* __restrict__ fr, const SkiTiles& tl) {
  int t = 0;
#pragma unroll
  for (int i = 0; i < D; ++i) t = t * tl.nt[i] + fr[i] / tl.E[i];
  return t;
}

template <int D>
__global__ void ski_tile_count_kernel(const int* __restrict__ first, int64_t n, SkiTiles tl, int* __restrict__ cnt) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) atomicAdd(cnt + ski_tile_of<D>(first + r * D, tl), 1);
}

// exclusive scan of cnt[0..m) into off[0..m] (one CTA; m <= 2^20, once per data update); cnt is reset to 0 for the fill pass
__global__ void __launch_bounds__(1024) ski_tile_scan_kernel(int* __restrict__ cnt, int m, int* __restrict__ off) {
  __shared__ int sh[1024];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < m; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = (i < m) ? cnt[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int add = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0;
      __syncthreads();
      sh[threadIdx.x] += add;
      __syncthreads();
    }
    if (i < m) { off[i] = carry + sh[threadIdx.x] - v; cnt[i] = 0; }
    __syncthreads();
    if (threadIdx.x == 1023) carry += sh[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) off[m] = carry;
}

// bucket fill: sorted copies of the per-point interpolation data + the permutation back to the caller's row order
template <int D>
__global__ void ski_tile_fill_kernel(const int* __restrict__ first, const float* __restrict__ wts, int64_t n, SkiTiles tl,
                                     const int* __restrict__ off, int* __restrict__ cursor, int* __restrict__ perm,
                                     int* __restrict__ first_s, float* __restrict__ wts_s) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int t = ski_tile_of<D>(first + r * D, tl);
  const int64_t pos = off[t] + atomicAdd(cursor + t, 1);
  perm[pos] = (int)r;
#pragma unroll
  for (int i = 0; i < D; ++i) first_s[pos * D + i] = first[r * D + i];
#pragma unroll
  for (int i = 0; i < D * 4; ++i) wts_s[pos * D * 4 + i] = wts[r * D * 4 + i];
}

// geometry of one tile's node block: base node, extent and row-major pitch per dimension (last dimension fastest)
template <int D>
struct SkiBlock {
  int base[D], ext[D], pitch[D];
  int nodes;
};
template <int D>
__device__ __forceinline__ SkiBlock<D> ski_block_of(int tile, const SkiGeom& g, const SkiTiles& tl) {
  SkiBlock<D> b;
  int rem = tile;
#pragma unroll
  for (int i = D - 1; i >= 0; --i) {
    const int tc = rem % tl.nt[i];
    rem /= tl.nt[i];
    b.base[i] = tc * tl.E[i];
    b.ext[i] = min(tl.E[i] + 3, g.G[i] - b.base[i]);
  }
  int pch = 1;
#pragma unroll
  for (int i = D - 1; i >= 0; --i) { b.pitch[i] = pch; pch *= b.ext[i]; }
  b.nodes = pch;
  return b;
}
// block-local node -> flat index into the global grid block
template <int D>
__device__ __forceinline__ int64_t ski_block_global(const SkiBlock<D>& b, const SkiGeom& g, int node) {
  int64_t idx = 0;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    const int c = node / b.pitch[i];
    node -= c * b.pitch[i];
    idx += (int64_t)(b.base[i] + c) * g.stride[i];
  }
  return idx;
}
// neighbour p (base-4 digits, dimension 0 most significant) of a point: block-local node and weight
template <int D>
__device__ __forceinline__ void ski_local_nnz(const int* __restrict__ fr, const float* __restrict__ wr, const SkiBlock<D>& b, int p, int& node, float& w) {
  node = 0;
  w = 1.f;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    const int c = (p >> (2 * (D - 1 - i))) & 3;
    node += (fr[i] - b.base[i] + c) * b.pitch[i];
    w *= wr[i * 4 + c];
  }
}

// U += W^T V: one CTA per (tile, part), node block in shared memory; a warp takes a point, its lanes = 8 neighbours x 4 column groups
template <int D>
__global__ void __launch_bounds__(256)
ski_scatter_tiled_kernel(const int* __restrict__ first_s, const float* __restrict__ wts_s, const int* __restrict__ perm,
                         const int* __restrict__ off, SkiGeom g, SkiTiles tl, int parts, const float* __restrict__ V16,
                         float* __restrict__ U) {
  constexpr int NNZ = 1 << (2 * D);
  extern __shared__ __align__(16) float blk[];   // [nodes][16]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, cg = lane & 3;
  // work item = (tile, part): crowded tiles (few tiles, many points: small grids in low dimension) are shared by `parts` CTAs
  for (int64_t wk = blockIdx.x; wk < (int64_t)tl.ntiles * parts; wk += gridDim.x) {
    const int tile = (int)(wk / parts), part = (int)(wk % parts);
    const int t0 = off[tile], tn = off[tile + 1] - t0;
    const int p0 = t0 + (int)((int64_t)tn * part / parts), p1 = t0 + (int)((int64_t)tn * (part + 1) / parts);
    if (p0 == p1) continue;
    const SkiBlock<D> b = ski_block_of<D>(tile, g, tl);
    __syncthreads();
    for (int e = tid; e < b.nodes * 4; e += 256) reinterpret_cast<float4*>(blk)[e] = make_float4(0, 0, 0, 0);
    __syncthreads();
    for (int p = p0 + warp; p < p1; p += 8) {
      int fr[D];
      float wr[D * 4];
#pragma unroll
      for (int i = 0; i < D; ++i) fr[i] = first_s[(int64_t)p * D + i];
#pragma unroll
      for (int i = 0; i < D * 4; ++i) wr[i] = wts_s[(int64_t)p * D * 4 + i];
      const float4 v = reinterpret_cast<const float4*>(V16 + (int64_t)perm[p] * TP)[cg];
      for (int q = lane >> 2; q < NNZ; q += 8) {
        int node;
        float w;
        ski_local_nnz<D>(fr, wr, b, q, node, w);
        if (w != 0.f) {
          float* dst = blk + node * TP + cg * 4;
          atomicAdd(dst + 0, w * v.x); atomicAdd(dst + 1, w * v.y); atomicAdd(dst + 2, w * v.z); atomicAdd(dst + 3, w * v.w);
        }
      }
    }
    __syncthreads();
    for (int e = tid; e < b.nodes * 4; e += 256) {
      const float4 x = reinterpret_cast<const float4*>(blk)[e];
      if (x.x != 0.f || x.y != 0.f || x.z != 0.f || x.w != 0.f)
        red_add_v4(U + ski_block_global<D>(b, g, e >> 2) * TP + (e & 3) * 4, x);
    }
  }
}

// out[perm[p]] = sum_q w_q U[node_q]: the tile's node block is staged in shared memory once
template <int D>
__global__ void __launch_bounds__(256)
ski_gather_tiled_kernel(const int* __restrict__ first_s, const float* __restrict__ wts_s, const int* __restrict__ perm,
                        const int* __restrict__ off, SkiGeom g, SkiTiles tl, int parts, const float* __restrict__ U,
                        float* __restrict__ out) {
  constexpr int NNZ = 1 << (2 * D);
  extern __shared__ __align__(16) float blk[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, cg = lane & 3;
  // work item = (tile, part): crowded tiles (few tiles, many points: small grids in low dimension) are shared by `parts` CTAs
  for (int64_t wk = blockIdx.x; wk < (int64_t)tl.ntiles * parts; wk += gridDim.x) {
    const int tile = (int)(wk / parts), part = (int)(wk % parts);
    const int t0 = off[tile], tn = off[tile + 1] - t0;
    const int p0 = t0 + (int)((int64_t)tn * part / parts), p1 = t0 + (int)((int64_t)tn * (part + 1) / parts);
    if (p0 == p1) continue;
    const SkiBlock<D> b = ski_block_of<D>(tile, g, tl);
    __syncthreads();
    for (int e = tid; e < b.nodes * 4; e += 256)
      reinterpret_cast<float4*>(blk)[e] = __ldg(reinterpret_cast<const float4*>(U + ski_block_global<D>(b, g, e >> 2) * TP) + (e & 3));
    __syncthreads();
    for (int p = p0 + warp; p < p1; p += 8) {
      int fr[D];
      float wr[D * 4];
#pragma unroll
      for (int i = 0; i < D; ++i) fr[i] = first_s[(int64_t)p * D + i];
#pragma unroll
      for (int i = 0; i < D * 4; ++i) wr[i] = wts_s[(int64_t)p * D * 4 + i];
      float4 acc = make_float4(0, 0, 0, 0);
      for (int q = lane >> 2; q < NNZ; q += 8) {
        int node;
        float w;
        ski_local_nnz<D>(fr, wr, b, q, node, w);
        const float4 u = *reinterpret_cast<const float4*>(blk + node * TP + cg * 4);
        acc.x = fmaf(w, u.x, acc.x); acc.y = fmaf(w, u.y, acc.y); acc.z = fmaf(w, u.z, acc.z); acc.w = fmaf(w, u.w, acc.w);
      }
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {   // lanes with the same column group: fixed tree => deterministic
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
        acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
      }
      if (lane < 4) reinterpret_cast<float4*>(out + (int64_t)perm[p] * TP)[cg] = acc;
    }
  }
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
__global__ void ski_toeplitz_kernel(float* __restrict__ T, int G, float step, float inv_ls, int kind, int deriv) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= G * G) return;
  const int a = e / G, b = e % G;
  const float r = fabsf((float)(a - b)) * step * inv_ls;   // |dx| / l
  float v;
  if (kind == GP_RBF) {
    v = expf(-0.5f * r * r);
    if (deriv) v *= r * r;                                 // l dk/dl = r^2 k  (functions/rbf_covariance.py:20-29)
  } else {
    const float nu2 = (kind == GP_MATERN12) ? 1.f : (kind == GP_MATERN32 ? 3.f : 5.f);
    const float rho = sqrtf(nu2) * r;
    const float ex = expf(-rho);
    if (!deriv) v = (kind == GP_MATERN12) ? ex : (kind == GP_MATERN32 ? (1.f + rho) * ex : (1.f + rho + rho * rho * (1.f / 3.f)) * ex);
    else        v = (kind == GP_MATERN12) ? rho * ex : (kind == GP_MATERN32 ? rho * rho * ex : (1.f + rho) * rho * rho * (1.f / 3.f) * ex);
  }                                                        // l dk/dl = -rho dk/drho  (functions/matern_covariance.py:27-56)
  T[e] = v;
}

// <a, b> over n floats -> one fp64 partial per CTA (fixed order inside the CTA and on the host: reproducible)
__global__ void __launch_bounds__(256) ski_dot_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n4,
                                                      double* __restrict__ part) {
  __shared__ double sh[256];
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    const float4 x = reinterpret_cast<const float4*>(a)[i], y = reinterpret_cast<const float4*>(b)[i];
    acc += (double)x.x * y.x + (double)x.y * y.y + (double)x.z * y.z + (double)x.w * y.w;
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}

static SkiTiles ski_tiles_of(const gp_ski_state* s, int d) {
  SkiTiles tl;
  tl.ntiles = s->ntiles;
  for (int i = 0; i < SKI_MAXD; ++i) { tl.E[i] = i < d ? s->tile_edge[i] : 1; tl.nt[i] = i < d ? s->tile_num[i] : 1; }
  return tl;
}

// bucket the points by tile (counting sort on the device): once per data update
static int ski_bucket_points(gp_plan* p, const SkiGeom& g) {
  gp_ski_state* s = p->ski;
  cudaStream_t st = p->stream;
  const int d = g.d;
  const int64_t n = p->n1;
  GP_REQUIRE(n < ((int64_t)1 << 31), GP_E_SHAPE, "SKI: n too large");
  static const int edge_by_d[SKI_MAXD + 1] = {0, 256, 32, 8, 3};   // (E + 3)^d nodes x 64 B <= 85 KB of shared memory
  int64_t nt = 1;
  for (int i = 0; i < d; ++i) {
    s->tile_edge[i] = edge_by_d[d];
    s->tile_num[i] = std::max(1, (int)cdiv(g.G[i] - 3, s->tile_edge[i]));
    nt *= s->tile_num[i];
  }
  GP_REQUIRE(nt <= (1 << 20), GP_E_SHAPE, "SKI: %lld point tiles (grid too fine for d=%d)", (long long)nt, d);
  s->ntiles = (int)nt;
  const SkiTiles tl = ski_tiles_of(s, d);
  GP_CHECK(s->tile_cnt.ensure(sizeof(int) * nt));
  GP_CHECK(s->tile_off.ensure(sizeof(int) * (nt + 1)));
  GP_CHECK(s->perm.ensure(sizeof(int) * n));
  GP_CHECK(s->first_s.ensure(sizeof(int) * n * d));
  GP_CHECK(s->wts_s.ensure(sizeof(float) * n * d * 4));
  GP_CUDA(cudaMemsetAsync(s->tile_cnt.p, 0, sizeof(int) * nt, st));
  const unsigned gb = (unsigned)cdiv(n, 256);
  int* cnt = s->tile_cnt.as<int>();
  int* off = s->tile_off.as<int>();
#define GP_SKI_BUCKET(DD)                                                                                                   \
  case DD:                                                                                                                  \
    ski_tile_count_kernel<DD><<<gb, 256, 0, st>>>(s->first.as<int>(), n, tl, cnt);                                          \
    ski_tile_scan_kernel<<<1, 1024, 0, st>>>(cnt, (int)nt, off);                                                            \
    ski_tile_fill_kernel<DD><<<gb, 256, 0, st>>>(s->first.as<int>(), s->wts.as<float>(), n, tl, off, cnt, s->perm.as<int>(), \
                                                 s->first_s.as<int>(), s->wts_s.as<float>());                               \
    break;
  switch (d) {
    GP_SKI_BUCKET(1)
    GP_SKI_BUCKET(2)
    GP_SKI_BUCKET(3)
    GP_SKI_BUCKET(4)
  }
#undef GP_SKI_BUCKET
  p->launches += 3;
  GP_CUDA(cudaGetLastError());
  return GP_OK;
}

constexpr int SKI_TILE_SMEM = 96 * 1024;
template <int D>
static int ski_tiled_attrs(gp_plan* p) {
  static bool done[64] = {};
  if (!done[p->device & 63]) {
    GP_CUDA(cudaFuncSetAttribute(ski_scatter_tiled_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, SKI_TILE_SMEM));
    GP_CUDA(cudaFuncSetAttribute(ski_gather_tiled_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, SKI_TILE_SMEM));
    done[p->device & 63] = true;
  }
  return GP_OK;
}
// CTAs per tile: ~1024 points each on average (no host read-back of the real counts: the split only balances load)
static int ski_tile_parts(const gp_plan* p) {
  const int64_t avg = p->n1 / std::max(1, p->ski->ntiles);
  return (int)std::min<int64_t>(256, std::max<int64_t>(1, cdiv(avg, 1024)));
}
// shared memory of one tile's node block
static size_t ski_tile_smem(const gp_ski_state* s, int d) {
  size_t nodes = 1;
  for (int i = 0; i < d; ++i) nodes *= (size_t)std::min(s->tile_edge[i] + 3, s->G[i]);
  return nodes * TP * sizeof(float);
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
  GP_CHECK(ski_bucket_points(p, g));
  size_t toff = 0;
  for (int i = 0; i < d; ++i) toff += (size_t)s->G[i] * s->G[i];
  GP_CHECK(s->T.ensure(sizeof(float) * toff));
  GP_CHECK(s->dT.ensure(sizeof(float) * toff));   // l_i dT_i/dl_i: the factors of the hyper-parameter gradients
  toff = 0;
  for (int i = 0; i < d; ++i) {
    const float l = (p->ls.size() == 1) ? p->ls[0] : p->ls[i];
    ski_toeplitz_kernel<<<(unsigned)cdiv((int64_t)s->G[i] * s->G[i], 256), 256, 0, st>>>(s->T.as<float>() + toff, s->G[i], s->step[i], 1.f / l, p->kind, 0);
    ski_toeplitz_kernel<<<(unsigned)cdiv((int64_t)s->G[i] * s->G[i], 256), 256, 0, st>>>(s->dT.as<float>() + toff, s->G[i], s->step[i], 1.f / l, p->kind, 1);
    p->launches += 2;
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
  (void)t;
  (void)n;
  float* A = s->gridA.as<float>();
  float* B = s->gridB.as<float>();
  const SkiTiles tl = ski_tiles_of(s, D);
  const size_t tsm = ski_tile_smem(s, D);
  GP_REQUIRE(tsm <= (size_t)SKI_TILE_SMEM, GP_E_SHAPE, "SKI: tile block of %zu bytes does not fit in shared memory", tsm);
  GP_CHECK(ski_tiled_attrs<D>(p));
  const int parts = ski_tile_parts(p);
  const unsigned tgrid = (unsigned)std::min<int64_t>((int64_t)tl.ntiles * parts, 4 * (int64_t)p->n_sm);
  GP_CUDA(cudaMemsetAsync(A, 0, sizeof(float) * g.M * TP, st));
  ski_scatter_tiled_kernel<D><<<tgrid, 256, tsm, st>>>(s->first_s.as<int>(), s->wts_s.as<float>(), s->perm.as<int>(), s->tile_off.as<int>(), g, tl, parts, V16, A);
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
  ski_gather_tiled_kernel<D><<<tgrid, 256, tsm, st>>>(s->first_s.as<int>(), s->wts_s.as<float>(), s->perm.as<int>(), s->tile_off.as<int>(), g, tl, parts, cur, OUT16);
  p->launches += 2 + D;
  GP_CUDA(cudaGetLastError());
  return GP_OK;
}

// Bilinear derivative of the interpolated operator (the reference reaches it through InterpolatedLinearOperator._bilinear_derivative
// -> the grid kernel's Toeplitz columns -> autograd, kernels/grid_kernel.py:138-177):
//   sum_ij (L_i . R_j) d(K_ski)_ij / d theta = < A, dK_uu/d theta B >,   A = W^T L, B = W^T R   (grid blocks [M][16]),
//   dK_uu / d l_i = (1 / l_i) T_0 x ... x (l_i dT_i/dl_i) x ... x T_{d-1}.
// total[0] += <A, K_uu B> (d/d outputscale), total[1 + i] += <A, (.. dT_i ..) B>: d + 1 sweeps of d mode products each.
template <int D>
static int ski_bilinear_d(gp_plan* p, const float* L16, const float* R16, double* total) {
  gp_ski_state* s = p->ski;
  cudaStream_t st = p->stream;
  const int64_t n = p->n1;
  SkiGeom g;
  g.d = D;
  g.M = 1;
  for (int i = D - 1; i >= 0; --i) { g.G[i] = s->G[i]; g.lo[i] = s->lo[i]; g.step[i] = s->step[i]; g.stride[i] = g.M; g.M *= s->G[i]; }
  (void)n;
  constexpr int DOT_BLOCKS = 296;
  const SkiTiles tl = ski_tiles_of(s, D);
  const size_t tsm = ski_tile_smem(s, D);
  GP_CHECK(ski_tiled_attrs<D>(p));
  const int parts = ski_tile_parts(p);
  const unsigned tgrid = (unsigned)std::min<int64_t>((int64_t)tl.ntiles * parts, 4 * (int64_t)p->n_sm);
  GP_CHECK(s->gridC.ensure(sizeof(float) * g.M * TP));
  GP_CHECK(s->gridD.ensure(sizeof(float) * g.M * TP));
  GP_CHECK(p->misc.ensure(sizeof(double) * DOT_BLOCKS * (D + 1)));
  float* A = s->gridC.as<float>();
  float* B = s->gridD.as<float>();
  double* part = p->misc.as<double>();
  GP_CUDA(cudaMemsetAsync(A, 0, sizeof(float) * g.M * TP, st));
  GP_CUDA(cudaMemsetAsync(B, 0, sizeof(float) * g.M * TP, st));
  ski_scatter_tiled_kernel<D><<<tgrid, 256, tsm, st>>>(s->first_s.as<int>(), s->wts_s.as<float>(), s->perm.as<int>(), s->tile_off.as<int>(), g, tl, parts, L16, A);
  ski_scatter_tiled_kernel<D><<<tgrid, 256, tsm, st>>>(s->first_s.as<int>(), s->wts_s.as<float>(), s->perm.as<int>(), s->tile_off.as<int>(), g, tl, parts, R16, B);
  p->launches += 2;
  for (int term = 0; term <= D; ++term) {          // term 0: K_uu ; term 1 + i: derivative factor in dimension i
    size_t toff = 0;
    const float* cur = B;
    float* bufs[2] = {s->gridA.as<float>(), s->gridB.as<float>()};
    for (int i = 0; i < D; ++i) {
      const int G = s->G[i];
      const int64_t inner = g.stride[i] * TP;
      const int64_t tot = g.M / G * TP;
      const int64_t nslab = cdiv(tot, SKI_MT);
      const size_t sh = sizeof(float) * ((size_t)G * (G + 1) + 4 + (size_t)G * SKI_MT);
      const float* Tm = ((term == 1 + i) ? s->dT.as<float>() : s->T.as<float>()) + toff;
      float* out = bufs[i & 1];
      ski_mode_kernel<<<(unsigned)std::min<int64_t>(nslab, 8 * p->n_sm), 256, sh, st>>>(Tm, G, cur, out, inner, tot, nslab);
      cur = out;
      toff += (size_t)G * G;
    }
    ski_dot_kernel<<<DOT_BLOCKS, 256, 0, st>>>(A, cur, g.M * TP / 4, part + (size_t)term * DOT_BLOCKS);
    p->launches += D + 1;
  }
  GP_CUDA(cudaGetLastError());
  std::vector<double> h((size_t)DOT_BLOCKS * (D + 1));
  GP_CUDA(cudaMemcpyAsync(h.data(), part, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, st));
  GP_CUDA(cudaStreamSynchronize(st));
  for (int term = 0; term <= D; ++term) {
    double acc = 0.0;
    for (int b = 0; b < DOT_BLOCKS; ++b) acc += h[(size_t)term * DOT_BLOCKS + b];
    total[term] += acc;
  }
  return GP_OK;
}

int ski_bilinear(gp_plan* p, const float* L16, const float* R16, double* total) {
  static bool attr_done[64] = {};
  if (!attr_done[p->device & 63]) {
    GP_CUDA(cudaFuncSetAttribute(ski_mode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
    attr_done[p->device & 63] = true;
  }
  switch (p->d) {
    case 1: return ski_bilinear_d<1>(p, L16, R16, total);
    case 2: return ski_bilinear_d<2>(p, L16, R16, total);
    case 3: return ski_bilinear_d<3>(p, L16, R16, total);
    case 4: return ski_bilinear_d<4>(p, L16, R16, total);
  }
  set_error("SKI backend supports 1 <= d <= 4 (d=%d)", p->d);
  return GP_E_SHAPE;
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
