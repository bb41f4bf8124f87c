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
// Row-wise traversal of a tile's node block for the exchanges with the global grid block: a "row" fixes the coordinates of
// dimensions 0 .. D-2 (one division chain per row instead of one per element), the lanes cover (last-dimension node, column
// group) pairs, 8 nodes x 4 groups per pass.  fn(local node, global flat index, column group).
template <int D, typename F>
__device__ __forceinline__ void ski_block_rows(const SkiBlock<D>& b, const SkiGeom& g, int warp, int nwarps, int lane, F fn) {
  const int last = b.ext[D - 1];
  const int nrows = b.nodes / last;
  for (int row = warp; row < nrows; row += nwarps) {
    int rem = row, node0 = 0;
    int64_t idx0 = (int64_t)b.base[D - 1] * g.stride[D - 1];
#pragma unroll
    for (int i = D - 2; i >= 0; --i) {
      const int c = rem % b.ext[i];
      rem /= b.ext[i];
      node0 += c * b.pitch[i];
      idx0 += (int64_t)(b.base[i] + c) * g.stride[i];
    }
    for (int c = lane >> 2; c < last; c += 8) fn(node0 + c, idx0 + (int64_t)c * g.stride[D - 1], lane & 3);
  }
}

// neighbour q (base-4 digits, dimension 0 most significant) of the warp's current point: block-local node and weight.  The
// point's 4 D weights live one per lane (lane i * 4 + c holds w_i[c]) and are fetched with shuffles: indexing a per-thread
// array with the run-time digit would put it in local memory (3 GB of L2 traffic per product at C5 in the first version).
// fr[i] = first node of the point in dimension i relative to the block.  All 32 lanes must call this together.
template <int D>
__device__ __forceinline__ void ski_local_nnz(const int (&fr)[D], float myw, const SkiBlock<D>& b, int q, int& node, float& w) {
  node = 0;
  w = 1.f;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    const int c = (q >> (2 * (D - 1 - i))) & 3;
    node += (fr[i] + c) * b.pitch[i];
    w *= __shfl_sync(0xffffffffu, myw, i * 4 + c);
  }
}
// Interpolation data of one sorted point as the warp holds it: lane l < 4 D has weight w_{l / 4}[l % 4], lane l < D the first
// node of dimension l.  Loaded one point ahead of its use (the loads are dependent -- perm -> V row -- and a tile's points are
// visited once: without the look-ahead every point costs a full L2 / HBM round trip).
struct SkiPoint {
  float w;
  int f;
};
template <int D>
__device__ __forceinline__ SkiPoint ski_load_point(const int* __restrict__ first_s, const float* __restrict__ wts_s, int64_t p, int lane) {
  SkiPoint pt;
  pt.w = (lane < 4 * D) ? wts_s[p * (4 * D) + lane] : 0.f;
  pt.f = (lane < D) ? first_s[p * D + lane] : 0;
  return pt;
}
template <int D>
__device__ __forceinline__ void ski_point_first(const SkiPoint& pt, const SkiBlock<D>& b, int (&fr)[D]) {
#pragma unroll
  for (int i = 0; i < D; ++i) fr[i] = __shfl_sync(0xffffffffu, pt.f, i) - b.base[i];
}

// U += W^T V: one CTA (4 warps) per (tile, part), node block in shared memory.  Warp w owns column group w (4 of the 16 columns)
// of EVERY node of the block and visits every point of the tile: its lanes are 32 of the point's 4^D neighbours, all distinct
// nodes, so the accumulation is a plain shared-memory read-modify-write -- no atomics (fp32 atomicAdd on shared memory is a
// compare-and-swap loop on this architecture: ATOMS.CAST.SPIN, 250 cycles per add under contention; the version built on it
// took 2.0 ms per product at C5).
constexpr int SKI_SC_THREADS = 128;
template <int D>
__global__ void __launch_bounds__(SKI_SC_THREADS)
ski_scatter_tiled_kernel(const int* __restrict__ first_s, const float* __restrict__ wts_s, const int* __restrict__ perm,
                         const int* __restrict__ off, SkiGeom g, SkiTiles tl, int parts, const float* __restrict__ V16,
                         float* __restrict__ U) {
  constexpr int NNZ = 1 << (2 * D);
  constexpr int NIT = (NNZ + 31) / 32;
  extern __shared__ __align__(16) float blk[];   // [4 column groups][nodes][4]: a warp's 32 lanes (32 nodes, one column group) then
                                                 // spread over all banks; with [nodes][16] they hit 4 banks (16-way conflicts, 1.3 ms)
  const int tid = threadIdx.x, lane = tid & 31, cg = tid >> 5;
  // work item = (tile, part): crowded tiles (few tiles, many points: small grids in low dimension) are shared by `parts` CTAs
  for (int64_t wk = blockIdx.x; wk < (int64_t)tl.ntiles * parts; wk += gridDim.x) {
    const int tile = (int)(wk / parts), part = (int)(wk % parts);
    const int t0 = off[tile], tn = off[tile + 1] - t0;
    const int p0 = t0 + (int)((int64_t)tn * part / parts), p1 = t0 + (int)((int64_t)tn * (part + 1) / parts);
    if (p0 == p1) continue;
    const SkiBlock<D> b = ski_block_of<D>(tile, g, tl);
    __syncthreads();
    for (int e = tid; e < b.nodes * 4; e += SKI_SC_THREADS) reinterpret_cast<float4*>(blk)[e] = make_float4(0, 0, 0, 0);
    __syncthreads();
    float* mine = blk + (size_t)cg * b.nodes * 4;   // this warp's plane
    // two-deep look-ahead: point data and V row of p + 1, permutation entry of p + 2
    SkiPoint pt = ski_load_point<D>(first_s, wts_s, p0, lane);
    float4 v = reinterpret_cast<const float4*>(V16 + (int64_t)perm[p0] * TP)[cg];
    int row_n = (p0 + 1 < p1) ? perm[p0 + 1] : 0;
    for (int p = p0; p < p1; ++p) {
      SkiPoint pt_n = pt;
      float4 v_n = v;
      int row_nn = 0;
      if (p + 1 < p1) {
        pt_n = ski_load_point<D>(first_s, wts_s, p + 1, lane);
        v_n = reinterpret_cast<const float4*>(V16 + (int64_t)row_n * TP)[cg];
        if (p + 2 < p1) row_nn = perm[p + 2];
      }
      int fr[D];
      ski_point_first<D>(pt, b, fr);
      int node[NIT];
      float w[NIT];
#pragma unroll
      for (int k = 0; k < NIT; ++k) {
        const int q = lane + 32 * k;
        ski_local_nnz<D>(fr, pt.w, b, q < NNZ ? q : 0, node[k], w[k]);
        if (q >= NNZ) w[k] = 0.f;
      }
      float4 u[NIT];
#pragma unroll
      for (int k = 0; k < NIT; ++k) u[k] = *reinterpret_cast<const float4*>(mine + node[k] * 4);
#pragma unroll
      for (int k = 0; k < NIT; ++k) {
        if (w[k] != 0.f) {   // lanes of one instruction hit distinct nodes; lanes with w = 0 (padding, one-hot edge cells) stay away
          u[k].x = fmaf(w[k], v.x, u[k].x); u[k].y = fmaf(w[k], v.y, u[k].y); u[k].z = fmaf(w[k], v.z, u[k].z); u[k].w = fmaf(w[k], v.w, u[k].w);
          *reinterpret_cast<float4*>(mine + node[k] * 4) = u[k];
        }
      }
      __syncwarp();
      pt = pt_n; v = v_n; row_n = row_nn;
    }
    __syncthreads();
    ski_block_rows<D>(b, g, cg, SKI_SC_THREADS / 32, lane, [&](int node, int64_t idx, int q4) {
      const float4 x = reinterpret_cast<const float4*>(blk)[q4 * b.nodes + node];
      if (x.x != 0.f || x.y != 0.f || x.z != 0.f || x.w != 0.f) red_add_v4(U + idx * TP + q4 * 4, x);
    });
  }
}

// out[perm[p]] = sum_q w_q U[node_q]: the tile's node block is staged in shared memory once; a warp takes every 8th point, its
// lanes = 8 neighbours x 4 column groups
template <int D>
__global__ void __launch_bounds__(256)
ski_gather_tiled_kernel(const int* __restrict__ first_s, const float* __restrict__ wts_s, const int* __restrict__ perm,
                        const int* __restrict__ off, SkiGeom g, SkiTiles tl, int parts, const float* __restrict__ U,
                        float* __restrict__ out) {
  constexpr int NNZ = 1 << (2 * D);
  extern __shared__ __align__(16) float blk[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, cg = lane & 3;
  for (int64_t wk = blockIdx.x; wk < (int64_t)tl.ntiles * parts; wk += gridDim.x) {
    const int tile = (int)(wk / parts), part = (int)(wk % parts);
    const int t0 = off[tile], tn = off[tile + 1] - t0;
    const int p0 = t0 + (int)((int64_t)tn * part / parts), p1 = t0 + (int)((int64_t)tn * (part + 1) / parts);
    if (p0 == p1) continue;
    const SkiBlock<D> b = ski_block_of<D>(tile, g, tl);
    __syncthreads();
    ski_block_rows<D>(b, g, warp, 8, lane, [&](int node, int64_t idx, int q4) {
      reinterpret_cast<float4*>(blk)[node * 4 + q4] = __ldg(reinterpret_cast<const float4*>(U + idx * TP) + q4);
    });
    // first point of this warp: loaded while the block is being staged
    SkiPoint pt = {0.f, 0};
    int row = 0;
    if (p0 + warp < p1) { pt = ski_load_point<D>(first_s, wts_s, p0 + warp, lane); row = perm[p0 + warp]; }
    __syncthreads();
    for (int p = p0 + warp; p < p1; p += 8) {
      SkiPoint pt_n = pt;
      int row_n = row;
      if (p + 8 < p1) { pt_n = ski_load_point<D>(first_s, wts_s, p + 8, lane); row_n = perm[p + 8]; }
      int fr[D];
      ski_point_first<D>(pt, b, fr);
      float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll 2
      for (int k = 0; k < (NNZ + 7) / 8; ++k) {        // warp-uniform trip count: the shuffles need all lanes
        const int q = (lane >> 2) + 8 * k;
        int node;
        float w;
        ski_local_nnz<D>(fr, pt.w, b, q < NNZ ? q : 0, node, w);
        if (q >= NNZ) w = 0.f;
        const float4 u = *reinterpret_cast<const float4*>(blk + node * TP + cg * 4);
        acc.x = fmaf(w, u.x, acc.x); acc.y = fmaf(w, u.y, acc.y); acc.z = fmaf(w, u.z, acc.z); acc.w = fmaf(w, u.w, acc.w);
      }
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {   // lanes with the same column group: fixed tree => deterministic
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
        acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
      }
      if (lane < 4) reinterpret_cast<float4*>(out + (int64_t)row * TP)[cg] = acc;
      pt = pt_n; row = row_n;
    }
  }
}

// mode product: tensor viewed as [outer][G][inner] (inner includes the 16 columns): out[o][i][x] = sum_k T[i][k] in[o][k][x],
// i.e. per 64-wide slab of the flattened (outer, inner) space one [G x G] . [G x 64] product.  Runs on the tensor cores as a
// 3xTF32 product (T = T_hi + T_lo and B = B_hi + B_lo with round-to-nearest tf32 parts; T_lo B_lo, 2^-22 relative, is dropped):
// fp32-level accuracy at a small multiple of the tf32 rate, so the pass is bound by streaming the grid block, not by FMAs
// (the fp32 CUDA-core version of this kernel took 249 us per pass at G = 100, M = 10^6: 13 TFLOP/s).  Warp-level
// mma.sync.m16n8k8 is the right tool for these skinny products (M = G <= 128 rows, one 64-column slab per CTA step): there is
// no accumulator reuse across slabs for a tcgen05 / TMEM pipeline to amortise.
constexpr int SKI_MT = 64;          // slab width (positions)
constexpr int SKI_BP = SKI_MT + 8;  // pitch of the staged slab: 72 = 8 mod 32 -> conflict-free B fragments
__device__ __forceinline__ uint32_t tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// shared memory: T [GM][TPITCH] fp32 (rows and columns padded with zeros to GM = 16 ceil(G/16), GK = 8 ceil(G/8); TPITCH = GK + 4
// = 12 mod 32 for G = 100: conflict-free A fragments; split into hi / lo when a fragment is loaded), the staged slab as
// B_hi / B_lo [GK][SKI_BP] (split once while staging: seven warps read every value)
__global__ void __launch_bounds__(256)
ski_mode_kernel(const float* __restrict__ T, int G, const float* __restrict__ in, float* __restrict__ out, int64_t inner, int64_t total,
                int64_t nslab) {
  extern __shared__ __align__(16) float smm[];
  const int GM = (G + 15) & ~15, GK = (G + 7) & ~7, TP_ = GK + 4;
  float* Ts = smm;                                                         // [GM][TP_]
  uint32_t* Bh = reinterpret_cast<uint32_t*>(Ts + (size_t)GM * TP_);        // [GK][SKI_BP]
  uint32_t* Bl = Bh + (size_t)GK * SKI_BP;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int r = warp; r < GM; r += 8)
    for (int c = lane; c < TP_; c += 32) Ts[r * TP_ + c] = (r < G && c < G) ? T[r * G + c] : 0.f;
  const int gr = lane >> 2, gc = lane & 3;      // fragment coordinates
  const int nmt = GM / 16, nks = GK / 8;
  const uint32_t inner32 = (uint32_t)inner, total32 = (uint32_t)total;   // the host checks total < 2^31
  const int j4 = tid & 15, kst = tid >> 4;       // staging: this thread's float4 column of the slab, first k row
  for (int64_t slab = blockIdx.x; slab < nslab; slab += gridDim.x) {
    // slab -> 64 consecutive positions q = o * inner + x of the flattened (outer, inner) space, q < total = outer * inner
    const uint32_t q0 = (uint32_t)slab * SKI_MT;
    __syncthreads();
    {
      // 4 consecutive positions share o (inner is a multiple of 16): one division per thread and slab, then a constant stride per k
      const uint32_t q = q0 + (uint32_t)j4 * 4;
      const bool live = q < total32;
      const uint32_t o = live ? q / inner32 : 0u, x = live ? q - o * inner32 : 0u;
      const float* src = in + ((size_t)o * G) * inner + x;
      for (int k = kst; k < GK; k += 16) {
        float4 v = make_float4(0, 0, 0, 0);
        if (live && k < G) v = *reinterpret_cast<const float4*>(src + (size_t)k * inner);
        uint4 h, l;
        h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
        l.x = tf32_rna(v.x - __uint_as_float(h.x)); l.y = tf32_rna(v.y - __uint_as_float(h.y));
        l.z = tf32_rna(v.z - __uint_as_float(h.z)); l.w = tf32_rna(v.w - __uint_as_float(h.w));
        *reinterpret_cast<uint4*>(&Bh[k * SKI_BP + j4 * 4]) = h;
        *reinterpret_cast<uint4*>(&Bl[k * SKI_BP + j4 * 4]) = l;
      }
    }
    __syncthreads();
    // warp tile: 2 row tiles (32 output rows) x 4 column tiles (32 positions): every B fragment feeds two row tiles, every A
    // fragment four column tiles (one row tile x 8 column tiles per warp needed 1.5 shared-memory loads per MMA)
    const int nh = warp & 1;                       // which half of the slab's 64 positions
    for (int mp = warp >> 1; 2 * mp < nmt; mp += 4) {
      const bool two = 2 * mp + 1 < nmt;
      float acc[2][4][4];
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[m][nt][j] = 0.f;
      const float* tp = Ts + (size_t)(mp * 32 + gr) * TP_ + gc;
      for (int ks = 0; ks < nks; ++ks) {
        const int k0 = ks * 8;
        uint32_t ah[2][4], al[2][4];
#pragma unroll
        for (int m = 0; m < 2; ++m) {
          const float* t2 = tp + (size_t)(two ? m : 0) * 16 * TP_;      // a missing second row tile re-reads the first (discarded)
          const float a[4] = {t2[k0], t2[k0 + 8 * TP_], t2[k0 + 4], t2[k0 + 8 * TP_ + 4]};
#pragma unroll
          for (int j = 0; j < 4; ++j) { ah[m][j] = tf32_rna(a[j]); al[m][j] = tf32_rna(a[j] - __uint_as_float(ah[m][j])); }
        }
        const uint32_t* bh = Bh + (size_t)(k0 + gc) * SKI_BP + nh * 32 + gr;
        const uint32_t* bl = Bl + (size_t)(k0 + gc) * SKI_BP + nh * 32 + gr;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const uint32_t bh0 = bh[nt * 8], bh1 = bh[nt * 8 + 4 * SKI_BP], bl0 = bl[nt * 8], bl1 = bl[nt * 8 + 4 * SKI_BP];
#pragma unroll
          for (int m = 0; m < 2; ++m) {
            mma_tf32_16x8x8(acc[m][nt], al[m], bh0, bh1);   // small terms first
            mma_tf32_16x8x8(acc[m][nt], ah[m], bl0, bl1);
            mma_tf32_16x8x8(acc[m][nt], ah[m], bh0, bh1);
          }
        }
      }
      // C fragment: (row gr, cols 2 gc, 2 gc + 1) and (row gr + 8, same cols) of every 8-column tile
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        if (m == 1 && !two) break;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const uint32_t q = q0 + (uint32_t)(nh * 32 + nt * 8 + 2 * gc);
          if (q < total32) {
            const uint32_t o = q / inner32, x = q - o * inner32;
            const int r0 = mp * 32 + m * 16 + gr;
            float* dst = out + ((size_t)o * G + r0) * inner + x;
            if (r0 < G) *reinterpret_cast<float2*>(dst) = make_float2(acc[m][nt][0], acc[m][nt][1]);
            if (r0 + 8 < G) *reinterpret_cast<float2*>(dst + 8 * (size_t)inner) = make_float2(acc[m][nt][2], acc[m][nt][3]);
          }
        }
      }
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

static size_t ski_mode_smem(int G) {
  const size_t GM = (G + 15) & ~15, GK = (G + 7) & ~7;
  return sizeof(float) * (GM * (GK + 4) + 2 * GK * SKI_BP);
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
  // (E + 3)^d nodes x 64 B of shared memory per CTA: 4 / 23 / 22 / 40 KB -> 5 to 8 CTAs per SM hide the dependent
  // shared-memory read-modify-write chain of the scatter and the per-point loads (E = 8 at d = 3 -- 85 KB, 2 CTAs per SM -- was 3x slower)
  static const int edge_by_d[SKI_MAXD + 1] = {0, 64, 16, 4, 2};
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
// CTAs per tile: ~256 points each on average (no host read-back of the real counts: the split only balances load)
static int ski_tile_parts(const gp_plan* p) {
  const int64_t avg = p->n1 / std::max(1, p->ski->ntiles);
  return (int)std::min<int64_t>(1024, std::max<int64_t>(1, cdiv(avg, 256)));
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
  const unsigned tgrid = (unsigned)std::min<int64_t>((int64_t)tl.ntiles * parts, 16 * (int64_t)p->n_sm);
  GP_CUDA(cudaMemsetAsync(A, 0, sizeof(float) * g.M * TP, st));
  ski_scatter_tiled_kernel<D><<<tgrid, SKI_SC_THREADS, tsm, st>>>(s->first_s.as<int>(), s->wts_s.as<float>(), s->perm.as<int>(), s->tile_off.as<int>(), g, tl, parts, V16, A);
  size_t toff = 0;
  float* cur = A;
  float* nxt = B;
  for (int i = 0; i < D; ++i) {
    const int G = s->G[i];
    const int64_t inner = g.stride[i] * TP;                // elements after mode i (incl. the 16 columns)
    const int64_t total = g.M / G * TP;                    // positions of the flattened (outer, inner) space
    const int64_t nslab = cdiv(total, SKI_MT);
    const size_t sh = ski_mode_smem(G);
    GP_REQUIRE(total < ((int64_t)1 << 31), GP_E_SHAPE, "SKI: grid block too large");
    ski_mode_kernel<<<(unsigned)std::min<int64_t>(nslab, 2 * p->n_sm), 256, sh, st>>>(s->T.as<float>() + toff, G, cur, nxt, inner, total, nslab);
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
  const unsigned tgrid = (unsigned)std::min<int64_t>((int64_t)tl.ntiles * parts, 16 * (int64_t)p->n_sm);
  GP_CHECK(s->gridC.ensure(sizeof(float) * g.M * TP));
  GP_CHECK(s->gridD.ensure(sizeof(float) * g.M * TP));
  GP_CHECK(p->misc.ensure(sizeof(double) * DOT_BLOCKS * (D + 1)));
  float* A = s->gridC.as<float>();
  float* B = s->gridD.as<float>();
  double* part = p->misc.as<double>();
  GP_CUDA(cudaMemsetAsync(A, 0, sizeof(float) * g.M * TP, st));
  GP_CUDA(cudaMemsetAsync(B, 0, sizeof(float) * g.M * TP, st));
  ski_scatter_tiled_kernel<D><<<tgrid, SKI_SC_THREADS, tsm, st>>>(s->first_s.as<int>(), s->wts_s.as<float>(), s->perm.as<int>(), s->tile_off.as<int>(), g, tl, parts, L16, A);
  ski_scatter_tiled_kernel<D><<<tgrid, SKI_SC_THREADS, tsm, st>>>(s->first_s.as<int>(), s->wts_s.as<float>(), s->perm.as<int>(), s->tile_off.as<int>(), g, tl, parts, R16, B);
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
      const size_t sh = ski_mode_smem(G);
      const float* Tm = ((term == 1 + i) ? s->dT.as<float>() : s->T.as<float>()) + toff;
      float* out = bufs[i & 1];
      GP_REQUIRE(tot < ((int64_t)1 << 31), GP_E_SHAPE, "SKI: grid block too large");
      ski_mode_kernel<<<(unsigned)std::min<int64_t>(nslab, 2 * p->n_sm), 256, sh, st>>>(Tm, G, cur, out, inner, tot, nslab);
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
    GP_CUDA(cudaFuncSetAttribute(ski_mode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
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
    GP_CUDA(cudaFuncSetAttribute(ski_mode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
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
