// kmv_tc.cu -- the fused kernel-matmul  out = K(X1,X2) V  on tcgen05 tensor cores (sm_100a).
//
// Replaces (reference, paths under /root/reference/gpytorch):
//   sq_dist GEMM + exp over N^2          kernels/kernel.py:26-49, functions/rbf_covariance.py:14-19
//   Matern poly*exp passes               functions/matern_covariance.py:21-47
//   dense K @ V inside linear_cg         lazy/lazy_evaluated_kernel_tensor.py:245-276 (chunked form)
// The N x N matrix K never exists in HBM: per 128 x 96 tile it lives in TMEM only.
//
// One CTA (320 threads) owns one work unit = (128-row tile of K) x (a contiguous range of 96-column
// tiles).  Per column tile j:
//   GEMM1  S  = A_i . B_j^T            tcgen05.mma kind::tf32, M=128 N=96 K=KP (3xTF32 split operands packed
//                                      by pack.cu so that S_ij = -0.5|z_i - z_j|^2 directly), S in TMEM
//   EPI    P  = cov(S)                 two epilogue warpgroups ping-pong: tcgen05.ld -> ex2/sqrt (MUFU) ->
//                                      P_hi/P_lo (tf32 split) -> tcgen05.st back to TMEM (P_hi in place of S)
//   GEMM2  O  = P_hi [V_hi;V_lo] (N=32) + P_lo V_hi (N=16)   A operand from TMEM, B = V^T tile in smem;
//          O is a fresh accumulator per tile, folded into fp32 registers by the epilogue warps
// Operands arrive by bulk TMA (cp.async.bulk, mbarrier complete_tx) from tiles pre-packed in HBM in the exact
// UMMA K-major no-swizzle layout, through a NS-deep smem ring.  Warp roles: 0-3 epilogue WG0, 4-7 epilogue
// WG1, 8 TMA producer, 9 TMEM allocator + MMA issuer.
//
// TMEM columns (512 allocated): [0,96) S0/P0hi  [96,192) P0lo  [192,288) S1/P1hi  [288,384) P1lo  [384,416) O0  [416,448) O1
#include "gp_common.cuh"
#include "tc_ptx.cuh"

namespace gp {

using namespace ptx;

constexpr int TC_THREADS = 320;
constexpr int COL_S0 = 0, COL_PLO0 = 96, COL_STAGE = 192, COL_O = 384, COL_OSTAGE = 32;
constexpr int V_TILE_BYTES = 2 * TILE_J * TP * 4;  // [96/4][32 rows: V_hi(16) | V_lo(16)][4] = 12288
constexpr int MAX_NS = 4;

struct TcBars {
  uint64_t a_full;
  uint64_t b_full[MAX_NS];
  uint64_t b_empty[MAX_NS];
  uint64_t s_full[2];
  uint64_t p_full[2];
  uint64_t o_full[2];
  uint32_t tmem_base;
  uint32_t pad;
};

template <int KIND>
__global__ void __launch_bounds__(TC_THREADS, 1)
kmv_tc_kernel(const float* __restrict__ XA, const float* __restrict__ XB, const float* __restrict__ Vt,
              float* __restrict__ partial, int KP, int NS, int64_t ntile_j, int64_t tiles_per_split,
              int64_t rows_pad, int same, int64_t row_begin, const int* __restrict__ done_flag) {
  if (done_flag && *done_flag) return;  // CTA-uniform, before any barrier / TMEM state exists
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = (int)warp_idx_uniform();
  const int lane = threadIdx.x & 31;
  const int64_t it = blockIdx.x;
  const int split = blockIdx.y;
  const int64_t jt0 = (int64_t)split * tiles_per_split;
  const int64_t jt1 = min(ntile_j, jt0 + tiles_per_split);
  const int T = (int)max((int64_t)0, jt1 - jt0);

  const uint32_t a_bytes = (uint32_t)KP * TILE_I * 4;
  const uint32_t b_bytes = (uint32_t)KP * TILE_J * 4;
  const uint32_t stage_bytes = b_bytes + V_TILE_BYTES;
  uint8_t* sA = smem;
  uint8_t* sStage = smem + a_bytes;
  TcBars* bars = reinterpret_cast<TcBars*>(smem + a_bytes + (size_t)NS * stage_bytes);

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars->a_full), 1);
    for (int s = 0; s < MAX_NS; ++s) {
      mbar_init(smem_u32(&bars->b_full[s]), 1);
      mbar_init(smem_u32(&bars->b_empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bars->s_full[s]), 1);
      mbar_init(smem_u32(&bars->p_full[s]), 128);
      mbar_init(smem_u32(&bars->o_full[s]), 1);
    }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc(smem_u32(&bars->tmem_base), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == 8) {
    // ===================== TMA producer (one lane) =====================
    if (lane == 0 && T > 0) {
      mbar_arrive_expect_tx(smem_u32(&bars->a_full), a_bytes);
      bulk_g2s(smem_u32(sA), XA + it * (int64_t)TILE_I * KP, a_bytes, smem_u32(&bars->a_full));
      int sb = 0;
      uint32_t par = 1;
      for (int u = 0; u < T; ++u) {
        mbar_wait(smem_u32(&bars->b_empty[sb]), par);
        const uint32_t full = smem_u32(&bars->b_full[sb]);
        uint8_t* st = sStage + (size_t)sb * stage_bytes;
        const int64_t jt = jt0 + u;
        mbar_arrive_expect_tx(full, stage_bytes);
        bulk_g2s(smem_u32(st), XB + jt * (int64_t)TILE_J * KP, b_bytes, full);
        bulk_g2s(smem_u32(st + b_bytes), Vt + jt * (int64_t)(2 * TILE_J * TP), V_TILE_BYTES, full);
        if (++sb == NS) { sb = 0; par ^= 1; }
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer: the whole warp runs converged, one elected lane issues ==========
    if (T > 0) {
      constexpr uint32_t IDESC1 = idesc_tf32(TILE_I, TILE_J);   // S  = A B^T          128 x 96
      constexpr uint32_t IDESC2A = idesc_tf32(TILE_I, 2 * TP);  // O += P_hi [V_hi;V_lo]^T  128 x 32
      constexpr uint32_t IDESC2B = idesc_tf32(TILE_I, TP);      // O += P_lo V_hi^T         128 x 16
      const int ksteps1 = KP / 8;
      const uint64_t a_desc0 = smem_desc(smem_u32(sA), TILE_I * 16, 128);
      mbar_wait(smem_u32(&bars->a_full), 0);
      int sb1 = 0, sb2 = 0;       // smem ring slots of GEMM1(u) and GEMM2(u-1)
      uint32_t par1 = 0;
      for (int u = 0; u <= T; ++u) {
        if (u < T) {  // GEMM1(u): S[u&1] = A . B^T
          mbar_wait(smem_u32(&bars->b_full[sb1]), par1);
          tc_fence_after();
          const uint64_t b_desc0 = smem_desc(smem_u32(sStage + (size_t)sb1 * stage_bytes), TILE_J * 16, 128);
          const uint32_t d_s = tmem + (uint32_t)((u & 1) * COL_STAGE + COL_S0);
          for (int ks = 0; ks < ksteps1; ++ks)
            mma_tf32_ss(d_s, a_desc0 + (uint64_t)(ks * ((2 * TILE_I * 16) >> 4)), b_desc0 + (uint64_t)(ks * ((2 * TILE_J * 16) >> 4)),
                        IDESC1, ks > 0 ? 1u : 0u);
          tc_commit(smem_u32(&bars->s_full[u & 1]));
          if (++sb1 == NS) { sb1 = 0; par1 ^= 1; }
        }
        if (u >= 1) {  // GEMM2(u-1): O[v&1] = P . V   (fresh accumulator every tile, see epilogue)
          const int v = u - 1;
          mbar_wait(smem_u32(&bars->p_full[v & 1]), (v >> 1) & 1);
          tc_fence_after();
          const uint64_t v_desc0 = smem_desc(smem_u32(sStage + (size_t)sb2 * stage_bytes + b_bytes), 2 * TP * 16, 128);
          const uint32_t p_hi = tmem + (uint32_t)((v & 1) * COL_STAGE + COL_S0);
          const uint32_t p_lo = tmem + (uint32_t)((v & 1) * COL_STAGE + COL_PLO0);
          const uint32_t d_o = tmem + (uint32_t)(COL_O + (v & 1) * COL_OSTAGE);
#pragma unroll
          for (int ks = 0; ks < TILE_J / 8; ++ks)
            mma_tf32_ts(d_o, p_hi + ks * 8, v_desc0 + (uint64_t)(ks * ((2 * 2 * TP * 16) >> 4)), IDESC2A, ks > 0 ? 1u : 0u);
#pragma unroll
          for (int ks = 0; ks < TILE_J / 8; ++ks)
            mma_tf32_ts(d_o, p_lo + ks * 8, v_desc0 + (uint64_t)(ks * ((2 * 2 * TP * 16) >> 4)), IDESC2B, 1u);
          tc_commit(smem_u32(&bars->b_empty[sb2]));     // smem slot (B + V) and P[v&1] are free again
          tc_commit(smem_u32(&bars->o_full[v & 1]));    // O[v&1] holds tile v's product
          if (++sb2 == NS) sb2 = 0;
        }
      }
    }
  } else {
    // ===================== epilogue warpgroups =====================
    const int wg = warp >> 2;          // 0 or 1: handles tiles u = wg, wg+2, ...
    const int q = warp & 3;            // TMEM lane quadrant of this warp
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const int64_t gi = row_begin + it * TILE_I + q * 32 + lane;  // global row of this thread
    const uint32_t t_s = tmem + lane_off + (uint32_t)(wg * COL_STAGE + COL_S0);
    const uint32_t t_lo = tmem + lane_off + (uint32_t)(wg * COL_STAGE + COL_PLO0);
    const uint32_t t_o = tmem + lane_off + (uint32_t)(COL_O + wg * COL_OSTAGE);
    // O is flushed into fp32 registers after EVERY tile: the tensor core's accumulator truncates on each add,
    // so long TMEM accumulation chains drift (1e-4 at N = 50k); 36 adds per tile keep it at ~1e-6.
    float acc[TP];
#pragma unroll
    for (int c = 0; c < TP; ++c) acc[c] = 0.f;
    int npend = 0;  // tiles whose O has not been folded into acc yet (0 or 1)
    for (int u = wg; u < T; u += 2) {
      mbar_wait(smem_u32(&bars->s_full[wg]), (u >> 1) & 1);
      tc_fence_after();
      const int64_t jbase = (jt0 + u) * TILE_J;
      const bool diag_tile = same && (row_begin + it * TILE_I < jbase + TILE_J) && (jbase < row_begin + (it + 1) * TILE_I);
#pragma unroll 1
      for (int ch = 0; ch < TILE_J / 32; ++ch) {
        uint32_t r[32], lo[32];
        GP_TMEM_LD32(t_s + ch * 32, r);
        tmem_wait_ld();
        if (diag_tile) {
          const int cd = (int)(gi - (jbase + ch * 32));
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (c == cd) r[c] = 0u;  // a_ii = 0 exactly (kernel.py:44-45 fills the diagonal with 0)
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          float p = cov_from_arg<KIND>(__uint_as_float(r[c]));
          uint32_t hi = (__float_as_uint(p) + 0x1000u) & 0xFFFFE000u;  // RN to tf32
          lo[c] = __float_as_uint(p - __uint_as_float(hi));
          r[c] = hi;
        }
        GP_TMEM_ST32(t_s + ch * 32, r);
        GP_TMEM_ST32(t_lo + ch * 32, lo);
      }
      if (npend) {  // fold the previous tile's O (GEMM2(u-2) finished long ago) before releasing P(u)
        mbar_wait(smem_u32(&bars->o_full[wg]), ((u - 2) >> 1) & 1);
        tc_fence_after();
        uint32_t o[32];
        GP_TMEM_LD32(t_o, o);
        tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < TP; ++c) acc[c] += __uint_as_float(o[c]) + __uint_as_float(o[TP + c]);
      }
      npend = 1;
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(smem_u32(&bars->p_full[wg]));  // GEMM2(u) may now overwrite O[wg] and read P[wg]
    }
    if (npend) {
      const int ulast = wg + 2 * ((T - 1 - wg) / 2);
      mbar_wait(smem_u32(&bars->o_full[wg]), (ulast >> 1) & 1);
      tc_fence_after();
      uint32_t o[32];
      GP_TMEM_LD32(t_o, o);
      tmem_wait_ld();
#pragma unroll
      for (int c = 0; c < TP; ++c) acc[c] += __uint_as_float(o[c]) + __uint_as_float(o[TP + c]);
    }
    // each warpgroup owns one partial slot: partial[split * 2 + wg][row][16]
    const int64_t row = it * TILE_I + q * 32 + lane;
    float4* dst = reinterpret_cast<float4*>(partial + (((int64_t)split * 2 + wg) * rows_pad + row) * TP);
#pragma unroll
    for (int qq = 0; qq < 4; ++qq) dst[qq] = make_float4(acc[4 * qq], acc[4 * qq + 1], acc[4 * qq + 2], acc[4 * qq + 3]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

static int tc_smem_bytes(int KP, int* ns_out) {
  const int a_bytes = KP * TILE_I * 4;
  const int stage = KP * TILE_J * 4 + V_TILE_BYTES;
  const int budget = 220 * 1024 - a_bytes - (int)sizeof(TcBars) - 1024;
  int ns = budget / stage;
  if (ns > MAX_NS) ns = MAX_NS;
  *ns_out = ns;
  return a_bytes + ns * stage + (int)sizeof(TcBars) + 64;
}

template <int KIND>
static int launch_tc_kind(gp_plan* p, const int* done_flag) {
  int ns = 0;
  int smem_bytes = tc_smem_bytes(p->KP, &ns);
  GP_REQUIRE(ns >= 2, GP_E_SHAPE, "tcgen05 path: smem ring too small for KP=%d", p->KP);
  static bool attr_done[4] = {false, false, false, false};
  if (!attr_done[KIND]) {
    GP_CUDA(cudaFuncSetAttribute(kmv_tc_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_done[KIND] = true;
  }
  int64_t rows_pad = p->ntile_i * TILE_I;
  dim3 grid((unsigned)p->ntile_i, (unsigned)p->nsplit);
  kmv_tc_kernel<KIND><<<grid, TC_THREADS, smem_bytes, p->stream>>>(
      p->XA.as<float>(), p->XB.as<float>(), p->Vtiles.as<float>(), p->partial.as<float>(), p->KP, ns, p->ntile_j,
      p->tiles_per_split, rows_pad, p->same ? 1 : 0, p->row_begin, done_flag);
  p->launches++;
  GP_CUDA(cudaGetLastError());
  return GP_OK;
}

int kmv_tc_launch(gp_plan* p, const int* done_flag) {
  switch (p->kind) {
    case GP_RBF: return launch_tc_kind<GP_RBF>(p, done_flag);
    case GP_MATERN12: return launch_tc_kind<GP_MATERN12>(p, done_flag);
    case GP_MATERN32: return launch_tc_kind<GP_MATERN32>(p, done_flag);
    case GP_MATERN52: return launch_tc_kind<GP_MATERN52>(p, done_flag);
  }
  set_error("bad kernel kind %d", p->kind);
  return GP_E_SHAPE;
}

}  // namespace gp
