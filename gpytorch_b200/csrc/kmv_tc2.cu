// kmv_tc2.cu -- second-generation fused kernel-matmul  out = K(X1,X2) V  on tcgen05 tensor cores (sm_100a).
//
// Replaces (reference, paths under /root/reference/gpytorch): the sq_dist GEMM + exp over N^2 (kernels/kernel.py:26-49,
// functions/rbf_covariance.py:14-19), the Matern poly*exp passes (functions/matern_covariance.py:21-47) and the dense
// K @ V inside linear_cg (lazy/lazy_evaluated_kernel_tensor.py:245-276).  K never exists in HBM.
//
// What changed against kmv_tc.cu (round 1, "v16": 0.92 ms at C2, profiles/NOTES_r01.md) and why (profiles/NOTES_r02.md):
//  * One CTA per SM owns a 256-row block: epilogue warpgroup w owns the 128-row half w, both halves share ONE B / V stream, so the
//    L2 -> SMEM traffic per K tile halves (v16 streamed ~3 GB per launch).
//  * Each row half ALTERNATES between two private TMEM slots (tile u in slot u & 1) and has its OWN issuer warp: GEMM2(u-1) +
//    GEMM1(u+1) of the other slot run underneath the epilogue of tile u.  The ~1000 cycles per tile that v16 lost were the single
//    issuer warp's scalar instruction stream, not the tensor pipe: running descriptors, a looped GEMM1, polling waits and the
//    issuers as the oldest warps of their sub-partitions brought the turn-around down to the epilogue time of a tile.
//  * No software pipelining across tiles in the epilogue: S(u+1) is requested only after P(u) is released (prefetching it
//    lengthens the release -> s_full turn-around it competes with; measured).
//  * GEMM1 takes its A operand from TMEM (TS mode): the 128 x KP tile of a row half is constant for the whole CTA, SS mode
//    re-read it from shared memory for every tile (48 cycles per MMA, smem-bandwidth bound; TS: 32).  KP <= 48 only; wider
//    feature vectors (KP <= 64) keep A in shared memory.
//  * 0, 2 or 4 of every 8 ex2 evaluations (template NPOLY) can run as a degree-5 polynomial on the FMA pipe (Cody-Waite split
//    with the magic-number rounding trick, packed f32x2 arithmetic; max relative error 1.8e-7, MUFU.EX2: ~1.7e-7): default 2 for
//    Matern (sqrt + ex2 per entry), 0 for RBF (not MUFU-bound any more: the extra issue slots cost more than they save).
//  * O is a single 16-column accumulator per row half (V_hi and V_lo products accumulate into the same columns), folded into
//    fp32 registers once per tile, late in the NEXT tile, after waiting for o_full of the tile it belongs to.
//
// TMEM columns (512, one CTA per SM):
//   slot (w, b) at (2w + b) * 96: S / P_hi [0,64) + P_lo (bf16 pairs) [64,96)        -> [0, 384)
//   O(w) at 384 + 16 w                                                               -> [384, 416)
//   A(w) at 416 + 48 w  (TS mode, KP <= 48)                                          -> [416, 512)
// Warp roles (640 threads): 0 / 1 MMA issuers of row half 0 / 1 (warp 0 also allocates TMEM), 2 TMA producer, 3 trace observer
// (trace builds only), 4-19 epilogue: (warp - 4) >> 3 = row half, ((warp - 4) >> 2) & 1 = column half of the 64-column tile,
// warp & 3 = TMEM lane quadrant.
#include "gp_common.cuh"
#include "tc_ptx.cuh"

namespace gp {

using namespace ptx;

namespace v2 {

constexpr int THREADS = 640;
constexpr int W_MMA = 0, W_PROD = 2, W_EPI = 4;   // warps 0 / 1 MMA issuers (0 also allocates TMEM), 2 TMA producer, 3 idle, 4-19 epilogue.
// The issuers are the OLDEST warps of their sub-partitions: with them at the highest warp ids the epilogue warps won the issue
// arbitration most of the time and one tile's ~110 issuer instructions took ~1200 cycles (profiles/NOTES_r02.md)
constexpr int SLOT_COLS = 96;                  // S / P_hi 64 + P_lo 32
constexpr int COL_O = 4 * SLOT_COLS;           // 384
constexpr int COL_A = COL_O + 2 * TP;          // 416
constexpr int A_COLS_MAX = 48;                 // per row half
constexpr int TMEM_COLS = 512;
static_assert(COL_A + 2 * A_COLS_MAX <= TMEM_COLS, "TMEM budget");
constexpr int V_TF32_BYTES = 2 * TILE_J * TP * 4;  // [64/4][32 rows: V_hi(16) | V_lo(16)][4 tf32] = 8192
constexpr int V_BF16_BYTES = TILE_J * TP * 2;      // [64/8][16 rows][8 bf16]                       = 2048
constexpr int V_TILE_BYTES = V_TF32_BYTES + V_BF16_BYTES;
constexpr int MAX_NS = 8;
constexpr int ROWS_CTA = 2 * TILE_I;           // 256

struct Bars {
  uint64_t a_full;             // SS mode: both A tiles landed in smem (TMA) ; TS mode: 256 epilogue threads stored A into TMEM
  uint64_t b_full[MAX_NS];
  uint64_t b_empty[MAX_NS];    // 2 arrivals: GEMM2(0,u) and GEMM2(1,u) have read the stage
  uint64_t s_full[2][2];       // [warpgroup][slot]
  uint64_t p_full[2][2];       // 256 arrivals (8 warps)
  uint64_t o_full[2];          // [warpgroup]
  uint32_t tmem_base;
  uint32_t pad;
};

// ---- ex2 on the FMA pipe ----------------------------------------------------------------------------------------------
// 2^x = 2^n * 2^f, n = rint(x) via the magic constant 1.5 * 2^23, f = x - n in [-0.5, 0.5], 2^f by a degree-5 minimax
// polynomial (max relative error 1.8e-7 in fp32 Horner form, tools/exp2_poly_fit.py), 2^n by adding n to the exponent field.
// Two elements per call in packed f32x2 arithmetic (FADD2 / FFMA2: half the issue slots).  x is clamped at -126 (2^-126
// is the smallest normal; the MUFU path flushes below that, the difference is < 1.2e-38 absolute).
__device__ __forceinline__ uint64_t pk(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ void ex2_poly2(float x0, float x1, float& y0, float& y1) {
  constexpr float MAGIC = 12582912.f;  // 1.5 * 2^23
  x0 = fmaxf(x0, -126.f);
  x1 = fmaxf(x1, -126.f);
  const uint64_t X = pk(x0, x1), M = pk(MAGIC, MAGIC);
  const uint64_t T = add2(X, M);       // low mantissa bits of T = n (two's complement)
  const uint64_t F = sub2(X, sub2(T, M));
  uint64_t P = pk(0.0013281626161187887f, 0.0013281626161187887f);
  P = fma2(P, F, pk(0.009675584733486176f, 0.009675584733486176f));
  P = fma2(P, F, pk(0.05550697445869446f, 0.05550697445869446f));
  P = fma2(P, F, pk(0.24022118747234344f, 0.24022118747234344f));
  P = fma2(P, F, pk(0.6931470036506653f, 0.6931470036506653f));
  P = fma2(P, F, pk(1.0000001192092896f, 1.0000001192092896f));
  float p0, p1, t0, t1;
  upk(P, p0, p1);
  upk(T, t0, t1);
  y0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  y1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

// covariance of one element with the ex2 on the MUFU (POLY = false) -- identical to cov_from_arg of gp_common.cuh except
// that the RBF exponent is not clamped at 0 (see kmv_tc.cu: a > 0 only through rounding, < 2e-6) -- and the pieces the
// polynomial variant needs: the argument of the ex2 and the factor that multiplies it.
template <int KIND>
__device__ __forceinline__ void cov_split(float a, float& earg, float& fac) {
  constexpr int K = (KIND >= GP_DERIV) ? KIND - GP_DERIV : KIND;
  constexpr bool D = KIND >= GP_DERIV;
  if (K == GP_RBF) {
    earg = a;
    fac = D ? (-2.f / LOG2E) * a : 1.f;                 // g = |dx/l|^2 k
  } else {
    const float m = fmaxf(-a, 0.f);
    const float rho = sqrt_approx(m);
    earg = -LOG2E * rho;
    if (K == GP_MATERN12) fac = D ? rho : 1.f;
    else if (K == GP_MATERN32) fac = D ? m : rho + 1.f;
    else fac = D ? (rho + 1.f) * m * 0.33333334f : fmaf(fmaf(rho, 0.33333334f, 1.f), rho, 1.f);
  }
}
template <int KIND>
__device__ __forceinline__ float cov_mufu(float a) {
  if (KIND == GP_RBF) return ex2_approx(a);
  float earg, fac;
  cov_split<KIND>(a, earg, fac);
  return fac * ex2_approx(earg);
}
// P = cov(S) for 8 columns; NPOLY of them (0, 2 or 4: columns 3,7 / 1,3,5,7) take their ex2 from the FMA pipe
template <int KIND, int NPOLY>
__device__ __forceinline__ void cov_group8(const uint32_t* __restrict__ s, float (&p)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const bool poly = (NPOLY == 4) ? (i & 1) : (NPOLY == 2 ? (i & 3) == 3 : false);
    if (!poly) p[i] = cov_mufu<KIND>(__uint_as_float(s[i]));
  }
  if (NPOLY >= 2) {
#pragma unroll
    for (int j = 0; j < NPOLY / 2; ++j) {
      const int i0 = (NPOLY == 4) ? 4 * j + 1 : 3, i1 = (NPOLY == 4) ? 4 * j + 3 : 7;
      float e0, f0, e1, f1, y0, y1;
      cov_split<KIND>(__uint_as_float(s[i0]), e0, f0);
      cov_split<KIND>(__uint_as_float(s[i1]), e1, f1);
      ex2_poly2(e0, e1, y0, y1);
      p[i0] = (KIND == GP_RBF) ? y0 : f0 * y0;
      p[i1] = (KIND == GP_RBF) ? y1 : f1 * y1;
    }
  }
}
// tf32 truncation (one LOP3 each), residual in [0, 2^-10 p) by one packed FADD2 per pair, rounded to bf16 (RN), two per
// TMEM column: P is kept to ~2^-19 relative, random sign
__device__ __forceinline__ void split_group8(const float (&p)[8], uint32_t* __restrict__ hi, uint32_t* __restrict__ lo) {
#pragma unroll
  for (int i = 0; i < 8; i += 2) {
    const uint32_t h0 = __float_as_uint(p[i]) & 0xFFFFE000u, h1 = __float_as_uint(p[i + 1]) & 0xFFFFE000u;
    float l0, l1;
    sub_f32x2(p[i], p[i + 1], __uint_as_float(h0), __uint_as_float(h1), l0, l1);
    lo[i >> 1] = pack_bf16x2(l0, l1);
    hi[i] = h0;
    hi[i + 1] = h1;
  }
}

// One pipeline step.  On entry: pc = P of the group to be stored (MUFU issued one step ago), sn = S of the group that goes
// through the MUFU now (its load was issued one step ago).  (1) wait for sn, apply the exact-diagonal fix-up, (2) put the
// load of the group after that in flight into sn2 (address t_ld), (3) MUFU / polynomial on sn -> pn, (4) split pc and store.
template <int KIND, int NPOLY, int NLOAD>
__device__ __forceinline__ void epi_step(uint32_t t_st_hi, uint32_t t_st_lo, uint32_t t_ld, const float (&pc)[8], uint32_t (&sn)[8],
                                         float (&pn)[8], uint32_t (&sn2)[8], uint32_t (&sn3)[8], bool diag, int cd_rel) {
  tmem_wait_ld();
  if (diag) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i == cd_rel) sn[i] = 0u;   // a_ii = 0 exactly (kernel.py:44-45 fills the diagonal with 0)
  }
  if (NLOAD >= 1) GP_TMEM_LD8(t_ld, sn2);
  if (NLOAD >= 2) GP_TMEM_LD8(t_ld + 8, sn3);
  cov_group8<KIND, NPOLY>(sn, pn);
  uint32_t hi[8], lo[4];
  split_group8(pc, hi, lo);
  GP_TMEM_ST8(t_st_hi, hi);
  GP_TMEM_ST4(t_st_lo, lo);
}

template <int KIND, int NPOLY, bool A_TMEM, bool TRACE>
__global__ void __launch_bounds__(THREADS, 1)
kmv_tc2_kernel(const float* __restrict__ XA, const float* __restrict__ XB, const float* __restrict__ Vt,
               float* __restrict__ partial, int KP, int NS, int64_t ntile_j, int64_t tiles_per_split, int64_t rows_pad,
               int same, int64_t row_begin, const int* __restrict__ done_flag, long long* __restrict__ trace) {
  if (done_flag && *done_flag) return;  // CTA-uniform, before any barrier / TMEM state exists
  const bool tr = TRACE && trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0;
#define GP_TR2(tile, ev) do { if (TRACE && tr && lane == 0 && (tile) < 256) trace[(tile) * 16 + (ev)] = clock64(); } while (0)
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = (int)warp_idx_uniform();
  const int lane = threadIdx.x & 31;
  const int64_t ip = blockIdx.x;                 // 256-row block
  const int split = blockIdx.y;
  const int64_t jt0 = (int64_t)split * tiles_per_split;
  const int64_t jt1 = min(ntile_j, jt0 + tiles_per_split);
  const int T = (int)max((int64_t)0, jt1 - jt0);

  const uint32_t a_bytes = (uint32_t)KP * TILE_I * 4;
  const uint32_t b_bytes = (uint32_t)KP * TILE_J * 4;
  const uint32_t stage_bytes = b_bytes + V_TILE_BYTES;
  uint8_t* sStage = smem;
  uint8_t* sA = smem + (size_t)NS * stage_bytes;               // SS mode only: two A tiles
  Bars* bars = reinterpret_cast<Bars*>(sA + (A_TMEM ? 0 : 2 * a_bytes));

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars->a_full), A_TMEM ? 512 : 1);
    for (int s = 0; s < MAX_NS; ++s) {
      mbar_init(smem_u32(&bars->b_full[s]), 1);
      mbar_init(smem_u32(&bars->b_empty[s]), 2);
    }
    for (int w = 0; w < 2; ++w) {
      for (int b = 0; b < 2; ++b) {
        mbar_init(smem_u32(&bars->s_full[w][b]), 1);
        mbar_init(smem_u32(&bars->p_full[w][b]), 256);
      }
      mbar_init(smem_u32(&bars->o_full[w]), 1);
    }
    fence_mbar_init();
  }
  if (warp == W_MMA) tmem_alloc(smem_u32(&bars->tmem_base), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == W_PROD) {
    // ===================== TMA producer (one lane) =====================
    if (lane == 0 && T > 0) {
      if (!A_TMEM) {
        mbar_arrive_expect_tx(smem_u32(&bars->a_full), 2 * a_bytes);
        bulk_g2s(smem_u32(sA), XA + (ip * 2) * (int64_t)TILE_I * KP, 2 * a_bytes, smem_u32(&bars->a_full));  // two consecutive tiles
      }
      int sb = 0;
      uint32_t par = 1;
      for (int u = 0; u < T; ++u) {
        mbar_wait(smem_u32(&bars->b_empty[sb]), par);
        const uint32_t full = smem_u32(&bars->b_full[sb]);
        uint8_t* st = sStage + (size_t)sb * stage_bytes;
        const int64_t jt = jt0 + u;
        mbar_arrive_expect_tx(full, stage_bytes);
        bulk_g2s(smem_u32(st), XB + jt * (int64_t)TILE_J * KP, b_bytes, full);
        bulk_g2s(smem_u32(st + b_bytes), reinterpret_cast<const uint8_t*>(Vt) + jt * (int64_t)V_TILE_BYTES, V_TILE_BYTES, full);
        if (++sb == NS) { sb = 0; par ^= 1; }
      }
    }
  } else if (warp < 2) {
    // ===================== MMA issuers: warp 0 serves row half 0, warp 1 row half 1 ==========
    // Program order per issuer:  G1(w,0) G1(w,1) ; for u: wait P(w,u) -> GEMM2(w,u) -> GEMM1(w,u+2).
    // One thread issues everything of a warpgroup, so the tensor pipe orders GEMM2(w,u) (reads P in slot u&1) before
    // GEMM1(w,u+2) (overwrites it), and GEMM2(w,u) (overwrites O(w)) comes after the epilogue's arrive on p_full(w,u), which
    // follows its fold of O(w) from tile u-1.  Two issuers because the per-tile scalar work of ONE warp (descriptor updates,
    // two barrier polls, three commits, 25 MMAs: ~1000 cycles of dependent-issue latency, profiles/NOTES_r02.md) was the
    // bottleneck of the first version of this kernel; everything that can be is a running value updated by constant adds.
    const int w = warp;
    if (T > 0) {
      constexpr uint32_t IDESC1 = idesc_tf32(TILE_I, TILE_J);    // S = A B^T               128 x 64
      constexpr uint32_t IDESC2A = idesc_tf32(TILE_I, TP);       // O (+)= P_hi V_hi^T / V_lo^T   (tf32) 128 x 16
      constexpr uint32_t IDESC2B = idesc_bf16(TILE_I, TP);       // O += P_lo V^T           (bf16) 128 x 16
      const int ksteps1 = KP / 8;
      const uint32_t stage_d = stage_bytes >> 4;                 // descriptor units
      const uint64_t bdesc_first = smem_desc(smem_u32(sStage), TILE_J * 16, 128);
      const uint64_t vdesc_first = smem_desc(smem_u32(sStage + b_bytes), 2 * TP * 16, 128);            // tf32 tile: rows 0-15 V_hi, 16-31 V_lo
      const uint64_t wdesc_first = smem_desc(smem_u32(sStage + b_bytes + V_TF32_BYTES), TP * 16, 128);  // bf16 tile, 16 rows
      const uint64_t a_desc0 = smem_desc(smem_u32(sA + (size_t)w * a_bytes), TILE_I * 16, 128);
      const uint32_t a_t = tmem + (uint32_t)(COL_A + w * A_COLS_MAX);
      const uint32_t slot0 = tmem + (uint32_t)((2 * w) * SLOT_COLS);
      const uint32_t d_o = tmem + (uint32_t)(COL_O + w * TP);
      const uint32_t ofull = smem_u32(&bars->o_full[w]);
      const uint32_t bfull0 = smem_u32(&bars->b_full[0]), bempty0 = smem_u32(&bars->b_empty[0]);
      const uint32_t sfull0 = smem_u32(&bars->s_full[w][0]), pfull0 = smem_u32(&bars->p_full[w][0]);
      mbar_wait(smem_u32(&bars->a_full), 0);
      tc_fence_after();
      // running state of the GEMM1 stream (tile g1): ring stage, its parity, descriptor of its B tile
      int g1 = 0, sb1 = 0;
      uint32_t par1 = 0;
      uint64_t bdesc1 = bdesc_first;
      auto issue_g1 = [&]() {
        mbar_poll(bfull0 + 8u * (uint32_t)sb1, par1);
        tc_fence_after();
        if (w == 0 && g1 >= 2) GP_TR2(g1 - 2, 10);
        const uint32_t d_s = slot0 + (uint32_t)((g1 & 1) * SLOT_COLS);
        const uint32_t sfull = sfull0 + 8u * (uint32_t)(g1 & 1);
        if (elect_one()) {
          // a REAL loop over the k-steps with running operands (4 instructions per MMA); the unrolled-and-predicated and the
          // switch-per-count forms cost the issuer ~500 cycles per tile
          uint64_t bd = bdesc1;
          if (A_TMEM) {
            uint32_t aa = a_t;
            mma_tf32_ts_1t(d_s, aa, bd, IDESC1, 0u);
#pragma unroll 1
            for (int ks = 1; ks < ksteps1; ++ks) {
              aa += 8;
              bd += (uint64_t)((2 * TILE_J * 16) >> 4);
              mma_tf32_ts_1t(d_s, aa, bd, IDESC1, 1u);
            }
          } else {
            uint64_t ad = a_desc0;
            mma_tf32_ss_1t(d_s, ad, bd, IDESC1, 0u);
#pragma unroll 1
            for (int ks = 1; ks < ksteps1; ++ks) {
              ad += (uint64_t)((2 * TILE_I * 16) >> 4);
              bd += (uint64_t)((2 * TILE_J * 16) >> 4);
              mma_tf32_ss_1t(d_s, ad, bd, IDESC1, 1u);
            }
          }
          tc_commit_1t(sfull);
        }
        __syncwarp();
        ++g1;
        bdesc1 += stage_d;
        if (++sb1 == NS) { sb1 = 0; par1 ^= 1; bdesc1 = bdesc_first; }
      };
      issue_g1();
      if (T > 1) issue_g1();
      // running state of the GEMM2 stream (tile u)
      int sb2 = 0;
      uint64_t vdesc = vdesc_first, wdesc = wdesc_first;
      uint32_t ppar = 0;
#pragma unroll 1
      for (int u = 0; u < T; ++u) {
        const uint32_t b = (uint32_t)(u & 1);
        mbar_poll(pfull0 + 8u * b, ppar);
        tc_fence_after();
        GP_TR2(u, w == 0 ? 1 : 6);
        const uint32_t p_hi = slot0 + b * SLOT_COLS;
        const uint32_t p_lo = p_hi + TILE_J;
        const uint32_t bempty = bempty0 + 8u * (uint32_t)sb2;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < TILE_J / 8; ++ks) {
            const uint64_t vd = vdesc + (uint64_t)(ks * ((2 * 2 * TP * 16) >> 4));
            mma_tf32_ts_1t(d_o, p_hi + ks * 8, vd, IDESC2A, ks > 0 ? 1u : 0u);                 // V_hi rows
            mma_tf32_ts_1t(d_o, p_hi + ks * 8, vd + (uint64_t)((TP * 16) >> 4), IDESC2A, 1u);   // V_lo rows (+256 B)
          }
#pragma unroll
          for (int ks = 0; ks < TILE_J / 16; ++ks)
            mma_bf16_ts_1t(d_o, p_lo + ks * 8, wdesc + (uint64_t)(ks * ((2 * TP * 16) >> 4)), IDESC2B, 1u);
          tc_commit_1t(bempty);   // (one of two arrivals) this warpgroup's GEMM2 has read the V stage
          tc_commit_1t(ofull);    // O(w) holds tile u's product
        }
        __syncwarp();
        if (w == 0) GP_TR2(u, 9);
        ppar ^= b;                // the parity of p_full[w][b] flips every second tile
        vdesc += stage_d;
        wdesc += stage_d;
        if (++sb2 == NS) { sb2 = 0; vdesc = vdesc_first; wdesc = wdesc_first; }
        if (g1 < T) issue_g1();   // refill the slot GEMM2(w,u) has just consumed (same thread => ordered)
        if (w == 0) GP_TR2(u, 7);
      }
    }
  } else if (warp == 3) {
    // ===================== observer (trace runs only): when do S(0,u) and O(0,u) really complete? =====================
    if (tr && T > 0) {
      if (lane == 0) {
        for (int u = 0; u < T && u < 256; ++u) {
          mbar_wait(smem_u32(&bars->s_full[0][u & 1]), (uint32_t)((u >> 1) & 1));
          trace[u * 16 + 0] = clock64();
        }
      } else if (lane == 1) {
        for (int u = 0; u < T && u < 256; ++u) {
          mbar_wait(smem_u32(&bars->o_full[0]), (uint32_t)(u & 1));
          trace[u * 16 + 8] = clock64();
        }
      }
    }
  } else if (warp >= W_EPI) {
    // ===================== epilogue warps: 16 = 2 warpgroups (row halves) x 2 column halves x 4 lane quadrants =========
    // epilogue warp i = warp - 4: warpgroup wg = i >> 3 (rows [128 wg, +128) of the block), column half h = (i >> 2) & 1 (columns [32 h, +32) of
    // every 64-column tile, and columns [8 h, +8) of O), lane quadrant q = i & 3.  Four warps per SM sub-partition hide the
    // TMEM-load / MUFU / TMEM-store latencies of one another (with two, the first version ran at half the MUFU rate).
    const int wg = (warp - W_EPI) >> 3;
    const int h = ((warp - W_EPI) >> 2) & 1;
    const int q = warp & 3;            // TMEM lane quadrant of this warp (warp % 4)
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const int64_t rloc = ip * ROWS_CTA + wg * TILE_I + q * 32 + lane;   // local (padded) row of this thread
    const int64_t gi = row_begin + rloc;                                // global row
    const uint32_t t_o = tmem + lane_off + (uint32_t)(COL_O + wg * TP + h * 8);
    const uint32_t t_slot0 = tmem + lane_off + (uint32_t)((2 * wg) * SLOT_COLS + h * 32);       // this warp's S / P_hi columns
    const uint32_t t_lo0 = tmem + lane_off + (uint32_t)((2 * wg) * SLOT_COLS + TILE_J + h * 16);  // ... and P_lo columns
    if (A_TMEM) {
      // this thread's row of the A tile: XA[tile][kc][row][4] -> TMEM columns COL_A + 48 wg + [0, KP); the two column halves
      // of a warpgroup store alternate 8-column blocks
      const float4* src = reinterpret_cast<const float4*>(XA + (ip * 2 + wg) * (int64_t)TILE_I * KP) + (q * 32 + lane);
      const uint32_t t_a = tmem + lane_off + (uint32_t)(COL_A + wg * A_COLS_MAX);
      if (T > 0) {
#pragma unroll
        for (int k8 = 0; k8 < A_COLS_MAX / 8; ++k8) {
          if (k8 * 8 < KP && (k8 & 1) == h) {
            const float4 v0 = __ldg(src + (size_t)(2 * k8) * TILE_I), v1 = __ldg(src + (size_t)(2 * k8 + 1) * TILE_I);
            uint32_t r[8] = {__float_as_uint(v0.x), __float_as_uint(v0.y), __float_as_uint(v0.z), __float_as_uint(v0.w),
                             __float_as_uint(v1.x), __float_as_uint(v1.y), __float_as_uint(v1.z), __float_as_uint(v1.w)};
            GP_TMEM_ST8(t_a + 8 * k8, r);
          }
        }
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(smem_u32(&bars->a_full));
      }
    }
    // O is folded into fp32 registers after EVERY tile: the tensor core's accumulator truncates on each add, so long
    // TMEM accumulation chains drift (1e-4 at N = 50k); 20 adds per tile keep the product at fp32 level.
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.f;
    if (T > 0) {
      // S of column group g (8 columns) of the current tile lives in register set g & 3: four sets, so a load never has to
      // wait for the MUFU ops that still read the set it overwrites.  P alternates between pa / pb.
      uint32_t s0[8], s1[8], s2[8], s3[8];
      float pa[8], pb[8];
      auto tile_diag = [&](int u, int& cd) -> bool {
        const int64_t jbase = (jt0 + u) * TILE_J + h * 32;             // first column of this warp's half tile
        const int64_t r0 = row_begin + ip * ROWS_CTA + wg * TILE_I;
        cd = (int)max((int64_t)-1000000, min((int64_t)1000000, gi - jbase));
        return same && (r0 < jbase + 32) && (jbase < r0 + TILE_I);
      };
      // No prefetch across tiles: S(u+1) is asked for only after P(u) has been released.  GEMM1(wg, u+1) can start no earlier
      // than the release of P(u-1) (it overwrites that slot) and, with the issuer's ~1000-cycle turn-around, needs the whole
      // time the warpgroup spends on tile u; the bubble at the start of a tile is filled by the other three warps of the
      // SM sub-partition.
#pragma unroll 1
      for (int u = 0; u < T; ++u) {
        const int b = u & 1;
        const uint32_t t_s = t_slot0 + (uint32_t)(b * SLOT_COLS);       // S, overwritten in place by P_hi
        const uint32_t t_lo = t_lo0 + (uint32_t)(b * SLOT_COLS);
        int cd;
        const bool diag = tile_diag(u, cd);
        mbar_wait(smem_u32(&bars->s_full[wg][b]), (uint32_t)((u >> 1) & 1));
        tc_fence_after();
        if (q == 0 && h == 0) GP_TR2(u, 2 + wg);
        GP_TMEM_LD8(t_s, s0);
        GP_TMEM_LD8(t_s + 8, s1);
        tmem_wait_ld();
        if (diag) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (i == cd) s0[i] = 0u;
        }
        GP_TMEM_LD8(t_s + 16, s2);
        cov_group8<KIND, NPOLY>(s0, pa);
        epi_step<KIND, NPOLY, 1>(t_s, t_lo, t_s + 24, pa, s1, pb, s3, s3, diag, cd - 8);       // MUFU g1, load g3, store g0
        epi_step<KIND, NPOLY, 0>(t_s + 8, t_lo + 4, t_s, pb, s2, pa, s0, s0, diag, cd - 16);   // MUFU g2, store g1
        epi_step<KIND, NPOLY, 0>(t_s + 16, t_lo + 8, t_s, pa, s3, pb, s0, s0, diag, cd - 24);  // MUFU g3, store g2
        // fold this warp's 8 columns of O(u-1) before P(u) is released (GEMM2(wg, u) overwrites O): GEMM2(wg, u-1) was issued
        // when the warpgroup released tile u-1, most of a tile-time ago
        if (u >= 1) {
          mbar_wait(smem_u32(&bars->o_full[wg]), (uint32_t)((u - 1) & 1));
          tc_fence_after();
          uint32_t o[8];
          GP_TMEM_LD8(t_o, o);
          tmem_wait_ld();
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[c] += __uint_as_float(o[c]);
        }
        {
          uint32_t hi[8], lo[4];
          split_group8(pb, hi, lo);
          GP_TMEM_ST8(t_s + 24, hi);
          GP_TMEM_ST4(t_lo + 12, lo);
        }
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(smem_u32(&bars->p_full[wg][b]));  // GEMM2(wg, u) may now read P and overwrite O(wg)
        if (q == 0 && h == 0) GP_TR2(u, 4 + wg);
      }
      // the last tile's product is still in TMEM
      mbar_wait(smem_u32(&bars->o_full[wg]), (uint32_t)((T - 1) & 1));
      tc_fence_after();
      uint32_t o[8];
      GP_TMEM_LD8(t_o, o);
      tmem_wait_ld();
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] += __uint_as_float(o[c]);
    }
    float4* dst = reinterpret_cast<float4*>(partial + ((int64_t)split * rows_pad + rloc) * TP + h * 8);
    dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem, TMEM_COLS);
  }
#undef GP_TR2
}

static int smem_bytes(int KP, bool a_tmem, int* ns_out) {
  const int a_bytes = a_tmem ? 0 : 2 * KP * TILE_I * 4;
  const int stage = KP * TILE_J * 4 + V_TILE_BYTES;
  int ns = (226 * 1024 - a_bytes - (int)sizeof(Bars) - 1024) / stage;
  if (ns > MAX_NS) ns = MAX_NS;
  *ns_out = ns;
  return a_bytes + ns * stage + (int)sizeof(Bars) + 64;
}

template <int KIND, int NPOLY, bool A_TMEM, bool TRACE = false>
static int launch_one(gp_plan* p, const int* done_flag) {
  int ns = 0;
  const int smem = smem_bytes(p->KP, A_TMEM, &ns);
  GP_REQUIRE(ns >= 4, GP_E_SHAPE, "tcgen05 path: smem ring too small for KP=%d", p->KP);
  static bool attr_done[64] = {};   // function attributes are per device
  const int dev_slot = p->device & 63;
  if (!attr_done[dev_slot]) {
    GP_CUDA(cudaFuncSetAttribute(kmv_tc2_kernel<KIND, NPOLY, A_TMEM, TRACE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_done[dev_slot] = true;
  }
  dim3 grid((unsigned)(p->rows_pad / ROWS_CTA), (unsigned)p->nsplit);
  kmv_tc2_kernel<KIND, NPOLY, A_TMEM, TRACE><<<grid, THREADS, smem, p->stream>>>(
      p->XA.as<float>(), p->XB.as<float>(), vtiles_ptr(p), partial_ptr(p), p->KP, ns, p->ntile_j,
      p->tiles_per_split, p->rows_pad, p->same ? 1 : 0, p->row_begin, done_flag, p->tc_trace);
  p->launches++;
  GP_CUDA(cudaGetLastError());
  return GP_OK;
}

template <int KIND>
static int launch_kind(gp_plan* p, const int* done_flag) {
  const bool a_tmem = p->KP <= A_COLS_MAX;
  const int npoly = p->npoly;
  if (KIND == GP_RBF && p->tc_trace && a_tmem && npoly == 0) return launch_one<GP_RBF, 0, true, true>(p, done_flag);   // tools/tc_trace2.py
#define GP_V2_CASE(NP)                                                   \
  case NP:                                                               \
    return a_tmem ? launch_one<KIND, NP, true>(p, done_flag) : launch_one<KIND, NP, false>(p, done_flag);
  switch (npoly) {
    GP_V2_CASE(0)
    GP_V2_CASE(2)
    GP_V2_CASE(4)
  }
#undef GP_V2_CASE
  set_error("bad polynomial share %d (0, 2 or 4 of 8)", npoly);
  return GP_E_SHAPE;
}

}  // namespace v2

int kmv_tc2_launch_kind(gp_plan* p, int kind, const int* done_flag) {
  switch (kind) {
    case GP_RBF: return v2::launch_kind<GP_RBF>(p, done_flag);
    case GP_MATERN12: return v2::launch_kind<GP_MATERN12>(p, done_flag);
    case GP_MATERN32: return v2::launch_kind<GP_MATERN32>(p, done_flag);
    case GP_MATERN52: return v2::launch_kind<GP_MATERN52>(p, done_flag);
    case GP_DERIV + GP_RBF: return v2::launch_kind<GP_DERIV + GP_RBF>(p, done_flag);
    case GP_DERIV + GP_MATERN12: return v2::launch_kind<GP_DERIV + GP_MATERN12>(p, done_flag);
    case GP_DERIV + GP_MATERN32: return v2::launch_kind<GP_DERIV + GP_MATERN32>(p, done_flag);
    case GP_DERIV + GP_MATERN52: return v2::launch_kind<GP_DERIV + GP_MATERN52>(p, done_flag);
  }
  set_error("bad kernel kind %d", kind);
  return GP_E_SHAPE;
}

}  // namespace gp
