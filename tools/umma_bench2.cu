// umma_bench2.cu -- what does one tile's tensor work REALLY cost inside the fused K.V kernel?
// The round-1 microbenchmark (umma_bench.cu) timed one MMA shape back to back on an idle SM.  The first version of
// kmv_tc2.cu showed the issuer thread needing ~1000 cycles for 25 MMAs whose floor is ~320 cycles, with the tensor pipe only
// 24 % active.  This benchmark replays the exact per-tile MMA sequences of the kernels (GEMM2 variants + GEMM1 TS / SS), with and
// without the epilogue's TMEM traffic (tcgen05.ld / st from four other warps), and reports cycles per tile.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I gpytorch_b200/csrc tools/umma_bench2.cu -o umma_bench2
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "tc_ptx.cuh"
using namespace gp::ptx;

__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {  // fp16 x fp16 -> fp32, K-major
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_f16_ts_1t(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

struct Bars { uint64_t done; uint64_t stop; uint64_t d2; uint64_t d3; uint64_t d4; uint32_t tmem; uint32_t pad; };

// SEQ: which per-tile MMA sequence the issuer replays (TMEM map as in kmv_tc2.cu: slots of 96 columns, O at 384, A at 416)
//  0: v2     GEMM2 = 16 x tf32 TS N=16 (8 k-steps x {V_hi, V_lo}) + 4 x bf16 TS N=16 ; GEMM1 = 5 x tf32 TS N=64
//  1: v16    GEMM2 =  8 x tf32 TS N=32 + 4 x bf16 TS N=16                             ; GEMM1 = 5 x tf32 SS N=64
//  2: f16    GEMM2 =  4 x f16 TS N=32 + 4 x f16 TS N=16                               ; GEMM1 = 5 x tf32 TS N=64
//  3: only GEMM1 TS (5 x N=64)     4: only GEMM1 SS     5: only GEMM2 of v2      6: only GEMM2 of v16     7: only GEMM2 f16
//  8: v2 GEMM2 but every MMA into its own accumulator (no dependent accumulation chain)
//  9: 20 x tf32 TS N=16 with ONE A block (operand re-use)
// 10: GEMM2 = 8 x tf32 TS N=32 (wide O) + 4 x bf16 N=16 ; GEMM1 TS
template <int SEQ, int TRAFFIC>
__global__ void __launch_bounds__(576, 1) bench_kernel(int tiles, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Bars* bars = reinterpret_cast<Bars*>(smem + 96 * 1024);
  const int warp = (int)warp_idx_uniform();
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bars->done), 1); mbar_init(smem_u32(&bars->stop), 1); mbar_init(smem_u32(&bars->d2), 1); mbar_init(smem_u32(&bars->d3), 1); mbar_init(smem_u32(&bars->d4), 1); fence_mbar_init(); }
  if (warp == 17) tmem_alloc(smem_u32(&bars->tmem), 512);
  for (int i = threadIdx.x; i < 24 * 1024; i += 576) reinterpret_cast<float*>(smem)[i] = 0.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = bars->tmem;
  volatile int* stopflag = reinterpret_cast<volatile int*>(smem + 97 * 1024);
  if (threadIdx.x == 0) *stopflag = 0;
  __syncthreads();
  if (warp == 16) {
    const uint64_t adesc = smem_desc(smem_u32(smem), 128 * 16, 128);               // A tile (SS): [KP/4][128][4]
    const uint64_t bdesc = smem_desc(smem_u32(smem + 32 * 1024), 64 * 16, 128);     // B tile: [KP/4][64][4]
    const uint64_t vdesc = smem_desc(smem_u32(smem + 48 * 1024), 2 * 16 * 16, 128); // V tf32 tile: [16][32][4]
    const uint64_t wdesc = smem_desc(smem_u32(smem + 56 * 1024), 16 * 16, 128);     // V bf16 tile: [8][16][8]
    constexpr uint32_t ID1 = gp::ptx::idesc_tf32(128, 64), ID16 = gp::ptx::idesc_tf32(128, 16), ID32 = gp::ptx::idesc_tf32(128, 32);
    constexpr uint32_t IB16 = gp::ptx::idesc_bf16(128, 16), IH32 = idesc_f16(128, 32), IH16 = idesc_f16(128, 16);
    long long t0 = clock64();
    if (elect_one()) {
      for (int u = 0; u < tiles; ++u) {
        const uint32_t slot = tmem + (uint32_t)((u & 3) * 96);
        const uint32_t p_hi = slot, p_lo = slot + 64, d_o = tmem + 384 + (uint32_t)((u & 1) * 16), a_t = tmem + 416;
        // ---- GEMM2 ----
        if (SEQ == 0 || SEQ == 5 || SEQ == 11) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            mma_tf32_ts_1t(d_o, p_hi + ks * 8, vdesc + (uint64_t)(ks * 64), ID16, ks > 0);
            mma_tf32_ts_1t(d_o, p_hi + ks * 8, vdesc + (uint64_t)(ks * 64 + 16), ID16, 1);
          }
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) mma_bf16_ts_1t(d_o, p_lo + ks * 8, wdesc + (uint64_t)(ks * 32), IB16, 1);
          if (SEQ == 11) { tc_commit_1t(smem_u32(&bars->stop)); tc_commit_1t(smem_u32(&bars->stop) + 16); }
        } else if (SEQ == 1 || SEQ == 6 || SEQ == 10) {
          const uint32_t d_w = tmem + 384 + (uint32_t)((u & 1) * 32);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) mma_tf32_ts_1t(d_w, p_hi + ks * 8, vdesc + (uint64_t)(ks * 64), ID32, ks > 0);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) mma_bf16_ts_1t(d_w, p_lo + ks * 8, wdesc + (uint64_t)(ks * 32), IB16, 1);
        } else if (SEQ == 2 || SEQ == 7) {
          const uint32_t d_w = tmem + 384 + (uint32_t)((u & 1) * 32);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) mma_f16_ts_1t(d_w, p_hi + ks * 8, vdesc + (uint64_t)(ks * 64), IH32, ks > 0);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) mma_f16_ts_1t(d_w, p_hi + 32 + ks * 8, wdesc + (uint64_t)(ks * 32), IH16, 1);
        } else if (SEQ == 8) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            mma_tf32_ts_1t(tmem + 256 + (uint32_t)(ks * 32), p_hi + ks * 8, vdesc + (uint64_t)(ks * 64), ID16, 1);
            mma_tf32_ts_1t(tmem + 256 + (uint32_t)(ks * 32 + 16), p_hi + ks * 8, vdesc + (uint64_t)(ks * 64 + 16), ID16, 1);
          }
        } else if (SEQ == 9) {
#pragma unroll
          for (int ks = 0; ks < 20; ++ks) mma_tf32_ts_1t(d_o, p_hi, vdesc, ID16, 1);
        }
        // ---- GEMM1 (next tile of this slot) ----
        if (SEQ == 0 || SEQ == 2 || SEQ == 3 || SEQ == 10 || SEQ == 11) {
#pragma unroll
          for (int ks = 0; ks < 5; ++ks) mma_tf32_ts_1t(slot, a_t + ks * 8, bdesc + (uint64_t)(ks * 128), ID1, ks > 0);
          if (SEQ == 11) tc_commit_1t(smem_u32(&bars->stop) + 24);
        } else if (SEQ == 1 || SEQ == 4) {
#pragma unroll
          for (int ks = 0; ks < 5; ++ks) mma_tf32_ss_1t(slot, adesc + (uint64_t)(ks * 256), bdesc + (uint64_t)(ks * 128), ID1, ks > 0);
        }
      }
      tc_commit_1t(smem_u32(&bars->done));
    }
    __syncwarp();
    long long t1 = clock64();
    mbar_wait(smem_u32(&bars->done), 0);
    long long t2 = clock64();
    *stopflag = 1;
    if (lane == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  } else if (warp < 16 && TRAFFIC && (TRAFFIC >= 2 || warp < 4)) {
    // the epilogue's load on the SM: per 8 columns one LDTM.x8, one STTM.x8, one STTM.x4, 8 MUFU, ~20 ALU/FMA instructions.
    // TRAFFIC 1: one warp per sub-partition, 2: four warps per sub-partition (the real kernel), 3: as 2 without the TMEM
    // accesses, 4: as 2 without the MUFU / ALU work
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t cbase = (uint32_t)((warp >> 2) * 24);
    uint32_t s[8], lo[4];
    float acc = 0.f;
    long long n = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = 0x3f000000u + lane + i;
    while (!*stopflag) {
#pragma unroll 1
      for (int g = 0; g < 8; ++g) {
        const uint32_t t = tmem + lane_off + cbase + (uint32_t)(((n & 3) * 96) + (g & 1) * 8);
        if (TRAFFIC != 3) { GP_TMEM_LD8(t, s); tmem_wait_ld(); }
        if (TRAFFIC != 4) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float e = ex2a(__uint_as_float(s[i]) * 0.5f);
            acc += e;
            s[i] = __float_as_uint(e) & 0xFFFFE000u;
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) lo[i] = s[i] ^ s[i + 4];
        if (TRAFFIC != 3) {
          GP_TMEM_ST8(t, s);
          GP_TMEM_ST4(tmem + lane_off + cbase + (uint32_t)(((n & 3) * 96) + 16 + (g & 1) * 4), lo);
        } else {
          acc += __uint_as_float(lo[0] ^ lo[1] ^ lo[2] ^ lo[3]);
        }
      }
      if (TRAFFIC != 3) tmem_wait_st();
      ++n;
    }
    if (acc == 12345.f) out[3] = n;
    if (lane == 0 && warp == 0) out[2] = n;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 17) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int SEQ, int TRAFFIC>
static void run(long long* d_out, const char* name) {
  cudaFuncSetAttribute(bench_kernel<SEQ, TRAFFIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  long long h[2][4];
  int tiles[2] = {64, 256};
  for (int i = 0; i < 2; ++i) {
    cudaMemset(d_out, 0, 32);
    bench_kernel<SEQ, TRAFFIC><<<1, 576, 100 * 1024>>>(tiles[i], d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
    cudaMemcpy(h[i], d_out, 32, cudaMemcpyDeviceToHost);
  }
  printf("%-70s traffic=%d: issue %.1f cyc/tile, complete %.1f cyc/tile (epilogue tiles/warp during run: %lld)\n", name, (int)TRAFFIC,
         (h[1][0] - h[0][0]) / 192.0, (h[1][1] - h[0][1]) / 192.0, h[1][2]);
}

int main() {
  long long* d_out; cudaMalloc(&d_out, 32);
#define RUN(S, NAME) run<S, 0>(d_out, NAME); run<S, 2>(d_out, NAME);
  RUN(0, "v2 tile: GEMM2 v2 + GEMM1 TS")
  RUN(11, "v2 tile + 3 tcgen05.commit per tile")
  return 0;
}
