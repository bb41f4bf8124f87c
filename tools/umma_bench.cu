// umma_bench.cu -- microbenchmarks that decide the GEMM2 design of kmv_tc.cu:
//  (1) cycles per tcgen05.mma for A-from-TMEM (TS) and smem (SS) operands at several N, kind::tf32 and kind::f16
//  (2) correctness of kind::f16 with A in TMEM packed two fp16 per 32-bit column (which half is the lower k?)
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "tc_ptx.cuh"
using namespace gp::ptx;

__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {  // fp16 x fp16 -> fp32, K-major
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_f16_ts_1t(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

struct Bars { uint64_t done; uint32_t tmem; uint32_t pad; };

// MODE 0: TS tf32, 1: SS tf32, 2: TS f16.  32 MMAs fully unrolled (uniform-register operands, like kmv_tc.cu),
// rotating over NACC independent accumulators; repeated `outer` times.
template <int MODE, int N, int NACC>
__global__ void __launch_bounds__(128, 1) bench_kernel(int outer, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Bars* bars = reinterpret_cast<Bars*>(smem + 96 * 1024);
  const int warp = (int)warp_idx_uniform();
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bars->done), 1); fence_mbar_init(); }
  if (warp == 1) tmem_alloc(smem_u32(&bars->tmem), 512);
  for (int i = threadIdx.x; i < 24 * 1024; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = bars->tmem;
  if (warp == 0) {
    const uint64_t adesc = smem_desc(smem_u32(smem), 128 * 16, 128);
    const uint64_t bdesc = smem_desc(smem_u32(smem + 32 * 1024), N * 16, 128);
    constexpr uint32_t id_t = idesc_tf32(128, N), id_h = idesc_f16(128, N);
    long long t0 = clock64();
    if (elect_one()) {
      for (int o = 0; o < outer; ++o) {
#pragma unroll
        for (int r = 0; r < 32; ++r) {
          const uint32_t dcol = tmem + 256 + (uint32_t)((r % NACC) * 64);
          if (MODE == 0) mma_tf32_ts_1t(dcol, tmem + (r & 7) * 8, bdesc + (uint64_t)((r & 3) * 64), id_t, 1);
          else if (MODE == 1) mma_tf32_ss_1t(dcol, adesc + (uint64_t)((r & 3) * 256), bdesc + (uint64_t)((r & 3) * 64), id_t, 1);
          else mma_f16_ts_1t(dcol, tmem + (r & 7) * 8, bdesc + (uint64_t)((r & 3) * 64), id_h, 1);
        }
      }
      tc_commit_1t(smem_u32(&bars->done));
    }
    __syncwarp();
    long long t1 = clock64();
    mbar_wait(smem_u32(&bars->done), 0);
    long long t2 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int MODE, int N, int NACC>
static void run_bench(long long* d_out) {
  const char* names[] = {"TS tf32", "SS tf32", "TS f16 "};
  cudaFuncSetAttribute(bench_kernel<MODE, N, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  long long h[2][2];
  int outers[2] = {2, 8};
  for (int i = 0; i < 2; ++i) {
    bench_kernel<MODE, N, NACC><<<1, 128, 100 * 1024>>>(outers[i], d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s N=%d: %s\n", names[MODE], N, cudaGetErrorString(e)); exit(1); }
    cudaMemcpy(h[i], d_out, 16, cudaMemcpyDeviceToHost);
  }
  printf("%s M=128 N=%3d nacc=%d: issue %.1f cyc/mma, complete %.1f cyc/mma\n", names[MODE], N, NACC, (h[1][0] - h[0][0]) / 192.0,
         (h[1][1] - h[0][1]) / 192.0);
}

// f16 TS correctness: D[128 x 16] = P[128 x 32 (fp16, packed in 16 tmem columns)] . V[16 x 32]^T
__global__ void __launch_bounds__(128, 1) f16_kernel(const __half* Vpk, float* D, int swap) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Bars* bars = reinterpret_cast<Bars*>(smem + 8192);
  const int warp = (int)warp_idx_uniform(), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bars->done), 1); fence_mbar_init(); }
  if (warp == 1) tmem_alloc(smem_u32(&bars->tmem), 512);
  for (int i = threadIdx.x; i < 16 * 32; i += 128) reinterpret_cast<__half*>(smem)[i] = Vpk[i];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = bars->tmem;
  const int row = warp * 32 + lane;
  // P[row][k] = (row % 7 + 1) * 0.125 + k * 0.0625  (exact in fp16)
  uint32_t r[16];
  for (int c = 0; c < 16; ++c) {
    float e0 = (row % 7 + 1) * 0.125f + (2 * c) * 0.0625f, e1 = (row % 7 + 1) * 0.125f + (2 * c + 1) * 0.0625f;
    __half2 h = swap ? __floats2half2_rn(e1, e0) : __floats2half2_rn(e0, e1);  // .x = low 16 bits
    r[c] = *reinterpret_cast<uint32_t*>(&h);
  }
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(tmem + ((uint32_t)(warp * 32) << 16)),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
  tmem_wait_st();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (warp == 0) {
    // V smem layout K-major no-swizzle, 16-byte chunks = 8 halves: [k/8][n (16)][8]
    if (elect_one()) {
      for (int ks = 0; ks < 2; ++ks) {
        uint64_t bd = smem_desc(smem_u32(smem) + ks * 2 * (16 * 16), 16 * 16, 128);
        mma_f16_ts_1t(tmem + 64, tmem + ks * 8, bd, idesc_f16(128, 16), ks > 0);
      }
      tc_commit_1t(smem_u32(&bars->done));
    }
    __syncwarp();
  }
  mbar_wait(smem_u32(&bars->done), 0);
  tc_fence_after();
  uint32_t o[16];
  GP_TMEM_LD16(tmem + ((uint32_t)(warp * 32) << 16) + 64, o);
  tmem_wait_ld();
  for (int c = 0; c < 16; ++c) D[row * 16 + c] = __uint_as_float(o[c]);
  tc_fence_before(); __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d_out; cudaMalloc(&d_out, 16);
  run_bench<0, 16, 1>(d_out); run_bench<0, 16, 2>(d_out); run_bench<0, 16, 4>(d_out);
  run_bench<0, 32, 1>(d_out); run_bench<0, 32, 2>(d_out); run_bench<0, 64, 1>(d_out); run_bench<0, 128, 1>(d_out); run_bench<0, 256, 1>(d_out);
  run_bench<1, 16, 1>(d_out); run_bench<1, 64, 1>(d_out); run_bench<1, 64, 2>(d_out); run_bench<1, 128, 1>(d_out); run_bench<1, 256, 1>(d_out);
  run_bench<2, 16, 1>(d_out); run_bench<2, 32, 1>(d_out); run_bench<2, 32, 2>(d_out); run_bench<2, 64, 1>(d_out);
  // f16 packing test
  std::vector<__half> V(16 * 32), Vpk(16 * 32);
  for (int c = 0; c < 16; ++c) for (int k = 0; k < 32; ++k) V[c * 32 + k] = __float2half((float)((c * 3 + k * 5) % 11 - 5) * 0.25f);
  for (int c = 0; c < 16; ++c) for (int k = 0; k < 32; ++k) Vpk[((k / 8) * 16 + c) * 8 + k % 8] = V[c * 32 + k];
  __half* dV; float* dD; cudaMalloc(&dV, Vpk.size() * 2); cudaMalloc(&dD, 128 * 16 * 4);
  cudaMemcpy(dV, Vpk.data(), Vpk.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 1024);
  for (int swap = 0; swap < 2; ++swap) {
    f16_kernel<<<1, 128, 16 * 1024>>>(dV, dD, swap);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("f16 swap=%d: %s\n", swap, cudaGetErrorString(e)); return 1; }
    std::vector<float> D(128 * 16); cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double err = 0;
    for (int r = 0; r < 128; ++r) for (int c = 0; c < 16; ++c) {
      double acc = 0; for (int k = 0; k < 32; ++k) acc += ((r % 7 + 1) * 0.125 + k * 0.0625) * (double)__half2float(V[c * 32 + k]);
      err = fmax(err, fabs(acc - D[r * 16 + c]));
    }
    printf("f16 TS packing: low-half-is-%s-k  max|err| = %.3e %s\n", swap ? "odd" : "even", err, err < 1e-3 ? "OK" : "MISMATCH");
  }
  return 0;
}
