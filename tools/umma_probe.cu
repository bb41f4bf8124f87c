// umma_probe.cu -- standalone check of the tcgen05 building blocks used by kmv_tc.cu:
//   bulk TMA (cp.async.bulk) of pre-packed K-major no-swizzle tiles, tcgen05.mma kind::tf32 with both
//   operands in smem (GEMM1) and with the A operand in TMEM (GEMM2), tcgen05.ld / tcgen05.st round trip.
// usage: umma_probe [variant]   variant 0: LBO = K-chunk stride, SBO = 8-row-group stride (what kmv_tc.cu uses)
//                               variant 1: the two swapped (diagnostic only)
// Prints max |err| of D1 = A B^T (128x96, K=40) and D2 = tf32(0.5 D1) V^T (128x16) against a CPU evaluation.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "tc_ptx.cuh"

using namespace gp::ptx;

constexpr int M = 128, N1 = 96, N2 = 16;

struct Bars { uint64_t full, mma1, pfull, mma2; uint32_t tmem; uint32_t pad; };

__global__ void __launch_bounds__(128, 1) probe_kernel(const float* Apk, const float* Bpk, const float* Vpk, float* D1, float* D2, int K, int variant) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t a_bytes = K * M * 4, b_bytes = K * N1 * 4, v_bytes = N1 * N2 * 4;
  uint8_t* sA = smem; uint8_t* sB = sA + a_bytes; uint8_t* sV = sB + b_bytes;
  Bars* bars = reinterpret_cast<Bars*>(sV + v_bytes);
  const int warp = (int)warp_idx_uniform(), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars->full), 1); mbar_init(smem_u32(&bars->mma1), 1);
    mbar_init(smem_u32(&bars->pfull), 128); mbar_init(smem_u32(&bars->mma2), 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&bars->tmem), 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = bars->tmem;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(smem_u32(&bars->full), a_bytes + b_bytes + v_bytes);
    bulk_g2s(smem_u32(sA), Apk, a_bytes, smem_u32(&bars->full));
    bulk_g2s(smem_u32(sB), Bpk, b_bytes, smem_u32(&bars->full));
    bulk_g2s(smem_u32(sV), Vpk, v_bytes, smem_u32(&bars->full));
  }
  if (warp == 0) {  // converged warp, elected lane issues (same pattern as kmv_tc.cu)
    mbar_wait(smem_u32(&bars->full), 0);
    tc_fence_after();
    for (int ks = 0; ks < K / 8; ++ks) {
      uint32_t la = M * 16, lb = N1 * 16, s8 = 128;
      uint64_t ad = variant == 0 ? smem_desc(smem_u32(sA) + ks * 2 * la, la, s8) : smem_desc(smem_u32(sA) + ks * 2 * la, s8, la);
      uint64_t bd = variant == 0 ? smem_desc(smem_u32(sB) + ks * 2 * lb, lb, s8) : smem_desc(smem_u32(sB) + ks * 2 * lb, s8, lb);
      mma_tf32_ss(tmem + 0, ad, bd, idesc_tf32(M, N1), ks > 0);
    }
    tc_commit(smem_u32(&bars->mma1));
  }
  mbar_wait(smem_u32(&bars->mma1), 0);
  tc_fence_after();
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
  const int row = warp * 32 + lane;
  for (int ch = 0; ch < 3; ++ch) {
    uint32_t r[32];
    GP_TMEM_LD32(tmem + lane_off + ch * 32, r);
    tmem_wait_ld();
    for (int c = 0; c < 32; ++c) {
      D1[row * N1 + ch * 32 + c] = __uint_as_float(r[c]);
      r[c] = __float_as_uint(0.5f * __uint_as_float(r[c])) & 0xFFFFE000u;
    }
    GP_TMEM_ST32(tmem + lane_off + 96 + ch * 32, r);
  }
  tmem_wait_st();
  tc_fence_before();
  mbar_arrive(smem_u32(&bars->pfull));
  if (warp == 0) {
    mbar_wait(smem_u32(&bars->pfull), 0);
    tc_fence_after();
    for (int ks = 0; ks < N1 / 8; ++ks) {
      uint32_t lv = N2 * 16, s8 = 128;
      uint64_t bd = variant == 0 ? smem_desc(smem_u32(sV) + ks * 2 * lv, lv, s8) : smem_desc(smem_u32(sV) + ks * 2 * lv, s8, lv);
      mma_tf32_ts(tmem + 384, tmem + 96 + ks * 8, bd, idesc_tf32(M, N2), ks > 0);
    }
    tc_commit(smem_u32(&bars->mma2));
  }
  mbar_wait(smem_u32(&bars->mma2), 0);
  tc_fence_after();
  uint32_t o[16];
  GP_TMEM_LD16(tmem + lane_off + 384, o);
  tmem_wait_ld();
  for (int c = 0; c < 16; ++c) D2[row * N2 + c] = __uint_as_float(o[c]);
  tc_fence_before(); __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

static float tf32r(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
#include <string.h>

int main(int argc, char** argv) {
  int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int K = 40;
  std::vector<float> A(M * K), B(N1 * K), V(N2 * N1);
  srand(1);
  auto rnd = []() { return tf32r((float)rand() / RAND_MAX * 2.f - 1.f); };
  for (auto& x : A) x = rnd();
  for (auto& x : B) x = rnd();
  for (auto& x : V) x = rnd();
  // K-major no-swizzle packing: [k/4][rows][4]
  std::vector<float> Apk(M * K), Bpk(N1 * K), Vpk(N2 * N1);
  for (int r = 0; r < M; ++r) for (int k = 0; k < K; ++k) Apk[((k / 4) * M + r) * 4 + k % 4] = A[r * K + k];
  for (int r = 0; r < N1; ++r) for (int k = 0; k < K; ++k) Bpk[((k / 4) * N1 + r) * 4 + k % 4] = B[r * K + k];
  for (int c = 0; c < N2; ++c) for (int j = 0; j < N1; ++j) Vpk[((j / 4) * N2 + c) * 4 + j % 4] = V[c * N1 + j];
  float *dA, *dB, *dV, *dD1, *dD2;
  cudaMalloc(&dA, Apk.size() * 4); cudaMalloc(&dB, Bpk.size() * 4); cudaMalloc(&dV, Vpk.size() * 4);
  cudaMalloc(&dD1, M * N1 * 4); cudaMalloc(&dD2, M * N2 * 4);
  cudaMemcpy(dA, Apk.data(), Apk.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, Bpk.data(), Bpk.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dV, Vpk.data(), Vpk.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD1, 0, M * N1 * 4); cudaMemset(dD2, 0, M * N2 * 4);
  int smem = K * M * 4 + K * N1 * 4 + N1 * N2 * 4 + sizeof(Bars) + 64;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe_kernel<<<1, 128, smem>>>(dA, dB, dV, dD1, dD2, K, variant);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("variant %d: CUDA error %s\n", variant, cudaGetErrorString(e)); return 2; }
  std::vector<float> D1(M * N1), D2(M * N2);
  cudaMemcpy(D1.data(), dD1, D1.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(D2.data(), dD2, D2.size() * 4, cudaMemcpyDeviceToHost);
  double e1 = 0, e2 = 0, e2b = 0;
  for (int r = 0; r < M; ++r) {
    std::vector<double> s(N1);
    for (int c = 0; c < N1; ++c) {
      double acc = 0; for (int k = 0; k < K; ++k) acc += (double)A[r * K + k] * B[c * K + k];
      s[c] = acc; e1 = fmax(e1, fabs(acc - D1[r * N1 + c]));
    }
    for (int c = 0; c < N2; ++c) {
      double acc = 0, acc_gpu = 0;
      for (int j = 0; j < N1; ++j) { acc += (double)tf32r(0.5f * (float)s[j]) * V[c * N1 + j]; acc_gpu += (double)tf32r(0.5f * D1[r * N1 + j]) * V[c * N1 + j]; }
      e2 = fmax(e2, fabs(acc - D2[r * N2 + c])); e2b = fmax(e2b, fabs(acc_gpu - D2[r * N2 + c]));
    }
  }
  printf("variant %d: GEMM1(SS) max|err| = %.3e   GEMM2(TS) max|err| = %.3e (vs own D1: %.3e)   %s\n", variant, e1, e2, e2b,
         (e1 < 1e-4 && e2b < 1e-4) ? "OK" : "MISMATCH");
  return (e1 < 1e-4 && e2b < 1e-4) ? 0 : 1;
}
