// mufu_bench.cu -- measures the MUFU.EX2 issue rate of one SM (the denominator of the fused K.V kernel's roofline) as a
// function of the number of resident warps per SM sub-partition, (a) for pure independent ex2 chains and (b) for the
// epilogue's instruction mix (ex2 + LOP3 + FADD2/2 + F2FP/2 per element).  Standalone: nvcc -arch=sm_100a, run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MIX>
__global__ void k(float* out, const float* in, int iters, long long* cycles) {
  float s[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) s[i] = in[(threadIdx.x + i) & 63];
  uint32_t sink = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      float p0 = ex2(s[i]), p1 = ex2(s[i + 1]);
      if (MIX) {
        uint32_t h0 = __float_as_uint(p0) & 0xFFFFE000u, h1 = __float_as_uint(p1) & 0xFFFFE000u;
        float l0 = p0 - __uint_as_float(h0), l1 = p1 - __uint_as_float(h1);
        uint32_t pk;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(l1), "f"(l0));
        sink ^= pk;               // LOP3 (3-input with the line below)
        sink ^= h0 + h1;          // IADD
        s[i] = p0 - 1.0f; s[i + 1] = p1 - 1.0f;
      } else {
        s[i] = p0 - 1.0f; s[i + 1] = p1 - 1.0f;   // keeps the argument in (-1, 1]: ex2 never over/underflows
      }
    }
  }
  long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) acc += s[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + __uint_as_float(sink & 0xff);
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
  float *out, *in; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&in, 64 * 4); cudaMalloc(&cyc, 148 * 8);
  float h[64]; for (int i = 0; i < 64; ++i) h[i] = -0.01f * i;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  const int iters = 4000;
  for (int mix = 0; mix < 2; ++mix)
    for (int warps : {1, 2, 4, 8, 12, 16, 32}) {
      long long hc[148];
      for (int rep = 0; rep < 2; ++rep) {
        if (mix) k<1><<<148, warps * 32>>>(out, in, iters, cyc); else k<0><<<148, warps * 32>>>(out, in, iters, cyc);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
      double c = 0; for (int i = 0; i < 148; ++i) c += hc[i]; c /= 148;
      double n = (double)warps * 32 * 16 * iters;
      printf("%s warps/SM=%2d (%.2f per SMSP): %.1f cycles -> %.2f ex2/clk/SM\n", mix ? "mix " : "pure", warps, warps / 4.0, c, n / c);
    }
  cudaError_t e = cudaGetLastError();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
