// MUFU throughput of ex2.approx in fp32, f16x2 and bf16x2 form (lane-results per clock per SM).  nvcc -arch=sm_100a ex2_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(int iters, long long* cyc, uint32_t* sink) {
  uint32_t a[8];
  for (int i = 0; i < 8; ++i) a[i] = (MODE == 0) ? __float_as_uint(-0.001f * (threadIdx.x + i)) : (0xbc00bc00u + threadIdx.x * 65537u + i);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(a[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(a[i]));
    }
  }
  long long t1 = clock64();
  uint32_t s = 0;
  for (int i = 0; i < 8; ++i) s ^= a[i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  long long* cyc; uint32_t* sink;
  cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 148 * 1024 * 4);
  const int iters = 2048;
  for (int warps : {4, 8, 16}) {
    for (int mode = 0; mode < 3; ++mode) {
      if (mode == 0) k<0><<<148, warps * 32>>>(iters, cyc, sink);
      if (mode == 1) k<1><<<148, warps * 32>>>(iters, cyc, sink);
      if (mode == 2) k<2><<<148, warps * 32>>>(iters, cyc, sink);
      cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
      const double results = (double)iters * 8 * warps * 32 * (mode == 0 ? 1 : 2);
      printf("[ex2 rate] %s warps/SM %2d: %.2f lane-results/clk/SM (%.0f clk)\n", mode == 0 ? "f32   " : (mode == 1 ? "f16x2 " : "bf16x2"), warps, results / c, c);
    }
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
