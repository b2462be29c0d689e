// Shared helpers of the standalone self-test (host buffers, RNG, reporting).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/storygen_b200.h"

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

inline uint64_t g_seed = 0x1234567ull;
inline float frand() {  // uniform [-1, 1)
  g_seed = g_seed * 6364136223846793005ull + 1442695040888963407ull;
  return static_cast<float>((g_seed >> 40) & 0xFFFFFF) / 8388608.0f - 1.0f;
}

struct HostBf {
  std::vector<float> f;  // bf16-rounded values as float
  std::vector<__nv_bfloat16> h;
  __nv_bfloat16* d = nullptr;
  void init(size_t n, float scale) {
    f.resize(n);
    h.resize(n);
    for (size_t i = 0; i < n; ++i) {
      h[i] = __float2bfloat16(frand() * scale);
      f[i] = __bfloat162float(h[i]);
    }
    CK(cudaMalloc(&d, n * 2));
    CK(cudaMemcpy(d, h.data(), n * 2, cudaMemcpyHostToDevice));
  }
  void zeros(size_t n) {
    f.assign(n, 0.f);
    h.assign(n, __float2bfloat16(0.f));
    CK(cudaMalloc(&d, n * 2));
    CK(cudaMemset(d, 0, n * 2));
  }
  void fetch() {
    CK(cudaMemcpy(h.data(), d, h.size() * 2, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < h.size(); ++i) f[i] = __bfloat162float(h[i]);
  }
  ~HostBf() {
    if (d) cudaFree(d);
  }
};
struct HostF {
  std::vector<float> f;
  float* d = nullptr;
  void init(size_t n, float scale) {
    f.resize(n);
    for (size_t i = 0; i < n; ++i) f[i] = frand() * scale;
    CK(cudaMalloc(&d, n * 4));
    CK(cudaMemcpy(d, f.data(), n * 4, cudaMemcpyHostToDevice));
  }
  void zeros(size_t n) {
    f.assign(n, 0.f);
    CK(cudaMalloc(&d, n * 4));
    CK(cudaMemset(d, 0, n * 4));
  }
  void fetch() { CK(cudaMemcpy(f.data(), d, f.size() * 4, cudaMemcpyDeviceToHost)); }
  ~HostF() {
    if (d) cudaFree(d);
  }
};

inline int g_fail = 0;
inline const char* g_filter = nullptr;
inline bool want(const char* name) { return g_filter == nullptr || strstr(name, g_filter) != nullptr; }

inline void report(const char* name, int rc, double max_err, double max_ref, double tol) {
  cudaError_t e = cudaDeviceSynchronize();
  bool ok = (rc == 0) && (e == cudaSuccess) && (max_err <= tol) && (max_err == max_err);
  printf("[%s] %-44s rc=%d cuda=%s max_err=%.4g (ref max %.4g, tol %.3g)\n", ok ? "PASS" : "FAIL", name,
         rc, cudaGetErrorName(e), max_err, max_ref, tol);
  fflush(stdout);
  if (!ok) ++g_fail;
  if (e != cudaSuccess) {
    printf("sticky CUDA error, aborting remaining cases\n");
    exit(3);
  }
}

