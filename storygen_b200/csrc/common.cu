#include "common.cuh"

#include <atomic>
#include <mutex>

namespace sg {

static std::atomic<int64_t> g_launches{0};

int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev;
}

// SM count of the CURRENT device (cached per device: a process may drive several GPUs)
int num_sms() {
  static std::atomic<int> cache[kMaxDevices];
  const int dev = current_device();
  if (dev < 0 || dev >= kMaxDevices) return 148;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE property of a kernel: set it once per (kernel, device).
// `done` is the kernel's bit mask of configured devices; the attribute call is idempotent, so two threads racing on
// the first launch both set it and both succeed.
int ensure_dynamic_smem(const void* kernel, int bytes, std::atomic<uint64_t>* done) {
  const int dev = current_device();
  const uint64_t bit = (dev >= 0 && dev < kMaxDevices) ? (1ull << dev) : 0;
  if (bit != 0 && (done->load(std::memory_order_acquire) & bit)) return 0;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return static_cast<int>(e);
  if (bit != 0) done->fetch_or(bit, std::memory_order_release);
  return 0;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  });
  return fn;
}

int encode_tmap_bf16(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return SG_ERR_DRIVER;
  if (reinterpret_cast<uintptr_t>(ptr) & 15) return SG_ERR_ALIGN;
  cuuint64_t d[5];
  cuuint64_t s[4];
  cuuint32_t b[5];
  cuuint32_t e[5];
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    b[i] = box[i];
    e[i] = 1;
    if (box[i] == 0 || box[i] > 256) return SG_ERR_SHAPE;
  }
  for (int i = 0; i < rank - 1; ++i) {
    s[i] = strides_bytes[i];
    if (s[i] % 16 != 0) return SG_ERR_ALIGN;
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank),
                  const_cast<void*>(ptr), d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return 1000 + static_cast<int>(r);
  return 0;
}

}  // namespace sg

extern "C" int sg_abi_version(void) { return SG_ABI_VERSION; }
extern "C" int sg_sizeof_gemm_args(void) { return static_cast<int>(sizeof(sg_gemm_args)); }
extern "C" int sg_sizeof_attn_args(void) { return static_cast<int>(sizeof(sg_attn_args)); }
extern "C" int64_t sg_launch_count(void) { return sg::g_launches.load(); }
extern "C" const char* sg_error_string(int code) {
  switch (code) {
    case 0: return "ok";
    case SG_ERR_SHAPE: return "unsupported or inconsistent shape";
    case SG_ERR_ALIGN: return "pointer/stride not 16-byte aligned";
    case SG_ERR_DRIVER: return "cuTensorMapEncodeTiled driver entry point not found";
    case SG_ERR_ARG: return "null pointer or bad argument";
    default:
      if (code >= 1000) return "cuTensorMapEncodeTiled failed (code - 1000 = CUresult)";
      if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
      return "unknown error";
  }
}
