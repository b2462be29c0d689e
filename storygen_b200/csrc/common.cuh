// Host-side helpers shared by all launchers: tensor-map encoding (driver entry point resolved at run
// time so the .so has no link-time libcuda dependency), SM count, launch counter.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <atomic>

#include "../../include/storygen_b200.h"
#include "ptx.cuh"

namespace sg {

constexpr int kMaxDevices = 64;
int current_device();
int num_sms();  // of the current device
void count_launch(int n = 1);
// one-time (per kernel and device, thread-safe) opt-in to > 48 KB of dynamic shared memory; returns 0 or a cudaError
int ensure_dynamic_smem(const void* kernel, int bytes, std::atomic<uint64_t>* done);

// rank in [2,5]; dims/strides in elements of bf16 -> strides given in BYTES (rank-1 entries, dim0 is
// contiguous). Returns 0 or an SG_ERR_* / cudaError code.
int encode_tmap_bf16(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle);

}  // namespace sg
