// Attention backward for sm_100a (training of StoryGen's image cross-attention, train_StorySalon_stage2.py:305-332):
//   dV = P^T dO,   dS = scale * P o (dO V^T - delta),   dK = dS^T Q,   dQ = dS K,      P = 2^(scale_log2 S - lse2),
// FlashAttention-2 style: one CTA per (128-key tile, head, lane) looping over the 128-query tiles; P is recomputed from
// the forward's row statistics, nothing of size Nq x Nk ever exists in memory.
//
//   warp 0     : TMA        K_j, V_j once; Q_i, dO_i through a (1 or 2 deep) ring
//   warp 1     : MMA issuer S = Q_i K_j^T and dP = dO_i V_j^T (SS, M = 128 queries, N = 128 keys);  then, once the compute
//                           warps have published P and dS (bf16, in shared memory):
//                           dV += P^T dO_i, dK += dS^T Q_i (A = the [query x key] tile read MN-major, M = 128 keys),
//                           dQ_i = dS K_j (A = the same tile read K-major, M = 128 queries)
//   warps 4..7 : compute    one thread per query row: S / dP rows from TMEM in 32-key chunks, P = exp2(...),
//                           dS = scale P (dP - delta), bf16 -> swizzled shared memory; dQ_i rows from TMEM ->
//                           red.global.add.v4.f32 into the caller's fp32 accumulator (query tiles receive contributions
//                           from every key tile); at the end dK_j / dV_j rows (thread = key row) -> bf16 -> global
// TMEM: S | dP | dV | dK | dQ (dQ aliases S when 3 x head dim does not fit beside them).
// Channels past the head's own inside the last 64-channel swizzle atom belong to the next head: the padded 16-byte chunk of
// Q_i and dO_i is zeroed in shared memory (so S, dP, dK, dV see zeros there) and the padded dQ columns are never stored.
#include "common.cuh"

namespace sg {

constexpr int kBwdAtom = 128 * 128;  // [128 rows x 64 bf16], swizzle 128B

struct AttnBwdKParams {
  CUtensorMap q_map, k_map, v_map, do_map;
  const float* lse;    // [lanes][heads][nq] log2 domain
  const float* delta;  // [lanes][heads][nq]
  float* dq_accum;     // fp32 [lanes][nq][C]
  bf16* dk;
  bf16* dv;
  long long dk_stride_n, dk_stride_t;
  int nq, nk, n_heads;
  float scale, scale_log2;
};

template <int D>
struct BwdCfg {
  static constexpr int NA = (D + 63) / 64;
  static constexpr int KQ = (D + 15) / 16;       // k-steps of S / dP
  static constexpr int DP = (D + 15) & ~15;      // accumulator columns of dV / dK / dQ
  static constexpr int kTile = NA * kBwdAtom;    // one [128 x D] operand tile
  static constexpr int kQStages = (NA == 1) ? 2 : 1;
  static constexpr int kColS = 0, kColDP = 128, kColDV = 256, kColDK = 256 + DP;
  static constexpr bool kAliasDQ = (256 + 3 * DP > 512);
  static constexpr int kColDQ = kAliasDQ ? 0 : 256 + 2 * DP;
  static constexpr int kSmem = 2 * kTile /*K V*/ + 2 * kQStages * kTile /*Q dO*/ + 4 * kBwdAtom /*P dS*/ + 1024 + 256;
  static_assert(256 + 2 * DP <= 512, "head dim too large for the TMEM layout");
  static_assert(kSmem <= 231424, "shared memory");
};

template <int D>
__global__ void __launch_bounds__(256, 1) attn_bwd_kernel(const __grid_constant__ AttnBwdKParams p) {
  using Cfg = BwdCfg<D>;
  constexpr int NA = Cfg::NA, KQ = Cfg::KQ, DP = Cfg::DP, QS = Cfg::kQStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* k_smem = smem;
  uint8_t* v_smem = k_smem + Cfg::kTile;
  uint8_t* q_smem = v_smem + Cfg::kTile;            // [QS]
  uint8_t* do_smem = q_smem + QS * Cfg::kTile;      // [QS]
  uint8_t* p_smem = do_smem + QS * Cfg::kTile;      // [128 q x 128 k] as 2 atoms
  uint8_t* ds_smem = p_smem + 2 * kBwdAtom;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ds_smem + 2 * kBwdAtom);
  uint64_t* kv_full = bars;            // [1]
  uint64_t* qd_full = bars + 1;        // [2]
  uint64_t* qd_ready = bars + 3;       // [2] pads zeroed
  uint64_t* qd_empty = bars + 5;       // [2]
  uint64_t* sp_full = bars + 7;        // S, dP complete
  uint64_t* pds_ready = bars + 8;      // P, dS in shared memory (and S / dP read)
  uint64_t* acc_done = bars + 9;       // dV / dK / dQ MMAs of the tile complete
  uint64_t* dq_free = bars + 10;       // dQ tile read
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int j0 = blockIdx.x * 128;  // first key of this CTA
  const int head = blockIdx.y, img = blockIdx.z;
  const int TQ = (p.nq + 127) / 128;

  if (threadIdx.x == 0) {
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qd_full[i], 1);
      mbar_init(&qd_ready[i], 4);
      mbar_init(&qd_empty[i], 1);
    }
    mbar_init(sp_full, 1);
    mbar_init(pds_ready, 4);
    mbar_init(acc_done, 1);
    mbar_init(dq_free, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    // ===================== TMA =====================
    mbar_arrive_expect_tx_w(kv_full, 2 * Cfg::kTile);
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      tma_load_3d_w(k_smem + a * kBwdAtom, &p.k_map, kv_full, head * D + 64 * a, j0, img);
      tma_load_3d_w(v_smem + a * kBwdAtom, &p.v_map, kv_full, head * D + 64 * a, j0, img);
    }
    for (int i = 0; i < TQ; ++i) {
      const int s = i % QS;
      mbar_wait(&qd_empty[s], (((i / QS) & 1) ^ 1));
      mbar_arrive_expect_tx_w(&qd_full[s], 2 * Cfg::kTile);
#pragma unroll
      for (int a = 0; a < NA; ++a) {
        tma_load_3d_w(q_smem + s * Cfg::kTile + a * kBwdAtom, &p.q_map, &qd_full[s], head * D + 64 * a, i * 128, img);
        tma_load_3d_w(do_smem + s * Cfg::kTile + a * kBwdAtom, &p.do_map, &qd_full[s], head * D + 64 * a, i * 128, img);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
    auto idesc_acc = [](int a, uint32_t a_mn) {  // N = the head's channels inside atom a, rounded to 16
      const int rem = DP - 64 * a;
      return make_idesc_bf16(128, static_cast<uint32_t>(rem >= 64 ? 64 : rem), a_mn, 1);
    };
    mbar_wait(kv_full, 0);
    const uint32_t k_addr = smem_u32(k_smem), v_addr = smem_u32(v_smem), p_addr = smem_u32(p_smem), ds_addr = smem_u32(ds_smem);
    for (int i = 0; i < TQ; ++i) {
      const int s = i % QS;
      const uint32_t q_addr = smem_u32(q_smem + s * Cfg::kTile), do_addr = smem_u32(do_smem + s * Cfg::kTile);
      mbar_wait(&qd_ready[s], (i / QS) & 1);
      if (Cfg::kAliasDQ && i > 0) mbar_wait(dq_free, (i - 1) & 1);  // dQ_{i-1} lives in S's columns
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < KQ; ++ks) {
        const uint32_t off = (ks >> 2) * kBwdAtom + (ks & 3) * 32;
        umma_ss_w(tmem_base + Cfg::kColS, make_smem_desc(q_addr + off, 16, 1024, kLayoutSW128),
                  make_smem_desc(k_addr + off, 16, 1024, kLayoutSW128), idesc_s, ks != 0);
      }
#pragma unroll
      for (int ks = 0; ks < KQ; ++ks) {
        const uint32_t off = (ks >> 2) * kBwdAtom + (ks & 3) * 32;
        umma_ss_w(tmem_base + Cfg::kColDP, make_smem_desc(do_addr + off, 16, 1024, kLayoutSW128),
                  make_smem_desc(v_addr + off, 16, 1024, kLayoutSW128), idesc_s, ks != 0);
      }
      umma_commit_w(sp_full);
      mbar_wait(pds_ready, i & 1);
      if (!Cfg::kAliasDQ && i > 0) mbar_wait(dq_free, (i - 1) & 1);
      tc_fence_after();
      // dV += P^T dO_i ; dK += dS^T Q_i   (A read MN-major: M = keys, K = the tile's 128 query rows, 16 per k-step)
#pragma unroll
      for (int a = 0; a < NA; ++a) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          umma_ss_w(tmem_base + Cfg::kColDV + 64 * a, make_smem_desc(p_addr + ks * 2048, kBwdAtom, 1024, kLayoutSW128),
                    make_smem_desc(do_addr + a * kBwdAtom + ks * 2048, kBwdAtom, 1024, kLayoutSW128), idesc_acc(a, 1),
                    !(i == 0 && ks == 0));
        }
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          umma_ss_w(tmem_base + Cfg::kColDK + 64 * a, make_smem_desc(ds_addr + ks * 2048, kBwdAtom, 1024, kLayoutSW128),
                    make_smem_desc(q_addr + a * kBwdAtom + ks * 2048, kBwdAtom, 1024, kLayoutSW128), idesc_acc(a, 1),
                    !(i == 0 && ks == 0));
        }
        // dQ_i = dS K_j   (A read K-major: M = queries, K = the 128 keys)
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t off = (ks >> 2) * kBwdAtom + (ks & 3) * 32;
          umma_ss_w(tmem_base + Cfg::kColDQ + 64 * a, make_smem_desc(ds_addr + off, 16, 1024, kLayoutSW128),
                    make_smem_desc(k_addr + a * kBwdAtom + ks * 2048, kBwdAtom, 1024, kLayoutSW128), idesc_acc(a, 0), ks != 0);
        }
      }
      umma_commit_w(acc_done);
      umma_commit_w(&qd_empty[s]);
    }
  } else if (warp >= 4) {
    // ===================== compute =====================
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;  // query row of the tile (loop) / key row of the CTA (epilogue) == TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t p_row = smem_u32(p_smem) + r * 128, ds_row = smem_u32(ds_smem) + r * 128;
    const int C = p.n_heads * D;
    auto zero_pads = [&](int i) {
      const int s = i % QS;
      mbar_wait(&qd_full[s], (i / QS) & 1);
      if constexpr (D % 16 != 0) {
        constexpr int atom = D / 64, chunk = (D % 64) / 8;
        const uint32_t off = s * Cfg::kTile + atom * kBwdAtom + r * 128 + ((chunk ^ (r & 7)) << 4);
        sts128(smem_u32(q_smem) + off, make_uint4(0, 0, 0, 0));
        sts128(smem_u32(do_smem) + off, make_uint4(0, 0, 0, 0));
        fence_proxy_async_smem();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&qd_ready[s]);
    };
    zero_pads(0);
    for (int i = 0; i < TQ; ++i) {
      const int q = i * 128 + r;
      const bool q_ok = q < p.nq;
      const long long srow = (static_cast<long long>(img) * p.n_heads + head) * p.nq + q;
      // rows past nq: lse = +inf -> P = dS = 0
      const float lse2 = q_ok ? __ldg(p.lse + srow) : INFINITY;
      const float delta = q_ok ? __ldg(p.delta + srow) : 0.f;
      mbar_wait(sp_full, i & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t sv[32], dv[32];
        tmem_ld32(tmem_base + lane_base + Cfg::kColS + c * 32, sv);
        tmem_ld32(tmem_base + lane_base + Cfg::kColDP + c * 32, dv);
        tmem_ld_wait();
        uint32_t pw[16], dw[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          const float p0 = fast_exp2(fmaf(__uint_as_float(sv[2 * t]), p.scale_log2, -lse2));
          const float p1 = fast_exp2(fmaf(__uint_as_float(sv[2 * t + 1]), p.scale_log2, -lse2));
          const float d0 = p0 * (__uint_as_float(dv[2 * t]) - delta) * p.scale;
          const float d1 = p1 * (__uint_as_float(dv[2 * t + 1]) - delta) * p.scale;
          pw[t] = pack_bf16x2(p0, p1);
          dw[t] = pack_bf16x2(d0, d1);
        }
        // keys [32c, 32c + 32) of this query row: atom c >> 1, 16-byte chunks (c & 1) * 4 + t
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const uint32_t off = (c >> 1) * kBwdAtom + ((((c & 1) * 4 + t) ^ (r & 7)) << 4);
          sts128(p_row + off, make_uint4(pw[4 * t], pw[4 * t + 1], pw[4 * t + 2], pw[4 * t + 3]));
          sts128(ds_row + off, make_uint4(dw[4 * t], dw[4 * t + 1], dw[4 * t + 2], dw[4 * t + 3]));
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_ready);
      if (i + 1 < TQ) zero_pads(i + 1);  // lets the tensor core start S / dP of the next tile behind this tile's MMAs
      // dQ_i rows -> fp32 accumulator
      mbar_wait(acc_done, i & 1);
      tc_fence_after();
      float* dq_row = p.dq_accum + (static_cast<long long>(img) * p.nq + q) * C + head * D;
#pragma unroll
      for (int c = 0; c < DP / 16; ++c) {
        uint32_t qv[16];
        tmem_ld16(tmem_base + lane_base + Cfg::kColDQ + c * 16, qv);
        tmem_ld_wait();
        if (q_ok) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (c * 16 + g * 4 < D) {
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dq_row + c * 16 + g * 4),
                           "f"(__uint_as_float(qv[4 * g])), "f"(__uint_as_float(qv[4 * g + 1])),
                           "f"(__uint_as_float(qv[4 * g + 2])), "f"(__uint_as_float(qv[4 * g + 3]))
                           : "memory");
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_free);
    }
    // dK_j, dV_j rows (thread = key row) -> bf16
    const int key = j0 + r;
    bf16* dk_row = p.dk + static_cast<long long>(img) * p.dk_stride_n + static_cast<long long>(key) * p.dk_stride_t + head * D;
    bf16* dv_row = p.dv + static_cast<long long>(img) * p.dk_stride_n + static_cast<long long>(key) * p.dk_stride_t + head * D;
#pragma unroll
    for (int c = 0; c < DP / 16; ++c) {
      uint32_t kv_[16], vv[16];
      tmem_ld16(tmem_base + lane_base + Cfg::kColDK + c * 16, kv_);
      tmem_ld16(tmem_base + lane_base + Cfg::kColDV + c * 16, vv);
      tmem_ld_wait();
      if (key < p.nk) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (c * 16 + g * 8 < D) {
            uint4 a, b;
            a.x = pack_bf16x2(__uint_as_float(kv_[8 * g + 0]), __uint_as_float(kv_[8 * g + 1]));
            a.y = pack_bf16x2(__uint_as_float(kv_[8 * g + 2]), __uint_as_float(kv_[8 * g + 3]));
            a.z = pack_bf16x2(__uint_as_float(kv_[8 * g + 4]), __uint_as_float(kv_[8 * g + 5]));
            a.w = pack_bf16x2(__uint_as_float(kv_[8 * g + 6]), __uint_as_float(kv_[8 * g + 7]));
            b.x = pack_bf16x2(__uint_as_float(vv[8 * g + 0]), __uint_as_float(vv[8 * g + 1]));
            b.y = pack_bf16x2(__uint_as_float(vv[8 * g + 2]), __uint_as_float(vv[8 * g + 3]));
            b.z = pack_bf16x2(__uint_as_float(vv[8 * g + 4]), __uint_as_float(vv[8 * g + 5]));
            b.w = pack_bf16x2(__uint_as_float(vv[8 * g + 6]), __uint_as_float(vv[8 * g + 7]));
            *reinterpret_cast<uint4*>(dk_row + c * 16 + g * 8) = a;
            *reinterpret_cast<uint4*>(dv_row + c * 16 + g * 8) = b;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// delta[n, h, q] = sum_c dO[n, q, h D + c] * O[n, q, h D + c]   (one thread per (n, q, h); consecutive threads = heads)
__global__ void attn_bwd_delta_kernel(const bf16* __restrict__ o, const bf16* __restrict__ d_o, long long stride_n,
                                      long long stride_t, int n_lanes, int n_heads, int head_dim, int nq,
                                      float* __restrict__ delta) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(n_lanes) * nq * n_heads;
  if (idx >= total) return;
  const int h = static_cast<int>(idx % n_heads);
  const long long nq_i = idx / n_heads;
  const int q = static_cast<int>(nq_i % nq), n = static_cast<int>(nq_i / nq);
  const long long off = n * stride_n + q * stride_t + static_cast<long long>(h) * head_dim;
  float s = 0.f;
  for (int c = 0; c < head_dim; c += 8) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(o + off + c));
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(d_o + off + c));
    const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2 = unpack_bf16x2(a.z), a3 = unpack_bf16x2(a.w);
    const float2 b0 = unpack_bf16x2(b.x), b1 = unpack_bf16x2(b.y), b2 = unpack_bf16x2(b.z), b3 = unpack_bf16x2(b.w);
    s += a0.x * b0.x + a0.y * b0.y + a1.x * b1.x + a1.y * b1.y + a2.x * b2.x + a2.y * b2.y + a3.x * b3.x + a3.y * b3.y;
  }
  delta[(static_cast<long long>(n) * n_heads + h) * nq + q] = s;
}

template <int D>
static int launch_attn_bwd(const AttnBwdKParams& kp, int nk, int heads, int lanes, cudaStream_t stream) {
  using Cfg = BwdCfg<D>;
  static std::atomic<uint64_t> configured{0};
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(attn_bwd_kernel<D>), Cfg::kSmem, &configured)) return rc;
  dim3 grid((nk + 127) / 128, heads, lanes);
  attn_bwd_kernel<D><<<grid, 256, Cfg::kSmem, stream>>>(kp);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

static int make_tok_map_bwd(CUtensorMap* m, const void* ptr, int channels, int tokens, int lanes, int64_t stride_t,
                            int64_t stride_n) {
  uint64_t dims[3] = {static_cast<uint64_t>(channels), static_cast<uint64_t>(tokens), static_cast<uint64_t>(lanes)};
  uint64_t strides[2] = {static_cast<uint64_t>(stride_t) * 2, static_cast<uint64_t>(stride_n) * 2};
  uint32_t box[3] = {64, 128, 1};
  return encode_tmap_bf16(m, ptr, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace sg

using namespace sg;

extern "C" int sg_attention_bwd(const sg_attn_bwd_args* a, void* stream_v) {
  if (a == nullptr || !a->q || !a->k || !a->v || !a->o || !a->d_o || !a->lse || !a->delta || !a->dq_accum || !a->dk || !a->dv)
    return SG_ERR_ARG;
  if (a->nq < 1 || a->nk < 1 || a->n_lanes < 1 || a->n_heads < 1) return SG_ERR_SHAPE;
  switch (a->head_dim) {
    case 32: case 40: case 64: case 80: case 128: break;
    default: return SG_ERR_SHAPE;
  }
  if ((reinterpret_cast<uintptr_t>(a->dk) & 15) || (reinterpret_cast<uintptr_t>(a->dv) & 15) || (a->dk_stride_t % 8) ||
      (a->dk_stride_n % 8) || (reinterpret_cast<uintptr_t>(a->dq_accum) & 15) || (a->o_stride_t % 8) || (a->o_stride_n % 8) ||
      (reinterpret_cast<uintptr_t>(a->o) & 15) || (reinterpret_cast<uintptr_t>(a->d_o) & 15))
    return SG_ERR_ALIGN;
  const int C = a->n_heads * a->head_dim;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  {
    const long long total = static_cast<long long>(a->n_lanes) * a->nq * a->n_heads;
    attn_bwd_delta_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
        static_cast<const bf16*>(a->o), static_cast<const bf16*>(a->d_o), a->o_stride_n, a->o_stride_t, a->n_lanes, a->n_heads,
        a->head_dim, a->nq, a->delta);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  AttnBwdKParams kp;
  memset(&kp, 0, sizeof(kp));
  int rc;
  if ((rc = make_tok_map_bwd(&kp.q_map, a->q, C, a->nq, a->n_lanes, a->q_stride_t, a->q_stride_n)) != 0) return rc;
  if ((rc = make_tok_map_bwd(&kp.k_map, a->k, C, a->nk, a->n_lanes, a->k_stride_t, a->k_stride_n)) != 0) return rc;
  if ((rc = make_tok_map_bwd(&kp.v_map, a->v, C, a->nk, a->n_lanes, a->v_stride_t, a->v_stride_n)) != 0) return rc;
  if ((rc = make_tok_map_bwd(&kp.do_map, a->d_o, C, a->nq, a->n_lanes, a->o_stride_t, a->o_stride_n)) != 0) return rc;
  kp.lse = a->lse;
  kp.delta = a->delta;
  kp.dq_accum = a->dq_accum;
  kp.dk = static_cast<bf16*>(a->dk);
  kp.dv = static_cast<bf16*>(a->dv);
  kp.dk_stride_n = a->dk_stride_n;
  kp.dk_stride_t = a->dk_stride_t;
  kp.nq = a->nq;
  kp.nk = a->nk;
  kp.n_heads = a->n_heads;
  kp.scale = a->scale;
  kp.scale_log2 = a->scale * 1.4426950408889634f;
  switch (a->head_dim) {
    case 32: return launch_attn_bwd<32>(kp, a->nk, a->n_heads, a->n_lanes, stream);
    case 40: return launch_attn_bwd<40>(kp, a->nk, a->n_heads, a->n_lanes, stream);
    case 64: return launch_attn_bwd<64>(kp, a->nk, a->n_heads, a->n_lanes, stream);
    case 80: return launch_attn_bwd<80>(kp, a->nk, a->n_heads, a->n_lanes, stream);
    default: return launch_attn_bwd<128>(kp, a->nk, a->n_heads, a->n_lanes, stream);
  }
}
