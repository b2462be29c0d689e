// Persistent warp-specialised tcgen05 GEMM / implicit-GEMM convolution for sm_100a.
//
//   warp 0      : TMA producer  (A tiles gathered from NHWC sources with per-segment pixel shifts,
//                                B = pre-packed weights), kStages-deep smem ring
//   warp 1      : MMA issuer    (one elected lane; tcgen05.mma M=128 x N=BN x K=16, fp32 in TMEM,
//                                two accumulator stages so the epilogue of tile i overlaps tile i+1)
//   warps 2..9  : epilogue      two warpgroups, each thread owns one accumulator row; the 32-column chunks
//                               of a tile alternate between the groups (2 warps per SM sub-partition hide
//                               each other's latency); residual rows, bias vectors and LayerNorm statistics are
//                               fetched before the accumulator is awaited; tcgen05.ld (double-buffered) ->
//                               LN correction / bias / time-embedding / residual / GEGLU / statistics in registers
//                               -> swizzled smem slab -> one TMA bulk store per 32-row x 32-column chunk
//
// Tile = 128 output pixels (a bw x bh x bn box of the NHWC output grid) x BN output channels.
// K loop = list of segments (source tensor, dx, dy, first channel, #64-channel blocks): this one
// kernel is Linear, 1x1 conv, 3x3 conv (stride 1 or, via phase-strided sources, stride 2), conv over
// a channel-concat of two tensors, and conv + fused 1x1 shortcut.  See include/storygen_b200.h.
#include <stdlib.h>

#include "gemm_common.cuh"

namespace sg {

// CG = 2: the CTA-pair form.  Two CTAs of a cluster (the two SMs of a TPC) work on two vertically adjacent M tiles and the
// same N tile: each loads its own A tile and HALF of the B tile, the leader (cluster rank 0) issues tcgen05.mma.cta_group::2
// with M = 256 and commits to the barriers of both CTAs, each CTA's epilogue drains its own 128 accumulator rows.  Per
// k-block a CTA stages 16 KB + 64 B x BN instead of 16 KB + 128 B x BN: the operand ingest that bounds the K-heavy
// convolutions (and a good part of their power) drops by a quarter to a third.
template <int BN, int MODE, int RT, int CG, int EG>
__device__ __forceinline__ void gemm_kernel_impl(const GemmKParams& p) {
  using Cfg = GemmCfg<BN, RT, CG, EG>;
  constexpr int kEpiWarps = Cfg::kEpiWarps;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for SWIZZLE_128B tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + Cfg::kStages * Cfg::kABytes;
  uint8_t* epi_scratch = smem + Cfg::kStages * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_scratch + Cfg::kScratchBytes);
  uint64_t* full_bar = bars;                      // [kStages]
  uint64_t* empty_bar = bars + Cfg::kStages;      // [kStages]
  uint64_t* tfull_bar = bars + 2 * Cfg::kStages;  // [2]
  uint64_t* tempty_bar = tfull_bar + 2;           // [2]
  uint64_t* res_bar = tempty_bar + 2;             // [kEpiWarps] residual slabs landed (RT = 1)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + kEpiWarps);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // broadcast: the compiler treats it as warp-uniform
  const int lane = threadIdx.x & 31;
  const int num_m_tiles = p.tiles_x * p.tiles_y * p.tiles_n;
  // work items: CG = 1: (M tile, N tile); CG = 2: (pair of M tiles, N tile), this CTA taking M tile 2 * pair + rank
  const int num_tiles = ((num_m_tiles + CG - 1) / CG) * p.num_n_tiles;
  const int cta_rank = (CG == 2) ? static_cast<int>(cluster_ctarank()) : 0;
  const int first_item = blockIdx.x / CG, item_step = gridDim.x / CG;

  if (threadIdx.x == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kEpiWarps * CG);  // one arrive per epilogue warp (of both CTAs of a pair)
    }
    for (int i = 0; i < kEpiWarps; ++i) mbar_init(&res_bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (CG == 2) tmem_alloc2(tmem_slot, Cfg::kTmemCols);
    else tmem_alloc(tmem_slot, Cfg::kTmemCols);
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all();  // the peer's barriers are initialised before anything is sent to them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);  // warp-uniform for the compiler

  if (warp == 0) {
    // ===================== TMA producer (whole warp, one elected lane issues: see ptx.cuh) =====================
    if (lane == 0) tma_prefetch_desc(&p.b_map);
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = first_item; tile < num_tiles; tile += item_step) {
      const int mt = (tile / p.num_n_tiles) * CG + cta_rank;  // past the last M tile (odd count): TMA zero-fills
      const int nt = tile % p.num_n_tiles;
      const int tx = mt % p.tiles_x;
      const int ty = (mt / p.tiles_x) % p.tiles_y;
      const int tn = mt / (p.tiles_x * p.tiles_y);
      const int x0 = tx * p.bw, y0 = ty * p.bh, n0 = tn * p.bn;
      // CG = 2: this CTA's half of the N tile's (possibly ragged) columns
      const int ncols = (p.N - nt * BN) < BN ? (p.N - nt * BN) : BN;
      const int b_row0 = nt * BN + cta_rank * (ncols / 2);
      int kb = 0;
      for (int s = 0; s < p.num_segs; ++s) {
        const sg_seg seg = p.segs[s];
        const CUtensorMap* amap = &p.a_maps[seg.src];
        for (int j = 0; j < seg.nkb; ++j, ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if constexpr (CG == 2) {
            // both CTAs' bytes complete on the LEADER's barrier
            if (cta_rank == 0) mbar_arrive_expect_tx_w(&full_bar[stage], 2 * Cfg::kStageBytes);
            const uint32_t lb = leader_bar_addr(&full_bar[stage]);
            tma_load_4d_2cta_w(smem_a + stage * Cfg::kABytes, amap, lb, seg.c0 + j * kBK, x0 + seg.dx, y0 + seg.dy, n0);
            tma_load_2d_2cta_w(smem_b + stage * Cfg::kBBytes, &p.b_map, lb, kb * kBK, b_row0);
            if (++stage == Cfg::kStages) {
              stage = 0;
              phase ^= 1;
            }
            continue;
          }
          mbar_arrive_expect_tx_w(&full_bar[stage], Cfg::kStageBytes);
          tma_load_4d_w(smem_a + stage * Cfg::kABytes, amap, &full_bar[stage], seg.c0 + j * kBK, x0 + seg.dx,
                        y0 + seg.dy, n0);
          tma_load_2d_w(smem_b + stage * Cfg::kBBytes, &p.b_map, &full_bar[stage], kb * kBK, nt * BN);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp, one elected lane issues; CG = 2: the leader CTA only) =====================
    if (CG == 1 || cta_rank == 0) {
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    int mma_tile = 0;
    const uint32_t a_addr0 = smem_u32(smem_a), b_addr0 = smem_u32(smem_b);
    for (int tile = first_item; tile < num_tiles; tile += item_step) {
      // the last N tile may be ragged (e.g. N = 640 as 256 + 256 + 128): issue only the columns that exist
      const int nt_mma = tile % p.num_n_tiles;
      const int ncols = (p.N - nt_mma * BN) < BN ? (p.N - nt_mma * BN) : BN;
      const uint32_t idesc = make_idesc_bf16(kBM * CG, static_cast<uint32_t>((ncols + 15) & ~15), 0, 0);
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      if (p.dbg != nullptr && blockIdx.x == 0 && lane == 0 && mma_tile < 8) p.dbg[(1 * 8 + mma_tile) * 8 + 0] = clock64();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < p.total_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint64_t a_desc = make_smem_desc(a_addr0 + stage * Cfg::kABytes, 16, 1024, kLayoutSW128);
        const uint64_t b_desc = make_smem_desc(b_addr0 + stage * Cfg::kBBytes, 16, 1024, kLayoutSW128);
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k) {
          // advance 16 bf16 = 32 bytes inside the swizzle atom: +2 in the (>>4) address field
          if constexpr (CG == 2) umma_ss2_w(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
          else umma_ss_w(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
        }
        if constexpr (CG == 2) umma_commit2_w(&empty_bar[stage]);
        else umma_commit_w(&empty_bar[stage]);
        if (++stage == Cfg::kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if constexpr (CG == 2) umma_commit2_w(&tfull_bar[acc]);
      else umma_commit_w(&tfull_bar[acc]);
      if (p.dbg != nullptr && blockIdx.x == 0 && lane == 0 && mma_tile < 8) p.dbg[(1 * 8 + mma_tile) * 8 + 1] = clock64();
      ++mma_tile;
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    }
  } else {
    // ===================== epilogue (see gemm_common.cuh::epilogue_tile) =====================
    const int quarter = warp & 3;     // TMEM lane quarter this warp may access
    const int grp = (warp - 2) >> 2;  // epilogue warpgroup 0 .. EG-1
    uint8_t* scratch = epi_scratch + (warp - 2) * Cfg::kWarpScratch;
    int acc = 0;
    uint32_t acc_phase = 0;
    int tile_ord = 0;
    EpiState est;
    if (lane == 0) {
      tma_prefetch_desc(&p.out_map);
      if (RT == 1) tma_prefetch_desc(&p.res_map);
    }
    for (int tile = first_item; tile < num_tiles; tile += item_step) {
      const int mt = (tile / p.num_n_tiles) * CG + cta_rank;
      const int nt = tile % p.num_n_tiles;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
      if (CG == 1 || mt < num_m_tiles) {
        epilogue_tile<BN, MODE, RT, CG, EG>(p, mt, nt, t_row, scratch, quarter, grp, lane, &tfull_bar[acc], acc_phase, est,
                                        &res_bar[warp - 2], tile_ord, warp);
      } else {
        mbar_wait(&tfull_bar[acc], acc_phase);  // odd M-tile count: the pair's second half has no rows
      }
      ++tile_ord;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2 && cta_rank != 0) mbar_arrive_cluster(&tempty_bar[acc], 0);  // the leader's MMA warp waits for both CTAs
        else mbar_arrive(&tempty_bar[acc]);
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    // the staging slabs must outlive the TMA stores that read them
    if (lane == 0) bulk_wait_group_read<0>();
  }

  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all();  // the peer may still read this CTA's operands / signal its barriers
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc2(tmem_base, Cfg::kTmemCols);
    else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

constexpr int kWideLinear = 3, kWideGeglu = 4;  // epilogue warpgroups of the small-K kernels (pick_epi_groups)

template <int BN, int MODE, int RT, int EG>
__global__ void __launch_bounds__(64 + 128 * EG, 1) gemm_kernel(const __grid_constant__ GemmKParams p) {
  gemm_kernel_impl<BN, MODE, RT, 1, EG>(p);
}
template <int BN, int MODE, int EG>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 128 * EG, 1) gemm_pair_kernel(const __grid_constant__ GemmKParams p) {
  gemm_kernel_impl<BN, MODE, 0, 2, EG>(p);
}

template <int BN, int MODE, int RT = 0, int EG = 2>
static int launch_gemm_mode(const GemmKParams& kp, int num_tiles, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, RT, 1, EG>;
  static std::atomic<uint64_t> configured{0};
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(gemm_kernel<BN, MODE, RT, EG>), Cfg::kSmemBytes, &configured)) return rc;
  int grid = num_tiles < num_sms() ? num_tiles : num_sms();
  gemm_kernel<BN, MODE, RT, EG><<<grid, Cfg::kThreads, Cfg::kSmemBytes, stream>>>(kp);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

template <int BN, int MODE, int EG = 2>
static int launch_gemm_pair(const GemmKParams& kp, int num_pair_tiles, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, 0, 2, EG>;
  static std::atomic<uint64_t> configured{0};
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(gemm_pair_kernel<BN, MODE, EG>), Cfg::kSmemBytes, &configured)) return rc;
  const int pairs = num_sms() / 2;
  const int grid = 2 * (num_pair_tiles < pairs ? num_pair_tiles : pairs);
  gemm_pair_kernel<BN, MODE, EG><<<grid, Cfg::kThreads, Cfg::kSmemBytes, stream>>>(kp);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

// eg > 2: the wide-epilogue instantiations (pick_epi_groups)
template <int BN>
static int launch_gemm(const GemmKParams& kp, int num_tiles, bool res_tma, int eg, cudaStream_t stream) {
  if constexpr (BN == 160) {
    if (kp.mode == SG_EPI_LINEAR && eg == kWideLinear)
      return res_tma ? launch_gemm_mode<BN, SG_EPI_LINEAR, 1, kWideLinear>(kp, num_tiles, stream)
                     : launch_gemm_mode<BN, SG_EPI_LINEAR, 0, kWideLinear>(kp, num_tiles, stream);
  }
  if constexpr (BN == 256) {
    if (kp.mode == SG_EPI_GEGLU && eg == kWideGeglu) return launch_gemm_mode<BN, SG_EPI_GEGLU, 0, kWideGeglu>(kp, num_tiles, stream);
  }
  if constexpr (BN <= 160) {
    if (res_tma && kp.mode == SG_EPI_LINEAR) return launch_gemm_mode<BN, SG_EPI_LINEAR, 1>(kp, num_tiles, stream);
  }
  switch (kp.mode) {
    case SG_EPI_GEGLU: return launch_gemm_mode<BN, SG_EPI_GEGLU>(kp, num_tiles, stream);
    case SG_EPI_NCHW_F32:
      if constexpr (BN == 64) return launch_gemm_mode<BN, SG_EPI_NCHW_F32>(kp, num_tiles, stream);
      return SG_ERR_SHAPE;
    default: return launch_gemm_mode<BN, SG_EPI_LINEAR>(kp, num_tiles, stream);
  }
}

// N tile width: the main loop is bound by the bytes an SM must pull into shared memory per MMA (A 16 KB + B 128 B x BN
// per k-block, ~80-100 B/clk/SM achievable: profiles/r01_ubench_tma_ingest.log), so wide tiles win: 256 with a
// ragged last tile from N = 512 on.  Small problems (deep UNet levels: M = 768 / 1536 rows) are the exception: when
// the tile count would leave more than half of the SMs idle, halve the tile width until the grid fills them.
static int pick_bn_tile(const sg_gemm_args* a) {
  int bn_tile = (a->N % 256 == 0 || a->N >= 512) ? 256 : ((a->N % 160 == 0) ? 160 : ((a->N % 128 == 0) ? 128 : 64));
  // small K: the epilogue, not the main loop, bounds the tile; 160-wide tiles keep two staging slabs per epilogue warp
  // (the TMA store of one chunk drains while the next is computed) and halve the accumulator a warp walks per tile
  if (a->Ktot <= 1280 && a->N % 160 == 0 && a->mode == SG_EPI_LINEAR) bn_tile = 160;
  const long long m_tiles = static_cast<long long>((a->out_w + a->bw - 1) / a->bw) * ((a->out_h + a->bh - 1) / a->bh) *
                            ((a->out_n + a->bn - 1) / a->bn);
  while (bn_tile > 64 && bn_tile % 128 == 0 && m_tiles * ((a->N + bn_tile - 1) / bn_tile) * 2 <= num_sms()) bn_tile /= 2;
  return bn_tile;
}

// Epilogue warpgroups: the small-K GEMMs (K <= 1280: every Linear of the transformer blocks) are bound by the epilogue's
// dependent chains, not by the main loop, and run with more than two epilogue warpgroups: three for the 160-wide LINEAR
// tiles (5 chunks: 2 + 2 + 1), four for the 256-wide GEGLU tiles (8 chunks: 2 each).  Measured against two warpgroups
// (profiles/r02_gemm_epi_groups.log): QKV projection 90 -> 75 us, GEGLU 188 -> 162 us, out-projection + residual 45 -> 40 us.
static int pick_epi_groups(const sg_gemm_args* a, int bn_tile) {
  if (a->Ktot > 1280) return 2;
  if (bn_tile == 160 && a->mode == SG_EPI_LINEAR) return kWideLinear;
  if (bn_tile == 256 && a->mode == SG_EPI_GEGLU) return kWideGeglu;
  return 2;
}

}  // namespace sg

using namespace sg;

extern "C" int sg_gemm_row_stat_parts(const sg_gemm_args* a) {
  if (a == nullptr || a->bw < 1 || a->bh < 1 || a->bn < 1 || a->N < 1) return SG_ERR_ARG;
  const int bn_tile = pick_bn_tile(a);
  return pick_epi_groups(a, bn_tile) * static_cast<int>((a->N + bn_tile - 1) / bn_tile);
}

extern "C" int sg_gemm(const sg_gemm_args* a, void* stream_v) {
  if (a == nullptr || a->b == nullptr || a->out == nullptr) return SG_ERR_ARG;
  if (a->num_src < 1 || a->num_src > SG_MAX_SRC || a->num_segs < 1 || a->num_segs > SG_MAX_SEGS)
    return SG_ERR_ARG;
  if (a->bw * a->bh * a->bn != kBM || a->bw < 1 || a->bh < 1 || a->bn < 1) return SG_ERR_SHAPE;  // => all powers of two
  if (a->N % 32 != 0 || a->Ktot % kBK != 0) return SG_ERR_SHAPE;
  if (a->mode != SG_EPI_LINEAR && a->mode != SG_EPI_GEGLU && a->mode != SG_EPI_NCHW_F32) return SG_ERR_ARG;
  if (a->mode == SG_EPI_NCHW_F32 && (a->out_channels < 1 || a->out_channels > 32 || a->N != 32)) return SG_ERR_SHAPE;
  if ((a->row_stats_out || a->col_stats_out) && a->mode != SG_EPI_LINEAR) return SG_ERR_ARG;
  if (a->ln_stats && (!a->ln_csum || a->ln_parts < 1 || a->mode == SG_EPI_NCHW_F32)) return SG_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(a->row_stats_out) & 7) || (reinterpret_cast<uintptr_t>(a->col_stats_out) & 15) ||
      (reinterpret_cast<uintptr_t>(a->ln_stats) & 7))
    return SG_ERR_ALIGN;
  if ((reinterpret_cast<uintptr_t>(a->out) & 15) || (a->mode != SG_EPI_NCHW_F32 && (a->ldo % 8)) ||
      (a->residual && ((reinterpret_cast<uintptr_t>(a->residual) & 15) || (a->ldr % 8))))
    return SG_ERR_ALIGN;

  GemmKParams kp;
  memset(&kp, 0, sizeof(kp));
  int total_kb = 0;
  for (int s = 0; s < a->num_segs; ++s) {
    const sg_seg& sgm = a->segs[s];
    if (sgm.src < 0 || sgm.src >= a->num_src || sgm.nkb < 1 || sgm.c0 % 8 != 0) return SG_ERR_ARG;
    if (sgm.c0 + sgm.nkb * kBK > a->src[sgm.src].dim_c) return SG_ERR_SHAPE;
    kp.segs[s] = sgm;
    total_kb += sgm.nkb;
  }
  if (static_cast<int64_t>(total_kb) * kBK != a->Ktot) return SG_ERR_SHAPE;

  for (int i = 0; i < a->num_src; ++i) {
    const sg_src& s = a->src[i];
    uint64_t dims[4] = {static_cast<uint64_t>(s.dim_c), static_cast<uint64_t>(s.dim_w),
                        static_cast<uint64_t>(s.dim_h), static_cast<uint64_t>(s.dim_n)};
    uint64_t strides[3] = {static_cast<uint64_t>(s.stride_w) * 2, static_cast<uint64_t>(s.stride_h) * 2,
                           static_cast<uint64_t>(s.stride_n) * 2};
    uint32_t box[4] = {static_cast<uint32_t>(kBK), static_cast<uint32_t>(a->bw),
                       static_cast<uint32_t>(a->bh), static_cast<uint32_t>(a->bn)};
    int rc = encode_tmap_bf16(&kp.a_maps[i], s.ptr, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != 0) return rc;
  }

  const int bn_tile = pick_bn_tile(a);
  const int eg = pick_epi_groups(a, bn_tile);
  {
    uint64_t dims[2] = {static_cast<uint64_t>(a->Ktot), static_cast<uint64_t>(a->N)};
    uint64_t strides[1] = {static_cast<uint64_t>(a->Ktot) * 2};
    uint32_t box[2] = {static_cast<uint32_t>(kBK), static_cast<uint32_t>(bn_tile)};
    int rc = encode_tmap_bf16(&kp.b_map, a->b, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != 0) return rc;
  }

  kp.num_segs = a->num_segs;
  kp.total_kb = total_kb;
  kp.out_w = a->out_w;
  kp.out_h = a->out_h;
  kp.out_n = a->out_n;
  kp.bw = a->bw;
  kp.bh = a->bh;
  kp.bn = a->bn;
  kp.bw_log2 = 31 - __builtin_clz(static_cast<unsigned>(a->bw));
  kp.bh_log2 = 31 - __builtin_clz(static_cast<unsigned>(a->bh));
  kp.tiles_x = (a->out_w + a->bw - 1) / a->bw;
  kp.tiles_y = (a->out_h + a->bh - 1) / a->bh;
  kp.tiles_n = (a->out_n + a->bn - 1) / a->bn;
  kp.num_n_tiles = static_cast<int>((a->N + bn_tile - 1) / bn_tile);
  kp.N = static_cast<int>(a->N);
  kp.mode = a->mode;
  bool res_tma = false;
  if (a->mode != SG_EPI_NCHW_F32) {
    // output view for the epilogue's TMA stores: (column, x, y, n), one box = 32 rows (a warp's slab of the tile:
    // sx x sy x sn pixels) x 32 (GEGLU: 16) columns, swizzle matching the slab's row pitch
    const int out_cols = (a->mode == SG_EPI_GEGLU) ? static_cast<int>(a->N / 2) : static_cast<int>(a->N);
    const long long osw = (a->out_stride_w || a->out_stride_h || a->out_stride_n) ? a->out_stride_w : a->ldo;
    const long long osh = (a->out_stride_w || a->out_stride_h || a->out_stride_n) ? a->out_stride_h : static_cast<long long>(a->out_w) * a->ldo;
    const long long osn = (a->out_stride_w || a->out_stride_h || a->out_stride_n) ? a->out_stride_n
                                                                                 : static_cast<long long>(a->out_h) * a->out_w * a->ldo;
    if ((osw % 8) || (osh % 8) || (osn % 8)) return SG_ERR_ALIGN;
    const int sx = a->bw < 32 ? a->bw : 32;
    const int sy = (a->bh < 32 / sx) ? a->bh : 32 / sx;
    const int sn = 32 / (sx * sy);
    uint64_t dims[4] = {static_cast<uint64_t>(out_cols), static_cast<uint64_t>(a->out_w), static_cast<uint64_t>(a->out_h),
                        static_cast<uint64_t>(a->out_n)};
    uint64_t strides[3] = {static_cast<uint64_t>(osw) * 2, static_cast<uint64_t>(osh) * 2, static_cast<uint64_t>(osn) * 2};
    uint32_t box[4] = {a->mode == SG_EPI_GEGLU ? 16u : 32u, static_cast<uint32_t>(sx), static_cast<uint32_t>(sy),
                       static_cast<uint32_t>(sn)};
    int rc = encode_tmap_bf16(&kp.out_map, a->out, 4, dims, strides, box,
                              a->mode == SG_EPI_GEGLU ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc != 0) return rc;
    // small-K GEMMs with a residual (the transformer blocks' out-projections and proj_out): their epilogue is the critical
    // path, so the residual comes in through TMA slabs instead of row-per-thread global reads
    res_tma = (a->mode == SG_EPI_LINEAR) && (a->residual != nullptr) && (a->Ktot <= 1280) && (bn_tile <= 160);
    if (res_tma) {
      uint64_t rdims[4] = {static_cast<uint64_t>(a->N), static_cast<uint64_t>(a->out_w), static_cast<uint64_t>(a->out_h),
                           static_cast<uint64_t>(a->out_n)};
      uint64_t rstrides[3] = {static_cast<uint64_t>(a->ldr) * 2, static_cast<uint64_t>(a->out_w) * a->ldr * 2,
                              static_cast<uint64_t>(a->out_h) * a->out_w * a->ldr * 2};
      rc = encode_tmap_bf16(&kp.res_map, a->residual, 4, rdims, rstrides, box, CU_TENSOR_MAP_SWIZZLE_64B);
      if (rc != 0) return rc;
    }
  }
  kp.out_channels = a->out_channels;
  kp.res_scale = a->res_scale;
  kp.bias = a->bias;
  kp.lane_bias = a->lane_bias;
  kp.ld_lane_bias = a->ld_lane_bias > 0 ? a->ld_lane_bias : a->N;
  kp.residual = static_cast<const bf16*>(a->residual);
  kp.ldr = a->ldr;
  kp.out = static_cast<bf16*>(a->out);
  kp.ldo = a->ldo;
  kp.dbg = reinterpret_cast<long long*>(a->debug_timeline);
  kp.M = static_cast<long long>(a->out_w) * a->out_h * a->out_n;
  kp.row_stats_out = a->row_stats_out;
  kp.ln_stats = a->ln_stats;
  kp.ln_csum = a->ln_csum;
  kp.ln_parts = a->ln_parts;
  kp.ln_inv_k = 1.0f / static_cast<float>(a->Ktot);
  kp.ln_eps = a->ln_eps;
  kp.col_stats_out = a->col_stats_out;

  const int num_tiles = kp.tiles_x * kp.tiles_y * kp.tiles_n * kp.num_n_tiles;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  // CTA pairs (cta_group::2) for the K-heavy GEMMs with enough tile pairs to fill the machine: their main loop is bound
  // by operand ingest / power, and a pair stages each B byte once instead of twice.  SG_GEMM_PAIR=0 disables (A/B runs).
  static const bool pair_enabled = [] {
    const char* e = getenv("SG_GEMM_PAIR");
    return e == nullptr || e[0] != '0';
  }();
  const int num_m_tiles = kp.tiles_x * kp.tiles_y * kp.tiles_n;
  const int num_pair_tiles = ((num_m_tiles + 1) / 2) * kp.num_n_tiles;
  // ... and for wide outputs at any K (N >= 960: the fused QKV projection, the GEGLU up-projection): every M tile re-reads
  // the whole weight matrix from L2, so their main loop is bound by B traffic, which a pair halves
  if (pair_enabled && (a->mode == SG_EPI_LINEAR || a->mode == SG_EPI_GEGLU) && !res_tma && (total_kb >= 32 || a->N >= 960) &&
      num_pair_tiles >= num_sms() / 2 && (bn_tile == 256 || (bn_tile == 160 && a->mode == SG_EPI_LINEAR))) {
    // the pair's B box holds half of the N tile's rows
    uint64_t dims[2] = {static_cast<uint64_t>(a->Ktot), static_cast<uint64_t>(a->N)};
    uint64_t strides[1] = {static_cast<uint64_t>(a->Ktot) * 2};
    uint32_t box[2] = {static_cast<uint32_t>(kBK), static_cast<uint32_t>(bn_tile / 2)};
    int rc = encode_tmap_bf16(&kp.b_map, a->b, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != 0) return rc;
    if (a->mode == SG_EPI_GEGLU) {
      if (eg == kWideGeglu) return launch_gemm_pair<256, SG_EPI_GEGLU, kWideGeglu>(kp, num_pair_tiles, stream);
      return launch_gemm_pair<256, SG_EPI_GEGLU>(kp, num_pair_tiles, stream);
    }
    if (bn_tile == 256) return launch_gemm_pair<256, SG_EPI_LINEAR>(kp, num_pair_tiles, stream);
    if (eg == kWideLinear) return launch_gemm_pair<160, SG_EPI_LINEAR, kWideLinear>(kp, num_pair_tiles, stream);
    return launch_gemm_pair<160, SG_EPI_LINEAR>(kp, num_pair_tiles, stream);
  }
  switch (bn_tile) {
    case 256: return launch_gemm<256>(kp, num_tiles, false, eg, stream);
    case 160: return launch_gemm<160>(kp, num_tiles, res_tma, eg, stream);
    case 128: return launch_gemm<128>(kp, num_tiles, res_tma, eg, stream);
    default: return launch_gemm<64>(kp, num_tiles, res_tma, eg, stream);
  }
}
