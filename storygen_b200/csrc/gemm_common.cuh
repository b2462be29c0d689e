// Shared pieces of the tcgen05 GEMM kernels: parameters and the epilogue.
#pragma once
#include "common.cuh"

namespace sg {


constexpr int kBM = 128;
constexpr int kBK = 64;  // one 128-byte swizzle atom of bf16


struct GemmKParams {
  CUtensorMap a_maps[SG_MAX_SRC];
  CUtensorMap b_map;
  CUtensorMap out_map;  // (col, x, y, n) view of the output, box = (32 | 16 columns) x (one warp's 32-row slab of the tile)
  CUtensorMap res_map;  // same view of the residual (RT = 1 kernels: residual slabs fetched by TMA)
  sg_seg segs[SG_MAX_SEGS];
  int num_segs;
  int total_kb;
  int out_w, out_h, out_n;
  int bw, bh, bn;
  int bw_log2, bh_log2;  // bw, bh are powers of two (bw * bh * bn == 128)
  int tiles_x, tiles_y, tiles_n;  // M tiles = tiles_x * tiles_y * tiles_n
  int num_n_tiles;
  int N;
  int mode;
  int out_channels;  // SG_EPI_NCHW_F32: number of valid output channels (<= 32)
  float res_scale;
  const float* bias;
  const float* lane_bias;
  long long ld_lane_bias;
  const bf16* residual;
  long long ldr;
  bf16* out;
  long long ldo;
  long long* dbg;  // bring-up aid: phase clocks of CTA 0 (NULL in production)
  // fused normalisation hooks (include/storygen_b200.h): all optional
  long long M;            // total output rows (stride of the row-statistics parts)
  float* row_stats_out;   // [EG * num_n_tiles][M][2]  (sg_gemm_row_stat_parts)
  const float* ln_stats;  // [ln_parts][M][2]
  const float* ln_csum;   // [N]
  int ln_parts;
  float ln_inv_k, ln_eps;
  float* col_stats_out;   // [num_m_tiles * 4][N][2]
};

// Per epilogue warp: kSlabs staging slabs of [32 rows x 64 B] (the TMA store's source) + the bias / csum vectors of the
// warp's (at most 4) 32-column chunks of a tile.
//   RT = 0: 2 slabs used alternately (1 for 256-wide tiles, whose deep operand ring needs the shared memory); a
//           residual is read by the threads themselves, row by row (fine when the main loop hides the epilogue);
//   RT = 1: one slab per chunk of the tile; the residual's slabs are fetched INTO them by TMA at the start of the tile,
//           every thread adds its row in place and the same slab is the TMA store's source.  For the small-K GEMMs of
//           the transformer blocks (K = C or 2C), whose epilogue is the critical path: a row-per-thread global read costs
//           one L1 wavefront per 16 bytes, the bulk copy none.
//   CG = 2: CTA pair (cta_group::2): this CTA stages only its half of the B tile.
//   EG = number of epilogue warpgroups (2..4), warpgroup g taking the 32-column chunks g, g + EG, ...  Two suffice when the
//           main loop is the critical path (K-heavy convolutions); the small-K GEMMs of the transformer blocks are bound by
//           the epilogue's dependent chains (tcgen05.ld -> arithmetic -> shared -> TMA store), which more resident warps hide.
template <int BN, int RT = 0, int CG = 1, int EG = 2>
struct GemmCfg {
  static constexpr int kEpiWarps = 4 * EG;
  static constexpr int kThreads = 64 + 128 * EG;  // TMA warp + MMA warp + EG epilogue warpgroups
  static constexpr int kABytes = kBM * kBK * 2;  // 16 KB
  static constexpr int kBBytes = (BN / CG) * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kMine = (BN / 32 + EG - 1) / EG;  // chunks of a tile per epilogue warp
  static constexpr int kSlabs = RT ? kMine : ((BN == 256) ? 1 : 2);
  static constexpr int kVecBytes = 2 * 4 * 128;  // bias[4][32] + csum[4][32] floats
  static constexpr int kWarpScratch = kSlabs * 2048 + kVecBytes;
  static constexpr int kScratchBytes = kEpiWarps * kWarpScratch;
  static constexpr int kBarBytes = 320;   // 2 x kStages + 4 pipeline barriers, kEpiWarps residual barriers, the TMEM slot
  static constexpr int kMaxSmem = 231424;  // 227 KB - 1 KB (reserved)
  static constexpr int kFit = (kMaxSmem - 1024 /*align*/ - kBarBytes - kScratchBytes) / kStageBytes;
  static constexpr int kStages = kFit > 8 ? 8 : kFit;
  static constexpr int kTmemCols = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);
  static constexpr int kSmemBytes = kStages * kStageBytes + kScratchBytes + 1024 /*align*/ + kBarBytes;
  static_assert(kStages >= 3, "too few pipeline stages");
  static_assert(kSmemBytes <= kMaxSmem, "exceeds the 227 KB - 1 KB (reserved) dynamic shared memory limit");
};

// erf-GELU with ONE MUFU: for z = |x| / sqrt(2), erfc(z) = 2^(z * P7(z)) with a degree-7 minimax fit of log2(erfc)
// on [0, 6] (erfc(6) ~ 2e-17), and gelu(x) = relu(x) - 0.5 |x| erfc(z): exact erf semantics (no tanh approximation),
// |error| <= 7e-6 absolute and <= 6e-5 relative even in the negative tail, i.e. far below the bf16 output rounding.
// 13 FMA-pipe instructions + 1 MUFU (Abramowitz-Stegun 7.1.26 needed 2 MUFU + ~16; erff() ~35).
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fminf(fabsf(x) * 0.70710678f, 6.0f);
  float q = fmaf(z, -1.2884705e-05f, 3.6014518e-04f);
  q = fmaf(q, z, -4.3878024e-03f);
  q = fmaf(q, z, 3.1177560e-02f);
  q = fmaf(q, z, -1.4790413e-01f);
  q = fmaf(q, z, -9.2054057e-01f);
  q = fmaf(q, z, -1.6271738f);
  const float e = fast_exp2(q * z);                    // erfc(|x| / sqrt(2))
  return fmaf(-0.70710678f * z, e, fmaxf(x, 0.f));  // 0.5 |x| = z / sqrt(2)
}

// two gelu_erf at once with packed fp32 arithmetic (same polynomial): 17 instead of 34 issue slots per pair
__device__ __forceinline__ float2 gelu_erf2(float2 x) {
  const float2 z = make_float2(fminf(fabsf(x.x) * 0.70710678f, 6.0f), fminf(fabsf(x.y) * 0.70710678f, 6.0f));
  float2 q = ffma2(z, make_float2(-1.2884705e-05f, -1.2884705e-05f), make_float2(3.6014518e-04f, 3.6014518e-04f));
  q = ffma2(q, z, make_float2(-4.3878024e-03f, -4.3878024e-03f));
  q = ffma2(q, z, make_float2(3.1177560e-02f, 3.1177560e-02f));
  q = ffma2(q, z, make_float2(-1.4790413e-01f, -1.4790413e-01f));
  q = ffma2(q, z, make_float2(-9.2054057e-01f, -9.2054057e-01f));
  q = ffma2(q, z, make_float2(-1.6271738f, -1.6271738f));
  const float2 t = fmul2(q, z);
  const float2 e = make_float2(fast_exp2(t.x), fast_exp2(t.y));  // erfc(|x| / sqrt(2))
  return ffma2(fmul2(z, make_float2(-0.70710678f, -0.70710678f)), e, make_float2(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f)));
}

// State an epilogue warp carries from tile to tile: which staging buffer comes next (the TMA stores of earlier tiles may
// still be reading the other one).
struct EpiState {
  uint32_t store_seq = 0;
  uint32_t res_phase = 0;  // RT = 1: parity of this warp's residual barrier
};

// One output tile of the epilogue, executed by one epilogue warp (quarter = TMEM lane quarter it may access = rows
// [32 quarter, 32 quarter + 32) of the tile, grp = epilogue warpgroup taking the 32-column chunks c = grp, grp + EG, ...).
// t_row = TMEM address of this warp's lanes at the accumulator's first column; tfull / tfull_phase = "accumulator
// complete" barrier.
//
// Everything is row-per-thread, exactly as tcgen05.ld 32x32b delivers the accumulator: thread = output row, 32 columns
// per chunk in registers.  LayerNorm-of-A correction, bias, per-image bias, GEGLU, residual and the per-row statistics
// are register arithmetic; the bf16 chunk is then written to a swizzled shared-memory slab [32 rows x 64 B] and leaves
// through ONE TMA bulk store per chunk (coalescing, strided / clipped outputs and the ragged edges are the TMA unit's
// business; the next chunk is processed while the store drains).  Per-channel statistics (GroupNorm of the consumer)
// are column sums of that same slab.  Before the accumulator is awaited the warp has already fetched its residual row
// segments, the bias / csum vectors of all its chunks and its row's LayerNorm statistics, so their latency overlaps
// the tile's main loop; TMEM loads are double-buffered.
template <int BN, int MODE, int RT, int CG = 1, int EG = 2>
__device__ __forceinline__ void epilogue_tile(const GemmKParams& p, const int mt, const int nt, const uint32_t t_row,
                                              uint8_t* scratch, const int quarter, const int grp, const int lane,
                                              uint64_t* tfull, const uint32_t tfull_phase, EpiState& st, uint64_t* res_bar,
                                              const int tile_ord = 0, const int warp_id = 0) {
  using Cfg = GemmCfg<BN, RT, CG, EG>;
  static_assert(RT == 0 || MODE == SG_EPI_LINEAR, "TMA residual slabs: LINEAR epilogue only");
  const bool dbg_on = (p.dbg != nullptr) && blockIdx.x == 0 && lane == 0 && tile_ord < 8;
#define SG_GT(slot) do { if (dbg_on) p.dbg[(warp_id * 8 + tile_ord) * 8 + (slot)] = clock64(); } while (0)
  SG_GT(0);
  const int r = quarter * 32 + lane;
  constexpr int kChunks = BN / 32;
  constexpr int kMine = Cfg::kMine;  // chunks per warp (<= 4)
  constexpr int kVB = (EG == 2) ? 2 : 1;  // TMEM-load / residual register sets (3-4 warpgroups: 112-128 registers per thread)
  const int tx = mt % p.tiles_x;
  const int ty = (mt / p.tiles_x) % p.tiles_y;
  const int tn = mt / (p.tiles_x * p.tiles_y);
  const int x = tx * p.bw + (r & (p.bw - 1));
  const int y = ty * p.bh + ((r >> p.bw_log2) & (p.bh - 1));
  const int n = tn * p.bn + (r >> (p.bw_log2 + p.bh_log2));
  const bool row_ok = (x < p.out_w) && (y < p.out_h) && (n < p.out_n);
  const long long m = (static_cast<long long>(n) * p.out_h + y) * p.out_w + x;  // output row index
  const uint32_t scratch_u32 = smem_u32(scratch);
  const uint32_t bias_u32 = scratch_u32 + Cfg::kSlabs * 2048;
  const uint32_t csum_u32 = bias_u32 + 4 * 128;
  const float* lane_bias_row =
      (p.lane_bias != nullptr) ? p.lane_bias + static_cast<long long>(n) * p.ld_lane_bias : nullptr;
  const bool use_ln = (p.ln_stats != nullptr);  // kernel-uniform
  const bool use_res = (MODE == SG_EPI_LINEAR) && (p.residual != nullptr);
  const bool row_out = (MODE == SG_EPI_LINEAR) && (p.row_stats_out != nullptr);
  const bool col_out = (MODE == SG_EPI_LINEAR) && (p.col_stats_out != nullptr);

  // slab of this warp inside the tile box: rows [32 quarter, +32) = sx x sy x sn pixels starting at (x0, y0, n0)
  const int r0 = quarter * 32;
  const int slab_x = tx * p.bw + (r0 & (p.bw - 1));
  const int slab_y = ty * p.bh + ((r0 >> p.bw_log2) & (p.bh - 1));
  const int slab_n = tn * p.bn + (r0 >> (p.bw_log2 + p.bh_log2));

  // ---- before the accumulator is awaited: everything that comes from global memory -------------------------------
  // RT = 0: the row's residual segment of the NEXT chunk is fetched while the current one is processed (the first
  // before the accumulator is awaited); two register sets instead of one per chunk
  [[maybe_unused]] uint4 res[kVB][4];
  [[maybe_unused]] auto fetch_res = [&](const int ci, uint4 (&dst)[4]) {
    const int col0 = nt * BN + (grp + EG * ci) * 32;
    if (use_res && row_ok && grp + EG * ci < kChunks && col0 < p.N) {
#pragma unroll
      for (int i = 0; i < 4; ++i) dst[i] = __ldg(reinterpret_cast<const uint4*>(p.residual + m * p.ldr + col0) + i);
    }
  };
  if constexpr (RT == 1) {
    __syncwarp();  // every lane is done with the previous tile's slabs (column statistics read them)
    // residual slabs by TMA into the slabs the output will leave from; slab ci is free once the store that read it in the
    // previous tile is done (stores retire in order: wait for all but the kMine - 1 - ci youngest)
    if (use_res && lane == 0) {
      int nvalid = 0;
#pragma unroll
      for (int ci = 0; ci < kMine; ++ci)
        if (grp + EG * ci < kChunks && nt * BN + (grp + EG * ci) * 32 < p.N) ++nvalid;
      mbar_arrive_expect_tx(res_bar, static_cast<uint32_t>(nvalid) * 2048u);
#pragma unroll
      for (int ci = 0; ci < kMine; ++ci) {
        const int col0 = nt * BN + (grp + EG * ci) * 32;
        if (grp + EG * ci < kChunks && col0 < p.N) {
          bulk_wait_group_read_n(kMine - 1 - ci);
          tma_load_4d(reinterpret_cast<void*>(scratch + ci * 2048), &p.res_map, res_bar, col0, slab_x, slab_y, slab_n);
        }
      }
    }
  }
  if constexpr (MODE == SG_EPI_LINEAR && RT == 0) fetch_res(0, res[0]);
  // previous tile's readers of the vectors are done (they finished the tile before arriving on tempty / looping here)
  __syncwarp();
#pragma unroll
  for (int ci = 0; ci < kMine; ++ci) {
    const int col0 = nt * BN + (grp + EG * ci) * 32;
    const bool ok = (grp + EG * ci < kChunks) && (col0 < p.N);
    if (p.bias != nullptr) sts32(bias_u32 + ci * 128 + lane * 4, ok ? __float_as_uint(__ldg(p.bias + col0 + lane)) : 0u);
    if (use_ln) sts32(csum_u32 + ci * 128 + lane * 4, ok ? __float_as_uint(__ldg(p.ln_csum + col0 + lane)) : 0u);
  }
  float ln_mean = 0.f, ln_rstd = 1.f;
  if (use_ln && row_ok) {
    // fused LayerNorm of the A operand: mean / rstd of this thread's row from the producer's row partials
    float s = 0.f, q = 0.f;
    for (int pp = 0; pp < p.ln_parts; ++pp) {
      const float2 e = __ldg(reinterpret_cast<const float2*>(p.ln_stats) + static_cast<long long>(pp) * p.M + m);
      s += e.x;
      q += e.y;
    }
    ln_mean = s * p.ln_inv_k;
    ln_rstd = rsqrtf(fmaxf(q * p.ln_inv_k - ln_mean * ln_mean, 0.f) + p.ln_eps);
  }
  __syncwarp();
  float2 rs2_acc = make_float2(0.f, 0.f), rq2_acc = make_float2(0.f, 0.f);  // per-row (sum, sumsq) of this warp's columns
  bool res_waited = false;

  SG_GT(1);
  mbar_wait(tfull, tfull_phase);
  tc_fence_after();
  SG_GT(2);
  uint32_t v[kVB][32];
  if (kVB == 2 && grp < kChunks) tmem_ld32(t_row + grp * 32, v[0]);
#pragma unroll
  for (int ci = 0; ci < kMine; ++ci) {
    const int c = grp + EG * ci;
    if (c < kChunks) {  // warp-uniform
      if constexpr (kVB == 1) {
        tmem_ld32(t_row + c * 32, v[0]);
        if constexpr (MODE == SG_EPI_LINEAR && RT == 0) {
          if (ci > 0) fetch_res(ci, res[0]);
        }
      }
      tmem_ld_wait();
      if constexpr (kVB == 2) {
        if (ci + 1 < kMine && c + EG < kChunks) {  // overlap the math below
          tmem_ld32(t_row + (c + EG) * 32, v[(ci + 1) & 1]);
          if constexpr (MODE == SG_EPI_LINEAR && RT == 0) fetch_res(ci + 1, res[(ci + 1) & 1]);
        }
      }
      const int col0 = nt * BN + c * 32;
      if (col0 < p.N) {  // warp-uniform
        // the chunk as 16 column pairs: every step below is packed fp32 (f32x2) or bf16x2 arithmetic, one issue slot per
        // two elements (the small-K GEMMs are bound by the epilogue's instruction count)
        float2 f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = make_float2(__uint_as_float(v[ci % kVB][2 * j]), __uint_as_float(v[ci % kVB][2 * j + 1]));
        if (use_ln) {
          // LN(x) W^T + b' = rstd (x W'^T - mean csum) + b': acc holds x W'^T
          const float2 nm = make_float2(-ln_mean, -ln_mean), rs2 = make_float2(ln_rstd, ln_rstd);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 cc = lds128(csum_u32 + ci * 128 + j * 16);  // same address in every lane: broadcast
            f[2 * j] = ffma2(make_float2(__uint_as_float(cc.x), __uint_as_float(cc.y)), nm, f[2 * j]);
            f[2 * j + 1] = ffma2(make_float2(__uint_as_float(cc.z), __uint_as_float(cc.w)), nm, f[2 * j + 1]);
          }
          if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint4 bb = lds128(bias_u32 + ci * 128 + j * 16);
              f[2 * j] = ffma2(f[2 * j], rs2, make_float2(__uint_as_float(bb.x), __uint_as_float(bb.y)));
              f[2 * j + 1] = ffma2(f[2 * j + 1], rs2, make_float2(__uint_as_float(bb.z), __uint_as_float(bb.w)));
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = fmul2(f[j], rs2);
          }
        } else if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 bb = lds128(bias_u32 + ci * 128 + j * 16);
            f[2 * j] = fadd2(f[2 * j], make_float2(__uint_as_float(bb.x), __uint_as_float(bb.y)));
            f[2 * j + 1] = fadd2(f[2 * j + 1], make_float2(__uint_as_float(bb.z), __uint_as_float(bb.w)));
          }
        }
        if (lane_bias_row != nullptr && row_ok) {
          const float4* b4 = reinterpret_cast<const float4*>(lane_bias_row + col0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bb = __ldg(b4 + j);
            f[2 * j] = fadd2(f[2 * j], make_float2(bb.x, bb.y));
            f[2 * j + 1] = fadd2(f[2 * j + 1], make_float2(bb.z, bb.w));
          }
        }
        if constexpr (MODE == SG_EPI_NCHW_F32) {
          // final conv: fp32 NCHW store of the first out_channels columns (coalesced along x)
          if (c == 0 && row_ok) {
            float* of = reinterpret_cast<float*>(p.out);
            const long long plane = static_cast<long long>(p.out_h) * p.out_w;
            const long long base =
                static_cast<long long>(n) * p.out_channels * plane + static_cast<long long>(y) * p.out_w + x;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (2 * j < p.out_channels) of[base + (2 * j) * plane] = f[j].x;
              if (2 * j + 1 < p.out_channels) of[base + (2 * j + 1) * plane] = f[j].y;
            }
          }
        } else {
          // ---- bf16 chunk of this row: 16 (LINEAR) / 8 (GEGLU) packed words
          constexpr int kWords = (MODE == SG_EPI_GEGLU) ? 8 : 16;
          uint32_t w[kWords];
          if constexpr (MODE == SG_EPI_GEGLU) {
            // columns [0,16) = value, [16,32) = gate (weights pre-interleaved on the host)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float2 o = fmul2(f[j], gelu_erf2(f[8 + j]));
              w[j] = pack_bf16x2(o.x, o.y);
            }
          } else {
            // (acc + bias) is rounded to bf16, then the residual is added with ONE rounding (bf16x2 fma), like the reference's
            // separate Linear and add ops under autocast
#pragma unroll
            for (int j = 0; j < 16; ++j) w[j] = pack_bf16x2(f[j].x, f[j].y);
            if (use_res) {
              const uint32_t sc = pack_bf16x2(p.res_scale, p.res_scale);
              if constexpr (RT == 1) {
                // this row of the residual slab (TMA wrote it with the store's swizzle)
                if (!res_waited) {
                  mbar_wait(res_bar, st.res_phase);
                  res_waited = true;
                }
                const uint32_t rrow = scratch_u32 + ci * 2048 + lane * 64;
                const int sw = (lane >> 1) & 3;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const uint4 q = lds128(rrow + ((j ^ sw) << 4));
                  w[4 * j] = hfma2_bf16(sc, q.x, w[4 * j]);
                  w[4 * j + 1] = hfma2_bf16(sc, q.y, w[4 * j + 1]);
                  w[4 * j + 2] = hfma2_bf16(sc, q.z, w[4 * j + 2]);
                  w[4 * j + 3] = hfma2_bf16(sc, q.w, w[4 * j + 3]);
                }
              } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const uint4 q = res[ci % kVB][j];
                  w[4 * j] = hfma2_bf16(sc, q.x, w[4 * j]);
                  w[4 * j + 1] = hfma2_bf16(sc, q.y, w[4 * j + 1]);
                  w[4 * j + 2] = hfma2_bf16(sc, q.z, w[4 * j + 2]);
                  w[4 * j + 3] = hfma2_bf16(sc, q.w, w[4 * j + 3]);
                }
              }
            }
            if (row_out) {
              // per-row (sum, sumsq) of the values as stored
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float2 a = bf16x2_to_float2(w[j]);
                rs2_acc = fadd2(rs2_acc, a);
                rq2_acc = ffma2(a, a, rq2_acc);
              }
            }
            if (!row_ok) {  // rows outside the image: TMA clips them; zeros keep them out of the column statistics
#pragma unroll
              for (int j = 0; j < 16; ++j) w[j] = 0u;
            }
          }
          // ---- staging slab: [32 rows x 64 B] swizzle-64B (LINEAR) / [32 rows x 32 B] swizzle-32B (GEGLU)
          // GEGLU slabs are 1 KB: two of them alternate inside one 2 KB region even when the configuration has a single one
          constexpr int kAlt = (MODE == SG_EPI_GEGLU) ? 2 * Cfg::kSlabs : Cfg::kSlabs;
          constexpr uint32_t kSlabPitch = (MODE == SG_EPI_GEGLU) ? 1024u : 2048u;
          const uint32_t buf = scratch_u32 + (RT == 1 ? ci * 2048u : (st.store_seq % kAlt) * kSlabPitch);
          if constexpr (RT == 0) {
            // the TMA store that last read this slab must be done with it
            if (lane == 0) bulk_wait_group_read<kAlt - 1>();
            __syncwarp();
          } else {
            // without a residual nothing has waited for the previous tile's store of this slab yet
            if (!use_res) {
              if (lane == 0) bulk_wait_group_read<kMine - 1>();
              __syncwarp();
            }
          }
          if constexpr (MODE == SG_EPI_GEGLU) {
            const uint32_t row = buf + lane * 32;
            const int sw = (lane >> 2) & 1;
            sts128(row + ((0 ^ sw) << 4), make_uint4(w[0], w[1], w[2], w[3]));
            sts128(row + ((1 ^ sw) << 4), make_uint4(w[4], w[5], w[6], w[7]));
          } else {
            const uint32_t row = buf + lane * 64;
            const int sw = (lane >> 1) & 3;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              sts128(row + ((j ^ sw) << 4), make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]));
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            const int ocol = (MODE == SG_EPI_GEGLU) ? (col0 >> 1) : col0;
            tma_store_4d(&p.out_map, buf, ocol, slab_x, slab_y, slab_n);
            bulk_commit_group();
          }
          ++st.store_seq;
          if constexpr (MODE == SG_EPI_LINEAR) {
            if (col_out) {
              // per-channel (sum, sumsq) over this warp's 32 rows, read back from the slab (the rounded values the
              // consumer's GroupNorm will see): lane = (row parity, column pair); rows 2k / 2k+1 are 64 B apart, so the
              // two half-warps hit disjoint banks
              const int cp = lane & 15, hf = lane >> 4;
              float2 cs = make_float2(0.f, 0.f), cq = make_float2(0.f, 0.f);
#pragma unroll
              for (int k = 0; k < 16; ++k) {
                const int rr = 2 * k + hf;
                const float2 e = bf16x2_to_float2(lds32(buf + rr * 64 + ((((cp >> 2) ^ ((rr >> 1) & 3))) << 4) + (cp & 3) * 4));
                cs = fadd2(cs, e);
                cq = ffma2(e, e, cq);
              }
              const float s0 = cs.x + __shfl_xor_sync(0xffffffffu, cs.x, 16);
              const float q0 = cq.x + __shfl_xor_sync(0xffffffffu, cq.x, 16);
              const float s1 = cs.y + __shfl_xor_sync(0xffffffffu, cs.y, 16);
              const float q1 = cq.y + __shfl_xor_sync(0xffffffffu, cq.y, 16);
              if (lane < 16)
                *reinterpret_cast<float4*>(reinterpret_cast<float2*>(p.col_stats_out) +
                                           (static_cast<long long>(mt) * 4 + quarter) * p.N + col0 + 2 * cp) =
                    make_float4(s0, q0, s1, q1);
            }
          }
        }
      }
    }
  }
  if constexpr (RT == 1) {
    if (use_res) {
      if (!res_waited) mbar_wait(res_bar, st.res_phase);  // (tile without a valid chunk for this warp: nothing was requested)
      st.res_phase ^= 1;
    }
  }
  if constexpr (MODE == SG_EPI_LINEAR) {
    if (row_out && row_ok)
      reinterpret_cast<float2*>(p.row_stats_out)[static_cast<long long>(nt * EG + grp) * p.M + m] = make_float2(rs2_acc.x + rs2_acc.y, rq2_acc.x + rq2_acc.y);
  }
  SG_GT(3);
#undef SG_GT
}

}  // namespace sg
