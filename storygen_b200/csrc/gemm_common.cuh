// Shared pieces of the tcgen05 GEMM kernels (1-CTA gemm.cu, CTA-pair gemm2.cu): parameters and the epilogue.
#pragma once
#include "common.cuh"

namespace sg {


constexpr int kBM = 128;
constexpr int kBK = 64;  // one 128-byte swizzle atom of bf16
constexpr int kGemmThreads = 320;  // TMA warp + MMA warp + 2 epilogue warpgroups
constexpr int kEpiWarps = 8;


struct GemmKParams {
  CUtensorMap a_maps[SG_MAX_SRC];
  CUtensorMap b_map;
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
  int transpose_epi;  // 1: coalesce global traffic through a smem transpose (small-K GEMMs); 0: direct row stores
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
  float* row_stats_out;   // [2 * num_n_tiles][M][2]
  const float* ln_stats;  // [ln_parts][M][2]
  const float* ln_csum;   // [N]
  int ln_parts;
  float ln_inv_k, ln_eps;
  float* col_stats_out;   // [num_m_tiles * 4][N][2]
  long long osw, osh, osn;  // output element strides of (x, y, n)  (dense: ldo, out_w * ldo, out_h * out_w * ldo)
};

template <int BN>
struct GemmCfg {
  static constexpr int kABytes = kBM * kBK * 2;  // 16 KB
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (200 * 1024) / kStageBytes > 8 ? 8 : (200 * 1024) / kStageBytes;
  static constexpr int kTmemCols = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);
  static constexpr int kWarpScratch = 2048 + 256;  // per epilogue warp: [32 rows x 64 B] transpose buffer + 32 bias + 32 csum floats
  static constexpr int kScratchBytes = kEpiWarps * kWarpScratch;
  static constexpr int kSmemBytes = kStages * kStageBytes + kScratchBytes + 1024 /*align*/ + 128 /*barriers*/;
  static_assert(kSmemBytes <= 231424, "exceeds the 227 KB - 1 KB (reserved) dynamic shared memory limit");
};

// erf-GELU with ONE MUFU: for z = |x| / sqrt(2), erfc(z) = 2^(z * P7(z)) with a degree-7 minimax fit of log2(erfc)
// on [0, 6] (erfc(6) ~ 2e-17), and gelu(x) = relu(x) - 0.5 |x| erfc(z): exact erf semantics (no tanh approximation),
// |error| <= 7e-6 absolute and <= 6e-5 relative even in the negative tail, i.e. far below the bf16 output rounding.
// 13 FMA-pipe instructions + 1 MUFU (Abramowitz-Stegun 7.1.26 needed 2 MUFU + ~16; erff() ~35): the GEGLU epilogue
// is MUFU / FMA bound for K <= 640.
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


// Column sums over the 32 lanes of a warp of v[0..31] (lane = row, index = column): after the 5 exchange levels lane L
// holds the total of column L in v[0].  31 shuffles instead of 5 x 32.
__device__ __forceinline__ float warp_column_sums(float (&v)[32], const int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? v[i] : v[i + s];
      const float keep = up ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

// One output tile of the epilogue, executed by one epilogue warp (quarter = TMEM lane quarter it may access,
// grp = epilogue warpgroup 0/1 taking the 32-column chunks c = grp, grp+2, ...).  t_row = TMEM address of this
// warp's lanes at the accumulator's first column; tfull / tfull_phase = "accumulator complete" barrier.
template <int BN, int MODE>
__device__ __forceinline__ void epilogue_tile(const GemmKParams& p, const int mt, const int nt, const uint32_t t_row,
                                              uint8_t* scratch, const int quarter, const int grp, const int lane,
                                              uint64_t* tfull, const uint32_t tfull_phase, const int tile_ord = 0, const int warp_id = 0) {
  const bool dbg_on = (p.dbg != nullptr) && blockIdx.x == 0 && lane == 0 && tile_ord < 8;
#define SG_GT(slot) do { if (dbg_on) p.dbg[(warp_id * 8 + tile_ord) * 8 + (slot)] = clock64(); } while (0)
  SG_GT(0);
    // Phase A (row-per-thread, as delivered by tcgen05.ld 32x32b): + bias + time embedding, [GEGLU], -> bf16.
  // Phase B (after a swizzled transpose through this warp's smem scratch): every global access of the
  // warp covers 8 rows x 64 B (GEGLU: 16 rows x 32 B) instead of 32 rows x 16 B; the residual is added and
  // the result stored there.  (acc + bias) is rounded to bf16 before the residual add, exactly like the
  // reference's separate Linear / add ops under autocast.
  const int r = quarter * 32 + lane;
  constexpr int kChunks = BN / 32;
  constexpr int kMine = (kChunks + 1) / 2;                     // chunks per warpgroup (<= 4)
  constexpr int kLpr = (MODE == SG_EPI_GEGLU) ? 2 : 4;         // lanes per row in phase B
  constexpr int kIter = (MODE == SG_EPI_GEGLU) ? 2 : 4;        // phase-B iterations per chunk
  constexpr int kRpi = 32 / kLpr;                              // rows per phase-B iteration
  constexpr int kRowBytes = (MODE == SG_EPI_GEGLU) ? 32 : 64;  // bf16 bytes per row per chunk
  const int tx = mt % p.tiles_x;
  const int ty = (mt / p.tiles_x) % p.tiles_y;
  const int tn = mt / (p.tiles_x * p.tiles_y);
  auto row_coords = [&](int rr, int& x, int& y, int& n) {
    x = tx * p.bw + (rr & (p.bw - 1));
    y = ty * p.bh + ((rr >> p.bw_log2) & (p.bh - 1));
    n = tn * p.bn + (rr >> (p.bw_log2 + p.bh_log2));
  };
  // shared-window address of 16-byte piece `pc` of row `row` (XOR swizzle keeps both phases conflict free)
  const uint32_t scratch_u32 = smem_u32(scratch);
  auto sc_addr = [&](int buf, int row, int pc) -> uint32_t {
    const int sw = (MODE == SG_EPI_GEGLU) ? ((row >> 2) & 1) : ((row >> 1) & 3);
    return scratch_u32 + buf * 2048 + row * kRowBytes + ((pc ^ sw) << 4);
  };
  // bias (+ per-image time-embedding bias) of a chunk: one coalesced 128-byte load per warp, prefetched one
  // chunk ahead into a register, staged in this warp's scratch and read back as broadcasts
  const uint32_t bias_u32 = scratch_u32 + 2048;
  int x, y, n;
  row_coords(r, x, y, n);
  const bool row_ok = (x < p.out_w) && (y < p.out_h) && (n < p.out_n);
  const float* lane_bias_row =
      (p.lane_bias != nullptr) ? p.lane_bias + static_cast<long long>(n) * p.ld_lane_bias : nullptr;
  // phase-B rows of this lane
  long long mB[kIter], oB[kIter];  // row index (residual / statistics) and output element offset
  bool okB[kIter];
  if constexpr (MODE != SG_EPI_NCHW_F32) {
#pragma unroll
    for (int i = 0; i < kIter; ++i) {
      int xb, yb, nb;
      row_coords(quarter * 32 + i * kRpi + lane / kLpr, xb, yb, nb);
      okB[i] = (xb < p.out_w) && (yb < p.out_h) && (nb < p.out_n);
      mB[i] = (static_cast<long long>(nb) * p.out_h + yb) * p.out_w + xb;
      oB[i] = nb * p.osn + yb * p.osh + xb * p.osw;
    }
  }
  // Residual prefetch (coalesced, phase-B mapping) BEFORE the accumulator is awaited: its DRAM latency
  // overlaps the tile's main loop.
  uint4 res[kMine][4];
  const bool use_res = (MODE == SG_EPI_LINEAR) && (p.residual != nullptr);
  const bool tr = (p.transpose_epi != 0);  // kernel-uniform
  if constexpr (MODE == SG_EPI_LINEAR) {
    if (use_res) {
#pragma unroll
      for (int ci = 0; ci < kMine; ++ci) {
        const int c = grp + 2 * ci;
        const int col0 = nt * BN + c * 32;
        if (c < kChunks && col0 < p.N) {
          if (tr) {  // phase-B mapping: 8 rows x 64 B per warp instruction
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (okB[i])
                res[ci][i] = __ldg(reinterpret_cast<const uint4*>(p.residual + mB[i] * p.ldr + col0 + (lane & 3) * 8));
          } else if (row_ok) {  // own row, 4 x 16 B
            const long long m = (static_cast<long long>(n) * p.out_h + y) * p.out_w + x;
#pragma unroll
            for (int i = 0; i < 4; ++i)
              res[ci][i] = __ldg(reinterpret_cast<const uint4*>(p.residual + m * p.ldr + col0) + i);
          }
        }
      }
    }
  }

  float bias_pre = 0.f;
  if (p.bias != nullptr && grp < kChunks && nt * BN + grp * 32 < p.N) bias_pre = __ldg(p.bias + nt * BN + grp * 32 + lane);
  // fused LayerNorm of the A operand: mean / rstd of this thread's row from the producer's row partials
  const bool use_ln = (p.ln_stats != nullptr);  // kernel-uniform
  const uint32_t csum_u32 = scratch_u32 + 2048 + 128;
  float ln_mean = 0.f, ln_rstd = 1.f, csum_pre = 0.f;
  if (use_ln) {
    if (row_ok) {
      const long long m = (static_cast<long long>(n) * p.out_h + y) * p.out_w + x;
      float s = 0.f, q = 0.f;
      for (int pp = 0; pp < p.ln_parts; ++pp) {
        const float2 e = __ldg(reinterpret_cast<const float2*>(p.ln_stats) + static_cast<long long>(pp) * p.M + m);
        s += e.x;
        q += e.y;
      }
      ln_mean = s * p.ln_inv_k;
      ln_rstd = rsqrtf(fmaxf(q * p.ln_inv_k - ln_mean * ln_mean, 0.f) + p.ln_eps);
    }
    if (grp < kChunks && nt * BN + grp * 32 < p.N) csum_pre = __ldg(p.ln_csum + nt * BN + grp * 32 + lane);
  }
  // statistics of the OUTPUT for the consumers' LayerNorm (per row) / GroupNorm (per channel): LINEAR mode only
  const bool row_out = (MODE == SG_EPI_LINEAR) && (p.row_stats_out != nullptr);
  const bool col_out = (MODE == SG_EPI_LINEAR) && (p.col_stats_out != nullptr);
  float rs_acc[4] = {0.f, 0.f, 0.f, 0.f}, rq_acc[4] = {0.f, 0.f, 0.f, 0.f};  // direct path uses [0]
  SG_GT(1);
  mbar_wait(tfull, tfull_phase);
  tc_fence_after();
  SG_GT(2);
  // TMEM loads are double-buffered where registers allow (the LINEAR mode holds the residual prefetch;
  // there the second warp of the sub-partition hides the TMEM latency instead)
  constexpr bool kDouble = (MODE != SG_EPI_LINEAR);
  uint32_t v[kDouble ? 2 : 1][32];
  if (kDouble && grp < kChunks) tmem_ld32(t_row + grp * 32, v[0]);
#pragma unroll
  for (int ci = 0; ci < kMine; ++ci) {
    const int c = grp + 2 * ci;
    if (c < kChunks) {  // warp-uniform
      if constexpr (!kDouble) tmem_ld32(t_row + c * 32, v[0]);
      tmem_ld_wait();
      if (kDouble && ci + 1 < kMine && c + 2 < kChunks)
        tmem_ld32(t_row + (c + 2) * 32, v[(ci + 1) & 1]);  // overlaps the math below
      const int col0 = nt * BN + c * 32;
      if (col0 < p.N) {  // warp-uniform
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[kDouble ? (ci & 1) : 0][j]);
        if (use_ln) {
          // LN(x) W^T = rstd (x W'^T - mean csum): acc holds x W'^T
          sts32(csum_u32 + lane * 4, __float_as_uint(csum_pre));
          if (ci + 1 < kMine && c + 2 < kChunks && col0 + 64 < p.N) csum_pre = __ldg(p.ln_csum + col0 + 64 + lane);
          __syncwarp();
          const float nm = -ln_mean;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 cc = lds128(csum_u32 + j * 16);
            f[4 * j + 0] = ln_rstd * fmaf(nm, __uint_as_float(cc.x), f[4 * j + 0]);
            f[4 * j + 1] = ln_rstd * fmaf(nm, __uint_as_float(cc.y), f[4 * j + 1]);
            f[4 * j + 2] = ln_rstd * fmaf(nm, __uint_as_float(cc.z), f[4 * j + 2]);
            f[4 * j + 3] = ln_rstd * fmaf(nm, __uint_as_float(cc.w), f[4 * j + 3]);
          }
          __syncwarp();  // every lane has read the vector before the next chunk overwrites it
        }
        if (p.bias != nullptr) {
          sts32(bias_u32 + lane * 4, __float_as_uint(bias_pre));
          if (ci + 1 < kMine && c + 2 < kChunks && col0 + 64 < p.N) bias_pre = __ldg(p.bias + col0 + 64 + lane);
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 bb = lds128(bias_u32 + j * 16);  // same address in every lane: broadcast
            f[4 * j + 0] += __uint_as_float(bb.x);
            f[4 * j + 1] += __uint_as_float(bb.y);
            f[4 * j + 2] += __uint_as_float(bb.z);
            f[4 * j + 3] += __uint_as_float(bb.w);
          }
          __syncwarp();  // every lane has read the vector before the next chunk overwrites it
        }
        if (lane_bias_row != nullptr && row_ok) {
          const float4* b4 = reinterpret_cast<const float4*>(lane_bias_row + col0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 bb = __ldg(b4 + j);
            f[4 * j + 0] += bb.x;
            f[4 * j + 1] += bb.y;
            f[4 * j + 2] += bb.z;
            f[4 * j + 3] += bb.w;
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
            for (int j = 0; j < 32; ++j)
              if (j < p.out_channels) of[base + j * plane] = f[j];
          }
        } else if (MODE == SG_EPI_LINEAR && !tr) {
          // direct path (large K: the epilogue is hidden behind the main loop and extra smem traffic would
          // steal operand bandwidth from the tensor core): own row, 4 x 16-byte stores
          // (acc + bias) is rounded to bf16 before the residual add (as the reference's separate ops do); the
          // statistics are taken from the fp32 value just before the final rounding
          float fin[32];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 q;
            q.x = pack_bf16x2(f[8 * j + 0], f[8 * j + 1]);
            q.y = pack_bf16x2(f[8 * j + 2], f[8 * j + 3]);
            q.z = pack_bf16x2(f[8 * j + 4], f[8 * j + 5]);
            q.w = pack_bf16x2(f[8 * j + 6], f[8 * j + 7]);
            float2 a0 = unpack_bf16x2(q.x), a1 = unpack_bf16x2(q.y), a2 = unpack_bf16x2(q.z), a3 = unpack_bf16x2(q.w);
            if (use_res) {
              const uint4 rq = res[ci][j];
              float2 r0 = unpack_bf16x2(rq.x), r1 = unpack_bf16x2(rq.y), r2 = unpack_bf16x2(rq.z), r3 = unpack_bf16x2(rq.w);
              a0.x = fmaf(p.res_scale, r0.x, a0.x); a0.y = fmaf(p.res_scale, r0.y, a0.y);
              a1.x = fmaf(p.res_scale, r1.x, a1.x); a1.y = fmaf(p.res_scale, r1.y, a1.y);
              a2.x = fmaf(p.res_scale, r2.x, a2.x); a2.y = fmaf(p.res_scale, r2.y, a2.y);
              a3.x = fmaf(p.res_scale, r3.x, a3.x); a3.y = fmaf(p.res_scale, r3.y, a3.y);
              q.x = pack_bf16x2(a0.x, a0.y);
              q.y = pack_bf16x2(a1.x, a1.y);
              q.z = pack_bf16x2(a2.x, a2.y);
              q.w = pack_bf16x2(a3.x, a3.y);
            }
            fin[8 * j + 0] = a0.x; fin[8 * j + 1] = a0.y; fin[8 * j + 2] = a1.x; fin[8 * j + 3] = a1.y;
            fin[8 * j + 4] = a2.x; fin[8 * j + 5] = a2.y; fin[8 * j + 6] = a3.x; fin[8 * j + 7] = a3.y;
            if (row_ok) {
              reinterpret_cast<uint4*>(p.out + (n * p.osn + y * p.osh + x * p.osw) + col0)[j] = q;
            }
          }
          if (row_out) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              rs_acc[0] += fin[j];
              rq_acc[0] = fmaf(fin[j], fin[j], rq_acc[0]);
            }
          }
          if (col_out) {
            // per-channel (sum, sumsq) over this warp's 32 rows (rows outside the image contribute zero)
            float sq[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              fin[j] = row_ok ? fin[j] : 0.f;
              sq[j] = fin[j] * fin[j];
            }
            const float cs = warp_column_sums(fin, lane), cq = warp_column_sums(sq, lane);
            reinterpret_cast<float2*>(p.col_stats_out)[(static_cast<long long>(mt) * 4 + quarter) * p.N + col0 + lane] =
                make_float2(cs, cq);
          }
        } else {
          const int buf = 0;
          // ---- phase A -> scratch
          if constexpr (MODE == SG_EPI_LINEAR) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 q;
              q.x = pack_bf16x2(f[8 * j + 0], f[8 * j + 1]);
              q.y = pack_bf16x2(f[8 * j + 2], f[8 * j + 3]);
              q.z = pack_bf16x2(f[8 * j + 4], f[8 * j + 5]);
              q.w = pack_bf16x2(f[8 * j + 6], f[8 * j + 7]);
              sts128(sc_addr(buf, lane, j), q);
            }
          } else {
            // GEGLU: columns [0,16) = value, [16,32) = gate (weights pre-interleaved on the host)
            float g[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) g[j] = f[j] * gelu_erf(f[16 + j]);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              uint4 q;
              q.x = pack_bf16x2(g[8 * j + 0], g[8 * j + 1]);
              q.y = pack_bf16x2(g[8 * j + 2], g[8 * j + 3]);
              q.z = pack_bf16x2(g[8 * j + 4], g[8 * j + 5]);
              q.w = pack_bf16x2(g[8 * j + 6], g[8 * j + 7]);
              sts128(sc_addr(buf, lane, j), q);
            }
          }
          __syncwarp();
          // ---- phase B: coalesced residual add + store
          const int ocol = (MODE == SG_EPI_GEGLU) ? (col0 >> 1) : col0;
          [[maybe_unused]] float ccs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, ccq[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int i = 0; i < kIter; ++i) {
            uint4 q = lds128(sc_addr(buf, i * kRpi + lane / kLpr, lane % kLpr));
            if constexpr (MODE == SG_EPI_LINEAR) {
             if (use_res || row_out || col_out) {  // kernel-uniform
              float2 a0 = unpack_bf16x2(q.x), a1 = unpack_bf16x2(q.y), a2 = unpack_bf16x2(q.z), a3 = unpack_bf16x2(q.w);
              if (use_res && okB[i]) {
                const uint4 rq = res[ci][i];
                float2 r0 = unpack_bf16x2(rq.x), r1 = unpack_bf16x2(rq.y), r2 = unpack_bf16x2(rq.z), r3 = unpack_bf16x2(rq.w);
                a0.x = fmaf(p.res_scale, r0.x, a0.x); a0.y = fmaf(p.res_scale, r0.y, a0.y);
                a1.x = fmaf(p.res_scale, r1.x, a1.x); a1.y = fmaf(p.res_scale, r1.y, a1.y);
                a2.x = fmaf(p.res_scale, r2.x, a2.x); a2.y = fmaf(p.res_scale, r2.y, a2.y);
                a3.x = fmaf(p.res_scale, r3.x, a3.x); a3.y = fmaf(p.res_scale, r3.y, a3.y);
                q.x = pack_bf16x2(a0.x, a0.y);
                q.y = pack_bf16x2(a1.x, a1.y);
                q.z = pack_bf16x2(a2.x, a2.y);
                q.w = pack_bf16x2(a3.x, a3.y);
              }
              if (row_out || col_out) {
                const float fv[8] = {a0.x, a0.y, a1.x, a1.y, a2.x, a2.y, a3.x, a3.y};
                if (row_out) {
                  // this lane holds 8 of the row's 32 columns; the 4 lanes of a row are neighbours
                  float s = 0.f, qq = 0.f;
#pragma unroll
                  for (int k = 0; k < 8; ++k) {
                    s += fv[k];
                    qq = fmaf(fv[k], fv[k], qq);
                  }
                  s += __shfl_xor_sync(0xffffffffu, s, 1);
                  qq += __shfl_xor_sync(0xffffffffu, qq, 1);
                  s += __shfl_xor_sync(0xffffffffu, s, 2);
                  qq += __shfl_xor_sync(0xffffffffu, qq, 2);
                  rs_acc[i] += s;
                  rq_acc[i] += qq;
                }
                if (col_out && okB[i]) {
#pragma unroll
                  for (int k = 0; k < 8; ++k) {
                    ccs[k] += fv[k];
                    ccq[k] = fmaf(fv[k], fv[k], ccq[k]);
                  }
                }
              }
             }
            }
            if (okB[i]) *reinterpret_cast<uint4*>(p.out + oB[i] + ocol + (lane % kLpr) * 8) = q;
          }
          if constexpr (MODE == SG_EPI_LINEAR) {
            if (col_out) {
              // lanes l, l+4, ..., l+28 hold the same 8 columns for different rows: fold them, lanes 0..3 store
#pragma unroll
              for (int k = 0; k < 8; ++k) {
#pragma unroll
                for (int o = 4; o <= 16; o <<= 1) {
                  ccs[k] += __shfl_xor_sync(0xffffffffu, ccs[k], o);
                  ccq[k] += __shfl_xor_sync(0xffffffffu, ccq[k], o);
                }
              }
              if (lane < 4) {
                float4* dst = reinterpret_cast<float4*>(
                    reinterpret_cast<float2*>(p.col_stats_out) + (static_cast<long long>(mt) * 4 + quarter) * p.N + col0 + lane * 8);
#pragma unroll
                for (int k = 0; k < 4; ++k) dst[k] = make_float4(ccs[2 * k], ccq[2 * k], ccs[2 * k + 1], ccq[2 * k + 1]);
              }
            }
          }
          __syncwarp();  // phase-B reads done before the next chunk's phase-A writes
        }
      }
    }
  }
  if constexpr (MODE == SG_EPI_LINEAR) {
    if (row_out) {
      float2* dst = reinterpret_cast<float2*>(p.row_stats_out) + static_cast<long long>(nt * 2 + grp) * p.M;
      if (tr) {
        if ((lane & 3) == 0) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (okB[i]) dst[mB[i]] = make_float2(rs_acc[i], rq_acc[i]);
        }
      } else if (row_ok) {
        dst[(static_cast<long long>(n) * p.out_h + y) * p.out_w + x] = make_float2(rs_acc[0], rq_acc[0]);
      }
    }
  }
  SG_GT(3);
#undef SG_GT
}

}  // namespace sg
