// HBM-bound kernels of the hot path: GroupNorm (two-phase), LayerNorm, timestep embedding + small
// dense layers, nearest 2x up-sample, conv_in / conv_out (4 channels), add_noise, CFG + DDIM update.
// All activations are channel-last bf16 and are moved as 16-byte vectors (8 channels); statistics
// are fp32 and are reduced with warp shuffles + fixed-order partial sums (deterministic).
#include "common.cuh"

namespace sg {

// x * sigmoid(x) with two MUFU ops (ex2, rcp.approx) and four FMA-pipe instructions: an IEEE division costs ~12 more and a
// slow-path call per element, which made the GroupNorm apply pass instruction-bound (ncu: issue 60 %, XU 56 %, DRAM 23 %).
// rcp.approx is 1 ulp -- far below the bf16 rounding of the result; exp(-x) = inf (x < -88) gives rcp = +0, result -0.
__device__ __forceinline__ float silu_f(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return x * r;
}

__device__ __forceinline__ void unpack8(const uint4& q, float (&f)[8]) {
  float2 a = unpack_bf16x2(q.x), b = unpack_bf16x2(q.y), c = unpack_bf16x2(q.z), d = unpack_bf16x2(q.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 q;
  q.x = pack_bf16x2(f[0], f[1]); q.y = pack_bf16x2(f[2], f[3]);
  q.z = pack_bf16x2(f[4], f[5]); q.w = pack_bf16x2(f[6], f[7]);
  return q;
}

// ------------------------------------------------------------------------------------------------
// GroupNorm
// ------------------------------------------------------------------------------------------------
constexpr int kGnMaxChunks = 128;
// pixel chunks per image (same split for the statistics and the apply pass): about one resident wave of blocks over
// the whole launch (6 blocks x 148 SMs), at least 4 pixels per thread, so neither the per-block prologue/tail nor
// a handful of fat blocks dominates
static inline int gn_chunks(int hw, int n_img, int PS) {
  int by_work = hw / (4 * PS);
  int by_wave = (6 * num_sms() + n_img - 1) / n_img;
  int c = by_work < by_wave ? by_work : by_wave;
  return c < 1 ? 1 : (c > kGnMaxChunks ? kGnMaxChunks : c);
}
// block = V * PS threads: V = C/8 channel vectors, PS pixel lanes
static inline void gn_block_shape(int C, int* V, int* PS) {
  *V = C / 8;
  int ps = 320 / *V;
  *PS = ps < 1 ? 1 : ps;
}

// partial[n][chunk][group][2] = (sum, sumsq) over the chunk's pixels
__global__ void gn_stats_kernel(const bf16* __restrict__ x0, int c0, const bf16* __restrict__ x1, int c1, int hw,
                                int groups, float* __restrict__ partial) {
  extern __shared__ float sm[];  // [PS][C][2]
  const int C = c0 + c1, V = C / 8, PS = blockDim.x / V;
  const int v = threadIdx.x % V, pl = threadIdx.x / V;
  const int chunks = gridDim.x, chunk = blockIdx.x, n = blockIdx.y;
  const int p_begin = static_cast<int>(static_cast<long long>(hw) * chunk / chunks);
  const int p_end = static_cast<int>(static_cast<long long>(hw) * (chunk + 1) / chunks);
  const bool from0 = (v * 8 < c0);
  const bf16* base = from0 ? x0 + static_cast<long long>(n) * hw * c0 + v * 8
                           : x1 + static_cast<long long>(n) * hw * c1 + (v * 8 - c0);
  const int cs = from0 ? c0 : c1;
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
  // 4 independent 16-byte loads in flight per thread (memory-level parallelism), then the arithmetic
  for (int p = p_begin + pl; p < p_end; p += 4 * PS) {
    uint4 raw[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (p + u * PS < p_end) raw[u] = __ldg(reinterpret_cast<const uint4*>(base + static_cast<long long>(p + u * PS) * cs));
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (p + u * PS < p_end) {
        float f[8];
        unpack8(raw[u], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          s[i] += f[i];
          q[i] += f[i] * f[i];
        }
      }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sm[(pl * C + v * 8 + i) * 2 + 0] = s[i];
    sm[(pl * C + v * 8 + i) * 2 + 1] = q[i];
  }
  __syncthreads();
  // one full warp per group: lanes stride over the (pixel lane, channel) partials, then a shuffle reduction
  const int cpg = C / groups, items = PS * cpg;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, full_warps = blockDim.x >> 5;
  if (warp < full_warps) {
    for (int g = warp; g < groups; g += full_warps) {
      float ts = 0.f, tq = 0.f;
      for (int it = lane; it < items; it += 32) {
        const int pp = it / cpg, c = g * cpg + it % cpg;
        const float2 e = *reinterpret_cast<const float2*>(&sm[(pp * C + c) * 2]);
        ts += e.x;
        tq += e.y;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        ts += __shfl_xor_sync(0xffffffffu, ts, o);
        tq += __shfl_xor_sync(0xffffffffu, tq, o);
      }
      if (lane == 0)
        *reinterpret_cast<float2*>(partial + ((static_cast<long long>(n) * chunks + chunk) * groups + g) * 2) =
            make_float2(ts, tq);
    }
  }
}

__global__ void gn_apply_kernel(const bf16* __restrict__ x0, int c0, const bf16* __restrict__ x1, int c1, int hw,
                                int groups, int chunks, float eps, const float* __restrict__ partial,
                                const float* __restrict__ gamma, const float* __restrict__ beta, int silu,
                                bf16* __restrict__ y) {
  __shared__ float s_mean[64], s_rstd[64];
  __shared__ float2 s_red[320];
  const int C = c0 + c1, V = C / 8, PS = blockDim.x / V;
  const int v = threadIdx.x % V, pl = threadIdx.x / V;
  const int n = blockIdx.y;
  const int cpg = C / groups;
  {  // reduce the per-chunk partials: thread (row, g) sums chunks row, row + rows, ...; then rows are folded
    const int rows_all = blockDim.x / groups, rows_cap = 320 / groups;
    const int rows = rows_all < rows_cap ? rows_all : rows_cap, g = threadIdx.x % groups, row = threadIdx.x / groups;
    if (row < rows) {
      float ts = 0.f, tq = 0.f;
      for (int ch = row; ch < chunks; ch += rows) {
        const float2 e = *reinterpret_cast<const float2*>(partial + ((static_cast<long long>(n) * chunks + ch) * groups + g) * 2);
        ts += e.x;
        tq += e.y;
      }
      s_red[row * groups + g] = make_float2(ts, tq);
    }
    __syncthreads();
    if (threadIdx.x < groups) {
      float ts = 0.f, tq = 0.f;
      for (int r = 0; r < rows; ++r) {
        ts += s_red[r * groups + threadIdx.x].x;
        tq += s_red[r * groups + threadIdx.x].y;
      }
      const float cnt = static_cast<float>(hw) * cpg;
      const float mean = ts / cnt;
      const float var = fmaxf(tq / cnt - mean * mean, 0.f);
      s_mean[threadIdx.x] = mean;
      s_rstd[threadIdx.x] = rsqrtf(var + eps);
    }
  }
  __syncthreads();
  const int slabs = gridDim.x, slab = blockIdx.x;
  const int p_begin = static_cast<int>(static_cast<long long>(hw) * slab / slabs);
  const int p_end = static_cast<int>(static_cast<long long>(hw) * (slab + 1) / slabs);
  const bool from0 = (v * 8 < c0);
  const bf16* base = from0 ? x0 + static_cast<long long>(n) * hw * c0 + v * 8
                           : x1 + static_cast<long long>(n) * hw * c1 + (v * 8 - c0);
  const int cs = from0 ? c0 : c1;
  float a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = v * 8 + i;
    const int g = c / cpg;
    const float ga = gamma[c] * s_rstd[g];
    a[i] = ga;
    b[i] = beta[c] - s_mean[g] * ga;
  }
  bf16* ybase = y + static_cast<long long>(n) * hw * C + v * 8;
  for (int p = p_begin + pl; p < p_end; p += 4 * PS) {
    uint4 raw[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (p + u * PS < p_end) raw[u] = __ldg(reinterpret_cast<const uint4*>(base + static_cast<long long>(p + u * PS) * cs));
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (p + u * PS < p_end) {
        float f[8];
        unpack8(raw[u], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float t = fmaf(f[i], a[i], b[i]);
          f[i] = silu ? silu_f(t) : t;
        }
        *reinterpret_cast<uint4*>(ybase + static_cast<long long>(p + u * PS) * C) = pack8(f);
      }
  }
}

// GroupNorm statistics from the PRODUCER's epilogue (sg_gemm col_stats_out): reduce the per-(32-row quarter, channel)
// partials of image n / group g to (mean, rstd).  One block per (group, image); every thread sums a fixed subset in a
// fixed order and the block tree is fixed, so the result is deterministic; fp64 accumulation keeps E[x^2] - mean^2
// from cancelling when |mean| >> std.
struct GnSrc {
  const float2* parts;
  int c, tiles_xy, bn, reps, m_tiles;
};
constexpr int kGnFinThreads = 256;
__global__ void __launch_bounds__(kGnFinThreads) gn_finalize_kernel(const GnSrc s0, const GnSrc s1, int hw, int groups, float eps,
                                                                    float2* __restrict__ mean_rstd) {
  __shared__ double s_s[kGnFinThreads / 32], s_q[kGnFinThreads / 32];
  const int g = blockIdx.x, n = blockIdx.y;
  const int C = s0.c + s1.c, cpg = C / groups;
  const int ch_lo = g * cpg, ch_hi = ch_lo + cpg;
  double s = 0.0, q = 0.0;
#pragma unroll
  for (int si = 0; si < 2; ++si) {
    const GnSrc& S = si == 0 ? s0 : s1;
    const int base = si == 0 ? 0 : s0.c;
    // channels of this group that live in source si
    const int lo = (ch_lo > base ? ch_lo : base) - base;
    const int hi = (ch_hi < base + S.c ? ch_hi : base + S.c) - base;
    if (S.c == 0 || hi <= lo) continue;
    const int nch = hi - lo;
    // quarter-blocks (32 rows) of image n: nq = 4 / bn per tile (bn = 1, 2 or 4 images per tile), q0 = first one of the image
    const int nq_log = S.bn == 1 ? 2 : (S.bn == 2 ? 1 : 0), nq = 1 << nq_log, q0 = (n % S.bn) << nq_log;
    const long long tile0 = static_cast<long long>(n / S.bn) * S.tiles_xy;
    const int rows = S.tiles_xy << nq_log;
    // thread = (channel of the group, block lane): no division inside the loop; consecutive threads read consecutive channels
    const int tpc = kGnFinThreads / nch, my_ch = threadIdx.x % nch, my_b = threadIdx.x / nch;
    if (my_b >= tpc) continue;
    auto fetch = [&](const int rep, const int bi) {
      if (bi >= rows) return make_float2(0.f, 0.f);
      const long long blk = static_cast<long long>(rep) * S.m_tiles * 4 + (tile0 + (bi >> nq_log)) * 4 + q0 + (bi & (nq - 1));
      return __ldg(S.parts + blk * S.c + lo + my_ch);
    };
    // the kernel is a chain of dependent L2 round trips: four independent loads in flight per thread
    for (int rep = 0; rep < S.reps; ++rep)
      for (int bi = my_b; bi < rows; bi += 4 * tpc) {
        const float2 e0 = fetch(rep, bi), e1 = fetch(rep, bi + tpc), e2 = fetch(rep, bi + 2 * tpc), e3 = fetch(rep, bi + 3 * tpc);
        s += (static_cast<double>(e0.x) + static_cast<double>(e1.x)) + (static_cast<double>(e2.x) + static_cast<double>(e3.x));
        q += (static_cast<double>(e0.y) + static_cast<double>(e1.y)) + (static_cast<double>(e2.y) + static_cast<double>(e3.y));
      }
  }
  // fixed-order reduction: butterfly inside the warp, then the 8 warp sums in order
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if ((threadIdx.x & 31) == 0) {
    s_s[threadIdx.x >> 5] = s;
    s_q[threadIdx.x >> 5] = q;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0.0, tq = 0.0;
#pragma unroll
    for (int w = 0; w < kGnFinThreads / 32; ++w) {
      ts += s_s[w];
      tq += s_q[w];
    }
    const double cnt = static_cast<double>(hw) * cpg;
    const double mean = ts / cnt;
    double var = tq / cnt - mean * mean;
    var = var > 0.0 ? var : 0.0;
    mean_rstd[static_cast<long long>(n) * groups + g] =
        make_float2(static_cast<float>(mean), static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps))));
  }
}

// y = [SiLU]((x - mean) * rstd * gamma + beta) with per-(image, group) statistics computed elsewhere
__global__ void gn_apply_ms_kernel(const bf16* __restrict__ x0, int c0, const bf16* __restrict__ x1, int c1, int hw,
                                   int groups, const float2* __restrict__ mean_rstd, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, int silu, bf16* __restrict__ y) {
  const int C = c0 + c1, V = C / 8, PS = blockDim.x / V;
  const int v = threadIdx.x % V, pl = threadIdx.x / V;
  const int n = blockIdx.y;
  const int cpg = C / groups;
  const int slabs = gridDim.x, slab = blockIdx.x;
  const int p_begin = static_cast<int>(static_cast<long long>(hw) * slab / slabs);
  const int p_end = static_cast<int>(static_cast<long long>(hw) * (slab + 1) / slabs);
  const bool from0 = (v * 8 < c0);
  const bf16* base = from0 ? x0 + static_cast<long long>(n) * hw * c0 + v * 8
                           : x1 + static_cast<long long>(n) * hw * c1 + (v * 8 - c0);
  const int cs = from0 ? c0 : c1;
  // first loads in flight before the (dependent) statistics / affine parameters are fetched
  uint4 raw[4];
  const int p0 = p_begin + pl;
#pragma unroll
  for (int u = 0; u < 4; ++u)
    if (p0 + u * PS < p_end) raw[u] = __ldg(reinterpret_cast<const uint4*>(base + static_cast<long long>(p0 + u * PS) * cs));
  float a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = v * 8 + i;
    const float2 mr = __ldg(mean_rstd + static_cast<long long>(n) * groups + c / cpg);
    const float ga = __ldg(gamma + c) * mr.y;
    a[i] = ga;
    b[i] = __ldg(beta + c) - mr.x * ga;
  }
  bf16* ybase = y + static_cast<long long>(n) * hw * C + v * 8;
  for (int p = p0; p < p_end; p += 4 * PS) {
    uint4 nxt[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (p + (4 + u) * PS < p_end)
        nxt[u] = __ldg(reinterpret_cast<const uint4*>(base + static_cast<long long>(p + (4 + u) * PS) * cs));
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (p + u * PS < p_end) {
        float f[8];
        unpack8(raw[u], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float t = fmaf(f[i], a[i], b[i]);
          f[i] = silu ? silu_f(t) : t;
        }
        *reinterpret_cast<uint4*>(ybase + static_cast<long long>(p + u * PS) * C) = pack8(f);
      }
#pragma unroll
    for (int u = 0; u < 4; ++u) raw[u] = nxt[u];
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, row held in registers (C <= 1280), two affine outputs
// ------------------------------------------------------------------------------------------------
template <int MAXV>  // vectors of 8 per lane
__global__ void layernorm_kernel(const bf16* __restrict__ x, long long rows, int C, float eps,
                                 const float* __restrict__ ga, const float* __restrict__ ba, bf16* __restrict__ ya,
                                 const float* __restrict__ gb, const float* __restrict__ bb, bf16* __restrict__ yb) {
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int V = C / 8;
  const bf16* xr = x + row * C;
  float f[MAXV][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = lane + 32 * i;
    if (v < V) {
      uint4 raw = __ldg(reinterpret_cast<const uint4*>(xr + v * 8));
      unpack8(raw, f[i]);
#pragma unroll
      for (int k = 0; k < 8; ++k) sum += f[i][k];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / C;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = lane + 32 * i;
    if (v < V) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float d = f[i][k] - mean;
        sq += d * d;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / C + eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = lane + 32 * i;
    if (v < V) {
      float o[8];
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(ga + v * 8));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(ga + v * 8 + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(ba + v * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(ba + v * 8 + 4));
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = (f[i][k] - mean) * rstd * gg[k] + bv[k];
      *reinterpret_cast<uint4*>(ya + row * C + v * 8) = pack8(o);
      if (yb != nullptr) {
        const float4 h0 = __ldg(reinterpret_cast<const float4*>(gb + v * 8));
        const float4 h1 = __ldg(reinterpret_cast<const float4*>(gb + v * 8 + 4));
        const float4 c0 = __ldg(reinterpret_cast<const float4*>(bb + v * 8));
        const float4 c1 = __ldg(reinterpret_cast<const float4*>(bb + v * 8 + 4));
        const float hg[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
        const float hb[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = (f[i][k] - mean) * rstd * hg[k] + hb[k];
        *reinterpret_cast<uint4*>(yb + row * C + v * 8) = pack8(o);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// timestep embedding
// ------------------------------------------------------------------------------------------------
__global__ void timestep_features_kernel(const float* __restrict__ t, int n, int dim, float* __restrict__ out) {
  const int half = dim / 2;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * half) return;
  const int i = idx % half, r = idx / half;
  // exponent = -ln(10000) * i / half   (downscale_freq_shift = 0)
  const float freq = expf(-9.210340371976184f * static_cast<float>(i) / static_cast<float>(half));
  const float arg = t[r] * freq;
  out[r * dim + i] = cosf(arg);          // flip_sin_to_cos=True -> [cos | sin]
  out[r * dim + half + i] = sinf(arg);
}

// out[r][m] = act_out( sum_k W[m][k] * act_in(x[r][k]) + b[m] ) for a handful of rows (image lanes).
// One block per SM: all n activation rows are staged ONCE into shared memory (fp32, act_in applied), then every warp
// walks output features two at a time: the two weight rows are streamed from HBM/L2 exactly once (bf16, 8-byte
// coalesced loads kept in registers), each x value read from smem feeds both features, and the per-lane partial
// sums of up to kSlMaxRows rows are folded with shuffles.  Element mapping inside a row: float4 index
// q = i * 32 + lane (i < 10) -> conflict-free LDS.128.  Rows beyond kSlMaxRows are processed in further passes.
constexpr int kSlMaxRows = 24;  // rows per pass (accumulators: 2 features x 24 rows = 48 registers)
constexpr int kSlWarps = 16;
__global__ void __launch_bounds__(kSlWarps * 32, 1)
    small_linear_kernel(const float* __restrict__ x, int n, int k, const bf16* __restrict__ w,
                        const float* __restrict__ b, int m, int silu_in, int silu_out, float* __restrict__ out) {
  extern __shared__ float xs[];  // [min(n, kSlMaxRows)][k]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Q = k / 4;  // float4 per row
  for (int r0 = 0; r0 < n; r0 += kSlMaxRows) {
    const int rows = (n - r0) < kSlMaxRows ? (n - r0) : kSlMaxRows;
    __syncthreads();  // previous pass fully consumed
    for (int idx = threadIdx.x; idx < rows * Q; idx += blockDim.x) {
      float4 v = __ldg(reinterpret_cast<const float4*>(x + static_cast<long long>(r0) * k) + idx);
      if (silu_in) { v.x = silu_f(v.x); v.y = silu_f(v.y); v.z = silu_f(v.z); v.w = silu_f(v.w); }
      reinterpret_cast<float4*>(xs)[idx] = v;
    }
    __syncthreads();
    // feature pairs are dealt round-robin to all warps of all blocks
    for (int f0 = 2 * (blockIdx.x * kSlWarps + warp); f0 < m; f0 += 2 * gridDim.x * kSlWarps) {
      const bool two = (f0 + 1 < m);
      float acc0[kSlMaxRows], acc1[kSlMaxRows];
#pragma unroll
      for (int r = 0; r < kSlMaxRows; ++r) acc0[r] = acc1[r] = 0.f;
      uint2 w0[10], w1[10];
#pragma unroll
      for (int i = 0; i < 10; ++i) {
        const int q = i * 32 + lane;
        w0[i] = (q < Q) ? __ldg(reinterpret_cast<const uint2*>(w + static_cast<long long>(f0) * k) + q) : make_uint2(0, 0);
        w1[i] = (two && q < Q) ? __ldg(reinterpret_cast<const uint2*>(w + static_cast<long long>(f0 + 1) * k) + q)
                               : make_uint2(0, 0);
      }
#pragma unroll
      for (int i = 0; i < 10; ++i) {
        const int q = i * 32 + lane;
        if (q < Q) {
          const float2 a0 = unpack_bf16x2(w0[i].x), a1 = unpack_bf16x2(w0[i].y);
          const float2 c0 = unpack_bf16x2(w1[i].x), c1 = unpack_bf16x2(w1[i].y);
#pragma unroll
          for (int r = 0; r < kSlMaxRows; ++r) {
            if (r < rows) {
              const float4 xv = reinterpret_cast<const float4*>(xs + r * k)[q];
              acc0[r] = fmaf(a0.x, xv.x, fmaf(a0.y, xv.y, fmaf(a1.x, xv.z, fmaf(a1.y, xv.w, acc0[r]))));
              acc1[r] = fmaf(c0.x, xv.x, fmaf(c0.y, xv.y, fmaf(c1.x, xv.z, fmaf(c1.y, xv.w, acc1[r]))));
            }
          }
        }
      }
      const float b0 = b ? b[f0] : 0.f, b1 = (b && two) ? b[f0 + 1] : 0.f;
#pragma unroll
      for (int r = 0; r < kSlMaxRows; ++r) {
        if (r < rows) {
          float s0 = acc0[r], s1 = acc1[r];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
          }
          if (lane == 0) {
            float* dst = out + static_cast<long long>(r0 + r) * m + f0;
            const float y0 = s0 + b0;
            dst[0] = silu_out ? silu_f(y0) : y0;
            if (two) {
              const float y1 = s1 + b1;
              dst[1] = silu_out ? silu_f(y1) : y1;
            }
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// up-sample, conv_in, conv_out
// ------------------------------------------------------------------------------------------------
__global__ void upsample2x_kernel(const uint4* __restrict__ x, int n, int h, int w, int v /*C/8*/,
                                  uint4* __restrict__ y) {
  const long long total = static_cast<long long>(n) * (2 * h) * (2 * w) * v;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int vc = static_cast<int>(i % v);
    long long p = i / v;
    const int ox = static_cast<int>(p % (2 * w));
    p /= (2 * w);
    const int oy = static_cast<int>(p % (2 * h));
    const int nn = static_cast<int>(p / (2 * h));
    y[i] = __ldg(x + ((static_cast<long long>(nn) * h + (oy >> 1)) * w + (ox >> 1)) * v + vc);
  }
}

// one thread per pixel: gathers its 3x3 x cin patch (zero padded) into one 128-byte bf16 row
__global__ void im2col3x3_nchw_kernel(const float* __restrict__ x, int n, int cin, int h, int w,
                                      bf16* __restrict__ out) {
  const long long pix = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (pix >= static_cast<long long>(n) * h * w) return;
  const int px = static_cast<int>(pix % w);
  const int py = static_cast<int>((pix / w) % h);
  const int nn = static_cast<int>(pix / (static_cast<long long>(w) * h));
  __align__(16) bf16 row[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) row[i] = __float2bfloat16(0.f);
  for (int ky = 0; ky < 3; ++ky) {
    const int yy = py + ky - 1;
    if (yy < 0 || yy >= h) continue;
    for (int kx = 0; kx < 3; ++kx) {
      const int xx = px + kx - 1;
      if (xx < 0 || xx >= w) continue;
      for (int ci = 0; ci < cin; ++ci)
        row[(ky * 3 + kx) * cin + ci] = __float2bfloat16(x[((static_cast<long long>(nn) * cin + ci) * h + yy) * w + xx]);
    }
  }
  uint4* dst = reinterpret_cast<uint4*>(out + pix * 64);
  const uint4* src = reinterpret_cast<const uint4*>(row);
#pragma unroll
  for (int i = 0; i < 8; ++i) dst[i] = src[i];
}

// ------------------------------------------------------------------------------------------------
// scheduler arithmetic
// ------------------------------------------------------------------------------------------------
__global__ void add_noise_kernel(const float* __restrict__ x, const float* __restrict__ eps,
                                 const float* __restrict__ acp, const int* __restrict__ t, int n, int chw,
                                 float* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(n) * chw) return;
  const int r = static_cast<int>(i / chw);
  const float a = acp[t[r]];
  out[i] = sqrtf(a) * x[i] + sqrtf(1.0f - a) * eps[i];
}

__global__ void cfg_ddim_kernel(const float* __restrict__ eps3, float* __restrict__ lat, long long total,
                                float s_img, float s_txt, const float* __restrict__ alphas) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float alpha_t = alphas[0], alpha_prev = alphas[1];
  const float u = eps3[i], im = eps3[total + i], al = eps3[2 * total + i];
  const float e = u + s_img * (im - u) + s_txt * (al - im);
  const float x = lat[i];
  const float x0 = (x - sqrtf(1.0f - alpha_t) * e) / sqrtf(alpha_t);
  lat[i] = sqrtf(alpha_prev) * x0 + sqrtf(1.0f - alpha_prev) * e;
}

}  // namespace sg

using namespace sg;

#define SG_LAUNCH_END()   \
  count_launch();         \
  return static_cast<int>(cudaGetLastError())

extern "C" int sg_groupnorm_ws_floats(int32_t n_img, int32_t groups) { return n_img * groups * 2 * kGnMaxChunks; }

extern "C" int sg_groupnorm_stats(const void* x0, int32_t c0, const void* x1, int32_t c1, int32_t n_img,
                                  int32_t hw, int32_t groups, float eps, float* stats, void* stream) {
  (void)eps;
  const int C = c0 + c1;
  if (!x0 || !stats || (c1 > 0 && !x1)) return SG_ERR_ARG;
  if (C % 8 || c0 % 8 || C % groups || groups > 64 || C / 8 > 1024) return SG_ERR_SHAPE;
  int V, PS;
  gn_block_shape(C, &V, &PS);
  const int chunks = gn_chunks(hw, n_img, PS);
  const size_t smem = static_cast<size_t>(PS) * C * 2 * sizeof(float);
  gn_stats_kernel<<<dim3(chunks, n_img), V * PS, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(x0), c0, static_cast<const bf16*>(x1), c1, hw, groups, stats);
  SG_LAUNCH_END();
}

extern "C" int sg_groupnorm_apply(const void* x0, int32_t c0, const void* x1, int32_t c1, int32_t n_img,
                                  int32_t hw, int32_t groups, float eps, const float* stats, const float* gamma,
                                  const float* beta, int32_t silu, void* y, void* stream) {
  const int C = c0 + c1;
  if (!x0 || !stats || !gamma || !beta || !y || (c1 > 0 && !x1)) return SG_ERR_ARG;
  if (C % 8 || c0 % 8 || C % groups || groups > 64 || C / 8 > 1024) return SG_ERR_SHAPE;
  int V, PS;
  gn_block_shape(C, &V, &PS);
  const int chunks = gn_chunks(hw, n_img, PS);
  gn_apply_kernel<<<dim3(chunks, n_img), V * PS, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(x0), c0, static_cast<const bf16*>(x1), c1, hw, groups, chunks, eps, stats, gamma,
      beta, silu, static_cast<bf16*>(y));
  SG_LAUNCH_END();
}

extern "C" int sg_groupnorm_finalize(const sg_gn_parts* src0, const sg_gn_parts* src1, int32_t n_img, int32_t hw,
                                     int32_t groups, float eps, float* mean_rstd, void* stream) {
  if (!src0 || !src0->parts || !mean_rstd || (src1 && !src1->parts)) return SG_ERR_ARG;
  GnSrc s[2];
  memset(s, 0, sizeof(s));
  const sg_gn_parts* in[2] = {src0, src1};
  int C = 0;
  for (int i = 0; i < 2; ++i) {
    if (!in[i]) {
      s[i].bn = 1;
      continue;
    }
    const sg_gn_parts& p = *in[i];
    if (p.c < 1 || p.tiles_xy < 1 || p.reps < 1 || p.m_tiles < 1) return SG_ERR_SHAPE;
    if (p.imgs_per_tile != 1 && p.imgs_per_tile != 2 && p.imgs_per_tile != 4) return SG_ERR_SHAPE;
    // the quarters of all reps must cover the image exactly once
    if (static_cast<long long>(p.reps) * p.tiles_xy * 128 / p.imgs_per_tile < hw) return SG_ERR_SHAPE;
    if (reinterpret_cast<uintptr_t>(p.parts) & 7) return SG_ERR_ALIGN;
    s[i].parts = reinterpret_cast<const float2*>(p.parts);
    s[i].c = p.c;
    s[i].tiles_xy = p.tiles_xy;
    s[i].bn = p.imgs_per_tile;
    s[i].reps = p.reps;
    s[i].m_tiles = p.m_tiles;
    C += p.c;
  }
  if (groups < 1 || C % groups || n_img < 1) return SG_ERR_SHAPE;
  gn_finalize_kernel<<<dim3(groups, n_img), kGnFinThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      s[0], s[1], hw, groups, eps, reinterpret_cast<float2*>(mean_rstd));
  SG_LAUNCH_END();
}

extern "C" int sg_groupnorm_apply_ms(const void* x0, int32_t c0, const void* x1, int32_t c1, int32_t n_img,
                                     int32_t hw, int32_t groups, const float* mean_rstd, const float* gamma,
                                     const float* beta, int32_t silu, void* y, void* stream) {
  const int C = c0 + c1;
  if (!x0 || !mean_rstd || !gamma || !beta || !y || (c1 > 0 && !x1)) return SG_ERR_ARG;
  if (C % 8 || c0 % 8 || C % groups || groups > 64 || C / 8 > 1024) return SG_ERR_SHAPE;
  int V, PS;
  gn_block_shape(C, &V, &PS);
  const int chunks = gn_chunks(hw, n_img, PS);
  gn_apply_ms_kernel<<<dim3(chunks, n_img), V * PS, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(x0), c0, static_cast<const bf16*>(x1), c1, hw, groups,
      reinterpret_cast<const float2*>(mean_rstd), gamma, beta, silu, static_cast<bf16*>(y));
  SG_LAUNCH_END();
}

extern "C" int sg_layernorm(const void* x, int64_t rows, int32_t c, float eps, const float* gamma_a,
                            const float* beta_a, void* y_a, const float* gamma_b, const float* beta_b, void* y_b,
                            void* stream) {
  if (!x || !gamma_a || !beta_a || !y_a) return SG_ERR_ARG;
  if (y_b && (!gamma_b || !beta_b)) return SG_ERR_ARG;
  if (c % 8 || c > 1280) return SG_ERR_SHAPE;
  const int warps = 8;
  const unsigned grid = static_cast<unsigned>((rows + warps - 1) / warps);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bf16* xp = static_cast<const bf16*>(x);
  bf16* ya = static_cast<bf16*>(y_a);
  bf16* yb = static_cast<bf16*>(y_b);
  if (c <= 512)
    layernorm_kernel<2><<<grid, warps * 32, 0, st>>>(xp, rows, c, eps, gamma_a, beta_a, ya, gamma_b, beta_b, yb);
  else if (c <= 768)
    layernorm_kernel<3><<<grid, warps * 32, 0, st>>>(xp, rows, c, eps, gamma_a, beta_a, ya, gamma_b, beta_b, yb);
  else
    layernorm_kernel<5><<<grid, warps * 32, 0, st>>>(xp, rows, c, eps, gamma_a, beta_a, ya, gamma_b, beta_b, yb);
  SG_LAUNCH_END();
}

extern "C" int sg_timestep_features(const float* t, int32_t n, int32_t dim, float* out, void* stream) {
  if (!t || !out) return SG_ERR_ARG;
  if (dim % 2) return SG_ERR_SHAPE;
  const int total = n * dim / 2;
  timestep_features_kernel<<<(total + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(t, n, dim, out);
  SG_LAUNCH_END();
}

extern "C" int sg_small_linear(const float* x, int32_t n, int32_t k, const void* w, const float* b, int32_t m,
                               int32_t silu_in, int32_t silu_out, float* out, void* stream) {
  if (!x || !w || !out) return SG_ERR_ARG;
  if (k % 8 || k > 1280) return SG_ERR_SHAPE;
  const int rows = n < kSlMaxRows ? n : kSlMaxRows;
  const size_t smem = static_cast<size_t>(rows) * k * sizeof(float);
  static std::atomic<uint64_t> configured{0};
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(small_linear_kernel),
                                   kSlMaxRows * 1280 * static_cast<int>(sizeof(float)), &configured))
    return rc;
  int blocks = (m + 2 * kSlWarps - 1) / (2 * kSlWarps);
  if (blocks > num_sms()) blocks = num_sms();
  small_linear_kernel<<<blocks, kSlWarps * 32, smem, static_cast<cudaStream_t>(stream)>>>(
      x, n, k, static_cast<const bf16*>(w), b, m, silu_in, silu_out, out);
  SG_LAUNCH_END();
}

extern "C" int sg_upsample2x(const void* x, int32_t n, int32_t h, int32_t w, int32_t c, void* y, void* stream) {
  if (!x || !y) return SG_ERR_ARG;
  if (c % 8) return SG_ERR_SHAPE;
  const long long total = static_cast<long long>(n) * 4 * h * w * (c / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  upsample2x_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), n, h, w, c / 8, static_cast<uint4*>(y));
  SG_LAUNCH_END();
}

extern "C" int sg_im2col3x3_nchw(const float* x_nchw, int32_t n, int32_t cin, int32_t h, int32_t w, void* patches,
                                 void* stream) {
  if (!x_nchw || !patches) return SG_ERR_ARG;
  if (cin < 1 || 9 * cin > 64) return SG_ERR_SHAPE;
  const long long pixels = static_cast<long long>(n) * h * w;
  im2col3x3_nchw_kernel<<<static_cast<unsigned>((pixels + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      x_nchw, n, cin, h, w, static_cast<bf16*>(patches));
  SG_LAUNCH_END();
}

extern "C" int sg_add_noise(const float* x, const float* eps, const float* acp, const int32_t* t, int32_t n,
                            int32_t chw, float* out, void* stream) {
  if (!x || !eps || !acp || !t || !out) return SG_ERR_ARG;
  const long long total = static_cast<long long>(n) * chw;
  add_noise_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, eps, acp, t, n, chw, out);
  SG_LAUNCH_END();
}

extern "C" int sg_cfg_ddim_step(const float* eps3, float* latents, int32_t b, int32_t chw, float s_img,
                                float s_txt, const float* alphas, void* stream) {
  if (!eps3 || !latents || !alphas) return SG_ERR_ARG;
  const long long total = static_cast<long long>(b) * chw;
  cfg_ddim_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      eps3, latents, total, s_img, s_txt, alphas);
  SG_LAUNCH_END();
}
