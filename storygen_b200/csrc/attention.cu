// Fused attention core for sm_100a: o = softmax(scale * Q K^T) V per (image lane, head); S / P / O resident in
// TMEM, K/V streamed by TMA through a ring of 128-key tiles.
//
// QT == 2 (Nq > 128, head dim <= 128): two 128-query tiles per CTA, 12 warps
//   warp 0      : TMA loader   both Q tiles once; K0 V0 K1 V1 ... through the ring (shared by both query tiles)
//   warp 1      : Q K^T issuer S_t = Q_t K_j^T (tcgen05.mma SS, M=128 N=128) as soon as K_j has landed and the
//                              softmax warps have READ S_t(j-1) [head dims <= 64: P has its own TMEM columns]
//   warp 2      : P V issuer   O_t += P_t V_j (tcgen05.mma TS: P read from TMEM, V MN-major, N = the head's channels)
//                              (head dims > 64: warp 1 issues both, in order)
//   warp 3      : V fixer      (head dims <= 64) writes a column of ones into channel D of every V tile, so that
//                              column D of O accumulates the softmax row sum on the tensor core
//                              (warps 0..3 shrink to 64 registers, the softmax warps grow to 216: setmaxnreg)
//   warps 4..11 : softmax      one warpgroup per query tile, one thread per query row (= TMEM lane): S row in
//                              registers (streamed: chunk c of S(j+1) replaces chunk c of S(j), and its row max is
//                              taken under the exps of S(j)), lazy rescale of the O row (only when the max grew by
//                              > 2^8), exp2 (MUFU; a fixed subset of the pairs as a packed cubic on the FMA pipe),
//                              P (bf16) -> TMEM, final O / l -> global
//   All issue loops are warp-uniform with one elected lane issuing (ptx.cuh: *_w wrappers).
// QT == 1 (Nq <= 128 or head dim 160): one query tile, warps 0/1 = TMA / MMA, warps 2..5 softmax, S double-buffered.
//
// Q/K/V are read in place from (lane, token, channel) tensors; a head's [tokens x d] slice is fetched as
// ceil(d/64) swizzle-128B atoms of 64 channels starting at channel h*d.  Channels past d inside the last atom
// belong to the next head (or are TMA zero fill): they are neutralised by zeroing the padded Q chunk in smem
// (QK^T) and by never storing the corresponding O columns (PV).
#include <type_traits>

#include "common.cuh"

namespace sg {

constexpr int kAtomBytes = 128 * 128;  // [128 rows x 64 bf16], swizzle 128B
#ifndef SG_ATTN_PINGPONG
#define SG_ATTN_PINGPONG 1
#endif
constexpr bool kPingPong = SG_ATTN_PINGPONG != 0;  // staggered exp phases of the two softmax warpgroups
#ifndef SG_ATTN_HANDOFF
#define SG_ATTN_HANDOFF 0
#endif
constexpr int kHandoff = SG_ATTN_HANDOFF;
#ifndef SG_ATTN_POLY_EVERY
#define SG_ATTN_POLY_EVERY 8
#endif
constexpr int kPolyPairs = SG_ATTN_POLY_EVERY > 0 ? SG_ATTN_POLY_EVERY / 2 : 1;
// head dims <= 64, two query tiles (the d = 40 level, a third of a denoising step): the software-pipelined softmax loop
#ifndef SG_ATTN2_PINGPONG
#define SG_ATTN2_PINGPONG 0
#endif
constexpr bool kPingPong2 = SG_ATTN2_PINGPONG != 0;
#ifndef SG_ATTN2_POLY_MASK
#define SG_ATTN2_POLY_MASK 0x11
#endif
// bit (i % 8) set: pair i of a 32-column chunk takes both exps on the FMA pipe (0x11 = 1 in 4, 0x49 = 3 in 8, 0x55 = 1 in 2)
constexpr unsigned kPolyMask2 = SG_ATTN2_POLY_MASK;
#ifndef SG_ATTN2_SKEW
#define SG_ATTN2_SKEW 0
#endif
// one-time phase offset between the two softmax warpgroups: group 1 starts its first tile's exps when group 0 has issued
// chunk (kSkew2 - 1) of its first tile (0 = no offset), so that one group's per-tile tail (P publish, TMEM round trips, row
// max reduction: no MUFU work) falls into the other group's exp phase instead of both idling the exp unit together
constexpr int kSkew2 = SG_ATTN2_SKEW;
#ifndef SG_ATTN2_ORDERED
#define SG_ATTN2_ORDERED 1
#endif
// 1: the exp loop of a chunk is written as an explicit 3-stage software pipeline (scale pair i+2 | exp pair i | pack pair i-1)
// of order-preserving asm statements, so MUFU and FMA/ALU instructions alternate finely instead of in ptxas' bursts
constexpr bool kOrdered2 = SG_ATTN2_ORDERED != 0;
#ifndef SG_ATTN_MC_POLY_MASK
#define SG_ATTN_MC_POLY_MASK 0x1
#endif
// attn_mc_kernel: bit b set = pair b of every 4 column pairs takes the packed cubic on the FMA pipe instead of MUFU.EX2
constexpr unsigned kMcPolyMask = SG_ATTN_MC_POLY_MASK;
constexpr int kPolyEvery = SG_ATTN_POLY_EVERY;  // 0 = all exps on MUFU; 4 = 1 in 4 on the FMA pipe; 2 = 1 in 2  // the other warpgroup may start its exps once chunk kHandoff (of 4) is issued

struct AttnKParams {
  CUtensorMap q_map, k_map, v_map;
  bf16* o;
  long long o_stride_n, o_stride_t;
  const int* kv_map;
  float* lse;  // optional [lanes][heads][nq]
  long long* dbg;  // optional phase timestamps of CTA (0,0,0): [warp][tile][8] (bring-up aid, NULL in production)
  int nq, nk, n_heads;
  float scale_log2;  // scale * log2(e)
};

// QT = number of 128-query tiles per CTA.
//   QT == 1: S double-buffered (QK^T of tile j+1 overlaps the softmax of tile j); any head dim <= 192.
//   QT == 2: two query tiles ping-pong on the tensor core, one S buffer each, two softmax warpgroups
//            (2 warps per SM sub-partition keep the MUFU pipe busy), K/V tiles fetched once for both.
//   SHORT   : QT == 1 specialisation for Nk <= 128 (the 77-token text cross-attention): one KV tile, so one S buffer,
//             2 smem slots and 256 TMEM columns suffice and two CTAs share an SM (their load / MMA / softmax
//             latencies overlap; a single-tile CTA is pure latency).
template <int D, int QT, bool SHORT = false>
struct AttnCfg {
  static constexpr int NA = (D + 63) / 64;  // 64-channel atoms per tile
  static constexpr int KQ = (D + 15) / 16;  // k-steps of Q K^T
  static constexpr int kTileBytes = NA * kAtomBytes;
  static constexpr int kSlots = SHORT ? 2 : ((QT == 2) ? (NA == 1 ? 8 : 4) : (NA == 1 ? 8 : (NA == 2 ? 5 : 3)));
  static constexpr int kFirstSoftmaxWarp = (QT == 2) ? 4 : 2;  // QT == 2: warpgroup 0 = {TMA, MMA, 2 idle}
  static constexpr int kThreads = 32 * kFirstSoftmaxWarp + 128 * QT;
  static constexpr int kSmemBytes = kTileBytes * (QT + kSlots) + 1024 + 512;
  // kSepP (two query tiles, one 64-channel atom): P gets its own TMEM columns instead of aliasing S, so the
  // tensor core may overwrite S_t with Q K_{j+1}^T as soon as the softmax warps have READ S_t(j) -- the
  // PV_j -> QK_{j+1} -> commit -> wake-up round trip (~1200 clk measured) leaves the softmax critical path.
  //   columns: S_t at t*128 | P_t at 256 + t*64 | O_t at 384 + t*64
  static constexpr bool kSepP = (QT == 2 && NA == 1);
  // kSumTC: the softmax denominator comes from the tensor core.  Channel D of every V tile (the first channel past
  // the head's own, inside the same 64-channel atom) is overwritten with ones in shared memory and the P V MMA is
  // widened to cover it, so O[:, D] = sum_k P[:, k]: 64 packed adds per row and tile leave the FMA pipe, and the
  // lazy O rescale keeps the sum consistent for free.
  static constexpr bool kSumTC = kSepP && (D % 64 != 0);
  static constexpr int kOnesChunk = (D % 64) / 8;  // 16-byte chunk of the V rows that holds channel D
  static constexpr int kColP = 256;
  static constexpr int kColO = SHORT ? 128 : (kSepP ? 384 : 256);
  static constexpr int kTmemCols = SHORT ? 256 : 512;
  static constexpr int kMinCtas = SHORT ? 2 : 1;
  static_assert(QT == 1 || NA <= 2, "two query tiles need 2 x 64*NA O columns next to 256 S columns");
  static_assert(!SHORT || (QT == 1 && NA <= 2), "SHORT: one query tile, S 128 + O <= 128 TMEM columns");
};

#define SG_TS(slot)                                                                                   \
  do {                                                                                                \
    if (dbg_on && lane == 0 && j < 16) p.dbg[(warp * 16 + j) * 8 + (slot)] = clock64();               \
  } while (0)

// DBG: the clock64 timeline of CTA (0,0,0) (sg_attn_args.debug_timeline) is a separate instantiation, so that the
// production kernels carry no stamp code in their inner loops.
template <int D, int QT, bool DBG = false, bool SHORT = false>
__global__ void __launch_bounds__(AttnCfg<D, QT, SHORT>::kThreads, AttnCfg<D, QT, SHORT>::kMinCtas)
    attn_kernel(const __grid_constant__ AttnKParams p) {
  using Cfg = AttnCfg<D, QT, SHORT>;
  constexpr int NA = Cfg::NA;
  constexpr int kSlots = Cfg::kSlots;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* q_smem = smem;
  uint8_t* kv_smem = smem + QT * Cfg::kTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kTileBytes * (QT + kSlots));
  uint64_t* kv_full = bars;              // [kSlots]
  uint64_t* kv_empty = bars + kSlots;    // [kSlots]
  uint64_t* q_full = bars + 2 * kSlots;  // [2] TMA -> softmax warps
  uint64_t* q_ready = q_full + 2;        // [2] softmax warps -> MMA (Q pad zeroed)
  uint64_t* s_full = q_ready + 2;        // [2] MMA -> softmax
  uint64_t* p_full = s_full + 2;         // [2] softmax -> MMA
  uint64_t* o_done = p_full + 2;         // [2] MMA -> softmax (PV retired)
  uint64_t* s_free = o_done + 2;         // [2] softmax -> MMA (S row-blocks are in registers; kSepP only)
  uint64_t* v_ready = s_free + 2;        // [kSlots] V fixer -> PV issuer (kSumTC)
  uint64_t* p_free = v_ready + kSlots;   // [2][4] PV issuer -> softmax: the k-steps that read chunk c of P_t have retired (kSepP)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_free + 8);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // broadcast: provably warp-uniform role index
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (128 * QT);
  const int head = blockIdx.y;
  const int img = blockIdx.z;
  const int T = (p.nk + 127) / 128;
  const bool dbg_on = DBG && (p.dbg != nullptr) && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kSlots; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
      mbar_init(&v_ready[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_ready[i], 4);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_done[i], 1);
      mbar_init(&s_free[i], 4);
    }
    for (int i = 0; i < 8; ++i) mbar_init(&p_free[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);  // warp-uniform for the compiler

  if (warp < Cfg::kFirstSoftmaxWarp) {
  // register re-distribution between warpgroups (QT == 2): warpgroup 0 = {TMA, MMA, 2 idle warps} gives its
  // registers to the 8 softmax warps, which hold a 128-wide S row per thread
  if constexpr (QT == 2) asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
  if (warp == 0) {
    // ===================== TMA loader (whole warp; one elected lane issues, see ptx.cuh) =====================
    {
      const int kv_img = (p.kv_map != nullptr) ? p.kv_map[img] : img;
#pragma unroll
      for (int t = 0; t < QT; ++t) {
        mbar_arrive_expect_tx_w(&q_full[t], Cfg::kTileBytes);
#pragma unroll
        for (int a = 0; a < NA; ++a)
          tma_load_3d_w(q_smem + t * Cfg::kTileBytes + a * kAtomBytes, &p.q_map, &q_full[t], head * D + 64 * a,
                      q0 + 128 * t, img);
      }
      int slot = 0;
      uint32_t phase = 0;
      auto load_tile = [&](const CUtensorMap* map, int key0) {
        mbar_wait(&kv_empty[slot], phase ^ 1);
        mbar_arrive_expect_tx_w(&kv_full[slot], Cfg::kTileBytes);
#pragma unroll
        for (int a = 0; a < NA; ++a)
          tma_load_3d_w(kv_smem + slot * Cfg::kTileBytes + a * kAtomBytes, map, &kv_full[slot], head * D + 64 * a,
                      key0, kv_img);
        if (++slot == kSlots) {
          slot = 0;
          phase ^= 1;
        }
      };
      if constexpr (QT == 1) {
        // consumption order of the MMA warp: K0 K1 V0 K2 V1 ...
        for (int s = 0; s <= T; ++s) {
          if (s < T) load_tile(&p.k_map, s * 128);
          if (s >= 1) load_tile(&p.v_map, (s - 1) * 128);
        }
      } else {
        // K0 V0 K1 V1 ...
        for (int s = 0; s < T; ++s) {
          load_tile(&p.k_map, s * 128);
          load_tile(&p.v_map, s * 128);
        }
      }
    }
  } else if (warp == 1 || (Cfg::kSepP && warp == 2)) {
    // ===================== MMA issuer(s) (whole warp; one elected lane issues) =====================
    {
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
      // B (= V) is MN-major; only the head's own channels are multiplied: atom a covers min(64, D - 64a) channels
      // rounded up to the UMMA N granularity of 16 (d = 40 -> N = 48 instead of 64: fewer wasted tensor FLOPs)
      auto idesc_pv_of = [](int a) {
        const int rem = D - 64 * a;
        const int n = rem >= 64 ? 64 : ((rem + (Cfg::kSumTC ? 16 : 15)) & ~15);  // kSumTC: + the ones column
        return make_idesc_bf16(128, static_cast<uint32_t>(n), 0, 1);
      };
      auto qk = [&](uint32_t q_addr, uint32_t k_addr, uint32_t s_col) {
#pragma unroll
        for (int ks = 0; ks < Cfg::KQ; ++ks) {
          const uint32_t off = (ks >> 2) * kAtomBytes + (ks & 3) * 32;
          umma_ss_w(tmem_base + s_col, make_smem_desc(q_addr + off, 16, 1024, kLayoutSW128),
                  make_smem_desc(k_addr + off, 16, 1024, kLayoutSW128), idesc_qk, ks != 0);
        }
      };
      auto pv = [&](uint32_t p_col, uint32_t v_addr, uint32_t o_col, bool first) {
#pragma unroll
        for (int a = 0; a < NA; ++a) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            // V tile rows = keys (128 B pitch); 16 keys per k-step = 2048 B
            umma_ts_w(tmem_base + o_col + 64 * a, tmem_base + p_col + ks * 8,
                    make_smem_desc(v_addr + a * kAtomBytes + ks * 2048, kAtomBytes, 1024, kLayoutSW128), idesc_pv_of(a),
                    !(first && ks == 0));
          }
        }
      };
      if constexpr (QT == 1) {
        int slot = 0;
        uint32_t phase = 0;
        const uint32_t q_addr = smem_u32(q_smem);
        auto issue_qk = [&](int j) {
          mbar_wait(&kv_full[slot], phase);
          tc_fence_after();
          qk(q_addr, smem_u32(kv_smem + slot * Cfg::kTileBytes), (j & 1) * 128);
          umma_commit_w(&kv_empty[slot]);
          umma_commit_w(&s_full[j & 1]);
          if (++slot == kSlots) {
            slot = 0;
            phase ^= 1;
          }
        };
        mbar_wait(&q_ready[0], 0);
        tc_fence_after();
        issue_qk(0);
        for (int j = 0; j < T; ++j) {
          if (j + 1 < T) issue_qk(j + 1);
          mbar_wait(&p_full[j & 1], (j >> 1) & 1);
          mbar_wait(&kv_full[slot], phase);
          tc_fence_after();
          pv((j & 1) * 128, smem_u32(kv_smem + slot * Cfg::kTileBytes), Cfg::kColO, j == 0);
          umma_commit_w(&kv_empty[slot]);
          umma_commit_w(&o_done[0]);
          if (++slot == kSlots) {
            slot = 0;
            phase ^= 1;
          }
        }
      } else if constexpr (!Cfg::kSepP) {
        auto slot_of = [](int i) { return i % kSlots; };
        auto phase_of = [](int i) { return static_cast<uint32_t>((i / kSlots) & 1); };
        auto tile_addr = [&](int i) { return smem_u32(kv_smem + slot_of(i) * Cfg::kTileBytes); };
        mbar_wait(&kv_full[slot_of(0)], phase_of(0));
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&q_ready[t], 0);
          tc_fence_after();
          qk(smem_u32(q_smem + t * Cfg::kTileBytes), tile_addr(0), t * 128);
          umma_commit_w(&s_full[t]);
        }
        umma_commit_w(&kv_empty[slot_of(0)]);
        for (int j = 0; j < T; ++j) {
          const int iv = 2 * j + 1, ik = 2 * j + 2;
          const bool more = (j + 1 < T);
          mbar_wait(&kv_full[slot_of(iv)], phase_of(iv));
          if (more) mbar_wait(&kv_full[slot_of(ik)], phase_of(ik));
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            mbar_wait(&p_full[t], j & 1);
            tc_fence_after();
            pv(t * 128, tile_addr(iv), Cfg::kColO + t * 64 * NA, j == 0);
            umma_commit_w(&o_done[t]);
            if (more) {
              qk(smem_u32(q_smem + t * Cfg::kTileBytes), tile_addr(ik), t * 128);
              umma_commit_w(&s_full[t]);
            }
          }
          umma_commit_w(&kv_empty[slot_of(iv)]);
          if (more) umma_commit_w(&kv_empty[slot_of(ik)]);
        }
      } else {
        // kSepP: TWO issuing threads, one in-order stream each (warp 1: Q K^T, warp 2: P V).  A shared-memory
        // operation of the issuing thread (mbarrier test, LDS) right after a tcgen05.mma costs ~100 clk
        // (profiles/r01_ubench_issue_stream.log), so one thread polling four event sources starved the tensor
        // pipe; two blocking streams need one wait per 3 / 8 MMAs and overlap each other.
        //   QK_t(j): K_j landed, softmax warps have read S_t(j-1) [s_free];  PV_t(j): V_j landed, P_t(j) published
        auto slot_of = [](int i) { return i % kSlots; };
        auto phase_of = [](int i) { return static_cast<uint32_t>((i / kSlots) & 1); };
        auto tile_addr = [&](int i) { return smem_u32(kv_smem + slot_of(i) * Cfg::kTileBytes); };
        if (warp == 1) {
          for (int j = 0; j < T; ++j) {
            const int ik = 2 * j;
            mbar_wait(&kv_full[slot_of(ik)], phase_of(ik));
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              if (j == 0) mbar_wait(&q_ready[t], 0);
              else mbar_wait(&s_free[t], (j - 1) & 1);
              tc_fence_after();
              if (dbg_on && lane == 0 && j < 16) p.dbg[(1 * 16 + j) * 8 + t * 2] = clock64();
              qk(smem_u32(q_smem + t * Cfg::kTileBytes), tile_addr(ik), t * 128);
              umma_commit_w(&s_full[t]);
            }
            umma_commit_w(&kv_empty[slot_of(ik)]);  // both tiles have used K_j
          }
        } else {
          for (int j = 0; j < T; ++j) {
            const int iv = 2 * j + 1;
            mbar_wait(Cfg::kSumTC ? &v_ready[slot_of(iv)] : &kv_full[slot_of(iv)], phase_of(iv));
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              mbar_wait(&p_full[t], j & 1);
              tc_fence_after();
              if (dbg_on && lane == 0 && j < 16) p.dbg[(1 * 16 + j) * 8 + t * 2 + 1] = clock64();
              // P V_j in four pairs of k-steps; after each pair the 32 keys (16 TMEM columns) of P_t it read are free for
              // the softmax warps to overwrite with P_t(j+1): they never wait for the whole P V to retire
#pragma unroll
              for (int ks = 0; ks < 8; ++ks) {
                umma_ts_w(tmem_base + Cfg::kColO + t * 64, tmem_base + Cfg::kColP + t * 64 + ks * 8,
                          make_smem_desc(tile_addr(iv) + ks * 2048, kAtomBytes, 1024, kLayoutSW128), idesc_pv_of(0),
                          !(j == 0 && ks == 0));
                if (ks & 1) umma_commit_w(&p_free[t * 4 + (ks >> 1)]);
              }
              umma_commit_w(&o_done[t]);
            }
            umma_commit_w(&kv_empty[slot_of(iv)]);  // both tiles have used V_j
          }
        }
      }
    }
  } else if (Cfg::kSumTC && warp == 3) {
    // ===================== V fixer: ones into channel D of every V tile =====================
    // TMA's writes are visible once kv_full completes; the generic-proxy stores below are ordered before the
    // tensor core's reads by fence.proxy.async + the v_ready arrive the PV issuer waits for.
    const uint4 ones = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);  // 8 x bf16 1.0
    for (int j = 0; j < T; ++j) {
      const int iv = 2 * j + 1, slot = iv % kSlots;
      mbar_wait(&kv_full[slot], static_cast<uint32_t>((iv / kSlots) & 1));
      const uint32_t tile = smem_u32(kv_smem + slot * Cfg::kTileBytes) + (D / 64) * kAtomBytes;
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        const int row = rr * 32 + lane;
        sts128(tile + row * 128 + ((Cfg::kOnesChunk ^ (row & 7)) << 4), ones);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&v_ready[slot]);
    }
  }
  } else {
    // ===================== softmax / correction / epilogue =====================
    if constexpr (QT == 2) asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    const int quarter = warp & 3;
    const int tq = (QT == 2) ? ((warp - Cfg::kFirstSoftmaxWarp) >> 2) : 0;  // query tile of this warpgroup
    const int r = quarter * 32 + lane;                 // query row within the tile == TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t o_tmem = tmem_base + lane_base + Cfg::kColO + tq * 64 * NA;
    uint8_t* my_q = q_smem + tq * Cfg::kTileBytes;

    mbar_wait(&q_full[tq], 0);
    if constexpr (D % 16 != 0) {
      // zero Q channels [D, D+8): 16-byte chunk (D/8 % 8) of atom D/64, XOR-swizzled by row
      constexpr int atom = D / 64;
      constexpr int chunk = (D % 64) / 8;
      uint4* dst = reinterpret_cast<uint4*>(my_q + atom * kAtomBytes + r * 128 + ((chunk ^ (r & 7)) << 4));
      *dst = make_uint4(0, 0, 0, 0);
      fence_proxy_async_smem();
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&q_ready[tq]);

    float m_used = 0.f, l_run = 0.f;
    if constexpr (Cfg::kSepP) {
      // ------------------------------------------------------------------------------------------------------
      // Software-pipelined online softmax (two query tiles per CTA, head dims <= 64).
      // Invariant at the top of tile j: sv = S_t(j) in registers (ragged tail already masked), mt = its scaled row
      // max.  Inside the exps of chunk c of tile j, chunk c-1 of S_t(j+1) -- loaded into the registers that chunk
      // of S_t(j) just left -- is folded into the next tile's row max, so the FMNMX stream rides under the MUFU
      // stream and no phase of the loop leaves the exp unit idle: per tile only the last chunk's 11 FMNMX3, the
      // 8-to-1 reduction and the P hand-over remain outside the exps.  The exp2 of the pairs selected by kPolyMask2
      // run as a packed cubic on the FMA pipe; the row sum comes from the tensor core (kSumTC).
      // ------------------------------------------------------------------------------------------------------
      const uint32_t s_tmem = tmem_base + lane_base + tq * 128;
      const uint32_t p_tmem = tmem_base + lane_base + Cfg::kColP + tq * 64;
      uint64_t* od = &o_done[tq];
      const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
      uint32_t sv[4][32];
      float mxn[8];
      auto max_reset = [&]() {
#pragma unroll
        for (int i = 0; i < 8; ++i) mxn[i] = -INFINITY;
      };
      auto max_chunk = [&](uint32_t (&v)[32]) {
#pragma unroll
        for (int i = 0; i < 32; ++i) mxn[i & 7] = fmaxf(mxn[i & 7], __uint_as_float(v[i]));
      };
      auto max_final = [&]() {
        return fmaxf(fmaxf(fmaxf(mxn[0], mxn[1]), fmaxf(mxn[2], mxn[3])),
                     fmaxf(fmaxf(mxn[4], mxn[5]), fmaxf(mxn[6], mxn[7]))) * p.scale_log2;
      };
      auto mask_chunk = [&](uint32_t (&v)[32], const int c, const int valid) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c * 32 + i >= valid) v[i] = 0xff800000u;  // -inf
      };
      if (kPingPong2 && tq == 1) named_bar_arrive(1, 256);  // warpgroup 0 owns the exp unit first

      // prologue: S_t(0) -> registers, row max
      mbar_wait(&s_full[tq], 0);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(s_tmem + c * 32, sv[c]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[tq]);  // the tensor core may overwrite S_t with Q_t K_1^T
      max_reset();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (T == 1 && (p.nk % 128) != 0) mask_chunk(sv[c], c, p.nk);
        max_chunk(sv[c]);
      }
      float mt = max_final();

      auto tile = [&](const int j, auto has_next_tag, auto mask_next_tag) {
        constexpr bool kHasNext = decltype(has_next_tag)::value;
        constexpr bool kMaskNext = decltype(mask_next_tag)::value;
        SG_TS(0);
        float alpha = 1.f;
        bool need = false;
        if (j == 0) {
          m_used = mt;
        } else if (mt > m_used + 8.f) {  // lazy: rescale O (and with it the row sum) only when the max grew by > 2^8
          alpha = fast_exp2(m_used - mt);
          m_used = mt;
          need = true;
        }
        const float2 nm2 = make_float2(-m_used, -m_used);
        [[maybe_unused]] const int valid_next = p.nk - (j + 1) * 128;
        [[maybe_unused]] float2 sum2[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f),
                                           make_float2(0.f, 0.f)};
        max_reset();
        if (kPingPong2) named_bar_sync(1 + tq, 256);
        if (kSkew2 > 0 && j == 0 && tq == 1) named_bar_sync(3, 256);
        SG_TS(4);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t pk[16];
          bool pf_done = true;
          [[maybe_unused]] bool sf_done = true;
          if constexpr (kOrdered2 && Cfg::kSumTC) {
            float2 xq[16], eq[16];
            xq[0] = ffma2_o(make_float2(__uint_as_float(sv[c][0]), __uint_as_float(sv[c][1])), sc2, nm2);
            xq[1] = ffma2_o(make_float2(__uint_as_float(sv[c][2]), __uint_as_float(sv[c][3])), sc2, nm2);
#pragma unroll
            for (int t = 0; t < 16; ++t) {
              if (t == 8) {
                pf_done = (j > 0) ? mbar_test(&p_free[tq * 4 + c], (j - 1) & 1) : true;
                if constexpr (kHasNext) sf_done = (c == 0) ? mbar_test(&s_full[tq], (j + 1) & 1) : true;
                if constexpr (kHasNext) {
                  if (c >= 1) {
                    tmem_ld_wait();
                    if constexpr (kMaskNext) mask_chunk(sv[c - 1], c - 1, valid_next);
                    max_chunk(sv[c - 1]);
                  }
                }
              }
              if (t + 2 < 16)
                xq[t + 2] = ffma2_o(make_float2(__uint_as_float(sv[c][2 * t + 4]), __uint_as_float(sv[c][2 * t + 5])), sc2, nm2);
              if ((kPolyMask2 >> (t & 7)) & 1u) {
                eq[t] = exp2_fma2(xq[t]);
              } else {
                eq[t].x = fast_exp2_o(xq[t].x);
                eq[t].y = fast_exp2_o(xq[t].y);
              }
              if (t >= 1) pk[t - 1] = pack_bf16x2_o(eq[t - 1].x, eq[t - 1].y);
            }
            pk[15] = pack_bf16x2_o(eq[15].x, eq[15].y);
          } else {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            if (i == 16) {
              // Probe the barriers this chunk needs at its end NOW (non-blocking test_wait, ~100+ clk of latency) and look at
              // the answers after the remaining exps: a blocking try_wait at the point of use costs that latency every time,
              // even when the phase completed long ago, and there are five such waits per tile.
              pf_done = (j > 0) ? mbar_test(&p_free[tq * 4 + c], (j - 1) & 1) : true;
              if constexpr (kHasNext) sf_done = (c == 0) ? mbar_test(&s_full[tq], (j + 1) & 1) : true;
            }
            if constexpr (kHasNext) {
              if (c >= 1 && i == 16) {
                // chunk c-1 of S(j+1) was requested a chunk ago: fold it into the next tile's row max
                tmem_ld_wait();
                if constexpr (kMaskNext) mask_chunk(sv[c - 1], c - 1, valid_next);
                max_chunk(sv[c - 1]);
              }
            }
            const float2 x = ffma2(make_float2(__uint_as_float(sv[c][i]), __uint_as_float(sv[c][i + 1])), sc2, nm2);
            float2 e;
            if ((kPolyMask2 >> ((i >> 1) & 7)) & 1u) {
              e = exp2_fma2(x);
            } else {
              e.x = fast_exp2(x.x);
              e.y = fast_exp2(x.y);
            }
            if constexpr (!Cfg::kSumTC) sum2[(i >> 1) & 3] = fadd2(sum2[(i >> 1) & 3], e);
            pk[i >> 1] = pack_bf16x2(e.x, e.y);
          }
          }
          if (j > 0) {
            if (c == 0 && __any_sync(0xffffffffu, need)) {
              // rare: the row max grew by more than 2^8 -> rescale the O row (and with it the row sum) once P V_{j-1} has retired
              mbar_wait(od, (j - 1) & 1);
              tc_fence_after();
#pragma unroll
              for (int oc = 0; oc < 2; ++oc) {
                uint32_t ov[32];
                tmem_ld32(o_tmem + oc * 32, ov);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
                tmem_st32(o_tmem + oc * 32, ov);
              }
            }
            // the two k-steps of P V_{j-1} that read this chunk's columns of P have retired (issued ~a tile ago)
            if (!pf_done) mbar_wait(&p_free[tq * 4 + c], (j - 1) & 1);
            tc_fence_after();
          }
          tmem_st16(p_tmem + c * 16, pk);
          if (kPingPong2 && c == kHandoff && !(tq == 1 && j == T - 1)) named_bar_arrive(2 - tq, 256);
          if (kSkew2 > 0 && j == 0 && tq == 0 && c == kSkew2 - 1) named_bar_arrive(3, 256);
          if constexpr (kHasNext) {
            if (c == 0) {
              if (!sf_done) mbar_wait(&s_full[tq], (j + 1) & 1);
              tc_fence_after();
              SG_TS(1);
            }
            tmem_ld32(s_tmem + c * 32, sv[c]);  // S(j+1), chunk c -> the registers chunk c of S(j) just left
          }
        }
        SG_TS(5);
        if constexpr (!Cfg::kSumTC) {
          const float2 s01 = fadd2(sum2[0], sum2[1]), s23 = fadd2(sum2[2], sum2[3]);
          l_run = l_run * alpha + ((s01.x + s01.y) + (s23.x + s23.y));
        }
        // tail: the TMEM pipe is in order, so once the load of S(j+1)'s last chunk has landed the P stores issued before it
        // have landed too; the last 11 FMNMX3 of the next row max run before the (then free) store wait
        if constexpr (kHasNext) {
          tmem_ld_wait();
          if constexpr (kMaskNext) mask_chunk(sv[3], 3, valid_next);
          max_chunk(sv[3]);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&p_full[tq]);
          if constexpr (kHasNext) mbar_arrive(&s_free[tq]);  // S_t(j+1) is in registers
        }
        if constexpr (kHasNext) mt = max_final();
        SG_TS(6);
      };
      const bool ragged = (p.nk % 128) != 0;
      for (int j = 0; j + 2 < T; ++j) tile(j, std::true_type{}, std::false_type{});
      if (T >= 2) {
        if (ragged) tile(T - 2, std::true_type{}, std::true_type{});
        else tile(T - 2, std::true_type{}, std::false_type{});
      }
      tile(T - 1, std::false_type{}, std::false_type{});

      // final: O / l -> global (only the d valid channels of this head)
      mbar_wait(od, (T - 1) & 1);
      tc_fence_after();
      if constexpr (Cfg::kSumTC) {
        l_run = __uint_as_float(tmem_ld1(o_tmem + D));  // column D of O = sum_k P[:, k] (ones column of V)
        tmem_ld_wait();
      }
    } else {
    if constexpr (QT == 2) {
      if (kPingPong && tq == 1) named_bar_arrive(1, 256);  // warpgroup 0 owns the exp unit first
    }
    // one KV tile of the online softmax; kMasked is only instantiated for the LAST tile (ragged nk), so the
    // 128-element mask (ISETP + SEL per element once if-converted) stays out of the steady-state loop
    uint32_t sv[4][32];
    auto softmax_tile = [&](const int j, auto masked_tag) {
      constexpr bool kMasked = decltype(masked_tag)::value;
      const int sb = (QT == 2) ? tq : (j & 1);                         // S / P buffer
      [[maybe_unused]] const uint32_t s_par = (QT == 2) ? (j & 1) : ((j >> 1) & 1);
      uint64_t* od = &o_done[(QT == 2) ? tq : 0];
      const uint32_t s_tmem = tmem_base + lane_base + sb * 128;
      mbar_wait(&s_full[sb], s_par);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(s_tmem + c * 32, sv[c]);
      tmem_ld_wait();
      if constexpr (kMasked) {
        const int valid = p.nk - j * 128;  // keys valid in this tile (>= 1)
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i >= valid) sv[c][i] = 0xff800000u;  // -inf
      }
      // row max: 8 independent chains
      float mx8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) mx8[i] = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) mx8[i & 7] = fmaxf(mx8[i & 7], __uint_as_float(sv[c][i]));
      const float mx = fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])),
                             fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7])));
      const float mt = mx * p.scale_log2;
      float alpha = 1.f;
      bool need = false;
      if (j == 0) {
        m_used = mt;
      } else if (mt > m_used + 8.f) {
        alpha = fast_exp2(m_used - mt);
        m_used = mt;
        need = true;
      }
      auto wait_and_rescale = [&]() {
        // PV_{j-1} must have retired before P_j overwrites P_{j-1}/P_{j-2} and before the O row may be
        // rescaled.  o_done phases are consumed strictly in order (one wait per tile): an mbarrier parity
        // wait on a phase that has not started yet would return immediately.
        if (j > 0) {
          mbar_wait(od, (j - 1) & 1);
          tc_fence_after();
          if (__any_sync(0xffffffffu, need)) {
#pragma unroll
            for (int c = 0; c < 2 * NA; ++c) {
              uint32_t ov[32];
              tmem_ld32(o_tmem + c * 32, ov);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
              tmem_st32(o_tmem + c * 32, ov);
            }
          }
        }
      };
      if constexpr (QT == 2) wait_and_rescale();  // already satisfied: QK_{j} was issued after PV_{j-1}
      float2 sum2[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      const float2 nm2 = make_float2(-m_used, -m_used), sc2 = make_float2(p.scale_log2, p.scale_log2);
      if constexpr (QT == 2) {
        // Warpgroup ping-pong on the MUFU pipe: the two softmax warpgroups share each sub-partition's
        // 4-lane/clk exp unit.  Left alone they fall into lockstep (both in the exp phase, then both idle
        // behind the MMA chain); with the exp phase made exclusive, one group's TMEM load / row max / P store
        // / wait for the tensor core overlaps the other group's exps.  Barrier 1 + t admits warpgroup t.
        if (kPingPong) named_bar_sync(1 + tq, 256);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            // packed FFMA2 / FADD2: one issue slot scales two scores / accumulates two exps
            const float2 x = ffma2(make_float2(__uint_as_float(sv[c][i]), __uint_as_float(sv[c][i + 1])), sc2, nm2);
            // the MUFU pipe (16 exp2 / clk / SM) bounds this loop: every kPolyEvery-th exp runs on the FMA pipe instead
            const float e0 = fast_exp2(x.x);
            const float e1 = (kPolyEvery > 0 && ((i >> 1) % kPolyPairs) == 0) ? exp2_fma(x.y) : fast_exp2(x.y);
            sum2[(i >> 1) & 3] = fadd2(sum2[(i >> 1) & 3], make_float2(e0, e1));
            pk[i >> 1] = pack_bf16x2(e0, e1);
          }
          tmem_st16(s_tmem + c * 16, pk);
          // admit the other warpgroup (its tile j if we are group 0, tile j+1 if we are group 1)
          if (kPingPong && c == kHandoff && !(tq == 1 && j == T - 1)) named_bar_arrive(2 - tq, 256);
        }
      } else if constexpr (SHORT) {
        // single KV tile: no previous PV to overlap, P (aliasing S, which is in registers) is stored chunk by chunk
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float2 x = ffma2(make_float2(__uint_as_float(sv[c][i]), __uint_as_float(sv[c][i + 1])), sc2, nm2);
            const float e0 = fast_exp2(x.x);
            const float e1 = fast_exp2(x.y);
            sum2[(i >> 1) & 3] = fadd2(sum2[(i >> 1) & 3], make_float2(e0, e1));
            pk[i >> 1] = pack_bf16x2(e0, e1);
          }
          tmem_st16(s_tmem + c * 16, pk);
        }
      } else {
        uint32_t pk[2][32];
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float2 x = ffma2(make_float2(__uint_as_float(sv[c][i]), __uint_as_float(sv[c][i + 1])), sc2, nm2);
            const float e0 = fast_exp2(x.x);
            const float e1 = fast_exp2(x.y);
            sum2[(i >> 1) & 3] = fadd2(sum2[(i >> 1) & 3], make_float2(e0, e1));
            pk[c >> 1][(c & 1) * 16 + (i >> 1)] = pack_bf16x2(e0, e1);
          }
        wait_and_rescale();  // late: the exps above overlap PV_{j-1}
        tmem_st32(s_tmem, pk[0]);
        tmem_st32(s_tmem + 32, pk[1]);
      }
      const float2 s01 = fadd2(sum2[0], sum2[1]), s23 = fadd2(sum2[2], sum2[3]);
      l_run = l_run * alpha + ((s01.x + s01.y) + (s23.x + s23.y));
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[sb]);
    };
    for (int j = 0; j + 1 < T; ++j) softmax_tile(j, std::false_type{});
    if (p.nk % 128 != 0) {
      softmax_tile(T - 1, std::true_type{});
    } else {
      softmax_tile(T - 1, std::false_type{});
    }

    // final: O / l  -> global (only the d valid channels of this head)
    mbar_wait(&o_done[(QT == 2) ? tq : 0], (T - 1) & 1);
    tc_fence_after();
    }
    const float inv_l = 1.0f / l_run;
    const int q = q0 + 128 * tq + r;
    if (p.lse != nullptr && q < p.nq)  // log2-domain log-sum-exp of the row (for the backward pass)
      p.lse[(static_cast<long long>(img) * p.n_heads + head) * p.nq + q] = m_used + log2f(l_run);
    bf16* orow = p.o + static_cast<long long>(img) * p.o_stride_n + static_cast<long long>(q) * p.o_stride_t +
                 head * D;
#pragma unroll
    for (int c = 0; c < 2 * NA; ++c) {
      if (c * 32 < D) {
        uint32_t ov[32];
        tmem_ld32(o_tmem + c * 32, ov);
        tmem_ld_wait();
        if (q < p.nq) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (c * 32 + g * 8 < D) {
              uint4 w;
              w.x = pack_bf16x2(__uint_as_float(ov[8 * g + 0]) * inv_l, __uint_as_float(ov[8 * g + 1]) * inv_l);
              w.y = pack_bf16x2(__uint_as_float(ov[8 * g + 2]) * inv_l, __uint_as_float(ov[8 * g + 3]) * inv_l);
              w.z = pack_bf16x2(__uint_as_float(ov[8 * g + 4]) * inv_l, __uint_as_float(ov[8 * g + 5]) * inv_l);
              w.w = pack_bf16x2(__uint_as_float(ov[8 * g + 6]) * inv_l, __uint_as_float(ov[8 * g + 7]) * inv_l);
              *reinterpret_cast<uint4*>(orow + c * 32 + g * 8) = w;
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// =================================================================================================================
// attn_mc_kernel: the "many small CTAs" form of the core for head dims < 64 (d = 40 / 32: the 64x64-latent level, 5/6 of
// the attention time of a step).
//
// With d = 40 a 128 x 128 score tile costs the tensor core ~350 clk but the exp unit 1024 (16 exp2 / clk / SM) and the
// softmax warps ~600 issue slots each: the loop is bound by how well the softmax instruction streams of ONE sub-partition
// overlap each other's latencies (TMEM round trips, MUFU queue, barrier wake-ups).  attn_kernel<40, 2> keeps 2 softmax warps
// per sub-partition (two query tiles of one CTA) and measures 0.60 exp-unit utilisation however its loop is scheduled
// (profiles/r02_attn_variants.log).  This kernel shrinks the CTA instead: ONE 128-query tile, 64-key tiles, S (64 columns)
// with P written over it in place and O (d + ones column -> 48 columns) in 128 TMEM columns, Q + a 4-slot ring of 8 KB
// K / V tiles in 49 KB of shared memory, 8 warps with 96 registers for the 4 softmax warps -- so FOUR CTAs share an SM:
// 4 softmax warps per sub-partition, each from an independent CTA at an independent point of its loop, and the
// hardware scheduler does the interleaving the two-tile kernel tried to do by hand.
//   warp 0: TMA (Q once; K_j, V_j through the ring)        warp 1: tcgen05 issuer (QK_j -> S; PV_j: O += P V_j)
//   warp 2: writes the ones column into every V tile       warps 4-7: online softmax, thread = query row
// Per tile and CTA the chain is serial (S_j -> softmax -> P_j -> PV_j, QK_{j+1} -> S_{j+1}); the other three CTAs fill
// the gaps.  The row sum comes from the tensor core (ones column of V, as in attn_kernel), the rescale of O is lazy.
// =================================================================================================================
template <int D>
struct AttnMcCfg {
  static_assert(D < 64 && D % 8 == 0, "one 64-channel atom with room for the ones column");
  static constexpr int KQ = (D + 15) / 16;            // k-steps of Q K^T
  static constexpr int kNPV = (D + 16) & ~15;         // head channels + the ones column, UMMA N granularity
  static constexpr int kOnesChunk = D / 8;            // 16-byte chunk of a V row that holds channel D
  static constexpr int kTK = 64;                      // keys per tile
  static constexpr int kQBytes = 128 * 128;           // [128 queries x 64 channels] swizzle 128B
  static constexpr int kKvBytes = kTK * 128;          // [64 keys x 64 channels]
  static constexpr int kSlots = 4;                    // ring order K0 V0 K1 V1 ...
  static constexpr int kThreads = 256;
  static constexpr int kCtasPerSm = 4;
  static constexpr int kColO = 64;                    // S / P at column 0
  static constexpr int kTmemCols = 128;
  static constexpr int kSmemBytes = kQBytes + kSlots * kKvBytes + 1024 + 256;
  static_assert(kColO + kNPV <= kTmemCols, "S + O must fit 128 TMEM columns");
};

template <int D>
__global__ void __launch_bounds__(AttnMcCfg<D>::kThreads, AttnMcCfg<D>::kCtasPerSm)
    attn_mc_kernel(const __grid_constant__ AttnKParams p) {
  using Cfg = AttnMcCfg<D>;
  constexpr int kSlots = Cfg::kSlots;
  constexpr int TK = Cfg::kTK;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* q_smem = smem;
  uint8_t* kv_smem = smem + Cfg::kQBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(kv_smem + kSlots * Cfg::kKvBytes);
  uint64_t* kv_full = bars;                  // [kSlots] TMA -> MMA / V fixer
  uint64_t* kv_empty = bars + kSlots;        // [kSlots] MMA -> TMA
  uint64_t* v_ready = bars + 2 * kSlots;     // [kSlots] V fixer -> MMA
  uint64_t* q_full = bars + 3 * kSlots;      // TMA -> softmax warps
  uint64_t* q_ready = q_full + 1;            // softmax warps -> MMA (Q pad zeroed)
  uint64_t* s_full = q_full + 2;             // MMA -> softmax: S_j complete (and P V_{j-1} retired)
  uint64_t* p_full = q_full + 3;             // softmax -> MMA: P_j published
  uint64_t* o_done = q_full + 4;             // MMA -> softmax: last P V retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_full + 5);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const int head = blockIdx.y;
  const int img = blockIdx.z;
  const int T = (p.nk + TK - 1) / TK;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kSlots; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
      mbar_init(&v_ready[i], 1);
    }
    mbar_init(q_full, 1);
    mbar_init(q_ready, 4);
    mbar_init(s_full, 1);
    mbar_init(p_full, 4);
    mbar_init(o_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  auto slot_of = [](int i) { return i % kSlots; };
  auto phase_of = [](int i) { return static_cast<uint32_t>((i / kSlots) & 1); };

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    if (warp == 0) {
      // ===================== TMA loader =====================
      const int kv_img = (p.kv_map != nullptr) ? p.kv_map[img] : img;
      mbar_arrive_expect_tx_w(q_full, Cfg::kQBytes);
      tma_load_3d_w(q_smem, &p.q_map, q_full, head * D, q0, img);
      for (int i = 0; i < 2 * T; ++i) {
        const int slot = slot_of(i);
        mbar_wait(&kv_empty[slot], phase_of(i) ^ 1);
        mbar_arrive_expect_tx_w(&kv_full[slot], Cfg::kKvBytes);
        tma_load_3d_w(kv_smem + slot * Cfg::kKvBytes, (i & 1) ? &p.v_map : &p.k_map, &kv_full[slot], head * D,
                      (i >> 1) * TK, kv_img);
      }
    } else if (warp == 1) {
      // ===================== tcgen05 issuer =====================
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, TK, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, Cfg::kNPV, 0, 1);  // B (= V) is MN-major
      const uint32_t q_addr = smem_u32(q_smem);
      auto tile_addr = [&](int i) { return smem_u32(kv_smem + slot_of(i) * Cfg::kKvBytes); };
      auto qk = [&](int i) {
        mbar_wait(&kv_full[slot_of(i)], phase_of(i));
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < Cfg::KQ; ++ks)
          umma_ss_w(tmem_base, make_smem_desc(q_addr + ks * 32, 16, 1024, kLayoutSW128),
                    make_smem_desc(tile_addr(i) + ks * 32, 16, 1024, kLayoutSW128), idesc_qk, ks != 0);
        umma_commit_w(&kv_empty[slot_of(i)]);
        umma_commit_w(s_full);
      };
      mbar_wait(q_ready, 0);
      qk(0);
      for (int j = 0; j < T; ++j) {
        const int iv = 2 * j + 1;
        mbar_wait(&v_ready[slot_of(iv)], phase_of(iv));
        mbar_wait(p_full, j & 1);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < TK / 16; ++ks)  // 16 keys per k-step: 8 TMEM columns of P, 2048 B of the V tile
          umma_ts_w(tmem_base + Cfg::kColO, tmem_base + ks * 8,
                    make_smem_desc(tile_addr(iv) + ks * 2048, kAtomBytes, 1024, kLayoutSW128), idesc_pv,
                    !(j == 0 && ks == 0));
        umma_commit_w(&kv_empty[slot_of(iv)]);
        if (j + 1 < T) qk(iv + 1);  // overwrites P_j: the tensor pipe executes in issue order
        else umma_commit_w(o_done);
      }
    } else if (warp == 2) {
      // ===================== V fixer: ones into channel D of every V tile (row sums on the tensor core) =====================
      const uint4 ones = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);  // 8 x bf16 1.0
      for (int j = 0; j < T; ++j) {
        const int iv = 2 * j + 1, slot = slot_of(iv);
        mbar_wait(&kv_full[slot], phase_of(iv));
        const uint32_t tile = smem_u32(kv_smem + slot * Cfg::kKvBytes);
#pragma unroll
        for (int rr = 0; rr < TK / 32; ++rr) {
          const int row = rr * 32 + lane;
          sts128(tile + row * 128 + ((Cfg::kOnesChunk ^ (row & 7)) << 4), ones);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&v_ready[slot]);
      }
    }
  } else {
    // ===================== online softmax: thread = query row =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 96;");
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;  // query row within the tile == TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t s_tmem = tmem_base + lane_base;
    const uint32_t o_tmem = tmem_base + lane_base + Cfg::kColO;

    mbar_wait(q_full, 0);
    if constexpr (D % 16 != 0) {
      // zero Q channels [D, D+8): Q K^T runs over KQ * 16 channels, K's channels there belong to the next head
      uint4* dst = reinterpret_cast<uint4*>(q_smem + r * 128 + (((D / 8) ^ (r & 7)) << 4));
      *dst = make_uint4(0, 0, 0, 0);
      fence_proxy_async_smem();
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(q_ready);

    float m_used = 0.f;
    const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
    // one 64-key tile; kMasked is only instantiated for the LAST tile (ragged nk), so the mask stays out of the steady state
    auto tile = [&](const int j, auto masked_tag) {
      constexpr bool kMasked = decltype(masked_tag)::value;
      uint32_t sv[2][32];
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      tmem_ld32(s_tmem, sv[0]);
      tmem_ld32(s_tmem + 32, sv[1]);
      tmem_ld_wait();
      if constexpr (kMasked) {
        const int valid = p.nk - j * TK;  // keys valid in the last tile (>= 1)
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i >= valid) sv[c][i] = 0xff800000u;  // -inf
      }
      float mx8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) mx8[i] = fmaxf(__uint_as_float(sv[0][i]), __uint_as_float(sv[1][i]));
#pragma unroll
      for (int i = 8; i < 32; ++i)
        mx8[i & 7] = fmaxf(mx8[i & 7], fmaxf(__uint_as_float(sv[0][i]), __uint_as_float(sv[1][i])));
      const float mt = fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])),
                             fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7]))) * p.scale_log2;
      // lazy rescale: O (and its ones column = the running row sum) is only rescaled when the maximum grew by > 2^8
      bool need = false;
      float alpha = 1.f;
      if (j == 0) {
        m_used = mt;
      } else if (mt > m_used + 8.f) {
        alpha = fast_exp2(m_used - mt);
        m_used = mt;
        need = true;
      }
      if (j > 0 && __any_sync(0xffffffffu, need)) {
        // S_j complete implies P V_{j-1} retired (issue order), so O is quiescent; 16 columns at a time (register budget)
#pragma unroll 1
        for (int c = 0; c < Cfg::kNPV / 16; ++c) {
          uint32_t ow[16];
          tmem_ld16(o_tmem + c * 16, ow);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) ow[i] = __float_as_uint(__uint_as_float(ow[i]) * alpha);
          tmem_st16(o_tmem + c * 16, ow);
        }
      }
      const float2 nm2 = make_float2(-m_used, -m_used);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float2 x = ffma2(make_float2(__uint_as_float(sv[c][i]), __uint_as_float(sv[c][i + 1])), sc2, nm2);
          // the exp unit bounds the loop: the pairs selected by kMcPolyMask run as a packed cubic on the FMA pipe
          float2 e;
          if ((kMcPolyMask >> ((i >> 1) & 3)) & 1u) e = exp2_fma2(x);
          else e = make_float2(fast_exp2(x.x), fast_exp2(x.y));
          pk[i >> 1] = pack_bf16x2(e.x, e.y);
        }
        tmem_st16(s_tmem + c * 16, pk);  // P_j over S_j (this thread's row only; all of S_j is in registers)
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    };
    for (int j = 0; j + 1 < T; ++j) tile(j, std::false_type{});
    if (p.nk % TK != 0) tile(T - 1, std::true_type{});
    else tile(T - 1, std::false_type{});

    // final: O / l -> global (only the d valid channels of this head)
    mbar_wait(o_done, 0);
    tc_fence_after();
    const float l_run = __uint_as_float(tmem_ld1(o_tmem + D));  // column D of O = sum_k P[:, k]
    tmem_ld_wait();
    const float inv_l = 1.0f / l_run;
    const int q = q0 + r;
    if (p.lse != nullptr && q < p.nq)
      p.lse[(static_cast<long long>(img) * p.n_heads + head) * p.nq + q] = m_used + log2f(l_run);
    bf16* orow = p.o + static_cast<long long>(img) * p.o_stride_n + static_cast<long long>(q) * p.o_stride_t + head * D;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      if (c * 32 < D) {
        uint32_t ov[32];
        if (c == 0) {
          tmem_ld32(o_tmem, ov);
        } else {
          uint32_t ow[16];
          tmem_ld16(o_tmem + 32, ow);
#pragma unroll
          for (int i = 0; i < 16; ++i) ov[i] = ow[i];
        }
        tmem_ld_wait();
        if (q < p.nq) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (c * 32 + g * 8 < D) {
              uint4 w;
              w.x = pack_bf16x2(__uint_as_float(ov[8 * g + 0]) * inv_l, __uint_as_float(ov[8 * g + 1]) * inv_l);
              w.y = pack_bf16x2(__uint_as_float(ov[8 * g + 2]) * inv_l, __uint_as_float(ov[8 * g + 3]) * inv_l);
              w.z = pack_bf16x2(__uint_as_float(ov[8 * g + 4]) * inv_l, __uint_as_float(ov[8 * g + 5]) * inv_l);
              w.w = pack_bf16x2(__uint_as_float(ov[8 * g + 6]) * inv_l, __uint_as_float(ov[8 * g + 7]) * inv_l);
              *reinterpret_cast<uint4*>(orow + c * 32 + g * 8) = w;
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// =================================================================================================================
// attn_text_kernel: the text cross-attention (77 tokens) at head dims < 64, one K/V tile of up to 80 keys.
// A CTA is a pure latency chain (TMA -> Q K^T -> softmax -> P V -> store, ~4 us) with almost no work, so what counts is how
// many chains an SM holds: S (80 columns, P written over it in place) + O (d + ones column -> 48 columns) = 128 TMEM columns,
// 36 KB of shared memory, 6 warps -- four CTAs per SM instead of the two of attn_kernel<D, 1, SHORT>.  The softmax makes two
// passes over S in TMEM (row max, then exp / pack in 16-column chunks), so a thread never holds more than one chunk.
//   warp 0: TMA (Q, K, V) + the ones column of V      warp 1: tcgen05 issuer      warps 2-5: softmax, thread = query row
// =================================================================================================================
template <int D>
struct AttnTextCfg {
  static_assert(D < 64 && D % 8 == 0, "one 64-channel atom with room for the ones column");
  static constexpr int KQ = (D + 15) / 16;
  static constexpr int kNPV = (D + 16) & ~15;
  static constexpr int kOnesChunk = D / 8;
  static constexpr int kTK = 80;                      // keys per tile (5 UMMA k-steps of P V; Q K^T with N = 80)
  static constexpr int kQBytes = 128 * 128;
  static constexpr int kKvBytes = kTK * 128;
  static constexpr int kThreads = 192;
  static constexpr int kCtasPerSm = 4;
  static constexpr int kColO = kTK;
  static constexpr int kTmemCols = 128;
  static constexpr int kSmemBytes = kQBytes + 2 * kKvBytes + 1024 + 128;
  static_assert(kColO + kNPV <= kTmemCols, "S + O must fit 128 TMEM columns");
};

template <int D>
__global__ void __launch_bounds__(AttnTextCfg<D>::kThreads, AttnTextCfg<D>::kCtasPerSm)
    attn_text_kernel(const __grid_constant__ AttnKParams p) {
  using Cfg = AttnTextCfg<D>;
  constexpr int TK = Cfg::kTK;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* q_smem = smem;
  uint8_t* k_smem = smem + Cfg::kQBytes;
  uint8_t* v_smem = k_smem + Cfg::kKvBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(v_smem + Cfg::kKvBytes);
  uint64_t* q_full = bars;       // TMA -> softmax warps
  uint64_t* k_full = bars + 1;   // TMA -> MMA
  uint64_t* v_full = bars + 2;   // TMA -> ones column
  uint64_t* v_ready = bars + 3;  // ones column -> MMA
  uint64_t* q_ready = bars + 4;  // softmax warps -> MMA (Q pad zeroed)
  uint64_t* s_full = bars + 5;   // MMA -> softmax
  uint64_t* p_full = bars + 6;   // softmax -> MMA
  uint64_t* o_done = bars + 7;   // MMA -> softmax
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const int head = blockIdx.y;
  const int img = blockIdx.z;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    mbar_init(k_full, 1);
    mbar_init(v_full, 1);
    mbar_init(v_ready, 1);
    mbar_init(q_ready, 4);
    mbar_init(s_full, 1);
    mbar_init(p_full, 4);
    mbar_init(o_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    // ===================== TMA loader, then the ones column of V =====================
    const int kv_img = (p.kv_map != nullptr) ? p.kv_map[img] : img;
    mbar_arrive_expect_tx_w(q_full, Cfg::kQBytes);
    tma_load_3d_w(q_smem, &p.q_map, q_full, head * D, q0, img);
    mbar_arrive_expect_tx_w(k_full, Cfg::kKvBytes);
    tma_load_3d_w(k_smem, &p.k_map, k_full, head * D, 0, kv_img);
    mbar_arrive_expect_tx_w(v_full, Cfg::kKvBytes);
    tma_load_3d_w(v_smem, &p.v_map, v_full, head * D, 0, kv_img);
    mbar_wait(v_full, 0);
    const uint4 ones = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);  // 8 x bf16 1.0
    const uint32_t tile = smem_u32(v_smem);
#pragma unroll
    for (int rr = 0; rr < (TK + 31) / 32; ++rr) {
      const int row = rr * 32 + lane;
      if (row < TK) sts128(tile + row * 128 + ((Cfg::kOnesChunk ^ (row & 7)) << 4), ones);
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive(v_ready);
  } else if (warp == 1) {
    // ===================== tcgen05 issuer =====================
    constexpr uint32_t idesc_qk = make_idesc_bf16(128, TK, 0, 0);
    constexpr uint32_t idesc_pv = make_idesc_bf16(128, Cfg::kNPV, 0, 1);  // B (= V) is MN-major
    const uint32_t q_addr = smem_u32(q_smem), k_addr = smem_u32(k_smem), v_addr = smem_u32(v_smem);
    mbar_wait(q_ready, 0);
    mbar_wait(k_full, 0);
    tc_fence_after();
#pragma unroll
    for (int ks = 0; ks < Cfg::KQ; ++ks)
      umma_ss_w(tmem_base, make_smem_desc(q_addr + ks * 32, 16, 1024, kLayoutSW128),
                make_smem_desc(k_addr + ks * 32, 16, 1024, kLayoutSW128), idesc_qk, ks != 0);
    umma_commit_w(s_full);
    mbar_wait(v_ready, 0);
    mbar_wait(p_full, 0);
    tc_fence_after();
#pragma unroll
    for (int ks = 0; ks < TK / 16; ++ks)
      umma_ts_w(tmem_base + Cfg::kColO, tmem_base + ks * 8, make_smem_desc(v_addr + ks * 2048, kAtomBytes, 1024, kLayoutSW128),
                idesc_pv, ks != 0);
    umma_commit_w(o_done);
  } else {
    // ===================== softmax: thread = query row =====================
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t s_tmem = tmem_base + lane_base;
    const uint32_t o_tmem = tmem_base + lane_base + Cfg::kColO;

    mbar_wait(q_full, 0);
    if constexpr (D % 16 != 0) {
      uint4* dst = reinterpret_cast<uint4*>(q_smem + r * 128 + (((D / 8) ^ (r & 7)) << 4));
      *dst = make_uint4(0, 0, 0, 0);
      fence_proxy_async_smem();
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(q_ready);

    mbar_wait(s_full, 0);
    tc_fence_after();
    // pass 1: row max over the valid keys
    float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int c = 0; c < TK / 16; ++c) {
      uint32_t sv[16];
      tmem_ld16(s_tmem + c * 16, sv);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (c * 16 + i < p.nk) mx[i & 3] = fmaxf(mx[i & 3], __uint_as_float(sv[i]));
    }
    const float m_used = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * p.scale_log2;
    const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), nm2 = make_float2(-m_used, -m_used);
    // pass 2: P = 2^(scale * S - max) as bf16 over S, chunk by chunk (P columns [8c, 8c + 8) lie below the S columns still unread)
#pragma unroll
    for (int c = 0; c < TK / 16; ++c) {
      uint32_t sv[16], pk[8];
      tmem_ld16(s_tmem + c * 16, sv);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        const float2 x = ffma2(make_float2(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])), sc2, nm2);
        const float e0 = (c * 16 + i < p.nk) ? fast_exp2(x.x) : 0.f;
        const float e1 = (c * 16 + i + 1 < p.nk) ? fast_exp2(x.y) : 0.f;
        pk[i >> 1] = pack_bf16x2(e0, e1);
      }
      tmem_st8(s_tmem + c * 8, pk);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(p_full);

    // O / l -> global (only the d valid channels of this head)
    mbar_wait(o_done, 0);
    tc_fence_after();
    const int q = q0 + r;
    bf16* orow = p.o + static_cast<long long>(img) * p.o_stride_n + static_cast<long long>(q) * p.o_stride_t + head * D;
    float inv_l = 0.f;
#pragma unroll
    for (int c = Cfg::kNPV / 16 - 1; c >= 0; --c) {  // the chunk with column D (the row sum) first
      uint32_t ov[16];
      tmem_ld16(o_tmem + c * 16, ov);
      tmem_ld_wait();
      if (c == D / 16) {
        const float l_run = __uint_as_float(ov[D % 16]);
        inv_l = 1.0f / l_run;
        if (p.lse != nullptr && q < p.nq)
          p.lse[(static_cast<long long>(img) * p.n_heads + head) * p.nq + q] = m_used + log2f(l_run);
      }
      if (q < p.nq) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (c * 16 + g * 8 < D) {
            uint4 w;
            w.x = pack_bf16x2(__uint_as_float(ov[8 * g + 0]) * inv_l, __uint_as_float(ov[8 * g + 1]) * inv_l);
            w.y = pack_bf16x2(__uint_as_float(ov[8 * g + 2]) * inv_l, __uint_as_float(ov[8 * g + 3]) * inv_l);
            w.z = pack_bf16x2(__uint_as_float(ov[8 * g + 4]) * inv_l, __uint_as_float(ov[8 * g + 5]) * inv_l);
            w.w = pack_bf16x2(__uint_as_float(ov[8 * g + 6]) * inv_l, __uint_as_float(ov[8 * g + 7]) * inv_l);
            *reinterpret_cast<uint4*>(orow + c * 16 + g * 8) = w;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int D, int QT, bool DBG = false, bool SHORT = false>
static int launch_attn(const AttnKParams& kp, int nq, int heads, int lanes, cudaStream_t stream) {
  using Cfg = AttnCfg<D, QT, SHORT>;
  if constexpr (!DBG && D == 40 && QT == 2) {
    if (kp.dbg != nullptr) return launch_attn<D, QT, true>(kp, nq, heads, lanes, stream);  // bring-up timeline build
  }
  static std::atomic<uint64_t> configured{0};
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(attn_kernel<D, QT, DBG, SHORT>), Cfg::kSmemBytes,
                                   &configured))
    return rc;
  dim3 grid((nq + 128 * QT - 1) / (128 * QT), heads, lanes);
  attn_kernel<D, QT, DBG, SHORT><<<grid, Cfg::kThreads, Cfg::kSmemBytes, stream>>>(kp);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

template <int D>
static int launch_attn_mc(const AttnKParams& kp, int nq, int heads, int lanes, cudaStream_t stream) {
  using Cfg = AttnMcCfg<D>;
  static std::atomic<uint64_t> configured{0};
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(attn_mc_kernel<D>), Cfg::kSmemBytes, &configured)) return rc;
  dim3 grid((nq + 127) / 128, heads, lanes);
  attn_mc_kernel<D><<<grid, Cfg::kThreads, Cfg::kSmemBytes, stream>>>(kp);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

// attn_mc_kernel is OPT-IN (SG_ATTN_MC=1).  Timed alone in short bursts (1965 MHz) it is 9 % faster than attn_kernel<40, 2>
// (1254 vs 1383 SM-clk per 128 x 128 tile, profiles/r02_attn_mc_bench.log); but a denoising step runs at the board's power
// cap, and there the joules per launch decide: 1.39 J at 1725 MHz against 1.35 J at 1880 MHz for the two-tile kernel in a
// sustained loop (profiles/r02_attn_power.log: twice the K / V traffic from L2, narrower MMAs), i.e. 3 % SLOWER, and the step
// time of bench.py is unchanged.  The two-tile kernel stays the default.
static bool use_attn_mc(int head_dim, int nk) {
  static const bool enabled = [] {
    const char* e = getenv("SG_ATTN_MC");
    return e != nullptr && e[0] == '1';
  }();
  return enabled && (head_dim == 40 || head_dim == 32) && nk > 128;
}

template <int D>
static int launch_attn_text(const AttnKParams& kp, int nq, int heads, int lanes, cudaStream_t stream) {
  using Cfg = AttnTextCfg<D>;
  static std::atomic<uint64_t> configured{0};
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(attn_text_kernel<D>), Cfg::kSmemBytes, &configured)) return rc;
  dim3 grid((nq + 127) / 128, heads, lanes);
  attn_text_kernel<D><<<grid, Cfg::kThreads, Cfg::kSmemBytes, stream>>>(kp);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

// head dims 40 / 32 with at most 80 keys (the 77-token text cross-attention): attn_text_kernel.  SG_ATTN_TEXT=0 disables (A/B runs).
static bool use_attn_text(int head_dim, int nk) {
  static const bool enabled = [] {
    const char* e = getenv("SG_ATTN_TEXT");
    return e == nullptr || e[0] != '0';
  }();
  return enabled && (head_dim == 40 || head_dim == 32) && nk <= 80;
}

// two query tiles per CTA whenever there are at least two (and the head dim leaves TMEM room)
template <int D>
static int launch_attn_auto(const AttnKParams& kp, int nq, int heads, int lanes, cudaStream_t stream) {
  if constexpr (D < 64) {
    if (use_attn_text(D, kp.nk)) return launch_attn_text<D>(kp, nq, heads, lanes, stream);
    if (use_attn_mc(D, kp.nk)) return launch_attn_mc<D>(kp, nq, heads, lanes, stream);
  }
  if constexpr (D <= 128) {
    if (kp.nk <= 128) return launch_attn<D, 1, false, true>(kp, nq, heads, lanes, stream);  // text cross-attention
    if (nq > 128) return launch_attn<D, 2>(kp, nq, heads, lanes, stream);
  }
  return launch_attn<D, 1>(kp, nq, heads, lanes, stream);
}

static int make_tok_map(CUtensorMap* m, const void* ptr, int channels, int tokens, int lanes, int64_t stride_t,
                        int64_t stride_n, uint32_t box_tokens = 128) {
  uint64_t dims[3] = {static_cast<uint64_t>(channels), static_cast<uint64_t>(tokens),
                      static_cast<uint64_t>(lanes)};
  uint64_t strides[2] = {static_cast<uint64_t>(stride_t) * 2, static_cast<uint64_t>(stride_n) * 2};
  uint32_t box[3] = {64, box_tokens, 1};
  return encode_tmap_bf16(m, ptr, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace sg

using namespace sg;

extern "C" int sg_attention(const sg_attn_args* a, void* stream_v) {
  if (a == nullptr || !a->q || !a->k || !a->v || !a->o) return SG_ERR_ARG;
  if (a->nq < 1 || a->nk < 1 || a->n_lanes < 1 || a->n_heads < 1) return SG_ERR_SHAPE;
  switch (a->head_dim) {
    case 32: case 40: case 64: case 80: case 128: case 160: break;
    default: return SG_ERR_SHAPE;
  }
  if ((reinterpret_cast<uintptr_t>(a->o) & 15) || (a->o_stride_t % 8) || (a->o_stride_n % 8)) return SG_ERR_ALIGN;
  const int C = a->n_heads * a->head_dim;
  const int kv_lanes = a->n_kv_lanes > 0 ? a->n_kv_lanes : a->n_lanes;
  AttnKParams kp;
  memset(&kp, 0, sizeof(kp));
  int rc;
  if ((rc = make_tok_map(&kp.q_map, a->q, C, a->nq, a->n_lanes, a->q_stride_t, a->q_stride_n)) != 0) return rc;
  // keys per K / V tile of the kernel that will run
  const uint32_t kv_box = use_attn_text(a->head_dim, a->nk) ? 80 : (use_attn_mc(a->head_dim, a->nk) ? 64 : 128);
  if ((rc = make_tok_map(&kp.k_map, a->k, C, a->nk, kv_lanes, a->k_stride_t, a->k_stride_n, kv_box)) != 0) return rc;
  if ((rc = make_tok_map(&kp.v_map, a->v, C, a->nk, kv_lanes, a->v_stride_t, a->v_stride_n, kv_box)) != 0) return rc;
  kp.o = static_cast<bf16*>(a->o);
  kp.o_stride_n = a->o_stride_n;
  kp.o_stride_t = a->o_stride_t;
  kp.kv_map = a->kv_map;
  kp.lse = a->lse;
  kp.dbg = reinterpret_cast<long long*>(a->debug_timeline);
  kp.nq = a->nq;
  kp.nk = a->nk;
  kp.n_heads = a->n_heads;
  kp.scale_log2 = a->scale * 1.4426950408889634f;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const int nq = a->nq, nh = a->n_heads, nl = a->n_lanes;
  switch (a->head_dim) {
    case 32: return launch_attn_auto<32>(kp, nq, nh, nl, stream);
    case 40: return launch_attn_auto<40>(kp, nq, nh, nl, stream);
    case 64: return launch_attn_auto<64>(kp, nq, nh, nl, stream);
    case 80: return launch_attn_auto<80>(kp, nq, nh, nl, stream);
    case 128: return launch_attn_auto<128>(kp, nq, nh, nl, stream);
    default: return launch_attn_auto<160>(kp, nq, nh, nl, stream);
  }
}
