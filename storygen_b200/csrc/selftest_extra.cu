// Self-test cases for the attention core and the small kernels (see selftest.cu).
#include "selftest_util.h"

static void run_attn_case(const char* name, int D, int H, int n_lanes, int n_kv_lanes, int nq, int nk,
                          float q_scale, bool fused_qkv, const std::vector<int>& kv_map_h) {
  if (!want(name)) return;
  const int C = H * D;
  HostBf Q, K, V, O, QKV;
  sg_attn_args a;
  memset(&a, 0, sizeof(a));
  const float* qf;
  const float* kf;
  const float* vf;
  long long q_st, k_st, v_st, q_sn, k_sn, v_sn;
  if (fused_qkv) {
    // self-attention reading a fused [lanes, tokens, 3C] projection output
    QKV.init(1ll * n_lanes * nq * 3 * C, q_scale);
    a.q = QKV.d;
    a.k = QKV.d + C;
    a.v = QKV.d + 2 * C;
    q_st = k_st = v_st = 3 * C;
    q_sn = k_sn = v_sn = 1ll * nq * 3 * C;
    qf = QKV.f.data();
    kf = QKV.f.data() + C;
    vf = QKV.f.data() + 2 * C;
  } else {
    Q.init(1ll * n_lanes * nq * C, q_scale);
    K.init(1ll * n_kv_lanes * nk * C, 1.f);
    V.init(1ll * n_kv_lanes * nk * C, 1.f);
    a.q = Q.d;
    a.k = K.d;
    a.v = V.d;
    q_st = k_st = v_st = C;
    q_sn = 1ll * nq * C;
    k_sn = v_sn = 1ll * nk * C;
    qf = Q.f.data();
    kf = K.f.data();
    vf = V.f.data();
  }
  O.zeros(1ll * n_lanes * nq * C);
  a.o = O.d;
  a.q_stride_n = q_sn;
  a.q_stride_t = q_st;
  a.k_stride_n = k_sn;
  a.k_stride_t = k_st;
  a.v_stride_n = v_sn;
  a.v_stride_t = v_st;
  a.o_stride_n = 1ll * nq * C;
  a.o_stride_t = C;
  a.n_lanes = n_lanes;
  a.n_heads = H;
  a.head_dim = D;
  a.nq = nq;
  a.nk = nk;
  a.n_kv_lanes = n_kv_lanes;
  a.scale = 1.0f / sqrtf(static_cast<float>(D));
  int* kv_map_d = nullptr;
  if (!kv_map_h.empty()) {
    CK(cudaMalloc(&kv_map_d, kv_map_h.size() * 4));
    CK(cudaMemcpy(kv_map_d, kv_map_h.data(), kv_map_h.size() * 4, cudaMemcpyHostToDevice));
    a.kv_map = kv_map_d;
  }
  int rc = sg_attention(&a, nullptr);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    report(name, rc, 1e30, 0, 0);
    return;
  }
  O.fetch();
  double max_err = 0, max_ref = 0;
  std::vector<float> s(nk);
  std::vector<float> acc(D);
  for (int n = 0; n < n_lanes; ++n) {
    const int kn = kv_map_h.empty() ? n : kv_map_h[n];
    for (int h = 0; h < H; ++h)
      for (int i = 0; i < nq; ++i) {
        const float* qp = qf + n * q_sn + i * q_st + h * D;
        float mx = -1e30f;
        for (int j = 0; j < nk; ++j) {
          const float* kp = kf + kn * k_sn + j * k_st + h * D;
          float d = 0.f;
          for (int c = 0; c < D; ++c) d += qp[c] * kp[c];
          s[j] = d * a.scale;
          mx = fmaxf(mx, s[j]);
        }
        float l = 0.f;
        for (int c = 0; c < D; ++c) acc[c] = 0.f;
        for (int j = 0; j < nk; ++j) {
          float pj = expf(s[j] - mx);
          l += pj;
          const float* vp = vf + kn * v_sn + j * v_st + h * D;
          for (int c = 0; c < D; ++c) acc[c] += pj * vp[c];
        }
        for (int c = 0; c < D; ++c) {
          float ref = acc[c] / l;
          float got = O.f[(1ll * n * nq + i) * C + h * D + c];
          max_err = fmax(max_err, fabs(static_cast<double>(ref) - got));
          max_ref = fmax(max_ref, fabs(static_cast<double>(ref)));
        }
      }
  }
  report(name, rc, max_err, max_ref, 0.02);
  if (kv_map_d) cudaFree(kv_map_d);
}

void test_ubench();
void test_ubench_mma();
void test_ubench_cp();
void test_ubench_mma_multi();
void test_ubench_mma_var();
void test_ubench_tma();
void test_ubench_softmax_mix();
void test_timeline();
void test_gemm_timeline();
void test_extra() {
  test_ubench();
  test_ubench_mma();
  test_ubench_cp();
  test_ubench_mma_multi();
  test_ubench_mma_var();
  test_ubench_tma();
  test_ubench_softmax_mix();
  test_timeline();
  test_gemm_timeline();
  run_attn_case("attn d40 self 2x8 nq=nk=256", 40, 8, 2, 2, 256, 256, 1.f, false, {});
  run_attn_case("attn d40 1 tile nq=128 nk=128 h=1", 40, 1, 1, 1, 128, 128, 1.f, false, {});
  run_attn_case("attn d40 text nq=256 nk=77 (masked tail)", 40, 8, 2, 2, 256, 77, 1.f, false, {});
  run_attn_case("attn d40 image nq=128 nk=768 kv_map", 40, 8, 3, 2, 128, 768, 1.f, false, {0, 1, 1});
  run_attn_case("attn d40 fused qkv nq=nk=384", 40, 8, 1, 1, 384, 384, 1.f, true, {});
  run_attn_case("attn d40 peaked scores (rescale path) nk=1024", 40, 2, 1, 1, 128, 1024, 12.f, false, {});
  run_attn_case("attn d80 self nq=nk=256", 80, 8, 2, 2, 256, 256, 1.f, false, {});
  run_attn_case("attn d80 text nk=77 fused-stride", 80, 8, 1, 1, 128, 77, 1.f, false, {});
  run_attn_case("attn d160 mid nq=64 nk=192", 160, 8, 2, 2, 64, 192, 1.f, false, {});
  run_attn_case("attn d160 self nq=nk=256 peaked", 160, 8, 1, 1, 256, 256, 6.f, false, {});
  run_attn_case("attn long d160 nq=256 nk=1536", 160, 4, 1, 1, 256, 1536, 3.f, false, {});
  run_attn_case("attn long d80 nq=256 nk=3072", 80, 2, 1, 1, 256, 3072, 3.f, false, {});
  run_attn_case("attn long d40 nq=512 nk=4096", 40, 2, 1, 1, 512, 4096, 4.f, false, {});
}

// ---------------------------------------------------------------------------------------------
// micro-benchmarks (./selftest "ubench"): TMEM read throughput, MUFU.EX2 throughput
// ---------------------------------------------------------------------------------------------
#include "common.cuh"
#include "ptx.cuh"
__global__ void ubench_tmem_ld(int iters, long long* cycles, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) sg::tmem_alloc(&slot, 512);
  sg::tc_fence_before();
  __syncthreads();
  sg::tc_fence_after();
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t v[4][32];
#pragma unroll
    for (int c = 0; c < 4; ++c) sg::tmem_ld32(base + ((it & 1) * 256) + c * 32, v[c]);
    sg::tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 4; ++c) acc += __uint_as_float(v[c][0]) + __uint_as_float(v[c][17]) + __uint_as_float(v[c][31]);
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) sg::tmem_dealloc(slot, 512);
}
__global__ void ubench_mufu(int iters, long long* cycles, float* sink) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = -0.001f * (threadIdx.x + i);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = sg::fast_exp2(x[i]) - 1.0009f;
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  float a = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) a += x[i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = a;
}
void test_timeline() {
  if (!want("timeline")) return;
  const int D = 40, H = 8, C = H * D, nq = 4096, nk = 4096, lanes = 3;
  HostBf Q, K, V, O;
  Q.init(1ll * lanes * nq * C, 1.f);
  K.init(1ll * lanes * nk * C, 1.f);
  V.init(1ll * lanes * nk * C, 1.f);
  O.zeros(1ll * lanes * nq * C);
  long long* dbg;
  CK(cudaMalloc(&dbg, 12 * 16 * 8 * 8));
  CK(cudaMemset(dbg, 0, 12 * 16 * 8 * 8));
  sg_attn_args a;
  memset(&a, 0, sizeof(a));
  a.q = Q.d; a.k = K.d; a.v = V.d; a.o = O.d;
  a.q_stride_n = 1ll * nq * C; a.q_stride_t = C; a.k_stride_n = 1ll * nk * C; a.k_stride_t = C;
  a.v_stride_n = 1ll * nk * C; a.v_stride_t = C; a.o_stride_n = 1ll * nq * C; a.o_stride_t = C;
  a.n_lanes = lanes; a.n_heads = H; a.head_dim = D; a.nq = nq; a.nk = nk; a.scale = 1.f / sqrtf(40.f);
  a.debug_timeline = reinterpret_cast<int64_t*>(dbg);
  for (int rep = 0; rep < 2; ++rep) { int rc = sg_attention(&a, nullptr); CK(cudaDeviceSynchronize()); if (rc) printf("rc=%d\n", rc); }
  std::vector<long long> h(12 * 16 * 8);
  CK(cudaMemcpy(h.data(), dbg, h.size() * 8, cudaMemcpyDeviceToHost));
  long long t0 = h[(4 * 16 + 2) * 8 + 0];
  printf("[timeline] clocks relative to warp4 tile2 start; slots: 0=tile start 1=S(j+1) ready (after chunk 0) 4=exps start 5=all chunks issued 6=P published + next row max done\n");
  for (int w : {4, 8}) for (int j = 2; j < 8; ++j) {
    printf("[timeline] warp %2d tile %d:", w, j);
    for (int k = 0; k < 7; ++k) printf(" %7lld", h[(w * 16 + j) * 8 + k] - t0);
    printf("\n");
  }
  for (int j = 2; j < 8; ++j)
    printf("[timeline] mma issue tile %d: QK_0 %7lld  PV_0 %7lld  QK_1 %7lld  PV_1 %7lld\n", j, h[(16 + j) * 8 + 0] - t0,
           h[(16 + j) * 8 + 1] - t0, h[(16 + j) * 8 + 2] - t0, h[(16 + j) * 8 + 3] - t0);
  cudaFree(dbg);
}
void test_gemm_timeline() {
  if (!want("gtimeline")) return;
  const int M = getenv("GT_M") ? atoi(getenv("GT_M")) : 98304, N = getenv("GT_N") ? atoi(getenv("GT_N")) : 2560,
            K = getenv("GT_K") ? atoi(getenv("GT_K")) : 320;
  HostBf A, B, O;
  A.init(1ll * M * K, 1.f);
  B.init(1ll * N * K, 0.05f);
  O.zeros(1ll * M * N);
  long long* dbg;
  CK(cudaMalloc(&dbg, 10 * 8 * 8 * 8));
  CK(cudaMemset(dbg, 0, 10 * 8 * 8 * 8));
  sg_gemm_args a;
  memset(&a, 0, sizeof(a));
  a.num_src = 1; a.src[0].ptr = A.d; a.src[0].dim_c = K; a.src[0].dim_w = M; a.src[0].dim_h = 1; a.src[0].dim_n = 1;
  a.src[0].stride_w = K; a.src[0].stride_h = 1ll * M * K; a.src[0].stride_n = 1ll * M * K;
  a.num_segs = 1; a.segs[0].src = 0; a.segs[0].nkb = K / 64;
  a.out_w = M; a.out_h = 1; a.out_n = 1; a.bw = 128; a.bh = 1; a.bn = 1;
  a.b = B.d; a.N = N; a.Ktot = K; a.mode = SG_EPI_LINEAR; a.out = O.d; a.ldo = N; a.debug_timeline = reinterpret_cast<int64_t*>(dbg);
  for (int rep = 0; rep < 2; ++rep) { int rc = sg_gemm(&a, nullptr); CK(cudaDeviceSynchronize()); if (rc) printf("rc=%d\n", rc); }
  std::vector<long long> h(10 * 8 * 8);
  CK(cudaMemcpy(h.data(), dbg, h.size() * 8, cudaMemcpyDeviceToHost));
  const long long t0 = h[(1 * 8 + 1) * 8 + 0];
  printf("[gtimeline] M=%d N=%d K=%d LINEAR; clocks relative to MMA start of tile 1\n", M, N, K);
  for (int j = 1; j < 6; ++j)
    printf("[gtimeline] tile %d: mma start %7lld issued %7lld | epi warp2: enter %7lld res-prefetch-done %7lld acc-ready %7lld done %7lld | epi warp6: enter %7lld acc-ready %7lld done %7lld\n", j,
           h[(8 + j) * 8 + 0] - t0, h[(8 + j) * 8 + 1] - t0, h[(2 * 8 + j) * 8 + 0] - t0, h[(2 * 8 + j) * 8 + 1] - t0,
           h[(2 * 8 + j) * 8 + 2] - t0, h[(2 * 8 + j) * 8 + 3] - t0, h[(6 * 8 + j) * 8 + 0] - t0, h[(6 * 8 + j) * 8 + 2] - t0,
           h[(6 * 8 + j) * 8 + 3] - t0);
  cudaFree(dbg);
}
void test_ubench() {
  if (!want("ubench")) return;
  long long* cyc;
  float* sink;
  CK(cudaMalloc(&cyc, 8 * 1024));
  CK(cudaMalloc(&sink, 4 * 1024 * 1024));
  for (int warps : {4, 8, 16}) {
    const int iters = 2000;
    ubench_tmem_ld<<<148, warps * 32>>>(iters, cyc, sink);
    CK(cudaDeviceSynchronize());
    long long h;
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    const double bytes = static_cast<double>(iters) * warps * 4 * 32 * 32 * 4;  // per SM
    printf("[ubench] tmem_ld 32x32b.x32: %2d warps/SM: %.1f cycles per 128-col row-block per warp, %.1f B/clk/SM\n", warps,
           static_cast<double>(h) / iters, bytes / h);
  }
  for (int warps : {4, 8, 16}) {
    const int iters = 4000;
    ubench_mufu<<<148, warps * 32>>>(iters, cyc, sink);
    CK(cudaDeviceSynchronize());
    long long h;
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("[ubench] ex2.approx.ftz.f32 (+1 FADD): %2d warps/SM: %.2f lane-results/clk/SM\n", warps,
           static_cast<double>(iters) * 16 * warps * 32 / h);
  }
  cudaFree(cyc);
  cudaFree(sink);
}

// tcgen05.mma issue/latency: `chains` independent accumulators used round-robin (1 = fully dependent stream)
template <int N>
__global__ void ubench_mma(int iters, int chains, int use_ts, long long* cycles) {
  extern __shared__ uint8_t smem_raw_u[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_u) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (16384 + 32768) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { sg::mbar_init(&bar, 1); sg::fence_barrier_init(); }
  sg::fence_proxy_async_smem();
  if (threadIdx.x < 32) sg::tmem_alloc(&slot, 512);
  sg::tc_fence_before();
  __syncthreads();
  sg::tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t idesc = sg::make_idesc_bf16(128, N, 0, 0);
    const uint64_t a_desc = sg::make_smem_desc(sg::smem_u32(smem), 16, 1024, sg::kLayoutSW128);
    const uint64_t b_desc = sg::make_smem_desc(sg::smem_u32(smem + 16384), 16, 1024, sg::kLayoutSW128);
    // accumulators at columns 0, N (2 chains) or 0, 128, 256, 384 (4 chains, N <= 128); chains is a power of two
    const uint32_t stride = (chains == 4) ? 128u : static_cast<uint32_t>(N);
    const uint32_t mask = static_cast<uint32_t>(chains - 1);
    const long long t0 = clock64();
#pragma unroll 8
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = slot + (static_cast<uint32_t>(it) & mask) * stride;
      const uint64_t koff = 2 * (it & 3);
      if (use_ts) sg::umma_ts(d, slot + 448, b_desc + koff, idesc, 1);
      else sg::umma_ss(d, a_desc + koff, b_desc + koff, idesc, 1);
    }
    sg::umma_commit(&bar);
    sg::mbar_wait(&bar, 0);
    const long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) sg::tmem_dealloc(slot, 512);
}
template <int N>
static void run_mma_bench(long long* cyc) {
  CK(cudaFuncSetAttribute(ubench_mma<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024));
  for (int ts = 0; ts < 2; ++ts)
    for (int chains : {1, 2, 4}) {
      if ((chains == 2 && 2 * N > 448) || (chains == 4 && N > 112)) continue;  // keep clear of the TS operand at col 448
      const int iters = 512;
      ubench_mma<N><<<148, 128, 60 * 1024>>>(iters, chains, ts, cyc);
      CK(cudaDeviceSynchronize());
      long long h;
      CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
      printf("[ubench] tcgen05.mma M128 N%-3d K16 %s: %d accumulator chain(s): %.1f clk per MMA (math %.0f clk)\n", N,
             ts ? "TS" : "SS", chains, static_cast<double>(h) / iters, 128.0 * N / 256.0);
    }
}
void test_ubench_mma() {
  if (!want("ubench")) return;
  long long* cyc;
  CK(cudaMalloc(&cyc, 8 * 1024));
  run_mma_bench<64>(cyc);
  run_mma_bench<128>(cyc);
  run_mma_bench<160>(cyc);
  run_mma_bench<256>(cyc);
  cudaFree(cyc);
}

// does the ~130 clk per-instruction floor belong to the issuing thread or to the tensor pipe?  `issuers` warps of one
// CTA (lane 0 of each) issue independent MMA streams into their own accumulators; optionally 2 CTAs per SM.
template <int M, int N>
__global__ void ubench_mma_multi(int iters, int issuers, int use_ts, long long* cycles) {
  extern __shared__ uint8_t smem_raw_u[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_u) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t bar[4];
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (16384 + 32768) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) sg::mbar_init(&bar[i], 1); sg::fence_barrier_init(); }
  sg::fence_proxy_async_smem();
  if (threadIdx.x < 32) sg::tmem_alloc(&slot, 256);
  sg::tc_fence_before();
  __syncthreads();
  sg::tc_fence_after();
  const int warp = threadIdx.x >> 5;
  long long t0 = 0, t1 = 0;
  if (warp < issuers && (threadIdx.x & 31) == 0) {
    const uint32_t idesc = sg::make_idesc_bf16(M, N, 0, 0);
    const uint64_t a_desc = sg::make_smem_desc(sg::smem_u32(smem), 16, 1024, sg::kLayoutSW128);
    const uint64_t b_desc = sg::make_smem_desc(sg::smem_u32(smem + 16384), 16, 1024, sg::kLayoutSW128);
    const uint32_t d = slot + static_cast<uint32_t>(warp) * 64u;  // N <= 64 per issuer when issuers > 1
    t0 = clock64();
#pragma unroll 8
    for (int it = 0; it < iters; ++it) {
      const uint64_t koff = 2 * (it & 3);
      if (use_ts) sg::umma_ts(d, slot + 192 + 16 * warp, b_desc + koff, idesc, 1);
      else sg::umma_ss(d, a_desc + koff, b_desc + koff, idesc, 1);
    }
    sg::umma_commit(&bar[warp]);
    sg::mbar_wait(&bar[warp], 0);
    t1 = clock64();
    if (warp == 0) cycles[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) sg::tmem_dealloc(slot, 256);
}
template <int M, int N>
static void run_mma_multi(long long* cyc, int issuers, int ctas_per_sm, int ts) {
  CK(cudaFuncSetAttribute(ubench_mma_multi<M, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024));
  const int iters = 512;
  ubench_mma_multi<M, N><<<148 * ctas_per_sm, 128, 60 * 1024>>>(iters, issuers, ts, cyc);
  CK(cudaDeviceSynchronize());
  long long h;
  CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
  printf("[ubench] tcgen05.mma M%-3d N%-3d K16 %s: %d issuing warp(s) x %d CTA/SM: %.1f clk per MMA per issuer -> %.1f clk per MMA per SM\n",
         M, N, ts ? "TS" : "SS", issuers, ctas_per_sm, static_cast<double>(h) / iters,
         static_cast<double>(h) / iters / (issuers * ctas_per_sm));
}
void test_ubench_mma_multi() {
  if (!want("ubench")) return;
  long long* cyc;
  CK(cudaMalloc(&cyc, 8 * 1024));
  for (int ts = 0; ts < 2; ++ts) {
    run_mma_multi<128, 64>(cyc, 1, 1, ts);
    run_mma_multi<128, 64>(cyc, 2, 1, ts);
    run_mma_multi<128, 64>(cyc, 4, 1, ts);
    run_mma_multi<128, 64>(cyc, 1, 2, ts);
    run_mma_multi<128, 64>(cyc, 2, 2, ts);
    run_mma_multi<128, 16>(cyc, 1, 1, ts);
    run_mma_multi<128, 32>(cyc, 1, 1, ts);
    run_mma_multi<64, 64>(cyc, 1, 1, ts);
    run_mma_multi<64, 64>(cyc, 2, 1, ts);
    run_mma_multi<64, 256>(cyc, 1, 1, ts);
    run_mma_multi<128, 256>(cyc, 1, 1, ts);
    run_mma_multi<128, 128>(cyc, 1, 2, ts);
  }
  cudaFree(cyc);
}

// what slows an MMA issue stream down?  V: 0 SS plain | 1 TS with the A address hoisted | 2 SS, accumulator address
// re-read from shared memory every MMA | 3 SS + mbarrier.test_wait every 4 MMAs | 4 SS + tcgen05.commit every 4 MMAs
// | 5 = 3 + 4 | 6 TS hoisted + commit every 8 + test_wait every 8
template <int N, int V>
__global__ void ubench_mma_var(int iters, long long* cycles) {
  extern __shared__ uint8_t smem_raw_u[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_u) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t bar, bar2, bar3;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (16384 + 32768) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { sg::mbar_init(&bar, 1); sg::mbar_init(&bar2, 1); sg::mbar_init(&bar3, 1); sg::fence_barrier_init(); }
  sg::fence_proxy_async_smem();
  if (threadIdx.x < 32) sg::tmem_alloc(&slot, 512);
  sg::tc_fence_before();
  __syncthreads();
  sg::tc_fence_after();
  if (threadIdx.x == 0) {
    sg::mbar_arrive(&bar3);  // phase 0 of bar3 is complete: test_wait(bar3, 0) returns true at once
    const uint32_t idesc = sg::make_idesc_bf16(128, N, 0, 0);
    const uint64_t a_desc = sg::make_smem_desc(sg::smem_u32(smem), 16, 1024, sg::kLayoutSW128);
    const uint64_t b_desc = sg::make_smem_desc(sg::smem_u32(smem + 16384), 16, 1024, sg::kLayoutSW128);
    const uint32_t d0 = slot, a_t = slot + 256;
    uint32_t sink = 0;
    const long long t0 = clock64();
#pragma unroll 8
    for (int it = 0; it < iters; ++it) {
      const uint64_t koff = 2 * (it & 3);
      if (V == 1 || V == 6) sg::umma_ts(d0, a_t, b_desc + koff, idesc, 1);
      else if (V == 2) sg::umma_ss(*reinterpret_cast<volatile uint32_t*>(&slot), a_desc + koff, b_desc + koff, idesc, 1);
      else sg::umma_ss(d0, a_desc + koff, b_desc + koff, idesc, 1);
      if ((V == 3 || V == 5) && (it & 3) == 3) sink += sg::mbar_test(&bar3, 0) ? 1u : 0u;
      if ((V == 4 || V == 5) && (it & 3) == 3) sg::umma_commit(&bar2);
      if (V == 6 && (it & 7) == 7) { sg::umma_commit(&bar2); sink += sg::mbar_test(&bar3, 0) ? 1u : 0u; }
    }
    sg::umma_commit(&bar);
    sg::mbar_wait(&bar, 0);
    const long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0 + (sink == 0xffffffffu ? 1 : 0);
  }
  __syncthreads();
  if (threadIdx.x < 32) sg::tmem_dealloc(slot, 512);
}
template <int N, int V>
static void run_mma_var(long long* cyc, const char* what) {
  CK(cudaFuncSetAttribute(ubench_mma_var<N, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024));
  const int iters = 512;
  ubench_mma_var<N, V><<<148, 128, 60 * 1024>>>(iters, cyc);
  CK(cudaDeviceSynchronize());
  long long h;
  CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
  printf("[ubench] issue-stream N%-3d %-52s: %.1f clk per MMA (math %.0f)\n", N, what, static_cast<double>(h) / iters, N / 2.0);
}
template <int N>
static void run_mma_vars(long long* cyc) {
  run_mma_var<N, 0>(cyc, "SS plain");
  run_mma_var<N, 1>(cyc, "TS, A address hoisted");
  run_mma_var<N, 2>(cyc, "SS, D address re-read from smem per MMA");
  run_mma_var<N, 3>(cyc, "SS + mbarrier.test_wait every 4");
  run_mma_var<N, 4>(cyc, "SS + tcgen05.commit every 4");
  run_mma_var<N, 5>(cyc, "SS + commit + test_wait every 4");
  run_mma_var<N, 6>(cyc, "TS hoisted + commit + test_wait every 8");
}
void test_ubench_mma_var() {
  if (!want("ubench2")) return;
  long long* cyc;
  CK(cudaMalloc(&cyc, 8 * 1024));
  run_mma_vars<64>(cyc);
  run_mma_vars<128>(cyc);
  run_mma_vars<160>(cyc);
  run_mma_vars<256>(cyc);
  cudaFree(cyc);
}

// how many bytes per clock can one SM pull into shared memory with TMA?  A ring of 4 x 32 KB stages ([256 rows x 64
// bf16] boxes of a row-major matrix), a consumer that frees a stage as soon as it lands, no tensor work at all.
__global__ void ubench_tma_ingest(const __grid_constant__ CUtensorMap map, int rows_total, int iters, int spread,
                                  long long* cycles) {
  extern __shared__ uint8_t smem_raw_u[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_u) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t full[4], empty[4];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) { sg::mbar_init(&full[i], 1); sg::mbar_init(&empty[i], 1); }
    sg::fence_barrier_init();
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    const long long t0 = clock64();
    // spread = 1: every CTA walks its own region (HBM stream); 0: all CTAs walk the same 8 MB (L2 hits)
    int row = spread ? static_cast<int>((static_cast<long long>(blockIdx.x) * 8192) % rows_total) : (blockIdx.x * 256) % 8192;
    for (int it = 0; it < iters; ++it) {
      const int st = it & 3;
      sg::mbar_wait(&empty[st], ((it >> 2) & 1) ^ 1);
      sg::mbar_arrive_expect_tx(&full[st], 32768);
      sg::tma_load_2d(smem + st * 32768, &map, &full[st], 0, row);
      row += 256;
      if (spread) { if (row >= rows_total) row = 0; } else if (row >= 65536) row = 0;
    }
    sg::mbar_wait(&full[(iters - 1) & 3], ((iters - 1) >> 2) & 1);  // last stage has landed
    cycles[blockIdx.x] = clock64() - t0;
  } else if (warp == 1 && lane == 0) {
    for (int it = 0; it < iters; ++it) {
      const int st = it & 3;
      sg::mbar_wait(&full[st], (it >> 2) & 1);
      sg::mbar_arrive(&empty[st]);
    }
  }
}
void test_ubench_tma() {
  if (!want("ubench3")) return;
  const int rows_total = 1 << 21;  // x 128 B = 256 MB
  void* buf;
  CK(cudaMalloc(&buf, static_cast<size_t>(rows_total) * 128));
  CK(cudaMemset(buf, 0, static_cast<size_t>(rows_total) * 128));
  CUtensorMap map;
  uint64_t dims[2] = {64, static_cast<uint64_t>(rows_total)};
  uint64_t strides[1] = {128};
  uint32_t box[2] = {64, 256};
  if (sg::encode_tmap_bf16(&map, buf, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B) != 0) { printf("tmap failed\n"); return; }
  long long* cyc;
  CK(cudaMalloc(&cyc, 8 * 1024));
  CK(cudaFuncSetAttribute(ubench_tma_ingest, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 32768 + 1024));
  const int iters = 2048;
  for (int spread = 0; spread < 2; ++spread)
    for (int ctas : {1, 16, 74, 148}) {
      ubench_tma_ingest<<<ctas, 64, 4 * 32768 + 1024>>>(map, rows_total, iters, spread, cyc);
      CK(cudaDeviceSynchronize());
      long long h[148];
      CK(cudaMemcpy(h, cyc, 8 * ctas, cudaMemcpyDeviceToHost));
      double mx = 0;
      for (int i = 0; i < ctas; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("[ubench] TMA ingest, %3d CTAs (1/SM), %s: %.1f B/clk/SM (slowest CTA), aggregate %.2f TB/s @1.96 GHz\n", ctas,
             spread ? "distinct regions (HBM)" : "shared 8 MB (L2)      ", 32768.0 * iters / mx,
             32768.0 * iters / mx * ctas * 1.96e9 / 1e12);
    }
  cudaFree(cyc);
  cudaFree(buf);
}

// softmax inner loop candidates: exp2 (+row-sum FADD) alone / + cvt.rn.bf16x2.f32 per pair / + integer
// round-half-up pack (2 IADD + PRMT per pair) / + exp2 replaced by a degree-3 polynomial on the FMA pipe for 1 in 4
template <int V>
__global__ void ubench_softmax_mix(int iters, long long* cycles, float* sink) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = -0.001f * (threadIdx.x + i);
  float sum = 0.f;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float e[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float a = fmaf(x[i], 1.0009f, -0.25f);
      if (V == 3 && (i & 3) == 3) {
        // 2^a for a <= 0: Cody-Waite split + cubic on the fractional part (FMA pipe only)
        const float fl = floorf(a);
        const float f = a - fl;
        float pz = fmaf(f, 0.0555054f, 0.2402265f);
        pz = fmaf(pz, f, 0.6931472f);
        pz = fmaf(pz, f, 1.0f);
        e[i] = __uint_as_float(__float_as_uint(pz) + (static_cast<int>(fl) << 23));
      } else {
        e[i] = sg::fast_exp2(a);
      }
      sum += e[i];
    }
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      if (V == 1) acc ^= sg::pack_bf16x2(e[i], e[i + 1]);
      if (V == 2 || V == 3) {
        const uint32_t lo = __float_as_uint(e[i]) + 0x8000u, hi = __float_as_uint(e[i + 1]) + 0x8000u;
        acc ^= __byte_perm(lo, hi, 0x7632);
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = e[i] - 1.5f;
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  float a = sum + __uint_as_float(acc & 0x3fffffffu);
#pragma unroll
  for (int i = 0; i < 16; ++i) a += x[i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = a;
}
template <int V>
static void run_softmax_mix(long long* cyc, float* sink, const char* what) {
  for (int warps : {4, 8}) {
    const int iters = 4000;
    ubench_softmax_mix<V><<<148, warps * 32>>>(iters, cyc, sink);
    CK(cudaDeviceSynchronize());
    long long h;
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("[ubench] softmax mix %-46s %d warps/SM: %.2f elements/clk/SM\n", what, warps,
           static_cast<double>(iters) * 16 * warps * 32 / h);
  }
}
void test_ubench_softmax_mix() {
  if (!want("ubench4")) return;
  long long* cyc;
  float* sink;
  CK(cudaMalloc(&cyc, 8 * 1024));
  CK(cudaMalloc(&sink, 4 * 1024 * 1024));
  run_softmax_mix<0>(cyc, sink, "FFMA + ex2 + FADD");
  run_softmax_mix<1>(cyc, sink, "... + cvt.rn.bf16x2.f32 per pair");
  run_softmax_mix<2>(cyc, sink, "... + integer pack (2 IADD + PRMT) per pair");
  run_softmax_mix<3>(cyc, sink, "... integer pack, 1 in 4 exps as FMA polynomial");
  cudaFree(cyc);
  cudaFree(sink);
}

// TS-mode GEMM feasibility: per k-block 4 x tcgen05.cp (smem A [128 x 16] -> TMEM, 128x256b) + 4 x TS MMA
template <int N>
__global__ void ubench_cp_ts(int iters, int with_cp, int with_mma, long long* cycles) {
  extern __shared__ uint8_t smem_raw_u2[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_u2) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (16384 + 32768) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { sg::mbar_init(&bar, 1); sg::fence_barrier_init(); }
  sg::fence_proxy_async_smem();
  if (threadIdx.x < 32) sg::tmem_alloc(&slot, 512);
  sg::tc_fence_before();
  __syncthreads();
  sg::tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t idesc = sg::make_idesc_bf16(128, N, 0, 0);
    const uint64_t a_desc = sg::make_smem_desc(sg::smem_u32(smem), 16, 1024, sg::kLayoutSW128);
    const uint64_t b_desc = sg::make_smem_desc(sg::smem_u32(smem + 16384), 16, 1024, sg::kLayoutSW128);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t a_tmem = slot + 256 + (it & 1) * 32;  // double-buffered A region: 8 columns per k16 step
      if (with_cp) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(a_tmem + k * 8), "l"(a_desc + 2 * k) : "memory");
      }
      if (with_mma) {
#pragma unroll
        for (int k = 0; k < 4; ++k) sg::umma_ts(slot, a_tmem + k * 8, b_desc + 2 * k, idesc, 1);
      }
    }
    sg::umma_commit(&bar);
    sg::mbar_wait(&bar, 0);
    cycles[blockIdx.x] = clock64() - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) sg::tmem_dealloc(slot, 512);
}
void test_ubench_cp() {
  if (!want("ubench")) return;
  long long* cyc;
  CK(cudaMalloc(&cyc, 8 * 1024));
  CK(cudaFuncSetAttribute(ubench_cp_ts<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024));
  CK(cudaFuncSetAttribute(ubench_cp_ts<160>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024));
  const int iters = 256;
  for (int mode = 0; mode < 3; ++mode) {
    const int cp = (mode != 1), mma = (mode != 0);
    long long h;
    ubench_cp_ts<256><<<148, 128, 60 * 1024>>>(iters, cp, mma, cyc);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("[ubench] k-block (K=64) N=256: %s%s: %.1f clk per k-block (SS would be 774, math 512)\n", cp ? "4x tcgen05.cp " : "",
           mma ? "4x TS mma" : "", static_cast<double>(h) / iters);
    ubench_cp_ts<160><<<148, 128, 60 * 1024>>>(iters, cp, mma, cyc);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("[ubench] k-block (K=64) N=160: %s%s: %.1f clk per k-block (SS would be 668, math 320)\n", cp ? "4x tcgen05.cp " : "",
           mma ? "4x TS mma" : "", static_cast<double>(h) / iters);
  }
  cudaFree(cyc);
}
