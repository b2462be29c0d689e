// Standalone device self-test of the C ABI (no Python, no torch): every case builds random bf16
// inputs, runs the launcher, and compares with a straightforward fp32 host loop over the SAME
// bf16-rounded inputs.  Used for bring-up under gpurun; the pytest suite covers the same kernels
// through ctypes against the oracle.
//   ./selftest [filter-substring]
#include "selftest_util.h"

static float gelu_h(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678f)); }

// ---------------------------------------------------------------------------------------------
// GEMM: generic host reference driven by the same sg_gemm_args
// ---------------------------------------------------------------------------------------------
struct SrcHost {
  const HostBf* t;
  sg_src s;
};

static void run_gemm_case(const char* name, std::vector<SrcHost>& srcs, std::vector<sg_seg>& segs, int out_w,
                          int out_h, int out_n, int bw, int bh, int bn, int N, int mode, bool use_bias,
                          bool use_lane_bias, bool use_res, float res_scale) {
  if (!want(name)) return;
  int Ktot = 0;
  for (auto& s : segs) Ktot += s.nkb * 64;
  const long long M = 1ll * out_w * out_h * out_n;
  const int n_out_cols = (mode == SG_EPI_GEGLU) ? N / 2 : N;
  HostBf B, R, O;
  HostF bias, lbias;
  B.init(static_cast<size_t>(N) * Ktot, 1.0f / sqrtf(static_cast<float>(Ktot)) * 2.0f);
  bias.init(N, 0.5f);
  lbias.init(static_cast<size_t>(out_n) * N, 0.5f);
  R.init(static_cast<size_t>(M) * n_out_cols, 1.0f);
  O.zeros(static_cast<size_t>(M) * n_out_cols);

  sg_gemm_args a;
  memset(&a, 0, sizeof(a));
  a.num_src = static_cast<int>(srcs.size());
  for (size_t i = 0; i < srcs.size(); ++i) a.src[i] = srcs[i].s;
  a.num_segs = static_cast<int>(segs.size());
  for (size_t i = 0; i < segs.size(); ++i) a.segs[i] = segs[i];
  a.out_w = out_w;
  a.out_h = out_h;
  a.out_n = out_n;
  a.bw = bw;
  a.bh = bh;
  a.bn = bn;
  a.b = B.d;
  a.N = N;
  a.Ktot = Ktot;
  a.mode = mode;
  a.res_scale = res_scale;
  a.bias = use_bias ? bias.d : nullptr;
  a.lane_bias = use_lane_bias ? lbias.d : nullptr;
  a.residual = use_res ? R.d : nullptr;
  a.ldr = n_out_cols;
  a.out = O.d;
  a.ldo = n_out_cols;
  int rc = sg_gemm(&a, nullptr);
  CK(cudaDeviceSynchronize());
  O.fetch();

  // host reference
  double max_err = 0, max_ref = 0;
  std::vector<float> acc(N);
  for (int n = 0; n < out_n; ++n)
    for (int y = 0; y < out_h; ++y)
      for (int x = 0; x < out_w; ++x) {
        const long long m = (1ll * n * out_h + y) * out_w + x;
        for (int c = 0; c < N; ++c) acc[c] = 0.f;
        int kbase = 0;
        for (auto& sgm : segs) {
          const SrcHost& sh = srcs[sgm.src];
          const int xx = x + sgm.dx, yy = y + sgm.dy;
          const bool inb = xx >= 0 && xx < sh.s.dim_w && yy >= 0 && yy < sh.s.dim_h && n < sh.s.dim_n;
          if (inb) {
            // offset of sh.s.ptr relative to the tensor base, in elements
            const long long base_off =
                (reinterpret_cast<const char*>(sh.s.ptr) - reinterpret_cast<const char*>(sh.t->d)) / 2;
            const float* ap = sh.t->f.data() + base_off + 1ll * n * sh.s.stride_n +
                              1ll * yy * sh.s.stride_h + 1ll * xx * sh.s.stride_w + sgm.c0;
            for (int c = 0; c < N; ++c) {
              const float* bp = B.f.data() + 1ll * c * Ktot + kbase;
              float s = 0.f;
              for (int k = 0; k < sgm.nkb * 64; ++k) s += ap[k] * bp[k];
              acc[c] += s;
            }
          }
          kbase += sgm.nkb * 64;
        }
        for (int c = 0; c < N; ++c) {
          if (use_bias) acc[c] += bias.f[c];
          if (use_lane_bias) acc[c] += lbias.f[1ll * n * N + c];
        }
        if (mode == SG_EPI_LINEAR) {
          for (int c = 0; c < N; ++c) {
            float ref = acc[c] + (use_res ? res_scale * R.f[m * N + c] : 0.f);
            float got = O.f[m * N + c];
            max_err = fmax(max_err, fabs(static_cast<double>(ref) - got));
            max_ref = fmax(max_ref, fabs(static_cast<double>(ref)));
          }
        } else {
          for (int c32 = 0; c32 < N; c32 += 32)
            for (int j = 0; j < 16; ++j) {
              float ref = acc[c32 + j] * gelu_h(acc[c32 + 16 + j]);
              float got = O.f[m * n_out_cols + c32 / 2 + j];
              max_err = fmax(max_err, fabs(static_cast<double>(ref) - got));
              max_ref = fmax(max_ref, fabs(static_cast<double>(ref)));
            }
        }
      }
  report(name, rc, max_err, max_ref, 0.02 * fmax(1.0, max_ref));
}

static sg_src plain_src(const HostBf& t, long long M, int K) {
  sg_src s;
  s.ptr = t.d;
  s.dim_c = K;
  s.dim_w = M;
  s.dim_h = 1;
  s.dim_n = 1;
  s.stride_w = K;
  s.stride_h = M * K;
  s.stride_n = M * K;
  return s;
}
static sg_src nhwc_src(const HostBf& t, int n, int h, int w, int c) {
  sg_src s;
  s.ptr = t.d;
  s.dim_c = c;
  s.dim_w = w;
  s.dim_h = h;
  s.dim_n = n;
  s.stride_w = c;
  s.stride_h = 1ll * w * c;
  s.stride_n = 1ll * h * w * c;
  return s;
}

static void test_gemm() {
  {  // plain, single tile, BN=64
    HostBf A;
    A.init(128 * 64, 1.f);
    std::vector<SrcHost> srcs = {{&A, plain_src(A, 128, 64)}};
    std::vector<sg_seg> segs = {{0, 0, 0, 0, 1}};
    run_gemm_case("gemm plain 128x64x64 (1 k-block)", srcs, segs, 128, 1, 1, 128, 1, 1, 64, SG_EPI_LINEAR,
                  false, false, false, 0.f);
  }
  {  // plain, multiple k blocks + M tail + bias + residual, BN=160
    HostBf A;
    A.init(300 * 320, 1.f);
    std::vector<SrcHost> srcs = {{&A, plain_src(A, 300, 320)}};
    std::vector<sg_seg> segs = {{0, 0, 0, 0, 5}};
    run_gemm_case("gemm plain 300x320x320 bias+res (BN160)", srcs, segs, 300, 1, 1, 128, 1, 1, 320,
                  SG_EPI_LINEAR, true, false, true, 2.0f);
  }
  {  // BN=256, many tiles (persistent loop with > 148 tiles), K=1280
    HostBf A;
    A.init(4096ll * 640, 1.f);
    std::vector<SrcHost> srcs = {{&A, plain_src(A, 4096, 640)}};
    std::vector<sg_seg> segs = {{0, 0, 0, 0, 10}};
    run_gemm_case("gemm plain 4096x1280x640 (BN256, 160 tiles)", srcs, segs, 4096, 1, 1, 128, 1, 1, 1280,
                  SG_EPI_LINEAR, true, false, false, 0.f);
  }
  {  // BN=128
    HostBf A;
    A.init(512ll * 128, 1.f);
    std::vector<SrcHost> srcs = {{&A, plain_src(A, 512, 128)}};
    std::vector<sg_seg> segs = {{0, 0, 0, 0, 2}};
    run_gemm_case("gemm plain 512x128x128 (BN128)", srcs, segs, 512, 1, 1, 128, 1, 1, 128, SG_EPI_LINEAR,
                  false, false, true, 1.0f);
  }
  {  // GEGLU epilogue
    HostBf A;
    A.init(256ll * 320, 1.f);
    std::vector<SrcHost> srcs = {{&A, plain_src(A, 256, 320)}};
    std::vector<sg_seg> segs = {{0, 0, 0, 0, 5}};
    run_gemm_case("gemm geglu 256x2560x320", srcs, segs, 256, 1, 1, 128, 1, 1, 2560, SG_EPI_GEGLU, true,
                  false, false, 0.f);
  }
  {  // two-source K concat (dual out-projection: [O_t | O_i] [W_t | W_i]^T + 2h)
    HostBf A0, A1;
    A0.init(256ll * 320, 1.f);
    A1.init(256ll * 320, 1.f);
    std::vector<SrcHost> srcs = {{&A0, plain_src(A0, 256, 320)}, {&A1, plain_src(A1, 256, 320)}};
    std::vector<sg_seg> segs = {{0, 0, 0, 0, 5}, {1, 0, 0, 0, 5}};
    run_gemm_case("gemm dual-source 256x320x(320+320) +2h", srcs, segs, 256, 1, 1, 128, 1, 1, 320,
                  SG_EPI_LINEAR, true, false, true, 2.0f);
  }
  {  // 3x3 conv 16x16, 64->64, n=3 (odd: tile overhang in n), time-embedding lane bias
    HostBf X;
    X.init(3ll * 16 * 16 * 64, 1.f);
    std::vector<SrcHost> srcs = {{&X, nhwc_src(X, 3, 16, 16, 64)}};
    std::vector<sg_seg> segs;
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) segs.push_back({0, kx - 1, ky - 1, 0, 1});
    run_gemm_case("conv3x3 n3 16x16 64->64 +temb", srcs, segs, 16, 16, 3, 16, 8, 1, 64, SG_EPI_LINEAR, true,
                  true, false, 0.f);
  }
  {  // 3x3 conv 8x8 with bn=2 boxes, 128->320
    HostBf X;
    X.init(3ll * 8 * 8 * 128, 1.f);
    std::vector<SrcHost> srcs = {{&X, nhwc_src(X, 3, 8, 8, 128)}};
    std::vector<sg_seg> segs;
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) segs.push_back({0, kx - 1, ky - 1, 0, 2});
    run_gemm_case("conv3x3 n3 8x8 128->320 (bn=2 box)", srcs, segs, 8, 8, 3, 8, 8, 2, 320, SG_EPI_LINEAR,
                  true, false, false, 0.f);
  }
  {  // 3x3 conv 64x64 (bw=64,bh=2), 64->64, + 1x1 shortcut over a concat of two raw sources
    HostBf X, S0, S1;
    X.init(1ll * 64 * 64 * 64, 1.f);
    S0.init(1ll * 64 * 64 * 128, 1.f);
    S1.init(1ll * 64 * 64 * 64, 1.f);
    std::vector<SrcHost> srcs = {{&X, nhwc_src(X, 1, 64, 64, 64)},
                                 {&S0, nhwc_src(S0, 1, 64, 64, 128)},
                                 {&S1, nhwc_src(S1, 1, 64, 64, 64)}};
    std::vector<sg_seg> segs;
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) segs.push_back({0, kx - 1, ky - 1, 0, 1});
    segs.push_back({1, 0, 0, 0, 2});
    segs.push_back({2, 0, 0, 0, 1});
    run_gemm_case("conv3x3 64x64 64->64 + shortcut(cat 128|64)", srcs, segs, 64, 64, 1, 64, 2, 1, 64,
                  SG_EPI_LINEAR, true, false, false, 0.f);
  }
  {  // stride-2 3x3 conv 16x16 -> 8x8 via 4 phase-strided views of the input
    HostBf X;
    const int n = 2, h = 16, w = 16, c = 64;
    X.init(1ll * n * h * w * c, 1.f);
    std::vector<SrcHost> srcs;
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        sg_src s;
        s.ptr = X.d + (1ll * py * w + px) * c;
        s.dim_c = c;
        s.dim_w = w / 2;
        s.dim_h = h / 2;
        s.dim_n = n;
        s.stride_w = 2ll * c;
        s.stride_h = 2ll * w * c;
        s.stride_n = 1ll * h * w * c;
        srcs.push_back({&X, s});
      }
    // input row = 2*oy + ky - 1: ky=0 -> phase 1, shift -1; ky=1 -> phase 0, shift 0; ky=2 -> phase 1, shift 0
    const int ph[3] = {1, 0, 1};
    const int sh[3] = {-1, 0, 0};
    std::vector<sg_seg> segs;
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) segs.push_back({ph[ky] * 2 + ph[kx], sh[kx], sh[ky], 0, 1});
    run_gemm_case("conv3x3 stride2 16x16->8x8 64->64 (phase views)", srcs, segs, 8, 8, n, 8, 8, 2, 64,
                  SG_EPI_LINEAR, true, false, false, 0.f);
  }
}

#ifdef SG_SELFTEST_EXTRA
void test_extra();
#endif

int main(int argc, char** argv) {
  if (argc > 1) g_filter = argv[1];
  int dev_count = 0;
  CK(cudaGetDeviceCount(&dev_count));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s sm_%d%d, %d SMs, abi %d\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount,
         sg_abi_version());
  test_gemm();
#ifdef SG_SELFTEST_EXTRA
  test_extra();
#endif
  printf("%s: %d failure(s), %lld launches\n", g_fail ? "SELFTEST FAILED" : "SELFTEST OK", g_fail,
         static_cast<long long>(sg_launch_count()));
  return g_fail ? 1 : 0;
}
