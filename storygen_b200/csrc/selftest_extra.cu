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

void test_extra() {
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
