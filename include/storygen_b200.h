/* storygen_b200 — C ABI of the B200 (sm_100a) kernels behind StoryGen's denoising hot path.
 *
 * The reference (haoningwu3639/StoryGen) ships no native code and no FFI: its hot path is a chain of
 * torch.nn / diffusers / xformers library calls.  Each entry point below replaces one group of those
 * calls; the comment on each names the reference call sites (file:line in /root/reference) it stands
 * in for.  INTEGRATION.md shows the ctypes binding used by storygen_b200/ (the Python mirror of the
 * reference's UNet2DConditionModel / StableDiffusionPipeline).
 *
 * Conventions
 *   - All pointers are DEVICE pointers unless stated otherwise; the caller owns every buffer
 *     (inputs, outputs, workspaces).  The library never allocates or frees device memory and keeps
 *     no mutable global state besides the lazily resolved driver entry point for tensor-map encoding.
 *   - Activations are bf16, channel-last: images are (n, y, x, c) = "NHWC", token streams are
 *     (n, token, c).  Statistics / biases / time embeddings are fp32.
 *   - Every launcher enqueues on `stream` (a cudaStream_t passed as void*), never synchronises, and
 *     is CUDA-graph capturable.
 *   - Return value: 0 on success; negative = invalid argument (SG_ERR_*); positive = cudaError_t of
 *     a failed launch/encode.  Nothing throws or aborts.
 */
#ifndef STORYGEN_B200_H
#define STORYGEN_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SG_ABI_VERSION 9

#define SG_ERR_SHAPE (-1)   /* unsupported / inconsistent shape            */
#define SG_ERR_ALIGN (-2)   /* pointer or stride not 16-byte aligned        */
#define SG_ERR_DRIVER (-3)  /* cuTensorMapEncodeTiled entry point not found */
#define SG_ERR_ARG (-4)     /* null pointer / bad enum                      */

/* ---- introspection ------------------------------------------------------------------------- */
int sg_abi_version(void);
/* sizeof() of the argument structs as compiled, so a binding can verify its mirror. */
int sg_sizeof_gemm_args(void);
int sg_sizeof_attn_args(void);
/* number of kernels this library has launched since load (host counter; bench "gpu_launches") */
int64_t sg_launch_count(void);
const char* sg_error_string(int code);

/* ---- tcgen05 GEMM / implicit-GEMM convolution ------------------------------------------------
 * out[m, :] = epilogue( sum_k A[m, k] * B[:, k] ),  A assembled on the fly from up to 4 NHWC source
 * tensors by TMA: the K loop walks `segs`; segment s contributes nkb*64 consecutive channels
 * [c0, c0 + 64*nkb) of source `src`, read at pixel (x + dx, y + dy) of the output pixel (x, y)
 * (out-of-image reads return 0 = zero padding).  A plain Linear is one segment with dx = dy = 0 over
 * a (1, 1, M, K) "image"; a 3x3 conv is 9 segments; conv + 1x1 shortcut + channel concat are just
 * more segments.  B is the pre-packed weight [N, Ktot] (bf16, K contiguous, K ordered as `segs`).
 *
 * Replaces: nn.Conv2d 3x3 / 1x1 and nn.Linear calls of ResnetBlock2D, Downsample2D, Upsample2D
 * (model/unet_2d_blocks.py:222-233,332-368,552-584), Transformer2DModel.proj_in/proj_out
 * (model/attention.py:59,83,100,118), CrossAttention.to_q/to_k/to_v/to_out (model/attention.py:
 * 175-223 via diffusers), FeedForward/GEGLU (model/attention.py:347-350,391-393), the skip
 * torch.cat (model/unet_2d_blocks.py:609,626,716) and the residual adds (model/attention.py:262,
 * 277,291-293,300; ResnetBlock2D output).
 */
typedef struct sg_src {
  const void* ptr;                    /* bf16, 16-byte aligned                              */
  int64_t dim_c, dim_w, dim_h, dim_n; /* extents (elements)                                 */
  int64_t stride_w, stride_h, stride_n; /* element strides (channel stride is 1); *2 % 16 == 0 */
} sg_src;

typedef struct sg_seg {
  int32_t src; /* index into sg_gemm_args.src */
  int32_t dx, dy;
  int32_t c0;  /* first channel                */
  int32_t nkb; /* number of 64-channel blocks  */
} sg_seg;

#define SG_MAX_SRC 4
#define SG_MAX_SEGS 24

#define SG_EPI_LINEAR 0 /* out = acc + bias + lane_bias + res_scale * residual            */
#define SG_EPI_GEGLU 1  /* B rows interleaved [16 value | 16 gate]; out = v * gelu_erf(g) */
#define SG_EPI_NCHW_F32 2 /* N == 32 (zero-padded weight rows); out = fp32 (n, out_channels, h, w) */

typedef struct sg_gemm_args {
  int32_t num_src;
  int32_t num_segs;
  sg_src src[SG_MAX_SRC];
  sg_seg segs[SG_MAX_SEGS];
  /* output pixel grid and the 128-row tile box (bw * bh * bn == 128) */
  int32_t out_w, out_h, out_n;
  int32_t bw, bh, bn;
  /* weights */
  const void* b; /* bf16 [N, Ktot] */
  int64_t N, Ktot;
  /* epilogue */
  int32_t mode;
  int32_t out_channels; /* SG_EPI_NCHW_F32 only */
  float res_scale;
  const float* bias;      /* [N] (pre-activation columns) or NULL */
  const float* lane_bias; /* [out_n, ld_lane_bias] or NULL (per-image bias: time embedding) */
  int64_t ld_lane_bias;
  const void* residual;   /* bf16 [M, ldr] or NULL (M = out_n*out_h*out_w rows)           */
  int64_t ldr;
  void* out; /* bf16 [M, ldo] */
  int64_t ldo;
  int64_t* debug_timeline; /* bring-up aid: device int64[10*8*8] phase clocks of CTA 0, or NULL */
  /* ---- fused normalisation hooks (all optional) ----------------------------------------------
   * LayerNorm is folded INTO the GEMMs either side of it (model/attention.py:250,268,283,298 `norm1..norm4` followed by
   * to_q / to_k / to_v / ff.net[0]):   LN(x) W^T + b = rstd_m (x W'^T - mean_m csum) + b',
   *   W' = W diag(gamma) (bf16),  csum[n] = sum_k W'[n,k],  b' = b + W beta   (packed once per model),
   * so the normalised activation is never written: the PRODUCER of x emits per-row partial (sum, sum of squares) of its
   * output (`row_stats_out`), the CONSUMER multiplies the raw x and applies mean / rstd per row in its epilogue.
   * GroupNorm statistics likewise come from the producer: `col_stats_out` receives per-channel partial sums of
   * 32-row blocks, reduced to (mean, rstd) per (image, group) by sg_groupnorm_finalize. */
  float* row_stats_out;  /* fp32 [sg_gemm_row_stat_parts(args)][M][2]: (sum, sumsq) of output row m over the columns of
                            one (N tile, epilogue warpgroup); modes LINEAR only */
  const float* ln_stats; /* fp32 [ln_parts][M][2]: row partials of the A operand (M = this GEMM's rows) or NULL */
  const float* ln_csum;  /* fp32 [N] column sums of W' (required with ln_stats)                                */
  int32_t ln_parts;
  float ln_eps;
  float* col_stats_out;  /* fp32 [M tiles * 4][N][2]: per (128-row tile, 32-row quarter, output channel) partial
                            (sum, sumsq) over the quarter's valid rows; modes LINEAR only                       */
  /* Strided output (LINEAR / GEGLU): element (n, y, x, col) is stored at out + n*out_stride_n + y*out_stride_h +
   * x*out_stride_w + col; all zero = dense rows of ldo elements.  Lets the four sub-pixel phases of
   * "nearest 2x up-sample -> conv3x3" (model/unet_2d_blocks.py:581-584 Upsample2D) be computed as four 2x2-tap
   * convolutions of the LOW-resolution tensor with pre-summed weights, each writing every second pixel of the
   * output: 4/9 of the FLOPs and no up-sampled tensor. */
  int64_t out_stride_w, out_stride_h, out_stride_n;
} sg_gemm_args;

int sg_gemm(const sg_gemm_args* args, void* stream);
/* number of row-statistics partials sg_gemm(args) writes per row (epilogue warpgroups x N tiles; depends on N, K and the tile grid) */
int sg_gemm_row_stat_parts(const sg_gemm_args* args);

/* ---- fused attention core (FlashAttention-style, tcgen05 + TMEM) -----------------------------
 * o[n, q, h*d:(h+1)*d] = softmax(scale * Q_h K_h^T) V_h for every (n, h); Q/K/V are read in place
 * from (n, token, channel) bf16 tensors with arbitrary row strides (so a fused QKV projection
 * output can be consumed directly).  Q lane n reads K/V lane kv_map[n] (identity when NULL), so
 * CFG lanes that share prior-frame features share one K/V projection.
 *
 * Replaces: CrossAttention core of attn1 / attn2 / attn3 — baddbmm + softmax + bmm or
 * xformers.ops.memory_efficient_attention (model/attention.py:255-260, 271-276, 285-290 via
 * diffusers CrossAttention; enabled at inference.py:58-64).
 */
typedef struct sg_attn_args {
  const void* q; /* element (n, i, c) at q + n*q_stride_n + i*q_stride_t + c */
  const void* k;
  const void* v;
  void* o;
  int64_t q_stride_n, q_stride_t;
  int64_t k_stride_n, k_stride_t;
  int64_t v_stride_n, v_stride_t;
  int64_t o_stride_n, o_stride_t;
  int32_t n_lanes, n_heads, head_dim; /* head_dim in {32, 40, 64, 80, 128, 160} */
  int32_t nq, nk;
  int32_t n_kv_lanes;    /* number of K/V lanes (0 = n_lanes)                                  */
  const int32_t* kv_map; /* optional device int32[n_lanes]: K/V lane of each Q lane; NULL = identity */
  float scale;        /* softmax scale (head_dim^-0.5)                                        */
  int64_t* debug_timeline; /* bring-up aid: device int64[12*16*8] phase clocks of CTA 0, or NULL     */
  float* lse; /* optional fp32 [n_lanes, n_heads, nq]: log2-domain log-sum-exp of every score row,
                 lse2 = log2(sum_k 2^(scale*log2(e)*s_k)), saved for sg_attention_bwd (training)  */
} sg_attn_args;

int sg_attention(const sg_attn_args* args, void* stream);

/* ---- attention backward (training of the image cross-attention, train_StorySalon_stage2.py:305-332) ----------
 * Given q, k, v, the forward output o, its row statistics lse (sg_attention) and the output gradient d_o, computes
 *   dV = P^T dO,  dK = dS^T Q,  dQ = dS K   with P = softmax(scale Q K^T), dS = scale * P o (dO V^T - rowsum(dO o O)),
 * recomputing P tile by tile on the tensor cores (FlashAttention-2 backward: one CTA per 128-key tile, looping over
 * the query tiles; S, dP, dV, dK, dQ accumulators in TMEM).  All tensors bf16 (lane, token, channel) with the strides
 * of sg_attn_args; dq_accum is an fp32 workspace [n_lanes, nq, n_heads * head_dim] that the CALLER zeroes: query-tile
 * contributions of different key tiles are added with red.global.add (the only non-deterministic summation order
 * of the library; dk / dv are written once).  head_dim in {32, 40, 64, 80, 128}; K/V lanes == Q lanes.
 * delta: fp32 workspace [n_lanes, n_heads, nq]. */
typedef struct sg_attn_bwd_args {
  const void* q; const void* k; const void* v; const void* o; const void* d_o;
  const float* lse;
  float* delta;
  float* dq_accum;
  void* dk; void* dv;
  int64_t q_stride_n, q_stride_t, k_stride_n, k_stride_t, v_stride_n, v_stride_t;
  int64_t o_stride_n, o_stride_t;    /* of o and d_o                    */
  int64_t dk_stride_n, dk_stride_t;  /* of dk and dv                    */
  int32_t n_lanes, n_heads, head_dim, nq, nk;
  float scale;
} sg_attn_bwd_args;
int sg_attention_bwd(const sg_attn_bwd_args* args, void* stream);

/* ---- normalisation / elementwise kernels (HBM-bound) ---------------------------------------- */

/* GroupNorm statistics over an NHWC tensor that is the channel-concat of up to two sources.
 * `stats` is an fp32 workspace of sg_groupnorm_ws_floats() floats holding fixed-order partial
 * (sum, sum of squares) per (image, pixel chunk, group); sg_groupnorm_apply finishes the reduction
 * (deterministic, no atomics).  Replaces the reduction half of nn.GroupNorm
 * (ResnetBlock2D.norm1/norm2, Transformer2DModel.norm model/attention.py:55,98, conv_norm_out
 * model/unet_2d_condition.py:259-262,477-479). */
int sg_groupnorm_ws_floats(int32_t n_img, int32_t groups); /* size of `stats` in floats */
int sg_groupnorm_stats(const void* x0, int32_t c0, const void* x1, int32_t c1, int32_t n_img,
                       int32_t hw, int32_t groups, float eps, float* stats, void* stream);
/* y = (x - mean) * rstd * gamma + beta, optionally followed by SiLU; x = concat(x0, x1) along c. */
int sg_groupnorm_apply(const void* x0, int32_t c0, const void* x1, int32_t c1, int32_t n_img,
                       int32_t hw, int32_t groups, float eps, const float* stats, const float* gamma,
                       const float* beta, int32_t silu, void* y, void* stream);

/* GroupNorm from producer-side statistics (no separate pass over the activations).
 * sg_groupnorm_finalize: reduces the col_stats_out partials of the GEMMs that produced x0 and x1 (the channel-concat
 * x0 | x1 of the skip connections, model/unet_2d_blocks.py:609,626,716) to mean_rstd[n_img][groups][2], in a fixed
 * order (deterministic), accumulating in fp64 so that E[x^2] - mean^2 does not cancel for channels with |mean| >> std.
 * Source i is described by sg_gn_parts: `parts` = `reps` consecutive col_stats_out buffers [(M tiles) * 4][c][2] of GEMMs
 * that tiled their (identical) output pixel grids in boxes of bw x bh x bn pixels (bw * bh * bn == 128); image n owns the
 * tiles [ (n / bn) * tiles_xy, (n / bn + 1) * tiles_xy ) and, inside each, the 4 / bn quarters starting at
 * (n % bn) * (4 / bn): tiles_xy = (tiles along x) * (tiles along y), imgs_per_tile = bn (1, 2 or 4: a 32-row quarter never
 * spans two images).  reps > 1: the tensor was written by several GEMMs (the 4 sub-pixel phases of an up-sampling conv),
 * each covering hw / reps pixels of every image.
 * sg_groupnorm_apply_ms: y = [SiLU]((x - mean) * rstd * gamma + beta) with those statistics. */
typedef struct sg_gn_parts {
  const float* parts;
  int32_t c;             /* channels of this source                */
  int32_t tiles_xy;
  int32_t imgs_per_tile;
  int32_t reps;
  int32_t m_tiles;       /* M tiles of ONE producing GEMM (stride between the `reps` buffers = m_tiles * 4 * c * 2 floats) */
} sg_gn_parts;
int sg_groupnorm_finalize(const sg_gn_parts* src0, const sg_gn_parts* src1 /* or NULL */, int32_t n_img, int32_t hw,
                          int32_t groups, float eps, float* mean_rstd, void* stream);
int sg_groupnorm_apply_ms(const void* x0, int32_t c0, const void* x1, int32_t c1, int32_t n_img, int32_t hw,
                          int32_t groups, const float* mean_rstd, const float* gamma, const float* beta,
                          int32_t silu, void* y, void* stream);

/* LayerNorm over the channel dim of a (rows, c) bf16 tensor; up to two affine outputs from one
 * statistics pass (norm2 and norm4 read the same tensor, model/attention.py:268,283). */
int sg_layernorm(const void* x, int64_t rows, int32_t c, float eps, const float* gamma_a,
                 const float* beta_a, void* y_a, const float* gamma_b, const float* beta_b,
                 void* y_b, void* stream);

/* Sinusoidal timestep features [cos | sin] (flip_sin_to_cos=True, freq_shift=0), fp32 (n, dim).
 * Replaces diffusers Timesteps (model/unet_2d_condition.py:392). */
int sg_timestep_features(const float* t, int32_t n, int32_t dim, float* out, void* stream);

/* Small dense layers on (n, k) rows where n is the image-lane count (time embedding MLP and the 22
 * time_emb_proj layers): out = act_out(W * act_in(x) + b); x, out fp32; W bf16 [m, k].
 * Replaces TimestepEmbedding and ResnetBlock2D.time_emb_proj (model/unet_2d_condition.py:398). */
int sg_small_linear(const float* x, int32_t n, int32_t k, const void* w, const float* b, int32_t m,
                    int32_t silu_in, int32_t silu_out, float* out, void* stream);

/* Nearest-neighbour 2x upsample NHWC bf16 (Upsample2D, model/unet_2d_blocks.py:581-584,656-658). */
int sg_upsample2x(const void* x, int32_t n, int32_t h, int32_t w, int32_t c, void* y, void* stream);

/* conv_in front end (model/unet_2d_condition.py:124,411): 3x3 patches of the NCHW fp32 latent
 * (cin <= 7) -> bf16 rows (n*h*w, 64): column (ky*3+kx)*cin + ci, zero padded.  conv_in itself is then
 * one sg_gemm with K = 64.  conv_out (model/unet_2d_condition.py:268,480) is sg_gemm with the
 * SG_EPI_NCHW_F32 epilogue. */
int sg_im2col3x3_nchw(const float* x_nchw, int32_t n, int32_t cin, int32_t h, int32_t w, void* patches,
                      void* stream);

/* x_t = sqrt(acp[t]) * x + sqrt(1 - acp[t]) * eps  (DDIMScheduler.add_noise at
 * model/pipeline.py:420-424); per-image timestep index t[n]; all fp32; numel per image = chw. */
int sg_add_noise(const float* x, const float* eps, const float* acp, const int32_t* t, int32_t n,
                 int32_t chw, float* out, void* stream);

/* Dual classifier-free guidance + eta=0 DDIM update (model/pipeline.py:456-461):
 *   e = u + s_img (i - u) + s_txt (a - i);  x0 = (x - sqrt(1-a_t) e)/sqrt(a_t);
 *   x' = sqrt(a_prev) x0 + sqrt(1 - a_prev) e.
 * eps3 = (3, b, chw) lanes [uncond | image | all]; latents updated in place; `alphas` is a DEVICE
 * pair (a_t, a_prev) so a captured CUDA graph of one step can be replayed for every timestep. */
int sg_cfg_ddim_step(const float* eps3, float* latents, int32_t b, int32_t chw, float s_img,
                     float s_txt, const float* alphas, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STORYGEN_B200_H */
