/*
 * evcdiff.h — C ABI of libevcdiff.so, the sm_100a (B200) kernel library behind the conditional
 * video-diffusion sampling path (models/__init__.py samplers, models/pndm.py,
 * models/better NCSN++ UNet, models/unet.py).
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name ends in _host;
 *  - the caller owns every buffer (the library never allocates device memory);
 *  - every entry point enqueues work on `stream` and returns without synchronising, so a whole
 *    sampling loop can be captured in one CUDA graph;
 *  - return value: 0 = ok, negative = evc_status (evc_last_error() gives a message);
 *    shape / alignment violations are errors, never fallbacks.  There is no CPU path.
 *  - activations are NHWC bf16 ("pixel rows" of C channels); sampler state x_t is NCHW fp32, the
 *    layout the reference samplers return (models/__init__.py:339-342).
 *
 * The reference has no FFI on this path except one pybind op,
 *   upfirdn2d(input, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1) -> Tensor
 *   (models/better/op/upfirdn2d.cpp:12-23), which evc_fir_resample replaces; all other entry points
 * replace the torch ops named next to them.
 */
#ifndef EVCDIFF_H_
#define EVCDIFF_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* evc_stream_t; /* cudaStream_t */

enum evc_status {
  EVC_OK = 0,
  EVC_ERR_INVALID = -1, /* bad shape / alignment / argument */
  EVC_ERR_CUDA = -2,    /* a CUDA runtime / driver call failed */
  EVC_ERR_UNSUPPORTED = -3
};

int evc_version(void);
const char* evc_last_error(void);
/* Launch the kernels of the sampling loop with programmatic dependent launch (the prologue of kernel N+1 overlaps the
 * tail of kernel N).  Pays off for small, launch-latency-bound batches.  The EVC_PDL environment variable overrides. */
void evc_set_pdl(int enabled);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
int64_t evc_launch_count(void);
/* sizeof() of the ABI structs as this library was compiled: 0 evc_tensor4, 1 evc_gemm_desc, 2 evc_attn_desc,
 * 3 evc_step_coef, 4 evc_pndm_coef; -1 for an unknown id.  Bindings check their own struct layouts against it at
 * load time (a shorter foreign struct would make the library read past its end). */
int64_t evc_struct_size(int32_t which);

/* ------------------------------------------------------------------------------------------------
 * Dense contractions: implicit-GEMM convolution / batched GEMM on tcgen05 tensor cores.
 * Replaces nn.Conv2d 3x3 / 1x1 (models/better/layers.py:89-113), NIN (layers.py:535-544), the
 * attention einsums (layerspp.py:239-243) and models/unet.py Nin / einsum (unet.py:49-63,114-119).
 *
 *   out[b, y, x, n] = alpha * ( sum_seg sum_tap sum_c  A_seg[b, y+dy, x+dx, c] * Wt[zb, n, k(seg,tap,c)]
 *                               + bias[n] + resid[b, y, x, n] )
 *
 * A segments: up to 3 bf16 tensors viewed as (B, H, W, C) (or (B, 2H, 2W, C) with stride 2) with element strides; taps = 1 (1x1 / GEMM)
 * or 9 (3x3, zero padding 1).  C must be a multiple of 8 (64 for full tensor-core efficiency).  Several segments implement a virtual
 * channel concat and the fused 1x1 skip branch of a residual block (extra K).
 * Wt: bf16 (w_batches, N, K_total), K contiguous, K order = segment-major, tap-major, channel-minor.
 * w_batches is 1 (shared weights) or B (per-sample B operand, e.g. K / V^T in attention; then the
 * output tile never mixes samples).
 * ---------------------------------------------------------------------------------------------- */
typedef struct evc_tensor4 {
  const void* ptr;    /* bf16 */
  int32_t C, W, H, B; /* extents, channel innermost */
  int64_t stride_w, stride_h, stride_b; /* element strides (channel stride is 1) */
} evc_tensor4;

enum evc_out_mode {
  EVC_OUT_BF16_ROWS = 0, /* out[(b*H*W + y*W + x) * out_ld + n]            bf16 */
  EVC_OUT_F32_ROWS = 1,  /* same addressing, fp32                              */
  EVC_OUT_BF16_T = 2,    /* out[b*out_bs + n*out_ld + (y*W + x)]           bf16 (channel-major / transposed) */
  EVC_OUT_F32_T = 3      /* same addressing, fp32 (NCHW result of the last conv) */
};

typedef struct evc_gemm_desc {
  int32_t n_seg;
  evc_tensor4 a[3];
  int32_t taps[3];
  const void* w;
  int32_t w_rows;    /* N */
  int32_t w_k;       /* K_total = sum_seg taps*C */
  int32_t w_batches; /* 1 or B */
  int64_t w_row_stride, w_batch_stride; /* elements */
  int32_t B, H, W;   /* output extent (== input extent, stride 1) */
  int32_t bn;        /* N tile: multiple of 16, 16..256 */
  void* out;
  int32_t out_mode;
  int64_t out_ld, out_bs;
  const float* bias;  /* N floats or NULL */
  const void* resid;  /* bf16 rows [(b*H*W + y*W + x) * resid_ld + n] or NULL */
  int64_t resid_ld;
  float alpha;
  int32_t max_ctas;   /* 0 = number of SMs */
  /* optional fused GroupNorm statistics of the stored output (EVC_OUT_BF16_ROWS only, H*W % 32 == 0):
   * stats[(b*N + n)*2 + {0,1}] += {sum, sum of squares} * 2^20 as 64-bit integers (order-independent, hence
   * deterministic); the caller zeroes the buffer before the launch.  Same format as evc_gn_stats. */
  int64_t* stats;
  /* convolution stride, 0/1 or 2 (models/unet.py:219 down-sampling conv): every A segment then has extent
   * (B, stride*H, stride*W, C) and tap (dy,dx) reads input pixel (stride*y + dy, stride*x + dx). */
  int32_t stride;
  /* 0 = automatic, 1 = one CTA per 128-row tile, 2 = CTA pair (tcgen05 cta_group::2, 256-row tile; the pair shares
   * one B tile, each CTA staging half of it).  2 needs shared weights (w_batches == 1) and bn % 16 == 0. */
  int32_t cta_group;
  /* Split-precision ("fp32-tolerance") operands, all NULL in the default bf16 mode.  Every bf16 tensor x is then a
   * pair (hi, lo) with hi = bf16(x), lo = bf16(x - hi) (16 mantissa bits); a product is accumulated as
   * hi*hi + hi*lo + lo*hi in fp32.  a_lo[s] / w_lo / resid_lo have the layout of a[s] / w / resid; a bf16 output is
   * written as the pair (out, out_lo).  Cost: 3x the MMA work and 2x the activation traffic. */
  evc_tensor4 a_lo[3];
  const void* w_lo;
  void* out_lo;
  const void* resid_lo;
  /* Fused GroupNorm apply (NULL / 0 = off): `out` receives SiLU(GN(acc + bias) * gamma' + beta') instead of the raw
   * convolution output, i.e. conv -> GroupNorm -> SiLU of ResnetBlockBigGANppGN (layerspp.py:611-613) in one launch,
   * without the raw tensor ever reaching memory.  The epilogue keeps each accumulator tile in tensor memory until all
   * tiles of its sample have added their statistics (gn_ticket[b] counts them), then normalises from it.
   * gn_ss = [gamma' (N) | beta' (N)] fp32, gamma' = 1 + scale when gn_adagn (AdaGN, layerspp.py:520-527); gn_ticket =
   * B int32 zeroed by the caller before every launch (like stats).  Needs: stats, bias, no residual, alpha = 1, bf16
   * row output, H*W a multiple of 128 with W <= 128, bn % 32 == 0 (bn <= 192 so that two tile slots fit in shared
   * memory), N % gn_groups == 0, and one CTA per SM (the grid
   * is at most the SM count, so every tile of a sample is in flight together). */
  const float* gn_ss;
  int32_t* gn_ticket;
  float gn_eps;
  int32_t gn_groups;
  int32_t gn_adagn;
  /* Split-K for launches with few M tiles (small batches; the 8x8 / 16x16 levels): split_k > 1 cuts the K loop of
   * every tile into split_k slices handled by different CTAs; each writes its fp32 partial tile to sk_ws, and the
   * CTA that arrives last on the tile's ticket adds the slices in the fixed order 0..split_k-1 (bit-reproducible
   * whatever the arrival order) and runs the normal epilogue (bias, residual, alpha, store, fused statistics).
   * 0 / 1 = off.  Needs bn % 32 == 0, no gn_ss, no split-precision planes.
   * sk_ws: split_k * roundup2(m_tiles) * 128 * (ceil(N / bn) * bn) floats (sk_ws_bytes is checked); may be shared
   * by plans that run one after the other on a stream.  sk_ticket: ceil(N / bn) * roundup2(m_tiles) int32, zero
   * before the first launch (the kernel leaves them zero), NOT shared between plans in flight. */
  int32_t split_k;
  void* sk_ws;
  int64_t sk_ws_bytes;
  int32_t* sk_ticket;
} evc_gemm_desc;

typedef struct evc_gemm_plan evc_gemm_plan;

/* Validates, encodes the TMA descriptors (host side, no device work) and returns a reusable plan. */
int evc_gemm_plan_create(const evc_gemm_desc* desc, evc_gemm_plan** plan);
/* bias_override: NULL = use desc->bias. Used for per-label bias tables (models/unet.py:87-88). */
int evc_gemm_plan_launch(const evc_gemm_plan* plan, const float* bias_override, evc_stream_t stream);
/* same, with the [gamma' | beta'] row of the current label for a plan created with gn_ss (NULL = desc->gn_ss) */
int evc_gemm_plan_launch_gn(const evc_gemm_plan* plan, const float* bias_override, const float* gn_ss_override,
                            evc_stream_t stream);
/* General form: out_override (NULL = desc->out) redirects a per-thread-store output (fp32 / transposed modes, e.g. the
 * final conv writing eps straight into a slot of the F-PNDM history ring, models/pndm.py:41-52) -- not valid for
 * plans that store through TMA (bf16 row outputs of whole tiles), whose descriptors hold the pointer. */
int evc_gemm_plan_launch_ex(const evc_gemm_plan* plan, const float* bias_override, const float* gn_ss_override,
                            void* out_override, evc_stream_t stream);
/* Synchronising debug query (not capturable): tile waits of the fused GroupNorm apply that gave up since the last
 * call. 0 unless a caller launched a gn_ss plan without zeroing gn_ticket; -1 without a device. */
int64_t evc_gemm_fault_count(void);
void evc_gemm_plan_destroy(evc_gemm_plan* plan);
/* 2 * M * N * K of one launch (dense FLOPs, for the roofline) */
double evc_gemm_plan_flops(const evc_gemm_plan* plan);
/* 1 or 2: the CTA grouping the plan selected */
int evc_gemm_plan_cta_group(const evc_gemm_plan* plan);

/* ------------------------------------------------------------------------------------------------
 * Fused self-attention forward (flash-style, tcgen05): out[b, q, h*d + :] = softmax(scale * Q_h K_h^T) V_h with
 *   Q_h = qk[b, q, h*d : (h+1)*d],  K_h = qk[b, k, C + h*d : C + (h+1)*d]   (qk: (B, N, >= 2C) bf16 rows, row stride qk_ld)
 *   V_h^T = vT[b, h*d : (h+1)*d, k]                                          (vT: (B, C, N) bf16, row stride vT_ld)
 *   or, with vT == NULL,  V_h = v[b, k, h*d : (h+1)*d]                       (v: (B, N, >= C) bf16 rows, row stride v_ld;
 *                                                                             e.g. columns 2C.. of one fused q|k|v projection)
 * Replaces einsum -> softmax -> einsum of AttnBlockpp / AttnBlock (layerspp.py:239-243, unet.py:114-119); the N x N
 * score matrix never leaves the SM.  Needs N % 64 == 0, d = C/heads % 64 == 0 (% 128 for 256 < d <= 384); head dims
 * above 384 (unet.py 'deeper': one head of 768) split the output columns of a head over several CTAs
 * (EVC_ERR_UNSUPPORTED otherwise: the caller then uses the batched-GEMM + evc_softmax_rows formulation).
 * ---------------------------------------------------------------------------------------------- */
typedef struct evc_attn_desc {
  const void* qk;
  int64_t qk_ld;
  const void* vT;
  int64_t vT_ld;
  void* out;       /* (B, N, out_ld >= C) bf16 rows */
  int64_t out_ld;
  int32_t B, N, C, heads;
  float scale;
  const void* v;   /* used when vT == NULL */
  int64_t v_ld;
} evc_attn_desc;
typedef struct evc_attn_plan evc_attn_plan;
int evc_attn_plan_create(const evc_attn_desc* desc, evc_attn_plan** plan);
int evc_attn_plan_launch(const evc_attn_plan* plan, evc_stream_t stream);
void evc_attn_plan_destroy(evc_attn_plan* plan);
double evc_attn_plan_flops(const evc_attn_plan* plan);

/* ------------------------------------------------------------------------------------------------
 * GroupNorm statistics and the fused normalise / AdaGN / affine / SiLU pass.
 * Replaces nn.GroupNorm + get_act_norm (layerspp.py:465-549) and Normalize+Swish (unet.py:44-46,90-95).
 * ---------------------------------------------------------------------------------------------- */
/* stats[(b*c_total + c_off + c)*2 + {0,1}] = {sum, sum of squares} * 2^20 (int64 fixed point) over the HW pixels of x
 * (bf16 rows, row stride ldx), C % 8 == 0.  Deterministic (no floating-point atomics).  `workspace` is caller-owned scratch of at least
 * evc_gn_stats_workspace() bytes whose first bytes (the ticket counters) must be zero before the FIRST use; the
 * kernel leaves them zero, so one workspace can be shared by every call on the same stream. */
int evc_gn_stats_workspace(int32_t B, int32_t HW, int32_t C, int64_t* bytes);
int evc_gn_stats(const void* x, int64_t ldx, int32_t B, int32_t HW, int32_t C, int64_t* stats, int32_t c_total,
                 int32_t c_off, void* workspace, int64_t workspace_bytes, evc_stream_t stream);

/* y[b, p, c] = act( (x[b,p,c] - mean_g) * rstd_g * (gamma'[c]) + beta'[c] ), group g = c / (C/groups) over the
 * concatenation [x0 | x1] (x1 may be NULL).  mean/rstd come from the per-channel sums `stats0` (B,C0,2) and
 * `stats1` (B,C1,2) (int64 fixed point, 2^20) written by evc_gn_stats or by the GEMM epilogue (biased variance,
 * eps); a group may straddle the concat boundary.
 *   adagn != 0: gamma' = 1 + ss[c], beta' = ss[C + c]          (ss = Dense_0(SiLU(temb)) row, 2*C floats)
 *   adagn == 0: gamma' = ss[c], beta' = ss[C + c]              (GroupNorm affine weight | bias)
 * silu != 0 applies x*sigmoid(x).  Output bf16 rows with row stride C (materialises the concat). */
int evc_gn_apply(const void* x0, int32_t C0, const void* x1, int32_t C1, int32_t B, int32_t HW, const int64_t* stats0,
                 const int64_t* stats1, int32_t groups, float eps, const float* ss, int32_t adagn, int32_t silu,
                 void* y, evc_stream_t stream);

/* Split-precision ("fp32-tolerance") variants: every bf16 tensor is a pair (x, x_lo) with value x + x_lo (see
 * evc_gemm_desc.a_lo).  Same arithmetic in fp32, results written back as pairs. */
int evc_gn_stats_split(const void* x, const void* x_lo, int64_t ldx, int32_t B, int32_t HW, int32_t C, int64_t* stats,
                       int32_t c_total, int32_t c_off, void* workspace, int64_t workspace_bytes, evc_stream_t stream);
int evc_gn_apply_split(const void* x0, const void* x0_lo, int32_t C0, const void* x1, const void* x1_lo, int32_t C1,
                       int32_t B, int32_t HW, const int64_t* stats0, const int64_t* stats1, int32_t groups, float eps,
                       const float* ss, int32_t adagn, int32_t silu, void* y, void* y_lo, evc_stream_t stream);
int evc_fir_resample_split(const void* x, const void* x_lo, void* y, void* y_lo, int32_t B, int32_t H, int32_t W,
                           int32_t C, int32_t up, evc_stream_t stream);
int evc_softmax_rows_split(const float* S, void* P, void* P_lo, int64_t rows, int32_t cols, evc_stream_t stream);
int evc_pack_nchw_split(const void* src, int32_t src_is_f64, int32_t B, int32_t C, int32_t HW, float scale, float shift,
                        void* dst, void* dst_lo, int32_t Cpad, int32_t c_off, evc_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * FIR [1,3,3,1] x2 resampling (upfirdn2d modes used by upsample_2d / downsample_2d,
 * up_or_down_sampling.py:196-258; kernels op/upfirdn2d_kernel.cu:107-207).  NHWC bf16 in/out.
 * up != 0: (B,H,W,C) -> (B,2H,2W,C); else (B,H,W,C) -> (B,H/2,W/2,C).
 * ---------------------------------------------------------------------------------------------- */
int evc_fir_resample(const void* x, void* y, int32_t B, int32_t H, int32_t W, int32_t C, int32_t up,
                     evc_stream_t stream);

/* Fused prologue of the up / down res blocks (layerspp.py:598-611: h = act(GN(x)); h = resample(h); x = resample(x)):
 * one read of x = [x0 | x1] produces y_act = FIR(SiLU(GN(x) * gamma' + beta')) (B,H',W',C0+C1) and the resampled skip
 * inputs y_raw0 (B,H',W',C0), y_raw1 (B,H',W',C1).  Arguments as evc_gn_apply / evc_fir_resample; the activation is
 * not rounded to bf16 before the filter.  Zero padding applies after the activation, as in upfirdn2d(h). */
int evc_gn_fir(const void* x0, int32_t C0, const void* x1, int32_t C1, int32_t B, int32_t H, int32_t W,
               const int64_t* stats0, const int64_t* stats1, int32_t groups, float eps, const float* ss, int32_t adagn,
               int32_t up, void* y_act, void* y_raw0, void* y_raw1, evc_stream_t stream);

/* nearest-neighbour x2 upsample, NHWC bf16 (models/unet.py:123-131) */
int evc_nearest_up2(const void* x, void* y, int32_t B, int32_t H, int32_t W, int32_t C, evc_stream_t stream);

/* row softmax: P[r, :] = softmax(S[r, :]) ; S fp32 (already scaled), P bf16 (layerspp.py:241, unet.py:117) */
int evc_softmax_rows(const float* S, void* P, int64_t rows, int32_t cols, evc_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Time embedding (layers.py:504-518, ncsnpp_more.py:277-281) — evaluated once per distinct label.
 * ---------------------------------------------------------------------------------------------- */
/* out[l, :] = [sin(t_l * f_j) | cos(t_l * f_j)], f_j = freqs[j], j < dim/2 ; fp32 */
int evc_timestep_embedding(const float* labels, const float* freqs, int32_t L, int32_t dim, float* out,
                           evc_stream_t stream);
/* y[l, n] = act_out( sum_k act_in(x[l,k]) * W[n,k] + b[n] ), fp32, act flags: 0 none, 1 SiLU */
int evc_linear_f32(const float* x, const float* W, const float* b, float* y, int32_t L, int32_t K, int32_t N,
                   int32_t act_in, int32_t act_out, evc_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Layout / sampler-state kernels.
 * ---------------------------------------------------------------------------------------------- */
/* NCHW (fp32 or fp64) -> channels [c_off, c_off+C) of an NHWC bf16 buffer with Cpad channels.
 * v = scale * src + shift (data_transform 2x-1 fused, function.py:62-63). */
int evc_pack_nchw(const void* src, int32_t src_is_f64, int32_t B, int32_t C, int32_t HW, float scale, float shift,
                  void* dst, int32_t Cpad, int32_t c_off, evc_stream_t stream);
int evc_fill_zero(void* p, int64_t bytes, evc_stream_t stream);

/* One sampler update, NCHW fp32 state (models/__init__.py:289-335 ddpm, :166-169 ddim, :196/:333-335 denoise):
 *   mode 0:  x0 = k0*(x - k1*eps); if clip: x0 = clamp(x0,-1,1);  x' = c_x0*x0 + c_x*x + c_eps*eps + c_noise*noise
 *   mode 1:  x' = x - k1*eps                                    (final denoise)
 * also writes bf16(x') into the NHWC UNet input `xin` (row stride Cpad, Cpad % 8 == 0): channels [0, CP) with
 * CP = C rounded up to 8 -- whole 16-byte chunks, the pad channels [C, CP) are written as zero, so the conditioning
 * frames of the virtual torch.cat (ncsnpp_more.py:256-257) live at channel CP onwards and the first conv's weight has
 * zero columns for the pad channels. noise may be NULL when c_noise == 0. x_out may alias x. Vectorised (float4
 * planes, 4 pixels per thread) when HW % 4 == 0 and all pointers are 16-byte aligned. */
typedef struct evc_step_coef {
  int32_t mode, clip;
  float k0, k1, c_x0, c_x, c_eps, c_noise;
} evc_step_coef;
int evc_sampler_update(const float* x, const float* eps, const float* noise, float* x_out, void* xin, int32_t B,
                       int32_t C, int32_t HW, int32_t Cpad, const evc_step_coef* coef_host, evc_stream_t stream);

/* PNDM transfer with a fused linear multistep combination (models/pndm.py:19-33, :15, :47):
 *   et = w_scale * sum_i w[i]*e[i] (n_e <= 4; evaluated left to right) ; x' = x + d*(p*x - q*et) ; clip -> clamp(x',-1,1)
 * writes x' (fp32 NCHW) to x_out, optionally et to et_out, and bf16(x') into xin. */
typedef struct evc_pndm_coef {
  int32_t n_e, clip;
  float w[4];
  float w_scale; /* et = w_scale * sum_i w[i]*e[i] : 1/6 (Runge-Kutta), 1/24 (Adams-Bashforth), 1 (plain) */
  float d, p, q;
} evc_pndm_coef;
int evc_pndm_update(const float* x, const float* const* e_host, float* x_out, float* et_out, void* xin, int32_t B,
                    int32_t C, int32_t HW, int32_t Cpad, const evc_pndm_coef* coef_host, evc_stream_t stream);

/* frames = clamp((x+1)/2, 0, 1), NCHW fp32 -> NCHW fp32 (inverse_data_transform, function.py:73-82) */
int evc_inverse_transform(const float* x, float* frames, int64_t n, evc_stream_t stream);
/* out = uint8(round(255 * frames)) -- the storage format of the data set (city_sender.py:487 divides it by 255);
 * used for the end-of-run gather of predicted frames (4x less NVLink traffic than fp32). */
int evc_frames_to_uint8(const float* frames, uint8_t* out, int64_t n, evc_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Sender-side accept decision on the GPU (next rows of the path): per-frame PSNR in float64
 * (city_sender.py:255-258 cal_psnr) and the accept-until-first-failure prefix of decide_5to5 (:353-374).
 * ---------------------------------------------------------------------------------------------- */
/* psnr[f] = 10 log10(maxvalue^2 / mean((a_f - b_f)^2)), frames of frame_elems fp32 values */
int evc_frame_psnr(const float* a, const float* b, int32_t n_frames, int64_t frame_elems, double maxvalue, double* psnr,
                   evc_stream_t stream);
/* counts[v] = number of leading frames of video v (n_frames scores each) passing the threshold */
int evc_accept_prefix(const double* score, int32_t n_videos, int32_t n_frames, double threshold,
                      int32_t higher_is_better, int32_t* counts, evc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* EVCDIFF_H_ */
