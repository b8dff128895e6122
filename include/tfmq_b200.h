/*
 * tfmq_b200 -- C ABI of the B200 (sm_100a) hot path for TFMQ-DM's w4a8 DDIM
 * denoising step and its PTQ calibration passes.
 *
 * The reference (ModelTC/TFMQ-DM) is pure PyTorch and has no FFI; the boundary a
 * maintainer binds is this header (ctypes stub: INTEGRATION.md).  Each entry
 * point names the reference code it replaces (paths relative to the reference
 * root).  All pointers are DEVICE pointers borrowed from the caller unless the
 * name says host; every call is asynchronous on `stream` (a cudaStream_t passed
 * as void*), allocates nothing, never synchronises and is CUDA-graph
 * capturable.  Return value: 0 = OK, otherwise a tfmq_status; the message is in
 * tfmq_last_error().  Unsupported shapes are an error, never a fallback.
 *
 * Activation layout is NHWC ("pixel-major"): element (n,y,x,c) of a tensor with
 * pixel pitch `ld` (elements) lives at ((n*H+y)*W+x)*ld + c.  A `ld` larger than
 * the channel count lets a tensor live inside a wider concat buffer.
 * Quantised activations ("codes") are u8 NHWC with a `halo`-pixel border that
 * holds the zero-point code, so zero padding of the de-quantised tensor
 * (quant/quant_layer.py:338 F.conv2d(padding=1)) is exact in the integer domain.
 */
#ifndef TFMQ_B200_H
#define TFMQ_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tfmq_ctx tfmq_ctx;

typedef enum {
  TFMQ_OK = 0,
  TFMQ_ERR_ARG = 1,         /* null pointer / bad enum / misaligned pointer */
  TFMQ_ERR_SHAPE = 2,       /* shape the kernels do not support */
  TFMQ_ERR_CUDA = 3,        /* CUDA runtime / driver error */
  TFMQ_ERR_UNAVAILABLE = 4  /* no sm_100 device / driver entry point missing */
} tfmq_status;

/* context: one per device per process, ONE DEVICE PER PROCESS (the launchers cache the kernels' opt-in shared-memory
 * attribute per process, matching the one-process-per-GPU model of the path); not thread-safe */
int tfmq_create(tfmq_ctx** out, int device);
int tfmq_destroy(tfmq_ctx* ctx);
const char* tfmq_last_error(tfmq_ctx* ctx);
/* ABI version of this header (bumped on any signature change) */
int tfmq_abi_version(void);
/* number of kernels launched through this context since creation */
int64_t tfmq_launch_count(tfmq_ctx* ctx);

/* ------------------------------------------------------------------------- *
 * Weight pre-pack.  Replaces the per-forward weight fake-quant of
 * UniformAffineQuantizer.forward (quant/quant_layer.py:211-227) and the hard
 * branch of AdaRoundQuantizer.forward (quant/adaptive_rounding.py:51-70):
 *   q = clamp(rint(w/delta) + zp, 0, 15)                (alpha == NULL)
 *   q = clamp(floor(w/delta) + (alpha >= 0) + zp, 0, 15) (alpha != NULL)
 * w is [cout][k] with k = taps*cin in (tap, cin) order (the caller permutes the
 * OIHW weight to OHWI first).  Outputs:
 *   codes  [cout][k]   u8, one code per byte (for checks / checkpoint export)
 *   packed [cout][k/2] two codes per byte; inside every 32-code group g byte i
 *          holds code g*32+i in the low nibble and code g*32+16+i in the high one
 *   wsum   [cout]      int32 sum over k of (q - zp)
 * delta, zp are per-output-channel [cout] (channel_wise=True everywhere in the
 * reference entry points).  k must be a multiple of 32.
 * ------------------------------------------------------------------------- */
int tfmq_pack_w4(tfmq_ctx* ctx, const float* w, const float* delta, const float* zp, const float* alpha_or_null,
                 int cout, int k, uint8_t* codes_or_null, uint8_t* packed, int32_t* wsum, void* stream);

/* ------------------------------------------------------------------------- *
 * GroupNorm statistics (torch.nn.GroupNorm as used by ddim/models/diffusion.py:
 * 31-33 and ldm/modules/diffusionmodules/util.py:214-216): per (image, group)
 * sum and sum of squares in double, accumulated with atomics into
 * stats[n][groups][2]; the caller zeroes `stats` (tfmq_fill_zero) beforehand.
 * ------------------------------------------------------------------------- */
int tfmq_gn_stats(tfmq_ctx* ctx, const float* x, int64_t ld, int n, int hw, int c, int groups, double* stats,
                  void* stream);
int tfmq_fill_zero(tfmq_ctx* ctx, void* p, size_t bytes, void* stream);

/* GroupNorm statistics of a convolution's OUTPUT, accumulated by the convolution's own epilogue
 * (fused when an output tile lies inside one image, otherwise by a follow-up tfmq_gn_stats_part launch).
 * The output's channel ch belongs to group (ch_off + ch) / cpg of the tensor that will be normalised
 * (ch_off != 0 when the output is the second part of a skip concatenation). */
typedef struct {
  double* stats;        /* [n][groups][2], accumulated */
  int cpg;              /* channels per group of the normalised tensor */
  int ch_off;           /* offset of this tensor's channel 0 inside the normalised tensor */
  int groups;
  int reserved;
} tfmq_gn_target;
/* partial statistics of x (c channels) into the group geometry above */
int tfmq_gn_stats_part(tfmq_ctx* ctx, const float* x, int64_t ld, int n, int hw, int c, const tfmq_gn_target* target,
                       void* stream);

/* ------------------------------------------------------------------------- *
 * Activation producer: [GroupNorm-apply] -> [SiLU] -> [act fake-quant codes].
 * Replaces GN + nonlinearity + UniformAffineQuantizer.forward on the input of
 * a QuantLayer (quant/quant_block.py:415-434,178-209; quant/quant_layer.py:
 * 223-226): code = clamp(rintf(x / delta) + zp, 0, 255) with a true fp32
 * divide and round-half-even.  (delta, zp) are read from device memory `aq`
 * (2 floats) so FSC can swap them per timestep without host involvement.
 * ------------------------------------------------------------------------- */
typedef struct {
  const float* src;      /* fp32 NHWC */
  int64_t src_ld;
  int n, h, w, c;        /* SOURCE extent */
  int upsample;          /* 1: nearest x2 (output is 2h x 2w), 0: none */
  /* GroupNorm (optional) */
  const double* gn_stats; /* [n][groups][2] from tfmq_gn_stats, or NULL */
  const float* gamma;     /* [c] */
  const float* beta;      /* [c] */
  int groups;
  float eps;
  int silu;              /* apply x*sigmoid(x) after GN */
  /* output: exactly one of dst_u8 / dst_f32 / (dst_hi, dst_lo) */
  const float* aq;       /* device (delta, zp) for dst_u8 */
  uint8_t* dst_u8;       /* [n][H+2*halo][W+2*halo][dst_c] codes, border = zp */
  int halo;              /* 0 or 1 */
  int dst_c;             /* channel count of the destination buffer */
  int dst_c_off;         /* first destination channel written */
  float* dst_f32;        /* fp32 NHWC (no halo) */
  int64_t dst_ld;
  /* third kind of output: fp16 hi / lo planes (hi = half(v), lo = half(v - hi)) for tfmq_conv_h16 */
  void* dst_hi;
  void* dst_lo;
  int64_t dst_h_ld;      /* in halves, multiple of 4 */
  /* token producers of the transformer blocks (quant/quant_block.py:248-299; ldm/modules/attention.py:37-44,205-216),
   * one NHWC "pixel" = one token:
   *   ln_gamma != NULL: LayerNorm over the c channels (eps ln_eps) instead of GroupNorm;
   *   geglu != 0: src rows hold 2c channels [value | gate], the output is value * gelu(gate) (c channels). */
  const float* ln_gamma;
  const float* ln_beta;
  float ln_eps;
  int geglu;
} tfmq_act_desc;
int tfmq_act_prepare(tfmq_ctx* ctx, const tfmq_act_desc* d, void* stream);

/* ------------------------------------------------------------------------- *
 * w4a8 convolution / linear as an implicit GEMM on tcgen05 (kind::i8).
 * Replaces QuantLayer.forward (quant/quant_layer.py:306-340) with use_wq and
 * use_aq on, plus the elementwise tail of the enclosing block:
 *   out = delta_a*delta_w[c] * (sum_k a_code*(q-zp_w) - zp_a*wsum[c]) + bias[c]
 *         [+ emb[n][c]]  (h + temb_proj(...)[:, :, None, None], quant_block.py:430)
 *         [+ res[pixel][c]]  (x + h, quant_block.py:443)
 * 3x3 (stride 1, pad 1; input has halo 1) and 1x1 / linear (halo 0).
 * ------------------------------------------------------------------------- */
typedef struct {
  const uint8_t* act;   /* u8 codes [n][h+2*halo][w+2*halo][cin] */
  int n, h, w, cin, cout;
  int ksize;            /* 1 or 3 */
  const uint8_t* packed; /* tfmq_pack_w4 output, [cout][ksize*ksize*cin/2] */
  const int32_t* wzp;    /* [cout] weight zero points; any integer: Scaler.MSE does not force the range to contain 0
                          * (quant/quant_layer.py:38-64), so a channel whose weights share one sign has zp < 0 or > 15 */
  const float* wdelta;   /* [cout] */
  const int32_t* wsum;   /* [cout] */
  const float* bias;     /* [cout] or NULL */
  const float* aq;       /* device (delta_a, zp_a) */
  const float* emb;      /* [n][cout] or NULL */
  int64_t emb_ld;
  const float* res;      /* fp32 NHWC or NULL (may alias out) */
  int64_t res_ld;
  float* out;            /* fp32 NHWC */
  int64_t out_ld;
  int n_stat;            /* 0..2 GroupNorm statistics targets fed by this output */
  tfmq_gn_target stat[2];
} tfmq_conv_w4a8_desc;
int tfmq_conv_w4a8(tfmq_ctx* ctx, const tfmq_conv_w4a8_desc* d, void* stream);

/* ------------------------------------------------------------------------- *
 * fp32-accurate convolution / linear on tcgen05 (kind::tf32, error-compensated
 * 3-pass split) for the layers the reference keeps in floating point
 * (quant/quant_model.py:57-58,103-120: skip/shortcut/op convs, Conv1d qkv and
 * proj_out, first/last layers) and for weight-only-quantised layers
 * (disable_aq).  w_hi/w_lo are [cout][ksize*ksize*cin] fp32 in (tap, cin)
 * order with w_hi = tf32(w), w_lo = w - w_hi (w_lo NULL: weights exact in tf32).
 *   out = wscale[c] * conv(x, w) + bias[c] [+ emb[n][c]] [+ res]
 * passes: 3 = fp32-accurate, 1 = plain tf32.
 * ------------------------------------------------------------------------- */
typedef struct {
  const float* x;       /* fp32 NHWC, no halo (zero padding via TMA OOB fill) */
  int64_t x_ld;
  int n, h, w, cin, cout; /* INPUT spatial extent */
  int ksize, stride;    /* (1|3), (1|2) */
  int pad_lo;           /* leading zero padding (1, or 0 for DDIM's asymmetric downsample) */
  int out_h, out_w;
  const float* w_hi;
  const float* w_lo;    /* or NULL */
  const float* wscale;  /* [cout] or NULL */
  const float* bias;    /* [cout] or NULL */
  const float* res;     /* or NULL (may alias out) */
  int64_t res_ld;
  float* out;
  int64_t out_ld;
  int passes;           /* 1 or 3 */
  const float* emb;     /* [n][cout] per-image add (time embedding) or NULL */
  int64_t emb_ld;
  int n_stat;
  tfmq_gn_target stat[2];
} tfmq_conv_fp_desc;
int tfmq_conv_fp(tfmq_ctx* ctx, const tfmq_conv_fp_desc* d, void* stream);

/* The same layers on kind::f16 (twice the tensor rate and half the operand bytes of kind::tf32): every fp32 operand
 * is split into two fp16 terms, hi = half(v) and lo = half(v - hi) (22 significand bits, like the tf32 hi/lo split),
 * and the product is hi*hi + lo*hi + hi*lo with fp32 accumulation.  Both operands arrive pre-split: the activation
 * planes are written by tfmq_act_prepare (dst_hi / dst_lo), the weight planes once at load time, scaled per output
 * channel by a power of two so that w_lo stays a normal fp16 (the inverse goes into wscale).
 * Range: |x| must stay below 65504 (the split saturates instead of overflowing).
 *   out = wscale[c] * conv(x_hi + x_lo, w_hi + w_lo) + bias[c] [+ emb[n][c]] [+ res] */
typedef struct {
  const void* x_hi;     /* fp16 NHWC, no halo (zero padding via TMA OOB fill) */
  const void* x_lo;
  int64_t x_ld;         /* in halves, multiple of 8 */
  int n, h, w, cin, cout; /* INPUT spatial extent; cin % 16 == 0 */
  int ksize, stride;    /* (1|3), (1|2) */
  int pad_lo;
  int out_h, out_w;
  const void* w_hi;     /* fp16 [cout][ksize*ksize*cin], (tap, cin) order */
  const void* w_lo;     /* or NULL (weights exact in fp16, e.g. integer codes) */
  const float* wscale;  /* [cout] or NULL */
  const float* bias;
  const float* res;
  int64_t res_ld;
  float* out;
  int64_t out_ld;
  const float* emb;
  int64_t emb_ld;
  int n_stat;
  tfmq_gn_target stat[2];
  /* optional: write the result as fp16 hi / lo planes (the split tfmq_act_prepare makes) INSTEAD of fp32 `out`, which may
   * then be NULL: the consumer is another tfmq_conv_h16 or tfmq_attention_h16 (the qkv projection of an attention block).
   * Pixel pitch out_h_ld in halves (multiple of 8).  Not combinable with res / n_stat. */
  void* out_hi;
  void* out_lo;
  int64_t out_h_ld;
  /* optional split-K for reductions over a very long K with few output tiles (the weight-gradient GEMMs of the AdaRound
   * reconstruction loop, quant/reconstruction.py:182-198 `err.backward()`): 0 / 1 = off, > 1 = that many K ranges per output
   * tile, < 0 = chosen by the library.  `out` is cleared and the partial tiles are added by TMA reduce (fp32 adds in arrival
   * order: not bit-reproducible run to run).  Not combinable with res / emb / n_stat / plane output. */
  int ksplit;
} tfmq_conv_h16_desc;
int tfmq_conv_h16(tfmq_ctx* ctx, const tfmq_conv_h16_desc* d, void* stream);

/* first / last convolution with <= 4 input or output channels, fp32 FFMA.
 * in : NCHW [n][cin<=4][h][w]  -> NHWC fp32 (ld)        (conv_in)
 * out: NHWC fp32 (ld)          -> NCHW [n][cout<=4][h][w] (conv_out)
 * weights OIHW fp32 as stored by torch. 3x3, stride 1, pad 1. */
int tfmq_conv_in(tfmq_ctx* ctx, const float* x_nchw, const float* w, const float* bias, int n, int h, int wd, int cin,
                 int cout, float* out, int64_t out_ld, void* stream);
int tfmq_conv_out(tfmq_ctx* ctx, const float* x, int64_t x_ld, const float* w, const float* bias, int n, int h, int wd,
                  int cin, int cout, float* out_nchw, void* stream);

/* ------------------------------------------------------------------------- *
 * Small-M linear (time-embedding MLP / Temporal Information Block,
 * quant/quant_block.py:52-64,101-115): out[m][o] = act_out(sum_i f(x[m][i]) * W[o][i] + b[o])
 *   f = identity | SiLU, optionally followed by u8 fake-quant with `aq`
 *   W = fp32 [o][i]  (w_f32)  or  w4 codes (packed/wzp/wdelta as above)
 * m <= 4096; one warp per output element.
 * ------------------------------------------------------------------------- */
typedef struct {
  const float* x;
  int64_t x_ld;
  int m, in_f, out_f;
  int silu_in;           /* apply SiLU to x first */
  const float* aq;       /* (delta, zp) -> quantise the (SiLU'd) input; NULL = fp input */
  const float* w_f32;    /* fp weights, or NULL */
  const uint8_t* codes;  /* [out_f][in_f] u8 weight codes (when w_f32 == NULL) */
  const float* wzp_f;    /* [out_f] */
  const float* wdelta;   /* [out_f] */
  const float* bias;     /* or NULL */
  float* out;
  int64_t out_ld;
} tfmq_linear_desc;
int tfmq_linear_small(tfmq_ctx* ctx, const tfmq_linear_desc* d, void* stream);

/* Several small-M linears in ONE launch: the per-block embedding projections of a Temporal Information Block
 * (quant/quant_block.py:58-63,108-114 -- a Python loop over `emb_layers`, one SiLU + quantise + F.linear per block) all
 * read the same embedding.  `tfmq_linear_grouped_plan` validates n HOST descriptors (same m) and fills
 * cta_start_host[0..n] (prefix sums of the CTAs each layer gets); the caller uploads the descriptors and the prefix array
 * once and replays `tfmq_linear_grouped` (descs_dev / cta_start_dev are DEVICE pointers; total_ctas = cta_start[n]). */
int tfmq_linear_grouped_plan(tfmq_ctx* ctx, const tfmq_linear_desc* descs_host, int n, int* cta_start_host);
int tfmq_linear_grouped(tfmq_ctx* ctx, const tfmq_linear_desc* descs_dev, const int* cta_start_dev, int n,
                        int total_ctas, int m, int max_in_f, void* stream);

/* sinusoidal timestep embedding.  style 0: DDIM [sin|cos], freq = exp(-ln(1e4) i/(half-1))
 * (ddim/models/diffusion.py:6-24); style 1: LDM [cos|sin], freq = exp(-ln(1e4) i/half)
 * (ldm/modules/diffusionmodules/util.py:151-171). t is fp32 [m]. */
int tfmq_timestep_embedding(tfmq_ctx* ctx, const float* t, int m, int dim, int style, float* out, void* stream);

/* ------------------------------------------------------------------------- *
 * Fused QK^T-softmax-PV attention in fp32 (the attention core is NOT quantised
 * in the reference as shipped: quant/quant_block.py:226,240,318,350,487,496).
 *   o[b,h,i,:] = sum_j softmax_j(scale * q[b,h,i,:].k[b,h,j,:]) v[b,h,j,:]
 * q/k/v/o are addressed as base + b*sb + h*sh + t*st + d (elements).
 * ------------------------------------------------------------------------- */
typedef struct {
  const float* q; int64_t q_sb, q_sh, q_st;
  const float* k; int64_t k_sb, k_sh, k_st;
  const float* v; int64_t v_sb, v_sh, v_st;
  float* o;       int64_t o_sb, o_sh, o_st;
  int b, heads, tq, tk, d;
  float scale;
  /* optional: write o as the fp16 hi / lo planes `tfmq_conv_h16` reads (same split as tfmq_act_prepare's dst_hi / dst_lo),
   * addressed with o_sb / o_sh / o_st, INSTEAD of fp32 `o` (which may then be NULL).  Saves the split launch between an
   * attention core and a floating-point proj_out conv.  Tensor-core kernels only (head dims 32, 40, 64, 80, 160, 256,
   * 384, 512, 576, 960 with 8-byte aligned operands); other shapes return TFMQ_ERR_SHAPE. */
  void* o_hi;
  void* o_lo;
} tfmq_attn_desc;
int tfmq_attention(tfmq_ctx* ctx, const tfmq_attn_desc* d, void* stream);

/* The same attention core on tcgen05 (UMMA, accumulators in tensor memory): S = Q K^T with both operands from shared
 * memory, O += P V with P read from TENSOR MEMORY (written there by the softmax warps, never staged in shared memory)
 * and V read as stored ([key][dim], MN-major).  Replaces the same reference code as tfmq_attention
 * (openaimodel.py:383-405, quant_block.py:212-245,474-505); fp32-accurate: q, k, v arrive PRE-SPLIT as fp16 hi / lo
 * planes (hi = half(x), lo = half(x - hi)) -- written once per tensor by tfmq_conv_h16 (out_hi / out_lo) or
 * tfmq_act_prepare (dst_hi / dst_lo) -- and every product is hi*hi + lo*hi + hi*lo with fp32 accumulation; softmax in
 * fp32.  Planes are addressed like tfmq_attn_desc (base + b*sb + h*sh + t*st + dim, in halves; strides multiples of 8,
 * bases 16-byte aligned).  Head dims: multiples of 8 from 16 to 64 (32 for LDM-4, 40 for SD v1.4's first level).
 * Output: fp32 `o`, or (o_hi, o_lo) planes for a following tfmq_conv_h16. */
typedef struct {
  const void* q_hi; const void* q_lo; int64_t q_sb, q_sh, q_st;
  const void* k_hi; const void* k_lo; int64_t k_sb, k_sh, k_st;
  const void* v_hi; const void* v_lo; int64_t v_sb, v_sh, v_st;
  float* o;       /* or NULL when o_hi / o_lo are given */
  void* o_hi;
  void* o_lo;
  int64_t o_sb, o_sh, o_st;
  int b, heads, tq, tk, d;
  float scale;
} tfmq_attn_h16_desc;
int tfmq_attention_h16(tfmq_ctx* ctx, const tfmq_attn_h16_desc* d, void* stream);

/* ------------------------------------------------------------------------- *
 * DDIM update (ddim/functions/denoising.py:31-37; ldm/models/diffusion/ddim.py:
 * 196-212), elementwise on [count] floats:
 *   x0 = (x - e*sqrt(1-a_t)) / sqrt(a_t);   x_prev = sqrt(a_prev)*x0 + c1*noise + c2*e
 * coef = device float[4] {sqrt(a_t), sqrt(1-a_t), sqrt(a_prev), c2}; c1*noise is
 * added when noise != NULL (eta > 0), and coef is then float[6]: c1 = coef[4], and coef[5] != 0 selects the LDM sampler's
 * summation order (sqrt(a_prev)*x0 + c2*e) + c1*noise (ldm/models/diffusion/ddim.py:205-211; sigma_t * randn there)
 * instead of the DDIM runner's (sqrt(a_prev)*x0 + c1*noise) + c2*e.  x0_out may be NULL.
 * ------------------------------------------------------------------------- */
int tfmq_ddim_update(tfmq_ctx* ctx, const float* x, const float* e, const float* noise, const float* coef,
                     int64_t count, float* x_prev, float* x0_out, void* stream);
/* e = e_u + s*(e_c - e_u) (classifier-free guidance, ldm/models/diffusion/ddim.py:178-185) */
int tfmq_cfg_combine(tfmq_ctx* ctx, const float* e_uncond, const float* e_cond, float s, int64_t count, float* out,
                     void* stream);

/* PLMS sampler (ldm/models/diffusion/plms.py:226-238): the Adams-Bashforth combination of the current noise prediction e0
 * with the stored ones (e1 = old_eps[-1], e2, e3), in the reference's operation order.  order 0: e0; 1: (e0 + e1) / 2 (second
 * half of the first, pseudo improved Euler, step); 2: (3 e0 - e1) / 2; 3: (23 e0 - 16 e1 + 5 e2) / 12;
 * 4: (55 e0 - 59 e1 + 37 e2 - 9 e3) / 24.  The latent update itself is tfmq_ddim_update with the combined prediction. */
int tfmq_plms_eps(tfmq_ctx* ctx, const float* e0, const float* e1, const float* e2, const float* e3, int order,
                  int64_t count, float* out, void* stream);

/* ------------------------------------------------------------------------- *
 * First-stage decode prologue: what happens to the sampled latent before the
 * decoder's conv_in, as ONE launch.
 *   z' = z * inv_scale                                   LatentDiffusion.decode_first_stage,
 *                                                        ldm/models/diffusion/ddpm.py:706-713 (1/scale_factor)
 *   codebook != NULL (VQModelInterface.decode, ldm/models/autoencoder.py:274-279; the quantiser is
 *   taming-transformers' VectorQuantizer2.forward, a dependency the reference does not vendor):
 *       j = argmin_j (|z'|^2 + |e_j|^2 - 2 z'.e_j) (first minimum), zq = z' + (e_j - z')
 *   else zq = z'                                         (AutoencoderKL.decode, :330-333; force_not_quantize)
 *   out[co] = bias[co] + sum_i w[co][i] zq[i]            post_quant_conv (1x1); w == NULL: out = zq
 * z: NCHW fp32 [n][c][hw], out: NCHW fp32 [n][c_out][hw], c, c_out <= 4; codebook [n_embed][c];
 * indices (optional): int32 [n*hw], the chosen code of every latent pixel.
 * ------------------------------------------------------------------------- */
int tfmq_first_stage_input(tfmq_ctx* ctx, const float* z, float inv_scale, const float* codebook, int n_embed,
                           const float* w, const float* bias, int n, int hw, int c, int c_out, float* out,
                           int32_t* indices, void* stream);

/* ------------------------------------------------------------------------- *
 * Calibration primitives.
 * ------------------------------------------------------------------------- */
/* per-row min/max of x[rows][cols] -> mm[rows][2] (rows=1: whole tensor) */
int tfmq_minmax_rows(tfmq_ctx* ctx, const float* x, int64_t rows, int64_t cols, float* mm, void* stream);
/* Scaler.MSE (quant/quant_layer.py:38-64): for each row, 80 shrink candidates,
 * score mean(|dq(x)-x|^2.4), first strict minimum wins.  Writes delta[rows], zp[rows]. */
int tfmq_mse_scale_search(tfmq_ctx* ctx, const float* x, int64_t rows, int64_t cols, int level, float* delta,
                          float* zp, void* stream);
/* act_momentum_update (quant/quant_layer.py:229-244): state = {x_min, x_max} EMA(0.95)
 * with this batch's min/max, then MINMAX -> aq = (delta, zp). All device-side. */
int tfmq_act_range_update(tfmq_ctx* ctx, const float* x, int64_t ld, int64_t pixels, int c, float momentum, int level,
                          float* state, float* aq, void* stream);
/* AdaRound soft weight (quant/adaptive_rounding.py:40-41,59-60,67-70):
 *   w_soft = delta*(clamp(floor(w/delta) + clamp(sigmoid(alpha)*1.2-0.1,0,1) + zp, 0, L-1) - zp) */
int tfmq_adaround_soft(tfmq_ctx* ctx, const float* w, const float* delta, const float* zp, const float* alpha,
                       int cout, int64_t k, int level, float* w_soft, void* stream);
/* One fused AdaRound optimiser step for one weight tensor (quant/reconstruction.py:
 * 182-198 + quant/reconstruction_util.py:36-91): chain dL/dw_soft -> dL/dalpha, add the
 * rounding-regulariser gradient lambda*d/dalpha sum(1-|2h-1|^b), Adam(lr, betas=(.9,.999),
 * eps=1e-8) update of alpha in place, and accumulate the regulariser value into
 * round_loss[0] (caller zeroes).  b <= 0 disables the regulariser (warm-up).
 * step is the 1-based Adam step count. */
int tfmq_adaround_step(tfmq_ctx* ctx, const float* w, const float* delta, const float* zp, float* alpha,
                       const float* grad_w, float* adam_m, float* adam_v, int cout, int64_t k, int level, int step,
                       float lr, float b, float lambda, float* round_loss, void* stream);
/* rec = sum over all elements of |pred-tgt|^2 / batch   (lp_loss p=2, quant_layer.py:146-156:
 * .sum(1).mean(), so `batch` = numel / shape[1]), and grad = 2*(pred-tgt)/batch.
 * loss[0] accumulated (caller zeroes). */
int tfmq_rec_loss(tfmq_ctx* ctx, const float* pred, const float* tgt, int64_t count, int batch, float* loss,
                  float* grad_or_null, void* stream);

/* Measured-peak helper for bench.py: dense u8 x s8 -> s32 GEMM on the same tcgen05
 * pipeline, no unpack / epilogue traffic beyond an s32 store.  a:[m][k] u8, b:[n][k] s8. */
int tfmq_gemm_i8_peak(tfmq_ctx* ctx, const uint8_t* a, const int8_t* b, int m, int n, int k, int32_t* out,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TFMQ_B200_H */
