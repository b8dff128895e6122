// Small FFMA kernels: the time-embedding MLP (M = batch rows) and the first / last
// convolutions whose 3-4 channel side does not fill a tensor-core tile.
#include "ctx.h"

namespace tfmq {

__device__ __forceinline__ float silu1(float v) { return v * (1.f / (1.f + expf(-v))); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Time-embedding MLP layer.  grid (out chunks of 64, row chunks of 8).  Phase 1: the 8 rows' inputs are
// transformed ONCE (SiLU, act fake-quant) into shared memory.  Phase 2: every warp produces 8 outputs; each
// weight is read once (4 consecutive features per lane, all loads of an output issued before use) and applied
// to the 8 rows.
constexpr int LIN_THREADS = 256;
constexpr int LIN_OUT_PER_CTA = 8;   // one output per warp: many small CTAs, short dependent-load chains
constexpr int LIN_ROWS = 8;
constexpr int LIN_GROUP_OUT_PER_CTA = 64;   // grouped launch: the input transform of a CTA (SiLU + exact quantiser: two thirds of the kernel at 32) is shared by 64 outputs

template <int OPC>
__device__ __forceinline__ void linear_small_body(const tfmq_linear_desc& d, const int bx, const int by, float* xs) {
  const int m0 = by * LIN_ROWS;
  const int rows = min(LIN_ROWS, d.m - m0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float dl = 1.f, z = 0.f;
  if (d.aq) dl = d.aq[0], z = d.aq[1];
  // phase 1: float4 loads, four in flight per thread (in_f % 4 == 0; x rows 16-byte aligned when x_ld % 4 == 0)
  const int nv4 = d.in_f >> 2;
  const bool vec = (d.x_ld & 3) == 0 && ((uintptr_t)d.x & 15) == 0;
#pragma unroll 4
  for (int i = threadIdx.x; i < LIN_ROWS * nv4; i += LIN_THREADS) {
    const int r = i / nv4, f4 = i - r * nv4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < rows) {
      const float* src = d.x + (long long)(m0 + r) * d.x_ld + 4 * f4;
      if (vec) v = *reinterpret_cast<const float4*>(src);
      else v = make_float4(src[0], src[1], src[2], src[3]);
      float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (d.silu_in) e[j] = silu1(e[j]);
        if (d.aq) {
          const float q = fminf(fmaxf(rintf(__fdiv_rn(e[j], dl)) + z, 0.f), 255.f);
          e[j] = d.w_f32 ? dl * (q - z) : (q - z);   // integer path keeps the exact (code - zp)
        }
      }
      v = make_float4(e[0], e[1], e[2], e[3]);
    }
    *reinterpret_cast<float4*>(xs + 4 * i) = v;
  }
  __syncthreads();
  const bool int_path = !d.w_f32 && d.aq;
  const int nv = d.in_f >> 2;                   // in_f is a multiple of 4
  for (int oo = 0; oo < OPC / 8; ++oo) {
    const int o = bx * OPC + warp * (OPC / 8) + oo;
    if (o >= d.out_f) break;
    float acc[LIN_ROWS];
#pragma unroll
    for (int r = 0; r < LIN_ROWS; ++r) acc[r] = 0.f;
    const float zw = d.w_f32 ? 0.f : d.wzp_f[o];
    for (int v0 = lane; v0 < nv; v0 += 32 * 4) {
      float4 wv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int v = v0 + 32 * u;
        wv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (v < nv) {
          if (d.w_f32) {
            wv[u] = *reinterpret_cast<const float4*>(d.w_f32 + (long long)o * d.in_f + 4 * v);
          } else {
            const uint32_t c4 = *reinterpret_cast<const uint32_t*>(d.codes + (long long)o * d.in_f + 4 * v);
            wv[u] = make_float4((float)(c4 & 255u) - zw, (float)((c4 >> 8) & 255u) - zw,
                                (float)((c4 >> 16) & 255u) - zw, (float)(c4 >> 24) - zw);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int v = v0 + 32 * u;
        if (v < nv) {
#pragma unroll
          for (int r = 0; r < LIN_ROWS; ++r) {
            const float4 xv = *reinterpret_cast<const float4*>(xs + r * d.in_f + 4 * v);
            acc[r] = fmaf(xv.x, wv[u].x, fmaf(xv.y, wv[u].y, fmaf(xv.z, wv[u].z, fmaf(xv.w, wv[u].w, acc[r]))));
          }
        }
      }
    }
    // integer path: every product and partial sum is an exact integer below 2^24 in magnitude per lane
    // (|code - zp| <= 255, |w| <= 15, in_f / 32 terms), so fp32 accumulation is exact
#pragma unroll
    for (int r = 0; r < LIN_ROWS; ++r) {
      const float t = warp_sum(acc[r]);
      if (lane == 0 && r < rows) {
        float res = t;
        if (!d.w_f32) res *= int_path ? (dl * d.wdelta[o]) : d.wdelta[o];
        if (d.bias) res += d.bias[o];
        d.out[(long long)(m0 + r) * d.out_ld + o] = res;
      }
    }
  }
}

__global__ void __launch_bounds__(LIN_THREADS) linear_small_kernel(const tfmq_linear_desc d) {
  extern __shared__ __align__(16) float xs[];   // [LIN_ROWS][in_f]: fp32 inputs, or (code - zp) as fp32
  linear_small_body<LIN_OUT_PER_CTA>(d, blockIdx.x, blockIdx.y, xs);
}

// Several layers in one launch (the 22 per-block embedding projections of a Temporal Information Block all read the
// same input): CTA x-range [cta_start[l], cta_start[l+1]) belongs to layer l.
__global__ void __launch_bounds__(LIN_THREADS) linear_grouped_kernel(const tfmq_linear_desc* __restrict__ descs,
                                                                     const int* __restrict__ cta_start, int n) {
  extern __shared__ __align__(16) float xs[];
  int l = 0;
  while (l + 1 < n && (int)blockIdx.x >= cta_start[l + 1]) ++l;
  const tfmq_linear_desc d = descs[l];
  linear_small_body<LIN_GROUP_OUT_PER_CTA>(d, blockIdx.x - cta_start[l], blockIdx.y, xs);
}

// conv_in: NCHW (cin<=4) -> NHWC, 3x3 pad 1.  Weights transposed in smem as ws[tap*cin][cout] so a thread reads float4 of 4
// consecutive output channels.  A thread owns FOUR consecutive pixels of an image row x 4 output channels per pass: one
// weight float4 feeds 16 FMAs (with one pixel per thread the kernel was bound by shared-memory bandwidth: 27 LDS.128 per 108
// FMAs, four wavefronts each: 97 us for a 59 MB output), the 3 x 6 input window of the four pixels lives in registers.
constexpr int CIN_PIX = 16;       // pixel quads per CTA pass (x 16 channel groups = 256 threads)
constexpr int CIN_GROUPS = 4;     // passes per CTA: the transposed weight staging is shared by 256 pixels
constexpr int CIN_Q = 4;          // pixels per thread
__global__ void __launch_bounds__(256) conv_in_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ bias, int n, int h, int wd, int cin,
                                                      int cout, float* __restrict__ out, long long out_ld) {
  extern __shared__ float ws[];                 // [cin*9][cout]
  const int K = cin * 9;
  for (int i = threadIdx.x; i < K * cout; i += 256) {
    const int co = i / K, k = i - co * K;       // w is [co][ci][ky][kx] = [co][k]
    ws[k * cout + co] = w[i];
  }
  __syncthreads();
  const int qpr = (wd + CIN_Q - 1) / CIN_Q;                     // pixel quads per image row
  const long long total_q = (long long)n * h * qpr;
  const int cg = threadIdx.x & 15;
  for (int grp = 0; grp < CIN_GROUPS; ++grp) {
    const long long q = ((long long)blockIdx.x * CIN_GROUPS + grp) * CIN_PIX + (threadIdx.x >> 4);
    if (q >= total_q) return;
    const int x0 = (int)(q % qpr) * CIN_Q;
    const int yy = (int)((q / qpr) % h);
    const int nn = (int)(q / ((long long)qpr * h));
    float in[4][3][CIN_Q + 2];                  // [ci][ky][column x0 - 1 ... x0 + 4]
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) {
      const float* xp = x + ((long long)nn * cin + (ci < cin ? ci : 0)) * h * wd;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int y = yy + ky - 1;
#pragma unroll
        for (int j = 0; j < CIN_Q + 2; ++j) {
          const int xq = x0 + j - 1;
          in[ci][ky][j] = (ci < cin && y >= 0 && y < h && xq >= 0 && xq < wd) ? xp[(long long)y * wd + xq] : 0.f;
        }
      }
    }
    const long long pix0 = ((long long)nn * h + yy) * wd + x0;
    for (int co = cg * 4; co < cout; co += 64) {
      const float4 b4 = bias ? *reinterpret_cast<const float4*>(bias + co) : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 acc[CIN_Q];
#pragma unroll
      for (int p = 0; p < CIN_Q; ++p) acc[p] = b4;
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        if (ci < cin) {
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const float4 w4 = *reinterpret_cast<const float4*>(ws + (ci * 9 + ky * 3 + kx) * cout + co);
#pragma unroll
              for (int p = 0; p < CIN_Q; ++p) {
                const float v = in[ci][ky][p + kx];
                acc[p].x = fmaf(v, w4.x, acc[p].x), acc[p].y = fmaf(v, w4.y, acc[p].y);
                acc[p].z = fmaf(v, w4.z, acc[p].z), acc[p].w = fmaf(v, w4.w, acc[p].w);
              }
            }
          }
        }
      }
#pragma unroll
      for (int p = 0; p < CIN_Q; ++p)
        if (x0 + p < wd) *reinterpret_cast<float4*>(out + (pix0 + p) * out_ld + co) = acc[p];
    }
  }
}

// conv_out: NHWC -> NCHW (cout<=4), 3x3 pad 1.  CTA = 8 warps x 8 pixels each; weights in smem as
// float4 per (tap, ci) over the (<=4) output channels; lanes stride over ci (coalesced input rows).
constexpr int COUT_PIX_PER_WARP = 8;
__global__ void __launch_bounds__(256) conv_out_kernel(const float* __restrict__ x, long long x_ld,
                                                       const float* __restrict__ w, const float* __restrict__ bias,
                                                       int n, int h, int wd, int cin, int cout,
                                                       float* __restrict__ out) {
  extern __shared__ float4 w4s[];               // [9][cin]
  for (int i = threadIdx.x; i < 9 * cin; i += 256) {
    const int tap = i / cin, ci = i - tap * cin;
    float t[4] = {0.f, 0.f, 0.f, 0.f};
    for (int co = 0; co < cout; ++co) t[co] = w[((long long)co * cin + ci) * 9 + tap];
    w4s[i] = make_float4(t[0], t[1], t[2], t[3]);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long total = (long long)n * h * wd;
  const long long p0 = (blockIdx.x * 8LL + warp) * COUT_PIX_PER_WARP;
  for (int pp = 0; pp < COUT_PIX_PER_WARP; ++pp) {
    const long long pix = p0 + pp;
    if (pix >= total) break;
    const int xx = (int)(pix % wd);
    const int yy = (int)((pix / wd) % h);
    const int nn = (int)(pix / ((long long)wd * h));
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int y = yy + ky - 1;
      if (y < 0 || y >= h) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xq = xx + kx - 1;
        if (xq < 0 || xq >= wd) continue;
        const float* xp = x + (((long long)nn * h + y) * wd + xq) * x_ld;
        const float4* wp = w4s + (ky * 3 + kx) * cin;
        for (int ci = lane; ci < cin; ci += 32) {
          const float v = xp[ci];
          const float4 w4 = wp[ci];
          acc.x = fmaf(v, w4.x, acc.x), acc.y = fmaf(v, w4.y, acc.y);
          acc.z = fmaf(v, w4.z, acc.z), acc.w = fmaf(v, w4.w, acc.w);
        }
      }
    }
    const float r0 = warp_sum(acc.x), r1 = warp_sum(acc.y), r2 = warp_sum(acc.z), r3 = warp_sum(acc.w);
    const float rl = lane == 0 ? r0 : lane == 1 ? r1 : lane == 2 ? r2 : r3;
    if (lane < cout) out[(((long long)nn * cout + lane) * h + yy) * wd + xx] = rl + (bias ? bias[lane] : 0.f);
  }
}

// conv_out for widths that are multiples of 8: a warp owns 8 consecutive pixels of one image row.  Per (input row,
// channel) it loads the 10 input columns it needs ONCE and each weight float4 once, and applies them to all 8 pixels
// (the pixel-at-a-time kernel above re-reads every input 9 times and every weight once per pixel).
__global__ void __launch_bounds__(256) conv_out_row8_kernel(const float* __restrict__ x, long long x_ld,
                                                            const float* __restrict__ w, const float* __restrict__ bias,
                                                            int n, int h, int wd, int cin, int cout,
                                                            float* __restrict__ out) {
  extern __shared__ float4 w4s[];               // [9][cin]
  for (int i = threadIdx.x; i < 9 * cin; i += 256) {
    const int tap = i / cin, ci = i - tap * cin;
    float t[4] = {0.f, 0.f, 0.f, 0.f};
    for (int co = 0; co < cout; ++co) t[co] = w[((long long)co * cin + ci) * 9 + tap];
    w4s[i] = make_float4(t[0], t[1], t[2], t[3]);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long total8 = (long long)n * h * (wd >> 3);          // groups of 8 pixels
  const long long grp = blockIdx.x * 8LL + warp;
  if (grp >= total8) return;
  const int gx = (int)(grp % (wd >> 3));
  const int yy = (int)((grp / (wd >> 3)) % h);
  const int nn = (int)(grp / ((long long)(wd >> 3) * h));
  const int x0 = gx * 8;
  float4 acc[8];
#pragma unroll
  for (int p = 0; p < 8; ++p) acc[p] = make_float4(0.f, 0.f, 0.f, 0.f);
  // (addressing in 32-bit offsets from one 64-bit row pointer, bounds tests only on the two edge columns: the version with a
  //  64-bit multiply and two compares per load executed ~410 instructions per (row, channel) step for 96 FMAs)
  const int ld = (int)x_ld;
  const bool left = x0 > 0, right = x0 + 8 < wd;
  for (int ky = 0; ky < 3; ++ky) {
    const int y = yy + ky - 1;
    if (y < 0 || y >= h) continue;
    const float* row = x + (((long long)nn * h + y) * wd + x0) * x_ld;      // column x0 of the input row
    for (int ci = lane; ci < cin; ci += 32) {
      const float* pc = row + ci;
      float v[10];
      v[0] = left ? pc[-ld] : 0.f;
#pragma unroll
      for (int j = 1; j < 9; ++j) v[j] = pc[(j - 1) * ld];
      v[9] = right ? pc[8 * ld] : 0.f;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const float4 w4 = w4s[(ky * 3 + kx) * cin + ci];
#pragma unroll
        for (int p = 0; p < 8; ++p) {
          const float a = v[p + kx];
          acc[p].x = fmaf(a, w4.x, acc[p].x), acc[p].y = fmaf(a, w4.y, acc[p].y);
          acc[p].z = fmaf(a, w4.z, acc[p].z), acc[p].w = fmaf(a, w4.w, acc[p].w);
        }
      }
    }
  }
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    const float r0 = warp_sum(acc[p].x), r1 = warp_sum(acc[p].y), r2 = warp_sum(acc[p].z), r3 = warp_sum(acc[p].w);
    const float rl = lane == 0 ? r0 : lane == 1 ? r1 : lane == 2 ? r2 : r3;
    if (lane < cout) out[(((long long)nn * cout + lane) * h + yy) * wd + x0 + p] = rl + (bias ? bias[lane] : 0.f);
  }
}

// First-stage decode prologue: latent scaling, optional nearest-codebook lookup with the straight-through arithmetic of
// the reference's quantiser, and the 1x1 post_quant_conv, NCHW -> NCHW with <= 4 channels.  One thread per latent pixel; the
// codebook is staged through shared memory in chunks of FS_CHUNK codes as (e0, e1, e2, e3) + |e|^2.
constexpr int FS_CHUNK = 1024;
__global__ void __launch_bounds__(128) first_stage_input_kernel(const float* __restrict__ z, float inv_scale,
                                                                const float* __restrict__ codebook, int n_embed,
                                                                const float* __restrict__ w, const float* __restrict__ bias,
                                                                long long total, int hw, int c, int c_out,
                                                                float* __restrict__ out, int* __restrict__ indices) {
  __shared__ float4 es[FS_CHUNK];
  __shared__ float ee[FS_CHUNK];
  const long long pix = (long long)blockIdx.x * 128 + threadIdx.x;
  const bool valid = pix < total;
  const long long nn = valid ? pix / hw : 0;
  const int p = valid ? (int)(pix - nn * hw) : 0;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (valid)
    for (int i = 0; i < c; ++i) v[i] = z[(nn * c + i) * hw + p] * inv_scale;
  if (codebook) {
    float zz = 0.f;
    for (int i = 0; i < c; ++i) zz += v[i] * v[i];
    float best = INFINITY;
    int best_j = 0;
    for (int j0 = 0; j0 < n_embed; j0 += FS_CHUNK) {
      __syncthreads();
      for (int j = threadIdx.x; j < FS_CHUNK; j += 128) {
        float e[4] = {0.f, 0.f, 0.f, 0.f};
        float s = 0.f;
        if (j0 + j < n_embed)
          for (int i = 0; i < c; ++i) {
            e[i] = codebook[(long long)(j0 + j) * c + i];
            s += e[i] * e[i];
          }
        es[j] = make_float4(e[0], e[1], e[2], e[3]);
        ee[j] = s;
      }
      __syncthreads();
      const int cnt = min(FS_CHUNK, n_embed - j0);
      for (int j = 0; j < cnt; ++j) {
        const float4 e = es[j];
        float dot = v[0] * e.x;
        dot = fmaf(v[1], e.y, dot), dot = fmaf(v[2], e.z, dot), dot = fmaf(v[3], e.w, dot);
        const float d = (zz + ee[j]) - 2.f * dot;
        if (d < best) best = d, best_j = j0 + j;      // strict <: the first minimum, as torch.argmin
      }
    }
    if (valid) {
      if (indices) indices[pix] = best_j;
      for (int i = 0; i < c; ++i) {
        const float e = codebook[(long long)best_j * c + i];
        v[i] = v[i] + (e - v[i]);                       // z + (z_q - z).detach()
      }
    }
  }
  if (!valid) return;
  for (int co = 0; co < c_out; ++co) {
    float acc = 0.f;
    if (w) {
      for (int i = 0; i < c; ++i) acc = fmaf(v[i], w[co * c + i], acc);
      if (bias) acc += bias[co];
    } else {
      acc = v[co];
    }
    out[(nn * c_out + co) * hw + p] = acc;
  }
}

}  // namespace tfmq

using namespace tfmq;

static int check_linear(tfmq_ctx* ctx, const tfmq_linear_desc* d) {
  TFMQ_REQUIRE(d && d->x && d->out, TFMQ_ERR_ARG, "linear_small: null pointer");
  TFMQ_REQUIRE(d->w_f32 || (d->codes && d->wzp_f && d->wdelta), TFMQ_ERR_ARG, "linear_small: weights missing");
  TFMQ_REQUIRE(d->m >= 0 && d->m <= 4096 && d->in_f > 0 && d->out_f > 0, TFMQ_ERR_SHAPE, "linear_small: m=%d", d->m);
  TFMQ_REQUIRE(d->in_f <= 1536 && d->in_f % 4 == 0, TFMQ_ERR_SHAPE, "linear_small: in_f %d (multiple of 4, <= 1536)",
               d->in_f);
  TFMQ_REQUIRE(d->w_f32 ? ((uintptr_t)d->w_f32 & 15) == 0 : ((uintptr_t)d->codes & 3) == 0, TFMQ_ERR_ARG,
               "linear_small: weights must be 16-byte (fp32) / 4-byte (codes) aligned");
  return TFMQ_OK;
}

extern "C" int tfmq_linear_small(tfmq_ctx* ctx, const tfmq_linear_desc* d, void* stream) {
  if (!ctx) return TFMQ_ERR_ARG;
  if (int rc = check_linear(ctx, d)) return rc;
  if (d->m == 0) return TFMQ_OK;
  dim3 grid((d->out_f + LIN_OUT_PER_CTA - 1) / LIN_OUT_PER_CTA, (d->m + LIN_ROWS - 1) / LIN_ROWS);
  linear_small_kernel<<<grid, LIN_THREADS, (size_t)LIN_ROWS * d->in_f * sizeof(float), tfmq_stream(stream)>>>(*d);
  TFMQ_LAUNCH_CHECK("linear_small");
  return TFMQ_OK;
}

extern "C" int tfmq_linear_grouped_plan(tfmq_ctx* ctx, const tfmq_linear_desc* descs_host, int n, int* cta_start_host) {
  if (!ctx) return TFMQ_ERR_ARG;
  TFMQ_REQUIRE(descs_host && cta_start_host && n >= 1 && n <= 256, TFMQ_ERR_ARG, "linear_grouped_plan: n=%d", n);
  int acc = 0;
  for (int l = 0; l < n; ++l) {
    if (int rc = check_linear(ctx, &descs_host[l])) return rc;
    TFMQ_REQUIRE(descs_host[l].m == descs_host[0].m, TFMQ_ERR_SHAPE, "linear_grouped_plan: layers differ in m");
    cta_start_host[l] = acc;
    acc += (descs_host[l].out_f + LIN_GROUP_OUT_PER_CTA - 1) / LIN_GROUP_OUT_PER_CTA;
  }
  cta_start_host[n] = acc;
  return TFMQ_OK;
}

extern "C" int tfmq_linear_grouped(tfmq_ctx* ctx, const tfmq_linear_desc* descs_dev, const int* cta_start_dev, int n,
                                   int total_ctas, int m, int max_in_f, void* stream) {
  if (!ctx) return TFMQ_ERR_ARG;
  TFMQ_REQUIRE(descs_dev && cta_start_dev && n >= 1 && n <= 256 && total_ctas >= 1, TFMQ_ERR_ARG, "linear_grouped: args");
  TFMQ_REQUIRE(m >= 0 && m <= 4096 && max_in_f > 0 && max_in_f <= 1536 && max_in_f % 4 == 0, TFMQ_ERR_SHAPE,
               "linear_grouped: m=%d max_in_f=%d", m, max_in_f);
  if (m == 0) return TFMQ_OK;
  dim3 grid(total_ctas, (m + LIN_ROWS - 1) / LIN_ROWS);
  linear_grouped_kernel<<<grid, LIN_THREADS, (size_t)LIN_ROWS * max_in_f * sizeof(float), tfmq_stream(stream)>>>(
      descs_dev, cta_start_dev, n);
  TFMQ_LAUNCH_CHECK("linear_grouped");
  return TFMQ_OK;
}

extern "C" int tfmq_conv_in(tfmq_ctx* ctx, const float* x_nchw, const float* w, const float* bias, int n, int h,
                            int wd, int cin, int cout, float* out, int64_t out_ld, void* stream) {
  if (!ctx) return TFMQ_ERR_ARG;
  TFMQ_REQUIRE(x_nchw && w && out, TFMQ_ERR_ARG, "conv_in: null pointer");
  TFMQ_REQUIRE(cin >= 1 && cin <= 4, TFMQ_ERR_SHAPE, "conv_in: cin %d > 4", cin);
  const size_t smem_in = (size_t)cout * cin * 9 * sizeof(float);
  TFMQ_REQUIRE(cout % 4 == 0 && out_ld % 4 == 0 && smem_in <= 160 * 1024, TFMQ_ERR_SHAPE, "conv_in: cout %d", cout);
  static size_t smem_in_set = 48 * 1024;
  if (smem_in > smem_in_set) {      // the first-stage decoder's conv_in (z -> 512 channels) needs 54 / 72 KB
    cudaError_t e = cudaFuncSetAttribute(conv_in_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_in);
    if (e != cudaSuccess) return tfmq_fail(ctx, TFMQ_ERR_CUDA, "conv_in: smem attr: %s", cudaGetErrorString(e));
    smem_in_set = smem_in;
  }
  TFMQ_REQUIRE(((uintptr_t)out & 15) == 0 && (!bias || ((uintptr_t)bias & 15) == 0), TFMQ_ERR_ARG,
               "conv_in: out / bias must be 16-byte aligned");
  const long long total = (long long)n * h * wd;
  if (total == 0) return TFMQ_OK;
  const long long total_q = (long long)n * h * ((wd + CIN_Q - 1) / CIN_Q);       // pixel quads
  conv_in_kernel<<<(unsigned)((total_q + CIN_PIX * CIN_GROUPS - 1) / (CIN_PIX * CIN_GROUPS)), 256, smem_in,
                   tfmq_stream(stream)>>>(x_nchw, w, bias, n, h, wd, cin, cout, out, out_ld);
  TFMQ_LAUNCH_CHECK("conv_in");
  return TFMQ_OK;
}

extern "C" int tfmq_conv_out(tfmq_ctx* ctx, const float* x, int64_t x_ld, const float* w, const float* bias, int n,
                             int h, int wd, int cin, int cout, float* out_nchw, void* stream) {
  if (!ctx) return TFMQ_ERR_ARG;
  TFMQ_REQUIRE(x && w && out_nchw, TFMQ_ERR_ARG, "conv_out: null pointer");
  TFMQ_REQUIRE(cout >= 1 && cout <= 4, TFMQ_ERR_SHAPE, "conv_out: cout %d > 4", cout);
  TFMQ_REQUIRE(9 * cin * 16 <= 48 * 1024, TFMQ_ERR_SHAPE, "conv_out: cin %d too large", cin);
  const long long total = (long long)n * h * wd;
  if (total == 0) return TFMQ_OK;
  if (wd % 8 == 0) {
    const long long groups = total / 8;
    conv_out_row8_kernel<<<(unsigned)((groups + 7) / 8), 256, (size_t)9 * cin * sizeof(float4), tfmq_stream(stream)>>>(
        x, x_ld, w, bias, n, h, wd, cin, cout, out_nchw);
    TFMQ_LAUNCH_CHECK("conv_out");
    return TFMQ_OK;
  }
  const long long per_cta = 8LL * COUT_PIX_PER_WARP;
  conv_out_kernel<<<(unsigned)((total + per_cta - 1) / per_cta), 256, (size_t)9 * cin * sizeof(float4),
                    tfmq_stream(stream)>>>(x, x_ld, w, bias, n, h, wd, cin, cout, out_nchw);
  TFMQ_LAUNCH_CHECK("conv_out");
  return TFMQ_OK;
}

extern "C" int tfmq_first_stage_input(tfmq_ctx* ctx, const float* z, float inv_scale, const float* codebook, int n_embed,
                                      const float* w, const float* bias, int n, int hw, int c, int c_out, float* out,
                                      int32_t* indices, void* stream) {
  if (!ctx) return TFMQ_ERR_ARG;
  TFMQ_REQUIRE(z && out, TFMQ_ERR_ARG, "first_stage_input: null pointer");
  TFMQ_REQUIRE(c >= 1 && c <= 4 && c_out >= 1 && c_out <= 4, TFMQ_ERR_SHAPE, "first_stage_input: channels %d -> %d (<= 4)", c,
               c_out);
  TFMQ_REQUIRE(w || c_out == c, TFMQ_ERR_SHAPE, "first_stage_input: no post_quant_conv but %d -> %d channels", c, c_out);
  TFMQ_REQUIRE(!codebook || n_embed >= 1, TFMQ_ERR_SHAPE, "first_stage_input: empty codebook");
  const long long total = (long long)n * hw;
  if (total == 0) return TFMQ_OK;
  first_stage_input_kernel<<<(unsigned)((total + 127) / 128), 128, 0, tfmq_stream(stream)>>>(
      z, inv_scale, codebook, n_embed, w, bias, total, hw, c, c_out, out, indices);
  TFMQ_LAUNCH_CHECK("first_stage_input");
  return TFMQ_OK;
}
