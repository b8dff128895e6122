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

// one warp per output element (m, o)
__global__ void __launch_bounds__(256) linear_small_kernel(const tfmq_linear_desc d) {
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= d.m * d.out_f) return;
  const int m = wid / d.out_f, o = wid - m * d.out_f;
  const float* x = d.x + (long long)m * d.x_ld;
  float r;
  if (d.w_f32) {
    const float* w = d.w_f32 + (long long)o * d.in_f;
    float acc = 0.f;
    for (int i = lane; i < d.in_f; i += 32) {
      float v = x[i];
      if (d.silu_in) v = silu1(v);
      if (d.aq) {
        const float dl = d.aq[0], z = d.aq[1];
        const float q = fminf(fmaxf(rintf(__fdiv_rn(v, dl)) + z, 0.f), 255.f);
        v = dl * (q - z);
      }
      acc = fmaf(v, w[i], acc);
    }
    r = warp_sum(acc);
  } else {
    const uint8_t* cw = d.codes + (long long)o * d.in_f;
    const int zw = (int)d.wzp_f[o];
    if (d.aq) {
      const float dl = d.aq[0], z = d.aq[1];
      const int za = (int)z;
      int acc = 0;
      for (int i = lane; i < d.in_f; i += 32) {
        float v = x[i];
        if (d.silu_in) v = silu1(v);
        const int qa = (int)fminf(fmaxf(rintf(__fdiv_rn(v, dl)) + z, 0.f), 255.f);
        acc += (qa - za) * ((int)cw[i] - zw);
      }
      r = (float)warp_sum_i(acc) * (dl * d.wdelta[o]);
    } else {
      float acc = 0.f;
      for (int i = lane; i < d.in_f; i += 32) {
        float v = x[i];
        if (d.silu_in) v = silu1(v);
        acc = fmaf(v, (float)((int)cw[i] - zw), acc);
      }
      r = warp_sum(acc) * d.wdelta[o];
    }
  }
  if (lane == 0) {
    if (d.bias) r += d.bias[o];
    d.out[(long long)m * d.out_ld + o] = r;
  }
}

// conv_in: NCHW (cin<=4) -> NHWC, 3x3 pad 1.  thread per (pixel, cout)
__global__ void __launch_bounds__(256) conv_in_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ bias, int n, int h, int wd, int cin,
                                                      int cout, float* __restrict__ out, long long out_ld) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long total = (long long)n * h * wd * cout;
  if (idx >= total) return;
  const int co = (int)(idx % cout);
  long long pix = idx / cout;
  const int xx = (int)(pix % wd);
  const int yy = (int)((pix / wd) % h);
  const int nn = (int)(pix / ((long long)wd * h));
  float acc = bias ? bias[co] : 0.f;
  for (int ci = 0; ci < cin; ++ci) {
    const float* xp = x + ((long long)nn * cin + ci) * h * wd;
    const float* wp = w + ((long long)co * cin + ci) * 9;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int y = yy + ky - 1;
      if (y < 0 || y >= h) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xq = xx + kx - 1;
        if (xq < 0 || xq >= wd) continue;
        acc = fmaf(xp[(long long)y * wd + xq], wp[ky * 3 + kx], acc);
      }
    }
  }
  out[pix * out_ld + co] = acc;
}

// conv_out: NHWC -> NCHW (cout<=4), 3x3 pad 1.  warp per output pixel, lanes over cin
__global__ void __launch_bounds__(256) conv_out_kernel(const float* __restrict__ x, long long x_ld,
                                                       const float* __restrict__ w, const float* __restrict__ bias,
                                                       int n, int h, int wd, int cin, int cout,
                                                       float* __restrict__ out) {
  const long long wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long total = (long long)n * h * wd;
  if (wid >= total) return;
  const int xx = (int)(wid % wd);
  const int yy = (int)((wid / wd) % h);
  const int nn = (int)(wid / ((long long)wd * h));
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int ky = 0; ky < 3; ++ky) {
    const int y = yy + ky - 1;
    if (y < 0 || y >= h) continue;
    for (int kx = 0; kx < 3; ++kx) {
      const int xq = xx + kx - 1;
      if (xq < 0 || xq >= wd) continue;
      const float* xp = x + (((long long)nn * h + y) * wd + xq) * x_ld;
      for (int ci = lane; ci < cin; ci += 32) {
        const float v = xp[ci];
#pragma unroll
        for (int co = 0; co < 4; ++co)
          if (co < cout) acc[co] = fmaf(v, w[(((long long)co * cin + ci) * 3 + ky) * 3 + kx], acc[co]);
      }
    }
  }
#pragma unroll
  for (int co = 0; co < 4; ++co) {
    const float r = warp_sum(acc[co]);
    if (lane == 0 && co < cout) out[(((long long)nn * cout + co) * h + yy) * wd + xx] = r + (bias ? bias[co] : 0.f);
  }
}

}  // namespace tfmq

using namespace tfmq;

extern "C" int tfmq_linear_small(tfmq_ctx* ctx, const tfmq_linear_desc* d, void* stream) {
  if (!ctx) return TFMQ_ERR_ARG;
  TFMQ_REQUIRE(d && d->x && d->out, TFMQ_ERR_ARG, "linear_small: null pointer");
  TFMQ_REQUIRE(d->w_f32 || (d->codes && d->wzp_f && d->wdelta), TFMQ_ERR_ARG, "linear_small: weights missing");
  TFMQ_REQUIRE(d->m >= 0 && d->m <= 4096 && d->in_f > 0 && d->out_f > 0, TFMQ_ERR_SHAPE, "linear_small: m=%d", d->m);
  if (d->m == 0) return TFMQ_OK;
  const long long warps = (long long)d->m * d->out_f;
  const int blocks = (int)((warps * 32 + 255) / 256);
  linear_small_kernel<<<blocks, 256, 0, tfmq_stream(stream)>>>(*d);
  TFMQ_LAUNCH_CHECK("linear_small");
  return TFMQ_OK;
}

extern "C" int tfmq_conv_in(tfmq_ctx* ctx, const float* x_nchw, const float* w, const float* bias, int n, int h,
                            int wd, int cin, int cout, float* out, int64_t out_ld, void* stream) {
  if (!ctx) return TFMQ_ERR_ARG;
  TFMQ_REQUIRE(x_nchw && w && out, TFMQ_ERR_ARG, "conv_in: null pointer");
  TFMQ_REQUIRE(cin >= 1 && cin <= 4, TFMQ_ERR_SHAPE, "conv_in: cin %d > 4", cin);
  const long long total = (long long)n * h * wd * cout;
  if (total == 0) return TFMQ_OK;
  conv_in_kernel<<<(unsigned)((total + 255) / 256), 256, 0, tfmq_stream(stream)>>>(x_nchw, w, bias, n, h, wd, cin,
                                                                                  cout, out, out_ld);
  TFMQ_LAUNCH_CHECK("conv_in");
  return TFMQ_OK;
}

extern "C" int tfmq_conv_out(tfmq_ctx* ctx, const float* x, int64_t x_ld, const float* w, const float* bias, int n,
                             int h, int wd, int cin, int cout, float* out_nchw, void* stream) {
  if (!ctx) return TFMQ_ERR_ARG;
  TFMQ_REQUIRE(x && w && out_nchw, TFMQ_ERR_ARG, "conv_out: null pointer");
  TFMQ_REQUIRE(cout >= 1 && cout <= 4, TFMQ_ERR_SHAPE, "conv_out: cout %d > 4", cout);
  const long long total = (long long)n * h * wd;
  if (total == 0) return TFMQ_OK;
  conv_out_kernel<<<(unsigned)((total * 32 + 255) / 256), 256, 0, tfmq_stream(stream)>>>(x, x_ld, w, bias, n, h, wd,
                                                                                        cin, cout, out_nchw);
  TFMQ_LAUNCH_CHECK("conv_out");
  return TFMQ_OK;
}
