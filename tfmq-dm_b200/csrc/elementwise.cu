// HBM-bound producer / consumer kernels around the tcgen05 GEMMs:
// GroupNorm statistics, the fused GN-apply + SiLU + activation-quantise producer,
// DDIM update, classifier-free-guidance combine, timestep embedding.
#include <cstdlib>

#include <cuda_fp16.h>

#include "ctx.h"
#include "ptx.cuh"

namespace tfmq {

// ---------------------------------------------------------------- GN statistics
// grid (chunks, n); each CTA reduces a run of pixels, threads own float4 channel
// vectors (coalesced rows), fp32 partials over <= a few hundred pixels, then
// double per channel -> per group -> one atomicAdd(double) pair per group.
constexpr int GN_THREADS = 256;
constexpr int GN_MAXVEC = 2;  // c <= 2048

// When c / 4 <= 256 the threads form a (channel vector tv, pixel lane tp) grid of tvn x k: with one pixel row per CTA pass a
// 224-channel tensor kept 56 of the 256 threads busy, one load in flight each (64 us for 59 MB); the k pixel lanes' fp32
// partials meet in shared memory and are added in a fixed order.
__global__ void __launch_bounds__(GN_THREADS) gn_stats_kernel(const float* __restrict__ x, long long ld, int hw, int c,
                                                              int groups, int pix_per_cta,
                                                              double* __restrict__ stats, int cpg, int ch_off, int tvn, int k) {
  extern __shared__ double sm[];  // [c] sum, [c] sumsq
  __shared__ float part[GN_THREADS * 8];   // [tp][sum x4 | sumsq x4][tv]
  const int n = blockIdx.y;
  const int p0 = blockIdx.x * pix_per_cta;
  int p1 = p0 + pix_per_cta;
  if (p1 > hw) p1 = hw;
  const int nvec = c >> 2;
  const int tv = threadIdx.x % tvn, tp = threadIdx.x / tvn;
  float s[GN_MAXVEC][4], ss[GN_MAXVEC][4];
#pragma unroll
  for (int v = 0; v < GN_MAXVEC; ++v)
#pragma unroll
    for (int j = 0; j < 4; ++j) s[v][j] = ss[v][j] = 0.f;
  const float* base = x + ((long long)n * hw) * ld;
  if (tp < k) {
    for (int p = p0 + tp; p < p1; p += k) {
      const float4* row = reinterpret_cast<const float4*>(base + (long long)p * ld);
#pragma unroll
      for (int v = 0; v < GN_MAXVEC; ++v) {
        const int iv = tv + v * GN_THREADS;
        if (iv < nvec) {
          const float4 f = row[iv];
          s[v][0] += f.x, s[v][1] += f.y, s[v][2] += f.z, s[v][3] += f.w;
          ss[v][0] += f.x * f.x, ss[v][1] += f.y * f.y, ss[v][2] += f.z * f.z, ss[v][3] += f.w * f.w;
        }
      }
    }
  }
  if (k == 1) {
#pragma unroll
    for (int v = 0; v < GN_MAXVEC; ++v) {
      const int iv = threadIdx.x + v * GN_THREADS;
      if (iv < nvec) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          sm[iv * 4 + j] = (double)s[v][j];
          sm[c + iv * 4 + j] = (double)ss[v][j];
        }
      }
    }
  } else {
    if (tp < k) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        part[(tp * 8 + j) * tvn + tv] = s[0][j];
        part[(tp * 8 + 4 + j) * tvn + tv] = ss[0][j];
      }
    }
    __syncthreads();
    for (int ch = threadIdx.x; ch < c; ch += GN_THREADS) {
      const int iv = ch >> 2, j = ch & 3;
      double a = 0.0, b = 0.0;
      for (int t = 0; t < k; ++t) {
        a += (double)part[(t * 8 + j) * tvn + iv];
        b += (double)part[(t * 8 + 4 + j) * tvn + iv];
      }
      sm[ch] = a;
      sm[c + ch] = b;
    }
  }
  __syncthreads();
  // channel ch of this tensor lies in group (ch_off + ch) / cpg of the normalised tensor
  const int g_lo = ch_off / cpg, g_hi = (ch_off + c - 1) / cpg;
  for (int g = g_lo + threadIdx.x; g <= g_hi; g += GN_THREADS) {
    const int lo = max(g * cpg - ch_off, 0), hi = min((g + 1) * cpg - ch_off, c);
    double a = 0, b = 0;
    for (int j = lo; j < hi; ++j) {
      a += sm[j];
      b += sm[c + j];
    }
    atomicAdd(&stats[((long long)n * groups + g) * 2], a);
    atomicAdd(&stats[((long long)n * groups + g) * 2 + 1], b);
  }
}

// ------------------------------------------------- GN-apply + SiLU + quantise
constexpr int ACT_THREADS = 256;

struct ActParams {
  tfmq_act_desc d;
  int out_h, out_w;     // destination interior extent
  int pix_per_cta;      // destination pixels (incl. halo) per CTA
  double inv_cnt;       // GroupNorm: 1 / (channels per group * h * w)
};

// x*sigmoid(x); __expf / __frcp_rn keep the result within ~2 ulp of the accurate form, far below the
// activation quantisation step, at a fraction of the instruction count
__device__ __forceinline__ float rcp_approx(float v) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ float silu_f(float v) { return v * rcp_approx(1.f + __expf(-v)); }

// code = clamp(rint(RN(t / delta)) + zp, 0, 255), bit-identical to the reference's fp32 divide + round,
// without paying for an IEEE division on every element: rint(t * (1/delta)) can differ from rint(RN(t/delta))
// only when t/delta lies within a few ulps of a rounding boundary; the exact remainder (one FMA) detects that
// case and only then the correctly rounded division is evaluated.
// Rounding and the float -> integer move use the 1.5*2^23 / 2^23 add tricks (full-rate FADD, round-half-even
// like rintf) instead of FRND / F2I, which share the quarter-rate pipe with the SiLU's EX2 / RCP.
// This version folds zero point and clamp in before the rounding: r = clamp(fma(t, 1/delta, zp), 0, 255) differs from
// RN(t / delta) + zp by < 1e-4 inside the code range (and both saturate outside it), the 1.5*2^23 add rounds it half-even
// and leaves the code in the low mantissa byte; 7 full-rate instructions per element instead of 14.
__device__ __noinline__ float quant_slow(float t, float delta, float zp) {
  const float code = fminf(fmaxf(rintf(__fdiv_rn(t, delta)) + zp, 0.f), 255.f);
  return __fadd_rn(code, 12582912.f);
}
// returns a word whose LOW BYTE is the code
__device__ __forceinline__ uint32_t quant1(float t, float delta, float inv, float zp) {
  const float r = fminf(fmaxf(fmaf(t, inv, zp), 0.f), 255.f);
  float m = __fadd_rn(r, 12582912.f);                    // 1.5 * 2^23: the sum's low mantissa bits are rint(r)
  const float nr = __fsub_rn(m, 12582912.f);
  // r is within 1e-4 of the reference's pre-rounding value: the two roundings can only differ when r is within that of a
  // half-integer; 5e-4 of margin sends 0.1 % of the elements to the exact path
  if (fabsf(r - nr) > 0.4995f) m = quant_slow(t, delta, zp);
  return __float_as_uint(m);
}
__device__ __forceinline__ uint32_t quant4(const float (&t)[4], float delta, float inv, float zp) {
  const uint32_t lo = __byte_perm(quant1(t[0], delta, inv, zp), quant1(t[1], delta, inv, zp), 0x0040);
  const uint32_t hi = __byte_perm(quant1(t[2], delta, inv, zp), quant1(t[3], delta, inv, zp), 0x0040);
  return __byte_perm(lo, hi, 0x5410);
}

// v = hi + lo (+ O(2^-22 |v|)): hi = half(v), lo = half(v - hi); saturating, so an out-of-range value cannot become inf
__device__ __forceinline__ uint32_t half_sat_bits(float v) {
  uint16_t h;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(v));
  return h;
}
__device__ __forceinline__ void split_h16x4(const float (&t)[4], uint2& hi, uint2& lo) {
  uint32_t hb[4], lb[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    hb[j] = half_sat_bits(t[j]);
    lb[j] = half_sat_bits(t[j] - __half2float(__ushort_as_half((unsigned short)hb[j])));
  }
  hi = make_uint2(hb[0] | (hb[1] << 16), hb[2] | (hb[3] << 16));
  lo = make_uint2(lb[0] | (lb[1] << 16), lb[2] | (lb[3] << 16));
}

__device__ __forceinline__ float gelu_f(float g) { return 0.5f * g * (1.f + erff(g * 0.70710678118654752440f)); }

// One warp per destination pixel (lanes = float4 channel vectors): the pixel decode / border test is per
// warp, not per element, and every global access is a full row.
// Specialised at compile time on (normalisation, SiLU, GEGLU, output kind): the kernel is instruction-issue bound
// (ncu: issue slots 70 % busy at 30 % of the DRAM rate), so per-element mode tests are not free.
enum { ACT_NORM_NONE = 0, ACT_NORM_GN = 1, ACT_NORM_LN = 2 };
enum { ACT_OUT_U8 = 0, ACT_OUT_F32 = 1, ACT_OUT_H16 = 2 };
template <int NORM, bool SILU, bool GEGLU, int OUT>
__global__ void __launch_bounds__(ACT_THREADS) act_prepare_kernel(const ActParams P) {
  extern __shared__ float sp[];  // [c] a = rstd*gamma, [c] b = beta - a*mean
  // programmatic dependent launch: the CTAs may be scheduled while the producer of `src` / `gn_stats` still drains
  griddep_launch_dependents();
  griddep_wait();
  const tfmq_act_desc& d = P.d;
  const int n = blockIdx.y;
  const int c = d.c;
  if (NORM == ACT_NORM_GN) {
    // Group moments from the double sums: three double multiply-adds per group (1 / count comes from the host; the
    // double-precision divide and square root this replaced cost ~200 DP instructions per group in EVERY CTA, on a part
    // with a 1/64-rate DP pipe), then rstd in fp32 as torch's own GroupNorm computes it.
    __shared__ float sg[2 * 64];
    const int cpg = c / d.groups;
    for (int g = threadIdx.x; g < d.groups; g += ACT_THREADS) {
      const double su = d.gn_stats[((long long)n * d.groups + g) * 2];
      const double sq = d.gn_stats[((long long)n * d.groups + g) * 2 + 1];
      const double mean = su * P.inv_cnt;
      double var = fma(-mean, mean, sq * P.inv_cnt);
      if (var < 0) var = 0;
      sg[2 * g] = 1.f / sqrtf((float)var + d.eps);
      sg[2 * g + 1] = (float)mean;
    }
    __syncthreads();
    for (int ch = threadIdx.x; ch < c; ch += ACT_THREADS) {
      const int g = ch / cpg;
      const float a = sg[2 * g] * d.gamma[ch];
      sp[ch] = a;
      sp[c + ch] = -a * sg[2 * g + 1] + d.beta[ch];
    }
    __syncthreads();
  }
  float delta = 1.f, zp = 0.f;
  if (OUT == ACT_OUT_U8) {
    delta = d.aq[0];
    zp = d.aq[1];
  }
  const float inv = __frcp_rn(delta);
  const int halo = (OUT == ACT_OUT_U8) ? d.halo : 0;
  const int Wp = P.out_w + 2 * halo, Hp = P.out_h + 2 * halo;
  const int npix = Wp * Hp;
  const int nvec = c >> 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NW = ACT_THREADS / 32;
  const int p0 = blockIdx.x * P.pix_per_cta;
  const int p1 = min(p0 + P.pix_per_cta, npix);
  const uint32_t zfill = (uint32_t)zp * 0x01010101u;
  constexpr bool gn = NORM == ACT_NORM_GN;
  for (int pp = p0 + warp; pp < p1; pp += NW) {
    const int yy = pp / Wp, xx = pp - yy * Wp;
    const int y = yy - halo, x = xx - halo;
    const bool border = (y < 0) | (x < 0) | (y >= P.out_h) | (x >= P.out_w);
    uint32_t* o8 = (OUT == ACT_OUT_U8)
                       ? reinterpret_cast<uint32_t*>(d.dst_u8 + ((long long)n * npix + pp) * d.dst_c + d.dst_c_off)
                       : nullptr;
    if (OUT == ACT_OUT_U8 && border) {
      for (int v = lane; v < nvec; v += 32) o8[v] = zfill;
      continue;
    }
    const int sy = d.upsample ? (y >> 1) : y, sx = d.upsample ? (x >> 1) : x;
    const float4* src = reinterpret_cast<const float4*>(d.src + (((long long)n * d.h + sy) * d.w + sx) * d.src_ld);
    float4* o32 = (OUT == ACT_OUT_F32) ? reinterpret_cast<float4*>(d.dst_f32 + ((long long)n * npix + pp) * d.dst_ld) : nullptr;
    uint2* ohi = (OUT == ACT_OUT_H16)
                     ? reinterpret_cast<uint2*>(static_cast<__half*>(d.dst_hi) + ((long long)n * npix + pp) * d.dst_h_ld)
                     : nullptr;
    uint2* olo = (OUT == ACT_OUT_H16)
                     ? reinterpret_cast<uint2*>(static_cast<__half*>(d.dst_lo) + ((long long)n * npix + pp) * d.dst_h_ld)
                     : nullptr;
    float ln_scale = 1.f, ln_shift = 0.f;
    if (NORM == ACT_NORM_LN) {
      // LayerNorm over the c channels of this token (nn.LayerNorm in BasicTransformerBlock, attention.py:205-207):
      // row moments in double, then y = (x * rstd - rstd * mean) * gamma + beta
      double s1 = 0.0, s2 = 0.0;
      for (int v = lane; v < nvec; v += 32) {
        const float4 f = src[v];
        s1 += (double)f.x + (double)f.y + (double)f.z + (double)f.w;
        s2 += (double)f.x * f.x + (double)f.y * f.y + (double)f.z * f.z + (double)f.w * f.w;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      const double mean = s1 / c;
      double var = s2 / c - mean * mean;
      if (var < 0) var = 0;
      ln_scale = (float)(1.0 / sqrt(var + (double)d.ln_eps));
      ln_shift = -ln_scale * (float)mean;
    }
    for (int v0 = lane; v0 < nvec; v0 += 64) {
      const int v1 = v0 + 32;
      const bool has1 = v1 < nvec;
      const float4 f0 = src[v0];
      const float4 f1 = has1 ? src[v1] : make_float4(0.f, 0.f, 0.f, 0.f);
      float t0[4] = {f0.x, f0.y, f0.z, f0.w}, t1[4] = {f1.x, f1.y, f1.z, f1.w};
      if (NORM == ACT_NORM_LN) {
        const float4 g0 = *reinterpret_cast<const float4*>(d.ln_gamma + v0 * 4), b0 = *reinterpret_cast<const float4*>(d.ln_beta + v0 * 4);
        t0[0] = fmaf(fmaf(t0[0], ln_scale, ln_shift), g0.x, b0.x), t0[1] = fmaf(fmaf(t0[1], ln_scale, ln_shift), g0.y, b0.y);
        t0[2] = fmaf(fmaf(t0[2], ln_scale, ln_shift), g0.z, b0.z), t0[3] = fmaf(fmaf(t0[3], ln_scale, ln_shift), g0.w, b0.w);
        if (has1) {
          const float4 g1 = *reinterpret_cast<const float4*>(d.ln_gamma + v1 * 4), b1 = *reinterpret_cast<const float4*>(d.ln_beta + v1 * 4);
          t1[0] = fmaf(fmaf(t1[0], ln_scale, ln_shift), g1.x, b1.x), t1[1] = fmaf(fmaf(t1[1], ln_scale, ln_shift), g1.y, b1.y);
          t1[2] = fmaf(fmaf(t1[2], ln_scale, ln_shift), g1.z, b1.z), t1[3] = fmaf(fmaf(t1[3], ln_scale, ln_shift), g1.w, b1.w);
        }
      }
      if (GEGLU) {
        // GEGLU (attention.py:37-44): the source row holds [value (c) | gate (c)]; out = value * gelu(gate), exact erf form
        const float4 q0 = src[nvec + v0];
        const float4 q1 = has1 ? src[nvec + v1] : make_float4(0.f, 0.f, 0.f, 0.f);
        const float gq0[4] = {q0.x, q0.y, q0.z, q0.w}, gq1[4] = {q1.x, q1.y, q1.z, q1.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          t0[j] *= gelu_f(gq0[j]);
          t1[j] *= gelu_f(gq1[j]);
        }
      }
      if (gn) {
        const float4 a0 = *reinterpret_cast<const float4*>(sp + v0 * 4), b0 = *reinterpret_cast<const float4*>(sp + c + v0 * 4);
        t0[0] = fmaf(t0[0], a0.x, b0.x), t0[1] = fmaf(t0[1], a0.y, b0.y);
        t0[2] = fmaf(t0[2], a0.z, b0.z), t0[3] = fmaf(t0[3], a0.w, b0.w);
        if (has1) {
          const float4 a1 = *reinterpret_cast<const float4*>(sp + v1 * 4), b1 = *reinterpret_cast<const float4*>(sp + c + v1 * 4);
          t1[0] = fmaf(t1[0], a1.x, b1.x), t1[1] = fmaf(t1[1], a1.y, b1.y);
          t1[2] = fmaf(t1[2], a1.z, b1.z), t1[3] = fmaf(t1[3], a1.w, b1.w);
        }
      }
      if (SILU) {
#pragma unroll
        for (int j = 0; j < 4; ++j) t0[j] = silu_f(t0[j]), t1[j] = silu_f(t1[j]);
      }
      if (OUT == ACT_OUT_U8) {
        o8[v0] = quant4(t0, delta, inv, zp);
        if (has1) o8[v1] = quant4(t1, delta, inv, zp);
      } else if (OUT == ACT_OUT_H16) {
        uint2 h, l;
        split_h16x4(t0, h, l);
        ohi[v0] = h, olo[v0] = l;
        if (has1) {
          split_h16x4(t1, h, l);
          ohi[v1] = h, olo[v1] = l;
        }
      } else {
        o32[v0] = make_float4(t0[0], t0[1], t0[2], t0[3]);
        if (has1) o32[v1] = make_float4(t1[0], t1[1], t1[2], t1[3]);
      }
    }
  }
}

// ------------------------------------------------- the same producer, flattened (GroupNorm / no normalisation)
// The warp-per-pixel kernel above spends ~40 issue slots per element at the LDM-4 shapes (ncu: issue-bound at a third of the
// DRAM rate): the per-pixel decode (division by the row length, border test, 64-bit addresses) is amortised over only
// c / 128 vectors per lane, a quarter of the lanes idle when c / 4 is not a multiple of 32, the SiLU's __expf carries a
// denormal-range fix-up and every element has its own branch to the exact-rounding path.  Here a thread owns float4 VECTORS of
// the flattened [destination pixels x c / 4] space: two multiply-high divisions per vector recover (pixel, channel vector),
// all lanes work, the exponential is one ex2.approx.ftz, the exact-rounding path is tested once per vector, and ACT_UNROLL
// vectors per thread are in flight.  The zero-point halo is filled by a separate short loop over the border pixels.
constexpr int ACT_UNROLL_DEFAULT = 2;

struct ActFlatParams {
  tfmq_act_desc d;
  int out_h, out_w;
  int nvec;                  // c / 4
  uint32_t nvec_magic;       // ceil(2^32 / nvec): idx / nvec = umulhi(idx, magic) for idx * nvec < 2^32
  uint32_t ow_magic;         // the same for out_w
  unsigned vec_per_cta;      // flattened vectors per CTA (a multiple of ACT_THREADS * ACT_UNROLL)
  unsigned vec_total;        // out_h * out_w * nvec
  int border_per_cta;        // halo pixels per CTA
  double inv_cnt;
};

__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// x * sigmoid(x) = x / (1 + 2^(-x log2 e)); a result of the exponential below the normal range flushes to 0 (sigmoid = 1)
__device__ __forceinline__ float silu_fast(float v) { return v * rcp_approx(1.f + ex2_ftz(v * -1.4426950408889634f)); }

template <bool GN, bool SILU, int OUT, int ACT_UNROLL>
__global__ void __launch_bounds__(ACT_THREADS) act_flat_kernel(const ActFlatParams P) {
  extern __shared__ float sp[];  // [c] a = rstd*gamma, [c] b = beta - a*mean
  griddep_launch_dependents();
  griddep_wait();
  const tfmq_act_desc& d = P.d;
  const int n = blockIdx.y;
  const int c = d.c;
  if (GN) {
    __shared__ float sg[2 * 64];
    const int cpg = c / d.groups;
    for (int g = threadIdx.x; g < d.groups; g += ACT_THREADS) {
      const double su = d.gn_stats[((long long)n * d.groups + g) * 2];
      const double sq = d.gn_stats[((long long)n * d.groups + g) * 2 + 1];
      const double mean = su * P.inv_cnt;
      double var = fma(-mean, mean, sq * P.inv_cnt);
      if (var < 0) var = 0;
      sg[2 * g] = 1.f / sqrtf((float)var + d.eps);
      sg[2 * g + 1] = (float)mean;
    }
    __syncthreads();
    for (int ch = threadIdx.x; ch < c; ch += ACT_THREADS) {
      const int g = ch / cpg;
      const float a = sg[2 * g] * d.gamma[ch];
      sp[ch] = a;
      sp[c + ch] = -a * sg[2 * g + 1] + d.beta[ch];
    }
    __syncthreads();
  }
  float delta = 1.f, zp = 0.f;
  if (OUT == ACT_OUT_U8) delta = d.aq[0], zp = d.aq[1];
  const float inv = __frcp_rn(delta);
  const int halo = (OUT == ACT_OUT_U8) ? d.halo : 0;
  const int Wp = P.out_w + 2 * halo;
  const int npix = Wp * (P.out_h + 2 * halo);
  const int nvec = P.nvec;
  const int up = d.upsample ? 1 : 0;
  const float* src_img = d.src + (long long)n * d.h * d.w * d.src_ld;
  const unsigned v_begin = blockIdx.x * P.vec_per_cta;
  const unsigned v_end = min(v_begin + P.vec_per_cta, P.vec_total);
  for (unsigned base = v_begin + threadIdx.x; base < v_end; base += ACT_THREADS * ACT_UNROLL) {
    float4 f[ACT_UNROLL];
    int vch[ACT_UNROLL];
    long long dpix[ACT_UNROLL];
    bool ok[ACT_UNROLL];
#pragma unroll
    for (int u = 0; u < ACT_UNROLL; ++u) {
      const unsigned idx = base + u * ACT_THREADS;
      ok[u] = idx < v_end;
      const unsigned pix = __umulhi(idx, P.nvec_magic);          // destination interior pixel (row-major oy, ox)
      vch[u] = (int)(idx - pix * (unsigned)nvec);
      const unsigned oy = __umulhi(pix, P.ow_magic);
      const unsigned ox = pix - oy * (unsigned)P.out_w;
      dpix[u] = (long long)n * npix + (long long)(oy + halo) * Wp + (ox + halo);
      const long long soff = ((long long)(oy >> up) * d.w + (ox >> up)) * d.src_ld + vch[u] * 4;
      f[u] = ok[u] ? *reinterpret_cast<const float4*>(src_img + soff) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < ACT_UNROLL; ++u) {
      float t[4] = {f[u].x, f[u].y, f[u].z, f[u].w};
      if (GN) {
        const float4 a = *reinterpret_cast<const float4*>(sp + vch[u] * 4), b = *reinterpret_cast<const float4*>(sp + c + vch[u] * 4);
        t[0] = fmaf(t[0], a.x, b.x), t[1] = fmaf(t[1], a.y, b.y), t[2] = fmaf(t[2], a.z, b.z), t[3] = fmaf(t[3], a.w, b.w);
      }
      if (SILU) {
#pragma unroll
        for (int j = 0; j < 4; ++j) t[j] = silu_fast(t[j]);
      }
      if (!ok[u]) continue;
      if (OUT == ACT_OUT_U8) {
        // fast rounding of the four codes; ONE test whether any of them sits within 2e-4 of a rounding boundary (the fast
        // value is within 1e-4 of the reference's pre-rounding value), and only then the exact divisions
        float m[4];
        bool slow = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float r = fminf(fmaxf(fmaf(t[j], inv, zp), 0.f), 255.f);
          m[j] = __fadd_rn(r, 12582912.f);
          slow |= fabsf(r - __fsub_rn(m[j], 12582912.f)) > 0.4998f;
        }
        if (slow) {
#pragma unroll
          for (int j = 0; j < 4; ++j) m[j] = quant_slow(t[j], delta, zp);
        }
        const uint32_t lo = __byte_perm(__float_as_uint(m[0]), __float_as_uint(m[1]), 0x0040);
        const uint32_t hi = __byte_perm(__float_as_uint(m[2]), __float_as_uint(m[3]), 0x0040);
        *reinterpret_cast<uint32_t*>(d.dst_u8 + dpix[u] * d.dst_c + d.dst_c_off + vch[u] * 4) = __byte_perm(lo, hi, 0x5410);
      } else if (OUT == ACT_OUT_H16) {
        uint2 h, l;
        split_h16x4(t, h, l);
        *reinterpret_cast<uint2*>(static_cast<__half*>(d.dst_hi) + dpix[u] * d.dst_h_ld + vch[u] * 4) = h;
        *reinterpret_cast<uint2*>(static_cast<__half*>(d.dst_lo) + dpix[u] * d.dst_h_ld + vch[u] * 4) = l;
      } else {
        *reinterpret_cast<float4*>(d.dst_f32 + dpix[u] * d.dst_ld + vch[u] * 4) = make_float4(t[0], t[1], t[2], t[3]);
      }
    }
  }
  if (OUT == ACT_OUT_U8 && halo) {
    // zero-point border: pixel b of the ring, b in [0, 2 Wp + 2 out_h): top row, bottom row, left column, right column
    const int nborder = 2 * Wp + 2 * P.out_h;
    const int b0 = blockIdx.x * P.border_per_cta, b1 = min(b0 + P.border_per_cta, nborder);
    const uint32_t zfill = (uint32_t)zp * 0x01010101u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int b = b0 + warp; b < b1; b += ACT_THREADS / 32) {
      int yy, xx;
      if (b < Wp) yy = 0, xx = b;
      else if (b < 2 * Wp) yy = P.out_h + 1, xx = b - Wp;
      else if (b < 2 * Wp + P.out_h) yy = b - 2 * Wp + 1, xx = 0;
      else yy = b - 2 * Wp - P.out_h + 1, xx = Wp - 1;
      uint32_t* o8 = reinterpret_cast<uint32_t*>(d.dst_u8 + ((long long)n * npix + (long long)yy * Wp + xx) * d.dst_c + d.dst_c_off);
      for (int v = lane; v < nvec; v += 32) o8[v] = zfill;
    }
  }
}


// ------------------------------------------------- the same producer with VECTOR-STATIONARY threads
// Both kernels above spend more than half of their issue slots on addressing (SASS of the flattened u8 kernel: ~55 of ~105
// instructions per float4 vector are multiply-high divisions, 64-bit IMAD.WIDE chains and the reload of the GroupNorm constants
// from shared memory), and the kernel is issue-bound at 2.9 TB/s.  Here blockDim.x = TV * k with TV = (c / 4) / R: a thread owns
// the SAME R channel vectors for its whole life and walks pixels with stride k.  Its GroupNorm constants are loaded once into
// registers, the (pixel, row) pair advances by additions, all offsets are 32-bit relative to per-image base pointers, and two
// pixels are in flight per thread.
struct ActStatParams {
  tfmq_act_desc d;
  int out_h, out_w;
  int tv;                    // threads along the channel-vector axis: (c / 4) / R
  int k;                     // pixels per CTA pass (blockDim.x = tv * k)
  int pix_per_cta;           // interior destination pixels per CTA
  int border_per_cta;        // halo pixels per CTA
  double inv_cnt;
};

template <bool GN, bool SILU, int OUT, int R, int U>
__global__ void __launch_bounds__(ACT_THREADS) act_stat_kernel(const ActStatParams P) {
  extern __shared__ float sp[];  // [c] a = rstd*gamma, [c] b = beta - a*mean
  griddep_launch_dependents();
  griddep_wait();
  const tfmq_act_desc& d = P.d;
  const int n = blockIdx.y;
  const int c = d.c;
  const int tv = threadIdx.x % P.tv, tp = threadIdx.x / P.tv;
  float4 ga[R], gb[R];
  if (GN) {
    __shared__ float sg[2 * 64];
    const int cpg = c / d.groups;
    for (int g = threadIdx.x; g < d.groups; g += blockDim.x) {
      const double su = d.gn_stats[((long long)n * d.groups + g) * 2];
      const double sq = d.gn_stats[((long long)n * d.groups + g) * 2 + 1];
      const double mean = su * P.inv_cnt;
      double var = fma(-mean, mean, sq * P.inv_cnt);
      if (var < 0) var = 0;
      sg[2 * g] = 1.f / sqrtf((float)var + d.eps);
      sg[2 * g + 1] = (float)mean;
    }
    __syncthreads();
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
      const int g = ch / cpg;
      const float a = sg[2 * g] * d.gamma[ch];
      sp[ch] = a;
      sp[c + ch] = -a * sg[2 * g + 1] + d.beta[ch];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < R; ++j) {
      ga[j] = *reinterpret_cast<const float4*>(sp + (tv + j * P.tv) * 4);
      gb[j] = *reinterpret_cast<const float4*>(sp + c + (tv + j * P.tv) * 4);
    }
  }
  float delta = 1.f, zp = 0.f;
  if (OUT == ACT_OUT_U8) delta = d.aq[0], zp = d.aq[1];
  const float inv = __frcp_rn(delta);
  const int halo = (OUT == ACT_OUT_U8) ? d.halo : 0;
  const int W = P.out_w, Wp = W + 2 * halo;
  const int npix_dst = Wp * (P.out_h + 2 * halo);
  const int up = d.upsample ? 1 : 0;
  // per-image bases; everything below is a 32-bit offset from them
  const float* src_img = d.src + (long long)n * d.h * d.w * d.src_ld + tv * 4;
  uint8_t* dst8 = (OUT == ACT_OUT_U8) ? d.dst_u8 + (long long)n * npix_dst * d.dst_c + d.dst_c_off + tv * 4 : nullptr;
  __half* dsth = (OUT == ACT_OUT_H16) ? static_cast<__half*>(d.dst_hi) + (long long)n * npix_dst * d.dst_h_ld + tv * 4 : nullptr;
  __half* dstl = (OUT == ACT_OUT_H16) ? static_cast<__half*>(d.dst_lo) + (long long)n * npix_dst * d.dst_h_ld + tv * 4 : nullptr;
  float* dstf = (OUT == ACT_OUT_F32) ? d.dst_f32 + (long long)n * npix_dst * d.dst_ld + tv * 4 : nullptr;
  const unsigned src_ld = (unsigned)d.src_ld, vstep = (unsigned)P.tv * 4u;
  const int p0 = blockIdx.x * P.pix_per_cta;
  const int p1 = min(p0 + P.pix_per_cta, P.out_h * W);
  const int k = P.k;
  int pix = p0 + tp;
  int oy = pix / W, ox = pix - oy * W;
  for (; pix < p1; pix += U * k) {
    float4 f[U][R];
    unsigned doff[U];
    bool ok[U];
    int oy_u = oy, ox_u = ox;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      ok[u] = pix + u * k < p1;
      const unsigned spix = up ? (unsigned)((oy_u >> 1) * d.w + (ox_u >> 1)) : (unsigned)(pix + u * k);
      const unsigned dpix = (unsigned)((oy_u + halo) * Wp + ox_u + halo);
      doff[u] = OUT == ACT_OUT_U8 ? dpix * (unsigned)d.dst_c : OUT == ACT_OUT_H16 ? dpix * (unsigned)d.dst_h_ld : dpix * (unsigned)d.dst_ld;
#pragma unroll
      for (int j = 0; j < R; ++j)
        f[u][j] = ok[u] ? *reinterpret_cast<const float4*>(src_img + spix * src_ld + j * vstep) : make_float4(0.f, 0.f, 0.f, 0.f);
      ox_u += k;
      while (ox_u >= W) ox_u -= W, ++oy_u;
    }
    oy = oy_u, ox = ox_u;
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int j = 0; j < R; ++j) {
        float t[4] = {f[u][j].x, f[u][j].y, f[u][j].z, f[u][j].w};
        if (GN) {
          t[0] = fmaf(t[0], ga[j].x, gb[j].x), t[1] = fmaf(t[1], ga[j].y, gb[j].y);
          t[2] = fmaf(t[2], ga[j].z, gb[j].z), t[3] = fmaf(t[3], ga[j].w, gb[j].w);
        }
        if (SILU) {
#pragma unroll
          for (int e = 0; e < 4; ++e) t[e] = silu_fast(t[e]);
        }
        if (!ok[u]) continue;
        if (OUT == ACT_OUT_U8) {
          // fast rounding of the four codes; ONE test whether any of them sits within 2e-4 of a rounding boundary (the fast
          // value is within 1e-4 of the reference's pre-rounding value), and only then the exact divisions
          float m[4];
          bool slow = false;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float r = fminf(fmaxf(fmaf(t[e], inv, zp), 0.f), 255.f);
            m[e] = __fadd_rn(r, 12582912.f);
            slow |= fabsf(r - __fsub_rn(m[e], 12582912.f)) > 0.4998f;
          }
          if (slow) {
#pragma unroll
            for (int e = 0; e < 4; ++e) m[e] = quant_slow(t[e], delta, zp);
          }
          const uint32_t lo = __byte_perm(__float_as_uint(m[0]), __float_as_uint(m[1]), 0x0040);
          const uint32_t hi = __byte_perm(__float_as_uint(m[2]), __float_as_uint(m[3]), 0x0040);
          *reinterpret_cast<uint32_t*>(dst8 + doff[u] + j * vstep) = __byte_perm(lo, hi, 0x5410);
        } else if (OUT == ACT_OUT_H16) {
          uint2 h, l;
          split_h16x4(t, h, l);
          *reinterpret_cast<uint2*>(dsth + doff[u] + j * vstep) = h;
          *reinterpret_cast<uint2*>(dstl + doff[u] + j * vstep) = l;
        } else {
          *reinterpret_cast<float4*>(dstf + doff[u] + j * vstep) = make_float4(t[0], t[1], t[2], t[3]);
        }
      }
    }
  }
  if (OUT == ACT_OUT_U8 && halo) {
    // zero-point border: pixel b of the ring, b in [0, 2 Wp + 2 out_h): top row, bottom row, left column, right column
    const int nborder = 2 * Wp + 2 * P.out_h;
    const int b0 = blockIdx.x * P.border_per_cta, b1 = min(b0 + P.border_per_cta, nborder);
    const uint32_t zfill = (uint32_t)zp * 0x01010101u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = (blockDim.x + 31) >> 5;
    const int nvec = c >> 2;
    for (int b = b0 + warp; b < b1; b += nwarp) {
      int yy, xx;
      if (b < Wp) yy = 0, xx = b;
      else if (b < 2 * Wp) yy = P.out_h + 1, xx = b - Wp;
      else if (b < 2 * Wp + P.out_h) yy = b - 2 * Wp + 1, xx = 0;
      else yy = b - 2 * Wp - P.out_h + 1, xx = Wp - 1;
      uint32_t* o8 = reinterpret_cast<uint32_t*>(d.dst_u8 + ((long long)n * npix_dst + (long long)yy * Wp + xx) * d.dst_c + d.dst_c_off);
      for (int v = lane; v < nvec; v += 32) o8[v] = zfill;
    }
  }
}

// ---------------------------------------------------------------- DDIM update
__global__ void ddim_update_kernel(const float* __restrict__ x, const float* __restrict__ e,
                                   const float* __restrict__ noise, const float* __restrict__ coef, long long count,
                                   float* __restrict__ x_prev, float* __restrict__ x0_out) {
  const float sa = coef[0], s1ma = coef[1], sap = coef[2], c2 = coef[3];
  const float c1 = noise ? coef[4] : 0.f;
  // coef[5] != 0 (read only with noise): the LDM sampler's order  (sqrt(a') x0 + c2 e) + c1 noise  (ddim.py:205-211) instead of
  // the DDIM runner's  (sqrt(a') x0 + c1 noise) + c2 e  (denoising.py:31-37)
  const bool dir_first = noise && coef[5] != 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x) {
    const float ev = e[i];
    // (x - e*sqrt(1-a)) / sqrt(a), then sqrt(a')*x0 + c1*noise + c2*e : same operation order as
    // ddim/functions/denoising.py:31-37 so the fp32 roundings agree; no FMA contraction
    const float x0 = __fdiv_rn(__fsub_rn(x[i], __fmul_rn(ev, s1ma)), sa);
    float r = __fmul_rn(sap, x0);
    if (dir_first) {
      r = __fadd_rn(__fadd_rn(r, __fmul_rn(c2, ev)), __fmul_rn(c1, noise[i]));
    } else {
      if (noise) r = __fadd_rn(r, __fmul_rn(c1, noise[i]));
      r = __fadd_rn(r, __fmul_rn(c2, ev));
    }
    x_prev[i] = r;
    if (x0_out) x0_out[i] = x0;
  }
}

__global__ void cfg_combine_kernel(const float* __restrict__ eu, const float* __restrict__ ec, float s, long long count,
                                   float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = __fadd_rn(eu[i], __fmul_rn(s, __fsub_rn(ec[i], eu[i])));
}

// PLMS multistep combination of the current and stored noise predictions (ldm/models/diffusion/plms.py:226-238), with the
// reference's operation order so the fp32 roundings agree:
//   order 0: e0                       (first half of the pseudo improved Euler step, and plain DDIM)
//   order 1: (e0 + e1) / 2            (second half: e1 = model output at x_prev)
//   order 2: (3 e0 - e1) / 2          order 3: (23 e0 - 16 e1 + 5 e2) / 12
//   order 4: (55 e0 - 59 e1 + 37 e2 - 9 e3) / 24        with e1, e2, e3 = old_eps[-1], [-2], [-3]
__global__ void plms_eps_kernel(const float* __restrict__ e0, const float* __restrict__ e1, const float* __restrict__ e2,
                                const float* __restrict__ e3, int order, long long count, float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x) {
    const float a = e0[i];
    float r = a;
    if (order == 1) {
      r = __fdiv_rn(__fadd_rn(a, e1[i]), 2.f);
    } else if (order == 2) {
      r = __fdiv_rn(__fsub_rn(__fmul_rn(3.f, a), e1[i]), 2.f);
    } else if (order == 3) {
      r = __fdiv_rn(__fadd_rn(__fsub_rn(__fmul_rn(23.f, a), __fmul_rn(16.f, e1[i])), __fmul_rn(5.f, e2[i])), 12.f);
    } else if (order >= 4) {
      r = __fdiv_rn(__fsub_rn(__fadd_rn(__fsub_rn(__fmul_rn(55.f, a), __fmul_rn(59.f, e1[i])), __fmul_rn(37.f, e2[i])),
                              __fmul_rn(9.f, e3[i])),
                    24.f);
    }
    out[i] = r;
  }
}

__global__ void timestep_embedding_kernel(const float* __restrict__ t, int m, int dim, int style,
                                          float* __restrict__ out) {
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m * half) return;
  const int r = i / half, j = i - r * half;
  float freq;
  if (style == 0)
    freq = expf((float)j * -(logf(10000.f) / (float)(half - 1)));
  else
    freq = expf(-logf(10000.f) * (float)j / (float)half);
  const float a = t[r] * freq;
  float* o = out + (long long)r * dim;
  if (style == 0) {
    o[j] = sinf(a);
    o[half + j] = cosf(a);
  } else {
    o[j] = cosf(a);
    o[half + j] = sinf(a);
  }
  if ((dim & 1) && j == 0) o[dim - 1] = 0.f;
}

}  // namespace tfmq

using namespace tfmq;

extern "C" int tfmq_fill_zero(tfmq_ctx* ctx, void* p, size_t bytes, void* stream) {
  if (!ctx) return TFMQ_ERR_ARG;
  TFMQ_REQUIRE(p || bytes == 0, TFMQ_ERR_ARG, "fill_zero: null pointer");
  cudaError_t e = cudaMemsetAsync(p, 0, bytes, tfmq_stream(stream));
  if (e != cudaSuccess) return tfmq_fail(ctx, TFMQ_ERR_CUDA, "fill_zero: %s", cudaGetErrorString(e));
  return TFMQ_OK;
}

extern "C" int tfmq_gn_stats(tfmq_ctx* ctx, const float* x, int64_t ld, int n, int hw, int c, int groups,
                             double* stats, void* stream) {
  if (!ctx) return TFMQ_ERR_ARG;
  TFMQ_REQUIRE(groups > 0 && c % groups == 0, TFMQ_ERR_SHAPE, "gn_stats: c=%d groups=%d", c, groups);
  tfmq_gn_target t;
  t.stats = stats, t.cpg = c / groups, t.ch_off = 0, t.groups = groups, t.reserved = 0;
  return tfmq_gn_stats_part(ctx, x, ld, n, hw, c, &t, stream);
}

extern "C" int tfmq_gn_stats_part(tfmq_ctx* ctx, const float* x, int64_t ld, int n, int hw, int c,
                                  const tfmq_gn_target* target, void* stream) {
  if (!ctx) return TFMQ_ERR_ARG;
  TFMQ_REQUIRE(x && target && target->stats, TFMQ_ERR_ARG, "gn_stats: null pointer");
  double* stats = target->stats;
  const int groups = target->groups;
  TFMQ_REQUIRE(c % 4 == 0 && ld % 4 == 0 && c <= 4 * GN_THREADS * GN_MAXVEC && target->cpg > 0 &&
                   target->ch_off >= 0 && (target->ch_off + c + target->cpg - 1) / target->cpg <= groups,
               TFMQ_ERR_SHAPE, "gn_stats: c=%d cpg=%d ch_off=%d groups=%d ld=%lld", c, target->cpg, target->ch_off, groups,
               (long long)ld);
  TFMQ_REQUIRE(((uintptr_t)x & 15) == 0, TFMQ_ERR_ARG, "gn_stats: x must be 16-byte aligned");
  if (n == 0 || hw == 0) return TFMQ_OK;
  int chunks = (ctx->sm_count * 4 + n - 1) / n;
  if (chunks > hw) chunks = hw;
  if (chunks < 1) chunks = 1;
  const int ppc = (hw + chunks - 1) / chunks;
  chunks = (hw + ppc - 1) / ppc;
  const int nvec = c / 4;
  const int tvn = nvec < GN_THREADS ? nvec : GN_THREADS, k = nvec < GN_THREADS ? GN_THREADS / nvec : 1;   // thread grid tvn x k
  gn_stats_kernel<<<dim3(chunks, n), GN_THREADS, 2 * c * sizeof(double), tfmq_stream(stream)>>>(
      x, ld, hw, c, groups, ppc, stats, target->cpg, target->ch_off, tvn, k);
  TFMQ_LAUNCH_CHECK("gn_stats");
  return TFMQ_OK;
}

extern "C" int tfmq_act_prepare(tfmq_ctx* ctx, const tfmq_act_desc* d, void* stream) {
  if (!ctx) return TFMQ_ERR_ARG;
  TFMQ_REQUIRE(d && d->src, TFMQ_ERR_ARG, "act_prepare: null pointer");
  TFMQ_REQUIRE((d->dst_u8 != nullptr) + (d->dst_f32 != nullptr) + (d->dst_hi != nullptr) == 1, TFMQ_ERR_ARG,
               "act_prepare: exactly one of dst_u8 / dst_f32 / dst_hi");
  TFMQ_REQUIRE(!d->dst_hi || (d->dst_lo && d->dst_h_ld % 4 == 0 && ((uintptr_t)d->dst_hi & 7) == 0 &&
                              ((uintptr_t)d->dst_lo & 7) == 0),
               TFMQ_ERR_ARG, "act_prepare: dst_lo missing or dst_h_ld / alignment");
  TFMQ_REQUIRE(d->c % 4 == 0 && d->src_ld % 4 == 0, TFMQ_ERR_SHAPE, "act_prepare: c/ld not multiple of 4");
  TFMQ_REQUIRE(((uintptr_t)d->src & 15) == 0, TFMQ_ERR_ARG, "act_prepare: src must be 16-byte aligned");
  if (d->dst_u8) {
    TFMQ_REQUIRE(d->aq, TFMQ_ERR_ARG, "act_prepare: aq required for u8 output");
    TFMQ_REQUIRE(d->dst_c % 4 == 0 && d->dst_c_off % 4 == 0 && d->dst_c_off + d->c <= d->dst_c, TFMQ_ERR_SHAPE,
                 "act_prepare: bad dst channel window");
    TFMQ_REQUIRE(d->halo == 0 || d->halo == 1, TFMQ_ERR_ARG, "act_prepare: halo");
    TFMQ_REQUIRE(((uintptr_t)d->dst_u8 & 3) == 0, TFMQ_ERR_ARG, "act_prepare: dst must be 4-byte aligned");
  } else if (d->dst_f32) {
    TFMQ_REQUIRE(d->dst_ld % 4 == 0 && ((uintptr_t)d->dst_f32 & 15) == 0, TFMQ_ERR_SHAPE, "act_prepare: dst_ld/align");
  }
  TFMQ_REQUIRE(!d->ln_gamma || (d->ln_beta && !d->gn_stats && !d->upsample && ((uintptr_t)d->ln_gamma & 15) == 0 &&
                                ((uintptr_t)d->ln_beta & 15) == 0),
               TFMQ_ERR_ARG, "act_prepare: LayerNorm needs ln_beta, 16-byte aligned parameters, no GN / upsample");
  TFMQ_REQUIRE(!d->geglu || (!d->gn_stats && !d->ln_gamma && !d->upsample && !d->silu), TFMQ_ERR_ARG,
               "act_prepare: GEGLU excludes GN / LN / SiLU / upsample");
  if (d->gn_stats) {
    TFMQ_REQUIRE(d->gamma && d->beta && d->groups > 0 && d->groups <= 64 && d->c % d->groups == 0, TFMQ_ERR_ARG,
                 "act_prepare: GroupNorm parameters (groups <= 64)");
    TFMQ_REQUIRE(!d->upsample, TFMQ_ERR_ARG, "act_prepare: GN with upsample unsupported");
  }
  if (d->n == 0) return TFMQ_OK;
  // vector-stationary kernel: c / 4 = R * tv with tv <= 256 threads along the channel axis, R in {1, 2}; every offset inside
  // one image must fit 32 bits
  static const bool stat_env = !(getenv("TFMQ_ACT_STAT") && atoi(getenv("TFMQ_ACT_STAT")) == 0);
  {
    const int nvec = d->c / 4;
    const int R = nvec <= ACT_THREADS ? 1 : 2;
    const long long out_h = d->upsample ? 2 * d->h : d->h, out_w = d->upsample ? 2 * d->w : d->w;
    const int halo_s = d->dst_u8 ? d->halo : 0;
    const long long dst_pitch = d->dst_u8 ? d->dst_c : d->dst_hi ? d->dst_h_ld : d->dst_ld;
    const bool fits = (long long)d->h * d->w * d->src_ld < (1ll << 31) &&
                      (out_h + 2 * halo_s) * (out_w + 2 * halo_s) * dst_pitch < (1ll << 31);
    if (stat_env && !d->ln_gamma && !d->geglu && nvec % R == 0 && nvec / R <= ACT_THREADS && fits) {
      ActStatParams S;
      S.d = *d;
      S.out_h = (int)out_h, S.out_w = (int)out_w;
      S.tv = nvec / R;
      S.k = ACT_THREADS / S.tv;
      const int threads = S.tv * S.k;
      static const int stat_mult = getenv("TFMQ_ACT_CTAS") ? atoi(getenv("TFMQ_ACT_CTAS")) : 4;   // measured best of 3..8 (tools/microbench_act.py)
      const int npix_s = (int)(out_h * out_w);
      int chunks = (ctx->sm_count * stat_mult + d->n - 1) / d->n;
      int ppc = (npix_s + chunks - 1) / chunks;
      static const int stat_u = getenv("TFMQ_ACT_INFLIGHT") ? atoi(getenv("TFMQ_ACT_INFLIGHT")) : 4;
      const int U = (R == 1 && stat_u == 4) ? 4 : 2;               // pixels in flight per thread
      ppc = (ppc + U * S.k - 1) / (U * S.k) * (U * S.k);           // whole passes of U k pixels
      chunks = (npix_s + ppc - 1) / ppc;
      S.pix_per_cta = ppc;
      const int nborder = halo_s ? 2 * (S.out_w + 2) + 2 * S.out_h : 0;
      S.border_per_cta = (nborder + chunks - 1) / chunks;
      S.inv_cnt = d->gn_stats ? 1.0 / ((double)(d->c / d->groups) * d->h * d->w) : 0.0;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(chunks, d->n), cfg.blockDim = dim3(threads);
      cfg.dynamicSmemBytes = d->gn_stats ? 2 * (size_t)d->c * sizeof(float) : 0, cfg.stream = tfmq_stream(stream);
      cudaLaunchAttribute attr[1];
      cfg.attrs = attr, cfg.numAttrs = (unsigned)tfmq_pdl_attr(&attr[0]);
      const int out = d->dst_u8 ? ACT_OUT_U8 : d->dst_hi ? ACT_OUT_H16 : ACT_OUT_F32;
      cudaError_t e;
#define STAT_BY_OUT_R(GN, SILU, RR, UU)                                                                \
  (out == ACT_OUT_U8 ? cudaLaunchKernelEx(&cfg, act_stat_kernel<GN, SILU, ACT_OUT_U8, RR, UU>, S)     \
   : out == ACT_OUT_H16 ? cudaLaunchKernelEx(&cfg, act_stat_kernel<GN, SILU, ACT_OUT_H16, RR, UU>, S) \
                        : cudaLaunchKernelEx(&cfg, act_stat_kernel<GN, SILU, ACT_OUT_F32, RR, UU>, S))
#define STAT_BY_OUT(GN, SILU) \
  (R == 2 ? STAT_BY_OUT_R(GN, SILU, 2, 2) : U == 4 ? STAT_BY_OUT_R(GN, SILU, 1, 4) : STAT_BY_OUT_R(GN, SILU, 1, 2))
      if (d->gn_stats) e = d->silu ? STAT_BY_OUT(true, true) : STAT_BY_OUT(true, false);
      else e = d->silu ? STAT_BY_OUT(false, true) : STAT_BY_OUT(false, false);
#undef STAT_BY_OUT
#undef STAT_BY_OUT_R
      if (e != cudaSuccess) return tfmq_fail(ctx, TFMQ_ERR_CUDA, "act_prepare: launch: %s", cudaGetErrorString(e));
      TFMQ_LAUNCH_CHECK("act_prepare");
      return TFMQ_OK;
    }
  }
  static const bool flat_env = !(getenv("TFMQ_ACT_FLAT") && atoi(getenv("TFMQ_ACT_FLAT")) == 0);   // 0: the warp-per-pixel kernel
  const long long out_pix = (long long)(d->upsample ? 4 : 1) * d->h * d->w;
  // (the multiply-high divisions are exact while dividend * divisor < 2^32)
  // measured (tools/microbench_act.py, LDM-4 shapes): the flattened kernel is 13-20 % faster for the fp16-split / fp32
  // outputs; with u8 output both kernels sit at the same ~3.1 TB/s (bound by the per-launch latency chain, not by issue
  // slots), so the u8 path keeps the warp-per-pixel kernel whose halo fill is part of the same loop
  static const int flat_u8 = getenv("TFMQ_ACT_FLAT_U8") ? atoi(getenv("TFMQ_ACT_FLAT_U8")) : 0;
  if (flat_env && (!d->dst_u8 || flat_u8) && !d->ln_gamma && !d->geglu && out_pix * (d->c / 4) * (d->c / 4) < (1ll << 32) &&
      out_pix * (d->upsample ? 2 * d->w : d->w) < (1ll << 32)) {
    ActFlatParams F;
    F.d = *d;
    F.out_h = d->upsample ? 2 * d->h : d->h;
    F.out_w = d->upsample ? 2 * d->w : d->w;
    F.nvec = d->c / 4;
    F.nvec_magic = (uint32_t)(((1ull << 32) + F.nvec - 1) / F.nvec);
    F.ow_magic = (uint32_t)(((1ull << 32) + F.out_w - 1) / F.out_w);
    // umulhi(x, ceil(2^32 / k)) == x / k holds while x * k < 2^32 (the error term x * (k - 2^32 mod k) / 2^32 stays below 1)
    F.vec_total = (unsigned)(out_pix * F.nvec);
    static const int flat_mult = getenv("TFMQ_ACT_CTAS") ? atoi(getenv("TFMQ_ACT_CTAS")) : 8;
    int chunks = (ctx->sm_count * flat_mult + d->n - 1) / d->n;
    static const int unroll_env = getenv("TFMQ_ACT_UNROLL") ? atoi(getenv("TFMQ_ACT_UNROLL")) : ACT_UNROLL_DEFAULT;
    const int unroll = unroll_env == 4 ? 4 : 2;                  // float4 vectors in flight per thread
    const unsigned gran = ACT_THREADS * (unsigned)unroll;
    unsigned vpc = (F.vec_total + chunks - 1) / chunks;
    vpc = (vpc + gran - 1) / gran * gran;
    chunks = (int)((F.vec_total + vpc - 1) / vpc);
    F.vec_per_cta = vpc;
    const int halo_f = d->dst_u8 ? d->halo : 0;
    const int nborder = halo_f ? 2 * (F.out_w + 2) + 2 * F.out_h : 0;
    F.border_per_cta = (nborder + chunks - 1) / chunks;
    F.inv_cnt = d->gn_stats ? 1.0 / ((double)(d->c / d->groups) * d->h * d->w) : 0.0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(chunks, d->n), cfg.blockDim = dim3(ACT_THREADS);
    cfg.dynamicSmemBytes = d->gn_stats ? 2 * (size_t)d->c * sizeof(float) : 0, cfg.stream = tfmq_stream(stream);
    cudaLaunchAttribute attr[1];
    cfg.attrs = attr, cfg.numAttrs = (unsigned)tfmq_pdl_attr(&attr[0]);
    const int out = d->dst_u8 ? ACT_OUT_U8 : d->dst_hi ? ACT_OUT_H16 : ACT_OUT_F32;
    cudaError_t e;
#define FLAT_BY_OUT_U(GN, SILU, U)                                                                   \
  (out == ACT_OUT_U8 ? cudaLaunchKernelEx(&cfg, act_flat_kernel<GN, SILU, ACT_OUT_U8, U>, F)        \
   : out == ACT_OUT_H16 ? cudaLaunchKernelEx(&cfg, act_flat_kernel<GN, SILU, ACT_OUT_H16, U>, F)    \
                        : cudaLaunchKernelEx(&cfg, act_flat_kernel<GN, SILU, ACT_OUT_F32, U>, F))
#define FLAT_BY_OUT(GN, SILU) (unroll == 4 ? FLAT_BY_OUT_U(GN, SILU, 4) : FLAT_BY_OUT_U(GN, SILU, 2))
    if (d->gn_stats) e = d->silu ? FLAT_BY_OUT(true, true) : FLAT_BY_OUT(true, false);
    else e = d->silu ? FLAT_BY_OUT(false, true) : FLAT_BY_OUT(false, false);
#undef FLAT_BY_OUT
#undef FLAT_BY_OUT_U
    if (e != cudaSuccess) return tfmq_fail(ctx, TFMQ_ERR_CUDA, "act_prepare: launch: %s", cudaGetErrorString(e));
    TFMQ_LAUNCH_CHECK("act_prepare");
    return TFMQ_OK;
  }
  ActParams P;
  P.d = *d;
  P.out_h = d->upsample ? 2 * d->h : d->h;
  P.out_w = d->upsample ? 2 * d->w : d->w;
  const int halo = d->dst_u8 ? d->halo : 0;
  const int npix = (P.out_h + 2 * halo) * (P.out_w + 2 * halo);
  static const int cta_mult = getenv("TFMQ_ACT_CTAS") ? atoi(getenv("TFMQ_ACT_CTAS")) : 4;   // CTAs per SM (tuning aid)
  int chunks = (ctx->sm_count * cta_mult + d->n - 1) / d->n;
  if (chunks > npix) chunks = npix;
  const int ppc = (npix + chunks - 1) / chunks;
  chunks = (npix + ppc - 1) / ppc;
  P.pix_per_cta = ppc;
  P.inv_cnt = d->gn_stats ? 1.0 / ((double)(d->c / d->groups) * d->h * d->w) : 0.0;
  const size_t smem = d->gn_stats ? 2 * (size_t)d->c * sizeof(float) : 0;
  const dim3 grid(chunks, d->n);
  cudaStream_t st = tfmq_stream(stream);
  const int out = d->dst_u8 ? ACT_OUT_U8 : d->dst_hi ? ACT_OUT_H16 : ACT_OUT_F32;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid, cfg.blockDim = dim3(ACT_THREADS), cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr, cfg.numAttrs = (unsigned)tfmq_pdl_attr(&attr[0]);
#define ACT_LAUNCH(NORM, SILU, GEGLU, OUT) \
  cudaLaunchKernelEx(&cfg, act_prepare_kernel<NORM, SILU, GEGLU, OUT>, P)
#define ACT_BY_OUT(NORM, SILU, GEGLU)                         \
  do {                                                        \
    if (out == ACT_OUT_U8) ACT_LAUNCH(NORM, SILU, GEGLU, ACT_OUT_U8);        \
    else if (out == ACT_OUT_H16) ACT_LAUNCH(NORM, SILU, GEGLU, ACT_OUT_H16); \
    else ACT_LAUNCH(NORM, SILU, GEGLU, ACT_OUT_F32);          \
  } while (0)
  if (d->geglu) ACT_BY_OUT(ACT_NORM_NONE, false, true);
  else if (d->ln_gamma) {
    TFMQ_REQUIRE(!d->silu, TFMQ_ERR_ARG, "act_prepare: LayerNorm + SiLU is not a combination of the path");
    ACT_BY_OUT(ACT_NORM_LN, false, false);
  } else if (d->gn_stats) {
    if (d->silu) ACT_BY_OUT(ACT_NORM_GN, true, false);
    else ACT_BY_OUT(ACT_NORM_GN, false, false);
  } else {
    if (d->silu) ACT_BY_OUT(ACT_NORM_NONE, true, false);
    else ACT_BY_OUT(ACT_NORM_NONE, false, false);
  }
#undef ACT_BY_OUT
#undef ACT_LAUNCH
  TFMQ_LAUNCH_CHECK("act_prepare");
  return TFMQ_OK;
}

extern "C" int tfmq_ddim_update(tfmq_ctx* ctx, const float* x, const float* e, const float* noise, const float* coef,
                                int64_t count, float* x_prev, float* x0_out, void* stream) {
  if (!ctx) return TFMQ_ERR_ARG;
  TFMQ_REQUIRE(x && e && coef && x_prev, TFMQ_ERR_ARG, "ddim_update: null pointer");
  if (count == 0) return TFMQ_OK;
  int blocks = (int)((count + 255) / 256);
  if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
  ddim_update_kernel<<<blocks, 256, 0, tfmq_stream(stream)>>>(x, e, noise, coef, count, x_prev, x0_out);
  TFMQ_LAUNCH_CHECK("ddim_update");
  return TFMQ_OK;
}

extern "C" int tfmq_plms_eps(tfmq_ctx* ctx, const float* e0, const float* e1, const float* e2, const float* e3,
                             int order, int64_t count, float* out, void* stream) {
  if (!ctx) return TFMQ_ERR_ARG;
  TFMQ_REQUIRE(e0 && out && order >= 0 && order <= 4, TFMQ_ERR_ARG, "plms_eps: null pointer / order");
  TFMQ_REQUIRE((order < 1 || e1) && (order < 3 || e2) && (order < 4 || e3), TFMQ_ERR_ARG,
               "plms_eps: order %d needs more stored predictions", order);
  if (count <= 0) return TFMQ_OK;
  const int blocks = (int)((count + 255) / 256 < 1184 ? (count + 255) / 256 : 1184);
  plms_eps_kernel<<<blocks, 256, 0, tfmq_stream(stream)>>>(e0, e1, e2, e3, order, count, out);
  TFMQ_LAUNCH_CHECK("plms_eps");
  return TFMQ_OK;
}

extern "C" int tfmq_cfg_combine(tfmq_ctx* ctx, const float* e_uncond, const float* e_cond, float s, int64_t count,
                                float* out, void* stream) {
  if (!ctx) return TFMQ_ERR_ARG;
  TFMQ_REQUIRE(e_uncond && e_cond && out, TFMQ_ERR_ARG, "cfg_combine: null pointer");
  if (count == 0) return TFMQ_OK;
  int blocks = (int)((count + 255) / 256);
  if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
  cfg_combine_kernel<<<blocks, 256, 0, tfmq_stream(stream)>>>(e_uncond, e_cond, s, count, out);
  TFMQ_LAUNCH_CHECK("cfg_combine");
  return TFMQ_OK;
}

extern "C" int tfmq_timestep_embedding(tfmq_ctx* ctx, const float* t, int m, int dim, int style, float* out,
                                       void* stream) {
  if (!ctx) return TFMQ_ERR_ARG;
  TFMQ_REQUIRE(t && out, TFMQ_ERR_ARG, "timestep_embedding: null pointer");
  TFMQ_REQUIRE(dim >= 4 && (style == 0 || style == 1), TFMQ_ERR_ARG, "timestep_embedding: dim/style");
  if (m == 0) return TFMQ_OK;
  const int total = m * (dim / 2);
  timestep_embedding_kernel<<<(total + 127) / 128, 128, 0, tfmq_stream(stream)>>>(t, m, dim, style, out);
  TFMQ_LAUNCH_CHECK("timestep_embedding");
  return TFMQ_OK;
}
