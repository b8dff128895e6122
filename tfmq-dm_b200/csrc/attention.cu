// Fused QK^T - softmax - PV attention in fp32 (flash-style: no T x T matrix in HBM).
//
// Three kernels:
//  * attn_h16_kernel<D>  : head dim D in {32,64,80}: warp-level tensor-core MMAs (m16n8k16 fp16) on
//    two-term fp16 splits of the fp32 operands, 3 products per term pair: fp32-accurate.
//  * attn_mma_kernel<D>  : head dim 40 (not a multiple of 16) or unaligned q: the same with m16n8k8 tf32.
//    The tcgen05/TMEM version of these kernels is the next step.
//  * attn_wide_kernel<DW, NW, KT> : one WIDE head (CIFAR d = 256, cin256 d = 384 / 576 / 960): the warps of a CTA split
//    the head dimension; same fp16-split tensor-core arithmetic.
//  * attn_generic_kernel<G> : any other head dim or unaligned operands: G lanes share one query, FFMA only.
#include <cuda_fp16.h>

#include "ctx.h"

namespace tfmq {

struct AttnP {
  tfmq_attn_desc a;
  int tk_tile;
};

// ------------------------------------------------------------------ generic FFMA kernel
template <int G>
__global__ void __launch_bounds__(128) attn_generic_kernel(const AttnP P) {
  extern __shared__ float sm[];
  const tfmq_attn_desc& a = P.a;
  const int d = a.d, TK = P.tk_tile;
  const int dp = d / G;  // dims per lane, <= 32, lane g owns dims g + G*i
  float* Ks = sm;
  float* Vs = sm + (size_t)TK * d;
  const int bh = blockIdx.y, b = bh / a.heads, h = bh % a.heads;
  constexpr int QPC = 128 / G;
  const int qi = blockIdx.x * QPC + threadIdx.x / G;
  const int g = threadIdx.x % G;
  const bool qvalid = qi < a.tq;
  const float* qp = a.q + (long long)b * a.q_sb + (long long)h * a.q_sh + (long long)(qvalid ? qi : 0) * a.q_st;
  const float* kp = a.k + (long long)b * a.k_sb + (long long)h * a.k_sh;
  const float* vp = a.v + (long long)b * a.v_sb + (long long)h * a.v_sh;
  float q[32], o[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    q[i] = (i < dp) ? qp[g + G * i] : 0.f;
    o[i] = 0.f;
  }
  float m = -INFINITY, l = 0.f;
  for (int k0 = 0; k0 < a.tk; k0 += TK) {
    __syncthreads();
    const int kn = min(TK, a.tk - k0);
    for (int i = threadIdx.x; i < TK * d; i += 128) {
      const int j = i / d, c = i - j * d;
      float kv = 0.f, vv = 0.f;
      if (j < kn) {
        kv = kp[(long long)(k0 + j) * a.k_st + c];
        vv = vp[(long long)(k0 + j) * a.v_st + c];
      }
      Ks[i] = kv;
      Vs[i] = vv;
    }
    __syncthreads();
    for (int j0 = 0; j0 < kn; j0 += 4) {
      float s[4];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const float* kr = Ks + (size_t)min(j0 + jj, TK - 1) * d + g;
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < dp) acc = fmaf(q[i], kr[G * i], acc);
#pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        s[jj] = (j0 + jj < kn) ? acc * a.scale : -INFINITY;
      }
      const float mnew = fmaxf(fmaxf(fmaxf(s[0], s[1]), fmaxf(s[2], s[3])), m);
      const float corr = expf(m - mnew);
      float p[4];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) p[jj] = expf(s[jj] - mnew);
      l = l * corr + ((p[0] + p[1]) + (p[2] + p[3]));
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if (i < dp) {
          float acc = o[i] * corr;
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) acc = fmaf(p[jj], Vs[(size_t)min(j0 + jj, TK - 1) * d + g + G * i], acc);
          o[i] = acc;
        }
      }
      m = mnew;
    }
  }
  if (qvalid) {
    float* op = a.o + (long long)b * a.o_sb + (long long)h * a.o_sh + (long long)qi * a.o_st;
    const float inv = 1.f / l;
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < dp) op[g + G * i] = o[i] * inv;
  }
}

// ------------------------------------------------------------------ tensor-core kernel
// CTA = 4 warps x 16 queries; K/V tiles of 64 keys staged in smem as tf32 hi / lo planes with
// row pitch D+4 floats (conflict-free fragment reads).  S = Q K^T and O += P V each as
// hi*hi + lo*hi + hi*lo.  The softmax probabilities come out of the S accumulator fragment in
// columns (2t, 2t+1); the PV MMA wants k-indices (t, t+4), so key (2t) is presented as k-index t
// and key (2t+1) as k-index t+4 -- a permutation of the summation order only.
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split_hi_lo(float v, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(v) & 0xFFFFE000u;
  lo = __float_as_uint(v - __uint_as_float(hi)) & 0xFFFFE000u;
}

constexpr int ATT_TK = 64;

template <int D>
__global__ void __launch_bounds__(128) attn_mma_kernel(const AttnP P) {
  constexpr int PITCH = D + 4;
  constexpr int KS = D / 8;   // k-steps of QK^T, n-tiles of PV
  extern __shared__ float sm[];
  float* Khi = sm;
  float* Klo = Khi + ATT_TK * PITCH;
  float* Vhi = Klo + ATT_TK * PITCH;
  float* Vlo = Vhi + ATT_TK * PITCH;
  const tfmq_attn_desc& a = P.a;
  const int bh = blockIdx.y, b = bh / a.heads, h = bh % a.heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int q0 = blockIdx.x * 64 + warp * 16;
  const float* qb = a.q + (long long)b * a.q_sb + (long long)h * a.q_sh;
  const float* kb = a.k + (long long)b * a.k_sb + (long long)h * a.k_sh;
  const float* vb = a.v + (long long)b * a.v_sb + (long long)h * a.v_sh;

  // Q fragments (A operand, rows g / g+8, cols t / t+4 of each 8-wide k-step), pre-scaled
  uint32_t qh[KS][4], ql[KS][4];
  {
    const int r0 = min(q0 + g, a.tq - 1), r1 = min(q0 + g + 8, a.tq - 1);
    const float* p0 = qb + (long long)r0 * a.q_st;
    const float* p1 = qb + (long long)r1 * a.q_st;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      split_hi_lo(p0[ks * 8 + t] * a.scale, qh[ks][0], ql[ks][0]);
      split_hi_lo(p1[ks * 8 + t] * a.scale, qh[ks][1], ql[ks][1]);
      split_hi_lo(p0[ks * 8 + t + 4] * a.scale, qh[ks][2], ql[ks][2]);
      split_hi_lo(p1[ks * 8 + t + 4] * a.scale, qh[ks][3], ql[ks][3]);
    }
  }
  float o[KS][4];
#pragma unroll
  for (int i = 0; i < KS; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  for (int k0 = 0; k0 < a.tk; k0 += ATT_TK) {
    __syncthreads();
    const int kn = min(ATT_TK, a.tk - k0);
    for (int i = threadIdx.x; i < ATT_TK * (D / 4); i += 128) {
      const int j = i / (D / 4), c4 = (i - j * (D / 4)) * 4;
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (j < kn) {
        kv = *reinterpret_cast<const float4*>(kb + (long long)(k0 + j) * a.k_st + c4);
        vv = *reinterpret_cast<const float4*>(vb + (long long)(k0 + j) * a.v_st + c4);
      }
      uint32_t hh[4], ll[4];
      split_hi_lo(kv.x, hh[0], ll[0]), split_hi_lo(kv.y, hh[1], ll[1]);
      split_hi_lo(kv.z, hh[2], ll[2]), split_hi_lo(kv.w, hh[3], ll[3]);
      *reinterpret_cast<uint4*>(Khi + j * PITCH + c4) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
      *reinterpret_cast<uint4*>(Klo + j * PITCH + c4) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
      split_hi_lo(vv.x, hh[0], ll[0]), split_hi_lo(vv.y, hh[1], ll[1]);
      split_hi_lo(vv.z, hh[2], ll[2]), split_hi_lo(vv.w, hh[3], ll[3]);
      *reinterpret_cast<uint4*>(Vhi + j * PITCH + c4) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
      *reinterpret_cast<uint4*>(Vlo + j * PITCH + c4) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
    }
    __syncthreads();

    // ---- S = Q K^T : 8 n-tiles of 8 keys
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      const uint32_t* kh = reinterpret_cast<const uint32_t*>(Khi) + (nt * 8 + g) * PITCH + t;
      const uint32_t* kl = reinterpret_cast<const uint32_t*>(Klo) + (nt * 8 + g) * PITCH + t;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const uint32_t bh0 = kh[ks * 8], bh1 = kh[ks * 8 + 4];
        const uint32_t bl0 = kl[ks * 8], bl1 = kl[ks * 8 + 4];
        mma_tf32(s[nt], ql[ks], bh0, bh1);
        mma_tf32(s[nt], qh[ks], bl0, bl1);
        mma_tf32(s[nt], qh[ks], bh0, bh1);
      }
    }
    // ---- mask the tail, online softmax (rows g and g+8; columns 2t, 2t+1 of each n-tile)
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int c = nt * 8 + 2 * t;
      if (c >= kn) s[nt][0] = -INFINITY, s[nt][2] = -INFINITY;
      if (c + 1 >= kn) s[nt][1] = -INFINITY, s[nt][3] = -INFINITY;
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float c0 = expf(m0 - mx0), c1 = expf(m1 - mx1);
    m0 = mx0, m1 = mx1;
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = expf(s[nt][0] - mx0);
      s[nt][1] = expf(s[nt][1] - mx0);
      s[nt][2] = expf(s[nt][2] - mx1);
      s[nt][3] = expf(s[nt][3] - mx1);
      rs0 += s[nt][0] + s[nt][1];
      rs1 += s[nt][2] + s[nt][3];
    }
    l0 = l0 * c0 + rs0;
    l1 = l1 * c1 + rs1;
#pragma unroll
    for (int i = 0; i < KS; ++i) {
      o[i][0] *= c0, o[i][1] *= c0;
      o[i][2] *= c1, o[i][3] *= c1;
    }
    // ---- O += P V : k-step = one n-tile of S (8 keys), n-tiles over D
#pragma unroll
    for (int kt = 0; kt < 8; ++kt) {
      uint32_t ph[4], pl[4];
      // A fragment: a0=(row g, k t) a1=(row g+8, k t) a2=(row g, k t+4) a3=(row g+8, k t+4)
      // with k-index t <-> key 2t and k-index t+4 <-> key 2t+1
      split_hi_lo(s[kt][0], ph[0], pl[0]);
      split_hi_lo(s[kt][2], ph[1], pl[1]);
      split_hi_lo(s[kt][1], ph[2], pl[2]);
      split_hi_lo(s[kt][3], ph[3], pl[3]);
      const uint32_t* vh = reinterpret_cast<const uint32_t*>(Vhi) + (kt * 8 + 2 * t) * PITCH + g;
      const uint32_t* vl = reinterpret_cast<const uint32_t*>(Vlo) + (kt * 8 + 2 * t) * PITCH + g;
#pragma unroll
      for (int nt = 0; nt < KS; ++nt) {
        const uint32_t bh0 = vh[nt * 8], bh1 = vh[PITCH + nt * 8];
        const uint32_t bl0 = vl[nt * 8], bl1 = vl[PITCH + nt * 8];
        mma_tf32(o[nt], pl, bh0, bh1);
        mma_tf32(o[nt], ph, bl0, bl1);
        mma_tf32(o[nt], ph, bh0, bh1);
      }
    }
  }
  // ---- normalise and store: C fragment rows g / g+8, cols 2t, 2t+1 of each 8-wide n-tile
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  float* ob = a.o + (long long)b * a.o_sb + (long long)h * a.o_sh;
  const int r0 = q0 + g, r1 = q0 + g + 8;
#pragma unroll
  for (int nt = 0; nt < KS; ++nt) {
    if (r0 < a.tq)
      *reinterpret_cast<float2*>(ob + (long long)r0 * a.o_st + nt * 8 + 2 * t) =
          make_float2(o[nt][0] * i0, o[nt][1] * i0);
    if (r1 < a.tq)
      *reinterpret_cast<float2*>(ob + (long long)r1 * a.o_st + nt * 8 + 2 * t) =
          make_float2(o[nt][2] * i1, o[nt][3] * i1);
  }
}

// ------------------------------------------------------------------ tensor-core kernel, fp16 split
// Same algorithm with m16n8k16 fp16 MMAs: every fp32 operand is split into two fp16 terms
// (hi = half(x), lo = half(x - hi): 22 significand bits, the same as the tf32 hi/lo split) and each product
// is hi*hi + lo*hi + hi*lo with fp32 accumulation.  One k16 instruction covers twice the depth of a k8 tf32
// instruction, so the tensor-pipe work halves.  K and V are both staged [key][D+8] (fp16 hi / lo planes, conflict-free
// 8-byte stores); the K fragments of S = Q K^T are plain 32-bit loads, the V fragments of O += P V come transposed out of
// `ldmatrix.x4.trans` (no scattered 2-byte transpose stores).  The raw fp32 K/V rows of the NEXT key tile are fetched into
// registers while the current tile is being multiplied.  The S accumulator fragment is directly the A fragment of the PV
// product (no permutation needed for k16).
__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split2_h(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x0, x1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// the saturating split tfmq_act_prepare writes (elementwise.cu::split_h16x4), for two values: bits of (hi0, hi1), (lo0, lo1)
__device__ __forceinline__ void split2_sat(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  uint16_t h0, h1, l0, l1;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h0) : "f"(x0));
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h1) : "f"(x1));
  const float r0 = x0 - __half2float(__ushort_as_half(h0)), r1 = x1 - __half2float(__ushort_as_half(h1));
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(l0) : "f"(r0));
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(l1) : "f"(r1));
  hi = (uint32_t)h0 | ((uint32_t)h1 << 16);
  lo = (uint32_t)l0 | ((uint32_t)l1 << 16);
}
// one (row, column pair) of the output: fp32, or the fp16 hi / lo planes a floating-point conv reads next
__device__ __forceinline__ void store_o2(const tfmq_attn_desc& a, long long off, float x0, float x1) {
  if (a.o_hi) {
    uint32_t hi, lo;
    split2_sat(x0, x1, hi, lo);
    *reinterpret_cast<uint32_t*>(static_cast<__half*>(a.o_hi) + off) = hi;
    *reinterpret_cast<uint32_t*>(static_cast<__half*>(a.o_lo) + off) = lo;
  } else {
    *reinterpret_cast<float2*>(a.o + off) = make_float2(x0, x1);
  }
}
__device__ __forceinline__ void split1_h(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn(x - __half2float(hi));
}

// 2^x on the SFU (MUFU.EX2, about 2 ulp); ex2(-inf) = 0
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// four 8x8 b16 matrices, transposed on the way out: thread (g, t) of matrix i gets elements (row 2t, col g), (row 2t+1, col g)
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
  const uint32_t addr = static_cast<uint32_t>(__cvta_generic_to_shared(smem_row));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

template <int D>
__global__ void __launch_bounds__(128) attn_h16_kernel(const AttnP P) {
  constexpr int KP = D + 8;          // K / V row pitch (halves): 16-byte aligned rows, conflict-free ldmatrix
  constexpr int KS = D / 16;         // k-steps of Q K^T
  constexpr int NT = D / 8;          // n-tiles of P V
  constexpr int ITEMS = ATT_TK * (D / 4) / 128;    // float4 pieces of a K (or V) tile per thread
  constexpr bool PREFETCH = ITEMS <= 4;            // registers for the next tile's raw rows (D = 32)
  static_assert(ATT_TK * (D / 4) % 128 == 0 && NT % 2 == 0, "tile geometry");
  extern __shared__ __half smh[];
  __half* Khi = smh;
  __half* Klo = Khi + ATT_TK * KP;
  __half* Vhi = Klo + ATT_TK * KP;   // [key][KP] like K
  __half* Vlo = Vhi + ATT_TK * KP;
  const tfmq_attn_desc& a = P.a;
  const int dr = a.d;                // real head dim <= D (d = 40 runs zero-padded as D = 48)
  const int bh = blockIdx.y, b = bh / a.heads, h = bh % a.heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int q0 = blockIdx.x * 64 + warp * 16;
  const float* qb = a.q + (long long)b * a.q_sb + (long long)h * a.q_sh;
  const float* kb = a.k + (long long)b * a.k_sb + (long long)h * a.k_sh;
  const float* vb = a.v + (long long)b * a.v_sb + (long long)h * a.v_sh;

  uint32_t qh[KS][4], ql[KS][4];
  const float qs = a.scale * 1.4426950408889634f;
  {
    const int r0 = min(q0 + g, a.tq - 1), r1 = min(q0 + g + 8, a.tq - 1);
    const float* p0 = qb + (long long)r0 * a.q_st;
    const float* p1 = qb + (long long)r1 * a.q_st;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const float2 z2 = make_float2(0.f, 0.f);
      const int ca = ks * 16 + 2 * t, cb = ca + 8;
      const float2 x0 = ca < dr ? *reinterpret_cast<const float2*>(p0 + ca) : z2;
      const float2 x1 = ca < dr ? *reinterpret_cast<const float2*>(p1 + ca) : z2;
      const float2 x2 = cb < dr ? *reinterpret_cast<const float2*>(p0 + cb) : z2;
      const float2 x3 = cb < dr ? *reinterpret_cast<const float2*>(p1 + cb) : z2;
      // scores are kept in the log2 domain (scale * log2 e folded into q): softmax is then one MUFU.EX2 per element
      split2_h(x0.x * qs, x0.y * qs, qh[ks][0], ql[ks][0]);
      split2_h(x1.x * qs, x1.y * qs, qh[ks][1], ql[ks][1]);
      split2_h(x2.x * qs, x2.y * qs, qh[ks][2], ql[ks][2]);
      split2_h(x3.x * qs, x3.y * qs, qh[ks][3], ql[ks][3]);
    }
  }
  float o[NT][4];
#pragma unroll
  for (int i = 0; i < NT; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  float4 kreg[PREFETCH ? ITEMS : 1], vreg[PREFETCH ? ITEMS : 1];
  auto fetch = [&](int k0, int it, float4& kv, float4& vv) {
    const int i = threadIdx.x + it * 128;
    const int j = i / (D / 4), c4 = (i - j * (D / 4)) * 4;
    kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
    if (k0 + j < a.tk && c4 < dr) {
      kv = *reinterpret_cast<const float4*>(kb + (long long)(k0 + j) * a.k_st + c4);
      vv = *reinterpret_cast<const float4*>(vb + (long long)(k0 + j) * a.v_st + c4);
    }
  };
  auto stage = [&](int it, const float4& kv, const float4& vv) {
    const int i = threadIdx.x + it * 128;
    const int j = i / (D / 4), c4 = (i - j * (D / 4)) * 4;
    uint32_t h01, l01, h23, l23;
    split2_h(kv.x, kv.y, h01, l01);
    split2_h(kv.z, kv.w, h23, l23);
    *reinterpret_cast<uint2*>(Khi + j * KP + c4) = make_uint2(h01, h23);
    *reinterpret_cast<uint2*>(Klo + j * KP + c4) = make_uint2(l01, l23);
    split2_h(vv.x, vv.y, h01, l01);
    split2_h(vv.z, vv.w, h23, l23);
    *reinterpret_cast<uint2*>(Vhi + j * KP + c4) = make_uint2(h01, h23);
    *reinterpret_cast<uint2*>(Vlo + j * KP + c4) = make_uint2(l01, l23);
  };
  if (PREFETCH) {
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) fetch(0, it, kreg[it], vreg[it]);
  }
  // ldmatrix row of this lane: matrix (lane >> 3) = (key half, n-tile parity), row (lane & 7)
  const int lm_key = ((lane >> 3) & 1) * 8 + (lane & 7), lm_dim = (lane >> 4) * 8;

  for (int k0 = 0; k0 < a.tk; k0 += ATT_TK) {
    __syncthreads();
    const int kn = min(ATT_TK, a.tk - k0);
    if (PREFETCH) {
#pragma unroll
      for (int it = 0; it < ITEMS; ++it) stage(it, kreg[it], vreg[it]);
    } else {
#pragma unroll 2
      for (int it = 0; it < ITEMS; ++it) {
        float4 kv, vv;
        fetch(k0, it, kv, vv);
        stage(it, kv, vv);
      }
    }
    __syncthreads();
    if (PREFETCH && k0 + ATT_TK < a.tk) {      // the next tile's rows travel while this one is multiplied
#pragma unroll
      for (int it = 0; it < ITEMS; ++it) fetch(k0 + ATT_TK, it, kreg[it], vreg[it]);
    }

    // ---- S = Q K^T
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      const uint32_t* kh = reinterpret_cast<const uint32_t*>(Khi + (nt * 8 + g) * KP) + t;
      const uint32_t* kl = reinterpret_cast<const uint32_t*>(Klo + (nt * 8 + g) * KP) + t;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const uint32_t bh0 = kh[ks * 8], bh1 = kh[ks * 8 + 4];
        const uint32_t bl0 = kl[ks * 8], bl1 = kl[ks * 8 + 4];
        mma_f16(s[nt], ql[ks], bh0, bh1);
        mma_f16(s[nt], qh[ks], bl0, bl1);
        mma_f16(s[nt], qh[ks], bh0, bh1);
      }
    }
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int c = nt * 8 + 2 * t;
      if (c >= kn) s[nt][0] = -INFINITY, s[nt][2] = -INFINITY;
      if (c + 1 >= kn) s[nt][1] = -INFINITY, s[nt][3] = -INFINITY;
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float c0 = ex2f(m0 - mx0), c1 = ex2f(m1 - mx1);
    m0 = mx0, m1 = mx1;
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = ex2f(s[nt][0] - mx0);
      s[nt][1] = ex2f(s[nt][1] - mx0);
      s[nt][2] = ex2f(s[nt][2] - mx1);
      s[nt][3] = ex2f(s[nt][3] - mx1);
      rs0 += s[nt][0] + s[nt][1];
      rs1 += s[nt][2] + s[nt][3];
    }
    l0 = l0 * c0 + rs0;
    l1 = l1 * c1 + rs1;
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      o[i][0] *= c0, o[i][1] *= c0;
      o[i][2] *= c1, o[i][3] *= c1;
    }
    // ---- O += P V : k-step = 16 keys = two n-tiles of S
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) {
      uint32_t ph[4], pl[4];
      split2_h(s[2 * kt][0], s[2 * kt][1], ph[0], pl[0]);
      split2_h(s[2 * kt][2], s[2 * kt][3], ph[1], pl[1]);
      split2_h(s[2 * kt + 1][0], s[2 * kt + 1][1], ph[2], pl[2]);
      split2_h(s[2 * kt + 1][2], s[2 * kt + 1][3], ph[3], pl[3]);
#pragma unroll
      for (int nt = 0; nt < NT; nt += 2) {
        // matrices: (keys kt*16 + 0..7, dims nt*8..), (keys +8.., same dims), then the same two for n-tile nt + 1
        uint32_t bh[4], bl[4];
        ldmatrix_x4_trans(bh, Vhi + (kt * 16 + lm_key) * KP + nt * 8 + lm_dim);
        ldmatrix_x4_trans(bl, Vlo + (kt * 16 + lm_key) * KP + nt * 8 + lm_dim);
        mma_f16(o[nt], pl, bh[0], bh[1]);
        mma_f16(o[nt], ph, bl[0], bl[1]);
        mma_f16(o[nt], ph, bh[0], bh[1]);
        mma_f16(o[nt + 1], pl, bh[2], bh[3]);
        mma_f16(o[nt + 1], ph, bl[2], bl[3]);
        mma_f16(o[nt + 1], ph, bh[2], bh[3]);
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  const long long ob = (long long)b * a.o_sb + (long long)h * a.o_sh;
  const int r0 = q0 + g, r1 = q0 + g + 8;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    if (nt * 8 + 2 * t >= dr) continue;
    if (r0 < a.tq) store_o2(a, ob + (long long)r0 * a.o_st + nt * 8 + 2 * t, o[nt][0] * i0, o[nt][1] * i0);
    if (r1 < a.tq) store_o2(a, ob + (long long)r1 * a.o_st + nt * 8 + 2 * t, o[nt][2] * i1, o[nt][3] * i1);
  }
}

// ------------------------------------------------------------------ tensor-core kernel for WIDE heads
// One head of 256 ... 960 channels (cin256's single-head SpatialTransformers, CIFAR's AttnBlock): the accumulators of
// a whole head do not fit one warp's registers, so the NW warps of a CTA split the HEAD DIMENSION instead of the
// queries.  A CTA owns 16 queries; warp w owns channels [w DW, (w + 1) DW):
//   * S = Q K^T : every warp multiplies its channel slice (a K-split of the product); the partial 16 x KT tiles go
//     through shared memory and every warp sums all of them in the same order, so all warps hold the SAME scores
//     and run the same online softmax (identical m, l);
//   * O += P V  : every warp multiplies P with its DW columns of V and keeps only those DW accumulator columns.
// No product is computed twice.  Same arithmetic as attn_h16_kernel (fp16 hi/lo split, 3 products, log2-domain softmax).
template <int DW, int NW, int KT, int QG>
__global__ void __launch_bounds__(NW * QG * 32) attn_wide_kernel(const AttnP P) {
  constexpr int KS = DW / 16;        // k-steps of this warp's slice of Q K^T
  constexpr int NT = DW / 8;         // n-tiles of this warp's slice of P V
  constexpr int ST = KT / 8;         // n-tiles of the score tile
  constexpr int THREADS = NW * QG * 32;   // QG groups of NW warps, one 16-query tile each, sharing the staged K / V tile
  static_assert(DW % 16 == 0 && NT % 2 == 0 && KT % 16 == 0, "geometry");
  const tfmq_attn_desc& a = P.a;
  const int d = a.d;                 // == DW * NW
  const int KP = d + 8;              // row pitch (halves): 16-byte aligned rows, conflict-free fragment loads
  extern __shared__ __half smh[];
  __half* Khi = smh;
  __half* Klo = Khi + KT * KP;
  __half* Vhi = Klo + KT * KP;
  __half* Vlo = Vhi + KT * KP;
  float* Sp_all = reinterpret_cast<float*>(Vlo + KT * KP);  // [QG][NW][16][KT] partial scores
  const int bh = blockIdx.y, b = bh / a.heads, h = bh % a.heads;
  const int qg = (threadIdx.x >> 5) / NW, warp = (threadIdx.x >> 5) % NW, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int q0 = (blockIdx.x * QG + qg) * 16;
  float* Sp = Sp_all + qg * NW * 16 * KT;
  const int dw0 = warp * DW;
  const float* qb = a.q + (long long)b * a.q_sb + (long long)h * a.q_sh;
  const float* kb = a.k + (long long)b * a.k_sb + (long long)h * a.k_sh;
  const float* vb = a.v + (long long)b * a.v_sb + (long long)h * a.v_sh;

  uint32_t qh[KS][4], ql[KS][4];
  const float qs = a.scale * 1.4426950408889634f;
  {
    const int r0 = min(q0 + g, a.tq - 1), r1 = min(q0 + g + 8, a.tq - 1);
    const float* p0 = qb + (long long)r0 * a.q_st + dw0;
    const float* p1 = qb + (long long)r1 * a.q_st + dw0;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const float2 x0 = *reinterpret_cast<const float2*>(p0 + ks * 16 + 2 * t);
      const float2 x1 = *reinterpret_cast<const float2*>(p1 + ks * 16 + 2 * t);
      const float2 x2 = *reinterpret_cast<const float2*>(p0 + ks * 16 + 2 * t + 8);
      const float2 x3 = *reinterpret_cast<const float2*>(p1 + ks * 16 + 2 * t + 8);
      split2_h(x0.x * qs, x0.y * qs, qh[ks][0], ql[ks][0]);
      split2_h(x1.x * qs, x1.y * qs, qh[ks][1], ql[ks][1]);
      split2_h(x2.x * qs, x2.y * qs, qh[ks][2], ql[ks][2]);
      split2_h(x3.x * qs, x3.y * qs, qh[ks][3], ql[ks][3]);
    }
  }
  float o[NT][4];
#pragma unroll
  for (int i = 0; i < NT; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const int lm_key = ((lane >> 3) & 1) * 8 + (lane & 7), lm_dim = (lane >> 4) * 8;
  const int vec_per_row = d >> 2;

  for (int k0 = 0; k0 < a.tk; k0 += KT) {
    __syncthreads();
    const int kn = min(KT, a.tk - k0);
    // stage the K and V rows of this key tile (all channels) as fp16 hi / lo planes
    for (int i = threadIdx.x; i < KT * vec_per_row; i += THREADS) {
      const int j = i / vec_per_row, c4 = (i - j * vec_per_row) * 4;
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (j < kn) {
        kv = *reinterpret_cast<const float4*>(kb + (long long)(k0 + j) * a.k_st + c4);
        vv = *reinterpret_cast<const float4*>(vb + (long long)(k0 + j) * a.v_st + c4);
      }
      uint32_t h01, l01, h23, l23;
      split2_h(kv.x, kv.y, h01, l01);
      split2_h(kv.z, kv.w, h23, l23);
      *reinterpret_cast<uint2*>(Khi + j * KP + c4) = make_uint2(h01, h23);
      *reinterpret_cast<uint2*>(Klo + j * KP + c4) = make_uint2(l01, l23);
      split2_h(vv.x, vv.y, h01, l01);
      split2_h(vv.z, vv.w, h23, l23);
      *reinterpret_cast<uint2*>(Vhi + j * KP + c4) = make_uint2(h01, h23);
      *reinterpret_cast<uint2*>(Vlo + j * KP + c4) = make_uint2(l01, l23);
    }
    __syncthreads();

    // ---- partial S over this warp's channel slice
    float s[ST][4];
#pragma unroll
    for (int nt = 0; nt < ST; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      const uint32_t* kh = reinterpret_cast<const uint32_t*>(Khi + (nt * 8 + g) * KP + dw0) + t;
      const uint32_t* kl = reinterpret_cast<const uint32_t*>(Klo + (nt * 8 + g) * KP + dw0) + t;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const uint32_t bh0 = kh[ks * 8], bh1 = kh[ks * 8 + 4];
        const uint32_t bl0 = kl[ks * 8], bl1 = kl[ks * 8 + 4];
        mma_f16(s[nt], ql[ks], bh0, bh1);
        mma_f16(s[nt], qh[ks], bl0, bl1);
        mma_f16(s[nt], qh[ks], bh0, bh1);
      }
    }
    // exchange: Sp[w][row][col], fragment (rows g, g + 8; cols nt * 8 + 2t, + 1)
    {
      float* mine = Sp + warp * 16 * KT;
#pragma unroll
      for (int nt = 0; nt < ST; ++nt) {
        *reinterpret_cast<float2*>(mine + g * KT + nt * 8 + 2 * t) = make_float2(s[nt][0], s[nt][1]);
        *reinterpret_cast<float2*>(mine + (g + 8) * KT + nt * 8 + 2 * t) = make_float2(s[nt][2], s[nt][3]);
      }
    }
    __syncthreads();
#pragma unroll
    for (int nt = 0; nt < ST; ++nt) {
      float2 a0 = make_float2(0.f, 0.f), a1 = a0;
      for (int w = 0; w < NW; ++w) {         // the same order in every warp: identical sums
        const float2 p0 = *reinterpret_cast<const float2*>(Sp + w * 16 * KT + g * KT + nt * 8 + 2 * t);
        const float2 p1 = *reinterpret_cast<const float2*>(Sp + w * 16 * KT + (g + 8) * KT + nt * 8 + 2 * t);
        a0.x += p0.x, a0.y += p0.y, a1.x += p1.x, a1.y += p1.y;
      }
      s[nt][0] = a0.x, s[nt][1] = a0.y, s[nt][2] = a1.x, s[nt][3] = a1.y;
    }
    // ---- online softmax (log2 domain), identical in every warp
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int nt = 0; nt < ST; ++nt) {
      const int c = nt * 8 + 2 * t;
      if (c >= kn) s[nt][0] = -INFINITY, s[nt][2] = -INFINITY;
      if (c + 1 >= kn) s[nt][1] = -INFINITY, s[nt][3] = -INFINITY;
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float c0 = ex2f(m0 - mx0), c1 = ex2f(m1 - mx1);
    m0 = mx0, m1 = mx1;
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < ST; ++nt) {
      s[nt][0] = ex2f(s[nt][0] - mx0);
      s[nt][1] = ex2f(s[nt][1] - mx0);
      s[nt][2] = ex2f(s[nt][2] - mx1);
      s[nt][3] = ex2f(s[nt][3] - mx1);
      rs0 += s[nt][0] + s[nt][1];
      rs1 += s[nt][2] + s[nt][3];
    }
    l0 = l0 * c0 + rs0;
    l1 = l1 * c1 + rs1;
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      o[i][0] *= c0, o[i][1] *= c0;
      o[i][2] *= c1, o[i][3] *= c1;
    }
    // ---- O[:, slice] += P V[:, slice]
#pragma unroll
    for (int kt = 0; kt < KT / 16; ++kt) {
      uint32_t ph[4], pl[4];
      split2_h(s[2 * kt][0], s[2 * kt][1], ph[0], pl[0]);
      split2_h(s[2 * kt][2], s[2 * kt][3], ph[1], pl[1]);
      split2_h(s[2 * kt + 1][0], s[2 * kt + 1][1], ph[2], pl[2]);
      split2_h(s[2 * kt + 1][2], s[2 * kt + 1][3], ph[3], pl[3]);
#pragma unroll
      for (int nt = 0; nt < NT; nt += 2) {
        uint32_t bh[4], bl[4];
        ldmatrix_x4_trans(bh, Vhi + (kt * 16 + lm_key) * KP + dw0 + nt * 8 + lm_dim);
        ldmatrix_x4_trans(bl, Vlo + (kt * 16 + lm_key) * KP + dw0 + nt * 8 + lm_dim);
        mma_f16(o[nt], pl, bh[0], bh[1]);
        mma_f16(o[nt], ph, bl[0], bl[1]);
        mma_f16(o[nt], ph, bh[0], bh[1]);
        mma_f16(o[nt + 1], pl, bh[2], bh[3]);
        mma_f16(o[nt + 1], ph, bl[2], bl[3]);
        mma_f16(o[nt + 1], ph, bh[2], bh[3]);
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  const long long ob = (long long)b * a.o_sb + (long long)h * a.o_sh + dw0;
  const int r0 = q0 + g, r1 = q0 + g + 8;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    if (r0 < a.tq) store_o2(a, ob + (long long)r0 * a.o_st + nt * 8 + 2 * t, o[nt][0] * i0, o[nt][1] * i0);
    if (r1 < a.tq) store_o2(a, ob + (long long)r1 * a.o_st + nt * 8 + 2 * t, o[nt][2] * i1, o[nt][3] * i1);
  }
}

template <int DW, int NW, int KT, int QG>
static int launch_wide(tfmq_ctx* ctx, const AttnP& P, cudaStream_t st) {
  const size_t smem = (size_t)(4 * KT * (P.a.d + 8)) * sizeof(__half) + (size_t)QG * NW * 16 * KT * sizeof(float);
  auto kern = attn_wide_kernel<DW, NW, KT, QG>;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return tfmq_fail(ctx, TFMQ_ERR_CUDA, "attention: smem attr: %s", cudaGetErrorString(e));
    smem_set = smem;
  }
  dim3 grid((P.a.tq + 16 * QG - 1) / (16 * QG), P.a.b * P.a.heads);
  kern<<<grid, NW * QG * 32, smem, st>>>(P);
  TFMQ_LAUNCH_CHECK("attention_wide");
  return TFMQ_OK;
}

template <int D>
static int launch_h16(tfmq_ctx* ctx, const AttnP& P, cudaStream_t st) {
  const size_t smem = (size_t)(4 * ATT_TK * (D + 8)) * sizeof(__half);
  auto kern = attn_h16_kernel<D>;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return tfmq_fail(ctx, TFMQ_ERR_CUDA, "attention: smem attr: %s", cudaGetErrorString(e));
    smem_set = smem;
  }
  dim3 grid((P.a.tq + 63) / 64, P.a.b * P.a.heads);
  kern<<<grid, 128, smem, st>>>(P);
  TFMQ_LAUNCH_CHECK("attention_h16");
  return TFMQ_OK;
}

template <int D>
static int launch_mma(tfmq_ctx* ctx, const AttnP& P, cudaStream_t st) {
  const size_t smem = (size_t)4 * ATT_TK * (D + 4) * sizeof(float);
  auto kern = attn_mma_kernel<D>;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return tfmq_fail(ctx, TFMQ_ERR_CUDA, "attention: smem attr: %s", cudaGetErrorString(e));
    smem_set = smem;
  }
  dim3 grid((P.a.tq + 63) / 64, P.a.b * P.a.heads);
  kern<<<grid, 128, smem, st>>>(P);
  TFMQ_LAUNCH_CHECK("attention_mma");
  return TFMQ_OK;
}

template <int G>
static int launch_generic(tfmq_ctx* ctx, AttnP& P, cudaStream_t st) {
  int tk = 12288 / P.a.d;
  if (tk > 32) tk = 32;
  if (tk < 4) tk = 4;
  tk &= ~3;
  P.tk_tile = tk;
  const size_t smem = (size_t)2 * tk * P.a.d * sizeof(float);
  auto kern = attn_generic_kernel<G>;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return tfmq_fail(ctx, TFMQ_ERR_CUDA, "attention: smem attr: %s", cudaGetErrorString(e));
    smem_set = smem;
  }
  constexpr int QPC = 128 / G;
  dim3 grid((P.a.tq + QPC - 1) / QPC, P.a.b * P.a.heads);
  kern<<<grid, 128, smem, st>>>(P);
  TFMQ_LAUNCH_CHECK("attention_generic");
  return TFMQ_OK;
}

}  // namespace tfmq

using namespace tfmq;

extern "C" int tfmq_attention(tfmq_ctx* ctx, const tfmq_attn_desc* d, void* stream) {
  if (!ctx) return TFMQ_ERR_ARG;
  TFMQ_REQUIRE(d && d->q && d->k && d->v && (d->o || d->o_hi), TFMQ_ERR_ARG, "attention: null pointer");
  TFMQ_REQUIRE(!d->o_hi || (d->o_lo && (((uintptr_t)d->o_hi | (uintptr_t)d->o_lo) & 3) == 0), TFMQ_ERR_ARG,
               "attention: o_lo missing or fp16 planes not 4-byte aligned");
  TFMQ_REQUIRE(d->d > 0 && d->d <= 1024, TFMQ_ERR_SHAPE, "attention: head dim %d", d->d);
  if (d->b == 0 || d->heads == 0 || d->tq == 0) return TFMQ_OK;
  TFMQ_REQUIRE(d->tk > 0, TFMQ_ERR_SHAPE, "attention: empty key set");
  TFMQ_REQUIRE((long long)d->b * d->heads <= 65535, TFMQ_ERR_SHAPE, "attention: b*heads > 65535");
  AttnP P;
  P.a = *d;
  P.tk_tile = ATT_TK;
  cudaStream_t st = tfmq_stream(stream);
  auto al16 = [](const void* p, int64_t s0, int64_t s1, int64_t s2) {
    return (((uintptr_t)p & 15) == 0) && s0 % 4 == 0 && s1 % 4 == 0 && s2 % 4 == 0;
  };
  const bool vec_ok = al16(d->k, d->k_sb, d->k_sh, d->k_st) && al16(d->v, d->v_sb, d->v_sh, d->v_st) &&
                      (((uintptr_t)d->o & 7) == 0) && d->o_sb % 2 == 0 && d->o_sh % 2 == 0 && d->o_st % 2 == 0;
  const bool planes = d->o_hi != nullptr;     // only the kernels below the next `if` write fp16 planes
  const bool q_ok = (((uintptr_t)d->q & 7) == 0) && d->q_sb % 2 == 0 && d->q_sh % 2 == 0 && d->q_st % 2 == 0;
  if (vec_ok) {
    switch (d->d) {
      case 32: if (q_ok) return launch_h16<32>(ctx, P, st); if (!planes) return launch_mma<32>(ctx, P, st); break;
      case 40: if (q_ok) return launch_h16<48>(ctx, P, st); if (!planes) return launch_mma<40>(ctx, P, st); break;   // zero-padded to 3 k16 steps
      case 64: if (q_ok) return launch_h16<64>(ctx, P, st); if (!planes) return launch_mma<64>(ctx, P, st); break;
      case 80: if (q_ok) return launch_h16<80>(ctx, P, st); if (!planes) return launch_mma<80>(ctx, P, st); break;
      case 160:
        if (q_ok) return launch_h16<160>(ctx, P, st);      // SD v1.4's deepest levels
        break;
      case 256: if (q_ok) return launch_wide<64, 4, 64, 2>(ctx, P, st); break;    // CIFAR AttnBlock (one head)
      case 384: if (q_ok) return launch_wide<96, 4, 48, 2>(ctx, P, st); break;    // cin256, 32x32
      case 512: if (q_ok) return launch_wide<128, 4, 32, 2>(ctx, P, st); break;   // first-stage decoder mid.attn_1 (one head)
      case 576: if (q_ok) return launch_wide<96, 6, 32, 1>(ctx, P, st); break;    // cin256, 16x16
      case 960: if (q_ok) return launch_wide<160, 6, 16, 1>(ctx, P, st); break;   // cin256, 8x8
      default: break;
    }
  }
  TFMQ_REQUIRE(!planes, TFMQ_ERR_SHAPE, "attention: fp16-plane output needs a tensor-core kernel (head dim %d, alignment)",
               d->d);
  int G = 1;
  while (G <= 32 && !(d->d % G == 0 && d->d / G <= 32)) G <<= 1;
  TFMQ_REQUIRE(G <= 32, TFMQ_ERR_SHAPE, "attention: head dim %d not supported", d->d);
  switch (G) {
    case 1: return launch_generic<1>(ctx, P, st);
    case 2: return launch_generic<2>(ctx, P, st);
    case 4: return launch_generic<4>(ctx, P, st);
    case 8: return launch_generic<8>(ctx, P, st);
    case 16: return launch_generic<16>(ctx, P, st);
    default: return launch_generic<32>(ctx, P, st);
  }
}
