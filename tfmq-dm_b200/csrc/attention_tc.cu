// Fused QK^T - softmax - PV attention on tcgen05 (5th-gen tensor cores, accumulators in TMEM), fp32-accurate.
//
// Replaces the attention core of QKVAttentionLegacy.forward (ldm/modules/diffusionmodules/openaimodel.py:383-405),
// QuantAttnBlock.forward (quant/quant_block.py:474-505) and cross_attn_forward (quant/quant_block.py:212-245), which the
// reference evaluates in fp32 (its q/k/v/softmax quantisers are inert).
//
// Operands arrive PRE-SPLIT as fp16 hi / lo planes (hi = half(x), lo = half(x - hi): 22 significand bits), written once
// per tensor by the producing convolution's epilogue (tfmq_conv_h16 `out_hi / out_lo`) or by tfmq_act_prepare -- not
// once per query block as in the mma.sync kernels of attention.cu.  Every product is hi*hi + lo*hi + hi*lo.
//
// One CTA = one (batch, head, 128-query block); two CTAs are resident per SM and fill each other's bubbles.
//   warps 0-3  softmax: thread r owns query row r = TMEM lane r.  Per key tile: tcgen05.ld S, running max in the log2
//              domain with LAZY rescaling (the accumulators are only rescaled when the max grew by more than 2^8),
//              p = ex2(s*c - m*c) on the SFU, p split into fp16 hi / lo and written back IN PLACE over S with tcgen05.st
//              (P never touches shared memory); the O accumulators (TMEM) are rescaled here when needed.
//   warp 4     TMA producer: Q once, then a ring of K / V tiles (4-D boxes of the plane tensors, 64B / 128B swizzle).
//   warp 5     UMMA issuer, owns TMEM:  S = Q K^T  (A, B from shared memory, K-major),
//              O += P V  (A = P from TENSOR MEMORY, B = V from shared memory, MN-major: V is read as stored, [key][dim]).
//              The small cross terms of O accumulate in their own TMEM columns (the tensor-core accumulator truncates).
// TMEM columns: [0, KT) S / P, [KT, KT + D) O main, [KT + D, KT + 2 D) O small terms.
#include <cuda_fp16.h>

#include "ctx.h"
#include "ptx.cuh"

namespace tfmq {

struct AttnTcP {
  float* o;
  __half* o_hi;
  __half* o_lo;
  long long o_sb, o_sh, o_st;
  int heads, tq, tk, d;
  float scale_log2e;       // softmax scale * log2(e)
};

constexpr int ATC_SOFTMAX_WARPS = 4;
constexpr int ATC_WARP_TMA = 4, ATC_WARP_MMA = 5;
constexpr int ATC_THREADS = 6 * 32;
constexpr int ATC_STAGES = 2;
constexpr float ATC_LAZY = 8.f;     // rescale only when the running max (log2 domain) grew by more than this

// A from tensor memory, B from shared memory
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// shared-memory matrix descriptor of an operand tile whose rows are ROWB bytes (64: 64B swizzle, 128: 128B swizzle):
// 8-row groups are 8 * ROWB bytes apart.  Used for both majors: K-major (rows = M / N index, the K extent lies inside a
// row) and MN-major (rows = K index, the N extent lies inside a row).
template <int ROWB>
__device__ __forceinline__ uint64_t smem_desc_rows(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((8 * ROWB) >> 4) << 32;             // stride byte offset: next 8-row group
  d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell)
  d |= (uint64_t)(ROWB == 128 ? 2 : 4) << 61;         // SWIZZLE_128B / SWIZZLE_64B
  return d;
}

// D = head dim as the tensor cores see it (a multiple of 16; a real head dim of 40 runs as 48 with TMA zero fill),
// ROWB = bytes of one operand row in shared memory (64 for D = 32, 128 for D <= 64), KT = keys per tile.
template <int D, int ROWB, int KT>
__global__ void __launch_bounds__(ATC_THREADS, 2)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQh, const __grid_constant__ CUtensorMap tmQl,
               const __grid_constant__ CUtensorMap tmKh, const __grid_constant__ CUtensorMap tmKl,
               const __grid_constant__ CUtensorMap tmVh, const __grid_constant__ CUtensorMap tmVl, const AttnTcP p) {
  static_assert(D % 16 == 0 && D * 2 <= ROWB && (ROWB == 64 || ROWB == 128), "operand geometry");
  static_assert(KT == 64 || KT == 128, "key tile");
  constexpr uint32_t Q_BYTES = 128u * ROWB;           // one plane of the 128-query tile
  constexpr uint32_t KV_BYTES = (uint32_t)KT * ROWB;  // one plane of a key tile
  constexpr uint32_t STAGE_BYTES = 4u * KV_BYTES;     // K hi, K lo, V hi, V lo
  constexpr uint32_t TMEM_COLS = (KT + 2 * D <= 128) ? 128u : (KT + 2 * D <= 256) ? 256u : 512u;
  constexpr int NCH = KT / 32;                        // 32-column chunks of a score tile

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* q_hi = smem;
  uint8_t* q_lo = smem + Q_BYTES;
  uint8_t* ring = smem + 2 * Q_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + ATC_STAGES * STAGE_BYTES);
  uint64_t* q_full = bars;                 // Q planes landed
  uint64_t* kv_full = bars + 1;            // [ATC_STAGES] K / V planes of the stage landed
  uint64_t* kv_empty = kv_full + ATC_STAGES;   // [ATC_STAGES] the UMMAs that read the stage retired
  uint64_t* s_full = kv_empty + ATC_STAGES;    // S tile complete in TMEM (and every earlier UMMA retired)
  uint64_t* p_full = s_full + 1;           // P written over S (and O rescaled) by all four softmax warps
  uint64_t* o_full = p_full + 1;           // last P V retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y, b = bh / p.heads, h = bh - b * p.heads;
  const int q0 = blockIdx.x * 128;
  const int ntiles = (p.tk + KT - 1) / KT;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < ATC_STAGES; ++s) mbar_init(&kv_full[s], 1), mbar_init(&kv_empty[s], 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, ATC_SOFTMAX_WARPS);
    mbar_init(o_full, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmQh), tma_prefetch_desc(&tmQl), tma_prefetch_desc(&tmKh);
    tma_prefetch_desc(&tmKl), tma_prefetch_desc(&tmVh), tma_prefetch_desc(&tmVl);
  }
  if (warp == ATC_WARP_MMA) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + (uint32_t)KT, tmem_o2 = tmem_o + (uint32_t)D;

  if (warp == ATC_WARP_TMA) {
    // ===================================================== TMA producer
    if (elect_one_sync()) {
      mbar_expect_tx(q_full, 2 * Q_BYTES);
      tma_load_4d(q_hi, &tmQh, q_full, 0, h, q0, b);
      tma_load_4d(q_lo, &tmQl, q_full, 0, h, q0, b);
    }
    __syncwarp();
    int s = 0;
    uint32_t par = 0;
    for (int j = 0; j < ntiles; ++j) {
      mbar_wait_relaxed(&kv_empty[s], par ^ 1u);
      if (elect_one_sync()) {
        uint8_t* st = ring + (size_t)s * STAGE_BYTES;
        mbar_expect_tx(&kv_full[s], STAGE_BYTES);
        tma_load_4d(st, &tmKh, &kv_full[s], 0, h, j * KT, b);
        tma_load_4d(st + KV_BYTES, &tmKl, &kv_full[s], 0, h, j * KT, b);
        tma_load_4d(st + 2 * KV_BYTES, &tmVh, &kv_full[s], 0, h, j * KT, b);
        tma_load_4d(st + 3 * KV_BYTES, &tmVl, &kv_full[s], 0, h, j * KT, b);
      }
      __syncwarp();
      if (++s == ATC_STAGES) s = 0, par ^= 1u;
    }
  } else if (warp == ATC_WARP_MMA) {
    // ===================================================== UMMA issuer
    const uint32_t idesc_pv = idesc_f16(128, D) | (1u << 16);     // B (= V) is MN-major
    const uint64_t d_qh = smem_desc_rows<ROWB>(smem_u32(q_hi)), d_ql = smem_desc_rows<ROWB>(smem_u32(q_lo));
    const uint64_t d_ring = smem_desc_rows<ROWB>(smem_u32(ring));
    constexpr uint32_t KV_U = KV_BYTES >> 4, STAGE_U = STAGE_BYTES >> 4;    // descriptor address units
    constexpr uint32_t V_KSTEP_U = (16u * ROWB) >> 4;                       // 16 keys further down an MN-major tile
    mbar_wait(q_full, 0);
    int s = 0;
    uint32_t par = 0;
    for (int j = 0; j < ntiles; ++j) {
      const int kn = min(KT, p.tk - j * KT);           // valid keys of this tile
      const int nks = (kn + 15) >> 4;                  // 16-key steps of P V
      const uint32_t idesc_qk = idesc_f16(128, (uint32_t)(nks * 16));
      mbar_wait(&kv_full[s], par);
      tc_fence_after();
      const uint64_t d_kh = d_ring + (uint32_t)s * STAGE_U, d_kl = d_kh + KV_U;
      const uint64_t d_vh = d_kh + 2 * KV_U, d_vl = d_kh + 3 * KV_U;
      if (elect_one_sync()) {
        // S = Q_lo K_hi + Q_hi K_lo + Q_hi K_hi   (the in-order tensor pipe has finished reading P_{j-1} by then)
#pragma unroll
        for (int k = 0; k < D / 16; ++k) umma_f16(tmem_base, d_ql + 2 * k, d_kh + 2 * k, idesc_qk, k > 0);
#pragma unroll
        for (int k = 0; k < D / 16; ++k) umma_f16(tmem_base, d_qh + 2 * k, d_kl + 2 * k, idesc_qk, 1);
#pragma unroll
        for (int k = 0; k < D / 16; ++k) umma_f16(tmem_base, d_qh + 2 * k, d_kh + 2 * k, idesc_qk, 1);
        umma_commit(s_full);
      }
      __syncwarp();
      mbar_wait(p_full, (uint32_t)j & 1u);
      tc_fence_after();
      if (elect_one_sync()) {
        // P of 32-key chunk c sits at columns [32 c, 32 c + 16) (hi) and [32 c + 16, 32 c + 32) (lo), two halves per column
        for (int ks = 0; ks < nks; ++ks) {
          const uint32_t a_hi = tmem_base + (uint32_t)(32 * (ks >> 1) + 8 * (ks & 1)), a_lo = a_hi + 16u;
          const uint64_t vh = d_vh + (uint32_t)ks * V_KSTEP_U, vl = d_vl + (uint32_t)ks * V_KSTEP_U;
          const uint32_t acc = (j > 0 || ks > 0) ? 1u : 0u;
          umma_f16_ts(tmem_o2, a_lo, vh, idesc_pv, acc);
          umma_f16_ts(tmem_o2, a_hi, vl, idesc_pv, 1);
          umma_f16_ts(tmem_o, a_hi, vh, idesc_pv, acc);
        }
        umma_commit(&kv_empty[s]);
        if (j == ntiles - 1) umma_commit(o_full);
      }
      __syncwarp();
      if (++s == ATC_STAGES) s = 0, par ^= 1u;
    }
  } else {
    // ===================================================== softmax warps: thread = query row = TMEM lane
    const int r = warp * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    const uint32_t t_s = tmem_base + lane_addr;
    const float c = p.scale_log2e;
    float m = -INFINITY, l = 0.f;          // running max (raw score units) and row sum
    for (int j = 0; j < ntiles; ++j) {
      const int kn = min(KT, p.tk - j * KT);
      mbar_wait(s_full, (uint32_t)j & 1u);
      tc_fence_after();
      // ---- pass 1: row maximum of the valid columns
      float mx = -INFINITY;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        if (ch * 32 < kn) {
          uint32_t v[32];
          tmem_ld32(t_s + (uint32_t)(ch * 32), v);
          tmem_ld_wait();
          if (ch * 32 + 32 <= kn) {
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (ch * 32 + i < kn) mx = fmaxf(mx, __uint_as_float(v[i]));
          }
        }
      }
      // ---- lazy rescale of the accumulators (everything issued before S_j, i.e. P V of tile j - 1, has retired)
      float alpha = 1.f;
      bool grow = false;
      if (j == 0) {
        m = mx;
      } else if ((mx - m) * c > ATC_LAZY) {
        alpha = ex2_approx((m - mx) * c);
        m = mx;
        grow = true;
      }
      if (__any_sync(0xffffffffu, grow)) {
        l *= alpha;
#pragma unroll
        for (int part = 0; part < 2 * D / 16; ++part) {     // O main and O small terms: 2 D consecutive columns
          uint32_t v[16];
          tmem_ld16(tmem_o + lane_addr + (uint32_t)(part * 16), v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
          tmem_st16(tmem_o + lane_addr + (uint32_t)(part * 16), v);
        }
      }
      // ---- pass 2: p = 2^(s c - m c), row sum, fp16 hi / lo split, written over S in place
      const float mc = m * c;
      float rs = 0.f;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        if (ch * 32 < kn) {
          uint32_t v[32], w[32];
          tmem_ld32(t_s + (uint32_t)(ch * 32), v);
          tmem_ld_wait();
          const bool tail = ch * 32 + 32 > kn;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float p0 = ex2_approx(fmaf(__uint_as_float(v[2 * i]), c, -mc));
            float p1 = ex2_approx(fmaf(__uint_as_float(v[2 * i + 1]), c, -mc));
            if (tail) {
              if (ch * 32 + 2 * i >= kn) p0 = 0.f;
              if (ch * 32 + 2 * i + 1 >= kn) p1 = 0.f;
            }
            rs += p0 + p1;
            // column i of the packed A operand holds keys 2i (low half) and 2i + 1 (high half)
            const __half2 hh = __floats2half2_rn(p0, p1);
            const float2 hf = __half22float2(hh);
            const __half2 ll = __floats2half2_rn(p0 - hf.x, p1 - hf.y);
            w[i] = *reinterpret_cast<const uint32_t*>(&hh);
            w[16 + i] = *reinterpret_cast<const uint32_t*>(&ll);
          }
          tmem_st32(t_s + (uint32_t)(ch * 32), w);
        }
      }
      l += rs;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // ---- epilogue: O = (O_main + O_small) / l
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv = 1.f / l;
    const int t = q0 + r;
    const long long off = (long long)b * p.o_sb + (long long)h * p.o_sh + (long long)t * p.o_st;
#pragma unroll
    for (int part = 0; part < D / 16; ++part) {
      uint32_t v[16], v2[16];
      tmem_ld16(tmem_o + lane_addr + (uint32_t)(part * 16), v);
      tmem_ld16(tmem_o2 + lane_addr + (uint32_t)(part * 16), v2);
      tmem_ld_wait();
      float f[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = (__uint_as_float(v[i]) + __uint_as_float(v2[i])) * inv;
      if (t < p.tq && part * 16 < p.d) {
        if (p.o_hi) {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            // the saturating split tfmq_act_prepare writes (elementwise.cu::split_h16x4)
            uint16_t h0, h1, l0, l1;
            asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h0) : "f"(f[2 * i]));
            asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h1) : "f"(f[2 * i + 1]));
            const float r0 = f[2 * i] - __half2float(__ushort_as_half(h0));
            const float r1 = f[2 * i + 1] - __half2float(__ushort_as_half(h1));
            asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(l0) : "f"(r0));
            asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(l1) : "f"(r1));
            hi[i] = (uint32_t)h0 | ((uint32_t)h1 << 16);
            lo[i] = (uint32_t)l0 | ((uint32_t)l1 << 16);
          }
          if (part * 16 + 16 <= p.d) {
            uint4* dh = reinterpret_cast<uint4*>(p.o_hi + off + part * 16);
            uint4* dl = reinterpret_cast<uint4*>(p.o_lo + off + part * 16);
            dh[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]), dh[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
            dl[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]), dl[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
          } else {                                          // head dim 40: the last 16-column part holds 8 real columns
            *reinterpret_cast<uint4*>(p.o_hi + off + part * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(p.o_lo + off + part * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
        } else {
          float4* dst = reinterpret_cast<float4*>(p.o + off + part * 16);
          const int nv = (part * 16 + 16 <= p.d) ? 4 : 2;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (i < nv) dst[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == ATC_WARP_MMA) tmem_dealloc(tmem_base, TMEM_COLS);
}

// 4-D map of one fp16 plane addressed as base + b*sb + h*sh + t*st + dim (halves): dims {d, heads, tokens, b}
static int encode_plane(tfmq_ctx* ctx, CUtensorMap* m, const void* base, int d, int heads, int tokens, int b, int64_t sb,
                        int64_t sh, int64_t st, int box_d, int box_t, int rowb) {
  cuuint64_t dims[4] = {(cuuint64_t)d, (cuuint64_t)heads, (cuuint64_t)tokens, (cuuint64_t)b};
  cuuint64_t str[3] = {(cuuint64_t)sh * 2, (cuuint64_t)st * 2, (cuuint64_t)sb * 2};
  cuuint32_t box[4] = {(cuuint32_t)box_d, 1, (cuuint32_t)box_t, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  // a degenerate dimension (one head / one image) may carry any stride: give it a legal one
  if (heads == 1) str[0] = (cuuint64_t)d * 2 >= 16 ? (((cuuint64_t)d * 2 + 15) & ~15ull) : 16;
  if (b == 1) str[2] = str[1] * (cuuint64_t)tokens;
  CUresult r = ctx->encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, str, box, es,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 rowb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return tfmq_fail(ctx, TFMQ_ERR_CUDA, "attention_h16: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return TFMQ_OK;
}

template <int D, int ROWB, int KT>
static int launch_attn_tc(tfmq_ctx* ctx, const tfmq_attn_h16_desc* d, cudaStream_t st) {
  CUtensorMap tm[6];
  const void* base[6] = {d->q_hi, d->q_lo, d->k_hi, d->k_lo, d->v_hi, d->v_lo};
  for (int i = 0; i < 6; ++i) {
    const bool isq = i < 2, isk = i >= 2 && i < 4;
    const int64_t sb = isq ? d->q_sb : isk ? d->k_sb : d->v_sb, sh = isq ? d->q_sh : isk ? d->k_sh : d->v_sh;
    const int64_t stt = isq ? d->q_st : isk ? d->k_st : d->v_st;
    int rc = encode_plane(ctx, &tm[i], base[i], d->d, d->heads, isq ? d->tq : d->tk, d->b, sb, sh, stt, ROWB / 2,
                          isq ? 128 : KT, ROWB);
    if (rc) return rc;
  }
  AttnTcP p{};
  p.o = d->o, p.o_hi = static_cast<__half*>(d->o_hi), p.o_lo = static_cast<__half*>(d->o_lo);
  p.o_sb = d->o_sb, p.o_sh = d->o_sh, p.o_st = d->o_st;
  p.heads = d->heads, p.tq = d->tq, p.tk = d->tk, p.d = d->d;
  p.scale_log2e = d->scale * 1.4426950408889634f;
  const size_t smem = 1024 + 2 * 128 * ROWB + (size_t)ATC_STAGES * 4 * KT * ROWB + 128;
  auto kern = attn_tc_kernel<D, ROWB, KT>;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return tfmq_fail(ctx, TFMQ_ERR_CUDA, "attention_h16: smem attr: %s", cudaGetErrorString(e));
    smem_set = smem;
  }
  dim3 grid((d->tq + 127) / 128, d->b * d->heads);
  kern<<<grid, ATC_THREADS, smem, st>>>(tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], p);
  TFMQ_LAUNCH_CHECK("attention_h16");
  return TFMQ_OK;
}

}  // namespace tfmq

using namespace tfmq;

extern "C" int tfmq_attention_h16(tfmq_ctx* ctx, const tfmq_attn_h16_desc* d, void* stream) {
  if (!ctx) return TFMQ_ERR_ARG;
  TFMQ_REQUIRE(d && d->q_hi && d->q_lo && d->k_hi && d->k_lo && d->v_hi && d->v_lo && (d->o || (d->o_hi && d->o_lo)),
               TFMQ_ERR_ARG, "attention_h16: null pointer");
  if (d->b == 0 || d->heads == 0 || d->tq == 0) return TFMQ_OK;
  TFMQ_REQUIRE(d->tk > 0, TFMQ_ERR_SHAPE, "attention_h16: empty key set");
  TFMQ_REQUIRE((long long)d->b * d->heads <= 65535, TFMQ_ERR_SHAPE, "attention_h16: b*heads > 65535");
  TFMQ_REQUIRE(d->d % 8 == 0 && d->d >= 16 && d->d <= 64, TFMQ_ERR_SHAPE,
               "attention_h16: head dim %d (supported: multiples of 8 from 16 to 64)", d->d);
  const void* ptrs[6] = {d->q_hi, d->q_lo, d->k_hi, d->k_lo, d->v_hi, d->v_lo};
  for (int i = 0; i < 6; ++i)
    TFMQ_REQUIRE(((uintptr_t)ptrs[i] & 15) == 0, TFMQ_ERR_ARG, "attention_h16: operand planes must be 16-byte aligned");
  const int64_t strides[9] = {d->q_sb, d->q_sh, d->q_st, d->k_sb, d->k_sh, d->k_st, d->v_sb, d->v_sh, d->v_st};
  for (int i = 0; i < 9; ++i)
    TFMQ_REQUIRE(strides[i] % 8 == 0, TFMQ_ERR_SHAPE, "attention_h16: operand strides must be multiples of 8 halves");
  if (d->o_hi)
    TFMQ_REQUIRE((((uintptr_t)d->o_hi | (uintptr_t)d->o_lo) & 15) == 0 && d->o_sb % 8 == 0 && d->o_sh % 8 == 0 &&
                     d->o_st % 8 == 0,
                 TFMQ_ERR_ARG, "attention_h16: output planes must be 16-byte aligned with strides in multiples of 8");
  else
    TFMQ_REQUIRE(((uintptr_t)d->o & 15) == 0 && d->o_sb % 4 == 0 && d->o_sh % 4 == 0 && d->o_st % 4 == 0, TFMQ_ERR_ARG,
                 "attention_h16: fp32 output must be 16-byte aligned with strides in multiples of 4");
  cudaStream_t st = tfmq_stream(stream);
  if (d->d <= 32) {
    if (d->d == 32) return launch_attn_tc<32, 64, 128>(ctx, d, st);
    return launch_attn_tc<32, 128, 64>(ctx, d, st);      // 16, 24: zero-filled up to the 128-byte operand row
  }
  if (d->d <= 48) return launch_attn_tc<48, 128, 64>(ctx, d, st);
  return launch_attn_tc<64, 128, 64>(ctx, d, st);
}
