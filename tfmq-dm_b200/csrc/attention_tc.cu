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
#include <algorithm>
#include <cstdio>
#include <utility>
#include <vector>

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
  int dbg;                 // TFMQ_ATTN_DBG.  timing experiments: 2 = no P V UMMAs, 4 = no Q K^T UMMAs, 16 = predicated softmax path, 64 = no S
                           // prefetch; race hunting (a 20 us sleep): 32 / 128 = in the rescale path, 256 / 512 = one slow warp per group
  long long* prof;         // per-CTA phase cycle counters [grid][16] (TFMQ_ATTN_PROF=1), else null
};

// (counters are accumulated in global memory by one thread: no registers held when profiling is off)
#define ATC_PROF(idx)                                       \
  if (prof_on) {                                            \
    const long long now_ = clock64();                       \
    prof_dst[idx] += now_ - prof_t;                         \
    prof_t = now_;                                          \
  }

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
// TMEM allocation with the column count as an IMMEDIATE: with a register operand the toolchain cannot tell how many
// columns a CTA takes and the launch is limited to one CTA per SM (measured: occupancy 1 at any shared-memory size)
template <uint32_t COLS>
__device__ __forceinline__ void tmem_alloc_imm(uint32_t* slot_in_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t COLS>
__device__ __forceinline__ void tmem_dealloc_imm(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float fmax3(float a, float b, float c) {      // FMNMX3
#ifdef ATC_NO_FMAX3
  return fmaxf(a, fmaxf(b, c));
#else
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
#endif
}
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
// ROWB = bytes of one operand row in shared memory (64 for D = 32, 128 for D <= 64), KT = keys per tile,
// NQ = 128-query blocks per CTA, ST = K / V ring depth.  One CTA per SM (a kernel that allocates tensor memory is given
// one CTA per SM by the launch machinery: measured occupancy 1 at any shared-memory size), so the concurrency an SM needs
// lives INSIDE the CTA: every query block has its own S / P and O columns in tensor memory, its own group of four softmax
// warps and its own UMMA warp, and all blocks share each K / V tile in shared memory.
//   warps [0, 4 NQ)       softmax, group g = query block g
//   warps [4 NQ, 5 NQ)    UMMA issuers, one per query block (the issue cost of a UMMA, ~50 clocks, is what bounds a
//                         single issuing warp: the P V products are only 16-32 clocks of tensor work each)
//   warp  5 NQ            TMA producer
// S is DOUBLE-BUFFERED per query block: Q K^T of tile j + 1 is issued before the warp waits for P of tile j, so the softmax
// warps (bound by the SFU: one ex2 per score) go from tile to tile without waiting for the tensor pipe.
// UMMA count per key tile and query block: 3 D / 16 for S, KT / 16 x 2 for O: the hi and lo planes of V lie next to each
// other in a stage, so ONE MN-major descriptor (leading byte offset = plane size) presents [V_hi | V_lo] as an N = 2 DN operand
// and P_hi [V_hi | V_lo] lands in the adjacent "main" and "small terms" accumulator columns; P_lo V_hi is the second product.
template <int D, int ROWB, int KT, int NQ, int ST>
__global__ void __launch_bounds__((5 * NQ + 1) * 32, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQh, const __grid_constant__ CUtensorMap tmQl,
               const __grid_constant__ CUtensorMap tmKh, const __grid_constant__ CUtensorMap tmKl,
               const __grid_constant__ CUtensorMap tmVh, const __grid_constant__ CUtensorMap tmVl, const AttnTcP p) {
  static_assert(D % 16 == 0 && D * 2 <= ROWB && (ROWB == 64 || ROWB == 128), "operand geometry");
  static_assert(KT == 64 || KT == 128, "key tile");
  constexpr int DN = ROWB / 2;                        // columns of one V plane as the stacked operand sees it (>= D)
  constexpr uint32_t Q_BYTES = 128u * ROWB;           // one plane of a 128-query block
  constexpr uint32_t KV_BYTES = (uint32_t)KT * ROWB;  // one plane of a key tile
  constexpr uint32_t STAGE_BYTES = 4u * KV_BYTES;     // K hi, K lo, V hi, V lo
  constexpr uint32_t QB_COLS = 2 * KT + 2 * DN;       // TMEM columns of one query block: S / P x 2, O main, O small terms
  constexpr uint32_t TMEM_COLS = (NQ * QB_COLS <= 128) ? 128u : (NQ * QB_COLS <= 256) ? 256u : 512u;
  static_assert(NQ * QB_COLS <= 512, "tensor memory");
  constexpr int WARP_MMA0 = 4 * NQ, WARP_TMA = 5 * NQ;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* q_smem = smem;                              // [NQ][hi, lo]
  uint8_t* ring = smem + NQ * 2 * Q_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + ST * STAGE_BYTES);
  uint64_t* q_full = bars;                 // Q planes of all query blocks landed
  uint64_t* kv_full = bars + 1;            // [ST] K / V planes of the stage landed
  uint64_t* kv_empty = kv_full + ST;       // [ST] the UMMAs of every query block that read the stage retired
  uint64_t* s_full = kv_empty + ST;        // [NQ][2] S tile complete in TMEM buffer 0 / 1
  uint64_t* p_full = s_full + 2 * NQ;      // [NQ][2] P written over S (and O rescaled) by the block's four softmax warps
  uint64_t* pv_done = p_full + 2 * NQ;     // [NQ] the block's P V of a tile retired (one phase per tile; waited for only by the
                                           // in-loop rescale, which is never more than one phase behind)
  uint64_t* o_full = pv_done + NQ;         // [NQ] the block's LAST P V retired (the softmax warps may reach the epilogue two
                                           // phases ahead of pv_done: a parity wait on it would alias)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + NQ);

  const long long t_entry = p.prof ? clock64() : 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y, b = bh / p.heads, h = bh - b * p.heads;
  const int q0 = blockIdx.x * 128 * NQ;
  const int nqv = min(NQ, (p.tq - q0 + 127) / 128);   // query blocks of this CTA that hold rows
  const int ntiles = (p.tk + KT - 1) / KT;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < ST; ++s) mbar_init(&kv_full[s], 1), mbar_init(&kv_empty[s], (uint32_t)nqv);
    for (int i = 0; i < 2 * NQ; ++i) mbar_init(&s_full[i], 1), mbar_init(&p_full[i], 4);
    for (int i = 0; i < NQ; ++i) mbar_init(&pv_done[i], 1), mbar_init(&o_full[i], 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmQh), tma_prefetch_desc(&tmQl), tma_prefetch_desc(&tmKh);
    tma_prefetch_desc(&tmKl), tma_prefetch_desc(&tmVh), tma_prefetch_desc(&tmVl);
  }
  if (warp == WARP_TMA) tmem_alloc_imm<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == WARP_TMA) {
    // ===================================================== TMA producer
    if (elect_one_sync()) {
      mbar_expect_tx(q_full, (uint32_t)nqv * 2 * Q_BYTES);
      for (int i = 0; i < nqv; ++i) {
        tma_load_4d(q_smem + (size_t)(2 * i) * Q_BYTES, &tmQh, q_full, 0, h, q0 + 128 * i, b);
        tma_load_4d(q_smem + (size_t)(2 * i + 1) * Q_BYTES, &tmQl, q_full, 0, h, q0 + 128 * i, b);
      }
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
      if (++s == ST) s = 0, par ^= 1u;
    }
  } else if (warp >= WARP_MMA0) {
    // ===================================================== UMMA issuer of query block i
    const int i = warp - WARP_MMA0;
    if (i < nqv) {
      // every operand below is warp-uniform and computed outside the elected region: the descriptors stay in uniform
      // registers and the UTCHMMAs issue back to back
      const uint32_t idesc_pv2 = idesc_f16(128, 2 * DN) | (1u << 16);   // B = [V_hi | V_lo], MN-major
      const uint32_t idesc_pv1 = idesc_f16(128, D) | (1u << 16);        // B = V_hi
      constexpr uint32_t Q_U = Q_BYTES >> 4, KV_U = KV_BYTES >> 4, STAGE_U = STAGE_BYTES >> 4;   // descriptor address units
      constexpr uint32_t V_KSTEP_U = (16u * ROWB) >> 4;                 // 16 keys further down an MN-major tile
      const uint64_t d_qh = smem_desc_rows<ROWB>(smem_u32(q_smem)) + (uint32_t)(2 * i) * Q_U, d_ql = d_qh + Q_U;
      const uint64_t d_ring = smem_desc_rows<ROWB>(smem_u32(ring));
      const uint64_t lbo_planes = (uint64_t)KV_U << 16;                 // leading byte offset: V_lo plane = next N atom
      const uint32_t t_q = tmem_base + (uint32_t)i * QB_COLS, t_o = t_q + 2u * KT, t_o2 = t_o + (uint32_t)DN;
      // S_buf = Q_lo K_hi + Q_hi K_lo + Q_hi K_hi of the key tile in stage s
      auto issue_qk = [&](int s, int kn, int buf) {
        const uint32_t idesc_qk = idesc_f16(128, (uint32_t)(((kn + 15) >> 4) * 16));
        const uint64_t d_kh = d_ring + (uint32_t)s * STAGE_U, d_kl = d_kh + KV_U;
        const uint32_t t_s = t_q + (uint32_t)(buf * KT);
        if (elect_one_sync()) {
          if (!(p.dbg & 4)) {
#pragma unroll
            for (int k = 0; k < D / 16; ++k) umma_f16(t_s, d_ql + 2 * k, d_kh + 2 * k, idesc_qk, k > 0);
#pragma unroll
            for (int k = 0; k < D / 16; ++k) umma_f16(t_s, d_qh + 2 * k, d_kl + 2 * k, idesc_qk, 1);
#pragma unroll
            for (int k = 0; k < D / 16; ++k) umma_f16(t_s, d_qh + 2 * k, d_kh + 2 * k, idesc_qk, 1);
          }
          umma_commit(&s_full[2 * i + buf]);
        }
        __syncwarp();
      };
      const bool prof_on = p.prof != nullptr && i == 0 && lane == 0;
      long long* prof_dst = p.prof + ((long long)blockIdx.y * gridDim.x + blockIdx.x) * 16 + 7;
      long long prof_t = prof_on ? clock64() : 0;
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      issue_qk(0, min(KT, p.tk), 0);
      ATC_PROF(0)
      int s = 0;
      uint32_t par = 0;
      for (int j = 0; j < ntiles; ++j) {
        const int kn = min(KT, p.tk - j * KT);           // valid keys of this tile
        const int nks = (kn + 15) >> 4;                  // 16-key steps of P V
        const int buf = j & 1;
        const int s1 = (s + 1 == ST) ? 0 : s + 1;
        const uint32_t par1 = (s + 1 == ST) ? par ^ 1u : par;
        if (j + 1 < ntiles && !(p.dbg & 64)) {
          // S of the NEXT tile into the other buffer (behind this block's P V of tile j - 1, which read P from it) before
          // waiting for the softmax warps: they find it ready when they finish tile j
          mbar_wait(&kv_full[s1], par1);
          tc_fence_after();
          ATC_PROF(0)
          issue_qk(s1, min(KT, p.tk - (j + 1) * KT), buf ^ 1);
          ATC_PROF(1)
        }
        const uint64_t d_vh = d_ring + (uint32_t)s * STAGE_U + 2 * KV_U, d_vhl = d_vh | lbo_planes;
        const uint32_t t_p = t_q + (uint32_t)(buf * KT);
        mbar_wait(&p_full[2 * i + buf], (uint32_t)(j >> 1) & 1u);
        tc_fence_after();
        ATC_PROF(2)
        const uint32_t acc0 = j > 0 ? 1u : 0u;
        if (elect_one_sync()) {
          // P of 16-key group ks sits at columns [16 ks, 16 ks + 8) (hi) and [16 ks + 8, 16 ks + 16) (lo), two halves per column
          if (!(p.dbg & 2)) {
#pragma unroll
            for (int ks = 0; ks < KT / 16; ++ks) {
              if (ks < nks) {
                // [O main | O small] (+)= P_hi [V_hi | V_lo];   O small += P_lo V_hi
                umma_f16_ts(t_o, t_p + (uint32_t)(16 * ks), d_vhl + (uint32_t)ks * V_KSTEP_U, idesc_pv2, ks > 0 ? 1u : acc0);
                umma_f16_ts(t_o2, t_p + (uint32_t)(16 * ks + 8), d_vh + (uint32_t)ks * V_KSTEP_U, idesc_pv1, 1);
              }
            }
          }
          umma_commit(&kv_empty[s]);
          umma_commit(&pv_done[i]);
          if (j == ntiles - 1) umma_commit(&o_full[i]);
        }
        __syncwarp();
        if (j + 1 < ntiles && (p.dbg & 64)) {
          mbar_wait(&kv_full[s1], par1);
          tc_fence_after();
          issue_qk(s1, min(KT, p.tk - (j + 1) * KT), buf ^ 1);
        }
        ATC_PROF(3)
        s = s1, par = par1;
      }
    }
  } else {
    // ===================================================== softmax warps: group = query block, thread = query row = TMEM lane
    const int qb = warp >> 2, wq = warp & 3;
    if (qb < nqv) {
      const int r = wq * 32 + lane;
      const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
      const uint32_t t_q = tmem_base + (uint32_t)qb * QB_COLS + lane_addr;
      const uint32_t t_o = t_q + 2u * KT, t_o2 = t_o + (uint32_t)DN;
      const float c = p.scale_log2e;
      float m = -INFINITY, l = 0.f;          // running max (raw score units) and row sum
      const bool prof_on = p.prof != nullptr && threadIdx.x == 0;
      long long* prof_dst = p.prof + ((long long)blockIdx.y * gridDim.x + blockIdx.x) * 16;
      long long prof_t = prof_on ? clock64() : 0;
      if (prof_on) prof_dst[6] = -prof_t, prof_dst[12] = prof_t - t_entry;     // [12]: CTA entry -> roles start (prologue)
      for (int j = 0; j < ntiles; ++j) {
        const int kn = min(KT, p.tk - j * KT);
        const int buf = j & 1;
        const uint32_t t_s = t_q + (uint32_t)(buf * KT);
        if ((p.dbg & 256) && wq == 3 && (j & 3) == 1) __nanosleep(20000);      // one slow warp per group (race hunting)
        mbar_wait(&s_full[2 * qb + buf], (uint32_t)(j >> 1) & 1u);
        tc_fence_after();
        if ((p.dbg & 512) && wq == 2 && (j & 3) == 2) __nanosleep(20000);
        if (prof_on && j == 0) prof_dst[13] = clock64() - prof_t;               // [13]: roles start -> first S tile
        ATC_PROF(0)
        // ---- the whole score row into registers (one TMEM round trip), row maximum of the valid columns
        uint32_t sv[KT];
#pragma unroll
        for (int ch = 0; ch < KT / 32; ++ch)
          if (ch * 32 < kn) tmem_ld32(t_s + (uint32_t)(ch * 32), *reinterpret_cast<uint32_t(*)[32]>(&sv[ch * 32]));
        tmem_ld_wait();
        float mx = -INFINITY;
        const bool full = kn == KT && !(p.dbg & 16);
        if (full) {
          float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int k = 0; k < KT; k += 8) {
            m4[0] = fmax3(m4[0], __uint_as_float(sv[k]), __uint_as_float(sv[k + 1]));
            m4[1] = fmax3(m4[1], __uint_as_float(sv[k + 2]), __uint_as_float(sv[k + 3]));
            m4[2] = fmax3(m4[2], __uint_as_float(sv[k + 4]), __uint_as_float(sv[k + 5]));
            m4[3] = fmax3(m4[3], __uint_as_float(sv[k + 6]), __uint_as_float(sv[k + 7]));
          }
          mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        } else {
#pragma unroll
          for (int k = 0; k < KT; ++k)
            if (k < kn) mx = fmaxf(mx, __uint_as_float(sv[k]));
        }
        // ---- lazy rescale of the accumulators, once the P V of tile j - 1 has retired (S of this tile was issued ahead of it)
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
          if (p.dbg & 128) __nanosleep(20000);
          mbar_wait(&pv_done[qb], (uint32_t)(j - 1) & 1u);
          if (p.dbg & 32) __nanosleep(20000);
          tc_fence_after();
          l *= alpha;
#pragma unroll
          for (int part = 0; part < 2 * DN / 16; ++part) {   // O main and O small terms: 2 DN consecutive columns
            uint32_t v[16];
            tmem_ld16(t_o + (uint32_t)(part * 16), v);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = __float_as_uint(__uint_as_float(v[k]) * alpha);
            tmem_st16(t_o + (uint32_t)(part * 16), v);
          }
        }
        ATC_PROF(1)
        // ---- p = 2^(s c - m c), row sum, fp16 hi / lo split; per 16-key group (= one K step of P V) written over S in place
        // as [8 columns hi | 8 columns lo].  hi = p with the low 13 significand bits cleared (exact in fp16 above its
        // subnormal range), lo = p - hi: one logic op and one add per element, and the SFU stays the busiest pipe.
        const float mc = m * c;
        float rs0 = 0.f, rs1 = 0.f;
        auto split_store = [&](int g, float (&pv)[16]) {
          uint32_t w[16];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float p0 = pv[2 * k], p1 = pv[2 * k + 1];
            rs0 += p0, rs1 += p1;
            const float h0 = __uint_as_float(__float_as_uint(p0) & 0xFFFFE000u);
            const float h1 = __uint_as_float(__float_as_uint(p1) & 0xFFFFE000u);
            // a packed column of the A operand holds keys 2k (low half) and 2k + 1 (high half)
            const __half2 hh = __floats2half2_rn(h0, h1);
            const __half2 ll = __floats2half2_rn(p0 - h0, p1 - h1);
            w[k] = *reinterpret_cast<const uint32_t*>(&hh);
            w[8 + k] = *reinterpret_cast<const uint32_t*>(&ll);
          }
          tmem_st16(t_s + (uint32_t)(g * 16), w);
        };
        if (full) {                      // the common case: no per-element predicates
#pragma unroll
          for (int g = 0; g < KT / 16; ++g) {
            float pv[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) pv[k] = ex2_approx(fmaf(__uint_as_float(sv[g * 16 + k]), c, -mc));
            split_store(g, pv);
          }
        } else {
#pragma unroll
          for (int g = 0; g < KT / 16; ++g) {
            if (g * 16 < kn) {
              float pv[16];
#pragma unroll
              for (int k = 0; k < 16; ++k)
                pv[k] = (g * 16 + k < kn) ? ex2_approx(fmaf(__uint_as_float(sv[g * 16 + k]), c, -mc)) : 0.f;
              split_store(g, pv);
            }
          }
        }
        l += rs0 + rs1;
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[2 * qb + buf]);
        ATC_PROF(2)
      }
      // ---- epilogue: O = (O_main + O_small) / l
      mbar_wait(&o_full[qb], 0);
      tc_fence_after();
      ATC_PROF(3)
      const float inv = 1.f / l;
      const int t = q0 + qb * 128 + r;
      const long long off = (long long)b * p.o_sb + (long long)h * p.o_sh + (long long)t * p.o_st;
#pragma unroll
      for (int part = 0; part < D / 16; ++part) {
        uint32_t v[16], v2[16];
        tmem_ld16(t_o + (uint32_t)(part * 16), v);
        tmem_ld16(t_o2 + (uint32_t)(part * 16), v2);
        tmem_ld_wait();
        float f[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) f[k] = (__uint_as_float(v[k]) + __uint_as_float(v2[k])) * inv;
        if (t < p.tq && part * 16 < p.d) {
          if (p.o_hi) {
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              // the saturating split tfmq_act_prepare writes (elementwise.cu::split_h16x4)
              uint16_t h0, h1, l0, l1;
              asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h0) : "f"(f[2 * k]));
              asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h1) : "f"(f[2 * k + 1]));
              const float r0 = f[2 * k] - __half2float(__ushort_as_half(h0));
              const float r1 = f[2 * k + 1] - __half2float(__ushort_as_half(h1));
              asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(l0) : "f"(r0));
              asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(l1) : "f"(r1));
              hi[k] = (uint32_t)h0 | ((uint32_t)h1 << 16);
              lo[k] = (uint32_t)l0 | ((uint32_t)l1 << 16);
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
            for (int k = 0; k < 4; ++k)
              if (k < nv) dst[k] = make_float4(f[4 * k], f[4 * k + 1], f[4 * k + 2], f[4 * k + 3]);
          }
        }
      }
      ATC_PROF(4)
      if (prof_on) prof_dst[6] += clock64();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == WARP_TMA) tmem_dealloc_imm<TMEM_COLS>(tmem_base);
  if (p.prof && threadIdx.x == 0) {
    long long* prof_dst = p.prof + ((long long)blockIdx.y * gridDim.x + blockIdx.x) * 16;
    prof_dst[14] = clock64() - t_entry;                                          // [14]: CTA entry -> exit
    prof_dst[11] = t_entry;
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    prof_dst[15] = smid;
  }
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

template <int D, int ROWB, int KT, int NQ, int ST>
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
  static const int dbg_env = getenv("TFMQ_ATTN_DBG") ? atoi(getenv("TFMQ_ATTN_DBG")) : 0;
  static const bool prof_env = getenv("TFMQ_ATTN_PROF") != nullptr;   // debug aid: in-kernel phase counters, synchronous
  p.dbg = dbg_env;
  p.prof = nullptr;
  const size_t smem = 1024 + (size_t)NQ * 2 * 128 * ROWB + (size_t)ST * 4 * KT * ROWB + 256;
  auto kern = attn_tc_kernel<D, ROWB, KT, NQ, ST>;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return tfmq_fail(ctx, TFMQ_ERR_CUDA, "attention_h16: smem attr: %s", cudaGetErrorString(e));
    // several CTAs per SM are the design point: ask for the full shared-memory carve-out
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    smem_set = smem;
  }
  if (prof_env) {
    int occ = -1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, (5 * NQ + 1) * 32, smem);
    fprintf(stderr, "[attn prof] occupancy %d CTAs / SM (%zu B dynamic smem, %d threads)\n", occ, smem, (5 * NQ + 1) * 32);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, kern);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, ctx->device);
    fprintf(stderr, "[attn prof] regs %d static smem %zu local %zu maxthreads %d maxdyn %d carveout %d | SM: smem %zu regs %d "
            "blocks %d reserved %zu\n", fa.numRegs, fa.sharedSizeBytes, fa.localSizeBytes, fa.maxThreadsPerBlock,
            fa.maxDynamicSharedSizeBytes, fa.preferredShmemCarveout, prop.sharedMemPerMultiprocessor, prop.regsPerMultiprocessor,
            prop.maxBlocksPerMultiProcessor, prop.reservedSharedMemPerBlock);
    for (size_t sm = 8192; sm <= smem; sm += 8192) {
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, (5 * NQ + 1) * 32, sm);
      fprintf(stderr, " %zuK:%d", sm / 1024, occ);
    }
    fprintf(stderr, "\n");
  }
  dim3 grid((d->tq + 128 * NQ - 1) / (128 * NQ), d->b * d->heads);
  const size_t nctas = (size_t)grid.x * grid.y;
  if (prof_env) {
    cudaMalloc(&p.prof, nctas * 16 * sizeof(long long));
    cudaMemsetAsync(p.prof, 0, nctas * 16 * sizeof(long long), st);
  }
  static const bool time_env = getenv("TFMQ_ATTN_TIME") != nullptr;   // debug aid: per-launch device time, synchronous
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (time_env) {
    cudaEventCreate(&ev0), cudaEventCreate(&ev1);
    cudaEventRecord(ev0, st);
  }
  kern<<<grid, (5 * NQ + 1) * 32, smem, st>>>(tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], p);
  TFMQ_LAUNCH_CHECK("attention_h16");
  if (time_env) {
    cudaEventRecord(ev1, st);
    cudaStreamSynchronize(st);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev0, ev1);
    cudaEventDestroy(ev0), cudaEventDestroy(ev1);
    fprintf(stderr, "[attn time] b %d heads %d tq %d tk %d d %d NQ %d: %.1f us\n", d->b, d->heads, d->tq, d->tk, d->d, NQ, ms * 1e3f);
  }
  if (prof_env) {
    std::vector<long long> hb(nctas * 16);
    cudaMemcpy(hb.data(), p.prof, hb.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(p.prof);
    double avg[16] = {0};
    for (size_t c = 0; c < nctas; ++c)
      for (int i = 0; i < 16; ++i) avg[i] += (double)hb[c * 16 + i] / (double)nctas;
    const double nt = (double)((d->tk + KT - 1) / KT);
    {
      double per_sm[256] = {0};
      for (size_t c = 0; c < nctas; ++c) per_sm[hb[c * 16 + 15] & 255] += (double)hb[c * 16 + 14];
      double mx = 0, sum = 0;
      for (int i = 0; i < 256; ++i) mx = per_sm[i] > mx ? per_sm[i] : mx, sum += per_sm[i];
      fprintf(stderr, "[attn prof] CTA lifetime avg %.0f clocks (prologue %.0f, first S after %.0f); busiest SM holds %.0f clocks "
              "of CTA lifetime, mean SM %.0f\n", avg[14], avg[12], avg[13], mx, sum / ctx->sm_count);
      // gaps between consecutive CTAs of one SM (clock64 is per SM)
      std::vector<std::vector<std::pair<long long, long long>>> tl(256);
      for (size_t c = 0; c < nctas; ++c) tl[hb[c * 16 + 15] & 255].push_back({hb[c * 16 + 11], hb[c * 16 + 14]});
      double gap_sum = 0, span_sum = 0;
      long long ngap = 0, nsm = 0;
      for (auto& v : tl) {
        if (v.empty()) continue;
        std::sort(v.begin(), v.end());
        for (size_t k = 1; k < v.size(); ++k) gap_sum += (double)(v[k].first - (v[k - 1].first + v[k - 1].second)), ++ngap;
        span_sum += (double)(v.back().first + v.back().second - v.front().first);
        ++nsm;
      }
      fprintf(stderr, "[attn prof] per SM: first entry -> last exit %.0f clocks on average; gap between a CTA's exit and the next "
              "CTA's entry %.0f clocks on average (%lld gaps)\n", span_sum / nsm, ngap ? gap_sum / ngap : 0.0, ngap);
    }
    fprintf(stderr,
            "[attn prof] b %d heads %d tq %d tk %d d %d KT %d NQ %d | softmax warp 0 per key tile (each phase includes ~250 clocks of "
            "counter overhead): wait_s %.0f max+rescale %.0f exp+store %.0f | wait_o %.0f epilogue %.0f | loop total %.0f | "
            "UMMA warp 0 per key tile: wait_kv %.0f issue_qk %.0f wait_p %.0f issue_pv %.0f\n",
            d->b, d->heads, d->tq, d->tk, d->d, KT, NQ, avg[0] / nt, avg[1] / nt, avg[2] / nt, avg[3], avg[4], avg[6],
            avg[7] / nt, avg[8] / nt, avg[9] / nt, avg[10] / nt);
  }
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
  // Two 128-query blocks per CTA (one CTA per SM) over key tiles of 64 keys, S double-buffered in tensor memory; a single
  // query block when there is only one.  TFMQ_ATTN_NQ overrides (experiments).
  static const int nq_env = getenv("TFMQ_ATTN_NQ") ? atoi(getenv("TFMQ_ATTN_NQ")) : 0;
  const int qblocks = (d->tq + 127) / 128;
  const int nq = nq_env ? nq_env : (qblocks >= 2 ? 2 : 1);
  if (d->d == 32) {
    if (nq >= 2) return launch_attn_tc<32, 64, 64, 2, 6>(ctx, d, st);
    return launch_attn_tc<32, 64, 64, 1, 6>(ctx, d, st);
  }
  // 128-byte operand rows: head dims 16, 24 (zero-filled), 40 (runs as 48), 48, 56, 64
  if (d->d < 32) return launch_attn_tc<32, 128, 64, 1, 3>(ctx, d, st);
  if (d->d <= 48) {
    if (nq >= 2) return launch_attn_tc<48, 128, 64, 2, 4>(ctx, d, st);
    return launch_attn_tc<48, 128, 64, 1, 4>(ctx, d, st);
  }
  if (nq >= 2) return launch_attn_tc<64, 128, 64, 2, 4>(ctx, d, st);
  return launch_attn_tc<64, 128, 64, 1, 4>(ctx, d, st);
}
