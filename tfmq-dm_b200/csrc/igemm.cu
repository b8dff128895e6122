// tcgen05 implicit-GEMM convolution kernels + their C-ABI launchers.
// See igemm.cuh for the pipeline description and DESIGN.md for the roofline.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_fp16.h>

#include "igemm.cuh"

#include "ctx.h"
#include "ptx.cuh"

namespace tfmq {

// ---------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------
// Persistent: grid = min(#tiles, #SMs); each CTA walks tiles t = blockIdx.x, +gridDim.x, ... with the
// M index fastest, so CTAs running together share one weight tile in L2.
//   warps 0-7   epilogue: TMEM -> registers -> swizzled smem chunk (+ TMA-prefetched residual) -> TMA store,
//               overlapped with the next tile's main loop through the second accumulator stage
//   warp 8      TMA producer
//   warp 9      UMMA issuer (owns TMEM: 1 or 2 accumulator stages)
//   warps 10-13 operand transform (int4 -> s8 expansion / tf32 hi-lo split); highest warp ids because the SM
//               arbiter favours them and they sit on the critical path of the main loop
// w4a8 pipeline: A ring (TMA -> UMMA), s8 B ring (transform warps -> UMMA); the packed int4 weights never sit in
// shared memory, they are prefetched from L2 into registers.  See the transform warps.
// GroupNorm partial sums of one epilogue chunk: thread = (column, block of rows); loads batched by full unrolling
template <int CW>
__device__ __forceinline__ float2 chunk_col_partial(const uint8_t* buf, int et) {
  constexpr int PARTS = IGEMM_EPI_WARPS * 32 / CW, ROWS = 128 / PARTS;
  const int col = et & (CW - 1), part = et / CW;
  const uint32_t piece = (uint32_t)(col >> 2), within = (uint32_t)(col & 3) * 4u;
  float v[ROWS];
#pragma unroll
  for (int rr = 0; rr < ROWS; ++rr) {
    const int row = part * ROWS + rr;
    const uint32_t rsw = (uint32_t)((CW == 32) ? (row & 7) : ((row >> 1) & 3));
    v[rr] = *reinterpret_cast<const float*>(buf + row * (CW * 4) + ((piece ^ rsw) << 4) + within);
  }
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int rr = 0; rr < ROWS; ++rr) s1 += v[rr], s2 = fmaf(v[rr], v[rr], s2);
  return make_float2(s1, s2);
}

template <int MODE, int CG>
__device__ __forceinline__ void umma_fp(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  if (MODE == MODE_F16 && CG == 2) umma_f16_2cta(tmem_d, adesc, bdesc, idesc, acc);
  else if (MODE == MODE_F16) umma_f16(tmem_d, adesc, bdesc, idesc, acc);
  else umma_tf32(tmem_d, adesc, bdesc, idesc, acc);
}

#define PROF_T(idx)                                         \
  if (prof_on) {                                            \
    const long long now_ = clock64();                       \
    prof_acc[idx] += now_ - prof_t;                         \
    prof_t = now_;                                          \
  }

template <int MODE, int CG>
__global__ void __launch_bounds__(IGEMM_THREADS, 1)
igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
             const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmOut,
             const __grid_constant__ CUtensorMap tmRes, const IgemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // dynamic smem is only guaranteed 16-B aligned: skip to the next 1024-B boundary (128B swizzle atoms).
  // Pointer arithmetic on the __shared__ array (not an integer round trip) keeps the address space
  // known to the compiler, so every access below is LDS/STS rather than a generic LD/ST.
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = p.stages;        // pipeline stages (w4a8: depth of the A ring)
  const int SU = p.u_stages;     // w4a8: depth of the unpacked-B ring
  const int ACC = p.acc_stages;
  constexpr bool W4 = (MODE == MODE_W4A8);
  constexpr bool FP = (MODE == MODE_TF32 || MODE == MODE_F16);   // fp32-accurate modes: 3 error-compensated products

  // w4a8:   [A ring: S x 16 KB][s8 B ring: SU x u_bytes][packed int4 ring: SP x p_bytes] | other modes: [S x stage_bytes]
  const int SP = p.p_stages;
  uint8_t* u_ring = smem + (size_t)S * IGEMM_A_BYTES;
  uint8_t* p_ring = u_ring + (size_t)SU * p.u_bytes;
  // NB epilogue chunk buffers: 2, or 3 when a residual is folded in (its TMA load then runs two chunks ahead)
  const int NB = p.epi_bufs;
  uint8_t* ebuf = W4 ? p_ring + (size_t)SP * p.p_bytes : smem + (size_t)S * p.stage_bytes;
  float4* chp = reinterpret_cast<float4*>(ebuf + NB * 128 * 32 * 4);        // [tile_n] per-channel constants
  // [stat_imgs][tile_n] per-image, per-channel (sum, sumsq): a tile of a small feature map spans several images
  float2* cstat = reinterpret_cast<float2*>(ebuf + NB * 128 * 32 * 4 + p.tile_n * 16);
  float2* cpart = cstat + p.stat_imgs * p.tile_n;     // [2][parts][chunk_w] row-block partials of the last two chunks
  uint64_t* bars = reinterpret_cast<uint64_t*>(ebuf + NB * 128 * 32 * 4 + p.tile_n * 16 + p.stat_imgs * p.tile_n * 8 +
                                               2 * 256 * 8);
  // barrier slots (8 bytes each, 64 slots; 128 when p.deep_bars): the two pipelines use the first 24 differently.
  // deep_bars (w4a8 with more than 4 slots in the s8 B or packed ring, TFMQ_IGEMM_USTAGES / _PSTAGES up to 8): the four ring
  // barrier arrays move to a second block of 64 slots; the default layout is untouched.
  const bool deep = W4 && p.deep_bars != 0;
  uint64_t* full_tma = bars;            // [<=8] TMA bytes landed (w4a8: the A tile; CTA pair: of both CTAs, at the leader)
  uint64_t* empty = bars + 8;           // [<=8] UMMAs that read the stage (w4a8: the A slot) retired
  uint64_t* full_xf = deep ? bars + 64 : bars + 16;   // [<=8] transform warps done (w4a8: s8 B slot written, [<=4], deep [<=8])
  uint64_t* empty_u = deep ? bars + 72 : bars + 20;   // w4a8 [<=4], deep [<=8]: UMMAs that read the s8 B slot retired
  uint64_t* acc_full = bars + 24;       // [2] accumulator stage complete
  uint64_t* acc_empty = bars + 26;      // [2] accumulator stage drained by the epilogue
  uint64_t* res_full = bars + 40;       // [<=3] residual chunk landed in the epilogue buffer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 30);
  uint64_t* full_p = deep ? bars + 80 : bars + 32;    // w4a8 [<=4], deep [<=8]: packed int4 tile landed
  uint64_t* empty_p = deep ? bars + 88 : bars + 36;   // w4a8 [<=4], deep [<=8]: the transform warps hold the packed tile in registers

  const int tiles_x = p.W / p.tw;
  const int tiles_y = p.H / p.th;
  const int tiles_m = tiles_x * tiles_y * ((p.n_img + p.tn - 1) / p.tn);
  const int tiles_n = p.cout / p.tile_n;
  // CG == 2: a cluster of two CTAs (one TPC) works on two M-adjacent tiles of the same N tile with
  // tcgen05.mma.cta_group::2 (M = 256): every CTA stages its own 128 pixels of A but only HALF of the N rows of B.
  // (the fp16-split mode as a pair: each CTA's TMA producer loads its own A planes and HALF of the weight rows of both B planes;
  //  the UMMAs then read 7.5 KB instead of 11 KB of shared memory per K = 16 step and CTA)
  static_assert(CG == 1 || MODE == MODE_W4A8 || MODE == MODE_F16, "the CTA-pair path exists for the w4a8 and fp16-split modes");
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
  const int unit0 = (CG == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;     // cluster (or CTA) index
  const int nunits = (CG == 2) ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int tiles_mu = tiles_m / CG;
  // split-K (MODE_F16, p.ksplit > 1): unit = (K range, N tile, M tile); the partial tiles meet in global memory through
  // TMA reduce-adds.  Reductions over the pixel axis (weight gradients) have few output tiles and very long K.
  const int KS = (MODE == MODE_F16) ? p.ksplit : 1;
  const int mn_units = tiles_mu * tiles_n;
  const int total_units = (p.dbg & 64) ? 0 : mn_units * KS;     // dbg 64: timing experiment, prologue + teardown only
  const int b_rows = p.b_rows[rank];       // weight rows of the N tile this CTA stages
  const int kchunks = (p.cin + p.kchunk - 1) / p.kchunk;
  const int taps = p.ksize * p.ksize;
  const int nkb = taps * kchunks;

  if (threadIdx.x == 0) {
    if (W4) {
      for (int s = 0; s < S; ++s) {
        mbar_init(&full_tma[s], (CG == 2 && rank == 0) ? 2 : 1);   // own TMA (+ the peer's "my A landed" arrive)
        mbar_init(&empty[s], 1);
      }
      for (int s = 0; s < SU; ++s) {
        mbar_init(&full_xf[s], p.xf_gw * CG);                      // one transform group, of both CTAs of a pair
        mbar_init(&empty_u[s], 1);
      }
      for (int s = 0; s < SP; ++s) {
        mbar_init(&full_p[s], 1);
        mbar_init(&empty_p[s], p.xf_gw);
      }
    } else {
      for (int s = 0; s < S; ++s) {
        mbar_init(&full_tma[s], (CG == 2 && rank == 0) ? 2 : 1);   // own TMA (+ the peer's "my tiles landed" arrive)
        mbar_init(&full_xf[s], IGEMM_XF_WARPS);
        mbar_init(&empty[s], 1);
      }
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], IGEMM_EPI_WARPS * CG);
    }
    for (int s = 0; s < NB; ++s) mbar_init(&res_full[s], 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (FP || CG == 2) tma_prefetch_desc(&tmB2);
    if (MODE == MODE_F16) tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmOut);
    if (p.res) tma_prefetch_desc(&tmRes);
  }
  if (warp == IGEMM_WARP_MMA) {
    if (CG == 2) tmem_alloc2(tmem_slot, (uint32_t)p.tmem_cols);
    else tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  }
  if (W4 && rank == CG - 1) {
    // the 16 B rows after the weight rows of every s8 B slot: first row = 0x01 bytes, the rest zero (swizzle-
    // invariant); in a CTA pair they are the last rows of the second CTA's half.  Accumulator column tile_n then
    // holds sum_k a[m][k], and the weight zero point is applied in the epilogue instead of per code.
    for (int i = threadIdx.x; i < SU * 16 * 8; i += IGEMM_THREADS) {
      const int st_i = i / 128, rem = i - st_i * 128;
      const uint32_t fill = (rem < 8) ? 0x01010101u : 0u;
      *reinterpret_cast<uint4*>(u_ring + (size_t)st_i * p.u_bytes + (size_t)b_rows * 128 + rem * 16) =
          make_uint4(fill, fill, fill, fill);
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all();    // the peer's barriers are initialised before anything arrives on them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // programmatic dependent launch: everything above overlapped the tail of the previous kernel; nothing below may touch
  // global memory before that kernel has completed
  griddep_launch_dependents();
  griddep_wait();

  const bool need_a_lo = FP && (p.pass_flags & PASS_LO_HI);
  const bool need_b_lo = FP && (p.pass_flags & PASS_HI_LO);
  const bool xf_split = (MODE == MODE_TF32) && need_a_lo;      // tf32: A lo is made in the kernel; f16: it arrives by TMA
  const bool two_acc = need_a_lo || need_b_lo;                 // tf32: separate accumulator for the small terms
  const uint32_t acc_cols = W4 ? (uint32_t)p.tile_n + 16u : (uint32_t)p.tile_n * (two_acc ? 2u : 1u);

  if (warp == IGEMM_WARP_TMA) {
    // ===================================================== TMA producer
    // (whole warp convergent, one elected lane issues: coordinates and addresses stay warp-uniform)
    uint32_t tx_bytes = IGEMM_A_BYTES;          // w4a8: the A tile only (B comes through the transform warps)
    if (MODE == MODE_I8) tx_bytes += (uint32_t)p.tile_n * 128u;
    if (FP) tx_bytes += (uint32_t)b_rows * 128u * (need_b_lo ? 2u : 1u);     // CTA pair: this CTA's half of the weight rows
    if (MODE == MODE_F16 && need_a_lo) tx_bytes += IGEMM_A_BYTES;
    int s = 0;
    uint32_t par = 0;
    for (int ut = unit0; ut < total_units; ut += nunits) {
      const int ks = (KS > 1) ? ut / mn_units : 0, umn = (KS > 1) ? ut - ks * mn_units : ut;
      const int kb0 = (KS > 1) ? (int)((long long)ks * nkb / KS) : 0, kb1 = (KS > 1) ? (int)((long long)(ks + 1) * nkb / KS) : nkb;
      int mt = (umn % tiles_mu) * CG + (int)rank;
      const int c_out0 = (umn / tiles_mu) * p.tile_n;
      const int tx = mt % tiles_x;
      mt /= tiles_x;
      const int ty = mt % tiles_y;
      const int n0 = (mt / tiles_y) * p.tn, y0 = ty * p.th, x0 = tx * p.tw;
      int tap = kb0 / kchunks, kc = kb0 - tap * kchunks;
      int ky = tap / p.ksize, kx = tap - ky * p.ksize;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait_relaxed(&empty[s], par ^ 1u);
        uint8_t* st = W4 ? smem + (size_t)s * IGEMM_A_BYTES : smem + (size_t)s * p.stage_bytes;
        if (elect_one_sync()) {
          if (p.dbg & 1) {
            mbar_arrive(&full_tma[s]);
          } else {
            mbar_expect_tx(&full_tma[s], tx_bytes);
            tma_load_4d(st, &tmA, &full_tma[s], kc * p.kchunk, x0 * p.stride + kx + p.off, y0 * p.stride + ky + p.off,
                        n0);
            if (MODE == MODE_F16 && need_a_lo)
              tma_load_4d(st + p.offA_lo, &tmA2, &full_tma[s], kc * p.kchunk, x0 * p.stride + kx + p.off,
                          y0 * p.stride + ky + p.off, n0);
            if (!W4) {
              tma_load_2d(st + p.offB, &tmB, &full_tma[s], tap * p.cin + kc * p.kchunk, c_out0 + p.b_row0[rank]);
              if (need_b_lo)
                tma_load_2d(st + p.offB_lo, &tmB2, &full_tma[s], tap * p.cin + kc * p.kchunk, c_out0 + p.b_row0[rank]);
            }
          }
        }
        __syncwarp();
        if (++kc == kchunks) {
          kc = 0, ++tap;
          if (++kx == p.ksize) kx = 0, ++ky;
        }
        if (++s == S) s = 0, par ^= 1u;
      }
    }
  } else if (warp == IGEMM_WARP_TMB) {
    // ===================================================== TMA producer of the packed int4 weight tiles (w4a8)
    if (W4) {
      const CUtensorMap* tmBr = (CG == 2 && rank == 1) ? &tmB2 : &tmB;     // the two halves differ in box height
      const int b_row0 = p.b_row0[rank];
      const uint32_t tx_bytes = (uint32_t)b_rows * 64u;
      int s = 0;
      uint32_t par = 0;
      for (int ut = unit0; ut < total_units; ut += nunits) {
        const int c_out0 = (ut / tiles_mu) * p.tile_n;
        int kc = 0, tap = 0;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait_relaxed(&empty_p[s], par ^ 1u);
          if (elect_one_sync()) {
            mbar_expect_tx(&full_p[s], tx_bytes);
            // packed bytes: column = (tap*cin + kc*128)/2
            tma_load_2d(p_ring + (size_t)s * p.p_bytes, tmBr, &full_p[s], (tap * p.cin + kc * p.kchunk) >> 1,
                        c_out0 + b_row0);
          }
          __syncwarp();
          if (++kc == kchunks) kc = 0, ++tap;
          if (++s == SP) s = 0, par ^= 1u;
        }
      }
    }
  } else if (warp == IGEMM_WARP_MMA) {
   if (CG == 2 && rank != 0) {
    // ===================================================== CTA pair, second CTA: its UMMA warp only reports
    // "my A tile of slot s has landed" to the leader's barrier (the leader issues the UMMAs for both)
    const uint32_t remote = mapa_u32(smem_u32(full_tma), 0);
    int s = 0;
    uint32_t par = 0;
    for (int ut = unit0; ut < total_units; ut += nunits) {
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&full_tma[s], par);
        if (lane == 0) mbar_arrive_remote(remote + 8u * (uint32_t)s);
        __syncwarp();
        if (++s == S) s = 0, par ^= 1u;
      }
    }
   } else {
    // ===================================================== UMMA issuer (CTA pair: the leader CTA only)
    // The whole warp runs this loop convergently and every operand of the tcgen05 instructions is computed
    // outside the elected-thread region, so descriptors live in uniform registers and each UTC*MMA issues
    // without a per-instruction register shuffle; the tensor pipe starves on anything slower.
    const uint32_t umma_n = (uint32_t)p.tile_n + (W4 ? 16u : 0u);
    const uint32_t idesc = (MODE == MODE_TF32)  ? idesc_tf32(128, umma_n)
                           : (MODE == MODE_F16) ? idesc_f16(128 * CG, umma_n)
                                                : idesc_i8_u8s8(128 * CG, umma_n);
    const uint32_t smem_base = smem_u32(smem);
    const uint64_t d_a = smem_desc_sw128(smem_base);
    const uint64_t d_b = W4 ? smem_desc_sw128(smem_u32(u_ring)) : smem_desc_sw128(smem_base + p.offB);
    const uint64_t d_alo = smem_desc_sw128(smem_base + p.offA_lo);
    const uint64_t d_blo = smem_desc_sw128(smem_base + p.offB_lo);
    const uint32_t stage_units = W4 ? (IGEMM_A_BYTES >> 4) : (p.stage_bytes >> 4);   // descriptor address units per A slot
    const uint32_t u_units = p.u_bytes >> 4;
    const int nslice_last = (p.cin - (kchunks - 1) * p.kchunk) / p.kslice;   // valid 32-byte K slices of the last chunk
    const bool wait_xf = xf_split;              // tf32 with a split A: the transform warps' barrier gates the stage
    uint32_t tcount = 0;
    int s = 0, su = 0;
    uint32_t par = 0, s_units = 0, pu = 0, su_units = 0;
    const bool prof_on = p.prof != nullptr && lane == 0;
    long long prof_acc[6] = {0, 0, 0, 0, 0, 0};
    long long prof_t = prof_on ? clock64() : 0;
    for (int ut = unit0; ut < total_units; ut += nunits, ++tcount) {
      const uint32_t as = tcount % ACC;
      PROF_T(2)
      // the epilogue (of both CTAs of a pair) has drained this accumulator stage
      mbar_wait(&acc_empty[as], ((tcount / ACC) & 1u) ^ 1u);
      tc_fence_after();
      PROF_T(0)
      const uint32_t tmem_d = tmem_base + as * acc_cols;
      // tf32: the two small cross terms accumulate in their own TMEM columns so the (truncating)
      // tensor-core accumulator adds them to a small running sum, not to the large main one
      const uint32_t tmem_lo = tmem_d + (uint32_t)p.tile_n;
      uint32_t accumulate = 0, accumulate_lo = 0;
      const int ks = (KS > 1) ? ut / mn_units : 0;
      const int kb0 = (KS > 1) ? (int)((long long)ks * nkb / KS) : 0, kb1 = (KS > 1) ? (int)((long long)(ks + 1) * nkb / KS) : nkb;
      int kc = kb0 % kchunks;
      for (int kb = kb0; kb < kb1; ++kb) {
        PROF_T(2)
        if (W4) {
          mbar_wait(&full_tma[s], par);
          mbar_wait(&full_xf[su], pu);
        } else {
          // a transform warp arrives only after the stage's TMA barrier completed, so its barrier alone gates the stage
          if (wait_xf) mbar_wait(&full_xf[s], par);
          else mbar_wait(&full_tma[s], par);
        }
        tc_fence_after();
        PROF_T(1)
        const int nslice = (kc == kchunks - 1) ? nslice_last : 4;
        const uint64_t a_hi = d_a + s_units, b_hi = d_b + (W4 ? su_units : s_units);
        const uint64_t a_lo = d_alo + s_units, b_lo = d_blo + s_units;
        const bool last = kb == kb1 - 1;
        if (elect_one_sync()) {
          if (FP && nslice == 4) {
            // full channel chunk: unpredicated UMMAs, descriptors advanced in uniform registers (see the w4a8 branch below; with
            // three products the predicated form issues a k-block in ~1070 clocks against 672 of tensor-pipe time at N = 112)
            if (need_a_lo) {
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_fp<MODE, CG>(tmem_lo, a_lo + 2 * k, b_hi + 2 * k, idesc, (k > 0) | accumulate_lo);
            }
            if (need_b_lo) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_fp<MODE, CG>(tmem_lo, a_hi + 2 * k, b_lo + 2 * k, idesc, (k > 0) | accumulate_lo | (need_a_lo ? 1u : 0u));
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_fp<MODE, CG>(tmem_d, a_hi + 2 * k, b_hi + 2 * k, idesc, (k > 0) | accumulate);
          } else if (FP) {
            if (need_a_lo) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (k < nslice) umma_fp<MODE, CG>(tmem_lo, a_lo + 2 * k, b_hi + 2 * k, idesc, (k > 0) | accumulate_lo);
            }
            if (need_b_lo) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (k < nslice)
                  umma_fp<MODE, CG>(tmem_lo, a_hi + 2 * k, b_lo + 2 * k, idesc, (k > 0) | accumulate_lo | (need_a_lo ? 1u : 0u));
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (k < nslice) umma_fp<MODE, CG>(tmem_d, a_hi + 2 * k, b_hi + 2 * k, idesc, (k > 0) | accumulate);
          } else if (nslice == 4) {
            // full channel chunk (every k-block when cin is a multiple of 128): four unpredicated UMMAs, descriptors advanced in
            // uniform registers -- the predicated form below re-materialises both descriptors per UMMA (25 R2UR per k-block),
            // and the issue loop is what bounds the narrow tiles of the small maps (profiles/r2w_epilogue_probes.md)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (CG == 2) umma_i8_2cta(tmem_d, a_hi + 2 * k, b_hi + 2 * k, idesc, (k > 0) | accumulate);
              else umma_i8(tmem_d, a_hi + 2 * k, b_hi + 2 * k, idesc, (k > 0) | accumulate);
            }
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (k < nslice) {
                if (CG == 2) umma_i8_2cta(tmem_d, a_hi + 2 * k, b_hi + 2 * k, idesc, (k > 0) | accumulate);
                else umma_i8(tmem_d, a_hi + 2 * k, b_hi + 2 * k, idesc, (k > 0) | accumulate);
              }
          }
          if (CG == 2) {          // slot / accumulator hand-over goes to the same barrier of both CTAs
            umma_commit_2cta(&empty[s], 3);
            if (W4) umma_commit_2cta(&empty_u[su], 3);
            if (last) umma_commit_2cta(&acc_full[as], 3);
          } else {
            umma_commit(&empty[s]);
            if (W4) umma_commit(&empty_u[su]);
            if (last) umma_commit(&acc_full[as]);
          }
        }
        __syncwarp();
        accumulate = 1, accumulate_lo = 1;
        if (++kc == kchunks) kc = 0;
        s_units += stage_units;
        if (++s == S) s = 0, s_units = 0, par ^= 1u;
        if (W4) {
          su_units += u_units;
          if (++su == SU) su = 0, su_units = 0, pu ^= 1u;
        }
      }
    }
    if (prof_on)
      for (int i = 0; i < 6; ++i) p.prof[blockIdx.x * 16 + 10 + i] = prof_acc[i];
   }
  } else if (warp >= IGEMM_WARP_XF0) {
    // ===================================================== transform warps
    if (W4) {
      // int4 -> s8 expansion of the weight operand.  The packed tile of a k-block arrives by TMA in a small ring of
      // its own (dense [b_rows][64 B]); a transform group reads it into registers, gives the slot straight back, and
      // writes the 128B-swizzled K-major s8 tile into the next free slot of the B ring.  The three rings (A, packed B,
      // s8 B) have independent depths, so a slot is only held for as long as ITS data is in flight.  Two groups of two
      // warps work on alternate k-blocks (ring depths are even, so a group sees consecutive phases of every barrier
      // it waits on).  Codes stay unsigned (0..15): two ANDs and a shift per word; the zero point is folded out
      // through the ones row.  Thread tg of a group owns pieces (row (tg >> 2) + 16 i, K slice tg & 3).
      // Group width GW (p.xf_gw): two warps per group for wide tiles; for narrow tiles (<= 64 weight rows per CTA) FOUR groups
      // of one warp each take every fourth k-block -- a k-block of a 64-wide tile is a latency chain (wait for the packed tile,
      // LDS, wait for a free s8 slot, STS, proxy fence, arrive) of ~1000 clocks per group, so the number of groups in flight,
      // not the unpack work, sets the k-block rate of the small-map layers.  A group pass covers RP = 8 GW rows.
      constexpr int NP = (CG == 2) ? 8 : 15;            // pieces per thread: ceil(max rows / RP)
      const int GW = p.xf_gw, RP = 8 * GW;
      const int grp = (warp - IGEMM_WARP_XF0) / GW;
      const int tg = threadIdx.x - (IGEMM_WARP_XF0 + grp * GW) * 32;
      const int sub = tg & 3, row0 = tg >> 2;
      const uint32_t swz = (uint32_t)(row0 & 7);
      const uint32_t rd0 = (uint32_t)row0 * 64u + (uint32_t)sub * 16u;
      const uint32_t wr_lo = (uint32_t)row0 * 128u + (((2u * sub) ^ swz) << 4);
      const uint32_t wr_hi = (uint32_t)row0 * 128u + (((2u * sub + 1u) ^ swz) << 4);
      const uint32_t xf_remote = (CG == 2 && rank != 0) ? mapa_u32(smem_u32(full_xf), 0) : 0u;
      const int NG = IGEMM_XF_WARPS / GW;
      uint32_t it = 0;                                   // k-blocks of this CTA so far (all units)
      int su = 0, sp = 0;
      uint32_t pu = 0, pp = 0;
      for (int ut = unit0; ut < total_units; ut += nunits) {
        int kc = 0;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          if ((int)(it % NG) == grp) {
            int rem = p.cin - kc * p.kchunk;
            if (rem > p.kchunk) rem = p.kchunk;
            const bool active = sub < (rem >> 5) && !(p.dbg & 8);
            mbar_wait(&full_p[sp], pp);
            const uint8_t* src = p_ring + (size_t)sp * p.p_bytes + rd0;
            uint4 pk[NP];
            if (active) {
#pragma unroll
              for (int i = 0; i < NP; ++i)
                if (row0 + RP * i < b_rows) pk[i] = *reinterpret_cast<const uint4*>(src + i * RP * 64);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_p[sp]);    // release: the reads above are ordered before it
            mbar_wait(&empty_u[su], pu ^ 1u);
            uint8_t* dst = u_ring + (size_t)su * p.u_bytes;
            if (active) {
#pragma unroll
              for (int i = 0; i < NP; ++i) {
                if (row0 + RP * i < b_rows) {
                  const uint4 k4 = pk[i];
                  uint4 lo, hi;
                  lo.x = k4.x & 0x0F0F0F0Fu, lo.y = k4.y & 0x0F0F0F0Fu, lo.z = k4.z & 0x0F0F0F0Fu, lo.w = k4.w & 0x0F0F0F0Fu;
                  hi.x = (k4.x >> 4) & 0x0F0F0F0Fu, hi.y = (k4.y >> 4) & 0x0F0F0F0Fu;
                  hi.z = (k4.z >> 4) & 0x0F0F0F0Fu, hi.w = (k4.w >> 4) & 0x0F0F0F0Fu;
                  *reinterpret_cast<uint4*>(dst + wr_lo + i * RP * 128) = lo;
                  *reinterpret_cast<uint4*>(dst + wr_hi + i * RP * 128) = hi;
                }
              }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              if (CG == 2 && rank != 0) mbar_arrive_remote(xf_remote + 8u * (uint32_t)su);
              else mbar_arrive(&full_xf[su]);
            }
          }
          if (++kc == kchunks) kc = 0;
          if (++su == SU) su = 0, pu ^= 1u;
          if (++sp == SP) sp = 0, pp ^= 1u;
        }
      }
    } else if (xf_split) {
      // split the fp32 A tile into hi = a & 0xFFFFE000 (exactly a tf32) and lo = a - hi; the four warps share a stage
      const int t = threadIdx.x - IGEMM_WARP_XF0 * 32;
      constexpr int XF_T = IGEMM_XF_WARPS * 32;
      int s = 0;
      uint32_t par = 0;
      for (int ut = unit0; ut < total_units; ut += nunits) {
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_tma[s], par);
          uint8_t* st = smem + (size_t)s * p.stage_bytes;
          uint4* a = reinterpret_cast<uint4*>(st);
          uint4* al = reinterpret_cast<uint4*>(st + p.offA_lo);
#pragma unroll
          for (int i = 0; i < 1024 / XF_T; ++i) {
            const int idx = t + XF_T * i;
            uint4 v = a[idx], h, l;
            h.x = v.x & 0xFFFFE000u;
            h.y = v.y & 0xFFFFE000u;
            h.z = v.z & 0xFFFFE000u;
            h.w = v.w & 0xFFFFE000u;
            l.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(h.x));
            l.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(h.y));
            l.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(h.z));
            l.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(h.w));
            a[idx] = h;
            al[idx] = l;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&full_xf[s]);
          if (++s == S) s = 0, par ^= 1u;
        }
      }
    }
    // other modes: nothing to transform; the UMMA issuer waits on the TMA barrier alone
  } else {
    // ===================================================== epilogue warps (0..7)
    // All global I/O of the epilogue is TMA: the residual chunk is prefetched into a swizzled smem
    // buffer, every thread folds its part of an accumulator row into it (conflict-free float4 accesses), and
    // the buffer is stored back with one bulk tensor store.  Two buffers alternate across chunks.
    // Two warps share a TMEM lane quarter: warp w handles columns [half*CW/2, (half+1)*CW/2) of every chunk.
    constexpr int ET = IGEMM_EPI_WARPS * 32;
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = warp >> 2;             // which half of a chunk's columns
    const int et = threadIdx.x;             // 0..ET-1 within the epilogue group
    const int r = q * 32 + lane;            // accumulator row = pixel within the tile
    const int CW = p.chunk_w;               // 32 (128B swizzle) or 16 (64B swizzle) channels per chunk
    const int HC = CW >> 1;                 // columns per thread per chunk
    const int nchunks = p.tile_n / CW;
    const uint32_t chunk_bytes = 128u * (uint32_t)CW * 4u;
    // byte offset (within the row) of 16-byte piece k: (k*16) ^ sw16
    const uint32_t sw16 = (uint32_t)((CW == 32) ? (r & 7) : ((r >> 1) & 3)) << 4;
    const uint32_t row_off = (uint32_t)r * (uint32_t)(CW * 4);
    const int hw_t = p.th * p.tw;
    const bool emb_folded = p.emb && p.tn == 1;   // one image per tile: the embedding row joins the bias
    float a_scale = 1.f;
    int za = 0;
    if (MODE == MODE_W4A8) {
      a_scale = p.aq[0];
      za = (int)p.aq[1];
    }
    uint32_t tcount = 0, g = 0;             // tiles / chunks processed by this CTA
    uint32_t gb = 0, gph = 0;               // g % NB (chunk buffer) and (g / NB) & 1 (phase of its residual barrier)
    int chp_c0 = -1, chp_n0 = -1;           // N tile / image the per-channel constants in smem were computed for
    const int LA = NB - 1;                  // residual loads run LA chunks ahead of the fold
    const bool prof_on = p.prof != nullptr && et == 0;
    long long prof_acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long prof_t = prof_on ? clock64() : 0;
    const uint32_t acc_empty_remote = (CG == 2 && rank != 0) ? mapa_u32(smem_u32(acc_empty), 0) : 0u;
    for (int ut = unit0; ut < total_units; ut += nunits, ++tcount) {
      const int ks = (KS > 1) ? ut / mn_units : 0, umn = (KS > 1) ? ut - ks * mn_units : ut;
      int mt = (umn % tiles_mu) * CG + (int)rank;
      const int c_out0 = (umn / tiles_mu) * p.tile_n;
      const int tx = mt % tiles_x;
      mt /= tiles_x;
      const int ty = mt % tiles_y;
      const int n0 = (mt / tiles_y) * p.tn, y0 = ty * p.th, x0 = tx * p.tw;
      int n_l = n0 + r / hw_t;
      if (n_l >= p.n_img) n_l = p.n_img - 1;   // rows past the batch are clipped by the TMA store

      // residual prefetch for the first chunk overlaps the wait for the accumulator
      if (et == 0) {
        tma_store_wait_read<1>();             // the store that last used this buffer has read it
        if (p.res) {
          // all stores but the last one have been read, so the LA buffers after the last one's are free
          for (int j = 0; j < LA && j < nchunks; ++j) {
            const uint32_t b = (gb + (uint32_t)j) % (uint32_t)NB;
            mbar_expect_tx(&res_full[b], chunk_bytes);
            tma_load_4d(ebuf + b * chunk_bytes, &tmRes, &res_full[b], c_out0 + j * CW, x0, y0, n0);
          }
          // pull the rest of the residual tile into L2 while the main loop of this tile runs
          for (int ci = LA; ci < nchunks; ++ci) tma_prefetch_l2_4d(&tmRes, c_out0 + ci * CW, x0, y0, n0);
        }
      }
      // per-channel constants of this N tile; a persistent CTA walks the units M-fastest, so consecutive units usually
      // share the N tile (always, when one tile covers all output channels) and the constants are kept
      const bool chp_keep = tcount > 0 && c_out0 == chp_c0 && (!emb_folded || n0 == chp_n0) && KS == 1;
      chp_c0 = c_out0, chp_n0 = n0;
      if (!chp_keep) named_bar_sync(1, ET);   // previous tile is done with chp
      for (int ch = et; ch < (chp_keep ? 0 : p.tile_n); ch += ET) {
        const int c = c_out0 + ch;
        float sc = 1.f, bi = 0.f;
        int ws = 0, zw = 0;
        if (MODE == MODE_W4A8) {
          sc = a_scale * p.wscale[c];
          ws = za * p.wsum[c];
          zw = p.wzp[c];
        } else if (p.wscale) {
          sc = p.wscale[c];
        }
        if (p.bias && ks == 0) bi = p.bias[c];      // split-K: the bias joins the first partial tile only
        if (emb_folded) bi += p.emb[(long long)n0 * p.emb_ld + c];
        chp[ch] = make_float4(sc, bi, __int_as_float(ws), __int_as_float(zw));
      }
      const uint32_t as = tcount % ACC;
      PROF_T(0)
      mbar_wait_relaxed(&acc_full[as], (tcount / ACC) & 1u);
      tc_fence_after();
      PROF_T(1)
      const uint32_t tmem_d = tmem_base + as * acc_cols + ((uint32_t)(q * 32) << 16);
      named_bar_sync(1, ET);                  // chp visible; first buffer known free
      int a_sum = 0;                          // sum_k a[m][k] of this row (ones-row column)
      if (MODE == MODE_W4A8) {
        uint32_t sv[8];
        tmem_ld8(tmem_d + (uint32_t)p.tile_n, sv);
        tmem_ld_wait();
        a_sum = (int)sv[0];
      }

      for (int ci = 0; ci < nchunks; ++ci, ++g, gb = (gb + 1 == (uint32_t)NB) ? 0u : gb + 1, gph ^= (gb == 0u)) {
        const int c0 = ci * CW + half * HC;   // first column (within the tile) this thread handles
        uint8_t* buf = ebuf + gb * chunk_bytes;
        uint32_t v[16];
        if (CW == 32) {
          tmem_ld16(tmem_d + (uint32_t)c0, v);
        } else {
          tmem_ld8(tmem_d + (uint32_t)c0, *reinterpret_cast<uint32_t(*)[8]>(&v[0]));
        }
        if (FP && two_acc) {
          uint32_t v2[16];
          if (CW == 32) {
            tmem_ld16(tmem_d + (uint32_t)(p.tile_n + c0), v2);
          } else {
            tmem_ld8(tmem_d + (uint32_t)(p.tile_n + c0), *reinterpret_cast<uint32_t(*)[8]>(&v2[0]));
          }
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
        }
        tmem_ld_wait();
        if (ci == nchunks - 1) {
          // last TMEM read of this tile: hand the accumulator stage back to the UMMA issuer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CG == 2 && rank != 0) mbar_arrive_remote(acc_empty_remote + 8u * as);
            else mbar_arrive(&acc_empty[as]);
          }
        }
        PROF_T(2)
        if (p.res) mbar_wait(&res_full[gb], gph);
        PROF_T(3)
        const float* embp = (p.emb && !emb_folded) ? p.emb + (long long)n_l * p.emb_ld + c_out0 + c0 : nullptr;
        const int npiece = HC >> 2;           // 16-byte pieces per thread: 4 (CW 32) or 2 (CW 16)
        const int k0 = half * npiece;
        if (FP && p.out_planes) {
          // output as fp16 hi / lo planes (CW == 32): the chunk buffer holds a [128][32] half tile of each plane, 64-byte
          // rows under the 64B swizzle; this thread converts its 16 columns and writes two 16-byte pieces per plane
          uint32_t hw[8], lw[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 ca = chp[c0 + 2 * j], cb = chp[c0 + 2 * j + 1];
            const float f0 = __uint_as_float(v[2 * j]) * ca.x + ca.y, f1 = __uint_as_float(v[2 * j + 1]) * cb.x + cb.y;
            uint16_t h0, h1, l0, l1;      // the saturating split of tfmq_act_prepare (elementwise.cu::split_h16x4)
            asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h0) : "f"(f0));
            asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h1) : "f"(f1));
            const float r0 = f0 - __half2float(__ushort_as_half(h0)), r1 = f1 - __half2float(__ushort_as_half(h1));
            asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(l0) : "f"(r0));
            asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(l1) : "f"(r1));
            hw[j] = (uint32_t)h0 | ((uint32_t)h1 << 16);
            lw[j] = (uint32_t)l0 | ((uint32_t)l1 << 16);
          }
          const uint32_t sw4 = (uint32_t)((r >> 1) & 3);
          uint8_t* rh = buf + (uint32_t)r * 64u;
          uint8_t* rl = rh + 128u * 64u;
          const uint32_t pa = (((uint32_t)(2 * half)) ^ sw4) << 4, pb = (((uint32_t)(2 * half + 1)) ^ sw4) << 4;
          *reinterpret_cast<uint4*>(rh + pa) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          *reinterpret_cast<uint4*>(rh + pb) = make_uint4(hw[4], hw[5], hw[6], hw[7]);
          *reinterpret_cast<uint4*>(rl + pa) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
          *reinterpret_cast<uint4*>(rl + pb) = make_uint4(lw[4], lw[5], lw[6], lw[7]);
        }
        const bool fold_f32 = !(FP && p.out_planes) && !(p.dbg & 32);   // dbg 32: timing experiment without the fold
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (k < npiece && fold_f32) {
            float4* slot = reinterpret_cast<float4*>(buf + row_off + ((uint32_t)((k0 + k) << 4) ^ sw16));
            float4 o;
            if (MODE == MODE_I8) {
              o = make_float4(__uint_as_float(v[4 * k]), __uint_as_float(v[4 * k + 1]), __uint_as_float(v[4 * k + 2]),
                              __uint_as_float(v[4 * k + 3]));
            } else {
              float f[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float4 cp = chp[c0 + 4 * k + j];
                if (MODE == MODE_W4A8)
                  f[j] = (float)((int)v[4 * k + j] - __float_as_int(cp.w) * a_sum - __float_as_int(cp.z)) * cp.x + cp.y;
                else
                  f[j] = __uint_as_float(v[4 * k + j]) * cp.x + cp.y;
              }
              if (embp) {
                const float4 e4 = *reinterpret_cast<const float4*>(embp + 4 * k);
                f[0] += e4.x, f[1] += e4.y, f[2] += e4.z, f[3] += e4.w;
              }
              if (p.res) {
                const float4 r4 = *slot;
                f[0] += r4.x, f[1] += r4.y, f[2] += r4.z, f[3] += r4.w;
              }
              o = make_float4(f[0], f[1], f[2], f[3]);
            }
            *slot = o;
          }
        }
        PROF_T(4)
        fence_proxy_async_smem();
        named_bar_sync(1, ET);                // all 128 rows of the chunk are in smem
        PROF_T(5)
        // the store (and the next residual load) go out BEFORE the statistics pass over the same buffer: the TMA engine
        // works while the warps sum columns, and thread 0's warp is the one every other warp waits for at the next barrier
        if (et == 0) {
          if (p.dbg & 16) {
            // timing experiment: no store
          } else if (KS > 1) tma_reduce_add_4d(&tmOut, buf, c_out0 + ci * CW, x0, y0, n0);
          else tma_store_4d(&tmOut, buf, c_out0 + ci * CW, x0, y0, n0);
          if (FP && p.out_planes) tma_store_4d(&tmRes, buf + 128u * 64u, c_out0 + ci * CW, x0, y0, n0);   // lo plane
          tma_store_commit();
          if (ci + 1 < nchunks) {
            tma_store_wait_read<1>();         // the other buffer's store has finished reading
            if (p.res && ci + LA < nchunks) {
              // chunk ci + LA goes into the buffer of chunk ci - 1 (NB = 3) / ci - 1 (NB = 2), whose store was just awaited
              const uint32_t b = (gb + (uint32_t)LA) % (uint32_t)NB;
              mbar_expect_tx(&res_full[b], chunk_bytes);
              tma_load_4d(ebuf + b * chunk_bytes, &tmRes, &res_full[b], c_out0 + (ci + LA) * CW, x0, y0, n0);
            }
          }
        }
        if (p.n_stat > 0) {
          // GroupNorm statistics of the finished output values: thread = (column, block of rows) sums its rows;
          // the partials of the previous chunk are combined here too, in a fixed order (bit-reproducible)
          // (done by the LAST warps of the group: thread 0 issues the TMA store and waits for buffers in this phase)
          if (ci > 0 && et >= ET - CW * p.stat_imgs) {
            // image `im` of the tile owns the row blocks [im * ppi, (im + 1) * ppi)
            const int e2 = et - (ET - CW * p.stat_imgs);
            const int im = e2 / CW, col = e2 - im * CW, ppi = (ET / CW) / p.stat_imgs;
            const float2* pp = cpart + ((g - 1) & 1) * ET + im * ppi * CW + col;
            float a1 = 0.f, a2 = 0.f;
            for (int k = 0; k < ppi; ++k) a1 += pp[k * CW].x, a2 += pp[k * CW].y;
            cstat[im * p.tile_n + (ci - 1) * CW + col] = make_float2(a1, a2);
          }
          cpart[(g & 1) * ET + et] = (CW == 32) ? chunk_col_partial<32>(buf, et) : chunk_col_partial<16>(buf, et);   // [part][col]
        }
        PROF_T(6)
        PROF_T(7)
        if (!p.res) named_bar_sync(1, ET);    // without a residual barrier, publish "next buffer is free"
        PROF_T(8)
      }
      if (p.n_stat > 0) {
        named_bar_sync(1, ET);                // the last chunk's partials are written
        if (et >= ET - CW * p.stat_imgs) {
          const int e2 = et - (ET - CW * p.stat_imgs);
          const int im = e2 / CW, col = e2 - im * CW, ppi = (ET / CW) / p.stat_imgs;
          const float2* pp = cpart + ((g - 1) & 1) * ET + im * ppi * CW + col;
          float a1 = 0.f, a2 = 0.f;
          for (int k = 0; k < ppi; ++k) a1 += pp[k * CW].x, a2 += pp[k * CW].y;
          cstat[im * p.tile_n + (nchunks - 1) * CW + col] = make_float2(a1, a2);
        }
        named_bar_sync(1, ET);                // every chunk's sums are in cstat
        for (int k = 0; k < p.n_stat; ++k) {
          const tfmq_gn_target tg = p.stat[k];
          const int g_lo = (tg.ch_off + c_out0) / tg.cpg;
          const int g_hi = (tg.ch_off + c_out0 + p.tile_n - 1) / tg.cpg;
          const int ng = g_hi - g_lo + 1;
          for (int idx = et; idx < ng * p.stat_imgs; idx += ET) {
            const int im = idx / ng, gi = g_lo + (idx - im * ng);
            if (n0 + im >= p.n_img) continue;       // batch tail: rows past the batch carry no image
            const int ch_lo = max(gi * tg.cpg - tg.ch_off, c_out0) - c_out0;
            const int ch_hi = min((gi + 1) * tg.cpg - tg.ch_off, c_out0 + p.tile_n) - c_out0;
            // the group's channels of this tile in fp32 (each term is already an fp32 sum over 128 rows), one conversion per
            // group: the double-precision pipe is 1/64 rate and these adds sat on the epilogue's critical path; the running
            // sums across tiles stay in double
            float a1 = 0.f, a2 = 0.f;
            for (int ch = ch_lo; ch < ch_hi; ++ch) {
              const float2 v2 = cstat[im * p.tile_n + ch];
              a1 += v2.x;
              a2 += v2.y;
            }
            double* dst = tg.stats + ((long long)(n0 + im) * tg.groups + gi) * 2;
            atomicAdd(dst, (double)a1);
            atomicAdd(dst + 1, (double)a2);
          }
        }
      }
    }
    if (et == 0) tma_store_wait_all<0>();     // global writes complete before the CTA retires
    PROF_T(9)
    if (prof_on)
      for (int i = 0; i < 10; ++i) p.prof[blockIdx.x * 16 + i] = prof_acc[i];
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all();    // neither CTA leaves while its peer may still signal its barriers / read its TMEM
  else __syncthreads();
  if (warp == IGEMM_WARP_MMA) {
    if (CG == 2) tmem_dealloc2(tmem_base, (uint32_t)p.tmem_cols);
    else tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// largest multiple of 16 that divides cout and is <= limit
static int pick_tile_n(int cout, int limit = 256) {
  for (int t = limit; t >= 16; t -= 16)
    if (cout % t == 0) return t;
  return 0;
}

// N tile for a persistent grid: the widest tile has the best operand reuse, but small feature maps then
// leave most SMs idle (8x8x16 images = 8 M tiles).  Pick the divisor of cout that minimises
//   waves(tiles) x (k-blocks x (k_fix + k_col * tile_n) + epilogue(tile_n))   [cycles, fitted to the phase counters]
//   a tile that is not a multiple of 32 wide runs its epilogue in 16-column chunks (64B swizzle): twice the chunk iterations
static int pick_tile_n_balanced(int cout, int limit, int tiles_m, int nkb, int sm_count, double k_fix, double k_col,
                                int step = 16) {
  static const double cw16 = getenv("TFMQ_IGEMM_CW16_PENALTY") ? atof(getenv("TFMQ_IGEMM_CW16_PENALTY")) : 2.0;
  int best = 0;
  double best_cost = 0.0;
  for (int t = limit - limit % step; t >= step; t -= step) {
    if (cout % t) continue;
    const long long tiles = (long long)tiles_m * (cout / t);
    const double waves = (double)((tiles + sm_count - 1) / sm_count);
    const double cost = waves * (nkb * (k_fix + k_col * t) + 3000.0 + 30.0 * t * (t % 32 ? cw16 : 1.0));
    if (!best || cost < best_cost * 0.97) best = t, best_cost = cost;   // prefer the wider tile on near-ties
  }
  return best;
}

struct TileGeom {
  int th, tw, tn;
};
static bool pick_geom(int H, int W, TileGeom* g) {
  if (!is_pow2(W) || !is_pow2(H)) return false;
  if (W >= 128) {
    g->tw = 128, g->th = 1, g->tn = 1;
  } else {
    g->tw = W;
    int th = 128 / W;
    if (th > H) th = H;
    g->th = th;
    g->tn = 128 / (g->tw * g->th);
  }
  return g->tw * g->th * g->tn == 128;
}

static int encode(tfmq_ctx* ctx, CUtensorMap* m, CUtensorMapDataType dt, int rank, const void* base,
                  const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box, const cuuint32_t* estr,
                  CUtensorMapSwizzle sw) {
  CUresult r = ctx->encode_tiled(m, dt, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return tfmq_fail(ctx, TFMQ_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return TFMQ_OK;
}

template <int MODE, int CG>
static int launch_igemm(tfmq_ctx* ctx, const CUtensorMap& tmA, const CUtensorMap& tmA2, const CUtensorMap& tmB,
                        const CUtensorMap& tmB2, IgemmParams& p, cudaStream_t stream, const char* name) {
  // shared-memory plan
  if (CG == 1) p.b_rows[0] = p.tile_n, p.b_row0[0] = 0, p.b_rows[1] = 0, p.b_row0[1] = 0;
  if (p.stat_imgs < 1) p.stat_imgs = 1;
  // epilogue chunk buffers: a third one lets the residual's TMA loads run two chunks ahead of the fold (the K = 2016 layers
  // with a residual are epilogue-bound and spent a quarter of the epilogue waiting for it) -- if the main loop keeps its depth
  static const int eb_env = getenv("TFMQ_IGEMM_EPI_BUFS") ? atoi(getenv("TFMQ_IGEMM_EPI_BUFS")) : 0;
  p.epi_bufs = (eb_env == 2 || eb_env == 3) ? eb_env : (p.res ? 3 : 2);
  const uint32_t extra_base = 1024u /*alignment slack*/ + (uint32_t)p.tile_n * 16u /*chp*/ +
                              (uint32_t)(p.stat_imgs * p.tile_n) * 8u /*GN channel sums*/ + 2u * 256u * 8u /*GN partials*/ +
                              512u /*barriers*/;
  const uint32_t chunk_buf = 128u * 32u * 4u;
  uint32_t extra = extra_base + (uint32_t)p.epi_bufs * chunk_buf;
  uint32_t deep_extra = 0;
  p.deep_bars = 0;
  static const int stages_env = getenv("TFMQ_IGEMM_STAGES") ? atoi(getenv("TFMQ_IGEMM_STAGES")) : 0;   // experiments
  int stages;
  size_t smem;
  p.offA_lo = p.offB_lo = p.offP = p.offB = 0;
  p.u_stages = 0, p.u_bytes = 0, p.p_stages = 0, p.p_bytes = 0;
  if (MODE == MODE_W4A8) {
    // A ring + s8 B ring (CTA pair: each CTA holds half of the N rows of B, incl. the 16 ones / zero rows)
    p.u_bytes = ((uint32_t)(p.tile_n + 16) * 128u / (uint32_t)CG + 1023u) & ~1023u;
    static const int su_env = getenv("TFMQ_IGEMM_USTAGES") ? atoi(getenv("TFMQ_IGEMM_USTAGES")) : 0;
    p.u_stages = su_env > 0 ? su_env : (p.u_bytes * 4u <= 64u * 1024u ? 4 : 2);     // even: two transform groups
    // packed int4 ring: this CTA's weight rows x 64 B per k-block
    p.p_bytes = ((uint32_t)(CG == 2 ? p.b_rows[0] : p.tile_n) * 64u + 1023u) & ~1023u;
    static const int sp_env = getenv("TFMQ_IGEMM_PSTAGES") ? atoi(getenv("TFMQ_IGEMM_PSTAGES")) : 0;
    p.p_stages = sp_env > 0 ? sp_env : 4;
    if (p.u_stages < 2 || p.u_stages > 8 || (p.u_stages & 1) || p.p_stages < 1 || p.p_stages > 8)
      return tfmq_fail(ctx, TFMQ_ERR_ARG, "%s: ring depths (s8 B %d: even, 2..8; packed %d: 1..8)", name, p.u_stages, p.p_stages);
    p.deep_bars = (p.u_stages > 4 || p.p_stages > 4) ? 1 : 0;
    {
      // transform groups: four single-warp groups when a pass of 8 rows x NP pieces covers this CTA's weight rows
      static const int gw_env = getenv("TFMQ_IGEMM_XFGW") ? atoi(getenv("TFMQ_IGEMM_XFGW")) : 0;
      const int rows = CG == 2 ? (p.b_rows[0] > p.b_rows[1] ? p.b_rows[0] : p.b_rows[1]) : p.tile_n;
      const int np = CG == 2 ? 8 : 15;
      p.xf_gw = (rows <= 8 * np && p.u_stages % 4 == 0 && p.p_stages % 4 == 0) ? 1 : 2;
      if (gw_env == 2) p.xf_gw = 2;
    }
    deep_extra = p.deep_bars ? 512u : 0u;             // second block of 64 barrier slots
    extra += deep_extra;
    const long long left = (long long)ctx->max_smem_optin - extra - (long long)p.u_stages * p.u_bytes -
                           (long long)p.p_stages * p.p_bytes;
    stages = (int)(left / (long long)IGEMM_A_BYTES);
    if (p.epi_bufs == 3 && stages < 4 && !eb_env) {      // keep the A ring at least 4 deep (or as deep as it was)
      p.epi_bufs = 2;
      extra = extra_base + deep_extra + 2u * chunk_buf;
      stages = (int)((left + (long long)chunk_buf) / (long long)IGEMM_A_BYTES);
    }
    if (stages > 8) stages = 8;
    if (stages_env > 0 && stages_env < stages) stages = stages_env;
    if (stages < 2) return tfmq_fail(ctx, TFMQ_ERR_SHAPE, "%s: tile does not fit shared memory", name);
    p.stage_bytes = IGEMM_A_BYTES;
    smem = (size_t)stages * IGEMM_A_BYTES + (size_t)p.u_stages * p.u_bytes + (size_t)p.p_stages * p.p_bytes + extra;
  } else {
    const uint32_t bB = (uint32_t)(CG == 2 ? p.b_rows[0] : p.tile_n) * 128u;   // CTA pair: half of the weight rows per CTA
    uint32_t off = IGEMM_A_BYTES;
    constexpr bool FPM = (MODE == MODE_TF32 || MODE == MODE_F16);
    if (FPM && (p.pass_flags & PASS_LO_HI)) {
      p.offA_lo = off;
      off += IGEMM_A_BYTES;
    }
    p.offB = off;
    off += bB;
    if (FPM && (p.pass_flags & PASS_HI_LO)) {
      p.offB_lo = off;
      off += bB;
    }
    p.stage_bytes = (off + 1023u) & ~1023u;
    stages = (int)(((uint32_t)ctx->max_smem_optin - extra) / p.stage_bytes);
    if (p.epi_bufs == 3 && !eb_env) {
      const int stages2 = (int)(((uint32_t)ctx->max_smem_optin - extra + chunk_buf) / p.stage_bytes);
      if (stages2 > stages && stages < 4) {              // the third buffer would cost a pipeline stage
        p.epi_bufs = 2;
        extra -= chunk_buf;
        stages = stages2;
      }
    }
    if (stages > 8) stages = 8;
    if (stages_env > 0 && stages_env < stages) stages = stages_env;
    if (stages < 1) return tfmq_fail(ctx, TFMQ_ERR_SHAPE, "%s: tile does not fit shared memory", name);
    smem = (size_t)stages * p.stage_bytes + extra;
  }
  p.stages = stages;
  const int nkb = p.ksize * p.ksize * ((p.cin + p.kchunk - 1) / p.kchunk);
  int acc_cols = p.tile_n + (MODE == MODE_W4A8 ? 16 : 0);
  if ((MODE == MODE_TF32 || MODE == MODE_F16) && (p.pass_flags & (PASS_LO_HI | PASS_HI_LO))) acc_cols *= 2;
  p.acc_stages = (2 * acc_cols <= 512) ? 2 : 1;
  const int need = acc_cols * p.acc_stages;
  p.tmem_cols = need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512;

  auto kern = igemm_kernel<MODE, CG>;
  static size_t smem_set = 0;  // per template instance; one device per process
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return tfmq_fail(ctx, TFMQ_ERR_CUDA, "%s: smem attr: %s", name, cudaGetErrorString(e));
    smem_set = smem;
  }
  // epilogue tensor maps: output (and residual) as [cout][W][H][N] fp32 / s32 with pixel pitch ld
  p.chunk_w = (p.tile_n % 32 == 0) ? 32 : 16;
  CUtensorMap tmOut, tmRes;
  if (p.out_planes) {
    // fp16 hi / lo planes: tmOut = hi, tmRes = lo, [cout][W][H][N] halves with pixel pitch out_h_ld, 64-byte chunk rows
    if (p.chunk_w != 32) return tfmq_fail(ctx, TFMQ_ERR_SHAPE, "%s: plane output needs an N tile in multiples of 32", name);
    cuuint64_t dims[4] = {(cuuint64_t)p.cout, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.n_img};
    const cuuint64_t ld = (cuuint64_t)p.out_h_ld;
    cuuint64_t str[3] = {ld * 2, (cuuint64_t)p.W * ld * 2, (cuuint64_t)p.H * p.W * ld * 2};
    cuuint32_t box[4] = {32, (cuuint32_t)p.tw, (cuuint32_t)p.th, (cuuint32_t)p.tn};
    cuuint32_t es[4] = {1, 1, 1, 1};
    int rc = encode(ctx, &tmOut, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, p.out_hi, dims, str, box, es, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
    rc = encode(ctx, &tmRes, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, p.out_lo, dims, str, box, es, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
  } else {
    const bool i32 = (MODE == MODE_I8);
    const void* base = i32 ? (const void*)p.out_i32 : (const void*)p.out;
    const cuuint64_t ld = i32 ? (cuuint64_t)p.cout : (cuuint64_t)p.out_ld;
    cuuint64_t dims[4] = {(cuuint64_t)p.cout, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.n_img};
    cuuint64_t str[3] = {ld * 4, (cuuint64_t)p.W * ld * 4, (cuuint64_t)p.H * p.W * ld * 4};
    cuuint32_t box[4] = {(cuuint32_t)p.chunk_w, (cuuint32_t)p.tw, (cuuint32_t)p.th, (cuuint32_t)p.tn};
    cuuint32_t es[4] = {1, 1, 1, 1};
    const CUtensorMapSwizzle sw = p.chunk_w == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    int rc = encode(ctx, &tmOut, i32 ? CU_TENSOR_MAP_DATA_TYPE_INT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims,
                    str, box, es, sw);
    if (rc) return rc;
    tmRes = tmOut;
    if (p.res) {
      cuuint64_t rstr[3] = {(cuuint64_t)p.res_ld * 4, (cuuint64_t)p.W * p.res_ld * 4,
                            (cuuint64_t)p.H * p.W * p.res_ld * 4};
      rc = encode(ctx, &tmRes, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, p.res, dims, rstr, box, es, sw);
      if (rc) return rc;
    }
  }
  const int tiles_m = (p.W / p.tw) * (p.H / p.th) * ((p.n_img + p.tn - 1) / p.tn);
  if (MODE != MODE_F16 || p.ksplit < 1) p.ksplit = 1;
  if (p.ksplit > nkb) p.ksplit = nkb;
  if (p.ksplit > 1) {
    // split-K partial tiles are added into the output by TMA reduce: it starts from zero
    if (p.res || p.emb || p.n_stat || p.out_planes)
      return tfmq_fail(ctx, TFMQ_ERR_ARG, "%s: split-K takes no residual / embedding / statistics / plane output", name);
    cudaError_t e = cudaMemset2DAsync(p.out, (size_t)p.out_ld * 4, 0, (size_t)p.cout * 4, (size_t)p.n_img * p.H * p.W, stream);
    if (e != cudaSuccess) return tfmq_fail(ctx, TFMQ_ERR_CUDA, "%s: clearing the split-K output: %s", name, cudaGetErrorString(e));
  }
  const int total = tiles_m / CG * (p.cout / p.tile_n) * p.ksplit;   // tiles, or pairs of M-adjacent tiles (x K ranges)
  const int units = ctx->sm_count / CG;
  const int grid = (total < units ? total : units) * CG;
  static const bool time_env = getenv("TFMQ_IGEMM_TIME") != nullptr;   // debug aid: per-launch time, synchronous
  static const bool prof_env = getenv("TFMQ_IGEMM_PROF") != nullptr;   // debug aid: + in-kernel phase counters
  static const int dbg_env = getenv("TFMQ_IGEMM_DBG") ? atoi(getenv("TFMQ_IGEMM_DBG")) : 0;
  p.dbg = dbg_env;
  p.prof = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (prof_env) {
    cudaMalloc(&p.prof, (size_t)grid * 16 * sizeof(long long));
    cudaMemsetAsync(p.prof, 0, (size_t)grid * 16 * sizeof(long long), stream);
  }
  if (prof_env || time_env) {
    cudaEventCreate(&ev0);
    cudaEventCreate(&ev1);
    cudaEventRecord(ev0, stream);
  }
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid), cfg.blockDim = dim3(IGEMM_THREADS), cfg.dynamicSmemBytes = smem, cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (CG == 2) {
      attr[na].id = cudaLaunchAttributeClusterDimension;
      attr[na].val.clusterDim.x = 2, attr[na].val.clusterDim.y = 1, attr[na].val.clusterDim.z = 1;
      ++na;
    }
    if (!(prof_env || time_env)) na += tfmq_pdl_attr(&attr[na]);
    cfg.attrs = attr, cfg.numAttrs = (unsigned)na;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmA2, tmB, tmB2, tmOut, tmRes, p);
    if (e != cudaSuccess) return tfmq_fail(ctx, TFMQ_ERR_CUDA, "%s: launch: %s", name, cudaGetErrorString(e));
  }
  TFMQ_LAUNCH_CHECK(name);
  if (prof_env || time_env) {
    cudaEventRecord(ev1, stream);
    cudaStreamSynchronize(stream);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev0, ev1);
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    fprintf(stderr, "[igemm layer %s] n %d %dx%d cin %d cout %d k%d s%d passes 0x%x res %d emb %d stat %d : %.1f us\n", name,
            p.n_img, p.H, p.W, p.cin, p.cout, p.ksize, p.stride, p.pass_flags, p.res != nullptr, p.emb != nullptr,
            p.n_stat, ms * 1e3f);
  }
  if (prof_env) {
    std::vector<long long> hbuf((size_t)grid * 16);
    cudaMemcpy(hbuf.data(), p.prof, hbuf.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(p.prof);
    double avg[16] = {0};
    for (int b = 0; b < grid; ++b)
      for (int i = 0; i < 16; ++i) avg[i] += (double)hbuf[(size_t)b * 16 + i] / grid;
    const double tpc = (double)total / grid;
    fprintf(stderr,
            "[igemm prof %s] tiles/cta %.2f nkb %d tile_n %d | epi per tile: prologue %.0f wait_acc %.0f tmem_ld %.0f "
            "wait_res %.0f fold %.0f fence+bar %.0f stats %.0f store %.0f bar2 %.0f | tail %.0f | mma per tile: "
            "wait_acc_empty %.0f wait_full %.0f issue %.0f\n",
            name, tpc, nkb, p.tile_n, avg[0] / tpc, avg[1] / tpc, avg[2] / tpc, avg[3] / tpc, avg[4] / tpc, avg[5] / tpc,
            avg[6] / tpc, avg[7] / tpc, avg[8] / tpc, avg[9], avg[10] / tpc, avg[11] / tpc, avg[12] / tpc);
  }
  return TFMQ_OK;
}

}  // namespace tfmq

using namespace tfmq;

extern "C" int tfmq_conv_w4a8(tfmq_ctx* ctx, const tfmq_conv_w4a8_desc* d, void* stream) {
  if (!ctx) return TFMQ_ERR_ARG;
  TFMQ_REQUIRE(d && d->act && d->packed && d->wzp && d->wdelta && d->wsum && d->aq && d->out, TFMQ_ERR_ARG,
               "conv_w4a8: null pointer");
  TFMQ_REQUIRE(d->ksize == 1 || d->ksize == 3, TFMQ_ERR_SHAPE, "conv_w4a8: ksize %d", d->ksize);
  TFMQ_REQUIRE(d->cin % 32 == 0 && d->cin >= 32, TFMQ_ERR_SHAPE, "conv_w4a8: cin %d not a multiple of 32", d->cin);
  TFMQ_REQUIRE(d->cout % 16 == 0, TFMQ_ERR_SHAPE, "conv_w4a8: cout %d not a multiple of 16", d->cout);
  TFMQ_REQUIRE(d->out_ld % 4 == 0 && (!d->res || d->res_ld % 4 == 0), TFMQ_ERR_SHAPE, "conv_w4a8: ld not multiple of 4");
  TFMQ_REQUIRE(((uintptr_t)d->out & 15) == 0 && ((uintptr_t)d->act & 15) == 0 && ((uintptr_t)d->packed & 15) == 0 &&
                   (!d->res || ((uintptr_t)d->res & 15) == 0),
               TFMQ_ERR_ARG, "conv_w4a8: pointers must be 16-byte aligned");
  TileGeom g;
  TFMQ_REQUIRE(pick_geom(d->h, d->w, &g), TFMQ_ERR_SHAPE, "conv_w4a8: unsupported spatial %dx%d", d->h, d->w);
  IgemmParams p{};
  p.n_img = d->n, p.H = d->h, p.W = d->w, p.cin = d->cin, p.cout = d->cout;
  p.ksize = d->ksize, p.stride = 1, p.off = 0;
  p.th = g.th, p.tw = g.tw, p.tn = g.tn;
  // CTA pairs (cta_group::2) whenever the M tiles pair up; TFMQ_IGEMM_CG=1 forces single CTAs (debug / comparison)
  static const int cg_env = getenv("TFMQ_IGEMM_CG") ? atoi(getenv("TFMQ_IGEMM_CG")) : 2;
  const int tiles_m_all = (d->w / g.tw) * (d->h / g.th) * ((d->n + g.tn - 1) / g.tn);
  int cg = (cg_env == 2 && tiles_m_all % 2 == 0 && ctx->sm_count % 2 == 0) ? 2 : 1;
  {
    // + 16 rows for the activation-sum column, UMMA N <= 256
    const int nkb = d->ksize * d->ksize * ((d->cin + 127) / 128);
    p.tile_n = pick_tile_n_balanced(d->cout, 240, tiles_m_all / cg, nkb, ctx->sm_count / cg, 400.0, 2.6);
    if (cg == 2 && p.tile_n < 32) {       // the second CTA's half would hold no weight rows
      cg = 1;
      p.tile_n = pick_tile_n_balanced(d->cout, 240, tiles_m_all, nkb, ctx->sm_count, 400.0, 2.6);
    }
  }
  if (cg == 2) {
    const int half = (p.tile_n + 16) / 2;                   // B rows per CTA, a multiple of 8
    p.b_rows[0] = half, p.b_row0[0] = 0;
    p.b_rows[1] = p.tile_n - half, p.b_row0[1] = half;     // + the 16 ones / zero rows
  }
  p.kchunk = 128, p.kslice = 32;
  p.out = d->out, p.out_ld = d->out_ld, p.bias = d->bias, p.wscale = d->wdelta, p.wsum = d->wsum, p.wzp = d->wzp;
  p.aq = d->aq, p.emb = d->emb, p.emb_ld = d->emb_ld, p.res = d->res, p.res_ld = d->res_ld;
  TFMQ_REQUIRE(d->n_stat >= 0 && d->n_stat <= 2, TFMQ_ERR_ARG, "conv_w4a8: n_stat");
  // GroupNorm statistics in the epilogue; a tile of a small feature map spans tn images (per-image sums, <= 8 KB of smem)
  const bool fuse_stats = d->n_stat > 0 && g.tn * p.tile_n * 8 <= 8192;
  if (fuse_stats) {
    p.n_stat = d->n_stat;
    p.stat_imgs = g.tn;
    for (int i = 0; i < d->n_stat; ++i) p.stat[i] = d->stat[i];
  }

  const int halo = d->ksize == 3 ? 1 : 0;
  const cuuint64_t Hp = d->h + 2 * halo, Wp = d->w + 2 * halo;
  CUtensorMap tmA;
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->cin, Wp, Hp, (cuuint64_t)d->n};
    cuuint64_t str[3] = {(cuuint64_t)d->cin, Wp * d->cin, Hp * Wp * d->cin};
    cuuint32_t box[4] = {128, (cuuint32_t)g.tw, (cuuint32_t)g.th, (cuuint32_t)g.tn};
    cuuint32_t es[4] = {1, 1, 1, 1};
    int rc = encode(ctx, &tmA, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, d->act, dims, str, box, es,
                    CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  CUtensorMap tmB, tmB2;
  {
    // packed int4 weights [cout][ksize^2 * cin / 2 bytes]; box = this CTA's weight rows x 64 B (one 128-channel k-block)
    const cuuint64_t kbytes = (cuuint64_t)d->ksize * d->ksize * d->cin / 2;
    cuuint64_t dims[2] = {kbytes, (cuuint64_t)d->cout};
    cuuint64_t str[1] = {kbytes};
    cuuint32_t es[2] = {1, 1};
    for (int r = 0; r < cg; ++r) {
      cuuint32_t box[2] = {64, (cuuint32_t)(cg == 2 ? p.b_rows[r] : p.tile_n)};
      int rc = encode(ctx, r ? &tmB2 : &tmB, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d->packed, dims, str, box, es,
                      CU_TENSOR_MAP_SWIZZLE_NONE);
      if (rc) return rc;
    }
    if (cg == 1) tmB2 = tmB;
  }
  int rc = cg == 2 ? launch_igemm<MODE_W4A8, 2>(ctx, tmA, tmA, tmB, tmB2, p, tfmq_stream(stream), "conv_w4a8")
                   : launch_igemm<MODE_W4A8, 1>(ctx, tmA, tmA, tmB, tmB2, p, tfmq_stream(stream), "conv_w4a8");
  for (int i = 0; rc == TFMQ_OK && !fuse_stats && i < d->n_stat; ++i)
    rc = tfmq_gn_stats_part(ctx, d->out, d->out_ld, d->n, d->h * d->w, d->cout, &d->stat[i], stream);
  return rc;
}

extern "C" int tfmq_conv_fp(tfmq_ctx* ctx, const tfmq_conv_fp_desc* d, void* stream) {
  if (!ctx) return TFMQ_ERR_ARG;
  TFMQ_REQUIRE(d && d->x && d->w_hi && d->out, TFMQ_ERR_ARG, "conv_fp: null pointer");
  TFMQ_REQUIRE(d->ksize == 1 || d->ksize == 3, TFMQ_ERR_SHAPE, "conv_fp: ksize %d", d->ksize);
  TFMQ_REQUIRE(d->stride == 1 || d->stride == 2, TFMQ_ERR_SHAPE, "conv_fp: stride %d", d->stride);
  TFMQ_REQUIRE(d->cin % 8 == 0 && d->cin >= 8, TFMQ_ERR_SHAPE, "conv_fp: cin %d not a multiple of 8", d->cin);
  TFMQ_REQUIRE(d->cout % 16 == 0, TFMQ_ERR_SHAPE, "conv_fp: cout %d not a multiple of 16", d->cout);
  TFMQ_REQUIRE(d->x_ld % 4 == 0 && d->out_ld % 4 == 0 && (!d->res || d->res_ld % 4 == 0), TFMQ_ERR_SHAPE,
               "conv_fp: ld not multiple of 4");
  TFMQ_REQUIRE(((uintptr_t)d->out & 15) == 0 && ((uintptr_t)d->x & 15) == 0 && ((uintptr_t)d->w_hi & 15) == 0 &&
                   (!d->w_lo || ((uintptr_t)d->w_lo & 15) == 0) && (!d->res || ((uintptr_t)d->res & 15) == 0),
               TFMQ_ERR_ARG, "conv_fp: pointers must be 16-byte aligned");
  TFMQ_REQUIRE(d->passes == 1 || d->passes == 3, TFMQ_ERR_ARG, "conv_fp: passes %d", d->passes);
  TileGeom g;
  TFMQ_REQUIRE(pick_geom(d->out_h, d->out_w, &g), TFMQ_ERR_SHAPE, "conv_fp: unsupported spatial %dx%d", d->out_h,
               d->out_w);
  TFMQ_REQUIRE(g.tw * d->stride <= 256 && g.th * d->stride <= 256, TFMQ_ERR_SHAPE, "conv_fp: box too large");
  IgemmParams p{};
  p.n_img = d->n, p.H = d->out_h, p.W = d->out_w, p.cin = d->cin, p.cout = d->cout;
  p.ksize = d->ksize, p.stride = d->stride, p.off = d->ksize == 3 ? -d->pad_lo : 0;
  p.th = g.th, p.tw = g.tw, p.tn = g.tn;
  // two accumulator stages (each hi + lo) need tile_n <= 128; long main loops (3x3 convs) amortise an
  // un-overlapped epilogue better than they tolerate re-reading and re-splitting A once per N tile
  const int nkb_est = d->ksize * d->ksize * ((d->cin + 31) / 32);
  {
    const int tiles_m = (d->out_w / g.tw) * (d->out_h / g.th) * ((d->n + g.tn - 1) / g.tn);
    const int limit = (d->passes == 3 && nkb_est < 48) ? 128 : 256;
    p.tile_n = pick_tile_n_balanced(d->cout, limit, tiles_m, nkb_est, ctx->sm_count, 300.0, 2.0 * d->passes);
  }
  p.kchunk = 32, p.kslice = 8;
  p.pass_flags = PASS_HI_HI;
  if (d->passes == 3) p.pass_flags |= PASS_LO_HI | (d->w_lo ? PASS_HI_LO : 0);
  p.out = d->out, p.out_ld = d->out_ld, p.bias = d->bias, p.wscale = d->wscale;
  p.res = d->res, p.res_ld = d->res_ld;
  p.emb = d->emb, p.emb_ld = d->emb_ld;
  TFMQ_REQUIRE(d->n_stat >= 0 && d->n_stat <= 2, TFMQ_ERR_ARG, "conv_fp: n_stat");
  // GroupNorm statistics in the epilogue; a tile of a small feature map spans tn images (per-image sums, <= 8 KB of smem)
  const bool fuse_stats = d->n_stat > 0 && g.tn * p.tile_n * 8 <= 8192;
  if (fuse_stats) {
    p.n_stat = d->n_stat;
    p.stat_imgs = g.tn;
    for (int i = 0; i < d->n_stat; ++i) p.stat[i] = d->stat[i];
  }

  CUtensorMap tmA, tmB, tmB2;
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->cin, (cuuint64_t)d->w, (cuuint64_t)d->h, (cuuint64_t)d->n};
    cuuint64_t str[3] = {(cuuint64_t)d->x_ld * 4, (cuuint64_t)d->w * d->x_ld * 4,
                         (cuuint64_t)d->h * d->w * d->x_ld * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)(g.tw * d->stride), (cuuint32_t)(g.th * d->stride), (cuuint32_t)g.tn};
    cuuint32_t es[4] = {1, (cuuint32_t)d->stride, (cuuint32_t)d->stride, 1};
    // a box dimension of extent 1 must not carry a traversal stride
    if (g.tw == 1) box[1] = 1, es[1] = 1;
    if (g.th == 1) box[2] = 1, es[2] = 1;
    int rc = encode(ctx, &tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d->x, dims, str, box, es,
                    CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  const cuuint64_t kk = (cuuint64_t)d->ksize * d->ksize * d->cin;
  for (int i = 0; i < 2; ++i) {
    const float* w = i ? d->w_lo : d->w_hi;
    if (!w) {
      tmB2 = tmB;
      continue;
    }
    cuuint64_t dims[2] = {kk, (cuuint64_t)d->cout};
    cuuint64_t str[1] = {kk * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)p.tile_n};
    cuuint32_t es[2] = {1, 1};
    int rc = encode(ctx, i ? &tmB2 : &tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, w, dims, str, box, es,
                    CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  int rc = launch_igemm<MODE_TF32, 1>(ctx, tmA, tmA, tmB, tmB2, p, tfmq_stream(stream), "conv_fp");
  for (int i = 0; rc == TFMQ_OK && !fuse_stats && i < d->n_stat; ++i)
    rc = tfmq_gn_stats_part(ctx, d->out, d->out_ld, d->n, d->out_h * d->out_w, d->cout, &d->stat[i], stream);
  return rc;
}

extern "C" int tfmq_conv_h16(tfmq_ctx* ctx, const tfmq_conv_h16_desc* d, void* stream) {
  if (!ctx) return TFMQ_ERR_ARG;
  TFMQ_REQUIRE(d && d->x_hi && d->x_lo && d->w_hi && (d->out || d->out_hi), TFMQ_ERR_ARG, "conv_h16: null pointer");
  const bool planes = d->out_hi != nullptr;
  TFMQ_REQUIRE(!planes || (d->out_lo && !d->res && !d->emb && d->n_stat == 0 && d->out_h_ld % 8 == 0 && d->cout % 32 == 0 &&
                           (((uintptr_t)d->out_hi | (uintptr_t)d->out_lo) & 15) == 0),
               TFMQ_ERR_ARG, "conv_h16: plane output needs out_lo, 16-byte alignment, out_h_ld %% 8 == 0, cout %% 32 == 0 and "
                             "no residual / embedding / statistics");
  TFMQ_REQUIRE(d->ksize == 1 || d->ksize == 3, TFMQ_ERR_SHAPE, "conv_h16: ksize %d", d->ksize);
  TFMQ_REQUIRE(d->stride == 1 || d->stride == 2, TFMQ_ERR_SHAPE, "conv_h16: stride %d", d->stride);
  TFMQ_REQUIRE(d->cin % 16 == 0 && d->cin >= 16, TFMQ_ERR_SHAPE, "conv_h16: cin %d not a multiple of 16", d->cin);
  TFMQ_REQUIRE(d->cout % 16 == 0, TFMQ_ERR_SHAPE, "conv_h16: cout %d not a multiple of 16", d->cout);
  TFMQ_REQUIRE(d->x_ld % 8 == 0 && (planes || d->out_ld % 4 == 0) && (!d->res || d->res_ld % 4 == 0), TFMQ_ERR_SHAPE,
               "conv_h16: x_ld must be a multiple of 8 halves, out_ld / res_ld of 4 floats");
  TFMQ_REQUIRE((planes || ((uintptr_t)d->out & 15) == 0) && ((uintptr_t)d->x_hi & 15) == 0 && ((uintptr_t)d->x_lo & 15) == 0 &&
                   ((uintptr_t)d->w_hi & 15) == 0 && (!d->w_lo || ((uintptr_t)d->w_lo & 15) == 0) &&
                   (!d->res || ((uintptr_t)d->res & 15) == 0),
               TFMQ_ERR_ARG, "conv_h16: pointers must be 16-byte aligned");
  TileGeom g;
  TFMQ_REQUIRE(pick_geom(d->out_h, d->out_w, &g), TFMQ_ERR_SHAPE, "conv_h16: unsupported spatial %dx%d", d->out_h,
               d->out_w);
  TFMQ_REQUIRE(g.tw * d->stride <= 256 && g.th * d->stride <= 256, TFMQ_ERR_SHAPE, "conv_h16: box too large");
  IgemmParams p{};
  p.n_img = d->n, p.H = d->out_h, p.W = d->out_w, p.cin = d->cin, p.cout = d->cout;
  p.ksize = d->ksize, p.stride = d->stride, p.off = d->ksize == 3 ? -d->pad_lo : 0;
  p.th = g.th, p.tw = g.tw, p.tn = g.tn;
  const int nkb_est = d->ksize * d->ksize * ((d->cin + 63) / 64);
  // CTA pairs (cta_group::2, M = 256) whenever the M tiles pair up and split-K is off; TFMQ_IGEMM_F16_CG=1 forces single CTAs
  static const int f16_cg_env = getenv("TFMQ_IGEMM_F16_CG") ? atoi(getenv("TFMQ_IGEMM_F16_CG")) : 2;
  int cg = 1;
  {
    // two accumulator stages (each main + small-terms) need tile_n <= 128
    const int tiles_m = (d->out_w / g.tw) * (d->out_h / g.th) * ((d->n + g.tn - 1) / g.tn);
    const int limit = nkb_est < 24 ? 128 : 256;
    // measured per layer (profiles/r2l_conv_layers.md): pairs win where the main loop dominates (3x3 convs, K >= 896, the wide
    // qkv projections: 74 -> 67 us) and lose 5-15 % on the epilogue-bound 1x1 convs with K <= 672 and N <= 672
    static const int f16_cg_kb = getenv("TFMQ_IGEMM_F16_CG_KB") ? atoi(getenv("TFMQ_IGEMM_F16_CG_KB")) : 14;
    if (f16_cg_env == 2 && tiles_m % 2 == 0 && ctx->sm_count % 2 == 0 && (d->ksplit == 0 || d->ksplit == 1) &&
        (nkb_est >= f16_cg_kb || d->cout >= 1344))
      cg = 2;
    p.tile_n = pick_tile_n_balanced(d->cout, limit, tiles_m / cg, nkb_est, ctx->sm_count / cg, 300.0, 6.0, planes ? 32 : 16);
    if (d->ksplit != 0 && d->ksplit != 1) {
      // split-K: the widest N tile (fewest re-reads of the pixel operand), and as many K ranges as it takes to give every
      // SM a unit (ksplit < 0: chosen here), each at least 8 k-blocks long
      TFMQ_REQUIRE(!planes && !d->res && !d->emb && d->n_stat == 0, TFMQ_ERR_ARG,
                   "conv_h16: split-K takes no residual / embedding / statistics / plane output");
      int tn = 0;
      for (int t = 256; t >= 32 && !tn; t -= 32)
        if (d->cout % t == 0) tn = t;
      if (tn) p.tile_n = tn;
      const int units_mn = tiles_m * (d->cout / p.tile_n);
      int ks = d->ksplit;
      if (ks < 0) {
        ks = ctx->sm_count / units_mn;
        if (ks > nkb_est / 8) ks = nkb_est / 8;
      }
      p.ksplit = ks < 1 ? 1 : ks;
    }
  }
  p.out_planes = planes ? 1 : 0;
  p.out_hi = d->out_hi, p.out_lo = d->out_lo, p.out_h_ld = d->out_h_ld;
  p.kchunk = 64, p.kslice = 16;
  p.pass_flags = PASS_HI_HI | PASS_LO_HI | (d->w_lo ? PASS_HI_LO : 0);
  p.out = d->out, p.out_ld = d->out_ld, p.bias = d->bias, p.wscale = d->wscale;
  p.res = d->res, p.res_ld = d->res_ld;
  p.emb = d->emb, p.emb_ld = d->emb_ld;
  TFMQ_REQUIRE(d->n_stat >= 0 && d->n_stat <= 2, TFMQ_ERR_ARG, "conv_h16: n_stat");
  // GroupNorm statistics in the epilogue; a tile of a small feature map spans tn images (per-image sums, <= 8 KB of smem)
  const bool fuse_stats = d->n_stat > 0 && g.tn * p.tile_n * 8 <= 8192;
  if (fuse_stats) {
    p.n_stat = d->n_stat;
    p.stat_imgs = g.tn;
    for (int i = 0; i < d->n_stat; ++i) p.stat[i] = d->stat[i];
  }

  CUtensorMap tmA, tmA2, tmB, tmB2;
  for (int i = 0; i < 2; ++i) {
    cuuint64_t dims[4] = {(cuuint64_t)d->cin, (cuuint64_t)d->w, (cuuint64_t)d->h, (cuuint64_t)d->n};
    cuuint64_t str[3] = {(cuuint64_t)d->x_ld * 2, (cuuint64_t)d->w * d->x_ld * 2,
                         (cuuint64_t)d->h * d->w * d->x_ld * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)(g.tw * d->stride), (cuuint32_t)(g.th * d->stride), (cuuint32_t)g.tn};
    cuuint32_t es[4] = {1, (cuuint32_t)d->stride, (cuuint32_t)d->stride, 1};
    // a box dimension of extent 1 must not carry a traversal stride
    if (g.tw == 1) box[1] = 1, es[1] = 1;
    if (g.th == 1) box[2] = 1, es[2] = 1;
    int rc = encode(ctx, i ? &tmA2 : &tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, i ? d->x_lo : d->x_hi, dims, str, box, es,
                    CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  const cuuint64_t kk = (cuuint64_t)d->ksize * d->ksize * d->cin;
  for (int i = 0; i < 2; ++i) {
    const void* w = i ? d->w_lo : d->w_hi;
    if (!w) {
      tmB2 = tmB;
      continue;
    }
    cuuint64_t dims[2] = {kk, (cuuint64_t)d->cout};
    cuuint64_t str[1] = {kk * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)(p.tile_n / cg)};       // CTA pair: each CTA loads half of the weight rows
    cuuint32_t es[2] = {1, 1};
    int rc = encode(ctx, i ? &tmB2 : &tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, w, dims, str, box, es,
                    CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  if (cg == 2) p.b_rows[0] = p.b_rows[1] = p.tile_n / 2, p.b_row0[0] = 0, p.b_row0[1] = p.tile_n / 2;
  int rc = cg == 2 ? launch_igemm<MODE_F16, 2>(ctx, tmA, tmA2, tmB, tmB2, p, tfmq_stream(stream), "conv_h16")
                   : launch_igemm<MODE_F16, 1>(ctx, tmA, tmA2, tmB, tmB2, p, tfmq_stream(stream), "conv_h16");
  for (int i = 0; rc == TFMQ_OK && !fuse_stats && i < d->n_stat; ++i)
    rc = tfmq_gn_stats_part(ctx, d->out, d->out_ld, d->n, d->out_h * d->out_w, d->cout, &d->stat[i], stream);
  return rc;
}

extern "C" int tfmq_gemm_i8_peak(tfmq_ctx* ctx, const uint8_t* a, const int8_t* b, int m, int n, int k, int32_t* out,
                                 void* stream) {
  if (!ctx) return TFMQ_ERR_ARG;
  TFMQ_REQUIRE(a && b && out, TFMQ_ERR_ARG, "gemm_i8_peak: null pointer");
  TFMQ_REQUIRE(m % 128 == 0 && k % 128 == 0 && n % 16 == 0, TFMQ_ERR_SHAPE, "gemm_i8_peak: m%%128, k%%128, n%%16");
  IgemmParams p{};
  p.n_img = m, p.H = 1, p.W = 1, p.cin = k, p.cout = n, p.ksize = 1, p.stride = 1, p.off = 0;
  p.th = 1, p.tw = 1, p.tn = 128;
  p.tile_n = pick_tile_n(n);
  p.kchunk = 128, p.kslice = 32;
  p.out_i32 = out;
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[4] = {(cuuint64_t)k, 1, 1, (cuuint64_t)m};
    cuuint64_t str[3] = {(cuuint64_t)k, (cuuint64_t)k, (cuuint64_t)k};
    cuuint32_t box[4] = {128, 1, 1, 128};
    cuuint32_t es[4] = {1, 1, 1, 1};
    int rc = encode(ctx, &tmA, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, a, dims, str, box, es, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)n};
    cuuint64_t str[1] = {(cuuint64_t)k};
    cuuint32_t box[2] = {128, (cuuint32_t)p.tile_n};
    cuuint32_t es[2] = {1, 1};
    int rc = encode(ctx, &tmB, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, b, dims, str, box, es, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  return launch_igemm<MODE_I8, 1>(ctx, tmA, tmA, tmB, tmB, p, tfmq_stream(stream), "gemm_i8_peak");
}
