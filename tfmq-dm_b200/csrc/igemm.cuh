// Implicit-GEMM convolution / linear on tcgen05 for sm_100a.
//
//   D[128 pixels x tile_n couts] (TMEM) = sum over (tap, cin-chunk) A_tile * B_tile^T
//
// A tiles (activations, NHWC) arrive by 4-D TMA boxes [tn][th][tw][kchunk] so that one
// box is one 128-row, 128-byte-wide, 128B-swizzled K-major UMMA operand.  B tiles
// (weights, [cout][tap*cin]) arrive by 2-D TMA.  A warp-specialised pipeline:
//   warps 0-7   epilogue (TMEM -> registers -> smem -> TMA store), overlapped with the next tile's main loop
//   warp 8      TMA producer
//   warp 9      UMMA issuer, owns TMEM
//   warps 10-13 operand transform between TMA and UMMA (each owns whole pipeline stages)
//     MODE_W4A8 : unpack packed int4 codes -> (q - zp) s8 into the swizzled B tile
//     MODE_TF32 : split fp32 A into tf32 hi + lo planes (3-pass error compensation)
//     MODE_I8   : nothing (dense s8 weights; used for the measured int8 peak)
//     MODE_F16  : nothing: both operands arrive pre-split as fp16 hi + lo planes (kind::f16, 3 products)
// CG = 2 (w4a8 only): a cluster of two CTAs runs tcgen05.mma.cta_group::2 on two M-adjacent tiles (M = 256); each
// CTA stages its own A tile and half of the weight rows.
#pragma once
#include <cstdint>
#include <cuda.h>

#include "../../include/tfmq_b200.h"

namespace tfmq {

enum { MODE_W4A8 = 0, MODE_I8 = 1, MODE_TF32 = 2, MODE_F16 = 3 };
enum { PASS_HI_HI = 1, PASS_LO_HI = 2, PASS_HI_LO = 4 };

struct IgemmParams {
  int n_img, H, W;  // OUTPUT extent
  int cin, cout;
  int ksize, stride, off;  // input coord = out*stride + tap + off
  int th, tw, tn;          // tile = tn*th*tw = 128 output pixels
  int tile_n, stages, tmem_cols, acc_stages, chunk_w;
  int epi_bufs;              // epilogue chunk buffers (2, or 3: residual loads two chunks ahead)
  int deep_bars;             // w4a8: ring barrier arrays of 8 slots in a second barrier block (ring depths > 4)
  int kchunk, kslice;  // channels per k-block / per UMMA
  int xf_gw;                 // w4a8: warps per transform group (2, or 1 for narrow tiles: four groups in flight)
  int u_stages;              // w4a8: slots of the s8 B ring
  uint32_t u_bytes;          // w4a8: bytes per s8 B slot
  int p_stages;              // w4a8: slots of the packed int4 ring
  uint32_t p_bytes;          // w4a8: bytes per packed slot
  int b_rows[2], b_row0[2];  // weight rows of the N tile staged by CTA rank 0 / 1 of a pair (rank 0 only when CG == 1)
  uint32_t stage_bytes, offA_lo, offB, offB_lo, offP;
  int pass_flags;
  int ksplit;                // MODE_F16: the k-blocks of a tile are shared by ksplit units, partial tiles added by TMA reduce (out pre-zeroed)
  // epilogue
  float* out;
  long long out_ld;
  const float* bias;
  const float* wscale;  // [cout] per-channel scale (delta_w) or null
  const int32_t* wsum;
  const int32_t* wzp;
  const float* aq;  // (delta_a, zp_a) or null
  const float* emb;
  long long emb_ld;
  const float* res;
  long long res_ld;
  int32_t* out_i32;  // MODE_I8: raw accumulators [pixels][cout]
  int out_planes;    // fp modes: write fp16 hi / lo planes (out_hi, out_lo, pitch out_h_ld halves) instead of fp32 `out`
  void* out_hi;
  void* out_lo;
  long long out_h_ld;
  int n_stat;        // fused GroupNorm statistics of the output
  int stat_imgs;     // images a tile spans when statistics are fused (tn), else 1
  tfmq_gn_target stat[2];
  int dbg;           // debug, timing experiments only (TFMQ_IGEMM_DBG): 1 = no TMA loads of A, 8 = no int4 unpack, 16 = no TMA store,
                     // 32 = no epilogue fold, 64 = prologue + teardown only (tools/epi_probe.py)
  long long* prof;   // debug: per-CTA phase cycle counters [grid][16] (TFMQ_IGEMM_PROF=1), else null
};

constexpr int IGEMM_EPI_WARPS = 8;   // warps 0..7: two per TMEM lane quarter (each takes half of a chunk's columns)
constexpr int IGEMM_XF_WARPS = 4;    // operand transform
constexpr int IGEMM_XF_GROUP_WARPS = 2;   // w4a8: transform groups of 2 warps take alternate k-blocks
constexpr int IGEMM_WARP_TMA = IGEMM_EPI_WARPS, IGEMM_WARP_MMA = IGEMM_EPI_WARPS + 1, IGEMM_WARP_XF0 = IGEMM_EPI_WARPS + 2;
constexpr int IGEMM_WARP_TMB = IGEMM_WARP_XF0 + IGEMM_XF_WARPS;   // w4a8: TMA producer of the packed weight tiles
constexpr int IGEMM_THREADS = (IGEMM_EPI_WARPS + 3 + IGEMM_XF_WARPS) * 32;
constexpr uint32_t IGEMM_A_BYTES = 128 * 128;

template <int MODE, int CG>
__global__ void igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                             const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmOut,
                             const __grid_constant__ CUtensorMap tmRes, const IgemmParams p);

}  // namespace tfmq
