// Implicit-GEMM convolution / linear on tcgen05 for sm_100a.
//
//   D[128 pixels x tile_n couts] (TMEM) = sum over (tap, cin-chunk) A_tile * B_tile^T
//
// A tiles (activations, NHWC) arrive by 4-D TMA boxes [tn][th][tw][kchunk] so that one
// box is one 128-row, 128-byte-wide, 128B-swizzled K-major UMMA operand.  B tiles
// (weights, [cout][tap*cin]) arrive by 2-D TMA.  A warp-specialised pipeline:
//   warp 0    TMA producer
//   warp 1    UMMA issuer, owns TMEM
//   warps 2-5 operand transform between TMA and UMMA, then the epilogue
//     MODE_W4A8 : unpack packed int4 codes -> (q - zp) s8 into the swizzled B tile
//     MODE_TF32 : split fp32 A into tf32 hi + lo planes (3-pass error compensation)
//     MODE_I8   : nothing (dense s8 weights; used for the measured int8 peak)
#pragma once
#include <cstdint>
#include <cuda.h>

#include "../../include/tfmq_b200.h"

namespace tfmq {

enum { MODE_W4A8 = 0, MODE_I8 = 1, MODE_TF32 = 2 };
enum { PASS_HI_HI = 1, PASS_LO_HI = 2, PASS_HI_LO = 4 };

struct IgemmParams {
  int n_img, H, W;  // OUTPUT extent
  int cin, cout;
  int ksize, stride, off;  // input coord = out*stride + tap + off
  int th, tw, tn;          // tile = tn*th*tw = 128 output pixels
  int tile_n, stages, tmem_cols, acc_stages, chunk_w;
  int kchunk, kslice;  // channels per k-block / per UMMA
  uint32_t stage_bytes, offA_lo, offB, offB_lo, offP;
  int pass_flags;
  // epilogue
  float* out;
  long long out_ld;
  const float* bias;
  const float* wscale;  // [cout] per-channel scale (delta_w) or null
  const int32_t* wsum;
  const uint8_t* wzp;
  const float* aq;  // (delta_a, zp_a) or null
  const float* emb;
  long long emb_ld;
  const float* res;
  long long res_ld;
  int32_t* out_i32;  // MODE_I8: raw accumulators [pixels][cout]
  int n_stat;        // fused GroupNorm statistics of the output (only when tn == 1)
  tfmq_gn_target stat[2];
  int dbg;           // debug: bit0 = producer skips the TMA loads (timing experiments only, TFMQ_IGEMM_DBG)
  long long* prof;   // debug: per-CTA phase cycle counters [grid][16] (TFMQ_IGEMM_PROF=1), else null
};

constexpr int IGEMM_EPI_WARPS = 8;   // warps 0..7: two per TMEM lane quarter (each takes half of a chunk's columns)
constexpr int IGEMM_XF_WARPS = 4;    // last warps: operand transform
constexpr int IGEMM_WARP_TMA = IGEMM_EPI_WARPS, IGEMM_WARP_MMA = IGEMM_EPI_WARPS + 1, IGEMM_WARP_XF0 = IGEMM_EPI_WARPS + 2;
constexpr int IGEMM_THREADS = (IGEMM_EPI_WARPS + 2 + IGEMM_XF_WARPS) * 32;
constexpr uint32_t IGEMM_A_BYTES = 128 * 128;

template <int MODE>
__global__ void igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                             const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmOut,
                             const __grid_constant__ CUtensorMap tmRes, const IgemmParams p);

}  // namespace tfmq
