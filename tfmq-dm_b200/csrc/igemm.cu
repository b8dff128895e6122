// tcgen05 implicit-GEMM convolution kernels + their C-ABI launchers.
// See igemm.cuh for the pipeline description and DESIGN.md for the roofline.
#include "igemm.cuh"

#include "ctx.h"
#include "ptx.cuh"

namespace tfmq {

// ---------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------
// Persistent: grid = min(#tiles, #SMs); each CTA walks tiles t = blockIdx.x, +gridDim.x, ... with the
// M index fastest, so CTAs running together share one weight tile in L2.
//   warps 0-3   epilogue: TMEM -> registers -> swizzled smem chunk (+ TMA-prefetched residual) -> TMA store,
//               overlapped with the next tile's main loop through the second accumulator stage
//   warp 4      TMA producer
//   warp 5      UMMA issuer (owns TMEM: 1 or 2 accumulator stages)
//   warps 6-13  operand transform (int4 unpack / tf32 hi-lo split).  They are the critical path of the
//               main loop, so they get the highest warp ids (the SM arbiter favours high ids) and the
//               non-critical waits (producer, epilogue) back off with nanosleep instead of spinning.
template <int MODE>
__global__ void __launch_bounds__(IGEMM_THREADS, 1)
igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmOut,
             const __grid_constant__ CUtensorMap tmRes, const IgemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // dynamic smem is only guaranteed 16-B aligned: skip to the next 1024-B boundary (128B swizzle atoms).
  // Pointer arithmetic on the __shared__ array (not an integer round trip) keeps the address space
  // known to the compiler, so every access below is LDS/STS rather than a generic LD/ST.
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = p.stages;
  const int ACC = p.acc_stages;

  uint8_t* ebuf = smem + (size_t)S * p.stage_bytes;                       // 2 epilogue chunks [128][chunk_w] f32
  float4* chp = reinterpret_cast<float4*>(ebuf + 2 * 128 * 32 * 4);        // [tile_n] per-channel constants
  uint64_t* bars = reinterpret_cast<uint64_t*>(ebuf + 2 * 128 * 32 * 4 + 256 * 16);
  uint64_t* full_tma = bars;            // [S] TMA bytes landed
  uint64_t* full_xf = bars + S;         // [S] transform warps done
  uint64_t* empty = bars + 2 * S;       // [S] UMMAs that read the stage retired
  uint64_t* acc_full = bars + 3 * S;    // [2] accumulator stage complete
  uint64_t* acc_empty = bars + 3 * S + 2;  // [2] accumulator stage drained by the epilogue
  uint64_t* res_full = bars + 3 * S + 4;   // [2] residual chunk landed in the epilogue buffer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * S + 6);

  const int tiles_x = p.W / p.tw;
  const int tiles_y = p.H / p.th;
  const int tiles_m = tiles_x * tiles_y * ((p.n_img + p.tn - 1) / p.tn);
  const int tiles_n = p.cout / p.tile_n;
  const int total_tiles = tiles_m * tiles_n;
  const int kchunks = (p.cin + p.kchunk - 1) / p.kchunk;
  const int taps = p.ksize * p.ksize;
  const int nkb = taps * kchunks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_tma[s], 1);
      mbar_init(&full_xf[s], IGEMM_XF_WARPS);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 4);
      mbar_init(&res_full[s], 1);
    }
    mbar_fence_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (MODE == MODE_TF32) tma_prefetch_desc(&tmB2);
    tma_prefetch_desc(&tmOut);
    if (p.res) tma_prefetch_desc(&tmRes);
  }
  if (warp == 5) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  if (MODE == MODE_W4A8) {
    // rows [tile_n, tile_n+16) of every B stage: row tile_n = 0x01 bytes, the rest zero (swizzle-invariant)
    for (int i = threadIdx.x; i < S * 16 * 8; i += IGEMM_THREADS) {
      const int st_i = i / 128, rem = i - st_i * 128;
      const uint32_t fill = (rem < 8) ? 0x01010101u : 0u;
      *reinterpret_cast<uint4*>(smem + (size_t)st_i * p.stage_bytes + p.offB + (size_t)p.tile_n * 128 + rem * 16) =
          make_uint4(fill, fill, fill, fill);
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const bool need_a_lo = (MODE == MODE_TF32) && (p.pass_flags & PASS_LO_HI);
  const bool need_b_lo = (MODE == MODE_TF32) && (p.pass_flags & PASS_HI_LO);
  const bool two_acc = need_a_lo || need_b_lo;                 // tf32: separate accumulator for the small terms
  const uint32_t acc_cols = (MODE == MODE_W4A8) ? (uint32_t)p.tile_n + 16u : (uint32_t)p.tile_n * (two_acc ? 2u : 1u);

  if (warp == 4) {
    // ===================================================== TMA producer
    if (lane == 0) {
      uint32_t tx_bytes = IGEMM_A_BYTES;
      if (MODE == MODE_W4A8) tx_bytes += (uint32_t)p.tile_n * 64u;
      if (MODE == MODE_I8) tx_bytes += (uint32_t)p.tile_n * 128u;
      if (MODE == MODE_TF32) tx_bytes += (uint32_t)p.tile_n * 128u * (need_b_lo ? 2u : 1u);
      int s = 0;
      uint32_t par = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int mt = tile % tiles_m;
        const int c_out0 = (tile / tiles_m) * p.tile_n;
        const int tx = mt % tiles_x;
        mt /= tiles_x;
        const int ty = mt % tiles_y;
        const int n0 = (mt / tiles_y) * p.tn, y0 = ty * p.th, x0 = tx * p.tw;
        int kc = 0, kx = 0, ky = 0, tap = 0;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait_relaxed(&empty[s], par ^ 1u);
          uint8_t* st = smem + (size_t)s * p.stage_bytes;
          mbar_expect_tx(&full_tma[s], tx_bytes);
          tma_load_4d(st, &tmA, &full_tma[s], kc * p.kchunk, x0 * p.stride + kx + p.off,
                      y0 * p.stride + ky + p.off, n0);
          if (MODE == MODE_W4A8) {
            // packed bytes: column = (tap*cin + kc*128)/2
            tma_load_2d(st + p.offP, &tmB, &full_tma[s], (tap * p.cin + kc * p.kchunk) >> 1, c_out0);
          } else {
            tma_load_2d(st + p.offB, &tmB, &full_tma[s], tap * p.cin + kc * p.kchunk, c_out0);
            if (need_b_lo) tma_load_2d(st + p.offB_lo, &tmB2, &full_tma[s], tap * p.cin + kc * p.kchunk, c_out0);
          }
          if (++kc == kchunks) {
            kc = 0, ++tap;
            if (++kx == p.ksize) kx = 0, ++ky;
          }
          if (++s == S) s = 0, par ^= 1u;
        }
      }
    }
  } else if (warp == 5) {
    // ===================================================== UMMA issuer
    // W4A8: the s8 B tile carries 16 extra rows; row tile_n is all ones, so accumulator column tile_n
    // holds sum_k a[m][k] and the weight zero point can be applied in the epilogue instead of per code
    const uint32_t umma_n = (uint32_t)p.tile_n + (MODE == MODE_W4A8 ? 16u : 0u);
    const uint32_t idesc = (MODE == MODE_TF32) ? idesc_tf32(128, umma_n) : idesc_i8_u8s8(128, umma_n);
    uint32_t tcount = 0;
    int s = 0;
    uint32_t par = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
      const uint32_t as = tcount % ACC;
      mbar_wait(&acc_empty[as], ((tcount / ACC) & 1u) ^ 1u);   // epilogue has drained this stage
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * acc_cols;
      // tf32: the two small cross terms accumulate in their own TMEM columns so the (truncating)
      // tensor-core accumulator adds them to a small running sum, not to the large main one
      const uint32_t tmem_lo = tmem_d + (uint32_t)p.tile_n;
      uint32_t accumulate = 0, accumulate_lo = 0;
      int kc = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&full_tma[s], par);
        mbar_wait(&full_xf[s], par);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t st = smem_u32(smem + (size_t)s * p.stage_bytes);
          int rem = p.cin - kc * p.kchunk;
          if (rem > p.kchunk) rem = p.kchunk;
          const int nslice = rem / p.kslice;  // valid 32-byte K slices in this k-block
          const uint64_t a_hi = smem_desc_sw128(st);
          const uint64_t b_hi = smem_desc_sw128(st + p.offB);
          if (MODE == MODE_TF32) {
            const uint64_t a_lo = smem_desc_sw128(st + p.offA_lo);
            const uint64_t b_lo = smem_desc_sw128(st + p.offB_lo);
            if (need_a_lo)
              for (int k = 0; k < nslice; ++k) {
                umma_tf32(tmem_lo, a_lo + 2 * k, b_hi + 2 * k, idesc, accumulate_lo);
                accumulate_lo = 1;
              }
            if (need_b_lo)
              for (int k = 0; k < nslice; ++k) {
                umma_tf32(tmem_lo, a_hi + 2 * k, b_lo + 2 * k, idesc, accumulate_lo);
                accumulate_lo = 1;
              }
            for (int k = 0; k < nslice; ++k) {
              umma_tf32(tmem_d, a_hi + 2 * k, b_hi + 2 * k, idesc, accumulate);
              accumulate = 1;
            }
          } else {
            for (int k = 0; k < nslice; ++k) {
              umma_i8(tmem_d, a_hi + 2 * k, b_hi + 2 * k, idesc, accumulate);
              accumulate = 1;
            }
          }
          umma_commit(&empty[s]);
          if (kb == nkb - 1) umma_commit(&acc_full[as]);
        }
        __syncwarp();
        if (++kc == kchunks) kc = 0;
        if (++s == S) s = 0, par ^= 1u;
      }
    }
  } else if (warp >= 6) {
    // ===================================================== transform warps (6..13)
    const int t = threadIdx.x - 192;  // 0..255
    int s = 0;
    uint32_t par = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int c_out0 = (tile / tiles_m) * p.tile_n;
      if (MODE == MODE_W4A8) {
        // 256 threads: piece (row, K slice) = ((t >> 2) + 64 i, t & 3); fixed per-thread offsets,
        // +4096 B (packed) / +8192 B (s8 tile) per i.  Codes stay unsigned (0..15): two ANDs and a shift
        // per packed word; the zero point is folded out through the ones row (see the UMMA issuer).
        const int sub = t & 3;
        const uint32_t rd0 = p.offP + (uint32_t)(t >> 2) * 64u + (uint32_t)sub * 16u;
        const uint32_t swz = (uint32_t)((t >> 2) & 7);
        const uint32_t wr_lo = p.offB + (uint32_t)(t >> 2) * 128u + (((2u * sub) ^ swz) << 4);
        const uint32_t wr_hi = p.offB + (uint32_t)(t >> 2) * 128u + (((2u * sub + 1u) ^ swz) << 4);
        const int nrow_i = (p.tile_n - (t >> 2) + 63) >> 6;   // iterations with row < tile_n
        int kc = 0;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_tma[s], par);
          uint8_t* st = smem + (size_t)s * p.stage_bytes;
          int rem = p.cin - kc * p.kchunk;
          if (rem > p.kchunk) rem = p.kchunk;
          const int nslice = rem >> 5;
          if (sub < nslice) {
            uint4 pkv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (i < nrow_i) pkv[i] = *reinterpret_cast<const uint4*>(st + rd0 + i * 4096);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (i < nrow_i) {
                const uint4 pk = pkv[i];
                uint4 lo, hi;
                lo.x = pk.x & 0x0F0F0F0Fu, lo.y = pk.y & 0x0F0F0F0Fu, lo.z = pk.z & 0x0F0F0F0Fu, lo.w = pk.w & 0x0F0F0F0Fu;
                hi.x = (pk.x >> 4) & 0x0F0F0F0Fu, hi.y = (pk.y >> 4) & 0x0F0F0F0Fu;
                hi.z = (pk.z >> 4) & 0x0F0F0F0Fu, hi.w = (pk.w >> 4) & 0x0F0F0F0Fu;
                *reinterpret_cast<uint4*>(st + wr_lo + i * 8192) = lo;
                *reinterpret_cast<uint4*>(st + wr_hi + i * 8192) = hi;
              }
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&full_xf[s]);
          if (++kc == kchunks) kc = 0;
          if (++s == S) s = 0, par ^= 1u;
        }
      } else if (MODE == MODE_TF32) {
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_tma[s], par);
          if (need_a_lo) {
            uint8_t* st = smem + (size_t)s * p.stage_bytes;
            uint4* a = reinterpret_cast<uint4*>(st);
            uint4* al = reinterpret_cast<uint4*>(st + p.offA_lo);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int idx = t + 256 * i;
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
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&full_xf[s]);
          if (++s == S) s = 0, par ^= 1u;
        }
      } else {
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_tma[s], par);
          __syncwarp();
          if (lane == 0) mbar_arrive(&full_xf[s]);
          if (++s == S) s = 0, par ^= 1u;
        }
      }
    }
  } else {
    // ===================================================== epilogue warps (0..3)
    // All global I/O of the epilogue is TMA: the residual chunk is prefetched into a swizzled smem
    // buffer, every thread folds its accumulator row into it (conflict-free float4 accesses), and the
    // buffer is stored back with one bulk tensor store.  Two buffers alternate across chunks.
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int et = threadIdx.x;             // 0..127 within the epilogue group
    const int r = q * 32 + lane;            // accumulator row = pixel within the tile
    const int CW = p.chunk_w;               // 32 (128B swizzle) or 16 (64B swizzle) channels per chunk
    const int nchunks = p.tile_n / CW;
    const uint32_t chunk_bytes = 128u * (uint32_t)CW * 4u;
    // swizzled position of 16-byte piece k of this thread's row
    const int sw = (CW == 32) ? (r & 7) : ((r >> 1) & 3);
    const int hw_t = p.th * p.tw;
    float a_scale = 1.f;
    int za = 0;
    if (MODE == MODE_W4A8) {
      a_scale = p.aq[0];
      za = (int)p.aq[1];
    }
    uint32_t tcount = 0, g = 0;             // tiles / chunks processed by this CTA
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
      int mt = tile % tiles_m;
      const int c_out0 = (tile / tiles_m) * p.tile_n;
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
          mbar_expect_tx(&res_full[g & 1], chunk_bytes);
          tma_load_4d(ebuf + (g & 1) * chunk_bytes, &tmRes, &res_full[g & 1], c_out0, x0, y0, n0);
          // pull the rest of the residual tile into L2 while the main loop of this tile runs
          for (int ci = 1; ci < nchunks; ++ci) tma_prefetch_l2_4d(&tmRes, c_out0 + ci * CW, x0, y0, n0);
        }
      }
      // per-channel constants of this N tile
      named_bar_sync(1, 128);                 // previous tile is done with chp
      for (int ch = et; ch < p.tile_n; ch += 128) {
        const int c = c_out0 + ch;
        float sc = 1.f, bi = 0.f;
        int ws = 0;
        int zw = 0;
        if (MODE == MODE_W4A8) {
          sc = a_scale * p.wscale[c];
          ws = za * p.wsum[c];
          zw = (int)p.wzp[c];
        } else if (p.wscale) {
          sc = p.wscale[c];
        }
        if (p.bias) bi = p.bias[c];
        chp[ch] = make_float4(sc, bi, __int_as_float(ws), __int_as_float(zw));
      }
      const uint32_t as = tcount % ACC;
      mbar_wait_relaxed(&acc_full[as], (tcount / ACC) & 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * acc_cols + ((uint32_t)(q * 32) << 16);
      named_bar_sync(1, 128);                 // chp visible; first buffer known free
      int a_sum = 0;                          // sum_k a[m][k] of this row (ones-row column)
      if (MODE == MODE_W4A8) {
        uint32_t sv[16];
        tmem_ld16(tmem_d + (uint32_t)p.tile_n, sv);
        tmem_ld_wait();
        a_sum = (int)sv[0];
      }

      for (int ci = 0; ci < nchunks; ++ci, ++g) {
        const int c0 = ci * CW;
        uint8_t* buf = ebuf + (g & 1) * chunk_bytes;
        uint32_t v[32];
        tmem_ld16(tmem_d + (uint32_t)c0, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
        if (CW == 32) tmem_ld16(tmem_d + (uint32_t)(c0 + 16), *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
        if (MODE == MODE_TF32 && two_acc) {
          uint32_t v2[32];
          tmem_ld16(tmem_d + (uint32_t)(p.tile_n + c0), *reinterpret_cast<uint32_t(*)[16]>(&v2[0]));
          if (CW == 32)
            tmem_ld16(tmem_d + (uint32_t)(p.tile_n + c0 + 16), *reinterpret_cast<uint32_t(*)[16]>(&v2[16]));
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
        }
        tmem_ld_wait();
        if (ci == nchunks - 1) {
          // last TMEM read of this tile: hand the accumulator stage back to the UMMA issuer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[as]);
        }
        if (p.res) mbar_wait(&res_full[g & 1], (g >> 1) & 1u);
        const float* embp = p.emb ? p.emb + (long long)n_l * p.emb_ld + c_out0 + c0 : nullptr;
        const int npiece = CW / 4;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (k < npiece) {
            float4* slot = reinterpret_cast<float4*>(buf + r * (CW * 4) + ((k ^ sw) << 4));
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
        fence_proxy_async_smem();
        named_bar_sync(1, 128);               // all 128 rows of the chunk are in smem
        if (et == 0) {
          tma_store_4d(&tmOut, buf, c_out0 + c0, x0, y0, n0);
          tma_store_commit();
          if (ci + 1 < nchunks) {
            tma_store_wait_read<1>();         // the other buffer's store has finished reading
            if (p.res) {
              mbar_expect_tx(&res_full[(g + 1) & 1], chunk_bytes);
              tma_load_4d(ebuf + ((g + 1) & 1) * chunk_bytes, &tmRes, &res_full[(g + 1) & 1], c_out0 + c0 + CW, x0,
                          y0, n0);
            }
          }
        }
        if (!p.res) named_bar_sync(1, 128);   // without a residual barrier, publish "next buffer is free"
      }
    }
    if (et == 0) tma_store_wait_all<0>();     // global writes complete before the CTA retires
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
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

template <int MODE>
static int launch_igemm(tfmq_ctx* ctx, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmB2,
                        IgemmParams& p, cudaStream_t stream, const char* name) {
  // shared-memory plan
  const uint32_t bB = ((uint32_t)p.tile_n + (MODE == MODE_W4A8 ? 16u : 0u)) * 128u;
  uint32_t off = IGEMM_A_BYTES;
  p.offA_lo = p.offB_lo = p.offP = 0;
  if (MODE == MODE_TF32 && (p.pass_flags & PASS_LO_HI)) {
    p.offA_lo = off;
    off += IGEMM_A_BYTES;
  }
  p.offB = off;
  off += bB;
  if (MODE == MODE_TF32 && (p.pass_flags & PASS_HI_LO)) {
    p.offB_lo = off;
    off += bB;
  }
  if (MODE == MODE_W4A8) {
    p.offP = off;
    off += (uint32_t)p.tile_n * 64u;
  }
  p.stage_bytes = (off + 1023u) & ~1023u;
  const uint32_t extra = 1024u /*alignment slack*/ + 2u * 128u * 32u * 4u /*epilogue chunks*/ + 256u * 16u /*chp*/ +
                         256u /*barriers*/;
  const int kchunks = (p.cin + p.kchunk - 1) / p.kchunk;
  const int nkb = p.ksize * p.ksize * kchunks;
  int stages = (int)(((uint32_t)ctx->max_smem_optin - extra) / p.stage_bytes);
  if (stages > 6) stages = 6;
  if (stages < 1) return tfmq_fail(ctx, TFMQ_ERR_SHAPE, "%s: tile does not fit shared memory", name);
  (void)nkb;
  p.stages = stages;
  int acc_cols = p.tile_n + (MODE == MODE_W4A8 ? 16 : 0);
  if (MODE == MODE_TF32 && (p.pass_flags & (PASS_LO_HI | PASS_HI_LO))) acc_cols *= 2;
  p.acc_stages = (2 * acc_cols <= 512) ? 2 : 1;
  const int need = acc_cols * p.acc_stages;
  p.tmem_cols = need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512;
  const size_t smem = (size_t)stages * p.stage_bytes + extra;

  auto kern = igemm_kernel<MODE>;
  static size_t smem_set = 0;  // per template instance; one device per process
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return tfmq_fail(ctx, TFMQ_ERR_CUDA, "%s: smem attr: %s", name, cudaGetErrorString(e));
    smem_set = smem;
  }
  // epilogue tensor maps: output (and residual) as [cout][W][H][N] fp32 / s32 with pixel pitch ld
  p.chunk_w = (p.tile_n % 32 == 0) ? 32 : 16;
  CUtensorMap tmOut, tmRes;
  {
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
  const int total = tiles_m * (p.cout / p.tile_n);
  const int grid = total < ctx->sm_count ? total : ctx->sm_count;
  kern<<<grid, IGEMM_THREADS, smem, stream>>>(tmA, tmB, tmB2, tmOut, tmRes, p);
  TFMQ_LAUNCH_CHECK(name);
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
  p.tile_n = pick_tile_n(d->cout, 240);   // + 16 rows for the activation-sum column, UMMA N <= 256
  p.kchunk = 128, p.kslice = 32;
  p.out = d->out, p.out_ld = d->out_ld, p.bias = d->bias, p.wscale = d->wdelta, p.wsum = d->wsum, p.wzp = d->wzp;
  p.aq = d->aq, p.emb = d->emb, p.emb_ld = d->emb_ld, p.res = d->res, p.res_ld = d->res_ld;

  const int halo = d->ksize == 3 ? 1 : 0;
  const cuuint64_t Hp = d->h + 2 * halo, Wp = d->w + 2 * halo;
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->cin, Wp, Hp, (cuuint64_t)d->n};
    cuuint64_t str[3] = {(cuuint64_t)d->cin, Wp * d->cin, Hp * Wp * d->cin};
    cuuint32_t box[4] = {128, (cuuint32_t)g.tw, (cuuint32_t)g.th, (cuuint32_t)g.tn};
    cuuint32_t es[4] = {1, 1, 1, 1};
    int rc = encode(ctx, &tmA, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, d->act, dims, str, box, es,
                    CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    const cuuint64_t kbytes = (cuuint64_t)d->ksize * d->ksize * d->cin / 2;
    cuuint64_t dims[2] = {kbytes, (cuuint64_t)d->cout};
    cuuint64_t str[1] = {kbytes};
    cuuint32_t box[2] = {64, (cuuint32_t)p.tile_n};
    cuuint32_t es[2] = {1, 1};
    int rc = encode(ctx, &tmB, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d->packed, dims, str, box, es,
                    CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
  }
  return launch_igemm<MODE_W4A8>(ctx, tmA, tmB, tmB, p, tfmq_stream(stream), "conv_w4a8");
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
  p.tile_n = pick_tile_n(d->cout, (d->passes == 3 && nkb_est < 48) ? 128 : 256);
  p.kchunk = 32, p.kslice = 8;
  p.pass_flags = PASS_HI_HI;
  if (d->passes == 3) p.pass_flags |= PASS_LO_HI | (d->w_lo ? PASS_HI_LO : 0);
  p.out = d->out, p.out_ld = d->out_ld, p.bias = d->bias, p.wscale = d->wscale;
  p.res = d->res, p.res_ld = d->res_ld;
  p.emb = d->emb, p.emb_ld = d->emb_ld;

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
  return launch_igemm<MODE_TF32>(ctx, tmA, tmB, tmB2, p, tfmq_stream(stream), "conv_fp");
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
  return launch_igemm<MODE_I8>(ctx, tmA, tmB, tmB, p, tfmq_stream(stream), "gemm_i8_peak");
}
