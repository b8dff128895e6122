// PTQ calibration kernels: range statistics, the 80-candidate L2.4 scale search,
// running-stat activation ranges, AdaRound soft weights and the fused AdaRound step.
#include "ctx.h"

namespace tfmq {

// monotone float <-> int map so integer atomics implement float min / max
__device__ __forceinline__ int f2o(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float o2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int RED_THREADS = 256;

// mm_bits[row] = (ordered min, ordered max); init then atomically refined by chunks
__global__ void minmax_init_kernel(int* mm_bits, long long rows) {
  const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (r < rows) {
    mm_bits[2 * r] = f2o(INFINITY);
    mm_bits[2 * r + 1] = f2o(-INFINITY);
  }
}
// x addressed as row*row_stride + col (cols contiguous)
__global__ void __launch_bounds__(RED_THREADS) minmax_kernel(const float* __restrict__ x, long long row_stride,
                                                             long long cols, long long chunk, int* mm_bits,
                                                             int slot_per_row) {
  const long long row = blockIdx.y;
  const long long slot = slot_per_row ? row : 0;
  const long long c0 = blockIdx.x * chunk;
  long long c1 = c0 + chunk;
  if (c1 > cols) c1 = cols;
  const float* xr = x + row * row_stride;
  float mn = INFINITY, mx = -INFINITY;
  for (long long i = c0 + threadIdx.x; i < c1; i += RED_THREADS) {
    const float v = xr[i];
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
  mn = warp_min(mn);
  mx = warp_max(mx);
  __shared__ float smn[RED_THREADS / 32], smx[RED_THREADS / 32];
  if ((threadIdx.x & 31) == 0) smn[threadIdx.x >> 5] = mn, smx[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < RED_THREADS / 32; ++i) mn = fminf(mn, smn[i]), mx = fmaxf(mx, smx[i]);
    atomicMin(&mm_bits[2 * slot], f2o(mn));
    atomicMax(&mm_bits[2 * slot + 1], f2o(mx));
  }
}
__global__ void minmax_decode_kernel(int* mm_bits, long long rows) {
  const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (r < rows) {
    float* f = reinterpret_cast<float*>(mm_bits);
    f[2 * r] = o2f(mm_bits[2 * r]);
    f[2 * r + 1] = o2f(mm_bits[2 * r + 1]);
  }
}

// candidate i of Scaler.MSE: same double -> float path as the Python reference
__device__ __forceinline__ void mse_candidate(float x_min, float x_max, int i, int level, float* delta, float* zp) {
  const double f = 1.0 - ((double)i * 0.01);
  const double new_min = (double)x_min * f, new_max = (double)x_max * f;
  const float d = (float)((new_max - new_min) / (double)(level - 1));
  *delta = d;
  *zp = rintf(__fdiv_rn((float)(-new_min), d));
}

constexpr int MSE_CAND = 80;
// scores[row][80] (double, zeroed by caller) += sum over this chunk of |dq(x)-x|^2.4
__global__ void __launch_bounds__(RED_THREADS) mse_score_kernel(const float* __restrict__ x, long long cols,
                                                                long long chunk, int level,
                                                                const float* __restrict__ mm, double* scores) {
  const long long row = blockIdx.y;
  __shared__ float sd[MSE_CAND], sz[MSE_CAND];
  __shared__ double part[RED_THREADS / 32];
  if (threadIdx.x < MSE_CAND) mse_candidate(mm[2 * row], mm[2 * row + 1], threadIdx.x, level, &sd[threadIdx.x], &sz[threadIdx.x]);
  __syncthreads();
  const long long c0 = blockIdx.x * chunk;
  long long c1 = c0 + chunk;
  if (c1 > cols) c1 = cols;
  const float* xr = x + row * cols;
  const float top = (float)(level - 1);
  for (int cand = 0; cand < MSE_CAND; ++cand) {
    const float d = sd[cand], z = sz[cand];
    float acc = 0.f;
    for (long long i = c0 + threadIdx.x; i < c1; i += RED_THREADS) {
      const float v = xr[i];
      const float q = fminf(fmaxf(rintf(__fdiv_rn(v, d)) + z, 0.f), top);
      const float e = fabsf(__fmul_rn(d, q - z) - v);
      acc += powf(e, 2.4f);
    }
    double a = warp_sum_d((double)acc);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int i = 1; i < RED_THREADS / 32; ++i) a += part[i];
      atomicAdd(&scores[row * MSE_CAND + cand], a);
    }
    __syncthreads();
  }
}
__global__ void mse_pick_kernel(const double* __restrict__ scores, const float* __restrict__ mm, long long rows,
                                long long cols, int level, float* delta, float* zp) {
  const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (r >= rows) return;
  // the reference compares fp32 means with strict '<' starting from 1e10
  float best = 1e10f;
  int bi = 0;
  for (int i = 0; i < MSE_CAND; ++i) {
    const float s = (float)(scores[r * MSE_CAND + i] / (double)cols);
    if (s < best) best = s, bi = i;
  }
  mse_candidate(mm[2 * r], mm[2 * r + 1], bi, level, &delta[r], &zp[r]);
}

// act_momentum_update finalisation (single thread): EMA, then Scaler.MINMAX on the EMA range
__global__ void act_range_finalize_kernel(float* state, float* aq, float momentum, int level) {
  int* bits = reinterpret_cast<int*>(state);
  const float bmin = o2f(bits[2]), bmax = o2f(bits[3]);
  const float one_minus = (float)(1.0 - (double)momentum);
  const float xmin = __fadd_rn(__fmul_rn(state[0], momentum), __fmul_rn(bmin, one_minus));
  const float xmax = __fadd_rn(__fmul_rn(state[1], momentum), __fmul_rn(bmax, one_minus));
  state[0] = xmin;
  state[1] = xmax;
  bits[2] = f2o(INFINITY);
  bits[3] = f2o(-INFINITY);
  const double lo = fmin((double)xmin, 0.0), hi = fmax((double)xmax, 0.0);
  float d = (float)((hi - lo) / (double)(level - 1));
  if (d < 1e-8f) d = 1e-8f;
  aq[0] = d;
  aq[1] = rintf(__fdiv_rn((float)(-lo), d));
}

// ---------------------------------------------------------------- AdaRound
__device__ __forceinline__ float sigmoid_f(float a) { return 1.f / (1.f + expf(-a)); }

__global__ void adaround_soft_kernel(const float* __restrict__ w, const float* __restrict__ delta,
                                     const float* __restrict__ zp, const float* __restrict__ alpha, long long k,
                                     long long total, int level, float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / k;
    const float d = delta[row], z = zp[row];
    const float h = fminf(fmaxf(__fadd_rn(__fmul_rn(sigmoid_f(alpha[i]), 1.2f), -0.1f), 0.f), 1.f);
    const float q = fminf(fmaxf(floorf(__fdiv_rn(w[i], d)) + h + z, 0.f), (float)(level - 1));
    out[i] = __fmul_rn(d, q - z);
  }
}

__global__ void __launch_bounds__(RED_THREADS)
adaround_step_kernel(const float* __restrict__ w, const float* __restrict__ delta, const float* __restrict__ zp,
                     float* __restrict__ alpha, const float* __restrict__ grad_w, float* __restrict__ adam_m,
                     float* __restrict__ adam_v, long long k, long long total, int level, float step_size,
                     float inv_sqrt_bc2, float b, float lambda, float* round_loss) {
  const float beta1 = 0.9f, beta2 = 0.999f, eps = 1e-8f;
  double rl = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / k;
    const float d = delta[row], z = zp[row];
    const float a = alpha[i];
    const float sg = sigmoid_f(a);
    const float hr = __fadd_rn(__fmul_rn(sg, 1.2f), -0.1f);
    const float h = fminf(fmaxf(hr, 0.f), 1.f);
    const float dh = (hr >= 0.f && hr <= 1.f) ? 1.2f * sg * (1.f - sg) : 0.f;  // clamp passes grad on [0,1]
    const float pre = floorf(__fdiv_rn(w[i], d)) + h + z;
    const bool in_range = pre >= 0.f && pre <= (float)(level - 1);
    float g = in_range ? grad_w[i] * d * dh : 0.f;
    if (b > 0.f) {
      const float u = 2.f * h - 1.f;
      const float au = fabsf(u);
      rl += (double)(1.f - powf(au, b));
      const float sgn = u > 0.f ? 1.f : (u < 0.f ? -1.f : 0.f);
      g += -lambda * b * powf(au, b - 1.f) * sgn * 2.f * dh;
    }
    const float m = beta1 * adam_m[i] + (1.f - beta1) * g;
    const float v = beta2 * adam_v[i] + (1.f - beta2) * g * g;
    adam_m[i] = m;
    adam_v[i] = v;
    alpha[i] = a - step_size * (m / (sqrtf(v) * inv_sqrt_bc2 + eps));
  }
  if (b > 0.f) {
    rl = warp_sum_d(rl);
    __shared__ double part[RED_THREADS / 32];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = rl;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int i = 1; i < RED_THREADS / 32; ++i) rl += part[i];
      atomicAdd(round_loss, (float)(rl * (double)lambda));
    }
  }
}

__global__ void __launch_bounds__(RED_THREADS) rec_loss_kernel(const float* __restrict__ pred,
                                                               const float* __restrict__ tgt, long long count,
                                                               float inv_batch, float* loss, float* grad) {
  double acc = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x) {
    const float e = pred[i] - tgt[i];
    acc += (double)e * (double)e;
    if (grad) grad[i] = 2.f * e * inv_batch;
  }
  acc = warp_sum_d(acc);
  __shared__ double part[RED_THREADS / 32];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < RED_THREADS / 32; ++i) acc += part[i];
    atomicAdd(loss, (float)(acc * (double)inv_batch));
  }
}

}  // namespace tfmq

using namespace tfmq;

static int grid_for(tfmq_ctx* ctx, long long n) {
  long long b = (n + RED_THREADS - 1) / RED_THREADS;
  const long long cap = (long long)ctx->sm_count * 8;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

static int minmax_launch(tfmq_ctx* ctx, const float* x, long long row_stride, long long rows, long long cols, int* bits,
                         cudaStream_t st, bool init, int slot_per_row = 1) {
  if (init) {
    minmax_init_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(bits, rows);
    TFMQ_LAUNCH_CHECK("minmax_init");
  }
  long long chunks = (long long)ctx->sm_count * 8 / (rows > 0 ? rows : 1);
  if (chunks < 1) chunks = 1;
  long long chunk = (cols + chunks - 1) / chunks;
  if (chunk < 4096) chunk = 4096;
  chunks = (cols + chunk - 1) / chunk;
  TFMQ_REQUIRE(rows <= 65535, TFMQ_ERR_SHAPE, "minmax: rows %lld > 65535", rows);
  minmax_kernel<<<dim3((unsigned)chunks, (unsigned)rows), RED_THREADS, 0, st>>>(x, row_stride, cols, chunk, bits,
                                                                                slot_per_row);
  TFMQ_LAUNCH_CHECK("minmax");
  return TFMQ_OK;
}

extern "C" int tfmq_minmax_rows(tfmq_ctx* ctx, const float* x, int64_t rows, int64_t cols, float* mm, void* stream) {
  if (!ctx) return TFMQ_ERR_ARG;
  TFMQ_REQUIRE(x && mm, TFMQ_ERR_ARG, "minmax_rows: null pointer");
  if (rows == 0) return TFMQ_OK;
  TFMQ_REQUIRE(cols > 0, TFMQ_ERR_SHAPE, "minmax_rows: empty rows");
  cudaStream_t st = tfmq_stream(stream);
  int* bits = reinterpret_cast<int*>(mm);
  for (int64_t r0 = 0; r0 < rows; r0 += 65535) {
    const int64_t nr = rows - r0 < 65535 ? rows - r0 : 65535;
    int rc = minmax_launch(ctx, x + r0 * cols, cols, nr, cols, bits + 2 * r0, st, true);
    if (rc) return rc;
  }
  minmax_decode_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(bits, rows);
  TFMQ_LAUNCH_CHECK("minmax_decode");
  return TFMQ_OK;
}

extern "C" int tfmq_mse_scale_search(tfmq_ctx* ctx, const float* x, int64_t rows, int64_t cols, int level,
                                     float* delta, float* zp, void* stream) {
  if (!ctx) return TFMQ_ERR_ARG;
  TFMQ_REQUIRE(x && delta && zp, TFMQ_ERR_ARG, "mse_scale_search: null pointer");
  TFMQ_REQUIRE(level >= 2, TFMQ_ERR_ARG, "mse_scale_search: level");
  if (rows == 0) return TFMQ_OK;
  TFMQ_REQUIRE(cols > 0, TFMQ_ERR_SHAPE, "mse_scale_search: empty rows");
  cudaStream_t st = tfmq_stream(stream);
  // scratch: mm[rows][2] floats + scores[rows][80] doubles, stream-ordered allocation
  void* scratch = nullptr;
  const size_t mm_bytes = ((size_t)rows * 2 * sizeof(float) + 255) & ~(size_t)255;
  const size_t sc_bytes = (size_t)rows * MSE_CAND * sizeof(double);
  cudaError_t e = cudaMallocAsync(&scratch, mm_bytes + sc_bytes, st);
  if (e != cudaSuccess) return tfmq_fail(ctx, TFMQ_ERR_CUDA, "mse_scale_search: alloc: %s", cudaGetErrorString(e));
  float* mm = reinterpret_cast<float*>(scratch);
  double* scores = reinterpret_cast<double*>(reinterpret_cast<char*>(scratch) + mm_bytes);
  cudaMemsetAsync(scores, 0, sc_bytes, st);
  int rc = tfmq_minmax_rows(ctx, x, rows, cols, mm, stream);
  if (rc == TFMQ_OK) {
    for (int64_t r0 = 0; r0 < rows && rc == TFMQ_OK; r0 += 65535) {
      const int64_t nr = rows - r0 < 65535 ? rows - r0 : 65535;
      long long chunks = (long long)ctx->sm_count * 4 / nr;
      if (chunks < 1) chunks = 1;
      long long chunk = (cols + chunks - 1) / chunks;
      if (chunk < 2048) chunk = 2048;
      chunks = (cols + chunk - 1) / chunk;
      mse_score_kernel<<<dim3((unsigned)chunks, (unsigned)nr), RED_THREADS, 0, st>>>(
          x + r0 * cols, cols, chunk, level, mm + 2 * r0, scores + r0 * MSE_CAND);
      cudaError_t le = cudaGetLastError();
      if (le != cudaSuccess) rc = tfmq_fail(ctx, TFMQ_ERR_CUDA, "mse_score: %s", cudaGetErrorString(le));
      else ctx->launches++;
    }
  }
  if (rc == TFMQ_OK) {
    mse_pick_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(scores, mm, rows, cols, level, delta, zp);
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) rc = tfmq_fail(ctx, TFMQ_ERR_CUDA, "mse_pick: %s", cudaGetErrorString(le));
    else ctx->launches++;
  }
  cudaFreeAsync(scratch, st);
  return rc;
}

extern "C" int tfmq_act_range_update(tfmq_ctx* ctx, const float* x, int64_t ld, int64_t pixels, int c, float momentum,
                                     int level, float* state, float* aq, void* stream) {
  if (!ctx) return TFMQ_ERR_ARG;
  TFMQ_REQUIRE(x && state && aq, TFMQ_ERR_ARG, "act_range_update: null pointer");
  TFMQ_REQUIRE(pixels > 0 && c > 0, TFMQ_ERR_SHAPE, "act_range_update: empty tensor");
  cudaStream_t st = tfmq_stream(stream);
  int* bits = reinterpret_cast<int*>(state) + 2;
  if (ld == c) {
    int rc = minmax_launch(ctx, x, 0, 1, pixels * c, bits, st, false);
    if (rc) return rc;
  } else {
    // window into a wider buffer: one grid row per pixel, all reducing into the same slot
    for (int64_t p0 = 0; p0 < pixels; p0 += 65535) {
      const int64_t np = pixels - p0 < 65535 ? pixels - p0 : 65535;
      int rc = minmax_launch(ctx, x + p0 * ld, ld, np, c, bits, st, false, 0);
      if (rc) return rc;
    }
  }
  act_range_finalize_kernel<<<1, 1, 0, st>>>(state, aq, momentum, level);
  TFMQ_LAUNCH_CHECK("act_range_finalize");
  return TFMQ_OK;
}

extern "C" int tfmq_adaround_soft(tfmq_ctx* ctx, const float* w, const float* delta, const float* zp,
                                  const float* alpha, int cout, int64_t k, int level, float* w_soft, void* stream) {
  if (!ctx) return TFMQ_ERR_ARG;
  TFMQ_REQUIRE(w && delta && zp && alpha && w_soft, TFMQ_ERR_ARG, "adaround_soft: null pointer");
  const long long total = (long long)cout * k;
  if (total == 0) return TFMQ_OK;
  adaround_soft_kernel<<<grid_for(ctx, total), RED_THREADS, 0, tfmq_stream(stream)>>>(w, delta, zp, alpha, k, total,
                                                                                      level, w_soft);
  TFMQ_LAUNCH_CHECK("adaround_soft");
  return TFMQ_OK;
}

extern "C" int tfmq_adaround_step(tfmq_ctx* ctx, const float* w, const float* delta, const float* zp, float* alpha,
                                  const float* grad_w, float* adam_m, float* adam_v, int cout, int64_t k, int level,
                                  int step, float lr, float b, float lambda, float* round_loss, void* stream) {
  if (!ctx) return TFMQ_ERR_ARG;
  TFMQ_REQUIRE(w && delta && zp && alpha && grad_w && adam_m && adam_v && round_loss, TFMQ_ERR_ARG,
               "adaround_step: null pointer");
  TFMQ_REQUIRE(step >= 1, TFMQ_ERR_ARG, "adaround_step: step must be >= 1");
  const long long total = (long long)cout * k;
  if (total == 0) return TFMQ_OK;
  const double bc1 = 1.0 - pow(0.9, (double)step);
  const double bc2 = 1.0 - pow(0.999, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  adaround_step_kernel<<<grid_for(ctx, total), RED_THREADS, 0, tfmq_stream(stream)>>>(
      w, delta, zp, alpha, grad_w, adam_m, adam_v, k, total, level, step_size, inv_sqrt_bc2, b, lambda, round_loss);
  TFMQ_LAUNCH_CHECK("adaround_step");
  return TFMQ_OK;
}

extern "C" int tfmq_rec_loss(tfmq_ctx* ctx, const float* pred, const float* tgt, int64_t count, int batch, float* loss,
                             float* grad_or_null, void* stream) {
  if (!ctx) return TFMQ_ERR_ARG;
  TFMQ_REQUIRE(pred && tgt && loss && batch > 0, TFMQ_ERR_ARG, "rec_loss: bad argument");
  if (count == 0) return TFMQ_OK;
  rec_loss_kernel<<<grid_for(ctx, count), RED_THREADS, 0, tfmq_stream(stream)>>>(pred, tgt, count, 1.f / (float)batch,
                                                                                 loss, grad_or_null);
  TFMQ_LAUNCH_CHECK("rec_loss");
  return TFMQ_OK;
}
