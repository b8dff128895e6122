// Context management and the weight pre-pack entry point of the C ABI.
#include <cstdlib>
#include <cstring>
#include <new>

#include "ctx.h"

namespace tfmq {

// one CTA per output channel: quantise, pack two codes per byte, reduce sum(q - zp)
__global__ void __launch_bounds__(256) pack_w4_kernel(const float* __restrict__ w, const float* __restrict__ delta,
                                                      const float* __restrict__ zp, const float* __restrict__ alpha,
                                                      int k, uint8_t* __restrict__ codes, uint8_t* __restrict__ packed,
                                                      int32_t* __restrict__ wsum) {
  const int row = blockIdx.x;
  const float d = delta[row], z = zp[row];
  const float* wr = w + (long long)row * k;
  const float* ar = alpha ? alpha + (long long)row * k : nullptr;
  int local = 0;
  for (int b = threadIdx.x; b < k / 2; b += blockDim.x) {
    const int g = b >> 4, i = b & 15;
    int q2[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int idx = g * 32 + h * 16 + i;
      const float v = __fdiv_rn(wr[idx], d);
      float q = ar ? (floorf(v) + (ar[idx] >= 0.f ? 1.f : 0.f)) : rintf(v);
      q = fminf(fmaxf(q + z, 0.f), 15.f);
      q2[h] = (int)q;
      if (codes) codes[(long long)row * k + idx] = (uint8_t)q2[h];
      local += q2[h] - (int)z;
    }
    packed[(long long)row * (k / 2) + b] = (uint8_t)(q2[0] | (q2[1] << 4));
  }
  __shared__ int red[256];
  red[threadIdx.x] = local;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) wsum[row] = red[0];
}

}  // namespace tfmq

extern "C" int tfmq_abi_version(void) { return 6; }   // 4: tfmq_attention_h16, tfmq_conv_h16_desc.out_hi / out_lo; 5: tfmq_conv_h16_desc.ksplit; 6: int32 weight zero points, ddim_update coef[5]

extern "C" int tfmq_create(tfmq_ctx** out, int device) {
  if (!out) return TFMQ_ERR_ARG;
  *out = nullptr;
  tfmq_ctx* ctx = new (std::nothrow) tfmq_ctx();
  if (!ctx) return TFMQ_ERR_CUDA;
  memset(ctx, 0, sizeof(*ctx));
  ctx->device = device;
  *out = ctx;  // returned even on failure so the caller can read the message
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return tfmq_fail(ctx, TFMQ_ERR_UNAVAILABLE, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return tfmq_fail(ctx, TFMQ_ERR_UNAVAILABLE, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return tfmq_fail(ctx, TFMQ_ERR_UNAVAILABLE, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                     prop.major, prop.minor);
  ctx->sm_count = prop.multiProcessorCount;
  ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
    return tfmq_fail(ctx, TFMQ_ERR_UNAVAILABLE, "cuTensorMapEncodeTiled entry point not found");
  ctx->encode_tiled = reinterpret_cast<tfmq_encode_tiled_fn>(fn);
  return TFMQ_OK;
}

extern "C" int tfmq_destroy(tfmq_ctx* ctx) {
  delete ctx;
  return TFMQ_OK;
}

extern "C" const char* tfmq_last_error(tfmq_ctx* ctx) { return ctx ? ctx->err : "null context"; }

extern "C" int64_t tfmq_launch_count(tfmq_ctx* ctx) { return ctx ? ctx->launches : -1; }

extern "C" int tfmq_pack_w4(tfmq_ctx* ctx, const float* w, const float* delta, const float* zp,
                            const float* alpha_or_null, int cout, int k, uint8_t* codes_or_null, uint8_t* packed,
                            int32_t* wsum, void* stream) {
  if (!ctx) return TFMQ_ERR_ARG;
  TFMQ_REQUIRE(w && delta && zp && packed && wsum, TFMQ_ERR_ARG, "pack_w4: null pointer");
  TFMQ_REQUIRE(k > 0 && k % 32 == 0, TFMQ_ERR_SHAPE, "pack_w4: k=%d must be a positive multiple of 32", k);
  if (cout == 0) return TFMQ_OK;
  tfmq::pack_w4_kernel<<<cout, 256, 0, tfmq_stream(stream)>>>(w, delta, zp, alpha_or_null, k, codes_or_null, packed,
                                                              wsum);
  TFMQ_LAUNCH_CHECK("pack_w4");
  return TFMQ_OK;
}
