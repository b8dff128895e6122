// Internal context shared by the C-ABI entry points.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/tfmq_b200.h"

typedef CUresult (*tfmq_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                         const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                         CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                         CUtensorMapFloatOOBfill);

struct tfmq_ctx {
  int device;
  int sm_count;
  int max_smem_optin;
  tfmq_encode_tiled_fn encode_tiled;
  int64_t launches;
  char err[512];
};

static inline int tfmq_fail(tfmq_ctx* ctx, int status, const char* fmt, ...) {
  if (ctx) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
    va_end(ap);
  }
  return status;
}

#define TFMQ_REQUIRE(cond, status, ...) \
  do {                                  \
    if (!(cond)) return tfmq_fail(ctx, status, __VA_ARGS__); \
  } while (0)

// call after every kernel launch
#define TFMQ_LAUNCH_CHECK(name)                                                                        \
  do {                                                                                                 \
    cudaError_t e__ = cudaGetLastError();                                                              \
    if (e__ != cudaSuccess) return tfmq_fail(ctx, TFMQ_ERR_CUDA, "%s: %s", name, cudaGetErrorString(e__)); \
    ctx->launches++;                                                                                   \
  } while (0)

static inline cudaStream_t tfmq_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Programmatic dependent launch of the conv and producer kernels (TFMQ_PDL=0 switches it off): the next kernel's CTAs are
// scheduled, and run their prologue (barrier init, TMEM allocation, descriptor prefetch), while the tail of the previous
// kernel drains; every such kernel executes griddepcontrol.wait before it touches global memory.
static inline bool tfmq_pdl() {
  static const bool on = getenv("TFMQ_PDL") ? atoi(getenv("TFMQ_PDL")) != 0 : false;
  return on;
}
static inline int tfmq_pdl_attr(cudaLaunchAttribute* a) {
  if (!tfmq_pdl()) return 0;
  a->id = cudaLaunchAttributeProgrammaticStreamSerialization;
  a->val.programmaticStreamSerializationAllowed = 1;
  return 1;
}
