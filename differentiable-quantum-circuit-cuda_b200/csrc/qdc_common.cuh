// Common types, error convention and launch helpers for the B200 (sm_100a)
// differentiable statevector library.  One precision per library build, like
// the reference (-DQDC_F64 here <-> -DF64 at /root/reference/src/primitives.cu:11-29).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef QDC_F64
typedef double real_t;
typedef double2 cplx_t;
typedef double2 vec_t;  // one amplitude per 128-bit access
#define QDC_LV 0        // log2(amplitudes per vec_t)
#else
typedef float real_t;
typedef float2 cplx_t;
typedef float4 vec_t;   // two amplitudes per 128-bit access
#define QDC_LV 1
#endif
#define QDC_VA (1 << QDC_LV)  // amplitudes per vec_t
#define QDC_VR (2 * QDC_VA)   // reals per vec_t

union VecU {
  vec_t v;
  real_t r[QDC_VR];
};

// ---- error convention: NULL on success, heap message otherwise ------------
// Same contract as the reference's CUDA_CHECK (src/primitives.cu:32-49): the
// caller (QuantizedTensor::cuda_panic, src/quantized_tensor.rs:37-42) prints
// and never frees the string.
static inline char* qdc_errf(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
#include <stdarg.h>
static inline char* qdc_errf(const char* fmt, ...) {
  char* s = (char*)malloc(1024);
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(s, 1024, fmt, ap);
  va_end(ap);
  return s;
}

#define QDC_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t status_ = (cudaError_t)(call);                                      \
    if (status_ != cudaSuccess) {                                                   \
      return qdc_errf("CUDA ERROR: call of a function \"%s\" in line %d of file %s " \
                      "failed with %s.",                                            \
                      #call, __LINE__, __FILE__, cudaGetErrorName(status_));        \
    }                                                                               \
  } while (0)

#define QDC_TRY(expr)              \
  do {                             \
    const char* e_ = (expr);       \
    if (e_ != nullptr) return e_;  \
  } while (0)

// ---- index helpers -----------------------------------------------------
// Insert a zero bit at position `pos` of `i` (same role as the reference's
// INSERT_ZERO, src/primitives.cu:104-105, but 64-bit clean for n > 30).
__host__ __device__ __forceinline__ uint64_t ins0(uint64_t i, int pos) {
  const uint64_t low = i & ((1ull << pos) - 1ull);
  return ((i >> pos) << (pos + 1)) | low;
}

__host__ __device__ __forceinline__ uint32_t ins0_32(uint32_t i, int pos) {
  const uint32_t low = i & ((1u << pos) - 1u);
  return ((i >> pos) << (pos + 1)) | low;
}

// ---- gate parameter blocks (passed by value as kernel parameters; no
// __constant__ slots, so concurrent streams/threads do not race the way
// src/primitives.cu:109-111 + README.md:13 do) -----------------------------
struct GateQ1 {
  real_t re[4], im[4];
};
struct GateQ2 {
  real_t re[16], im[16];
};

struct DeviceInfo {
  int device;
  int sm_count;
};

static inline const char* qdc_device_info(DeviceInfo* out) {
  static thread_local DeviceInfo cached = {-1, 0};
  int dev = 0;
  QDC_CUDA(cudaGetDevice(&dev));
  if (cached.device != dev) {
    int sms = 0;
    QDC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cached.device = dev;
    cached.sm_count = sms;
  }
  *out = cached;
  return nullptr;
}
