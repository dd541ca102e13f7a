// Shared host/device helpers for libs3d_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/s3d.h"

namespace s3d {

void set_error(const char* fmt, ...);

#define S3D_CHECK_ARG(cond, ...)                         \
  do {                                                   \
    if (!(cond)) {                                       \
      s3d::set_error(__VA_ARGS__);                       \
      return S3D_ERR_INVALID;                            \
    }                                                    \
  } while (0)

#define S3D_CUDA(expr)                                                              \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      s3d::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return S3D_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

#define S3D_LAUNCH_CHECK() S3D_CUDA(cudaGetLastError())

int num_sms();   // SM count of the current device (cached)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float apply_act(float v, int act, float a) {
  switch (act) {
    case S3D_ACT_RELU:    return fmaxf(v, 0.f);
    case S3D_ACT_LEAKY:   return v > 0.f ? v : v * a;
    case S3D_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    case S3D_ACT_TANH:    return a * tanhf(v);
    default:              return v;
  }
}

// Out-of-line copy for epilogues that must stay small: 16 calls instead of 16 inlined switch bodies.
static __device__ __noinline__ float apply_act_slow(float v, int act, float a) { return apply_act(v, act, a); }

}  // namespace s3d
