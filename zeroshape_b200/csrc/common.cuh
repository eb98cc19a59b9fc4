// Shared helpers for the zeroshape_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include "../../include/zeroshape_b200.h"

namespace zs {

void set_error(const char* fmt, ...);
// kernels launched by this library since load (for bench.py's `gpu_launches` claim)
void count_launches(int n);

#define ZS_REQUIRE(cond, ...)                                   \
  do {                                                          \
    if (!(cond)) {                                              \
      zs::set_error(__VA_ARGS__);                               \
      return ZS_ERR_ARG;                                        \
    }                                                           \
  } while (0)

#define ZS_CUDA_CHECK_LAUNCH(name)                                                  \
  do {                                                                              \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ != cudaSuccess) {                                                       \
      zs::set_error("%s: CUDA launch failed: %s", name, cudaGetErrorString(e__));   \
      return ZS_ERR_CUDA;                                                           \
    }                                                                               \
    zs::count_launches(1);                                                          \
  } while (0)

#define ZS_CUDA_CALL(expr)                                                          \
  do {                                                                              \
    cudaError_t e__ = (expr);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      zs::set_error("%s failed: %s", #expr, cudaGetErrorString(e__));               \
      return ZS_ERR_CUDA;                                                           \
    }                                                                               \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// number of SMs of the current device (cached per thread)
int sm_count();

// ---- activations (match PyTorch fp32 CPU semantics) -----------------------------------------
__device__ __forceinline__ float act_gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
// torch.nn.Softplus(beta=100, threshold=20): x if beta*x > 20 else log1p(exp(beta*x))/beta
__device__ __forceinline__ float act_softplus100(float x) {
  float bx = 100.0f * x;
  return bx > 20.0f ? x : log1pf(expf(bx)) * 0.01f;
}
__device__ __forceinline__ float act_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case ZS_ACT_RELU: return fmaxf(x, 0.0f);
    case ZS_ACT_GELU: return act_gelu_erf(x);
    case ZS_ACT_SOFTPLUS100: return act_softplus100(x);
    case ZS_ACT_SIGMOID: return act_sigmoid(x);
    case ZS_ACT_CLAMP01: return fminf(fmaxf(x, 0.0f), 1.0f);
    default: return x;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace zs
