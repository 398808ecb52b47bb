// Shared helpers for libautovc_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/autovc_b200.h"

namespace avc {

void set_error(const char* fmt, ...);
extern std::atomic<unsigned long long> g_launches;

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

#define AVC_REQUIRE(cond, ...)                    \
  do {                                            \
    if (!(cond)) {                                \
      avc::set_error(__VA_ARGS__);                \
      return AVC_ERR_INVALID;                     \
    }                                             \
  } while (0)

#define AVC_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      avc::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return AVC_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

// after a <<<>>> launch
#define AVC_LAUNCHED()                    \
  do {                                    \
    avc::g_launches.fetch_add(1);         \
    AVC_CUDA(cudaGetLastError());         \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline size_t ceil_div(size_t a, size_t b) { return (a + b - 1) / b; }

int num_sms();

}  // namespace avc
