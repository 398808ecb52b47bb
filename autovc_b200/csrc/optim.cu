// Optimizer step of the training iteration (solver_encoder.py:130 builds torch.optim.Adam(lr, betas (0.9, 0.999), eps 1e-8,
// no weight decay); :300 calls .step()).  One launch updates every parameter tensor of the Generator: the host hands over a
// device table of (param, grad, exp_avg, exp_avg_sq) pointers plus a static chunk map, so the 74 tensors (28.3 M elements)
// cost one pass of 16 B read + 12 B written per element instead of torch's eight multi-tensor passes.
#include <math.h>

#include <algorithm>

#include "common.cuh"

namespace avc {

constexpr int ADAM_THREADS = 256;
constexpr int ADAM_CHUNK = 4096;       // elements per chunk: 4 float4 per thread

struct AdamTensor {
  float* p;
  const float* g;
  float* m;
  float* v;
  unsigned long long n;                // elements
};

// torch._multi_tensor_adam, element by element:
//   m = lerp(m, g, 1 - b1);  v = v * b2 + (1 - b2) * g * g;  p -= step_size * m / (sqrt(v) / sqrt(bc2) + eps)
__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float w1, float b2, float w2, float eps,
                                         float step_size, float bc2_sqrt, float gscale) {
  g *= gscale;
  m = fmaf(w1, g - m, m);
  v = fmaf(w2 * g, g, v * b2);
  const float denom = sqrtf(v) / bc2_sqrt + eps;
  p = fmaf(-step_size, m / denom, p);
}

__global__ void __launch_bounds__(ADAM_THREADS)
adam_multi_kernel(const AdamTensor* __restrict__ tab, const int2* __restrict__ chunks, int nchunks, float w1, float b2, float w2,
                  float eps, float step_size, float bc2_sqrt, float gscale) {
  for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const int2 ch = chunks[c];
    const AdamTensor t = tab[ch.x];
    const size_t start = (size_t)ch.y * ADAM_CHUNK;
    const size_t end = min(start + (size_t)ADAM_CHUNK, (size_t)t.n);
    const bool vec = ((((uintptr_t)t.p) | ((uintptr_t)t.g) | ((uintptr_t)t.m) | ((uintptr_t)t.v)) & 15) == 0;
    if (vec && end - start == ADAM_CHUNK) {
      float4 p4[4], g4[4], m4[4], v4[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const size_t i = start + (size_t)(k * ADAM_THREADS + threadIdx.x) * 4;
        p4[k] = *reinterpret_cast<const float4*>(t.p + i);
        g4[k] = __ldcs(reinterpret_cast<const float4*>(t.g + i));     // the gradient is dead after this pass
        m4[k] = *reinterpret_cast<const float4*>(t.m + i);
        v4[k] = *reinterpret_cast<const float4*>(t.v + i);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        adam_one(p4[k].x, g4[k].x, m4[k].x, v4[k].x, w1, b2, w2, eps, step_size, bc2_sqrt, gscale);
        adam_one(p4[k].y, g4[k].y, m4[k].y, v4[k].y, w1, b2, w2, eps, step_size, bc2_sqrt, gscale);
        adam_one(p4[k].z, g4[k].z, m4[k].z, v4[k].z, w1, b2, w2, eps, step_size, bc2_sqrt, gscale);
        adam_one(p4[k].w, g4[k].w, m4[k].w, v4[k].w, w1, b2, w2, eps, step_size, bc2_sqrt, gscale);
        const size_t i = start + (size_t)(k * ADAM_THREADS + threadIdx.x) * 4;
        *reinterpret_cast<float4*>(t.p + i) = p4[k];
        *reinterpret_cast<float4*>(t.m + i) = m4[k];
        *reinterpret_cast<float4*>(t.v + i) = v4[k];
      }
    } else {
      for (size_t i = start + threadIdx.x; i < end; i += ADAM_THREADS) {
        float p = t.p[i], m = t.m[i], v = t.v[i];
        adam_one(p, t.g[i], m, v, w1, b2, w2, eps, step_size, bc2_sqrt, gscale);
        t.p[i] = p; t.m[i] = m; t.v[i] = v;
      }
    }
  }
}

// solver_encoder.py:168-177 model_EMA: avg = ema * flat + (1 - ema) * flat, written back over the parameters.  In exact
// arithmetic that is the identity; in fp32 it is three roundings per element (two products, one sum) that move
// most parameters by an ulp, and a checkpoint holds the result.  Same operations, same order, no fma contraction.
__global__ void __launch_bounds__(ADAM_THREADS)
ema_blend_kernel(const AdamTensor* __restrict__ tab, const int2* __restrict__ chunks, int nchunks, float a, float b) {
  for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const int2 ch = chunks[c];
    const AdamTensor t = tab[ch.x];
    const size_t start = (size_t)ch.y * ADAM_CHUNK;
    const size_t end = min(start + (size_t)ADAM_CHUNK, (size_t)t.n);
    if ((((uintptr_t)t.p) & 15) == 0 && end - start == ADAM_CHUNK) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const size_t i = start + (size_t)(k * ADAM_THREADS + threadIdx.x) * 4;
        float4 v = *reinterpret_cast<const float4*>(t.p + i);
        v.x = __fadd_rn(__fmul_rn(a, v.x), __fmul_rn(b, v.x));
        v.y = __fadd_rn(__fmul_rn(a, v.y), __fmul_rn(b, v.y));
        v.z = __fadd_rn(__fmul_rn(a, v.z), __fmul_rn(b, v.z));
        v.w = __fadd_rn(__fmul_rn(a, v.w), __fmul_rn(b, v.w));
        *reinterpret_cast<float4*>(t.p + i) = v;
      }
    } else {
      for (size_t i = start + threadIdx.x; i < end; i += ADAM_THREADS) {
        const float x = t.p[i];
        t.p[i] = __fadd_rn(__fmul_rn(a, x), __fmul_rn(b, x));
      }
    }
  }
}

}  // namespace avc

using namespace avc;

extern "C" int avc_adam_chunk_elems(void) { return ADAM_CHUNK; }

extern "C" int avc_adam_step(const void* table, const void* chunks, int nchunks, double lr, double beta1, double beta2, double eps,
                             int step, float grad_scale, void* stream) {
  AVC_REQUIRE(table && chunks && nchunks > 0, "avc_adam_step: null table / no chunks");
  AVC_REQUIRE(step >= 1 && beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 && eps >= 0.0, "avc_adam_step: bad hyper-parameters");
  // the scalar prologue of torch.optim.adam._multi_tensor_adam (python floats = doubles), rounded to fp32 where torch does
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  const float step_size = (float)(lr / bc1);
  const float bc2_sqrt = (float)sqrt(bc2);
  const int grid = std::min(nchunks, num_sms() * 8);
  adam_multi_kernel<<<grid, ADAM_THREADS, 0, as_stream(stream)>>>(
      reinterpret_cast<const AdamTensor*>(table), reinterpret_cast<const int2*>(chunks), nchunks, (float)(1.0 - beta1), (float)beta2,
      (float)(1.0 - beta2), (float)eps, step_size, bc2_sqrt, grad_scale);
  AVC_LAUNCHED();
  return AVC_OK;
}

extern "C" int avc_ema_blend(const void* table, const void* chunks, int nchunks, double ema, void* stream) {
  AVC_REQUIRE(table && chunks && nchunks > 0, "avc_ema_blend: null table / no chunks");
  // torch multiplies an fp32 tensor by a python float in fp32 with the scalar cast to float; (1 - ema) is formed in double first
  const int grid = std::min(nchunks, num_sms() * 8);
  ema_blend_kernel<<<grid, ADAM_THREADS, 0, as_stream(stream)>>>(reinterpret_cast<const AdamTensor*>(table),
                                                                 reinterpret_cast<const int2*>(chunks), nchunks, (float)ema,
                                                                 (float)(1.0 - ema));
  AVC_LAUNCHED();
  return AVC_OK;
}
