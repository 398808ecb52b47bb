// fp32 CUDA-core GEMM mainloops (AVC_PREC_FP32 parity path).
//
// Two tile kernels share one 128x128x16 register-blocked outer-product microkernel:
//   NT + taps : C[m,n]      = sum_tap sum_k A[row(m,tap),k] * W[tap][n][k]       (conv fwd/dgrad, projections)
//   TN + taps : dW[tap][n,k] = sum_m dY[m,n] * X[row(m,tap),k]                    (weight gradients)
// row(m,tap) shifts the time index inside an utterance and reads zero outside [0,T): that is
// the zero padding of Conv1d(k=5,p=2) (model_vc_mel.py:28-31) and the h_{t-1} shift of dW_hh.
// The tensor-core (tcgen05) path in tc_gemm.cu has the same contracts.
#pragma once
#include "common.cuh"

namespace avc {

constexpr int SG_BM = 128, SG_BN = 128, SG_BK = 16, SG_THREADS = 256, SG_PAD = 4;

struct SimtSmem {
  float a[SG_BK][SG_BM + SG_PAD];
  float b[SG_BK][SG_BN + SG_PAD];
};

// acc[i][j]: rows (ty*4 + i%4 + 64*(i/4)), cols (tx*4 + j%4 + 64*(j/4)); tx = tid%16, ty = tid/16.
__device__ __forceinline__ void simt_microkernel(const SimtSmem& s, float (&acc)[8][8], int tx, int ty) {
#pragma unroll
  for (int kk = 0; kk < SG_BK; ++kk) {
    float a[8], b[8];
    *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(&s.a[kk][ty * 4]);
    *reinterpret_cast<float4*>(&a[4]) = *reinterpret_cast<const float4*>(&s.a[kk][64 + ty * 4]);
    *reinterpret_cast<float4*>(&b[0]) = *reinterpret_cast<const float4*>(&s.b[kk][tx * 4]);
    *reinterpret_cast<float4*>(&b[4]) = *reinterpret_cast<const float4*>(&s.b[kk][64 + tx * 4]);
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}

// Two-level summation: products are accumulated in `part` over SG_FLUSH slabs (64 k-values) and then
// folded into `acc`, which keeps fp32 rounding growth near that of a blocked (oneDNN/cuDNN-style)
// reduction instead of one K-long serial chain -- this is what holds the <=1e-4 parity through the
// 14 BatchNorm-amplified layers.
constexpr int SG_FLUSH = 4;
__device__ __forceinline__ void zero_acc(float (&a)[8][8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) a[i][j] = 0.f;
}
__device__ __forceinline__ void flush_acc(float (&acc)[8][8], float (&part)[8][8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc[i][j] += part[i][j];
      part[i][j] = 0.f;
    }
}

__device__ __forceinline__ int acc_row(int ty, int i) { return ty * 4 + (i & 3) + ((i >> 2) << 6); }
__device__ __forceinline__ int acc_col(int tx, int j) { return tx * 4 + (j & 3) + ((j >> 2) << 6); }

// Generic NT mainloop.  LoadA(m, tap, k) and LoadB(n, tap, k) return 0 outside their ranges.
// Both operands are K-contiguous; each thread fetches 8 A and 8 B scalars per 16-wide slab
// (16 consecutive k per row -> 64 B segments), register-prefetched one slab ahead.
template <class LoadA, class LoadB>
__device__ __forceinline__ void simt_mainloop_nt(SimtSmem& s, float (&acc)[8][8], int m0, int n0, int K, int ntaps,
                                                 const LoadA& loadA, const LoadB& loadB) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int lk = tid & 15, lr = tid >> 4;
  const int kslabs = (K + SG_BK - 1) / SG_BK;
  const int total = kslabs * ntaps;
  float ra[8], rb[8];
  auto gload = [&](int slab) {
    const int tap = slab / kslabs;
    const int k = (slab - tap * kslabs) * SG_BK + lk;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      ra[j] = loadA(m0 + lr + 16 * j, tap, k);
      rb[j] = loadB(n0 + lr + 16 * j, tap, k);
    }
  };
  if (total > 0) gload(0);
  float part[8][8];
  zero_acc(part);
  for (int slab = 0; slab < total; ++slab) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s.a[lk][lr + 16 * j] = ra[j];
      s.b[lk][lr + 16 * j] = rb[j];
    }
    __syncthreads();
    if (slab + 1 < total) gload(slab + 1);
    simt_microkernel(s, part, tx, ty);
    if ((slab & (SG_FLUSH - 1)) == SG_FLUSH - 1) flush_acc(acc, part);
    __syncthreads();
  }
  flush_acc(acc, part);
}

// Generic TN mainloop over reduction rows [r0, r1).  LoadA(r, n) = dY[r, n]; LoadB(r, tap, k) = X[row(r,tap), k].
// Both operands are contiguous along the *output* dims, so slabs are copied without transposition.
template <class LoadA, class LoadB>
__device__ __forceinline__ void simt_mainloop_tn(SimtSmem& s, float (&acc)[8][8], int n0, int k0, int r0, int r1, int tap,
                                                 const LoadA& loadA, const LoadB& loadB) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int lc = tid & 127, lr = tid >> 7;  // 2 rows x 128 columns per pass
  float ra[8], rb[8];
  auto gload = [&](int r) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int rr = r + lr + 2 * j;
      const bool ok = rr < r1;
      ra[j] = ok ? loadA(rr, n0 + lc) : 0.f;
      rb[j] = ok ? loadB(rr, tap, k0 + lc) : 0.f;
    }
  };
  if (r0 < r1) gload(r0);
  float part[8][8];
  zero_acc(part);
  int slab = 0;
  for (int r = r0; r < r1; r += SG_BK, ++slab) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s.a[lr + 2 * j][lc] = ra[j];
      s.b[lr + 2 * j][lc] = rb[j];
    }
    __syncthreads();
    if (r + SG_BK < r1) gload(r + SG_BK);
    simt_microkernel(s, part, tx, ty);
    if ((slab & (SG_FLUSH - 1)) == SG_FLUSH - 1) flush_acc(acc, part);
    __syncthreads();
  }
  flush_acc(acc, part);
}

// time-shifted row lookup shared by every tap kernel
struct TapRows {
  int T, shift0;
  // returns source row index, or -1 when the shifted frame falls outside the utterance
  __device__ __forceinline__ int operator()(int m, int tap) const {
    const int t = m % T;
    const int ts = t + shift0 + tap;
    return (ts >= 0 && ts < T) ? (m + shift0 + tap) : -1;
  }
};

}  // namespace avc
