// fp32 CUDA-core implementations of the GEMM-with-taps family + weight packers.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "simt_gemm.cuh"

namespace avc {

// ---------------------------------------------------------------------------------------
// NT + taps (conv fwd/dgrad, input projections, linear)
// ---------------------------------------------------------------------------------------
template <bool STATS, bool ACCUM>
__global__ void __launch_bounds__(SG_THREADS)
gemm_nt_taps_simt_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W,
                         const float* __restrict__ bias, float* __restrict__ C, int ldc, int M, int T, int N,
                         int K, int ntaps, int shift0, double* __restrict__ stats) {
  __shared__ SimtSmem s;
  const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
  const TapRows rows{T, shift0};
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  auto loadA = [&](int m, int tap, int k) -> float {
    if (m >= M || k >= K) return 0.f;
    const int r = rows(m, tap);
    return r >= 0 ? __ldg(A + (size_t)r * lda + k) : 0.f;
  };
  auto loadB = [&](int n, int tap, int k) -> float {
    return (n < N && k < K) ? __ldg(W + ((size_t)tap * N + n) * K + k) : 0.f;
  };
  simt_mainloop_nt(s, acc, m0, n0, K, ntaps, loadA, loadB);

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float bj[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int n = n0 + acc_col(tx, j);
    bj[j] = (bias != nullptr && n < N) ? bias[n] : 0.f;
  }
  float csum[8], csq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) csum[j] = csq[j] = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + acc_row(ty, i);
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + acc_col(tx, j);
      if (n < N) {
        float v = acc[i][j] + bj[j];
        if (ACCUM) v += C[(size_t)m * ldc + n];
        C[(size_t)m * ldc + n] = v;
        if (STATS) {
          csum[j] += v;
          csq[j] = fmaf(v, v, csq[j]);
        }
      }
    }
  }
  if (STATS) {
    // reduce the 16 row-groups (ty) of each column through shared memory, then one fp64 atomic per column
    float* red = &s.a[0][0];  // 2 * 16 * 128 floats = 16 KB <= sizeof(SimtSmem)
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      red[ty * 128 + acc_col(tx, j)] = csum[j];
      red[2048 + ty * 128 + acc_col(tx, j)] = csq[j];
    }
    __syncthreads();
    if (tid < 128) {
      const int n = n0 + tid;
      if (n < N) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          a += (double)red[r * 128 + tid];
          b += (double)red[2048 + r * 128 + tid];
        }
        atomicAdd(stats + n, a);
        atomicAdd(stats + N + n, b);
      }
    }
  }
}

int gemm_nt_taps_simt(const float* A, int lda, const float* W, const float* bias, float* C, int ldc, int nB, int T,
                      int N, int K, int ntaps, int shift0, double* stats, int accumulate, cudaStream_t st) {
  const int M = nB * T;
  dim3 grid(ceil_div(N, SG_BN), ceil_div(M, SG_BM));
  if (stats && accumulate) {
    set_error("avc_gemm_nt_taps: chan_stats and accumulate are mutually exclusive");
    return AVC_ERR_UNSUPPORTED;
  }
  if (stats)
    gemm_nt_taps_simt_kernel<true, false><<<grid, SG_THREADS, 0, st>>>(A, lda, W, bias, C, ldc, M, T, N, K, ntaps, shift0, stats);
  else if (accumulate)
    gemm_nt_taps_simt_kernel<false, true><<<grid, SG_THREADS, 0, st>>>(A, lda, W, bias, C, ldc, M, T, N, K, ntaps, shift0, nullptr);
  else
    gemm_nt_taps_simt_kernel<false, false><<<grid, SG_THREADS, 0, st>>>(A, lda, W, bias, C, ldc, M, T, N, K, ntaps, shift0, nullptr);
  AVC_LAUNCHED();
  return AVC_OK;
}

// ---------------------------------------------------------------------------------------
// TN + taps (weight gradients) : split over rows, partials to workspace, then a mapped reduce
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SG_THREADS)
gemm_tn_taps_simt_kernel(const float* __restrict__ dY, int ldy, const float* __restrict__ X, int ldx,
                         float* __restrict__ part, int M, int T, int N, int K, int ntaps, int shift0, int splits,
                         int rows_per_split) {
  __shared__ SimtSmem s;
  const int n0 = blockIdx.x * SG_BN, k0 = blockIdx.y * SG_BN;
  const int tap = blockIdx.z % ntaps, split = blockIdx.z / ntaps;
  const int r0 = split * rows_per_split;
  const int r1 = min(M, r0 + rows_per_split);
  const TapRows rows{T, shift0};
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  auto loadA = [&](int r, int n) -> float { return n < N ? __ldg(dY + (size_t)r * ldy + n) : 0.f; };
  auto loadB = [&](int r, int tp, int k) -> float {
    if (k >= K) return 0.f;
    const int src = rows(r, tp);
    return src >= 0 ? __ldg(X + (size_t)src * ldx + k) : 0.f;
  };
  simt_mainloop_tn(s, acc, n0, k0, r0, r1, tap, loadA, loadB);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float* out = part + ((size_t)split * ntaps + tap) * N * K;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int n = n0 + acc_row(ty, i);
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = k0 + acc_col(tx, j);
      if (k < K) out[(size_t)n * K + k] = acc[i][j];
    }
  }
}

// out index map shared with the tensor-core path
__device__ __forceinline__ size_t wgrad_out_index(int tap, int n, int k, int N, int K, int ntaps, int out_mode) {
  if (out_mode == 1) return ((size_t)n * K + k) * ntaps + tap;                 // PyTorch Conv1d (N, K, taps)
  if (out_mode == 2) {                                                          // LSTM: packed row u*4+g -> g*H+u
    const int H = N >> 2, u = n >> 2, g = n & 3;
    return (size_t)(g * H + u) * K + k;
  }
  return ((size_t)tap * N + n) * K + k;
}

__global__ void wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ dW, int N, int K, int ntaps,
                                    int splits, int out_mode, int accumulate) {
  const size_t total = (size_t)ntaps * N * K;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    float v = 0.f;
    for (int s = 0; s < splits; ++s) v += part[(size_t)s * total + i];
    const int k = (int)(i % K);
    const int n = (int)((i / K) % N);
    const int tap = (int)(i / ((size_t)K * N));
    const size_t o = wgrad_out_index(tap, n, k, N, K, ntaps, out_mode);
    dW[o] = accumulate ? dW[o] + v : v;
  }
}

// The split partials are summed in split order (deterministic), with the loads of several splits in flight.
// out_mode 1 (conv weight (Cout, Cin, taps)): one thread owns one (n, k) and writes its `ntaps` consecutive outputs, so
// both the partial reads (along k) and the stores (ntaps*4 contiguous bytes per thread) are coalesced.
template <int NTAPS>
__global__ void __launch_bounds__(256)
wgrad_reduce_conv_kernel(const float* __restrict__ part, float* __restrict__ dW, size_t nk, int splits, int accumulate) {
  const size_t total = nk * NTAPS;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nk; i += (size_t)gridDim.x * blockDim.x) {
    float v[NTAPS];
#pragma unroll
    for (int t = 0; t < NTAPS; ++t) v[t] = 0.f;
#pragma unroll 2
    for (int s = 0; s < splits; ++s) {
      float x[NTAPS];
#pragma unroll
      for (int t = 0; t < NTAPS; ++t) x[t] = part[(size_t)s * total + (size_t)t * nk + i];
#pragma unroll
      for (int t = 0; t < NTAPS; ++t) v[t] += x[t];
    }
#pragma unroll
    for (int t = 0; t < NTAPS; ++t) {
      const size_t o = i * NTAPS + t;
      dW[o] = accumulate ? dW[o] + v[t] : v[t];
    }
  }
}
// out_mode 0 / 2 with K % 4 == 0: float4 along k
__global__ void __launch_bounds__(256)
wgrad_reduce_vec4_kernel(const float4* __restrict__ part, float* __restrict__ dW, int N, int K, int ntaps, int splits,
                         int out_mode, int accumulate) {
  const size_t total4 = (size_t)ntaps * N * K / 4;
  const int K4 = K >> 2;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int s = 0; s < splits; ++s) {
      const float4 x = part[(size_t)s * total4 + i];
      v.x += x.x; v.y += x.y; v.z += x.z; v.w += x.w;
    }
    const int k = (int)(i % K4) * 4;
    const int n = (int)((i / K4) % N);
    const int tap = (int)(i / ((size_t)K4 * N));
    float4* o = reinterpret_cast<float4*>(dW + wgrad_out_index(tap, n, k, N, K, ntaps, out_mode));
    if (accumulate) {
      const float4 old = *o;
      v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w;
    }
    *o = v;
  }
}

static int tn_splits(int M, int N, int K, int ntaps) {
  const int tiles = ceil_div(N, SG_BN) * ceil_div(K, SG_BN) * ntaps;
  int splits = ceil_div(3 * num_sms(), tiles);
  const int max_splits = ceil_div(M, 512);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  return splits;
}

size_t gemm_tn_workspace_simt(int nB, int T, int N, int K, int ntaps) {
  return (size_t)tn_splits(nB * T, N, K, ntaps) * ntaps * N * K * sizeof(float);
}

int launch_wgrad_reduce(const float* part, float* dW, int N, int K, int ntaps, int splits, int out_mode, int accumulate,
                        cudaStream_t st) {
  const size_t total = (size_t)ntaps * N * K;
  const size_t cap = (size_t)num_sms() * 8;
  if (out_mode == 1 && (ntaps == 5 || ntaps == 1)) {
    const size_t nk = (size_t)N * K;
    const int blocks = (int)std::min<size_t>(ceil_div(nk, (size_t)256), cap);
    if (ntaps == 5) wgrad_reduce_conv_kernel<5><<<blocks, 256, 0, st>>>(part, dW, nk, splits, accumulate);
    else wgrad_reduce_conv_kernel<1><<<blocks, 256, 0, st>>>(part, dW, nk, splits, accumulate);
  } else if (out_mode != 1 && K % 4 == 0 && ((uintptr_t)part & 15) == 0 && ((uintptr_t)dW & 15) == 0) {
    const int blocks = (int)std::min<size_t>(ceil_div(total / 4, (size_t)256), cap);
    wgrad_reduce_vec4_kernel<<<blocks, 256, 0, st>>>((const float4*)part, dW, N, K, ntaps, splits, out_mode, accumulate);
  } else {
    const int blocks = (int)std::min<size_t>(ceil_div(total, (size_t)256), cap);
    wgrad_reduce_kernel<<<blocks, 256, 0, st>>>(part, dW, N, K, ntaps, splits, out_mode, accumulate);
  }
  AVC_LAUNCHED();
  return AVC_OK;
}

int gemm_tn_taps_simt(const float* dY, int ldy, const float* X, int ldx, float* dW, int nB, int T, int N, int K,
                      int ntaps, int shift0, int out_mode, int accumulate, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int M = nB * T;
  const int splits = tn_splits(M, N, K, ntaps);
  const size_t need = (size_t)splits * ntaps * N * K * sizeof(float);
  if (ws == nullptr || ws_bytes < need) {
    set_error("avc_gemm_tn_taps: workspace %zu < %zu", ws_bytes, need);
    return AVC_ERR_WORKSPACE;
  }
  int rps = ceil_div(M, splits);
  rps = ceil_div(rps, SG_BK) * SG_BK;
  dim3 grid(ceil_div(N, SG_BN), ceil_div(K, SG_BN), ntaps * splits);
  gemm_tn_taps_simt_kernel<<<grid, SG_THREADS, 0, st>>>(dY, ldy, X, ldx, (float*)ws, M, T, N, K, ntaps, shift0, splits, rps);
  AVC_LAUNCHED();
  return launch_wgrad_reduce((const float*)ws, dW, N, K, ntaps, splits, out_mode, accumulate, st);
}

// ---------------------------------------------------------------------------------------
// weight packers
// ---------------------------------------------------------------------------------------
// Output element: fmt 0 = fp32, 1 = bf16, 2 = fp16 (the 16-bit forms are what the tensor-core GEMMs read in place).
__device__ __forceinline__ void pack_store(void* base, size_t idx, float v, int fmt) {
  if (fmt == 0) {
    reinterpret_cast<float*>(base)[idx] = v;
  } else if (fmt == 2) {
    const __half h = __float2half_rn(v);
    reinterpret_cast<uint16_t*>(base)[idx] = *reinterpret_cast<const uint16_t*>(&h);
  } else {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    reinterpret_cast<uint16_t*>(base)[idx] = *reinterpret_cast<const uint16_t*>(&h);
  }
}

// Conv1d weight (Cout, Cin, taps) -> fwd [tap][Cout][ldf] and dgrad [taps-1-tap][Cin][ldd] (taps flipped; dX[t, ci] =
// sum_tap' sum_co dY[t + tap' - pad, co] * w[co, ci, taps-1-tap']).  One block moves a 32 co x 32 ci x taps tile through
// shared memory, so the global reads are runs of 32*taps floats and the writes runs of 32 elements; columns between the
// logical width and the leading dimension are zero-filled.
template <int NT_>   // taps known at compile time (0 = runtime value): constant trip counts let the loads be issued back to back
__global__ void __launch_bounds__(256)
pack_conv_tiled_kernel(const float* __restrict__ w, void* __restrict__ wf, int ldf, int fmt_f, void* __restrict__ wd, int ldd,
                       int fmt_d, int Cout, int Cin, int ntaps_rt) {
  extern __shared__ float pk_tile[];                 // [32][32 * ntaps + 1]
  const int ntaps = NT_ > 0 ? NT_ : ntaps_rt;
  const int span = 32 * ntaps, row = span + 1;
  const int ci0 = blockIdx.x * 32, co0 = blockIdx.y * 32;
#pragma unroll 4
  for (int e = threadIdx.x; e < 32 * span; e += 256) {
    const int co_l = e / span, j = e - co_l * span;
    const int co = co0 + co_l, ci = ci0 + j / ntaps;
    pk_tile[co_l * row + j] = (co < Cout && ci < Cin) ? __ldg(w + ((size_t)co * Cin + ci0) * ntaps + j) : 0.f;
  }
  __syncthreads();
#pragma unroll 4
  for (int e = threadIdx.x; e < ntaps * 1024; e += 256) {
    const int a = e & 31, b = (e >> 5) & 31, tap = e >> 10;
    if (wf) {                                        // lanes along ci
      const int ci = ci0 + a, co = co0 + b;
      if (co < Cout && ci < ldf) pack_store(wf, ((size_t)tap * Cout + co) * ldf + ci, pk_tile[b * row + a * ntaps + tap], fmt_f);
    }
    if (wd) {                                        // lanes along co
      const int co = co0 + a, ci = ci0 + b;
      if (ci < Cin && co < ldd)
        pack_store(wd, ((size_t)(ntaps - 1 - tap) * Cin + ci) * ldd + co, pk_tile[a * row + b * ntaps + tap], fmt_d);
    }
  }
}

// nn.LSTM weight (4H, I), rows g*H+u -> gate-interleaved p [u*4+g][ldp] and its transpose pT [k][ldpT]
__global__ void __launch_bounds__(256)
pack_lstm_tiled_kernel(const float* __restrict__ w, void* __restrict__ p, int ldp, int fmt_p, void* __restrict__ pT, int ldpT,
                       int fmt_pT, int H, int I) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.x * 32, pr0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int G = 4 * H;
  for (int i = ty; i < 32; i += 8) {
    const int pr = pr0 + i, k = k0 + tx;
    const float v = (pr < G && k < I) ? w[((size_t)(pr & 3) * H + (pr >> 2)) * I + k] : 0.f;
    tile[i][tx] = v;
    if (p && pr < G && k < ldp) pack_store(p, (size_t)pr * ldp + k, v, fmt_p);
  }
  __syncthreads();
  if (pT) {
    for (int i = ty; i < 32; i += 8) {
      const int k = k0 + i, pr = pr0 + tx;
      if (k < I && pr < ldpT) pack_store(pT, (size_t)k * ldpT + pr, pr < G ? tile[tx][i] : 0.f, fmt_pT);
    }
  }
}

__global__ void pack_lstm_bias_kernel(const float* __restrict__ b_ih, const float* __restrict__ b_hh,
                                      float* __restrict__ out, int H) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 4 * H) {
    const int g = i / H, u = i - g * H;
    out[u * 4 + g] = b_ih[i] + b_hh[i];
  }
}

__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int C) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < C) ? in[(size_t)r * C + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < C) out[(size_t)c * R + r] = tile[threadIdx.x][i];
  }
}

}  // namespace avc

using namespace avc;

static int pack_conv_impl(const float* w, void* wf, int ldf, int fmt_f, void* wd, int ldd, int fmt_d, int Cout, int Cin, int ntaps,
                          cudaStream_t st) {
  const size_t smem = (size_t)32 * (32 * ntaps + 1) * sizeof(float);
  if (smem > 48 * 1024) {
    set_error("avc_pack_conv_weight: %d taps exceed the packer's shared-memory tile", ntaps);
    return AVC_ERR_UNSUPPORTED;
  }
  const dim3 grid(ceil_div(std::max(Cin, ldf), 32), ceil_div(std::max(Cout, ldd), 32));
  if (ntaps == 5) pack_conv_tiled_kernel<5><<<grid, 256, smem, st>>>(w, wf, ldf, fmt_f, wd, ldd, fmt_d, Cout, Cin, ntaps);
  else if (ntaps == 1) pack_conv_tiled_kernel<1><<<grid, 256, smem, st>>>(w, wf, ldf, fmt_f, wd, ldd, fmt_d, Cout, Cin, ntaps);
  else pack_conv_tiled_kernel<0><<<grid, 256, smem, st>>>(w, wf, ldf, fmt_f, wd, ldd, fmt_d, Cout, Cin, ntaps);
  AVC_LAUNCHED();
  return AVC_OK;
}

extern "C" int avc_pack_conv_weight(const float* w, float* w_fwd, float* w_dgrad, int Cout, int Cin, int ntaps,
                                    void* stream) {
  AVC_REQUIRE(w && (w_fwd || w_dgrad) && Cout > 0 && Cin > 0 && ntaps > 0, "avc_pack_conv_weight: bad arguments");
  return pack_conv_impl(w, w_fwd, Cin, 0, w_dgrad, Cout, 0, Cout, Cin, ntaps, as_stream(stream));
}

extern "C" int avc_pack_conv_weight_h(const float* w, void* w_fwd, int ldf, int fmt_f, void* w_dgrad, int ldd, int fmt_d, int Cout,
                                      int Cin, int ntaps, void* stream) {
  AVC_REQUIRE(w && (w_fwd || w_dgrad) && Cout > 0 && Cin > 0 && ntaps > 0, "avc_pack_conv_weight_h: bad arguments");
  AVC_REQUIRE((!w_fwd || ldf >= Cin) && (!w_dgrad || ldd >= Cout) && fmt_f >= 0 && fmt_f <= 2 && fmt_d >= 0 && fmt_d <= 2,
              "avc_pack_conv_weight_h: bad leading dimension or format");
  return pack_conv_impl(w, w_fwd, w_fwd ? ldf : Cin, fmt_f, w_dgrad, w_dgrad ? ldd : Cout, fmt_d, Cout, Cin, ntaps, as_stream(stream));
}

static int pack_lstm_impl(const float* w, void* p, int ldp, int fmt_p, void* pT, int ldpT, int fmt_pT, int H, int I, cudaStream_t st) {
  pack_lstm_tiled_kernel<<<dim3(ceil_div(std::max(I, ldp), 32), ceil_div(std::max(4 * H, ldpT), 32)), 256, 0, st>>>(
      w, p, ldp, fmt_p, pT, ldpT, fmt_pT, H, I);
  AVC_LAUNCHED();
  return AVC_OK;
}

extern "C" int avc_pack_lstm_weight(const float* w, float* out_p, float* out_pT, int H, int I, void* stream) {
  AVC_REQUIRE(w && (out_p || out_pT) && H > 0 && I > 0, "avc_pack_lstm_weight: bad arguments");
  return pack_lstm_impl(w, out_p, I, 0, out_pT, 4 * H, 0, H, I, as_stream(stream));
}

extern "C" int avc_pack_lstm_weight_h(const float* w, void* out_p, int ldp, int fmt_p, void* out_pT, int ldpT, int fmt_pT, int H, int I,
                                      void* stream) {
  AVC_REQUIRE(w && (out_p || out_pT) && H > 0 && I > 0, "avc_pack_lstm_weight_h: bad arguments");
  AVC_REQUIRE((!out_p || ldp >= I) && (!out_pT || ldpT >= 4 * H) && fmt_p >= 0 && fmt_p <= 2 && fmt_pT >= 0 && fmt_pT <= 2,
              "avc_pack_lstm_weight_h: bad leading dimension or format");
  return pack_lstm_impl(w, out_p, out_p ? ldp : I, fmt_p, out_pT, out_pT ? ldpT : 4 * H, fmt_pT, H, I, as_stream(stream));
}

extern "C" int avc_pack_lstm_bias(const float* b_ih, const float* b_hh, float* out, int H, void* stream) {
  AVC_REQUIRE(b_ih && b_hh && out && H > 0, "avc_pack_lstm_bias: bad arguments");
  pack_lstm_bias_kernel<<<ceil_div(4 * H, 256), 256, 0, as_stream(stream)>>>(b_ih, b_hh, out, H);
  AVC_LAUNCHED();
  return AVC_OK;
}

extern "C" int avc_transpose(const float* in, float* out, int R, int C, void* stream) {
  AVC_REQUIRE(in && out && R > 0 && C > 0, "avc_transpose: bad arguments");
  dim3 grid(ceil_div(C, 32), ceil_div(R, 32));
  transpose_kernel<<<grid, dim3(32, 8), 0, as_stream(stream)>>>(in, out, R, C);
  AVC_LAUNCHED();
  return AVC_OK;
}
