// Kernels of the waveform variant (model_vc_wav.py:11-102, solver_encoder.py:264-290) that the GEMM-with-taps family and
// the BatchNorm kernels do not already cover:
//   * PReLU between a convolution and its BatchNorm (model_vc_wav.py:22-25, :44-47) with the BatchNorm's batch statistics
//     taken in the same pass, and its backward (input gradient + the single slope's gradient);
//   * row padding / un-padding of (B, T, C) tensors: the learned analysis filterbank Conv1d(1 -> 512, k = 1024, s = 256)
//     (model_vc_wav.py:18) is a 4-tap convolution over the waveform viewed as (B, L/256, 256) and the synthesis
//     ConvTranspose1d(512 -> 1, k = 1024, s = 256) (:52) is its adjoint, so both run on the taps-GEMM kernels once the
//     frame axis has been brought to a common length;
//   * (A, B, C) -> (B, A, C) permutation (filterbank weights (N, 1, taps*S) <-> packed [tap][N][S]);
//   * the SI-SNR term of solver_encoder.py:276-283 and its gradient;
//   * a whole-tensor sum (gradient of the synthesis layer's single bias).
#include "common.cuh"

namespace avc {

constexpr int PR_COLS = 32, PR_ROWS = 8;

// p = y > 0 ? y : a*y (torch.nn.PReLU, one shared slope); stats[c] += sum p, stats[C + c] += sum p^2
__global__ void __launch_bounds__(PR_COLS* PR_ROWS)
prelu_stats_fwd_kernel(const float* __restrict__ y, const float* __restrict__ slope, float* __restrict__ p, int M, int C,
                       int rows_per_block, double* __restrict__ stats) {
  __shared__ float sa[PR_ROWS][PR_COLS + 1], sb[PR_ROWS][PR_COLS + 1];
  const int c = blockIdx.x * PR_COLS + threadIdx.x;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  const float a = slope[0];
  float s = 0.f, q = 0.f;
  if (c < C) {
    for (int r = r0 + threadIdx.y; r < r1; r += PR_ROWS) {
      const size_t i = (size_t)r * C + c;
      const float v = y[i];
      const float o = v > 0.f ? v : a * v;
      p[i] = o;
      s += o;
      q = fmaf(o, o, q);
    }
  }
  if (!stats) return;
  sa[threadIdx.y][threadIdx.x] = s;
  sb[threadIdx.y][threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    double ds = 0.0, dq = 0.0;
#pragma unroll
    for (int i = 0; i < PR_ROWS; ++i) {
      ds += (double)sa[i][threadIdx.x];
      dq += (double)sb[i][threadIdx.x];
    }
    atomicAdd(stats + c, ds);
    atomicAdd(stats + C + c, dq);
  }
}

// dy = dp * (y > 0 ? 1 : a);  acc[0] += sum_{y <= 0} dp * y   (autograd of torch.prelu)
__global__ void __launch_bounds__(256)
prelu_bwd_kernel(const float* __restrict__ dp, const float* __restrict__ y, const float* __restrict__ slope,
                 float* __restrict__ dy, size_t n, double* __restrict__ acc) {
  const float a = slope[0];
  float s = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float g = dp[i], v = y[i];
    dy[i] = v > 0.f ? g : a * g;
    s += v > 0.f ? 0.f : g * v;
  }
  s = warp_sum(s);
  __shared__ float ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += (double)ws[i];
    atomicAdd(acc, t);
  }
}

__global__ void scalar_finalize_kernel(const double* __restrict__ acc, float* __restrict__ out, int accumulate) {
  const float v = (float)acc[0];
  out[0] = accumulate ? out[0] + v : v;
}

__global__ void __launch_bounds__(256)
sum_all_kernel(const float* __restrict__ x, size_t n, double* __restrict__ acc) {
  float s = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) s += x[i];
  s = warp_sum(s);
  __shared__ float ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += (double)ws[i];
    atomicAdd(acc, t);
  }
}

// dst (nB, Tdst, C) <- src (nB, Tsrc, C): rows t < min(Tsrc, Tdst) copied, the rest of dst zero-filled
__global__ void __launch_bounds__(256)
copy_rows3d_kernel(const float* __restrict__ src, int Tsrc, float* __restrict__ dst, int Tdst, int nB, int C) {
  const int M = nB * Tdst;
  for (int m = blockIdx.x; m < M; m += gridDim.x) {
    const int b = m / Tdst, t = m - b * Tdst;
    float* d = dst + (size_t)m * C;
    if (t < Tsrc) {
      const float* s = src + ((size_t)b * Tsrc + t) * C;
      for (int c = threadIdx.x; c < C; c += 256) d[c] = s[c];
    } else {
      for (int c = threadIdx.x; c < C; c += 256) d[c] = 0.f;
    }
  }
}

// out[b][a][c] = in[a][b][c]
__global__ void __launch_bounds__(256)
permute021_kernel(const float* __restrict__ in, float* __restrict__ out, int A, int B, int C) {
  const int rows = A * B;
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    const int a = r / B, b = r - a * B;
    const float* s = in + (size_t)r * C;
    float* d = out + ((size_t)b * A + a) * C;
    for (int c = threadIdx.x; c < C; c += 256) d[c] = s[c];
  }
}

// ---------------------------------------------------------------------------------------
// SI-SNR, solver_encoder.py:276-283 (est = x_identic, tgt = x_real, both (B, L)):
//   dot = sum est*tgt; en = sum tgt^2; scaled = dot*tgt/en; e = est - scaled; ratio = sum scaled^2 / sum e^2;
//   loss = -mean_b 10*log10(ratio_b)
// pass 1: per-utterance dot and energy; pass 2: per-utterance sum scaled^2 and sum e^2 with the reference's element-wise
// expressions (no algebraic shortcut: sum e^2 = |est|^2 - dot^2/en cancels catastrophically once est ~ tgt).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void block_add2(float a, float b, double* o0, double* o1) {
  a = warp_sum(a);
  b = warp_sum(b);
  __shared__ float wa[8], wb[8];
  if ((threadIdx.x & 31) == 0) {
    wa[threadIdx.x >> 5] = a;
    wb[threadIdx.x >> 5] = b;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0.0, tb = 0.0;
    for (int i = 0; i < 8; ++i) {
      ta += (double)wa[i];
      tb += (double)wb[i];
    }
    atomicAdd(o0, ta);
    atomicAdd(o1, tb);
  }
}

// grid (chunks, B); acc: double[4*B] = {dot, en, sum scaled^2, sum e^2} per utterance
__global__ void __launch_bounds__(256)
sisnr_pass1_kernel(const float* __restrict__ est, const float* __restrict__ tgt, int L, double* __restrict__ acc) {
  const int b = blockIdx.y;
  const float* e = est + (size_t)b * L;
  const float* t = tgt + (size_t)b * L;
  float d = 0.f, en = 0.f;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < L; i += gridDim.x * 256) {
    const float tv = t[i];
    d = fmaf(e[i], tv, d);
    en = fmaf(tv, tv, en);
  }
  block_add2(d, en, acc + 4 * b, acc + 4 * b + 1);
}

__global__ void __launch_bounds__(256)
sisnr_pass2_kernel(const float* __restrict__ est, const float* __restrict__ tgt, int L, double* __restrict__ acc) {
  const int b = blockIdx.y;
  const float* e = est + (size_t)b * L;
  const float* t = tgt + (size_t)b * L;
  const float dot = (float)acc[4 * b], en = (float)acc[4 * b + 1];
  float s2 = 0.f, n2 = 0.f;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < L; i += gridDim.x * 256) {
    const float sc = dot * t[i] / en;
    const float ee = e[i] - sc;
    s2 = fmaf(sc, sc, s2);
    n2 = fmaf(ee, ee, n2);
  }
  block_add2(s2, n2, acc + 4 * b + 2, acc + 4 * b + 3);
}

// saved[b] = {dot, en, A = sum scaled^2, N = sum e^2}; out[0] = -mean_b 10 log10(A/N)
__global__ void sisnr_finalize_kernel(const double* __restrict__ acc, int nB, float* __restrict__ saved, float* __restrict__ out) {
  double s = 0.0;
  for (int b = 0; b < nB; ++b) {
    const float dot = (float)acc[4 * b], en = (float)acc[4 * b + 1], A = (float)acc[4 * b + 2], N = (float)acc[4 * b + 3];
    saved[4 * b] = dot;
    saved[4 * b + 1] = en;
    saved[4 * b + 2] = A;
    saved[4 * b + 3] = N;
    s += (double)(10.f * log10f(A / N));
  }
  out[0] = (float)(-s / (double)nB);
}

// d loss / d est_i = gout * (-10 / (B ln 10)) * (2*alpha*t_i / A - 2*e_i / N), alpha = dot/en.
// (The reference's graph also carries -2*(sum_j e_j t_j)/en * t_i inside dN; sum_j e_j t_j = dot - alpha*en is zero up to
//  rounding and is dropped.)
__global__ void __launch_bounds__(256)
sisnr_bwd_kernel(const float* __restrict__ est, const float* __restrict__ tgt, const float* __restrict__ saved,
                 const float* __restrict__ gout, int nB, int L, float* __restrict__ dest, int accumulate) {
  const int b = blockIdx.y;
  const float dot = saved[4 * b], en = saved[4 * b + 1], A = saved[4 * b + 2], N = saved[4 * b + 3];
  const float k = gout[0] * (-10.f / ((float)nB * 2.302585092994046f));
  const float alpha = dot / en;
  const float ca = 2.f * alpha / A, cn = 2.f / N;
  const float* e = est + (size_t)b * L;
  const float* t = tgt + (size_t)b * L;
  float* d = dest + (size_t)b * L;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < L; i += gridDim.x * 256) {
    const float tv = t[i];
    const float ee = e[i] - dot * tv / en;
    const float g = k * (ca * tv - cn * ee);
    d[i] = accumulate ? d[i] + g : g;
  }
}

static int ew_blocks_w(size_t total) {
  return (int)std::min<size_t>(ceil_div(total, (size_t)256), (size_t)num_sms() * 16);
}

}  // namespace avc

using namespace avc;

extern "C" int avc_prelu_fwd(const float* y, const float* slope, float* p, double* chan_stats, int M, int C, void* stream) {
  AVC_REQUIRE(y && slope && p && M > 0 && C > 0, "avc_prelu_fwd: bad arguments");
  const int cb = ceil_div(C, PR_COLS);
  int rb = std::max(1, ceil_div(8 * num_sms(), cb));
  int rpb = std::max(64, ceil_div(M, rb));
  rb = ceil_div(M, rpb);
  prelu_stats_fwd_kernel<<<dim3(cb, rb), dim3(PR_COLS, PR_ROWS), 0, as_stream(stream)>>>(y, slope, p, M, C, rpb, chan_stats);
  AVC_LAUNCHED();
  return AVC_OK;
}

extern "C" int avc_prelu_bwd(const float* dp, const float* y, const float* slope, float* dy, float* dslope, int accumulate,
                             double* scratch, size_t n, void* stream) {
  AVC_REQUIRE(dp && y && slope && dy && dslope && scratch && n > 0, "avc_prelu_bwd: bad arguments");
  prelu_bwd_kernel<<<ew_blocks_w(n), 256, 0, as_stream(stream)>>>(dp, y, slope, dy, n, scratch);
  AVC_LAUNCHED();
  scalar_finalize_kernel<<<1, 1, 0, as_stream(stream)>>>(scratch, dslope, accumulate);
  AVC_LAUNCHED();
  return AVC_OK;
}

extern "C" int avc_sum_all(const float* x, size_t n, double* scratch, float* out, int accumulate, void* stream) {
  AVC_REQUIRE(x && scratch && out && n > 0, "avc_sum_all: bad arguments");
  sum_all_kernel<<<ew_blocks_w(n), 256, 0, as_stream(stream)>>>(x, n, scratch);
  AVC_LAUNCHED();
  scalar_finalize_kernel<<<1, 1, 0, as_stream(stream)>>>(scratch, out, accumulate);
  AVC_LAUNCHED();
  return AVC_OK;
}

extern "C" int avc_copy_rows3d(const float* src, int Tsrc, float* dst, int Tdst, int nB, int C, void* stream) {
  AVC_REQUIRE(src && dst && Tsrc > 0 && Tdst > 0 && nB > 0 && C > 0, "avc_copy_rows3d: bad arguments");
  const int M = nB * Tdst;
  copy_rows3d_kernel<<<std::min(M, num_sms() * 16), 256, 0, as_stream(stream)>>>(src, Tsrc, dst, Tdst, nB, C);
  AVC_LAUNCHED();
  return AVC_OK;
}

extern "C" int avc_permute021(const float* in, float* out, int A, int B, int C, void* stream) {
  AVC_REQUIRE(in && out && A > 0 && B > 0 && C > 0, "avc_permute021: bad arguments");
  permute021_kernel<<<std::min(A * B, num_sms() * 16), 256, 0, as_stream(stream)>>>(in, out, A, B, C);
  AVC_LAUNCHED();
  return AVC_OK;
}

extern "C" int avc_sisnr_fwd(const float* est, const float* tgt, int nB, int L, double* scratch, float* saved, float* out,
                             void* stream) {
  AVC_REQUIRE(est && tgt && scratch && saved && out && nB > 0 && L > 0, "avc_sisnr_fwd: bad arguments");
  const int chunks = std::max(1, std::min(ceil_div(L, 2048), ceil_div(num_sms() * 8, nB)));
  sisnr_pass1_kernel<<<dim3(chunks, nB), 256, 0, as_stream(stream)>>>(est, tgt, L, scratch);
  AVC_LAUNCHED();
  sisnr_pass2_kernel<<<dim3(chunks, nB), 256, 0, as_stream(stream)>>>(est, tgt, L, scratch);
  AVC_LAUNCHED();
  sisnr_finalize_kernel<<<1, 1, 0, as_stream(stream)>>>(scratch, nB, saved, out);
  AVC_LAUNCHED();
  return AVC_OK;
}

extern "C" int avc_sisnr_bwd(const float* est, const float* tgt, const float* saved, const float* gout, int nB, int L,
                             float* dest, int accumulate, void* stream) {
  AVC_REQUIRE(est && tgt && saved && gout && dest && nB > 0 && L > 0, "avc_sisnr_bwd: bad arguments");
  const int chunks = std::max(1, std::min(ceil_div(L, 2048), ceil_div(num_sms() * 8, nB)));
  sisnr_bwd_kernel<<<dim3(chunks, nB), 256, 0, as_stream(stream)>>>(est, tgt, saved, gout, nB, L, dest, accumulate);
  AVC_LAUNCHED();
  return AVC_OK;
}
