// AVC_PREC_FP32X3: fp32-accurate products on the tensor cores ("3xTF32").
//
// Every fp32 operand x is split into hi = x rounded to tf32's 11-bit significand and lo = x - hi (exact in fp32), and
//   a * b  ~=  a_hi*b_hi + a_hi*b_lo + a_lo*b_hi          (the dropped a_lo*b_lo term is <= 2^-22 |a b|)
// is accumulated in fp32 in TMEM by tcgen05.mma kind::tf32.  The three products are ONE GEMM over a three times longer
// reduction, so the existing taps-GEMM kernels (tc_gemm.cu) run unchanged on staged operands:
//   NT (conv forward / data gradient, projections, linear):  A' = [A_hi | A_hi | A_lo] and W' = [W_hi | W_lo | W_hi] along K
//   TN (weight gradients, reduction over rows):              dY' = [dY_hi ; dY_hi ; dY_lo] and X' = [X_hi ; X_lo ; X_hi]
//                                                            stacked along the utterance axis (3*nB utterances)
// Measured against the reference (emulated in the CPU oracle before this was written, then on the GPU goldens): the whole
// training step stays within 3e-5 max-abs of the reference's fp32 outputs -- inside its own 2e-5 distance to fp64.
//
// Recurrences with H > 64 (decoder LSTMs): per time step a split of h_{t-1} (or dG_{t+1}), ONE tensor-core GEMM against
// the pre-split W_hh and an elementwise gate kernel -- instead of a CUDA-core GEMM launch per step.
#include "common.cuh"

namespace avc {

int gemm_nt_taps_tc(const void*, int, int, const void*, int, int, const float*, float*, int, int, int, int, int, int, int, double*, int, int, int, void*, size_t, cudaStream_t, int chunk = 0,
                    int ksplit = 1, size_t c_split_stride = 0);
int gemm_tn_taps_tc(const void*, int, int, const void*, int, int, float*, int, int, int, int, int, int, int, int, int, int, void*, size_t, cudaStream_t, int chunk = 0);
size_t gemm_tn_workspace_tc(int, int, int, int, int, int, int chunk = 0);

static inline size_t align256x(size_t v) { return (v + 255) & ~(size_t)255; }
static inline int round_up_i(int v, int m) { return (v + m - 1) / m * m; }

__device__ __forceinline__ float tf32_hi(float x) {
  uint32_t u = __float_as_uint(x);
  u = (u + 0x0fffu + ((u >> 13) & 1u)) & 0xffffe000u;      // round to nearest even at 13 dropped bits
  return __uint_as_float(u);
}

// dst row r (pitch ldd >= 3C, zero padded) = [hi | hi | lo] (order 0: streamed operand) or [hi | lo | hi] (order 1: weights)
__global__ void __launch_bounds__(256)
split3_k_kernel(const float* __restrict__ src, size_t lds, float* __restrict__ dst, int ldd, size_t R, int C, int order) {
  for (size_t r = blockIdx.x; r < R; r += gridDim.x) {
    const float* s = src + r * lds;
    float* d = dst + r * (size_t)ldd;
    for (int c = threadIdx.x; c < C; c += 256) {
      const float x = s[c];
      const float hi = tf32_hi(x), lo = x - hi;
      d[c] = hi;
      d[C + c] = order ? lo : hi;
      d[2 * C + c] = order ? hi : lo;
    }
    for (int c = 3 * C + threadIdx.x; c < ldd; c += 256) d[c] = 0.f;
  }
}

// dst (3, M, ldd): plane 0 = hi, planes 1/2 = (hi, lo) for order 0, (lo, hi) for order 1; columns >= C zero
__global__ void __launch_bounds__(256)
split3_m_kernel(const float* __restrict__ src, size_t lds, float* __restrict__ dst, int ldd, size_t M, int C, int order) {
  const size_t plane = M * (size_t)ldd;
  for (size_t r = blockIdx.x; r < M; r += gridDim.x) {
    const float* s = src + r * lds;
    float* d = dst + r * (size_t)ldd;
    for (int c = threadIdx.x; c < ldd; c += 256) {
      const float x = c < C ? s[c] : 0.f;
      const float hi = tf32_hi(x), lo = x - hi;
      d[c] = hi;
      d[plane + c] = order ? lo : hi;
      d[2 * plane + c] = order ? hi : lo;
    }
  }
}

static int split_blocks(size_t rows) { return (int)std::min<size_t>(rows, (size_t)num_sms() * 16); }

// ---- NT ------------------------------------------------------------------------------------------------------
size_t gemm_nt_workspace_x3(int nB, int T, int N, int K, int ntaps) {
  const int Kp = round_up_i(3 * K, 32);
  return align256x((size_t)nB * T * Kp * 4) + align256x((size_t)ntaps * N * Kp * 4);
}

int gemm_nt_taps_x3(const float* A, int lda, const float* W, const float* bias, float* C, int ldc, int nB, int T, int N, int K,
                    int ntaps, int shift0, double* stats, int accumulate, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int Kp = round_up_i(3 * K, 32);
  const size_t M = (size_t)nB * T;
  const size_t offW = align256x(M * Kp * 4);
  if (!ws || ws_bytes < gemm_nt_workspace_x3(nB, T, N, K, ntaps)) {
    set_error("avc_gemm_nt_taps(fp32x3): workspace %zu < %zu", ws_bytes, gemm_nt_workspace_x3(nB, T, N, K, ntaps));
    return AVC_ERR_WORKSPACE;
  }
  float* A3 = (float*)ws;
  float* W3 = (float*)((uint8_t*)ws + offW);
  split3_k_kernel<<<split_blocks(M), 256, 0, st>>>(A, (size_t)lda, A3, Kp, M, K, 0);
  AVC_LAUNCHED();
  split3_k_kernel<<<split_blocks((size_t)ntaps * N), 256, 0, st>>>(W, (size_t)K, W3, Kp, (size_t)ntaps * N, K, 1);
  AVC_LAUNCHED();
  return gemm_nt_taps_tc(A3, 0, Kp, W3, 0, Kp, bias, C, ldc, nB, T, N, Kp, ntaps, shift0, stats, accumulate, 4, 1, nullptr, 0, st, 1);
}

// ---- TN ------------------------------------------------------------------------------------------------------
size_t gemm_tn_workspace_x3(int nB, int T, int N, int K, int ntaps) {
  const size_t M = (size_t)nB * T;
  const int Np = round_up_i(N, 4), Kp = round_up_i(K, 4);
  return align256x(3 * M * Np * 4) + align256x(3 * M * Kp * 4) + gemm_tn_workspace_tc(3 * nB, T, N, K, ntaps, 4, 1);
}

int gemm_tn_taps_x3(const float* dY, int ldy, const float* X, int ldx, float* dW, int nB, int T, int N, int K, int ntaps,
                    int shift0, int out_mode, int accumulate, void* ws, size_t ws_bytes, cudaStream_t st) {
  const size_t M = (size_t)nB * T;
  const int Np = round_up_i(N, 4), Kp = round_up_i(K, 4);
  const size_t offX = align256x(3 * M * Np * 4), offI = offX + align256x(3 * M * Kp * 4);
  const size_t need = gemm_tn_workspace_x3(nB, T, N, K, ntaps);
  if (!ws || ws_bytes < need) {
    set_error("avc_gemm_tn_taps(fp32x3): workspace %zu < %zu", ws_bytes, need);
    return AVC_ERR_WORKSPACE;
  }
  float* Y3 = (float*)ws;
  float* X3 = (float*)((uint8_t*)ws + offX);
  split3_m_kernel<<<split_blocks(M), 256, 0, st>>>(dY, (size_t)ldy, Y3, Np, M, N, 0);
  AVC_LAUNCHED();
  split3_m_kernel<<<split_blocks(M), 256, 0, st>>>(X, (size_t)ldx, X3, Kp, M, K, 1);
  AVC_LAUNCHED();
  return gemm_tn_taps_tc(Y3, 0, Np, X3, 0, Kp, dW, 3 * nB, T, N, K, ntaps, shift0, out_mode, accumulate, 4, 1,
                         (uint8_t*)ws + offI, ws_bytes - offI, st, 1);
}

// ---- recurrences (H > 64) ----------------------------------------------------------------------------------------
// Per time step: ONE chunked tensor-core GEMM (the step's B x K' operand against the pre-split W_hh; the reduction is cut into
// `ksplit` work items so that ~all SMs take part: B = 256 rows are only two 128-row tiles) and ONE gate kernel that sums the
// partial products, applies the cell algebra and writes the NEXT step's split operand [hi | hi | lo] in place.
struct X3Split { int tiles, ksplit; };
static X3Split x3_split(int rows, int N, int Kp) {
  X3Split r;
  r.tiles = ceil_div(rows, 128) * ceil_div(N, 128);
  const int kiters = Kp / 32;
  r.ksplit = std::max(1, std::min(num_sms() / std::max(1, r.tiles), kiters / 8));
  return r;
}

// gate algebra of one forward step: pre = P[b, t, :] + sum_s R[s][b, :], rows gate-interleaved (u*4 + g);
// A3 (nB x Kp): the split of h_t for the next step's GEMM
__global__ void __launch_bounds__(256)
lstm_gate_fwd_kernel(const float* __restrict__ P, const float* __restrict__ R, int nsplit, float* __restrict__ h_seq, int ldh,
                     float* __restrict__ gates, float* __restrict__ c_seq, float* __restrict__ A3, int Kp, int nB, int T, int H,
                     int t, int t_prev) {
  const int G = 4 * H;
  const size_t total = (size_t)nB * H;
  const size_t rs = (size_t)nB * G;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / H), u = (int)(i - (size_t)b * H);
    const size_t row = (size_t)b * T + t;
    float4 p = *reinterpret_cast<const float4*>(P + row * G + 4 * u);
    for (int s = 0; s < nsplit; ++s) {
      const float4 r = *reinterpret_cast<const float4*>(R + s * rs + (size_t)b * G + 4 * u);
      p.x += r.x; p.y += r.y; p.z += r.z; p.w += r.w;
    }
    const float gi = sigmoidf_acc(p.x), gf = sigmoidf_acc(p.y), gg = tanhf(p.z), go = sigmoidf_acc(p.w);
    const float cp = (t_prev >= 0) ? c_seq[((size_t)b * T + t_prev) * H + u] : 0.f;
    const float c = gf * cp + gi * gg;
    const float h = go * tanhf(c);
    *reinterpret_cast<float4*>(gates + row * G + 4 * u) = make_float4(gi, gf, gg, go);
    c_seq[row * H + u] = c;
    h_seq[row * ldh + u] = h;
    const float hi = tf32_hi(h);
    float* a = A3 + (size_t)b * Kp;
    a[u] = hi;
    a[H + u] = hi;
    a[2 * H + u] = h - hi;
  }
}

// gate algebra of one BPTT step: dh = dH[b, t, :] + sum_s R[s][b, :] (= dG_{t_next} W_hh); A3 (nB x Kp): the split of dG_t
__global__ void __launch_bounds__(256)
lstm_gate_bwd_kernel(const float* __restrict__ dH, int lddh, const float* __restrict__ R, int nsplit, const float* __restrict__ gates,
                     const float* __restrict__ c_seq, float* __restrict__ dP, float* __restrict__ dc_rec, float* __restrict__ A3,
                     int Kp, int nB, int T, int H, int t, int t_next, int t_prev) {
  const int G = 4 * H;
  const size_t total = (size_t)nB * H;
  const size_t rs = (size_t)nB * H;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / H), u = (int)(i - (size_t)b * H);
    const size_t row = (size_t)b * T + t;
    float dh = dH[row * lddh + u];
    for (int s = 0; s < nsplit; ++s) dh += R[s * rs + (size_t)b * H + u];
    const float4 a = *reinterpret_cast<const float4*>(gates + row * G + 4 * u);
    const float ct = c_seq[row * H + u];
    const float cp = (t_prev >= 0) ? c_seq[((size_t)b * T + t_prev) * H + u] : 0.f;
    const float tc = tanhf(ct);
    const float dc = dh * a.w * (1.f - tc * tc) + ((t_next >= 0) ? dc_rec[(size_t)b * H + u] : 0.f);
    float4 d;
    d.x = dc * a.z * a.x * (1.f - a.x);
    d.y = dc * cp * a.y * (1.f - a.y);
    d.z = dc * a.x * (1.f - a.z * a.z);
    d.w = dh * tc * a.w * (1.f - a.w);
    dc_rec[(size_t)b * H + u] = dc * a.y;
    *reinterpret_cast<float4*>(dP + row * G + 4 * u) = d;
    const float4 hi = make_float4(tf32_hi(d.x), tf32_hi(d.y), tf32_hi(d.z), tf32_hi(d.w));
    float* q = A3 + (size_t)b * Kp + 4 * u;
    *reinterpret_cast<float4*>(q) = hi;
    *reinterpret_cast<float4*>(q + G) = hi;
    *reinterpret_cast<float4*>(q + 2 * G) = make_float4(d.x - hi.x, d.y - hi.y, d.z - hi.z, d.w - hi.w);
  }
}

static int gate_blocks(size_t total) { return (int)std::min<size_t>(ceil_div(total, (size_t)256), (size_t)num_sms() * 8); }

// forward: W3 (4H x Kp, Kp = 3H rounded to 32), A3 (nB x Kp), R (ksplit x nB x 4H)
size_t lstm_fwd_workspace_x3(int nB, int T, int H) {
  const int Kp = round_up_i(3 * H, 32);
  const int S = x3_split(nB, 4 * H, Kp).ksplit;
  return align256x((size_t)4 * H * Kp * 4) + align256x((size_t)nB * Kp * 4) + align256x((size_t)S * nB * 4 * H * 4);
}

int lstm_seq_fwd_x3(const float* P, const float* Whh_p, float* h_seq, int ldh, float* gates, float* c_seq, int nB, int T, int H,
                    int reverse, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (reverse >= 2) {
    set_error("avc_lstm_seq_fwd(fp32x3): reverse=2/3 needs H <= 64");
    return AVC_ERR_UNSUPPORTED;
  }
  if (!ws || ws_bytes < lstm_fwd_workspace_x3(nB, T, H)) {
    set_error("avc_lstm_seq_fwd(fp32x3): workspace too small");
    return AVC_ERR_WORKSPACE;
  }
  const int G = 4 * H, Kp = round_up_i(3 * H, 32);
  const int S = x3_split(nB, G, Kp).ksplit;
  float* W3 = (float*)ws;
  float* A3 = (float*)((uint8_t*)ws + align256x((size_t)G * Kp * 4));
  float* R = (float*)((uint8_t*)A3 + align256x((size_t)nB * Kp * 4));
  split3_k_kernel<<<split_blocks(G), 256, 0, st>>>(Whh_p, (size_t)H, W3, Kp, (size_t)G, H, 1);
  AVC_LAUNCHED();
  AVC_CUDA(cudaMemsetAsync(A3, 0, (size_t)nB * Kp * 4, st));          // the K padding columns stay zero
  for (int step = 0; step < T; ++step) {
    const int t = reverse ? T - 1 - step : step;
    const int t_prev = step == 0 ? -1 : (reverse ? t + 1 : t - 1);
    if (t_prev >= 0) {
      // the utterances of one step form ONE "utterance" of nB frames for the taps-GEMM (no tap shift)
      if (int rc = gemm_nt_taps_tc(A3, 0, Kp, W3, 0, Kp, nullptr, R, G, 1, nB, G, Kp, 1, 0, nullptr, 0, 4, 1, nullptr, 0, st, 1, S,
                                   (size_t)nB * G))
        return rc;
    }
    lstm_gate_fwd_kernel<<<gate_blocks((size_t)nB * H), 256, 0, st>>>(P, R, t_prev >= 0 ? S : 0, h_seq, ldh, gates, c_seq, A3, Kp,
                                                                      nB, T, H, t, t_prev);
    AVC_LAUNCHED();
  }
  return AVC_OK;
}

// BPTT: W3 (H x Kp, Kp = 12H rounded to 32) from Whh_pT (H x 4H), A3 (nB x Kp), R (ksplit x nB x H), dc_rec (nB x H)
size_t lstm_bwd_workspace_x3(int nB, int T, int H) {
  const int Kp = round_up_i(12 * H, 32);
  const int S = x3_split(nB, H, Kp).ksplit;
  return align256x((size_t)H * Kp * 4) + align256x((size_t)nB * Kp * 4) + align256x((size_t)S * nB * H * 4) + align256x((size_t)nB * H * 4);
}

int lstm_seq_bwd_x3(const float* dH, int lddh, const float* Whh_pT, const float* gates, const float* c_seq, float* dP, int nB,
                    int T, int H, int reverse, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (reverse >= 2) {
    set_error("avc_lstm_seq_bwd(fp32x3): reverse=2/3 needs H <= 64");
    return AVC_ERR_UNSUPPORTED;
  }
  if (!ws || ws_bytes < lstm_bwd_workspace_x3(nB, T, H)) {
    set_error("avc_lstm_seq_bwd(fp32x3): workspace too small");
    return AVC_ERR_WORKSPACE;
  }
  const int G = 4 * H, Kp = round_up_i(3 * G, 32);
  const int S = x3_split(nB, H, Kp).ksplit;
  float* W3 = (float*)ws;
  float* A3 = (float*)((uint8_t*)ws + align256x((size_t)H * Kp * 4));
  float* R = (float*)((uint8_t*)A3 + align256x((size_t)nB * Kp * 4));
  float* dc_rec = (float*)((uint8_t*)R + align256x((size_t)S * nB * H * 4));
  split3_k_kernel<<<split_blocks(H), 256, 0, st>>>(Whh_pT, (size_t)G, W3, Kp, (size_t)H, G, 1);
  AVC_LAUNCHED();
  AVC_CUDA(cudaMemsetAsync(A3, 0, (size_t)nB * Kp * 4, st));
  for (int step = T - 1; step >= 0; --step) {
    const int t = reverse ? T - 1 - step : step;
    const int t_next = step == T - 1 ? -1 : (reverse ? t - 1 : t + 1);
    const int t_prev = step == 0 ? -1 : (reverse ? t + 1 : t - 1);
    if (t_next >= 0) {
      if (int rc = gemm_nt_taps_tc(A3, 0, Kp, W3, 0, Kp, nullptr, R, H, 1, nB, H, Kp, 1, 0, nullptr, 0, 4, 1, nullptr, 0, st, 1, S,
                                   (size_t)nB * H))
        return rc;
    }
    lstm_gate_bwd_kernel<<<gate_blocks((size_t)nB * H), 256, 0, st>>>(dH, lddh, R, t_next >= 0 ? S : 0, gates, c_seq, dP, dc_rec, A3,
                                                                      Kp, nB, T, H, t, t_next, t_prev);
    AVC_LAUNCHED();
  }
  return AVC_OK;
}

}  // namespace avc
