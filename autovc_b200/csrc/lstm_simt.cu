// LSTM recurrences (forward and BPTT), fp32 CUDA-core path.
// nn.LSTM semantics (model_vc_mel.py:61/:73, :90/:111, :104/:118): gates = P_t + h_{t-1} W_hh^T,
// i,f,o = sigmoid, g = tanh, c_t = f c_{t-1} + i g, h_t = o tanh(c_t), h_0 = c_0 = 0.
// Internal column order is gate-interleaved: column u*4+g, g in (i,f,g,o).
//
//  * H <= 64 (encoder BiLSTM, H = dim_neck): one launch runs the whole sequence; W_hh lives in
//    shared memory, each thread owns one gate column of one utterance.
//  * larger H (decoder lstm1/lstm2): one launch per timestep; the recurrent product is the
//    128x128x16 SIMT tile kernel and the gate nonlinearity + cell update are its epilogue (each
//    thread ends up holding complete (i,f,g,o) quads).  The persistent tensor-core recurrence
//    (lstm_tc.cu) replaces this per-step path in AVC_PREC_BF16 mode.
#include "simt_gemm.cuh"

namespace avc {

constexpr int SMALL_H_MAX = 64;

// ---------------------------------------------------------------------------------------
// small H: whole sequence in one kernel
// block = (4H threads) x (UPB utterances)
// ---------------------------------------------------------------------------------------
__global__ void lstm_small_fwd_kernel(const float* __restrict__ P, const float* __restrict__ Whh_p,
                                      float* __restrict__ h_seq, int ldh, float* __restrict__ gates,
                                      float* __restrict__ c_seq, int nB, int T, int H, int reverse) {
  extern __shared__ float sm[];
  const int G = 4 * H;
  int ldp = G;
  if (reverse >= 2) {   // both directions of a BiLSTM in one launch: blockIdx.y = direction, operands stacked [2]
    const int d = blockIdx.y;
    if (reverse == 3) {  // ... except P, which is one (nB, T, 2G) tensor [dir 0 gates | dir 1 gates] (a single N = 2G projection GEMM)
      P += d * G;
      ldp = 2 * G;
    } else {
      P += (size_t)d * nB * T * G;
    }
    Whh_p += (size_t)d * G * H;
    gates += (size_t)d * nB * T * G;
    c_seq += (size_t)d * nB * T * H;
    h_seq += d * H;
    reverse = d;
  }
  float* Ws = sm;                       // [G][H+1]
  float* hs = sm + (size_t)G * (H + 1);  // [UPB][H]
  const int j = threadIdx.x;            // gate column u*4+g
  const int ul = threadIdx.y;           // utterance within block
  const int b = blockIdx.x * blockDim.y + ul;
  const int u = j >> 2, g = j & 3;
  for (int i = threadIdx.y * G + j; i < G * H; i += G * blockDim.y) Ws[(i / H) * (H + 1) + (i % H)] = Whh_p[i];
  if (j < H) hs[ul * H + j] = 0.f;
  __syncthreads();
  float c = 0.f;
  const bool live = b < nB;
  const unsigned quad_base = ((threadIdx.y * blockDim.x + threadIdx.x) & 31) & ~3u;
  float p_next = live ? P[((size_t)b * T + (reverse ? T - 1 : 0)) * ldp + j] : 0.f;
  for (int step = 0; step < T; ++step) {
    const int t = reverse ? (T - 1 - step) : step;
    float acc = p_next;
    if (live && step + 1 < T)       // software prefetch: the next step's pre-activation is independent of the recurrence
      p_next = P[((size_t)b * T + (reverse ? t - 1 : t + 1)) * ldp + j];
    if (live) {
      const float* w = Ws + (size_t)j * (H + 1);
      const float* h = hs + ul * H;
#pragma unroll 8
      for (int k = 0; k < H; ++k) acc = fmaf(w[k], h[k], acc);
    }
    const float a = (g == 2) ? tanhf(acc) : sigmoidf_acc(acc);
    const float gi = __shfl_sync(0xffffffffu, a, quad_base + 0);
    const float gf = __shfl_sync(0xffffffffu, a, quad_base + 1);
    const float gg = __shfl_sync(0xffffffffu, a, quad_base + 2);
    const float go = __shfl_sync(0xffffffffu, a, quad_base + 3);
    c = gf * c + gi * gg;
    const float h_new = go * tanhf(c);
    __syncthreads();  // everyone has read hs
    if (live) {
      gates[((size_t)b * T + t) * G + j] = a;
      if (g == 0) {
        hs[ul * H + u] = h_new;
        h_seq[((size_t)b * T + t) * ldh + u] = h_new;
        c_seq[((size_t)b * T + t) * H + u] = c;
      }
    }
    __syncthreads();
  }
}

__global__ void lstm_small_bwd_kernel(const float* __restrict__ dH, int lddh, const float* __restrict__ Whh_p,
                                      const float* __restrict__ gates, const float* __restrict__ c_seq,
                                      float* __restrict__ dP, int nB, int T, int H, int reverse) {
  extern __shared__ float sm[];
  const int G = 4 * H;
  int ldp = G;
  if (reverse >= 2) {
    const int d = blockIdx.y;
    dH += d * H;
    Whh_p += (size_t)d * G * H;
    gates += (size_t)d * nB * T * G;
    c_seq += (size_t)d * nB * T * H;
    if (reverse == 3) {  // dP is one (nB, T, 2G) tensor [dir 0 | dir 1]: the operand of single N = 2G / K = 2G gradient GEMMs
      dP += d * G;
      ldp = 2 * G;
    } else {
      dP += (size_t)d * nB * T * G;
    }
    reverse = d;
  }
  float* Ws = sm;                              // [G][H+1]   (row j = gate column, col = hidden unit)
  float* dgs = Ws + (size_t)G * (H + 1);       // [UPB][G]   dG of the step processed just before
  const int j = threadIdx.x, ul = threadIdx.y;
  const int b = blockIdx.x * blockDim.y + ul;
  const int u = j >> 2, g = j & 3;
  for (int i = threadIdx.y * G + j; i < G * H; i += G * blockDim.y) Ws[(i / H) * (H + 1) + (i % H)] = Whh_p[i];
  dgs[ul * G + j] = 0.f;
  __syncthreads();
  const bool live = b < nB;
  float dc_rec = 0.f;
  const unsigned quad_base = ((threadIdx.y * blockDim.x + threadIdx.x) & 31) & ~3u;
  // operands of the step about to be processed, fetched one step ahead (they do not depend on the recurrence)
  float a_n = 0.f, ct_n = 0.f, cp_n = 0.f, dH_n = 0.f;
  auto fetch = [&](int step) {
    const int t = reverse ? (T - 1 - step) : step;
    const int t_prev = reverse ? t + 1 : t - 1;
    a_n = gates[((size_t)b * T + t) * G + j];
    ct_n = c_seq[((size_t)b * T + t) * H + u];
    cp_n = (step > 0) ? c_seq[((size_t)b * T + t_prev) * H + u] : 0.f;
    dH_n = dH[((size_t)b * T + t) * lddh + u];
  };
  if (live) fetch(T - 1);
  // BPTT visits timesteps in the opposite order of the forward walk
  for (int step = T - 1; step >= 0; --step) {
    const int t = reverse ? (T - 1 - step) : step;
    const float a = a_n, ct = ct_n, cp = cp_n, dHt = dH_n;
    if (live && step > 0) fetch(step - 1);
    // dh[u] = dH[b,t,u] + sum_j' dG_next[j'] * W[j'][u]: thread (u,g) sums the quarter j' in [g*H, (g+1)*H), quad-reduce
    float part = 0.f;
    {
      const float* dg = dgs + ul * G + g * H;
      const float* w = Ws + (size_t)(g * H) * (H + 1) + u;
#pragma unroll 8
      for (int jj = 0; jj < H; ++jj) part = fmaf(dg[jj], w[(size_t)jj * (H + 1)], part);
    }
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    const float dh = part + dHt;
    const float gi = __shfl_sync(0xffffffffu, a, quad_base + 0);
    const float gf = __shfl_sync(0xffffffffu, a, quad_base + 1);
    const float gg = __shfl_sync(0xffffffffu, a, quad_base + 2);
    const float go = __shfl_sync(0xffffffffu, a, quad_base + 3);
    const float tc = tanhf(ct);
    const float dc = dh * go * (1.f - tc * tc) + dc_rec;
    float d;
    if (g == 0) d = dc * gg * gi * (1.f - gi);
    else if (g == 1) d = dc * cp * gf * (1.f - gf);
    else if (g == 2) d = dc * gi * (1.f - gg * gg);
    else d = dh * tc * go * (1.f - go);
    dc_rec = dc * gf;
    __syncthreads();                 // every thread has read the previous step's dgs
    dgs[ul * G + j] = d;
    if (live) dP[((size_t)b * T + t) * ldp + j] = d;
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------
// H = 16 / 32 (the encoder BiLSTM of the two published configurations): same arithmetic in the same order as the generic
// kernels above, but the thread's W_hh row lives in registers, the recurrent vector is exchanged through a double-buffered
// shared tile read with 128-bit loads, and a step needs ONE block barrier instead of two.
// ---------------------------------------------------------------------------------------
template <int HT>
__global__ void __launch_bounds__(256)
lstm_small_fwd_kernel_t(const float* __restrict__ P, const float* __restrict__ Whh_p, float* __restrict__ h_seq, int ldh,
                        float* __restrict__ gates, float* __restrict__ c_seq, int nB, int T, int reverse) {
  constexpr int H = HT, G = 4 * HT, UPB = 256 / G;
  __shared__ __align__(16) float hs[2][UPB][HT];
  int ldp = G;
  if (reverse >= 2) {
    const int d = blockIdx.y;
    if (reverse == 3) {
      P += d * G;
      ldp = 2 * G;
    } else {
      P += (size_t)d * nB * T * G;
    }
    Whh_p += (size_t)d * G * H;
    gates += (size_t)d * nB * T * G;
    c_seq += (size_t)d * nB * T * H;
    h_seq += d * H;
    reverse = d;
  }
  const int j = threadIdx.x, ul = threadIdx.y;
  const int b = blockIdx.x * UPB + ul;
  const int u = j >> 2, g = j & 3;
  float w[HT];
#pragma unroll
  for (int k = 0; k < HT; ++k) w[k] = Whh_p[(size_t)j * H + k];
  if (j < H) hs[0][ul][j] = 0.f;
  __syncthreads();
  float c = 0.f;
  const bool live = b < nB;
  const unsigned quad_base = ((threadIdx.y * blockDim.x + threadIdx.x) & 31) & ~3u;
  // The step is a short dependent chain (16-32 FMAs, two transcendental calls, one barrier): with one pre-activation fetched one
  // step ahead it waited on the L2 / HBM latency of that load (0.56 us per step measured, r02).  PF steps are kept in flight.
  constexpr int PF = 4;
  float p_ring[PF];
#pragma unroll
  for (int i = 0; i < PF; ++i) p_ring[i] = (live && i < T) ? P[((size_t)b * T + (reverse ? T - 1 - i : i)) * ldp + j] : 0.f;
  for (int step0 = 0; step0 < T; step0 += PF) {
#pragma unroll
    for (int i = 0; i < PF; ++i) {
      const int step = step0 + i;
      if (step >= T) break;                     // uniform over the block
      const int t = reverse ? (T - 1 - step) : step;
      const int cur = step & 1;
      float acc = p_ring[i];
      if (live && step + PF < T) p_ring[i] = P[((size_t)b * T + (reverse ? t - PF : t + PF)) * ldp + j];
      const float4* h4 = reinterpret_cast<const float4*>(hs[cur][ul]);
#pragma unroll
      for (int k = 0; k < HT; k += 4) {
        const float4 hv = h4[k >> 2];
        acc = fmaf(w[k], hv.x, acc);
        acc = fmaf(w[k + 1], hv.y, acc);
        acc = fmaf(w[k + 2], hv.z, acc);
        acc = fmaf(w[k + 3], hv.w, acc);
      }
      const float a = (g == 2) ? tanhf(acc) : sigmoidf_acc(acc);
      const float gi = __shfl_sync(0xffffffffu, a, quad_base + 0);
      const float gf = __shfl_sync(0xffffffffu, a, quad_base + 1);
      const float gg = __shfl_sync(0xffffffffu, a, quad_base + 2);
      const float go = __shfl_sync(0xffffffffu, a, quad_base + 3);
      c = gf * c + gi * gg;
      const float h_new = go * tanhf(c);
      if (g == 0) hs[cur ^ 1][ul][u] = h_new;
      if (live) {
        gates[((size_t)b * T + t) * G + j] = a;
        if (g == 0) {
          h_seq[((size_t)b * T + t) * ldh + u] = h_new;
          c_seq[((size_t)b * T + t) * H + u] = c;
        }
      }
      __syncthreads();
    }
  }
}

template <int HT>
__global__ void __launch_bounds__(256)
lstm_small_bwd_kernel_t(const float* __restrict__ dH, int lddh, const float* __restrict__ Whh_p, const float* __restrict__ gates,
                        const float* __restrict__ c_seq, float* __restrict__ dP, int nB, int T, int reverse) {
  constexpr int H = HT, G = 4 * HT, UPB = 256 / G;
  __shared__ __align__(16) float dgs[2][UPB][G];
  int ldp = G;
  if (reverse >= 2) {
    const int d = blockIdx.y;
    dH += d * H;
    Whh_p += (size_t)d * G * H;
    gates += (size_t)d * nB * T * G;
    c_seq += (size_t)d * nB * T * H;
    if (reverse == 3) {  // dP is one (nB, T, 2G) tensor [dir 0 | dir 1]: the operand of single N = 2G / K = 2G gradient GEMMs
      dP += d * G;
      ldp = 2 * G;
    } else {
      dP += (size_t)d * nB * T * G;
    }
    reverse = d;
  }
  const int j = threadIdx.x, ul = threadIdx.y;
  const int b = blockIdx.x * UPB + ul;
  const int u = j >> 2, g = j & 3;
  float w[HT];                                  // W[g*H + jj][u]: the quarter of the reduction this thread sums
#pragma unroll
  for (int jj = 0; jj < HT; ++jj) w[jj] = Whh_p[(size_t)(g * H + jj) * H + u];
  dgs[0][ul][j] = 0.f;
  __syncthreads();
  const bool live = b < nB;
  float dc_rec = 0.f;
  const unsigned quad_base = ((threadIdx.y * blockDim.x + threadIdx.x) & 31) & ~3u;
  // PF steps of operands in flight (see lstm_small_fwd_kernel_t)
  constexpr int PF = 4;
  float a_r[PF], ct_r[PF], cp_r[PF], dH_r[PF];
  auto fetch = [&](int step, int i) {
    const int t = reverse ? (T - 1 - step) : step;
    const int t_prev = reverse ? t + 1 : t - 1;
    a_r[i] = gates[((size_t)b * T + t) * G + j];
    ct_r[i] = c_seq[((size_t)b * T + t) * H + u];
    cp_r[i] = (step > 0) ? c_seq[((size_t)b * T + t_prev) * H + u] : 0.f;
    dH_r[i] = dH[((size_t)b * T + t) * lddh + u];
  };
#pragma unroll
  for (int i = 0; i < PF; ++i) {
    a_r[i] = ct_r[i] = cp_r[i] = dH_r[i] = 0.f;
    if (live && T - 1 - i >= 0) fetch(T - 1 - i, i);
  }
  int cur = 0;
  for (int step0 = T - 1; step0 >= 0; step0 -= PF) {
#pragma unroll
    for (int i = 0; i < PF; ++i) {
      const int step = step0 - i;
      if (step < 0) break;                      // uniform over the block
      const int t = reverse ? (T - 1 - step) : step;
      const float a = a_r[i], ct = ct_r[i], cp = cp_r[i], dHt = dH_r[i];
      if (live && step - PF >= 0) fetch(step - PF, i);
      float part = 0.f;
      const float4* dg4 = reinterpret_cast<const float4*>(&dgs[cur][ul][g * H]);
#pragma unroll
      for (int jj = 0; jj < HT; jj += 4) {
        const float4 dv = dg4[jj >> 2];
        part = fmaf(dv.x, w[jj], part);
        part = fmaf(dv.y, w[jj + 1], part);
        part = fmaf(dv.z, w[jj + 2], part);
        part = fmaf(dv.w, w[jj + 3], part);
      }
      part += __shfl_xor_sync(0xffffffffu, part, 1);
      part += __shfl_xor_sync(0xffffffffu, part, 2);
      const float dh = part + dHt;
      const float gi = __shfl_sync(0xffffffffu, a, quad_base + 0);
      const float gf = __shfl_sync(0xffffffffu, a, quad_base + 1);
      const float gg = __shfl_sync(0xffffffffu, a, quad_base + 2);
      const float go = __shfl_sync(0xffffffffu, a, quad_base + 3);
      const float tc = tanhf(ct);
      const float dc = dh * go * (1.f - tc * tc) + dc_rec;
      float d;
      if (g == 0) d = dc * gg * gi * (1.f - gi);
      else if (g == 1) d = dc * cp * gf * (1.f - gf);
      else if (g == 2) d = dc * gi * (1.f - gg * gg);
      else d = dh * tc * go * (1.f - go);
      dc_rec = dc * gf;
      dgs[cur ^ 1][ul][j] = d;
      if (live) dP[((size_t)b * T + t) * ldp + j] = d;
      __syncthreads();
      cur ^= 1;
    }
  }
}

// ---------------------------------------------------------------------------------------
// large H: one launch per timestep
// ---------------------------------------------------------------------------------------
// gates tile: rows = utterances, cols = interleaved gate columns
__global__ void __launch_bounds__(SG_THREADS)
lstm_step_fwd_kernel(const float* __restrict__ P, const float* __restrict__ Whh_p, float* __restrict__ h_seq, int ldh,
                     float* __restrict__ gates, float* __restrict__ c_seq, int nB, int T, int H, int t, int t_prev) {
  __shared__ SimtSmem s;
  const int G = 4 * H;
  const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  if (t_prev >= 0) {
    auto loadA = [&](int b, int, int k) -> float {
      return (b < nB && k < H) ? h_seq[((size_t)b * T + t_prev) * ldh + k] : 0.f;
    };
    auto loadB = [&](int n, int, int k) -> float { return (n < G && k < H) ? __ldg(Whh_p + (size_t)n * H + k) : 0.f; };
    simt_mainloop_nt(s, acc, m0, n0, H, 1, loadA, loadB);
  }
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int b = m0 + acc_row(ty, i);
    if (b >= nB) continue;
    const size_t row = (size_t)b * T + t;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int n = n0 + tx * 4 + q * 64;  // first column of the quad
      if (n >= G) continue;
      const int u = n >> 2;
      const float4 p = *reinterpret_cast<const float4*>(P + row * G + n);
      const float gi = sigmoidf_acc(acc[i][q * 4 + 0] + p.x);
      const float gf = sigmoidf_acc(acc[i][q * 4 + 1] + p.y);
      const float gg = tanhf(acc[i][q * 4 + 2] + p.z);
      const float go = sigmoidf_acc(acc[i][q * 4 + 3] + p.w);
      const float cp = (t_prev >= 0) ? c_seq[((size_t)b * T + t_prev) * H + u] : 0.f;
      const float c = gf * cp + gi * gg;
      *reinterpret_cast<float4*>(gates + row * G + n) = make_float4(gi, gf, gg, go);
      c_seq[row * H + u] = c;
      h_seq[row * ldh + u] = go * tanhf(c);
    }
  }
}

// dh tile: rows = utterances, cols = hidden units; K = 4H over dG of the step processed before
__global__ void __launch_bounds__(SG_THREADS)
lstm_step_bwd_kernel(const float* __restrict__ dH, int lddh, const float* __restrict__ Whh_pT,
                     const float* __restrict__ gates, const float* __restrict__ c_seq, float* __restrict__ dP,
                     float* __restrict__ dc_rec, int nB, int T, int H, int t, int t_next, int t_prev) {
  __shared__ SimtSmem s;
  const int G = 4 * H;
  const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  if (t_next >= 0) {
    auto loadA = [&](int b, int, int k) -> float {
      return (b < nB && k < G) ? dP[((size_t)b * T + t_next) * G + k] : 0.f;
    };
    auto loadB = [&](int n, int, int k) -> float { return (n < H && k < G) ? __ldg(Whh_pT + (size_t)n * G + k) : 0.f; };
    simt_mainloop_nt(s, acc, m0, n0, G, 1, loadA, loadB);
  }
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int b = m0 + acc_row(ty, i);
    if (b >= nB) continue;
    const size_t row = (size_t)b * T + t;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int u = n0 + acc_col(tx, j);
      if (u >= H) continue;
      const float dh = acc[i][j] + dH[row * lddh + u];
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
    }
  }
}

int lstm_seq_fwd_simt(const float* P, const float* Whh_p, float* h_seq, int ldh, float* gates, float* c_seq, int nB,
                      int T, int H, int reverse, cudaStream_t st) {
  const int G = 4 * H;
  if (H == 16 || H == 32) {
    const int upb = 256 / G;
    const dim3 grid(ceil_div(nB, upb), reverse >= 2 ? 2 : 1), block(G, upb);
    if (H == 16) lstm_small_fwd_kernel_t<16><<<grid, block, 0, st>>>(P, Whh_p, h_seq, ldh, gates, c_seq, nB, T, reverse);
    else lstm_small_fwd_kernel_t<32><<<grid, block, 0, st>>>(P, Whh_p, h_seq, ldh, gates, c_seq, nB, T, reverse);
    AVC_LAUNCHED();
    return AVC_OK;
  }
  if (H <= SMALL_H_MAX && H % 8 == 0) {
    int upb = std::max(1, 256 / G);
    const size_t smem = ((size_t)G * (H + 1) + (size_t)upb * H) * sizeof(float);
    if (smem > 48 * 1024)
      AVC_CUDA(cudaFuncSetAttribute(lstm_small_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lstm_small_fwd_kernel<<<dim3(ceil_div(nB, upb), reverse >= 2 ? 2 : 1), dim3(G, upb), smem, st>>>(P, Whh_p, h_seq, ldh, gates, c_seq, nB, T, H, reverse);
    AVC_LAUNCHED();
    return AVC_OK;
  }
  if (reverse >= 2) {
    set_error("avc_lstm_seq_fwd: reverse=2/3 (both directions in one launch) needs H <= 64, H %% 8 == 0");
    return AVC_ERR_UNSUPPORTED;
  }
  dim3 grid(ceil_div(G, SG_BN), ceil_div(nB, SG_BM));
  for (int step = 0; step < T; ++step) {
    const int t = reverse ? T - 1 - step : step;
    const int t_prev = step == 0 ? -1 : (reverse ? t + 1 : t - 1);
    lstm_step_fwd_kernel<<<grid, SG_THREADS, 0, st>>>(P, Whh_p, h_seq, ldh, gates, c_seq, nB, T, H, t, t_prev);
    AVC_LAUNCHED();
  }
  return AVC_OK;
}

size_t lstm_bwd_workspace_simt(int nB, int T, int H) { return (size_t)nB * H * sizeof(float); }

int lstm_seq_bwd_simt(const float* dH, int lddh, const float* Whh_p, const float* Whh_pT, const float* gates,
                      const float* c_seq, float* dP, int nB, int T, int H, int reverse, void* ws, size_t ws_bytes,
                      cudaStream_t st) {
  const int G = 4 * H;
  if (H == 16 || H == 32) {
    const int upb = 256 / G;
    const dim3 grid(ceil_div(nB, upb), reverse >= 2 ? 2 : 1), block(G, upb);
    if (H == 16) lstm_small_bwd_kernel_t<16><<<grid, block, 0, st>>>(dH, lddh, Whh_p, gates, c_seq, dP, nB, T, reverse);
    else lstm_small_bwd_kernel_t<32><<<grid, block, 0, st>>>(dH, lddh, Whh_p, gates, c_seq, dP, nB, T, reverse);
    AVC_LAUNCHED();
    return AVC_OK;
  }
  if (H <= SMALL_H_MAX && H % 8 == 0) {
    int upb = std::max(1, 256 / G);
    const size_t smem = ((size_t)G * (H + 1) + (size_t)upb * G + (size_t)upb * H) * sizeof(float);
    if (smem > 48 * 1024)
      AVC_CUDA(cudaFuncSetAttribute(lstm_small_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lstm_small_bwd_kernel<<<dim3(ceil_div(nB, upb), reverse >= 2 ? 2 : 1), dim3(G, upb), smem, st>>>(dH, lddh, Whh_p, gates, c_seq, dP, nB, T, H, reverse);
    AVC_LAUNCHED();
    return AVC_OK;
  }
  if (reverse >= 2) {
    set_error("avc_lstm_seq_bwd: reverse=2/3 (both directions in one launch) needs H <= 64, H %% 8 == 0");
    return AVC_ERR_UNSUPPORTED;
  }
  if (!ws || ws_bytes < lstm_bwd_workspace_simt(nB, T, H)) {
    set_error("avc_lstm_seq_bwd: workspace too small");
    return AVC_ERR_WORKSPACE;
  }
  dim3 grid(ceil_div(H, SG_BN), ceil_div(nB, SG_BM));
  for (int step = T - 1; step >= 0; --step) {
    const int t = reverse ? T - 1 - step : step;
    const int t_next = step == T - 1 ? -1 : (reverse ? t - 1 : t + 1);  // processed after t in the forward walk
    const int t_prev = step == 0 ? -1 : (reverse ? t + 1 : t - 1);
    lstm_step_bwd_kernel<<<grid, SG_THREADS, 0, st>>>(dH, lddh, Whh_pT, gates, c_seq, dP, (float*)ws, nB, T, H, t, t_next, t_prev);
    AVC_LAUNCHED();
  }
  return AVC_OK;
}

}  // namespace avc
