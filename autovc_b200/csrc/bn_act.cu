// Train/eval BatchNorm1d + activation (+ residual) forward/backward, channel reductions, losses.
// Reference call sites: nn.BatchNorm1d model_vc_mel.py:57,:101,:139,:150,:160; F.relu :69,:115;
// torch.tanh :165; residual :197; F.mse_loss / F.l1_loss solver_encoder.py:230,:233,:236.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace avc {

// 16-bit stores for the "half" precision mode (fmt 1 = bf16, 2 = fp16): four values -> 8 bytes
__device__ __forceinline__ uint2 pack4_16(float a, float b, float c, float d, int fmt) {
  uint2 r;
  if (fmt == 2) {
    const __half2 lo = __floats2half2_rn(a, b), hi = __floats2half2_rn(c, d);
    r.x = *reinterpret_cast<const uint32_t*>(&lo);
    r.y = *reinterpret_cast<const uint32_t*>(&hi);
  } else {
    const __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
    r.x = *reinterpret_cast<const uint32_t*>(&lo);
    r.y = *reinterpret_cast<const uint32_t*>(&hi);
  }
  return r;
}
__device__ __forceinline__ uint16_t cvt1_16(float a, int fmt) {
  if (fmt == 2) {
    const __half h = __float2half_rn(a);
    return *reinterpret_cast<const uint16_t*>(&h);
  }
  const __nv_bfloat16 h = __float2bfloat16_rn(a);
  return *reinterpret_cast<const uint16_t*>(&h);
}

// dst16 (M, C) ld ldd  <-  src fp32 (M, C) ld lds
__global__ void cast16_kernel(const float* __restrict__ src, int lds, uint16_t* __restrict__ dst, int ldd, size_t M, int C, int fmt) {
  const size_t total = M * (size_t)C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / C;
    const int c = (int)(i % C);
    dst[r * ldd + c] = cvt1_16(src[r * lds + c], fmt);
  }
}
__global__ void cast16_vec4_kernel(const float4* __restrict__ src, uint2* __restrict__ dst, size_t total4, int fmt) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = src[i];
    dst[i] = pack4_16(v.x, v.y, v.z, v.w, fmt);
  }
}

// ---------------------------------------------------------------------------------------
// column reductions over (M, C): block = 32 columns x 8 row-lanes; fp32 per-thread partials over a
// bounded row chunk, fp64 across chunks (one atomic per column per block).
// ---------------------------------------------------------------------------------------
constexpr int CR_COLS = 32, CR_ROWS = 8;

template <int MODE>  // 0: sum & sumsq of x;  1: BN backward sums of g and g*xhat
__global__ void __launch_bounds__(CR_COLS* CR_ROWS)
col_reduce_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ z, const float* __restrict__ y,
                  const float* __restrict__ mean, const float* __restrict__ rstd, int M, int C, int rows_per_block,
                  int act, double* __restrict__ out) {
  __shared__ float sa[CR_ROWS][CR_COLS + 1], sb[CR_ROWS][CR_COLS + 1];
  const int c = blockIdx.x * CR_COLS + threadIdx.x;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  float a = 0.f, b = 0.f;
  if (c < C) {
    float mu = 0.f, rs = 0.f;
    if (MODE == 1) {
      mu = mean[c];
      rs = rstd[c];
    }
    if (MODE == 0) {
      // four independent row streams per thread keep enough loads in flight to approach HBM speed
      float a1 = 0.f, b1 = 0.f, a2 = 0.f, b2 = 0.f, a3 = 0.f, b3 = 0.f;
      int r = r0 + threadIdx.y;
      for (; r + 3 * CR_ROWS < r1; r += 4 * CR_ROWS) {
        const float v0 = x[(size_t)r * ldx + c];
        const float v1 = x[(size_t)(r + CR_ROWS) * ldx + c];
        const float v2 = x[(size_t)(r + 2 * CR_ROWS) * ldx + c];
        const float v3 = x[(size_t)(r + 3 * CR_ROWS) * ldx + c];
        a += v0; b = fmaf(v0, v0, b);
        a1 += v1; b1 = fmaf(v1, v1, b1);
        a2 += v2; b2 = fmaf(v2, v2, b2);
        a3 += v3; b3 = fmaf(v3, v3, b3);
      }
      for (; r < r1; r += CR_ROWS) {
        const float v = x[(size_t)r * ldx + c];
        a += v;
        b = fmaf(v, v, b);
      }
      a += a1 + a2 + a3;
      b += b1 + b2 + b3;
    } else {
      for (int r = r0 + threadIdx.y; r < r1; r += CR_ROWS) {
        const size_t i = (size_t)r * C + c;
        float g = x[i];  // dz
        if (act == AVC_ACT_RELU) g = z[i] > 0.f ? g : 0.f;
        else if (act == AVC_ACT_TANH) g *= (1.f - z[i] * z[i]);
        const float xh = (y[i] - mu) * rs;
        a += g;
        b = fmaf(g, xh, b);
      }
    }
  }
  sa[threadIdx.y][threadIdx.x] = a;
  sb[threadIdx.y][threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    double da = 0.0, db = 0.0;
#pragma unroll
    for (int i = 0; i < CR_ROWS; ++i) {
      da += (double)sa[i][threadIdx.x];
      db += (double)sb[i][threadIdx.x];
    }
    atomicAdd(out + c, da);
    atomicAdd(out + C + c, db);
  }
}

static void col_reduce_grid(int M, int C, dim3& grid, int& rows_per_block) {
  const int cb = ceil_div(C, CR_COLS);
  int rb = ceil_div(8 * num_sms(), cb);
  rows_per_block = ceil_div(M, rb);
  if (rows_per_block < 64) rows_per_block = 64;
  rb = ceil_div(M, rows_per_block);
  grid = dim3(cb, rb);
}

__global__ void bn_finalize_kernel(const double* __restrict__ stats, int M, int C, float eps, float momentum,
                                   float* __restrict__ mean, float* __restrict__ rstd, float* __restrict__ rmean,
                                   float* __restrict__ rvar) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mu = stats[c] / M;
  double var = stats[C + c] / M - mu * mu;
  if (var < 0.0) var = 0.0;
  mean[c] = (float)mu;
  rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (rmean) rmean[c] = (1.f - momentum) * rmean[c] + momentum * (float)mu;
  if (rvar) {
    const double unbiased = M > 1 ? var * ((double)M / (double)(M - 1)) : var;
    rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)unbiased;
  }
}

__global__ void bn_eval_stats_kernel(const float* __restrict__ rmean, const float* __restrict__ rvar, int C, float eps,
                                     float* __restrict__ mean, float* __restrict__ rstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  mean[c] = rmean[c];
  rstd[c] = 1.0f / sqrtf(rvar[c] + eps);
}

// the one expression every forward AND every recomputing backward kernel uses for the BatchNorm affine map
__device__ __forceinline__ float bn_affine(float y, float mu, float sc, float be) { return fmaf(y - mu, sc, be); }

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == AVC_ACT_RELU) return fmaxf(v, 0.f);
  if (act == AVC_ACT_TANH) return tanhf(v);
  return v;
}

// z = act(y*scale + shift) (+ residual); grid-stride over (M, C) with per-element channel lookup
__global__ void bn_act_fwd_kernel(const float* __restrict__ y, const float* __restrict__ mean,
                                  const float* __restrict__ rstd, const float* __restrict__ gamma,
                                  const float* __restrict__ beta, const float* __restrict__ res, float* __restrict__ z,
                                  size_t total, int C, int act) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const float sc = rstd[c] * gamma[c];
    float v = apply_act(bn_affine(y[i], mean[c], sc, beta[c]), act);
    if (res) v += res[i];
    z[i] = v;
  }
}

__global__ void bn_act_bwd_apply_kernel(const float* __restrict__ dz, const float* __restrict__ z,
                                        const float* __restrict__ y, const float* __restrict__ mean,
                                        const float* __restrict__ rstd, const float* __restrict__ gamma,
                                        const double* __restrict__ sums, float* __restrict__ dy, size_t total, int M,
                                        int C, int act) {
  const float invM = 1.0f / (float)M;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    float g = dz[i];
    if (act == AVC_ACT_RELU) g = z[i] > 0.f ? g : 0.f;
    else if (act == AVC_ACT_TANH) g *= (1.f - z[i] * z[i]);
    const float rs = rstd[c];
    const float xh = (y[i] - mean[c]) * rs;
    const float sg = (float)sums[c] * invM, sgx = (float)sums[C + c] * invM;
    dy[i] = gamma[c] * rs * (g - sg - xh * sgx);
  }
}

// ---------------------------------------------------------------------------------------
// Column-fixed variants (C % 4 == 0): block = 32 channel quads x 8 row lanes, each thread keeps the parameters of
// ITS four channels in registers and walks down the rows with four independent row streams in flight.  The
// grid-stride kernels above re-load 5 parameters per element and are L1-bound (r01b ncu: l1tex 83-90% busy, DRAM
// 31-41%); these are bound by HBM.  With z == nullptr the activation is recomputed from y (bitwise the forward's
// expression), which saves a whole (M, C) read in the two backward passes.
// ---------------------------------------------------------------------------------------
constexpr int BC_UNROLL = 4;

__global__ void __launch_bounds__(256)
bn_act_fwd_cols_kernel(const float4* __restrict__ y, const float* __restrict__ mean, const float* __restrict__ rstd,
                       const float* __restrict__ gamma, const float* __restrict__ beta, const float4* __restrict__ res,
                       float4* __restrict__ z, uint2* __restrict__ z16, int fmt16, uint2* __restrict__ z16b, int fmt16b, int M, int C,
                       int rows_per_block, int act) {
  const int cq = blockIdx.x * 32 + threadIdx.x;
  const int c = cq * 4;
  if (c >= C) return;
  float mu[4], sc[4], be[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    mu[j] = mean[c + j];
    sc[j] = rstd[c + j] * gamma[c + j];
    be[j] = beta[c + j];
  }
  const int C4 = C >> 2;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
  for (int r = r0 + threadIdx.y; r < r1; r += 8 * BC_UNROLL) {
    float4 v[BC_UNROLL], rr[BC_UNROLL];
#pragma unroll
    for (int u = 0; u < BC_UNROLL; ++u) {
      const int ru = r + 8 * u;
      if (ru < r1) {
        v[u] = y[(size_t)ru * C4 + cq];
        if (res) rr[u] = res[(size_t)ru * C4 + cq];
      }
    }
#pragma unroll
    for (int u = 0; u < BC_UNROLL; ++u) {
      const int ru = r + 8 * u;
      if (ru < r1) {
        float o[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = apply_act(bn_affine(o[j], mu[j], sc[j], be[j]), act);
        if (res) {
          o[0] += rr[u].x; o[1] += rr[u].y; o[2] += rr[u].z; o[3] += rr[u].w;
        }
        const size_t i = (size_t)ru * C4 + cq;
        if (z) z[i] = make_float4(o[0], o[1], o[2], o[3]);
        if (z16) z16[i] = pack4_16(o[0], o[1], o[2], o[3], fmt16);
        if (z16b) z16b[i] = pack4_16(o[0], o[1], o[2], o[3], fmt16b);
      }
    }
  }
}

// g = dz * act'(z) for four channels; z given, or recomputed from y when zv == nullptr
__device__ __forceinline__ void bn_bwd_g4(const float4& d4, const float4& y4, const float4* zv, const float (&mu)[4],
                                          const float (&rs)[4], const float (&sc)[4], const float (&be)[4], int act,
                                          float (&g)[4], float (&xh)[4]) {
  const float dv[4] = {d4.x, d4.y, d4.z, d4.w}, yv[4] = {y4.x, y4.y, y4.z, y4.w};
  float zz[4] = {0.f, 0.f, 0.f, 0.f};
  if (act != AVC_ACT_NONE) {
    if (zv) {
      zz[0] = zv->x; zz[1] = zv->y; zz[2] = zv->z; zz[3] = zv->w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) zz[j] = apply_act(bn_affine(yv[j], mu[j], sc[j], be[j]), act);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float gg = dv[j];
    if (act == AVC_ACT_RELU) gg = zz[j] > 0.f ? gg : 0.f;
    else if (act == AVC_ACT_TANH) gg *= (1.f - zz[j] * zz[j]);
    g[j] = gg;
    xh[j] = (yv[j] - mu[j]) * rs[j];
  }
}

__global__ void __launch_bounds__(256)
bn_bwd_reduce_cols_kernel(const float4* __restrict__ dz, const float4* __restrict__ z, const float4* __restrict__ y,
                          const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                          const float* __restrict__ beta, int M, int C, int rows_per_block, int act, double* __restrict__ out) {
  __shared__ float sa[8][128 + 4], sb[8][128 + 4];
  const int cq = blockIdx.x * 32 + threadIdx.x;
  const int c = cq * 4;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
  float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
  if (c < C) {
    float mu[4], rs[4], sc[4] = {0.f, 0.f, 0.f, 0.f}, be[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      mu[j] = mean[c + j];
      rs[j] = rstd[c + j];
      if (!z && act != AVC_ACT_NONE) {
        sc[j] = rs[j] * gamma[c + j];
        be[j] = beta[c + j];
      }
    }
    const int C4 = C >> 2;
    const bool ldz = z != nullptr && act != AVC_ACT_NONE;
    for (int r = r0 + threadIdx.y; r < r1; r += 8 * BC_UNROLL) {
      float4 d4[BC_UNROLL], y4[BC_UNROLL], z4[BC_UNROLL];
#pragma unroll
      for (int u = 0; u < BC_UNROLL; ++u) {
        const int ru = r + 8 * u;
        if (ru < r1) {
          const size_t i = (size_t)ru * C4 + cq;
          d4[u] = dz[i];
          y4[u] = y[i];
          if (ldz) z4[u] = z[i];
        }
      }
#pragma unroll
      for (int u = 0; u < BC_UNROLL; ++u) {
        if (r + 8 * u < r1) {
          float g[4], xh[4];
          bn_bwd_g4(d4[u], y4[u], ldz ? &z4[u] : nullptr, mu, rs, sc, be, act, g, xh);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            a[j] += g[j];
            b[j] = fmaf(g[j], xh[j], b[j]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    sa[threadIdx.y][threadIdx.x * 4 + j] = a[j];
    sb[threadIdx.y][threadIdx.x * 4 + j] = b[j];
  }
  __syncthreads();
  const int t = threadIdx.y * 32 + threadIdx.x;
  if (t < 128) {
    const int cc = blockIdx.x * 128 + t;
    if (cc < C) {
      double da = 0.0, db = 0.0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        da += (double)sa[i][t];
        db += (double)sb[i][t];
      }
      atomicAdd(out + cc, da);
      atomicAdd(out + C + cc, db);
    }
  }
}

__global__ void __launch_bounds__(256)
bn_bwd_apply_cols_kernel(const float4* __restrict__ dz, const float4* __restrict__ z, const float4* __restrict__ y,
                         const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                         const float* __restrict__ beta, const double* __restrict__ sums, float4* __restrict__ dy,
                         uint2* __restrict__ dy16, int fmt16, int M, int C, int rows_per_block, int act) {
  const int cq = blockIdx.x * 32 + threadIdx.x;
  const int c = cq * 4;
  if (c >= C) return;
  const float invM = 1.0f / (float)M;
  float mu[4], rs[4], sc[4] = {0.f, 0.f, 0.f, 0.f}, be[4] = {0.f, 0.f, 0.f, 0.f}, gr[4], sg[4], sgx[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    mu[j] = mean[c + j];
    rs[j] = rstd[c + j];
    gr[j] = gamma[c + j] * rs[j];
    sg[j] = (float)sums[c + j] * invM;
    sgx[j] = (float)sums[C + c + j] * invM;
    if (!z && act != AVC_ACT_NONE) {
      sc[j] = rs[j] * gamma[c + j];
      be[j] = beta[c + j];
    }
  }
  const int C4 = C >> 2;
  const bool ldz = z != nullptr && act != AVC_ACT_NONE;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
  for (int r = r0 + threadIdx.y; r < r1; r += 8 * BC_UNROLL) {
    float4 d4[BC_UNROLL], y4[BC_UNROLL], z4[BC_UNROLL];
#pragma unroll
    for (int u = 0; u < BC_UNROLL; ++u) {
      const int ru = r + 8 * u;
      if (ru < r1) {
        const size_t i = (size_t)ru * C4 + cq;
        d4[u] = dz[i];
        y4[u] = y[i];
        if (ldz) z4[u] = z[i];
      }
    }
#pragma unroll
    for (int u = 0; u < BC_UNROLL; ++u) {
      const int ru = r + 8 * u;
      if (ru < r1) {
        float g[4], xh[4], o[4];
        bn_bwd_g4(d4[u], y4[u], ldz ? &z4[u] : nullptr, mu, rs, sc, be, act, g, xh);
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = gr[j] * (g[j] - sg[j] - xh[j] * sgx[j]);
        const size_t i = (size_t)ru * C4 + cq;
        if (dy) dy[i] = make_float4(o[0], o[1], o[2], o[3]);
        if (dy16) dy16[i] = pack4_16(o[0], o[1], o[2], o[3], fmt16);
      }
    }
  }
}

// sum (and optionally sum of squares) of each column of x (M, C), row stride ldx4 float4s
template <bool SQ>
__global__ void __launch_bounds__(256)
col_sums_cols_kernel(const float4* __restrict__ x, int ldx4, int M, int C, int rows_per_block, double* __restrict__ out) {
  __shared__ float sa[8][128 + 4], sb[SQ ? 8 : 1][128 + 4];
  const int cq = blockIdx.x * 32 + threadIdx.x;
  const int c = cq * 4;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
  float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
  if (c < C) {
    for (int r = r0 + threadIdx.y; r < r1; r += 8 * BC_UNROLL) {
      float4 v[BC_UNROLL];
#pragma unroll
      for (int u = 0; u < BC_UNROLL; ++u)
        v[u] = (r + 8 * u < r1) ? x[(size_t)(r + 8 * u) * ldx4 + cq] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < BC_UNROLL; ++u) {
        a[0] += v[u].x; a[1] += v[u].y; a[2] += v[u].z; a[3] += v[u].w;
        if (SQ) {
          b[0] = fmaf(v[u].x, v[u].x, b[0]); b[1] = fmaf(v[u].y, v[u].y, b[1]);
          b[2] = fmaf(v[u].z, v[u].z, b[2]); b[3] = fmaf(v[u].w, v[u].w, b[3]);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    sa[threadIdx.y][threadIdx.x * 4 + j] = a[j];
    if (SQ) sb[threadIdx.y][threadIdx.x * 4 + j] = b[j];
  }
  __syncthreads();
  const int t = threadIdx.y * 32 + threadIdx.x;
  if (t < 128) {
    const int cc = blockIdx.x * 128 + t;
    if (cc < C) {
      double da = 0.0, db = 0.0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        da += (double)sa[i][t];
        if (SQ) db += (double)sb[i][t];
      }
      atomicAdd(out + cc, da);
      if (SQ) atomicAdd(out + C + cc, db);
    }
  }
}

// column sums of a 16-bit (bf16 / fp16) matrix, row stride ldx4 groups of four elements
__global__ void __launch_bounds__(256)
col_sums16_cols_kernel(const uint2* __restrict__ x, int ldx4, int M, int C, int rows_per_block, int fmt, double* __restrict__ out) {
  __shared__ float sa[8][128 + 4];
  const int cq = blockIdx.x * 32 + threadIdx.x;
  const int c = cq * 4;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
  float a[4] = {0.f, 0.f, 0.f, 0.f};
  if (c < C) {
    for (int r = r0 + threadIdx.y; r < r1; r += 8 * BC_UNROLL) {
      uint2 v[BC_UNROLL];
#pragma unroll
      for (int u = 0; u < BC_UNROLL; ++u) v[u] = (r + 8 * u < r1) ? x[(size_t)(r + 8 * u) * ldx4 + cq] : make_uint2(0u, 0u);
#pragma unroll
      for (int u = 0; u < BC_UNROLL; ++u) {
        float2 lo, hi;
        if (fmt == 2) {
          lo = __half22float2(*reinterpret_cast<const __half2*>(&v[u].x));
          hi = __half22float2(*reinterpret_cast<const __half2*>(&v[u].y));
        } else {
          lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v[u].x));
          hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v[u].y));
        }
        a[0] += lo.x; a[1] += lo.y; a[2] += hi.x; a[3] += hi.y;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) sa[threadIdx.y][threadIdx.x * 4 + j] = a[j];
  __syncthreads();
  const int t = threadIdx.y * 32 + threadIdx.x;
  if (t < 128) {
    const int cc = blockIdx.x * 128 + t;
    if (cc < C) {
      double da = 0.0;
#pragma unroll
      for (int i = 0; i < 8; ++i) da += (double)sa[i][t];
      atomicAdd(out + cc, da);
    }
  }
}

// grid for the column-fixed kernels: ceil(C/128) column blocks x enough row blocks for ~6 blocks per SM
static void cols_grid(int M, int C, dim3& grid, int& rows_per_block) {
  const int cb = ceil_div(C, 128);
  int rb = std::max(1, ceil_div(6 * num_sms(), cb));
  rows_per_block = std::max(8 * BC_UNROLL, ceil_div(ceil_div(M, rb), 8) * 8);
  rb = ceil_div(M, rows_per_block);
  grid = dim3(cb, rb);
}

__global__ void bn_param_grad_kernel(const double* __restrict__ sums, float* __restrict__ dgamma,
                                     float* __restrict__ dbeta, int C, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float dg = (float)sums[C + c], db = (float)sums[c];
  if (dgamma) dgamma[c] = accumulate ? dgamma[c] + dg : dg;
  if (dbeta) dbeta[c] = accumulate ? dbeta[c] + db : db;
}

__global__ void colsum_finalize_kernel(const double* __restrict__ sums, float* __restrict__ out, float* __restrict__ out2,
                                       int C, int out_mode, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  int o = c;
  if (out_mode == 2) {
    const int H = C >> 2;
    o = (c & 3) * H + (c >> 2);
  }
  const float v = (float)sums[c];
  out[o] = accumulate ? out[o] + v : v;
  if (out2) out2[o] = accumulate ? out2[o] + v : v;
}

// ---------------------------------------------------------------------------------------
// losses
// ---------------------------------------------------------------------------------------
template <bool L1>
__global__ void loss_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, size_t n,
                                double* __restrict__ acc) {
  float s = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float d = a[i] - b[i];
    s += L1 ? fabsf(d) : d * d;
  }
  s = warp_sum(s);
  __shared__ float ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += (double)ws[i];
    atomicAdd(acc, t);
  }
}
__global__ void loss_finalize_kernel(const double* __restrict__ acc, size_t n, float* __restrict__ out) {
  out[0] = (float)(acc[0] / (double)n);
}
__global__ void loss_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, size_t n,
                                const float* __restrict__ gout, int is_l1, float* __restrict__ da,
                                float* __restrict__ db, int accumulate) {
  const float scale = gout[0] / (float)n;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float d = a[i] - b[i];
    const float g = is_l1 ? (d > 0.f ? scale : (d < 0.f ? -scale : 0.f)) : 2.f * d * scale;
    if (da) da[i] = accumulate ? da[i] + g : g;
    if (db) db[i] = accumulate ? db[i] - g : -g;
  }
}

static int ew_blocks(size_t total) {
  return (int)std::min<size_t>(ceil_div(total, (size_t)256), (size_t)num_sms() * 16);
}

}  // namespace avc

using namespace avc;

static inline bool al16(const void* p) { return ((uintptr_t)p & 15) == 0; }

static int launch_col_sums(const float* x, int ldx, int M, int C, double* out, bool sq, cudaStream_t st) {
  if (C % 4 == 0 && ldx % 4 == 0 && al16(x)) {
    dim3 grid;
    int rpb;
    cols_grid(M, C, grid, rpb);
    if (sq) col_sums_cols_kernel<true><<<grid, dim3(32, 8), 0, st>>>((const float4*)x, ldx / 4, M, C, rpb, out);
    else col_sums_cols_kernel<false><<<grid, dim3(32, 8), 0, st>>>((const float4*)x, ldx / 4, M, C, rpb, out);
  } else {
    dim3 grid;
    int rpb;
    col_reduce_grid(M, C, grid, rpb);
    col_reduce_kernel<0><<<grid, dim3(CR_COLS, CR_ROWS), 0, st>>>(x, ldx, nullptr, nullptr, nullptr, nullptr, M, C, rpb, 0, out);
  }
  AVC_LAUNCHED();
  return AVC_OK;
}

extern "C" int avc_channel_stats(const float* x, int ldx, int M, int C, double* stats, void* stream) {
  AVC_REQUIRE(x && stats && M > 0 && C > 0 && ldx >= C, "avc_channel_stats: bad arguments");
  return launch_col_sums(x, ldx, M, C, stats, true, as_stream(stream));
}

extern "C" int avc_bn_finalize(const double* stats, int M, int C, float eps, float momentum, float* mean, float* rstd,
                               float* running_mean, float* running_var, void* stream) {
  AVC_REQUIRE(stats && mean && rstd && M > 0 && C > 0, "avc_bn_finalize: bad arguments");
  bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, as_stream(stream)>>>(stats, M, C, eps, momentum, mean, rstd,
                                                                       running_mean, running_var);
  AVC_LAUNCHED();
  return AVC_OK;
}

extern "C" int avc_bn_eval_stats(const float* running_mean, const float* running_var, int C, float eps, float* mean,
                                 float* rstd, void* stream) {
  AVC_REQUIRE(running_mean && running_var && mean && rstd && C > 0, "avc_bn_eval_stats: bad arguments");
  bn_eval_stats_kernel<<<ceil_div(C, 128), 128, 0, as_stream(stream)>>>(running_mean, running_var, C, eps, mean, rstd);
  AVC_LAUNCHED();
  return AVC_OK;
}

static int bn_fwd_impl(const float* y, const float* mean, const float* rstd, const float* gamma, const float* beta,
                       const float* residual, float* z, void* z16, int fmt16, int M, int C, int act, cudaStream_t st,
                       void* z16b = nullptr, int fmt16b = 0) {
  const bool vec = (C % 4 == 0) && al16(y) && (!z || al16(z)) && (!residual || al16(residual)) && (!z16 || ((uintptr_t)z16 & 7) == 0) &&
                   (!z16b || ((uintptr_t)z16b & 7) == 0);
  if (vec) {
    dim3 grid;
    int rpb;
    cols_grid(M, C, grid, rpb);
    bn_act_fwd_cols_kernel<<<grid, dim3(32, 8), 0, st>>>((const float4*)y, mean, rstd, gamma, beta, (const float4*)residual,
                                                         (float4*)z, (uint2*)z16, fmt16, (uint2*)z16b, fmt16b, M, C, rpb, act);
  } else {
    if (z16 || z16b || !z) {
      set_error("avc_bn_act_fwd_h: needs C %% 4 == 0 and 16-byte aligned tensors");
      return AVC_ERR_INVALID;
    }
    const size_t total = (size_t)M * C;
    bn_act_fwd_kernel<<<ew_blocks(total), 256, 0, st>>>(y, mean, rstd, gamma, beta, residual, z, total, C, act);
  }
  AVC_LAUNCHED();
  return AVC_OK;
}

extern "C" int avc_bn_act_fwd(const float* y, const float* mean, const float* rstd, const float* gamma,
                              const float* beta, const float* residual, float* z, int M, int C, int act, void* stream) {
  AVC_REQUIRE(y && mean && rstd && gamma && beta && z && M > 0 && C > 0, "avc_bn_act_fwd: bad arguments");
  return bn_fwd_impl(y, mean, rstd, gamma, beta, residual, z, nullptr, 0, M, C, act, as_stream(stream));
}

// z may be NULL when gamma/beta are given and the tensors allow the float4 path: the activation is then recomputed from y
static int bn_bwd_reduce_impl(const float* dz, const float* z, const float* y, const float* mean, const float* rstd,
                              const float* gamma, const float* beta, double* sums, int M, int C, int act, cudaStream_t st) {
  const bool vec = (C % 4 == 0) && al16(dz) && al16(y) && (!z || al16(z));
  if (vec) {
    dim3 grid;
    int rpb;
    cols_grid(M, C, grid, rpb);
    const bool recompute = gamma && beta;
    bn_bwd_reduce_cols_kernel<<<grid, dim3(32, 8), 0, st>>>((const float4*)dz, recompute ? nullptr : (const float4*)z,
                                                            (const float4*)y, mean, rstd, gamma, beta, M, C, rpb, act, sums);
    AVC_LAUNCHED();
    return AVC_OK;
  }
  if (!z) {
    set_error("avc_bn_act_bwd_reduce_y: z may only be NULL when C %% 4 == 0 and the tensors are 16-byte aligned");
    return AVC_ERR_INVALID;
  }
  dim3 grid;
  int rpb;
  col_reduce_grid(M, C, grid, rpb);
  col_reduce_kernel<1><<<grid, dim3(CR_COLS, CR_ROWS), 0, st>>>(dz, C, z, y, mean, rstd, M, C, rpb, act, sums);
  AVC_LAUNCHED();
  return AVC_OK;
}

static int bn_bwd_apply_impl(const float* dz, const float* z, const float* y, const float* mean, const float* rstd,
                             const float* gamma, const float* beta, const double* sums, float* dy, void* dy16, int fmt16,
                             float* dgamma, float* dbeta, int M, int C, int act, int accumulate, cudaStream_t st) {
  const bool vec = (C % 4 == 0) && al16(dz) && al16(y) && (!z || al16(z)) && (!dy || al16(dy)) && (!dy16 || ((uintptr_t)dy16 & 7) == 0);
  if (vec) {
    dim3 grid;
    int rpb;
    cols_grid(M, C, grid, rpb);
    bn_bwd_apply_cols_kernel<<<grid, dim3(32, 8), 0, st>>>((const float4*)dz, beta ? nullptr : (const float4*)z, (const float4*)y,
                                                           mean, rstd, gamma, beta, sums, (float4*)dy, (uint2*)dy16, fmt16, M, C,
                                                           rpb, act);
  } else {
    if (!z || dy16 || !dy) {
      set_error("avc_bn_act_bwd_apply: the scalar path needs z and an fp32 dy (C %% 4 != 0 or unaligned tensors)");
      return AVC_ERR_INVALID;
    }
    const size_t total = (size_t)M * C;
    bn_act_bwd_apply_kernel<<<ew_blocks(total), 256, 0, st>>>(dz, z, y, mean, rstd, gamma, sums, dy, total, M, C, act);
  }
  AVC_LAUNCHED();
  if (dgamma || dbeta) {
    bn_param_grad_kernel<<<ceil_div(C, 128), 128, 0, st>>>(sums, dgamma, dbeta, C, accumulate);
    AVC_LAUNCHED();
  }
  return AVC_OK;
}

extern "C" int avc_bn_act_bwd_reduce(const float* dz, const float* z, const float* y, const float* mean,
                                     const float* rstd, double* sums, int M, int C, int act, void* stream) {
  AVC_REQUIRE(dz && z && y && mean && rstd && sums && M > 0 && C > 0, "avc_bn_act_bwd_reduce: bad arguments");
  return bn_bwd_reduce_impl(dz, z, y, mean, rstd, nullptr, nullptr, sums, M, C, act, as_stream(stream));
}

extern "C" int avc_bn_act_bwd_apply(const float* dz, const float* z, const float* y, const float* mean,
                                    const float* rstd, const float* gamma, const double* sums, float* dy, float* dgamma,
                                    float* dbeta, int M, int C, int act, int accumulate, void* stream) {
  AVC_REQUIRE(dz && z && y && mean && rstd && gamma && sums && dy && M > 0 && C > 0, "avc_bn_act_bwd_apply: bad arguments");
  return bn_bwd_apply_impl(dz, z, y, mean, rstd, gamma, nullptr, sums, dy, nullptr, 0, dgamma, dbeta, M, C, act, accumulate,
                           as_stream(stream));
}

extern "C" int avc_bn_act_bwd_reduce_y(const float* dz, const float* z, const float* y, const float* mean, const float* rstd,
                                       const float* gamma, const float* beta, double* sums, int M, int C, int act, void* stream) {
  AVC_REQUIRE(dz && y && mean && rstd && gamma && beta && sums && M > 0 && C > 0, "avc_bn_act_bwd_reduce_y: bad arguments");
  return bn_bwd_reduce_impl(dz, z, y, mean, rstd, gamma, beta, sums, M, C, act, as_stream(stream));
}

extern "C" int avc_bn_act_bwd_apply_y(const float* dz, const float* z, const float* y, const float* mean, const float* rstd,
                                      const float* gamma, const float* beta, const double* sums, float* dy, void* dy16, int fmt16,
                                      float* dgamma, float* dbeta, int M, int C, int act, int accumulate, void* stream) {
  AVC_REQUIRE(dz && y && mean && rstd && gamma && beta && sums && (dy || dy16) && M > 0 && C > 0,
              "avc_bn_act_bwd_apply_y: bad arguments");
  AVC_REQUIRE(!dy16 || fmt16 == 1 || fmt16 == 2, "avc_bn_act_bwd_apply_y: fmt16 must be 1 (bf16) or 2 (fp16)");
  return bn_bwd_apply_impl(dz, z, y, mean, rstd, gamma, beta, sums, dy, dy16, fmt16, dgamma, dbeta, M, C, act, accumulate,
                           as_stream(stream));
}

extern "C" int avc_colsum(const float* x, int ldx, int M, int C, float* out, float* out2, int out_mode, int accumulate,
                          void* workspace, size_t workspace_bytes, void* stream) {
  AVC_REQUIRE(x && out && M > 0 && C > 0 && ldx >= C, "avc_colsum: bad arguments");
  AVC_REQUIRE(out_mode == 0 || (out_mode == 2 && C % 4 == 0), "avc_colsum: bad out_mode");
  if (!workspace || workspace_bytes < 2 * sizeof(double) * (size_t)C) {
    set_error("avc_colsum: workspace too small");
    return AVC_ERR_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  AVC_CUDA(cudaMemsetAsync(workspace, 0, 2 * sizeof(double) * (size_t)C, st));
  const int rc = launch_col_sums(x, ldx, M, C, (double*)workspace, false, st);
  if (rc) return rc;
  colsum_finalize_kernel<<<ceil_div(C, 128), 128, 0, st>>>((const double*)workspace, out, out2, C, out_mode, accumulate);
  AVC_LAUNCHED();
  return AVC_OK;
}

extern "C" int avc_colsum16(const void* x16, int fmt, int ldx, int M, int C, float* out, float* out2, int out_mode, int accumulate,
                            void* workspace, size_t workspace_bytes, void* stream) {
  AVC_REQUIRE(x16 && out && M > 0 && C > 0 && ldx >= C && (fmt == 1 || fmt == 2), "avc_colsum16: bad arguments");
  AVC_REQUIRE(C % 4 == 0 && ldx % 4 == 0 && ((uintptr_t)x16 & 7) == 0, "avc_colsum16: needs C, ldx multiples of 4 and an 8-byte aligned input");
  AVC_REQUIRE(out_mode == 0 || out_mode == 2, "avc_colsum16: bad out_mode");
  if (!workspace || workspace_bytes < 2 * sizeof(double) * (size_t)C) {
    set_error("avc_colsum16: workspace too small");
    return AVC_ERR_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  AVC_CUDA(cudaMemsetAsync(workspace, 0, 2 * sizeof(double) * (size_t)C, st));
  dim3 grid;
  int rpb;
  cols_grid(M, C, grid, rpb);
  col_sums16_cols_kernel<<<grid, dim3(32, 8), 0, st>>>((const uint2*)x16, ldx / 4, M, C, rpb, fmt, (double*)workspace);
  AVC_LAUNCHED();
  colsum_finalize_kernel<<<ceil_div(C, 128), 128, 0, st>>>((const double*)workspace, out, out2, C, out_mode, accumulate);
  AVC_LAUNCHED();
  return AVC_OK;
}

extern "C" int avc_mse_loss_fwd(const float* a, const float* b, size_t n, double* scratch, float* out, void* stream) {
  AVC_REQUIRE(a && b && scratch && out && n > 0, "avc_mse_loss_fwd: bad arguments");
  loss_fwd_kernel<false><<<ew_blocks(n), 256, 0, as_stream(stream)>>>(a, b, n, scratch);
  AVC_LAUNCHED();
  loss_finalize_kernel<<<1, 1, 0, as_stream(stream)>>>(scratch, n, out);
  AVC_LAUNCHED();
  return AVC_OK;
}

extern "C" int avc_l1_loss_fwd(const float* a, const float* b, size_t n, double* scratch, float* out, void* stream) {
  AVC_REQUIRE(a && b && scratch && out && n > 0, "avc_l1_loss_fwd: bad arguments");
  loss_fwd_kernel<true><<<ew_blocks(n), 256, 0, as_stream(stream)>>>(a, b, n, scratch);
  AVC_LAUNCHED();
  loss_finalize_kernel<<<1, 1, 0, as_stream(stream)>>>(scratch, n, out);
  AVC_LAUNCHED();
  return AVC_OK;
}

extern "C" int avc_loss_bwd(const float* a, const float* b, size_t n, const float* gout, int is_l1, float* da, float* db,
                            int accumulate, void* stream) {
  AVC_REQUIRE(a && b && gout && (da || db) && n > 0, "avc_loss_bwd: bad arguments");
  loss_bwd_kernel<<<ew_blocks(n), 256, 0, as_stream(stream)>>>(a, b, n, gout, is_l1, da, db, accumulate);
  AVC_LAUNCHED();
  return AVC_OK;
}

// ---- "half" mode producers: the same kernels additionally (or instead) emit the 16-bit operand copy the next GEMM reads ----
extern "C" int avc_cast16(const float* src, int lds, void* dst, int ldd, size_t M, int C, int fmt, void* stream) {
  AVC_REQUIRE(src && dst && M > 0 && C > 0 && lds >= C && ldd >= C && (fmt == 1 || fmt == 2), "avc_cast16: bad arguments");
  const size_t total = M * (size_t)C;
  if (lds == C && ldd == C && C % 4 == 0 && ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 8 == 0))
    cast16_vec4_kernel<<<ew_blocks(total / 4), 256, 0, as_stream(stream)>>>((const float4*)src, (uint2*)dst, total / 4, fmt);
  else
    cast16_kernel<<<ew_blocks(total), 256, 0, as_stream(stream)>>>(src, lds, (uint16_t*)dst, ldd, M, C, fmt);
  AVC_LAUNCHED();
  return AVC_OK;
}

extern "C" int avc_bn_act_fwd_h(const float* y, const float* mean, const float* rstd, const float* gamma, const float* beta,
                                const float* residual, float* z, void* z16, int fmt16, void* z16b, int fmt16b, int M, int C, int act,
                                void* stream) {
  AVC_REQUIRE(y && mean && rstd && gamma && beta && z16 && M > 0 && C > 0, "avc_bn_act_fwd_h: bad arguments");   // z (fp32) is optional
  AVC_REQUIRE(C % 4 == 0 && (fmt16 == 1 || fmt16 == 2) && (!z16b || fmt16b == 1 || fmt16b == 2), "avc_bn_act_fwd_h: C must be a multiple of 4");
  return bn_fwd_impl(y, mean, rstd, gamma, beta, residual, z, z16, fmt16, M, C, act, as_stream(stream), z16b, fmt16b);
}

extern "C" int avc_bn_act_bwd_apply_h(const float* dz, const float* z, const float* y, const float* mean, const float* rstd,
                                      const float* gamma, const double* sums, float* dy, void* dy16, int fmt16, float* dgamma,
                                      float* dbeta, int M, int C, int act, int accumulate, void* stream) {
  AVC_REQUIRE(dz && z && y && mean && rstd && gamma && sums && (dy || dy16) && M > 0 && C > 0, "avc_bn_act_bwd_apply_h: bad arguments");
  AVC_REQUIRE(C % 4 == 0 && (fmt16 == 1 || fmt16 == 2), "avc_bn_act_bwd_apply_h: C must be a multiple of 4");
  return bn_bwd_apply_impl(dz, z, y, mean, rstd, gamma, nullptr, sums, dy, dy16, fmt16, dgamma, dbeta, M, C, act, accumulate,
                           as_stream(stream));
}
