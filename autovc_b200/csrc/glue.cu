// Data-movement glue the reference does with squeeze/transpose/expand/cat/slicing:
// embedding concat (model_vc_mel.py:64-66), code down-sampling (:74-79), code up-sampling + target
// embedding concat (:186-192), and their gradients.
#include "common.cuh"

namespace avc {

// one block walks rows, threads walk the columns of a row: 32-bit index math only (a flat 64-bit i % C per element made
// this kernel ALU-bound at 17% of L1 / 4% of DRAM, r01b ncu)
__global__ void __launch_bounds__(256)
concat_bcast_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ e, float* __restrict__ out, int M, int T,
                    int Cx, int E) {
  const int C = Cx + E;
  for (int m = blockIdx.x; m < M; m += gridDim.x) {
    const float* xr = x + (size_t)m * ldx;
    const float* er = e + (size_t)(m / T) * E;
    float* orow = out + (size_t)m * C;
    for (int c = threadIdx.x; c < C; c += 256) orow[c] = c < Cx ? xr[c] : er[c - Cx];
  }
}

__global__ void codes_fwd_kernel(const float* __restrict__ enc, float* __restrict__ codes, int nB, int T, int n, int f) {
  const int J = T / f, n2 = 2 * n;
  const size_t total = (size_t)nB * J * n2;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % n2);
    const int j = (int)((i / n2) % J);
    const int b = (int)(i / ((size_t)n2 * J));
    const int t = c < n ? j * f + f - 1 : j * f;
    codes[i] = enc[((size_t)b * T + t) * n2 + c];
  }
}

__global__ void codes_bwd_kernel(const float* __restrict__ dcodes, float* __restrict__ denc, int nB, int T, int n, int f) {
  const int J = T / f, n2 = 2 * n;
  const size_t total = (size_t)nB * J * n2;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % n2);
    const int j = (int)((i / n2) % J);
    const int b = (int)(i / ((size_t)n2 * J));
    const int t = c < n ? j * f + f - 1 : j * f;
    denc[((size_t)b * T + t) * n2 + c] += dcodes[i];   // each (b,t,c) is touched by at most one code
  }
}

__global__ void __launch_bounds__(256)
upsample_concat_fwd_kernel(const float* __restrict__ codes, const float* __restrict__ c_trg, float* __restrict__ out, int nB,
                           int T, int n2, int f, int E) {
  const int C = n2 + E, J = T / f, M = nB * T;
  for (int m = blockIdx.x; m < M; m += gridDim.x) {
    const int b = m / T, t = m - b * T;
    const float* cr = codes + ((size_t)b * J + t / f) * n2;
    const float* er = c_trg + (size_t)b * E;
    float* orow = out + (size_t)m * C;
    for (int c = threadIdx.x; c < C; c += 256) orow[c] = c < n2 ? cr[c] : er[c - n2];
  }
}

__global__ void upsample_concat_bwd_kernel(const float* __restrict__ dout, int lddo, float* __restrict__ dcodes, int nB,
                                           int T, int n2, int f, int accumulate) {
  const int J = T / f;
  const size_t total = (size_t)nB * J * n2;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % n2);
    const int j = (int)((i / n2) % J);
    const int b = (int)(i / ((size_t)n2 * J));
    float s = 0.f;
    for (int t = j * f; t < (j + 1) * f; ++t) s += dout[((size_t)b * T + t) * lddo + c];
    dcodes[i] = accumulate ? dcodes[i] + s : s;
  }
}

__global__ void copy2d_kernel(const float* __restrict__ src, int ldsrc, float* __restrict__ dst, int lddst, int M, int C) {
  const size_t total = (size_t)M * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const size_t m = i / C;
    dst[m * lddst + c] = src[m * ldsrc + c];
  }
}

static int ew_blocks(size_t total) {
  return (int)std::min<size_t>(ceil_div(total, (size_t)256), (size_t)num_sms() * 16);
}

}  // namespace avc

using namespace avc;

extern "C" int avc_concat_bcast(const float* x, int ldx, const float* e, float* out, int nB, int T, int Cx, int E, void* stream) {
  AVC_REQUIRE(x && e && out && nB > 0 && T > 0 && Cx > 0 && E > 0 && ldx >= Cx, "avc_concat_bcast: bad arguments");
  concat_bcast_kernel<<<std::min(nB * T, num_sms() * 16), 256, 0, as_stream(stream)>>>(x, ldx, e, out, nB * T, T, Cx, E);
  AVC_LAUNCHED();
  return AVC_OK;
}

extern "C" int avc_codes_fwd(const float* enc, float* codes, int nB, int T, int n, int f, void* stream) {
  AVC_REQUIRE(enc && codes && nB > 0 && T > 0 && n > 0 && f > 0 && T % f == 0, "avc_codes_fwd: T must be a multiple of freq");
  codes_fwd_kernel<<<ew_blocks((size_t)nB * (T / f) * 2 * n), 256, 0, as_stream(stream)>>>(enc, codes, nB, T, n, f);
  AVC_LAUNCHED();
  return AVC_OK;
}

extern "C" int avc_codes_bwd(const float* dcodes, float* denc, int nB, int T, int n, int f, void* stream) {
  AVC_REQUIRE(dcodes && denc && nB > 0 && T > 0 && n > 0 && f > 0 && T % f == 0, "avc_codes_bwd: T must be a multiple of freq");
  codes_bwd_kernel<<<ew_blocks((size_t)nB * (T / f) * 2 * n), 256, 0, as_stream(stream)>>>(dcodes, denc, nB, T, n, f);
  AVC_LAUNCHED();
  return AVC_OK;
}

extern "C" int avc_upsample_concat_fwd(const float* codes, const float* c_trg, float* out, int nB, int T, int n2, int f, int E,
                                       void* stream) {
  AVC_REQUIRE(codes && c_trg && out && nB > 0 && T > 0 && n2 > 0 && f > 0 && E > 0 && T % f == 0,
              "avc_upsample_concat_fwd: bad arguments");
  upsample_concat_fwd_kernel<<<std::min(nB * T, num_sms() * 16), 256, 0, as_stream(stream)>>>(codes, c_trg, out, nB, T, n2, f, E);
  AVC_LAUNCHED();
  return AVC_OK;
}

extern "C" int avc_upsample_concat_bwd(const float* dout, int lddo, float* dcodes, int nB, int T, int n2, int f, int accumulate,
                                       void* stream) {
  AVC_REQUIRE(dout && dcodes && nB > 0 && T > 0 && n2 > 0 && f > 0 && T % f == 0 && lddo >= n2,
              "avc_upsample_concat_bwd: bad arguments");
  upsample_concat_bwd_kernel<<<ew_blocks((size_t)nB * (T / f) * n2), 256, 0, as_stream(stream)>>>(dout, lddo, dcodes, nB, T, n2, f, accumulate);
  AVC_LAUNCHED();
  return AVC_OK;
}

extern "C" int avc_copy2d(const float* src, int ldsrc, float* dst, int lddst, int M, int C, void* stream) {
  AVC_REQUIRE(src && dst && M > 0 && C > 0 && ldsrc >= C && lddst >= C, "avc_copy2d: bad arguments");
  copy2d_kernel<<<ew_blocks((size_t)M * C), 256, 0, as_stream(stream)>>>(src, ldsrc, dst, lddst, M, C);
  AVC_LAUNCHED();
  return AVC_OK;
}

// embeds.div(embeds.norm(p=2, dim=-1, keepdim=True)), model_bl.py:18-19: one warp per row
namespace avc {
__global__ void l2_normalize_rows_kernel(const float* __restrict__ x, float* __restrict__ out, int M, int C) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* xr = x + (size_t)row * C;
  float ss = 0.f;
  for (int c = lane; c < C; c += 32) ss = fmaf(xr[c], xr[c], ss);
  ss = warp_sum(ss);
  const float nrm = sqrtf(ss);
  for (int c = lane; c < C; c += 32) out[(size_t)row * C + c] = xr[c] / nrm;
}
}  // namespace avc

extern "C" int avc_l2_normalize_rows(const float* x, float* out, int M, int C, void* stream) {
  AVC_REQUIRE(x && out && M > 0 && C > 0, "avc_l2_normalize_rows: bad arguments");
  avc::l2_normalize_rows_kernel<<<avc::ceil_div(M, 8), 256, 0, avc::as_stream(stream)>>>(x, out, M, C);
  AVC_LAUNCHED();
  return AVC_OK;
}
