// Data-movement glue the reference does with squeeze/transpose/expand/cat/slicing:
// embedding concat (model_vc_mel.py:64-66), code down-sampling (:74-79), code up-sampling + target
// embedding concat (:186-192), and their gradients.
#include "common.cuh"

namespace avc {

__global__ void concat_bcast_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ e,
                                    float* __restrict__ out, int M, int T, int Cx, int E) {
  const int C = Cx + E;
  const size_t total = (size_t)M * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int m = (int)(i / C);
    out[i] = c < Cx ? x[(size_t)m * ldx + c] : e[(size_t)(m / T) * E + (c - Cx)];
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

__global__ void upsample_concat_fwd_kernel(const float* __restrict__ codes, const float* __restrict__ c_trg,
                                           float* __restrict__ out, int nB, int T, int n2, int f, int E) {
  const int C = n2 + E, J = T / f;
  const size_t total = (size_t)nB * T * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int t = (int)((i / C) % T);
    const int b = (int)(i / ((size_t)C * T));
    out[i] = c < n2 ? codes[((size_t)b * J + t / f) * n2 + c] : c_trg[(size_t)b * E + (c - n2)];
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
  concat_bcast_kernel<<<ew_blocks((size_t)nB * T * (Cx + E)), 256, 0, as_stream(stream)>>>(x, ldx, e, out, nB * T, T, Cx, E);
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
  upsample_concat_fwd_kernel<<<ew_blocks((size_t)nB * T * (n2 + E)), 256, 0, as_stream(stream)>>>(codes, c_trg, out, nB, T, n2, f, E);
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
