// Crop loader, the step before the hot path: data_loader.py:61-80 (`Utterances.__getitem__`: random
// utterance of a speaker, random crop of len_crop frames or zero-padding at the end) + the default collate of :90-102.
// The corpus lives in HBM as ONE ragged (sum F_i, n_bins) fp32 buffer with per-utterance row offsets and lengths; the host
// draws the random choices exactly like the reference (numpy stream) and hands over three int arrays per batch; this kernel
// gathers the B crops and the B speaker embeddings.  HBM-bound: 2 x 4 x T x n_bins bytes per crop, rows copied as float4
// when n_bins % 4 == 0 (every row of the corpus is then 16-byte aligned).
#include <algorithm>

#include "common.cuh"

namespace avc {

__global__ void __launch_bounds__(256)
crop_batch_kernel(const float* __restrict__ corpus, const long long* __restrict__ utt_row0, const int* __restrict__ utt_len,
                  const int* __restrict__ sel_utt, const int* __restrict__ sel_left, const float* __restrict__ emb_table,
                  const int* __restrict__ sel_spk, float* __restrict__ x_out, float* __restrict__ e_out, int B, int T, int n_bins,
                  int dim_emb) {
  const int b = blockIdx.y;
  const int u = sel_utt[b];
  const int F = utt_len[u];
  const int left = F > T ? sel_left[b] : 0;
  const int avail = min(T, F - left);                      // frames copied; the rest is zero padding (data_loader.py:70-73)
  const float* src = corpus + ((size_t)utt_row0[u] + left) * n_bins;
  float* dst = x_out + (size_t)b * T * n_bins;
  const size_t n_copy = (size_t)avail * n_bins, n_all = (size_t)T * n_bins;
  if ((n_bins & 3) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    const size_t c4 = n_copy >> 2, a4 = n_all >> 2;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < a4; i += (size_t)gridDim.x * blockDim.x)
      d4[i] = i < c4 ? __ldg(s4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_all; i += (size_t)gridDim.x * blockDim.x)
      dst[i] = i < n_copy ? __ldg(src + i) : 0.f;
  }
  if (blockIdx.x == 0) {
    const float* es = emb_table + (size_t)sel_spk[b] * dim_emb;
    for (int i = threadIdx.x; i < dim_emb; i += blockDim.x) e_out[(size_t)b * dim_emb + i] = __ldg(es + i);
  }
}

}  // namespace avc

using namespace avc;

extern "C" int avc_crop_batch(const float* corpus, const long long* utt_row0, const int* utt_len, const int* sel_utt,
                              const int* sel_left, const float* emb_table, const int* sel_spk, float* x_out, float* e_out, int B,
                              int T, int n_bins, int dim_emb, void* stream) {
  AVC_REQUIRE(corpus && utt_row0 && utt_len && sel_utt && sel_left && emb_table && sel_spk && x_out && e_out,
              "avc_crop_batch: null pointer");
  AVC_REQUIRE(B > 0 && T > 0 && n_bins > 0 && dim_emb > 0, "avc_crop_batch: bad shape B=%d T=%d n_bins=%d dim_emb=%d", B, T, n_bins, dim_emb);
  AVC_REQUIRE((n_bins & 3) != 0 || ((((uintptr_t)corpus) | ((uintptr_t)x_out)) & 15) == 0, "avc_crop_batch: corpus / output must be 16-byte aligned");
  const size_t per_crop = (size_t)T * n_bins / (((n_bins & 3) == 0) ? 4 : 1);
  const int gx = (int)std::max<size_t>(1, std::min<size_t>(ceil_div(per_crop, (size_t)256), 8));
  crop_batch_kernel<<<dim3(gx, B), 256, 0, as_stream(stream)>>>(corpus, utt_row0, utt_len, sel_utt, sel_left, emb_table, sel_spk, x_out,
                                                               e_out, B, T, n_bins, dim_emb);
  AVC_LAUNCHED();
  return AVC_OK;
}
