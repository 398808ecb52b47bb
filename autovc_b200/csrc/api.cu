// C-ABI front door: error plumbing and precision dispatch.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace avc {

static thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// fp32 CUDA-core implementations (gemm_simt.cu, lstm_simt.cu)
int gemm_nt_taps_simt(const float*, int, const float*, const float*, float*, int, int, int, int, int, int, int, double*, int, cudaStream_t);
int gemm_tn_taps_simt(const float*, int, const float*, int, float*, int, int, int, int, int, int, int, int, void*, size_t, cudaStream_t);
size_t gemm_tn_workspace_simt(int, int, int, int, int);
int lstm_seq_fwd_simt(const float*, const float*, float*, int, float*, float*, int, int, int, int, cudaStream_t);
int lstm_seq_bwd_simt(const float*, int, const float*, const float*, const float*, const float*, float*, int, int, int, int, void*, size_t, cudaStream_t);
size_t lstm_bwd_workspace_simt(int, int, int);
// tensor-core implementations (tc_gemm.cu)
int gemm_nt_taps_tc(const void*, int, int, const void*, int, int, const float*, float*, int, int, int, int, int, int, int, double*, int, int, int, void*, size_t, cudaStream_t, int chunk = 0,
                    int ksplit = 1, size_t c_split_stride = 0);
int gemm_tn_taps_tc(const void*, int, int, const void*, int, int, float*, int, int, int, int, int, int, int, int, int, int, void*, size_t, cudaStream_t, int chunk = 0);
size_t gemm_nt_workspace_tc(int, int, int, int, int, int);
size_t gemm_nt_workspace_h(int, int, int, int, int, int);
size_t gemm_tn_workspace_h(int, int, int, int, int, int, int);
size_t gemm_tn_workspace_tc(int, int, int, int, int, int, int chunk = 0);
// 3xTF32 (fp32x3.cu)
int gemm_nt_taps_x3(const float*, int, const float*, const float*, float*, int, int, int, int, int, int, int, double*, int, void*, size_t, cudaStream_t);
int gemm_tn_taps_x3(const float*, int, const float*, int, float*, int, int, int, int, int, int, int, int, void*, size_t, cudaStream_t);
size_t gemm_nt_workspace_x3(int, int, int, int, int);
size_t gemm_tn_workspace_x3(int, int, int, int, int);
int lstm_seq_fwd_x3(const float*, const float*, float*, int, float*, float*, int, int, int, int, void*, size_t, cudaStream_t);
int lstm_seq_bwd_x3(const float*, int, const float*, const float*, const float*, float*, int, int, int, int, void*, size_t, cudaStream_t);
size_t lstm_fwd_workspace_x3(int, int, int);
size_t lstm_bwd_workspace_x3(int, int, int);
constexpr int X3_SMALL_H = 64;      // up to here the whole-sequence CUDA-core kernels (W_hh in registers / shared memory) are used
// persistent tensor-core recurrences (lstm_tc.cu)
bool lstm_tc_supported(int H);
void lstm_tc_set_trace(unsigned long long* p);
void tc_gemm_set_trace(unsigned long long* p);
size_t lstm_tc_workspace(int nB, int T, int H, bool bwd);
int lstm_seq_tc(bool bwd, const void* W, const float* P, float* h_seq, int ldh, float* gates, float* c_seq, const float* dH,
                int lddh, float* dP, int nB, int T, int H, int reverse, void* ws, size_t ws_bytes, cudaStream_t st,
                void* aux16 = nullptr, int fmt16 = 0, int w_fmt = 0, void* aux16b = nullptr);

}  // namespace avc

using namespace avc;

extern "C" int avc_version(void) { return AVC_VERSION; }
extern "C" const char* avc_last_error(void) { return g_err; }
extern "C" unsigned long long avc_launch_count(void) { return g_launches.load(); }

extern "C" int avc_gemm_nt_taps(const float* A, int lda, const float* W, const float* bias, float* C, int ldc, int nB, int T,
                                int N, int K, int ntaps, int shift0, double* chan_stats, int accumulate, int prec, void* workspace,
                                size_t workspace_bytes, void* stream) {
  AVC_REQUIRE(A && W && C, "avc_gemm_nt_taps: null pointer");
  AVC_REQUIRE(nB > 0 && T > 0 && N > 0 && K > 0 && ntaps > 0, "avc_gemm_nt_taps: bad shape B=%d T=%d N=%d K=%d taps=%d", nB, T, N, K, ntaps);
  AVC_REQUIRE(lda >= K && ldc >= N, "avc_gemm_nt_taps: leading dimensions lda=%d < K=%d or ldc=%d < N=%d", lda, K, ldc, N);
  if (prec == AVC_PREC_FP32)
    return gemm_nt_taps_simt(A, lda, W, bias, C, ldc, nB, T, N, K, ntaps, shift0, chan_stats, accumulate, as_stream(stream));
  if (prec == AVC_PREC_BF16 || prec == AVC_PREC_TF32)
    return gemm_nt_taps_tc(A, 0, lda, W, 0, K, bias, C, ldc, nB, T, N, K, ntaps, shift0, chan_stats, accumulate,
                           prec == AVC_PREC_BF16 ? 2 : 4, 1, workspace, workspace_bytes, as_stream(stream));
  if (prec == AVC_PREC_FP32X3)
    return gemm_nt_taps_x3(A, lda, W, bias, C, ldc, nB, T, N, K, ntaps, shift0, chan_stats, accumulate, workspace, workspace_bytes,
                           as_stream(stream));
  set_error("avc_gemm_nt_taps: precision %d not available in this build", prec);
  return AVC_ERR_UNSUPPORTED;
}

extern "C" int avc_gemm_tn_taps(const float* dY, int ldy, const float* X, int ldx, float* dW, int nB, int T, int N, int K,
                                int ntaps, int shift0, int out_mode, int accumulate, int prec, void* workspace,
                                size_t workspace_bytes, void* stream) {
  AVC_REQUIRE(dY && X && dW, "avc_gemm_tn_taps: null pointer");
  AVC_REQUIRE(nB > 0 && T > 0 && N > 0 && K > 0 && ntaps > 0, "avc_gemm_tn_taps: bad shape");
  AVC_REQUIRE(ldy >= N && ldx >= K, "avc_gemm_tn_taps: bad leading dimensions");
  AVC_REQUIRE(out_mode == 0 || out_mode == 1 || (out_mode == 2 && ntaps == 1 && N % 4 == 0), "avc_gemm_tn_taps: bad out_mode");
  if (prec == AVC_PREC_FP32)
    return gemm_tn_taps_simt(dY, ldy, X, ldx, dW, nB, T, N, K, ntaps, shift0, out_mode, accumulate, workspace,
                             workspace_bytes, as_stream(stream));
  if (prec == AVC_PREC_BF16 || prec == AVC_PREC_TF32)
    return gemm_tn_taps_tc(dY, 0, ldy, X, 0, ldx, dW, nB, T, N, K, ntaps, shift0, out_mode, accumulate,
                           prec == AVC_PREC_BF16 ? 2 : 4, 1, workspace, workspace_bytes, as_stream(stream));
  if (prec == AVC_PREC_FP32X3)
    return gemm_tn_taps_x3(dY, ldy, X, ldx, dW, nB, T, N, K, ntaps, shift0, out_mode, accumulate, workspace, workspace_bytes,
                           as_stream(stream));
  set_error("avc_gemm_tn_taps: precision %d not available in this build", prec);
  return AVC_ERR_UNSUPPORTED;
}

extern "C" size_t avc_gemm_tn_workspace_bytes(int nB, int T, int N, int K, int ntaps, int prec) {
  if (prec == AVC_PREC_FP32X3) return gemm_tn_workspace_x3(nB, T, N, K, ntaps);
  if (prec == AVC_PREC_BF16 || prec == AVC_PREC_TF32) return gemm_tn_workspace_tc(nB, T, N, K, ntaps, prec == AVC_PREC_BF16 ? 2 : 4);
  return gemm_tn_workspace_simt(nB, T, N, K, ntaps);
}

extern "C" size_t avc_gemm_nt_workspace_bytes(int nB, int T, int N, int K, int ntaps, int prec) {
  if (prec == AVC_PREC_FP32X3) return gemm_nt_workspace_x3(nB, T, N, K, ntaps);
  if (prec == AVC_PREC_BF16 || prec == AVC_PREC_TF32) return gemm_nt_workspace_tc(nB, T, N, K, ntaps, prec == AVC_PREC_BF16 ? 2 : 4);
  return 0;
}

extern "C" int avc_lstm_seq_fwd(const float* P, const float* Whh_p, float* h_seq, int ldh, float* gates, float* c_seq, int nB,
                                int T, int H, int reverse, int prec, void* workspace, size_t workspace_bytes, void* stream) {
  AVC_REQUIRE(P && Whh_p && h_seq && gates && c_seq, "avc_lstm_seq_fwd: null pointer");
  AVC_REQUIRE(nB > 0 && T > 0 && H > 0 && ldh >= H, "avc_lstm_seq_fwd: bad shape");
  if ((prec == AVC_PREC_BF16 || prec == AVC_PREC_TF32) && lstm_tc_supported(H)) {
    AVC_REQUIRE(ldh % 4 == 0, "avc_lstm_seq_fwd(bf16): ldh must be a multiple of 4");
    return lstm_seq_tc(false, Whh_p, P, h_seq, ldh, gates, c_seq, nullptr, 0, nullptr, nB, T, H, reverse, workspace,
                       workspace_bytes, as_stream(stream));
  }
  if (prec == AVC_PREC_FP32X3 && H > X3_SMALL_H)
    return lstm_seq_fwd_x3(P, Whh_p, h_seq, ldh, gates, c_seq, nB, T, H, reverse, workspace, workspace_bytes, as_stream(stream));
  if (prec == AVC_PREC_FP32 || prec == AVC_PREC_BF16 || prec == AVC_PREC_TF32 || prec == AVC_PREC_FP32X3)
    return lstm_seq_fwd_simt(P, Whh_p, h_seq, ldh, gates, c_seq, nB, T, H, reverse, as_stream(stream));
  set_error("avc_lstm_seq_fwd: unknown precision %d", prec);
  return AVC_ERR_UNSUPPORTED;
}

extern "C" size_t avc_lstm_fwd_workspace_bytes(int nB, int T, int H, int prec) {
  if (prec == AVC_PREC_FP32X3) return H > X3_SMALL_H ? lstm_fwd_workspace_x3(nB, T, H) : 0;
  if ((prec == AVC_PREC_BF16 || prec == AVC_PREC_TF32) && lstm_tc_supported(H)) return lstm_tc_workspace(nB, T, H, false);
  return 0;
}

extern "C" int avc_lstm_seq_bwd(const float* dH, int lddh, const float* Whh_p, const float* Whh_pT, const float* gates,
                                const float* c_seq, float* dP, int nB, int T, int H, int reverse, int prec, void* workspace,
                                size_t workspace_bytes, void* stream) {
  AVC_REQUIRE(dH && Whh_p && Whh_pT && gates && c_seq && dP, "avc_lstm_seq_bwd: null pointer");
  AVC_REQUIRE(nB > 0 && T > 0 && H > 0 && lddh >= H, "avc_lstm_seq_bwd: bad shape");
  if ((prec == AVC_PREC_BF16 || prec == AVC_PREC_TF32) && lstm_tc_supported(H)) {
    AVC_REQUIRE(lddh % 4 == 0, "avc_lstm_seq_bwd(bf16): lddh must be a multiple of 4");
    return lstm_seq_tc(true, Whh_pT, nullptr, nullptr, 0, const_cast<float*>(gates), const_cast<float*>(c_seq), dH, lddh, dP,
                       nB, T, H, reverse, workspace, workspace_bytes, as_stream(stream));
  }
  if (prec == AVC_PREC_FP32X3 && H > X3_SMALL_H)
    return lstm_seq_bwd_x3(dH, lddh, Whh_pT, gates, c_seq, dP, nB, T, H, reverse, workspace, workspace_bytes, as_stream(stream));
  if (prec == AVC_PREC_FP32 || prec == AVC_PREC_BF16 || prec == AVC_PREC_TF32 || prec == AVC_PREC_FP32X3)
    return lstm_seq_bwd_simt(dH, lddh, Whh_p, Whh_pT, gates, c_seq, dP, nB, T, H, reverse, workspace, workspace_bytes,
                             as_stream(stream));
  set_error("avc_lstm_seq_bwd: unknown precision %d", prec);
  return AVC_ERR_UNSUPPORTED;
}

extern "C" size_t avc_lstm_bwd_workspace_bytes(int nB, int T, int H, int prec) {
  if (prec == AVC_PREC_FP32X3) return H > X3_SMALL_H ? lstm_bwd_workspace_x3(nB, T, H) : lstm_bwd_workspace_simt(nB, T, H);
  if ((prec == AVC_PREC_BF16 || prec == AVC_PREC_TF32) && lstm_tc_supported(H)) return lstm_tc_workspace(nB, T, H, true);
  return lstm_bwd_workspace_simt(nB, T, H);
}

extern "C" void avc_debug_set_trace(unsigned long long* device_buffer) {
  lstm_tc_set_trace(device_buffer);
  tc_gemm_set_trace(device_buffer);
}

// ---- GEMMs on operands that already live in HBM as 16-bit (formats: 0 = fp32 (staged to half_fmt), 1 = bf16, 2 = fp16) ----
extern "C" int avc_gemm_nt_taps_h(const void* A, int a_fmt, int lda, const float* W, const float* bias, float* C, int ldc, int nB,
                                  int T, int N, int K, int ntaps, int shift0, double* chan_stats, int accumulate, int half_fmt,
                                  void* workspace, size_t workspace_bytes, void* stream) {
  AVC_REQUIRE(A && W && C, "avc_gemm_nt_taps_h: null pointer");
  AVC_REQUIRE(nB > 0 && T > 0 && N > 0 && K > 0 && ntaps > 0 && lda >= K && ldc >= N, "avc_gemm_nt_taps_h: bad shape");
  AVC_REQUIRE(a_fmt >= 0 && a_fmt <= 2 && (half_fmt == 1 || half_fmt == 2), "avc_gemm_nt_taps_h: bad format code");
  return gemm_nt_taps_tc(A, a_fmt, lda, W, 0, K, bias, C, ldc, nB, T, N, K, ntaps, shift0, chan_stats, accumulate, 2, half_fmt,
                         workspace, workspace_bytes, as_stream(stream));
}

extern "C" int avc_gemm_nt_taps_hw(const void* A, int a_fmt, int lda, const void* W16, int w_fmt, int ldw, const float* bias, float* C,
                                   int ldc, int nB, int T, int N, int K, int ntaps, int shift0, double* chan_stats, int accumulate,
                                   void* workspace, size_t workspace_bytes, void* stream) {
  AVC_REQUIRE(A && W16 && C, "avc_gemm_nt_taps_hw: null pointer");
  AVC_REQUIRE(nB > 0 && T > 0 && N > 0 && K > 0 && ntaps > 0 && lda >= K && ldc >= N, "avc_gemm_nt_taps_hw: bad shape");
  AVC_REQUIRE(a_fmt >= 0 && a_fmt <= 2 && (w_fmt == 1 || w_fmt == 2), "avc_gemm_nt_taps_hw: bad format code");
  return gemm_nt_taps_tc(A, a_fmt, lda, W16, w_fmt, ldw, bias, C, ldc, nB, T, N, K, ntaps, shift0, chan_stats, accumulate, 2, w_fmt,
                         workspace, workspace_bytes, as_stream(stream));
}

extern "C" int avc_gemm_tn_taps_h(const void* dY, int y_fmt, int ldy, const void* X, int x_fmt, int ldx, float* dW, int nB, int T,
                                  int N, int K, int ntaps, int shift0, int out_mode, int accumulate, int half_fmt, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  AVC_REQUIRE(dY && X && dW, "avc_gemm_tn_taps_h: null pointer");
  AVC_REQUIRE(nB > 0 && T > 0 && N > 0 && K > 0 && ntaps > 0 && ldy >= N && ldx >= K, "avc_gemm_tn_taps_h: bad shape");
  AVC_REQUIRE(out_mode == 0 || out_mode == 1 || (out_mode == 2 && ntaps == 1 && N % 4 == 0), "avc_gemm_tn_taps_h: bad out_mode");
  AVC_REQUIRE(y_fmt >= 0 && y_fmt <= 2 && x_fmt >= 0 && x_fmt <= 2 && (half_fmt == 1 || half_fmt == 2), "avc_gemm_tn_taps_h: bad format code");
  return gemm_tn_taps_tc(dY, y_fmt, ldy, X, x_fmt, ldx, dW, nB, T, N, K, ntaps, shift0, out_mode, accumulate, 2, half_fmt, workspace,
                         workspace_bytes, as_stream(stream));
}

extern "C" size_t avc_gemm_nt_h_workspace_bytes(int nB, int T, int N, int K, int ntaps, int a_fmt) {
  return gemm_nt_workspace_h(nB, T, N, K, ntaps, a_fmt);
}
extern "C" size_t avc_gemm_tn_h_workspace_bytes(int nB, int T, int N, int K, int ntaps, int y_fmt, int x_fmt) {
  return gemm_tn_workspace_h(nB, T, N, K, ntaps, y_fmt, x_fmt);
}

// persistent recurrences that additionally emit the 16-bit operand copy the following GEMMs read ("half" mode)
extern "C" int avc_lstm_seq_fwd_h(const float* P, const void* Whh_p, int w_fmt, float* h_seq, int ldh, float* gates, float* c_seq,
                                  void* h16, int fmt16, void* h16b, int nB, int T, int H, int reverse, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  AVC_REQUIRE(P && Whh_p && h_seq && h16, "avc_lstm_seq_fwd_h: null pointer");
  AVC_REQUIRE((gates == nullptr) == (c_seq == nullptr), "avc_lstm_seq_fwd_h: gates and c_seq are saved together or not at all");
  AVC_REQUIRE(nB > 0 && T > 0 && lstm_tc_supported(H) && ldh >= H && ldh % 4 == 0 && (fmt16 == 1 || fmt16 == 2), "avc_lstm_seq_fwd_h: unsupported shape");
  AVC_REQUIRE(w_fmt == 0 || w_fmt == 1, "avc_lstm_seq_fwd_h: W_hh must be fp32 (0) or bf16 (1)");
  return lstm_seq_tc(false, Whh_p, P, h_seq, ldh, gates, c_seq, nullptr, 0, nullptr, nB, T, H, reverse, workspace, workspace_bytes,
                     as_stream(stream), h16, fmt16, w_fmt, h16b);
}

extern "C" int avc_lstm_seq_bwd_h(const float* dH, int lddh, const void* Whh_pT, int w_fmt, const float* gates, const float* c_seq,
                                  float* dP, void* dP16, int nB, int T, int H, int reverse, void* workspace, size_t workspace_bytes,
                                  void* stream) {
  AVC_REQUIRE(dH && Whh_pT && gates && c_seq && dP16, "avc_lstm_seq_bwd_h: null pointer");   // dP (fp32) is optional
  AVC_REQUIRE(nB > 0 && T > 0 && lstm_tc_supported(H) && lddh >= H && lddh % 4 == 0, "avc_lstm_seq_bwd_h: unsupported shape");
  AVC_REQUIRE(w_fmt == 0 || w_fmt == 1, "avc_lstm_seq_bwd_h: W_hh^T must be fp32 (0) or bf16 (1)");
  return lstm_seq_tc(true, Whh_pT, nullptr, nullptr, 0, const_cast<float*>(gates), const_cast<float*>(c_seq), dH, lddh, dP, nB, T, H,
                     reverse, workspace, workspace_bytes, as_stream(stream), dP16, 1, w_fmt);
}
