/* autovc_b200 — C-ABI of the B200-native AutoVC hot path (libautovc_b200.so).
 *
 * Drop-in boundary for ONE path of sebakeaaen/autovc: the model_vc_mel.Generator
 * forward/backward training step driven by solver_encoder.py and the make_spect.py log-mel
 * front-end.  The reference has no native layer (it calls torch.nn -> cuDNN/cuBLAS/oneDNN),
 * so each entry point below names the reference call site (file:line, relative to the
 * upstream tree) whose library call it replaces.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - plain C linkage, raw device pointers + sizes, no C++/torch types, no exceptions.
 *   - every function returns 0 on success, non-zero on failure; avc_last_error() returns a
 *     thread-local message.  Shape/alignment violations are errors, never silent fallbacks.
 *   - the library never allocates, frees or retains device memory: the caller (PyTorch's
 *     caching allocator) owns every buffer, including workspaces whose sizes the
 *     avc_*_workspace_bytes queries report.
 *   - all launches are asynchronous on `stream` (a cudaStream_t passed as void*); nothing
 *     synchronises.  Re-entrant; callable from the autograd worker thread.
 *   - activations are channels-last: (B, T, C) row-major, row m = b*T + t, channel stride 1,
 *     row stride `ld*` in elements (so column slices of wider buffers can be addressed).
 *   - `prec`: AVC_PREC_FP32 = CUDA-core FFMA, fp32 operands and accumulation (parity mode,
 *     <=1e-4 of the reference's fp32 path); AVC_PREC_BF16 / AVC_PREC_TF32 = tcgen05/TMEM tensor-core
 *     path with bf16 operands, or fp32 operands fetched by TMA as tf32 (no staging copies); fp32
 *     accumulation, fp32 statistics and fp32 recurrent state in both.  AVC_PREC_FP32X3 = fp32-accurate
 *     products on the tensor cores: every fp32 operand is split into a tf32 head and an fp32 remainder and
 *     a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi runs as ONE kind::tf32 GEMM over a three times longer
 *     reduction (operands staged in `workspace`); recurrences with H > 64 become one such GEMM per time step.
 *     Same <=1e-4 gate as AVC_PREC_FP32, measured at ~3e-5.
 */
#ifndef AUTOVC_B200_H_
#define AUTOVC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define AVC_VERSION 100

enum { AVC_PREC_FP32 = 0, AVC_PREC_BF16 = 1, AVC_PREC_TF32 = 2, AVC_PREC_HALF = 3, AVC_PREC_FP32X3 = 4 };
/* operand formats of the *_h entry points */
enum { AVC_FMT_FP32 = 0, AVC_FMT_BF16 = 1, AVC_FMT_FP16 = 2 };
enum { AVC_ACT_NONE = 0, AVC_ACT_RELU = 1, AVC_ACT_TANH = 2 };
enum {
  AVC_OK = 0,
  AVC_ERR_INVALID = 1,      /* bad shape / alignment / null pointer */
  AVC_ERR_CUDA = 2,         /* a CUDA runtime/driver call failed    */
  AVC_ERR_UNSUPPORTED = 3,  /* combination not implemented          */
  AVC_ERR_WORKSPACE = 4     /* workspace too small                  */
};

int avc_version(void);
const char* avc_last_error(void);
/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
unsigned long long avc_launch_count(void);

/* ---------------------------------------------------------------------------------------
 * GEMM family with time-shifted operand rows ("taps").  One kernel family implements
 *   Conv1d(k=5,p=2) forward        model_vc_mel.py:28-31 (ConvNorm.conv), called :69,:115,:165,:167
 *   Conv1d data gradient           autograd of the above (solver_encoder.py:294)
 *   LSTM input projections         x W_ih^T of nn.LSTM, model_vc_mel.py:73,:111,:118
 *   Linear(1024 -> n_bins)         model_vc_mel.py:10,:120
 * as  C[m, n] = bias[n] + sum_{tap<ntaps} sum_{k<K} A[row(m, tap), k] * W[tap][n][k],
 * where m = b*T + t and row(m, tap) = b*T + (t + shift0 + tap), rows outside [0,T) of the
 * same utterance reading as zero (= the convolution's zero padding).
 *   A: (nB*T, K) row stride lda.  W: packed [ntaps][N][K] (see avc_pack_conv_weight).
 *   C: (nB*T, N) row stride ldc.  bias may be NULL.
 * If chan_stats != NULL (double[2*N]: sum, sum of squares; caller zero-fills) the per-channel
 * batch statistics of C that train-mode BatchNorm1d needs (model_vc_mel.py:57) are accumulated.
 * accumulate != 0: C += result (sums the two directions' input gradients of the BiLSTM).
 */
int avc_gemm_nt_taps(const float* A, int lda, const float* W, const float* bias, float* C, int ldc,
                     int nB, int T, int N, int K, int ntaps, int shift0, double* chan_stats, int accumulate,
                     int prec, void* workspace, size_t workspace_bytes, void* stream);
/* bytes of workspace avc_gemm_nt_taps needs (0 for AVC_PREC_FP32; bf16 staging copies otherwise) */
size_t avc_gemm_nt_workspace_bytes(int nB, int T, int N, int K, int ntaps, int prec);

/* Weight gradients (conv wgrad, LSTM dW_ih / dW_hh, Linear dW), reduction over rows:
 *   dW[tap][n][k] = sum_m dY[m, n] * X[row(m, tap), k]
 * dY: (nB*T, N) ld ldy; X: (nB*T, K) ld ldx.  Deterministic split-M reduction through
 * `workspace`.  The result is written through an output map:
 *   out_mode 0: packed [ntaps][N][K];
 *   out_mode 1: PyTorch Conv1d layout (N, K, ntaps);
 *   out_mode 2: LSTM rows un-permuted: packed row r = u*4+g goes to row g*H+u (N = 4H), ntaps=1.
 * `accumulate` != 0 adds into dW instead of overwriting (second encoder pass). */
int avc_gemm_tn_taps(const float* dY, int ldy, const float* X, int ldx, float* dW,
                     int nB, int T, int N, int K, int ntaps, int shift0, int out_mode, int accumulate,
                     int prec, void* workspace, size_t workspace_bytes, void* stream);
size_t avc_gemm_tn_workspace_bytes(int nB, int T, int N, int K, int ntaps, int prec);

/* Weight repacking (once per optimizer step; weights are tiny next to activations).
 *   avc_pack_conv_weight: PyTorch (Cout, Cin, 5) ->  fwd  [tap][Cout][Cin]
 *                                                    dgrad [tap][Cin][Cout] with taps flipped
 *   avc_pack_lstm_weight: (4H, I) gate-blocked rows i,f,g,o -> gate-interleaved rows u*4+g:
 *                         out_p (4H, I) and its transpose out_pT (I, 4H); either may be NULL.
 *   avc_pack_lstm_bias:   b_ih + b_hh, interleaved.
 *   avc_transpose:        (R, C) -> (C, R).                                                   */
int avc_pack_conv_weight(const float* w, float* w_fwd, float* w_dgrad, int Cout, int Cin, int ntaps, void* stream);
int avc_pack_lstm_weight(const float* w, float* out_p, float* out_pT, int H, int I, void* stream);
int avc_pack_lstm_bias(const float* b_ih, const float* b_hh, float* out, int H, void* stream);
int avc_transpose(const float* in, float* out, int R, int C, void* stream);

/* ---------------------------------------------------------------------------------------
 * BatchNorm1d (train mode) + activation, model_vc_mel.py:57,:69,:115,:165-167 and the
 * residual add :197.
 */
/* Per-channel sum / sum-of-squares of x (M, C) ld ldx, accumulated into stats (double[2*C]). */
int avc_channel_stats(const float* x, int ldx, int M, int C, double* stats, void* stream);
/* stats -> mean, rstd (saved for backward), and the running-stat update of nn.BatchNorm1d
 * (momentum, unbiased variance).  running_* may be NULL (no update). */
int avc_bn_finalize(const double* stats, int M, int C, float eps, float momentum,
                    float* mean, float* rstd, float* running_mean, float* running_var, void* stream);
/* eval mode: mean/rstd from the running statistics. */
int avc_bn_eval_stats(const float* running_mean, const float* running_var, int C, float eps,
                      float* mean, float* rstd, void* stream);
/* z = act((y - mean) * rstd * gamma + beta) (+ residual).  y,z,(residual): (M, C). */
int avc_bn_act_fwd(const float* y, const float* mean, const float* rstd, const float* gamma, const float* beta,
                   const float* residual, float* z, int M, int C, int act, void* stream);
/* backward, two passes: (1) g = dz * act'(z); sums[0:C] += sum g, sums[C:2C] += sum g*xhat
 * (double[2*C], caller zero-fills); (2) dy = gamma*rstd*(g - sum_g/M - xhat*sum_gx/M),
 * dgamma = sum_gx, dbeta = sum_g (overwritten unless accumulate). */
int avc_bn_act_bwd_reduce(const float* dz, const float* z, const float* y, const float* mean, const float* rstd,
                          double* sums, int M, int C, int act, void* stream);
int avc_bn_act_bwd_apply(const float* dz, const float* z, const float* y, const float* mean, const float* rstd,
                         const float* gamma, const double* sums, float* dy, float* dgamma, float* dbeta,
                         int M, int C, int act, int accumulate, void* stream);
/* The same two passes with the activation RECOMPUTED from y (z = act((y-mean)*rstd*gamma+beta), the forward's own
 * expression) instead of read back: one (M, C) read less per pass.  z may be NULL when C % 4 == 0 and all tensors
 * are 16-byte aligned; otherwise it is required and read.  apply_y writes fp32 dy and/or a 16-bit copy dy16
 * (fmt16: AVC_FMT_BF16 / AVC_FMT_FP16), either may be NULL. */
int avc_bn_act_bwd_reduce_y(const float* dz, const float* z, const float* y, const float* mean, const float* rstd,
                            const float* gamma, const float* beta, double* sums, int M, int C, int act, void* stream);
int avc_bn_act_bwd_apply_y(const float* dz, const float* z, const float* y, const float* mean, const float* rstd,
                           const float* gamma, const float* beta, const double* sums, float* dy, void* dy16, int fmt16,
                           float* dgamma, float* dbeta, int M, int C, int act, int accumulate, void* stream);
/* column sums of x (M, C) -> out (C), used for Linear / LSTM bias gradients.
 * out_mode 0: out[c]; 2: LSTM un-permute (c = u*4+g -> g*H+u), written to BOTH out and out2
 * when out2 != NULL (b_ih and b_hh receive the same gradient). */
int avc_colsum(const float* x, int ldx, int M, int C, float* out, float* out2, int out_mode, int accumulate,
               void* workspace, size_t workspace_bytes, void* stream);

/* avc_colsum over a 16-bit (AVC_FMT_BF16 / AVC_FMT_FP16) matrix: C, ldx multiples of 4 (half mode: LSTM bias gradients
 * from the bf16 copy of dP, the same operand the weight-gradient GEMMs read). */
int avc_colsum16(const void* x16, int fmt, int ldx, int M, int C, float* out, float* out2, int out_mode, int accumulate,
                 void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * LSTM recurrences, nn.LSTM at model_vc_mel.py:61/:73 (encoder BiLSTM), :90/:111 (lstm1),
 * :104/:118 (lstm2).  P = x W_ih^T + b_ih + b_hh is computed by avc_gemm_nt_taps with the
 * interleaved packing, so P, gates and dP use column index u*4+g (g: 0=i,1=f,2=g,3=o).
 *   P      (nB, T, 4H)  pre-activations from the input projection
 *   Whh_p  (4H, H)      interleaved rows (forward);  Whh_pT (H, 4H) its transpose (backward)
 *   h_seq  (nB, T, H) with row stride ldh (so fwd/bwd directions can share one (B,T,2H) buffer)
 *   gates  (nB, T, 4H)  activated gates saved for backward;  c_seq (nB, T, H) cell states
 * reverse = 1 walks t = T-1 .. 0 (the *_reverse direction).  h0 = c0 = 0.  reverse = 2 (H <= 64 only): both
 * directions of a BiLSTM in ONE launch; P, Whh_p, gates, c_seq, dP are stacked [2][...] (forward first) and
 * direction d reads/writes h_seq / dH at column offset d*H of a shared (nB,T,2H) buffer (ldh = lddh = 2H).
 * reverse = 3: as 2, except that P and dP are ONE (nB, T, 8H) tensor each, [direction 0 gates | direction 1 gates] per row --
 * the output of a single N = 8H input-projection GEMM / the operand of single dW_ih (N = 8H) and dX (K = 8H) GEMMs.
 * AVC_PREC_FP32: H <= 64 runs the whole sequence in one launch (W_hh in shared memory), larger H
 * one launch per step.  AVC_PREC_BF16 with 128 <= H <= 1024, H % 64 == 0: ONE persistent cooperative
 * launch per layer-direction -- W_hh slices resident in shared memory across the SMs, tcgen05 MMAs
 * into TMEM, fused gate math, per-step release/acquire exchange of h_t (or dG_t) in bf16; cell state
 * and all saved tensors stay fp32.  Workspace sizes come from the *_workspace_bytes queries.
 */
int avc_lstm_seq_fwd(const float* P, const float* Whh_p, float* h_seq, int ldh, float* gates, float* c_seq,
                     int nB, int T, int H, int reverse, int prec,
                     void* workspace, size_t workspace_bytes, void* stream);
size_t avc_lstm_fwd_workspace_bytes(int nB, int T, int H, int prec);
/* BPTT: dH (nB,T,H) ld lddh = gradient w.r.t. h_seq from above -> dP (nB,T,4H). */
int avc_lstm_seq_bwd(const float* dH, int lddh, const float* Whh_p, const float* Whh_pT, const float* gates, const float* c_seq,
                     float* dP, int nB, int T, int H, int reverse, int prec,
                     void* workspace, size_t workspace_bytes, void* stream);
size_t avc_lstm_bwd_workspace_bytes(int nB, int T, int H, int prec);

/* ---------------------------------------------------------------------------------------
 * "half" mode (AVC_PREC_HALF): the GEMM operands already live in HBM as 16-bit copies that the PRODUCING kernels emit
 * next to (or instead of) their fp32 results, so the tensor cores run at the 16-bit rate with no staging pass:
 *   forward operands (activations, weights): fp16 -- 10 mantissa bits like tf32, half the bytes, twice the MMA rate;
 *   gradient operands (dY, dP):              bf16 -- fp32's exponent range, no loss scaling needed;
 *   both operands of one GEMM must share a format (kind::f16 traps on mixed bf16 x fp16 -- measured), so backward
 *   GEMMs take bf16 copies of the forward activations made on the fly.  Accumulation, statistics, saved state: fp32.
 * Operand format codes: AVC_FMT_FP32 (staged internally to `half_fmt`), AVC_FMT_BF16, AVC_FMT_FP16.  16-bit operands need
 * 16-byte aligned pointers and leading dimensions that are multiples of 8 elements.
 */
int avc_gemm_nt_taps_h(const void* A, int a_fmt, int lda, const float* W, const float* bias, float* C, int ldc,
                       int nB, int T, int N, int K, int ntaps, int shift0, double* chan_stats, int accumulate, int half_fmt,
                       void* workspace, size_t workspace_bytes, void* stream);
size_t avc_gemm_nt_h_workspace_bytes(int nB, int T, int N, int K, int ntaps, int a_fmt);
int avc_gemm_tn_taps_h(const void* dY, int y_fmt, int ldy, const void* X, int x_fmt, int ldx, float* dW,
                       int nB, int T, int N, int K, int ntaps, int shift0, int out_mode, int accumulate, int half_fmt,
                       void* workspace, size_t workspace_bytes, void* stream);
size_t avc_gemm_tn_h_workspace_bytes(int nB, int T, int N, int K, int ntaps, int y_fmt, int x_fmt);
/* dst16 (M, C) ld ldd <- src fp32 (M, C) ld lds */
int avc_cast16(const float* src, int lds, void* dst, int ldd, size_t M, int C, int fmt, void* stream);
/* avc_bn_act_fwd that also writes the 16-bit operand copies (C % 4 == 0): z16 in fmt16 (forward operand of the next GEMM)
 * and, when z16b != NULL, a second copy in fmt16b (the bf16 operand of the next layer's weight-gradient GEMM).
 * z (fp32) may be NULL when no consumer reads the activation in fp32. */
int avc_bn_act_fwd_h(const float* y, const float* mean, const float* rstd, const float* gamma, const float* beta,
                     const float* residual, float* z, void* z16, int fmt16, void* z16b, int fmt16b, int M, int C, int act,
                     void* stream);
/* avc_bn_act_bwd_apply that writes dy16 (and fp32 dy only when dy != NULL) */
int avc_bn_act_bwd_apply_h(const float* dz, const float* z, const float* y, const float* mean, const float* rstd,
                           const float* gamma, const double* sums, float* dy, void* dy16, int fmt16, float* dgamma,
                           float* dbeta, int M, int C, int act, int accumulate, void* stream);
/* persistent recurrences (128 <= H <= 1024, H % 64 == 0) with a 16-bit side output: h16 (nB,T,H) / dP16 (nB,T,4H, bf16) */
/* w_fmt: AVC_FMT_FP32 = Whh_p / Whh_pT are the fp32 packings of avc_pack_lstm_weight (converted per call);
 *        AVC_FMT_BF16 = the bf16 packings of avc_pack_lstm_weight_h (ld = H resp. 4H), read in place. */
/* h16b (optional, may be NULL): a bf16 copy of h next to h16 -- the operand of the dW_hh / next layer's dW_ih GEMMs.
 * gates and c_seq may BOTH be NULL (inference, conversion.py:91-92 under no_grad): nothing is saved for BPTT. */
int avc_lstm_seq_fwd_h(const float* P, const void* Whh_p, int w_fmt, float* h_seq, int ldh, float* gates, float* c_seq, void* h16,
                       int fmt16, void* h16b, int nB, int T, int H, int reverse, void* workspace, size_t workspace_bytes, void* stream);
/* (dP, the fp32 gate-gradient tensor, may be NULL: only dP16 is then written) */
int avc_lstm_seq_bwd_h(const float* dH, int lddh, const void* Whh_pT, int w_fmt, const float* gates, const float* c_seq, float* dP,
                       void* dP16, int nB, int T, int H, int reverse, void* workspace, size_t workspace_bytes, void* stream);
/* Weight packers that write the 16-bit (or fp32: fmt 0) layouts the tensor-core GEMMs read IN PLACE, with a leading
 * dimension (ldf >= Cin, ldd >= Cout, ldp >= I, ldpT >= 4H; multiples of 8 for 16-bit) whose tail columns are zero-filled:
 *   w_fwd [tap][Cout][ldf], w_dgrad [taps-1-tap][Cin][ldd];  out_p [u*4+g][ldp], out_pT [k][ldpT].  Either output may be NULL. */
int avc_pack_conv_weight_h(const float* w, void* w_fwd, int ldf, int fmt_f, void* w_dgrad, int ldd, int fmt_d, int Cout, int Cin,
                           int ntaps, void* stream);
int avc_pack_lstm_weight_h(const float* w, void* out_p, int ldp, int fmt_p, void* out_pT, int ldpT, int fmt_pT, int H, int I,
                           void* stream);
/* avc_gemm_nt_taps_h with W already packed as 16-bit [ntaps][N][ldw] (w_fmt bf16/fp16, ldw % 8 == 0): no per-call
 * staging of W.  A: fp32 (staged to w_fmt) or 16-bit of the SAME format.  Workspace: avc_gemm_nt_h_workspace_bytes. */
int avc_gemm_nt_taps_hw(const void* A, int a_fmt, int lda, const void* W16, int w_fmt, int ldw, const float* bias, float* C, int ldc,
                        int nB, int T, int N, int K, int ntaps, int shift0, double* chan_stats, int accumulate,
                        void* workspace, size_t workspace_bytes, void* stream);

/* Profiling hook: when set to a device buffer of >= 16*T uint64, CTA 0 of the next persistent recurrence
 * launches records %globaltimer stamps per step (slot layout in lstm_tc.cu).  NULL disables it. */
void avc_debug_set_trace(unsigned long long* device_buffer);

/* ---------------------------------------------------------------------------------------
 * Glue that the reference does with squeeze/transpose/expand/cat/slicing.
 */
/* model_vc_mel.py:64-66: out (B,T,Cx+E) = [x (B,T,Cx) ld ldx | e (B,E) broadcast over T]. */
int avc_concat_bcast(const float* x, int ldx, const float* e, float* out, int nB, int T, int Cx, int E, void* stream);
/* model_vc_mel.py:74-79,:201: codes (B, T/f, 2n) from enc (B,T,2n):
 *   codes[b,j,:n] = enc[b, j*f+f-1, :n];  codes[b,j,n:] = enc[b, j*f, n:]. */
int avc_codes_fwd(const float* enc, float* codes, int nB, int T, int n, int f, void* stream);
/* scatter-add of dcodes back into denc (B,T,2n), which the caller zero-fills. */
int avc_codes_bwd(const float* dcodes, float* denc, int nB, int T, int n, int f, void* stream);
/* model_vc_mel.py:186-192: out (B,T,2n+E) = [codes[b, t/f, :] | c_trg[b,:]]. */
int avc_upsample_concat_fwd(const float* codes, const float* c_trg, float* out, int nB, int T, int n2, int f, int E, void* stream);
/* dcodes[b,j,:] (+)= sum_{t in group j} dout[b,t,:n2]   (dout ld = lddo). */
int avc_upsample_concat_bwd(const float* dout, int lddo, float* dcodes, int nB, int T, int n2, int f, int accumulate, void* stream);
/* strided 2-D copy dst(M, C) ld lddst <- src(M, C) ld ldsrc (slice of the concat gradient). */
int avc_copy2d(const float* src, int ldsrc, float* dst, int lddst, int M, int C, void* stream);

/* ---------------------------------------------------------------------------------------
 * Losses, solver_encoder.py:230,:233 (F.mse_loss) and :236 (F.l1_loss), reduction='mean'.
 * out[0] = mean((a-b)^2) or mean(|a-b|).  scratch: double[1], caller zero-fills.
 * bwd: da = scale * 2(a-b)/n  or  scale * sign(a-b)/n  with scale read from *gout (device);
 * db = -da when db != NULL.  `accumulate` adds into da/db. */
int avc_mse_loss_fwd(const float* a, const float* b, size_t n, double* scratch, float* out, void* stream);
int avc_l1_loss_fwd(const float* a, const float* b, size_t n, double* scratch, float* out, void* stream);
int avc_loss_bwd(const float* a, const float* b, size_t n, const float* gout, int is_l1,
                 float* da, float* db, int accumulate, void* stream);

/* out[m, :] = x[m, :] / ||x[m, :]||_2   (model_bl.py:18-19, the d-vector's final normalisation; no epsilon, like the reference) */
int avc_l2_normalize_rows(const float* x, float* out, int M, int C, void* stream);

/* ---------------------------------------------------------------------------------------
 * Crop loader, data_loader.py:61-80 (`Utterances.__getitem__`) + default collate (:90-102), over a corpus resident in HBM:
 *   corpus    (sum_i F_i, n_bins) fp32, utterance u = rows utt_row0[u] .. utt_row0[u] + utt_len[u] - 1
 *   per crop b: utterance sel_utt[b]; F > T: frames sel_left[b] .. sel_left[b]+T-1 (:74-76); F < T: the utterance followed by
 *   zero frames (:70-73); F == T: as is.  e_out[b] = emb_table[sel_spk[b]] (dim_emb floats, :65).
 * The random draws (np.random.randint(2, len(list_uttrs)), np.random.randint(F - len_crop)) stay on the host so that the
 * reference's numpy stream is reproduced; x_out (B, T, n_bins), e_out (B, dim_emb). */
int avc_crop_batch(const float* corpus, const long long* utt_row0, const int* utt_len, const int* sel_utt, const int* sel_left,
                   const float* emb_table, const int* sel_spk, float* x_out, float* e_out, int B, int T, int n_bins,
                   int dim_emb, void* stream);

/* ---------------------------------------------------------------------------------------
 * Optimizer step, solver_encoder.py:130 (torch.optim.Adam(G.parameters(), lr): betas (0.9, 0.999), eps 1e-8, no weight
 * decay, no amsgrad) and :300 (.step()).  ONE launch updates every parameter tensor:
 *   g' = grad_scale * g;  m += (1-beta1)(g' - m);  v = beta2 v + (1-beta2) g'^2;
 *   p -= lr/(1-beta1^step) * m / (sqrt(v)/sqrt(1-beta2^step) + eps)      (torch's _multi_tensor_adam, fp32)
 * table : device array of n_tensors records {float* p; const float* g; float* m; float* v; uint64 n;} (40 bytes each);
 * chunks: device array of int2 {tensor index, chunk index within the tensor}, one entry per avc_adam_chunk_elems()
 *         elements of every tensor (static for a given parameter list);  step is 1-based.
 * grad_scale folds the 1/world_size of a summed data-parallel gradient (1.0 otherwise). */
int avc_adam_chunk_elems(void);
int avc_adam_step(const void* table, const void* chunks, int nchunks, double lr, double beta1, double beta2, double eps,
                  int step, float grad_scale, void* stream);

/* solver_encoder.py:168-177 Solver.model_EMA (called before every checkpoint, :334): every parameter p is overwritten with
 *   fl(fl(ema * p) + fl((1 - ema) * p))      (fp32, two products and one sum, no fma -- the reference's three ATen passes)
 * in ONE launch over all tensors.  table / chunks as for avc_adam_step (only the p and n columns are read). */
int avc_ema_blend(const void* table, const void* chunks, int nchunks, double ema, void* stream);

/* ---------------------------------------------------------------------------------------
 * make_spect front-end, make_spect.py:72-83 (spmel branch) + :30-48:
 *   y = filtfilt(butter(5, 30 Hz HP), wav) [fp64, scipy odd-extension padlen 18, lfilter_zi]
 *   wav' = 0.96*y + (dither - 0.5)*1e-6
 *   D = |rfft(hann_periodic * frames(reflect_pad(wav', 512), 1024, hop 256))|
 *   S = clip((20*log10(max(1e-5, D @ mel_basis)) - 16 + 100)/100, 0, 1)
 * wav, dither: (n_utt, max_len) float32 rows (dither = the uniform [0,1) draws of
 * make_spect.py:76, supplied by the host so the stream matches); lengths[i] <= max_len;
 * out: (n_utt, max_frames, 80) float32, frames beyond 1 + lengths[i]/256 are zero-filled
 * (= conversion.py:40-44 pad_seq).  mel_basis: (513, 80) float32 (librosa.filters.mel^T).
 * filt: double[18] = the Butterworth filter as 3 second-order sections, scipy layout (b0,b1,b2,1,a1,a2) per row
 * (scipy.signal.tf2sos(b, a)); zi: double[6] = scipy.signal.sosfilt_zi(sos).  The cascade realisation is what makes a
 * chunk-parallel IIR numerically possible; it matches scipy's (b, a)-form filtfilt to ~1e-6 on the waveform and to
 * < 1e-7 on the log-mel output.  lengths[i] must be >= 513 (numpy's multi-bounce reflection is not reproduced).
 * workspace: avc_logmel_workspace_bytes(n_utt, max_len).
 */
int avc_logmel_frontend(const float* wav, const float* dither, const int* lengths, int n_utt, int max_len,
                        const float* mel_basis, const double* filt, const double* zi,
                        float* out, int max_frames, void* workspace, size_t workspace_bytes, void* stream);
size_t avc_logmel_workspace_bytes(int n_utt, int max_len);
/* make_spect.py:84-86 (model_type 'stft'): same pipeline up to D, then S = clip((20*log10(max(1e-5, D)) - 16 + 100)/100, 0, 1)
 * on the 513 magnitudes themselves.  out: (n_utt, max_frames, 513) float32, frame-major (the reference saves the transpose);
 * same arguments and workspace as avc_logmel_frontend (mel_basis is only used to build the shared tables). */
int avc_logstft_frontend(const float* wav, const float* dither, const int* lengths, int n_utt, int max_len,
                         const float* mel_basis, const double* filt, const double* zi,
                         float* out, int max_frames, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * Waveform variant, model_vc_wav.py:11-102 (GeneratorWav, ConvTasNetEncoder / ConvTasNetDecoder) and the 'wav' branch
 * of the step, solver_encoder.py:264-290.  The convolutions themselves run on avc_gemm_{nt,tn}_taps:
 *   Conv1d(1 -> N, k = taps*S, stride S) on (B, L)             = taps-tap convolution over the view (B, L/S, S)
 *   ConvTranspose1d(N -> 1, k = taps*S, stride S)              = its adjoint (the data-gradient form)
 *   Conv1d / ConvTranspose1d(N -> N, k = 3, s = 1, p = 1)      = 3-tap forms with shift0 = -1
 * and these entry points supply what sits between them.
 */
/* nn.PReLU() (one slope, model_vc_wav.py:24,:46): p = y > 0 ? y : slope[0]*y over (M, C); if chan_stats != NULL
 * (double[2*C], caller zero-fills) the batch statistics the following BatchNorm1d needs (sum, sum of squares of p). */
int avc_prelu_fwd(const float* y, const float* slope, float* p, double* chan_stats, int M, int C, void* stream);
/* dy = dp * (y > 0 ? 1 : slope);  dslope[0] (+)= sum_{y <= 0} dp*y.  scratch: double[1], caller zero-fills. */
int avc_prelu_bwd(const float* dp, const float* y, const float* slope, float* dy, float* dslope, int accumulate,
                  double* scratch, size_t n, void* stream);
/* out[0] (+)= sum of all n elements (gradient of the synthesis layer's single bias).  scratch: double[1], zero-filled. */
int avc_sum_all(const float* x, size_t n, double* scratch, float* out, int accumulate, void* stream);
/* dst (nB, Tdst, C) <- src (nB, Tsrc, C): frames t < min(Tsrc, Tdst) copied, remaining frames of dst zero-filled
 * (brings the L/S waveform blocks and the L/S - taps + 1 frames of the filterbank onto one frame axis). */
int avc_copy_rows3d(const float* src, int Tsrc, float* dst, int Tdst, int nB, int C, void* stream);
/* out[b][a][c] = in[a][b][c]: filterbank weight (N, 1, taps*S) <-> packed [tap][N][S]. */
int avc_permute021(const float* in, float* out, int A, int B, int C, void* stream);
/* SI-SNR term, solver_encoder.py:276-283, est = x_identic (nB, L), tgt = x_real (nB, L):
 *   out[0] = -mean_b 10*log10( sum_t scaled^2 / sum_t (est - scaled)^2 ),  scaled = (sum est*tgt) * tgt / (sum tgt^2)
 * saved: float[4*nB] (dot, energy, numerator, denominator per utterance) for the backward; scratch: double[4*nB], zero-filled.
 * bwd: dest (+)= gout[0] * d out / d est   (tgt carries no gradient in the reference's step). */
int avc_sisnr_fwd(const float* est, const float* tgt, int nB, int L, double* scratch, float* saved, float* out, void* stream);
int avc_sisnr_bwd(const float* est, const float* tgt, const float* saved, const float* gout, int nB, int L,
                  float* dest, int accumulate, void* stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* AUTOVC_B200_H_ */
