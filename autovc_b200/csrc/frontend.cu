// make_spect front-end on the GPU: make_spect.py:72-83 (spmel branch) + butter_highpass :30-34 + pySTFT :36-48.
//
//   stage 1  filtfilt (fp64, scipy semantics: odd extension by padlen=18, steady-state initial state,
//            forward pass then the same filter run backwards; realised as a cascade of second-order
//            sections so that it can run chunk-parallel -- see below)                              :74
//   stage 2  wav' = 0.96*y + (dither - 0.5)*1e-6, stored fp32                                      :76
//   stage 3  per frame: reflect-pad(512) framing, periodic Hann, 1024-point FFT (two real frames
//            per complex transform), |.|, mel projection (only the non-zero band of each filter),
//            20*log10(max(1e-5,.)) - 16, (x+100)/100 clipped to [0,1]                              :78-83
// HBM-bound by design: every stage streams its input once; the FFT, window, mel and log work stay in
// shared memory / registers.
#include <stdlib.h>

#include "common.cuh"

namespace avc {

constexpr int FE_NFFT = 1024, FE_HOP = 256, FE_BINS = 513, FE_MELS = 80, FE_PADLEN = 18, FE_REFLECT = 512;
constexpr int FE_FRAMES_PER_CTA = 8;   // 4 complex FFTs per CTA, one warp each
constexpr int FE_SEG = 8;              // bins per work item of the mel projection
constexpr int FE_THREADS = 32 * (FE_FRAMES_PER_CTA / 2);

struct FeTables {           // lives at the head of the workspace
  float2 tw2[FE_NFFT];      // tw2[c*32 + b] = exp(-2*pi*i*(b*c)/1024): the inter-pass twiddles of the 32 x 32 decomposition
  float win[FE_NFFT];       // periodic Hann
  int2 band[FE_MELS];       // [first, last+1) non-zero FFT bin of each mel filter
  // balanced schedule of the mel projection: every filter's bin range cut into segments of <= FE_SEG bins; filter m owns the
  // segments [seg0[m], seg0[m+1]) (consecutive, so its partial sums are added in a fixed order)
  int4 seg[256];            // (filter, first bin, last bin + 1, -)
  int seg0[FE_MELS + 1];
  alignas(16) float segw[256][FE_SEG];  // the segment's filter weights, segment-major and zero-padded to FE_SEG bins: two 16-byte loads per
                            // segment instead of <= 8 scattered reads of mel_basis, and a branch-free 8-bin inner product
};

__global__ void fe_tables_kernel(const float* __restrict__ mel_basis, FeTables* __restrict__ tb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < FE_NFFT) {
    double s, c;
    sincospi(-2.0 * (double)(((i >> 5) * (i & 31)) & (FE_NFFT - 1)) / (double)FE_NFFT, &s, &c);
    tb->tw2[i] = make_float2((float)c, (float)s);
  }
  if (i < FE_NFFT) tb->win[i] = (float)(0.5 - 0.5 * cospi(2.0 * (double)i / (double)FE_NFFT));
  if (i < FE_MELS) {
    int lo = FE_BINS, hi = 0;
    for (int k = 0; k < FE_BINS; ++k)
      if (mel_basis[k * FE_MELS + i] != 0.f) {
        lo = min(lo, k);
        hi = max(hi, k + 1);
      }
    if (hi == 0) lo = 0;
    tb->band[i] = make_int2(lo, hi);
  }
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int n = 0;
    for (int m = 0; m < FE_MELS; ++m) {
      tb->seg0[m] = n;
      const int2 bd = tb->band[m];
      for (int k = bd.x; k < bd.y && n < 256; k += FE_SEG) {
        for (int j = 0; j < FE_SEG; ++j) tb->segw[n][j] = (k + j < bd.y) ? mel_basis[(k + j) * FE_MELS + m] : 0.f;
        tb->seg[n++] = make_int4(m, k, min(bd.y, k + FE_SEG), 0);
      }
    }
    tb->seg0[FE_MELS] = n;
  }
}

// --- stage 1+2: filtfilt as two chunk-parallel IIR sweeps (fp64) ----------------------------------------------
// The direct-form-II-transposed recurrence is linear in its state z:  z' = A z + B x,  y = z[0] + b0 x, so a chunk's END state
// is  A^L z_start + (end state from z = 0): chunks run independently from a zero state, a scan over the chunk states supplies
// the true start states, and a second pass re-runs every chunk from its true start state.  Mathematically identical to the
// sequential filter (rounding differs at the 1e-16 level).  (r01 did this in three launches per sweep with every thread walking
// its own 256-sample chunk straight from global memory -- 5.7 ms per 1024 utterances; the fused tile kernel below replaced it.)

// Cascade of 3 second-order sections, each direct-form-II-transposed (scipy.signal.sosfilt arithmetic):
//   y = b0 u + z0;  z0' = b1 u + z1 - a1 y;  z1' = b2 u - a2 y;  next section's input u = y.
// The companion (b, a) form has the same poles but a state-transition matrix whose powers reach 1e8 before they
// decay (5 poles clustered at |z| ~ 0.99), which makes any chunked carry numerically useless; the cascade's
// transition matrix stays O(50).  scipy's tf-form filtfilt and this cascade agree to ~1e-6 on the waveform and
// to < 1e-7 on the log-mel output (checked against the reference's bundled goldens).
constexpr int FE_NSEC = 3, FE_NST = 2 * FE_NSEC;
// Storage type of the forward sweep's output between the two sweeps.  The ARITHMETIC of both sweeps is fp64 like scipy's; rounding
// the stored intermediate to fp32 (6e-8 relative, ~6e-9 absolute on a 0.1-amplitude waveform) is two orders of magnitude below the
// 7e-7 by which scipy's own transfer-function filtfilt differs from the exact filter, and takes 8 of the 32 bytes per sample that
// the front-end moves through HBM off the bus.
typedef float fe_y1_t;
struct Df2t {
  double c[FE_NSEC][5];   // b0 b1 b2 a1 a2 per section
  double z[FE_NST];
  __device__ __forceinline__ double step(double x) {
    double u = x;
#pragma unroll
    for (int s = 0; s < FE_NSEC; ++s) {
      const double y = fma(c[s][0], u, z[2 * s]);
      z[2 * s] = fma(-c[s][3], y, fma(c[s][1], u, z[2 * s + 1]));
      z[2 * s + 1] = fma(-c[s][4], y, c[s][2] * u);
      u = y;
    }
    return u;
  }
  // filt: 3 rows of scipy sos coefficients (b0, b1, b2, a0 = 1, a1, a2)
  __device__ __forceinline__ void load(const double* __restrict__ filt) {
#pragma unroll
    for (int s = 0; s < FE_NSEC; ++s) {
      c[s][0] = filt[6 * s + 0];
      c[s][1] = filt[6 * s + 1];
      c[s][2] = filt[6 * s + 2];
      c[s][3] = filt[6 * s + 4];
      c[s][4] = filt[6 * s + 5];
    }
  }
};

// odd extension of x (length n) by FE_PADLEN at both ends; index i in [0, n + 2*padlen)
__device__ __forceinline__ double odd_ext(const float* __restrict__ x, int n, int i) {
  if (i < FE_PADLEN) return 2.0 * (double)x[0] - (double)x[FE_PADLEN - i];
  if (i < FE_PADLEN + n) return (double)x[i - FE_PADLEN];
  return 2.0 * (double)x[n - 1] - (double)x[n - 2 - (i - FE_PADLEN - n)];
}

// input sample i of the sweep: forward sweep reads the odd-extended waveform, backward sweep reads the forward
// output in reverse order
template <bool BACKWARD>
__device__ __forceinline__ double sweep_input(const float* __restrict__ x, const fe_y1_t* __restrict__ y1, int n, int ne, int i) {
  return BACKWARD ? (double)y1[ne - 1 - i] : odd_ext(x, n, i);
}

// --- stage 1+2, fused: one CTA per utterance walks the sweep tile by tile ------------------------------------------
// A CTA of 256 threads owns an utterance and processes tiles of 256 chunks x 32 samples = 8192 samples (first version:
// 128 x 64 -- same shared memory, half the warps and twice the serial chunk length; ncu: 11% warps active, 0.88 IPC):
//   load    the tile into shared memory as fp64 with coalesced (forward) / reversed-coalesced (backward) accesses,
//           one pad word per chunk so that the per-thread walks below are bank-conflict free;
//   pass 1  thread t runs its chunk from a zero state -> e[t]; thread 0 adds A^32 * (carry-in of the tile);
//   scan    Kogge-Stone over the 256 chunk states: w_i <- P[k] w_{i-2^k} + w_i with P[k] = A^(32*2^k) (all chunks share the
//           linear part, so a level is ONE 6x6 mat-vec per thread); w_i is then the true END state of chunk i;
//   pass 2  thread t re-runs its chunk from w_{t-1} and overwrites the tile with the outputs;
//   store   coalesced: forward y1 (fp64), backward 0.96*y + dither as fp32 (reversed).
// Every input is read once and every output written once per sweep.
constexpr int FE_TCH = 32;                      // samples per thread-chunk
constexpr int FE_TSH = 5;                       // log2(FE_TCH)
constexpr int FE_TNT = 256;                     // threads = chunks per tile
constexpr int FE_TILE = FE_TCH * FE_TNT;        // 8192 samples
constexpr int FE_TLEV = 8;                      // log2(FE_TNT)

// PW[k] = A^(FE_TCH * 2^k), k = 0..FE_TLEV-1 (row-major 6x6 each).  Single thread.
__global__ void fe_tile_powers_kernel(const double* __restrict__ filt, double* __restrict__ PW) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  constexpr int N = FE_NST;
  double M[N * N], R[N * N];
  Df2t f;
  f.load(filt);
  for (int k = 0; k < N; ++k) {
    for (int i = 0; i < N; ++i) f.z[i] = (i == k) ? 1.0 : 0.0;
    f.step(0.0);
    for (int i = 0; i < N; ++i) M[i * N + k] = f.z[i];
  }
  auto square = [&]() {
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < N; ++j) {
        double acc = 0.0;
        for (int k = 0; k < N; ++k) acc = fma(M[i * N + k], M[k * N + j], acc);
        R[i * N + j] = acc;
      }
    for (int i = 0; i < N * N; ++i) M[i] = R[i];
  };
  for (int L = FE_TCH; L > 1; L >>= 1) square();          // A^64
  for (int k = 0; k < FE_TLEV; ++k) {
    for (int i = 0; i < N * N; ++i) PW[k * N * N + i] = M[i];
    square();
  }
}

template <bool BACKWARD>
__global__ void __launch_bounds__(FE_TNT)
fe_iir_sweep_kernel(const float* __restrict__ wav, const float* __restrict__ dither, const int* __restrict__ lengths,
                    int max_len, const double* __restrict__ filt, const double* __restrict__ zi,
                    const double* __restrict__ PW, fe_y1_t* __restrict__ y1buf, float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int N = FE_NST;
  double* wst = reinterpret_cast<double*>(smem_raw);                  // [FE_TNT][N]  chunk states
  double* pw = wst + FE_TNT * N;                                  // [FE_TLEV][N*N]
  double* carry = pw + FE_TLEV * N * N;                               // [N] carry-in of the current tile
  // the tile holds what the sweep reads and writes in HBM -- fp32 values (the waveform / the fp32 intermediate in, the fp32
  // intermediate / the fp32 filtered waveform out); the filter states and all arithmetic are fp64.  As an fp64 array it took
  // 67.6 of the block's 94.5 KB and held the kernel at 2 blocks per SM (25 % occupancy, 37 % of the FP64 pipe: r02 ncu).
  float* tile = reinterpret_cast<float*>(carry + N);                   // [FE_TNT][FE_TCH + 1]
  const int u = blockIdx.x, tid = threadIdx.x;
  const int n = lengths[u];
  const float* x = wav + (size_t)u * max_len;
  const float* dz = dither + (size_t)u * max_len;
  float* o = out + (size_t)u * max_len;
  fe_y1_t* y1 = y1buf + (size_t)u * (max_len + 2 * FE_PADLEN);
  if (n <= FE_PADLEN) {   // scipy raises for such inputs; emit dither-only silence deterministically
    if (BACKWARD)
      for (int i = tid; i < n; i += FE_TNT) o[i] = (float)(((double)dz[i] - 0.5) * 1e-6);
    return;
  }
  const int ne = n + 2 * FE_PADLEN;
  for (int i = tid; i < FE_TLEV * N * N; i += FE_TNT) pw[i] = PW[i];
  if (tid < N) carry[tid] = zi[tid] * sweep_input<BACKWARD>(x, y1, n, ne, 0);   // scipy: zi * first sample of the sweep's input
  Df2t f;
  f.load(filt);
  float* mine = tile + tid * (FE_TCH + 1);
  for (int i0 = 0; i0 < ne; i0 += FE_TILE) {
    __syncthreads();                                      // previous tile fully stored; carry / pw visible
    // interior tiles (all but the first and the last one or two): branch-free, 16 independent loads in flight per thread --
    // with a rolled loop the 64 dependent-latency loads per thread made a tile take ~50 us
    const bool interior = i0 + FE_TILE <= ne && (BACKWARD || (i0 >= FE_PADLEN && i0 + FE_TILE <= FE_PADLEN + n));
    if (interior) {
      const float* xs = x + (i0 - FE_PADLEN);
      const fe_y1_t* ys = y1 + (ne - 1 - i0);
#pragma unroll 16
      for (int q = 0; q < FE_TCH; ++q) {
        const int s = q * FE_TNT + tid;
        tile[s + (s >> FE_TSH)] = BACKWARD ? (float)ys[-s] : xs[s];
      }
    } else {
#pragma unroll 8
      for (int q = 0; q < FE_TCH; ++q) {
        const int s = q * FE_TNT + tid;
        const int i = i0 + s;
        tile[s + (s >> FE_TSH)] = i < ne ? (float)sweep_input<BACKWARD>(x, y1, n, ne, i) : 0.f;   // zeros past the end: pure state decay
      }
    }
    __syncthreads();
    // pass 1: zero-state end state of my chunk
#pragma unroll
    for (int k = 0; k < N; ++k) f.z[k] = 0.0;
#pragma unroll 4
    for (int k = 0; k < FE_TCH; ++k) f.step((double)mine[k]);
    double w[N];
#pragma unroll
    for (int k = 0; k < N; ++k) w[k] = f.z[k];
    if (tid == 0) {                                       // chunk 0 starts from the tile's carry-in: end = A^64 carry + e_0
#pragma unroll
      for (int i = 0; i < N; ++i) {
        double acc = w[i];
#pragma unroll
        for (int k = 0; k < N; ++k) acc = fma(pw[i * N + k], carry[k], acc);
        w[i] = acc;
      }
    }
#pragma unroll
    for (int k = 0; k < N; ++k) wst[tid * N + k] = w[k];
    __syncthreads();
    // one state buffer, two barriers per level (read the neighbour, barrier, publish, barrier): the ping-pong pair cost 12 KB more
    // per block and with it the fourth resident block per SM
#pragma unroll 1
    for (int lev = 0; lev < FE_TLEV; ++lev) {
      const int d = 1 << lev;
      if (tid >= d) {
        const double* src = wst + (tid - d) * N;
        const double* P = pw + lev * N * N;
        double sv[N];
#pragma unroll
        for (int k = 0; k < N; ++k) sv[k] = src[k];
#pragma unroll
        for (int i = 0; i < N; ++i) {
          double acc = w[i];
#pragma unroll
          for (int k = 0; k < N; ++k) acc = fma(P[i * N + k], sv[k], acc);
          w[i] = acc;
        }
      }
      __syncthreads();
      if (tid >= d) {
#pragma unroll
        for (int k = 0; k < N; ++k) wst[tid * N + k] = w[k];
      }
      __syncthreads();
    }
    // pass 2: from the true start state (= end state of the previous chunk)
    if (tid == 0) {
#pragma unroll
      for (int k = 0; k < N; ++k) f.z[k] = carry[k];
    } else {
#pragma unroll
      for (int k = 0; k < N; ++k) f.z[k] = wst[(tid - 1) * N + k];
    }
#pragma unroll 4
    for (int k = 0; k < FE_TCH; ++k) mine[k] = (float)f.step((double)mine[k]);
    __syncthreads();                                      // every thread has read carry / its neighbour's state
    if (tid == FE_TNT - 1) {
#pragma unroll
      for (int k = 0; k < N; ++k) carry[k] = w[k];        // end state of the tile's last chunk
    }
    const bool interior_out = i0 + FE_TILE <= ne && (!BACKWARD || (i0 >= FE_PADLEN && i0 + FE_TILE <= FE_PADLEN + n));
    if (interior_out) {
      const int jb = (ne - 1 - i0) - FE_PADLEN;            // backward: output position of tile sample s is jb - s
#pragma unroll 16
      for (int q = 0; q < FE_TCH; ++q) {
        const int s = q * FE_TNT + tid;
        const double y = (double)tile[s + (s >> FE_TSH)];
        if (!BACKWARD) y1[i0 + s] = (fe_y1_t)y;
        else o[jb - s] = (float)(y * 0.96 + ((double)dz[jb - s] - 0.5) * 1e-6);
      }
    } else {
#pragma unroll 8
      for (int q = 0; q < FE_TCH; ++q) {
        const int s = q * FE_TNT + tid;
        const int i = i0 + s;
        const double y = (double)tile[s + (s >> FE_TSH)];
        if (!BACKWARD) {
          if (i < ne) y1[i] = (fe_y1_t)y;
        } else {
          const int j = (ne - 1 - i) - FE_PADLEN;
          if (i < ne && j >= 0 && j < n) o[j] = (float)(y * 0.96 + ((double)dz[j] - 0.5) * 1e-6);
        }
      }
    }
  }
}

// --- stage 3: framing + FFT + mel + log ---------------------------------------------------------------
// 1024-point complex FFT as 32 x 32 (n = 32a + b, k = c + 32d), ONE WARP per transform, 32 points per thread in registers:
//   pass 1  lane b: Y_b[c] = DFT32_a(x[32a + b]) * W1024^(b c)      (inputs read with lanes consecutive: conflict-free)
//   exchange through a [32][33] shared tile (row c, column b; written and read conflict-free)
//   pass 2  lane c: X[c + 32d] = DFT32_b(Y_b[c])                    (outputs written with lanes consecutive)
// Two shared-memory round trips per transform instead of the five of an in-place radix-4 loop, whose power-of-two strides
// also cost 4- to 16-way bank conflicts on data and twiddles (r01c ncu: 7.4 ms per 1024 utterances at 80% L1/TEX busy).
__constant__ float2 c_w32[16] = {
    {1.0f, -0.0f},
    {0.98078528040323043f, -0.19509032201612825f},
    {0.92387953251128674f, -0.38268343236508978f},
    {0.83146961230254524f, -0.55557023301960218f},
    {0.70710678118654757f, -0.70710678118654757f},
    {0.55557023301960229f, -0.83146961230254524f},
    {0.38268343236508984f, -0.92387953251128674f},
    {0.19509032201612833f, -0.98078528040323043f},
    {0.0f, -1.0f},
    {-0.19509032201612819f, -0.98078528040323043f},
    {-0.38268343236508973f, -0.92387953251128674f},
    {-0.55557023301960196f, -0.83146961230254546f},
    {-0.70710678118654746f, -0.70710678118654757f},
    {-0.83146961230254535f, -0.55557023301960218f},
    {-0.92387953251128674f, -0.38268343236508989f},
    {-0.98078528040323043f, -0.19509032201612861f},
};

__host__ __device__ constexpr int bitrev5(int v) {
  return ((v & 1) << 4) | ((v & 2) << 2) | (v & 4) | ((v & 8) >> 2) | ((v & 16) >> 4);
}
__device__ __forceinline__ float2 cmul(float2 a, float2 w) { return make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x); }

// forward 32-point DFT in registers (decimation in frequency): X[k] ends up in v[bitrev5(k)].  Fully unrolled: every index and
// every twiddle is a compile-time constant.
__device__ __forceinline__ void fft32(float2 (&v)[32]) {
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
#pragma unroll
    for (int g = 0; g < 32; g += 2 * half) {
#pragma unroll
      for (int j = 0; j < half; ++j) {
        const float2 a = v[g + j], b = v[g + j + half];
        v[g + j] = make_float2(a.x + b.x, a.y + b.y);
        const float2 d = make_float2(a.x - b.x, a.y - b.y);
        const int m = j * (16 / half);
        if (m == 0) v[g + j + half] = d;
        else if (m == 8) v[g + j + half] = make_float2(d.y, -d.x);      // * (-i)
        else v[g + j + half] = cmul(d, c_w32[m]);
      }
    }
  }
}

constexpr int FE_EXLD = 33;                          // padded row of the exchange tile
constexpr int FE_EX = 32 * FE_EXLD;                  // float2 per transform

constexpr int FE_PAIRS_PER_WARP = 5;                 // a warp walks 5 frame pairs: a CTA (4 warps) covers 40 consecutive frames
constexpr int FE_FRAMES_PER_BLOCK = FE_FRAMES_PER_CTA * FE_PAIRS_PER_WARP;

// reflect-padded sample j of an utterance of n samples (np.pad mode='reflect'; positions beyond one reflection read 0)
__device__ __forceinline__ float fe_reflect(const float* __restrict__ x, int n, int j) {
  if (j < 0) j = -j;
  if (j >= n) j = 2 * (n - 1) - j;
  return (j >= 0 && j < n) ? __ldg(x + j) : 0.f;
}

// Every warp is on its own: frames 2q, 2q+1 -> registers -> transform -> magnitudes -> the two frames' 160 mel outputs.
// No block barrier anywhere, so the 24 resident warps of an SM sit in different phases and hide each other's latencies
// (the block-synchronous version ran at 21% of the issue rate).
// STFT_OUT (make_spect.py:84-86, the 'stft' model type): the 513 magnitudes themselves go through the same log / clip and are
// written frame-major (n_utt, max_frames, 513) instead of the mel projection.
template <bool STFT_OUT>
__global__ void __launch_bounds__(FE_THREADS, 5)
fe_stft_mel_kernel(const float* __restrict__ sig, const int* __restrict__ lengths, int max_len,
                   const float* __restrict__ mel_basis, const FeTables* __restrict__ tb, float* __restrict__ out,
                   int max_frames) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int u = blockIdx.y;
  const int n = lengths[u];
  const int n_frames = n > FE_PADLEN ? 1 + n / FE_HOP : 0;
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float2* e = reinterpret_cast<float2*>(smem_raw) + wid * FE_EX;   // this warp's [32][33] tile; then X[k]; then (|A_k|, |B_k|)
  const float* x = sig + (size_t)u * max_len;
  const float* win = tb->win;
  const float2* tw2 = tb->tw2;
  for (int g = 0; g < FE_PAIRS_PER_WARP; ++g) {
    const int f0 = blockIdx.x * FE_FRAMES_PER_BLOCK + (wid * FE_PAIRS_PER_WARP + g) * 2;   // frames f0 (real part), f0+1 (imaginary)
    if (f0 >= max_frames) break;
    constexpr int NOUT = STFT_OUT ? FE_BINS : FE_MELS;
    float* o = out + ((size_t)u * max_frames + f0) * NOUT;
    const int nout = min(2, max_frames - f0) * NOUT;
    if (f0 >= n_frames) {   // zero padding frames (conversion.py:40-44 pad_seq)
      for (int i = lane; i < nout; i += 32) o[i] = 0.f;
      continue;
    }
    float2 v[32];
    const int j0 = f0 * FE_HOP - FE_REFLECT;
    if (j0 >= 0 && j0 + FE_HOP + FE_NFFT <= n) {
      const float* c0 = x + j0;
#pragma unroll
      for (int a = 0; a < 32; ++a) {
        const int k = 32 * a + lane;
        const float w = __ldg(win + k);
        v[a] = make_float2(w * __ldg(c0 + k), w * __ldg(c0 + FE_HOP + k));
      }
    } else {
#pragma unroll
      for (int a = 0; a < 32; ++a) {
        const int k = 32 * a + lane;
        const float w = __ldg(win + k);
        v[a] = make_float2(w * fe_reflect(x, n, j0 + k), w * fe_reflect(x, n, j0 + FE_HOP + k));
      }
    }
    fft32(v);
    __syncwarp();                                          // the previous pair's mel loop is done with the tile
#pragma unroll
    for (int c = 0; c < 32; ++c) e[c * FE_EXLD + lane] = cmul(v[bitrev5(c)], __ldg(tw2 + c * 32 + lane));
    __syncwarp();
#pragma unroll
    for (int b = 0; b < 32; ++b) v[b] = e[lane * FE_EXLD + b];
    __syncwarp();
    fft32(v);
#pragma unroll
    for (int d = 0; d < 32; ++d) e[lane + 32 * d] = v[bitrev5(d)];   // natural order X[k], k = lane + 32 d
    __syncwarp();
    // split the two real spectra and take magnitudes IN PLACE: slot k <- (|A_k|, |B_k|), k = 0..512.  Slot k is read and
    // written by one lane only and the mirror slots 1024-k > 512 are never written, so the loop needs no barrier.
    for (int k = lane; k < FE_BINS; k += 32) {
      const float2 a = e[k];
      const float2 b = e[(FE_NFFT - k) & (FE_NFFT - 1)];
      const float ar = 0.5f * (a.x + b.x), ai = 0.5f * (a.y - b.y);      // frame f0
      const float br = 0.5f * (a.y + b.y), bi = 0.5f * (b.x - a.x);      // frame f0+1
      e[k] = make_float2(sqrtf(ar * ar + ai * ai), sqrtf(br * br + bi * bi));
    }
    __syncwarp();
    if (STFT_OUT) {
      for (int i = lane; i < nout; i += 32) {
        const int fr = i >= FE_BINS ? 1 : 0, k = i - fr * FE_BINS;
        float val = 0.f;
        if (f0 + fr < n_frames) {
          const float2 mg = e[k];
          const float db = 20.f * log10f(fmaxf(1e-5f, fr ? mg.y : mg.x)) - 16.f;
          val = fminf(fmaxf((db + 100.f) / 100.f, 0.f), 1.f);
        }
        o[i] = val;
      }
      continue;
    }
    // mel projection of BOTH frames: the bin ranges of the 80 filters are cut into segments of <= 8 bins (tb->seg), a lane takes
    // every 32nd segment and sums it for the two frames at once (one 8-byte magnitude pair + one weight per bin), the partial
    // sums go to the dead upper half of the tile and each output adds its filter's partials in a fixed order.  The first version
    // gave every lane whole filters: 2 to 45 bins wide, so a warp always waited for its widest one (~130 dependent iterations per
    // frame pair, the hottest lines of the r01e profile); this is ~40.
    float2* part = e + 640;                                  // [<= 256] partial sums (frame f0, frame f0+1)
    const int nseg = tb->seg0[FE_MELS];
    for (int it = lane; it < nseg; it += 32) {
      const int4 sg = tb->seg[it];
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(tb->segw[it]));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(tb->segw[it]) + 1);
      const float wt[FE_SEG] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      // bins past the segment's end carry zero weights; the slots read there (<= 519) hold finite spectrum values
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int j = 0; j < FE_SEG; ++j) {
        const float2 mg = e[sg.y + j];
        a0 = fmaf(mg.x, wt[j], a0);
        a1 = fmaf(mg.y, wt[j], a1);
      }
      part[it] = make_float2(a0, a1);
    }
    __syncwarp();
    for (int i = lane; i < nout; i += 32) {
      const int fr = i >= FE_MELS ? 1 : 0, m = i - fr * FE_MELS;
      float val = 0.f;
      if (f0 + fr < n_frames) {
        float acc = 0.f;
        for (int it = tb->seg0[m]; it < tb->seg0[m + 1]; ++it) acc += fr ? part[it].y : part[it].x;
        const float db = 20.f * log10f(fmaxf(1e-5f, acc)) - 16.f;
        val = fminf(fmaxf((db + 100.f) / 100.f, 0.f), 1.f);
      }
      o[i] = val;
    }
  }
}

static size_t fe_tables_bytes() { return (sizeof(FeTables) + 255) / 256 * 256; }

}  // namespace avc

using namespace avc;

extern "C" size_t avc_logmel_workspace_bytes(int n_utt, int max_len) {
  if (n_utt <= 0 || max_len <= 0) return 0;
  const size_t fwd = ((size_t)n_utt * (max_len + 2 * FE_PADLEN) * sizeof(fe_y1_t) + 255) / 256 * 256;
  const size_t sig = ((size_t)n_utt * max_len * sizeof(float) + 255) / 256 * 256;
  return fe_tables_bytes() + sig + fwd + 4096;      // tail: the FE_TLEV tile powers (8 x 36 doubles)
}

static int fe_front(const float* wav, const float* dither, const int* lengths, int n_utt, int max_len, const float* mel_basis,
                    const double* filt, const double* zi, float* out, int max_frames, int out_bins, void* workspace,
                    size_t workspace_bytes, cudaStream_t st) {
  if (!workspace || workspace_bytes < avc_logmel_workspace_bytes(n_utt, max_len)) {
    set_error("avc_logmel_frontend: workspace too small");
    return AVC_ERR_WORKSPACE;
  }
  unsigned char* ws = (unsigned char*)workspace;
  FeTables* tb = (FeTables*)ws;
  float* sig = (float*)(ws + fe_tables_bytes());
  const size_t sig_b = ((size_t)n_utt * max_len * sizeof(float) + 255) / 256 * 256;
  const size_t fwd_b = ((size_t)n_utt * (max_len + 2 * FE_PADLEN) * sizeof(fe_y1_t) + 255) / 256 * 256;
  fe_y1_t* fwd = (fe_y1_t*)(ws + fe_tables_bytes() + sig_b);
  double* PW = (double*)(ws + fe_tables_bytes() + sig_b + fwd_b);
  fe_tables_kernel<<<ceil_div(FE_NFFT, 256), 256, 0, st>>>(mel_basis, tb);
  AVC_LAUNCHED();
  // one CTA per utterance, tiles staged through shared memory
  fe_tile_powers_kernel<<<1, 32, 0, st>>>(filt, PW);
  AVC_LAUNCHED();
  const size_t ism = ((size_t)FE_TNT * FE_NST + FE_TLEV * FE_NST * FE_NST + FE_NST) * sizeof(double) +
                     (size_t)FE_TNT * (FE_TCH + 1) * sizeof(float);
  AVC_CUDA(cudaFuncSetAttribute(fe_iir_sweep_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ism));
  AVC_CUDA(cudaFuncSetAttribute(fe_iir_sweep_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ism));
  fe_iir_sweep_kernel<false><<<n_utt, FE_TNT, ism, st>>>(wav, dither, lengths, max_len, filt, zi, PW, fwd, sig);
  AVC_LAUNCHED();
  fe_iir_sweep_kernel<true><<<n_utt, FE_TNT, ism, st>>>(wav, dither, lengths, max_len, filt, zi, PW, fwd, sig);
  AVC_LAUNCHED();
  const size_t smem = (size_t)(FE_THREADS / 32) * FE_EX * sizeof(float2);
  dim3 grid(ceil_div(max_frames, FE_FRAMES_PER_BLOCK), n_utt);
  if (out_bins == FE_MELS) {
    AVC_CUDA(cudaFuncSetAttribute(fe_stft_mel_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fe_stft_mel_kernel<false><<<grid, FE_THREADS, smem, st>>>(sig, lengths, max_len, mel_basis, tb, out, max_frames);
  } else {
    AVC_CUDA(cudaFuncSetAttribute(fe_stft_mel_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fe_stft_mel_kernel<true><<<grid, FE_THREADS, smem, st>>>(sig, lengths, max_len, mel_basis, tb, out, max_frames);
  }
  AVC_LAUNCHED();
  return AVC_OK;
}

extern "C" int avc_logmel_frontend(const float* wav, const float* dither, const int* lengths, int n_utt, int max_len,
                                   const float* mel_basis, const double* filt, const double* zi, float* out,
                                   int max_frames, void* workspace, size_t workspace_bytes, void* stream) {
  AVC_REQUIRE(wav && dither && lengths && mel_basis && filt && zi && out, "avc_logmel_frontend: null pointer");
  AVC_REQUIRE(n_utt > 0 && max_len > 0 && max_frames > 0, "avc_logmel_frontend: bad shape");
  AVC_REQUIRE(max_frames >= 1 + max_len / FE_HOP, "avc_logmel_frontend: max_frames %d < 1 + max_len/256 = %d", max_frames,
              1 + max_len / FE_HOP);
  return fe_front(wav, dither, lengths, n_utt, max_len, mel_basis, filt, zi, out, max_frames, FE_MELS, workspace, workspace_bytes,
                  as_stream(stream));
}

extern "C" int avc_logstft_frontend(const float* wav, const float* dither, const int* lengths, int n_utt, int max_len,
                                    const float* mel_basis, const double* filt, const double* zi, float* out,
                                    int max_frames, void* workspace, size_t workspace_bytes, void* stream) {
  AVC_REQUIRE(wav && dither && lengths && mel_basis && filt && zi && out, "avc_logstft_frontend: null pointer");
  AVC_REQUIRE(n_utt > 0 && max_len > 0 && max_frames > 0, "avc_logstft_frontend: bad shape");
  AVC_REQUIRE(max_frames >= 1 + max_len / FE_HOP, "avc_logstft_frontend: max_frames %d < 1 + max_len/256 = %d", max_frames,
              1 + max_len / FE_HOP);
  return fe_front(wav, dither, lengths, n_utt, max_len, mel_basis, filt, zi, out, max_frames, FE_BINS, workspace, workspace_bytes,
                  as_stream(stream));
}
