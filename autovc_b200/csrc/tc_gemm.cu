// Tensor-core (tcgen05 + TMEM + TMA) implementations of the GEMM-with-taps family, AVC_PREC_BF16.
//
// One persistent, warp-specialised kernel, two operand modes:
//   MODE_NT : C[m,n] = sum_tap sum_k A[row(m,tap),k] W[tap][n][k]   (conv fwd / dgrad, LSTM input projections, linear)
//             A tile  = TMA box (64 ch, 128 frames, 1 utterance) of the channels-last activation, K-major;
//             the time shift of a tap is a TMA coordinate offset, and frames outside [0,T) are zero-filled
//             by the TMA unit (= the conv's zero padding) -- no im2col, no halo copies.
//   MODE_TN : dW[tap][n,k] = sum_m dY[m,n] X[row(m,tap),k]          (weight gradients)
//             both operands are the same channels-last boxes (64 ch, 64 frames), consumed MN-major.
// Pipeline: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread tcgen05.mma issuer,
// warps 2..5 = epilogue (tcgen05.ld -> registers -> global), 6-stage smem ring (full/empty mbarriers),
// two TMEM accumulators (tmem_full/tmem_empty) so the epilogue of tile i overlaps the mainloop of i+1.
// bf16 operands, fp32 accumulation in TMEM; BatchNorm channel sums are reduced in the epilogue in
// fp32 per tile and accumulated in fp64.
#include "tc_common.cuh"

namespace avc {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 64;
constexpr int TC_STAGES = 6;
constexpr int TC_STAGE_A = TC_BM * TC_BK * 2, TC_STAGE_B = TC_BN * TC_BK * 2;   // 16 KB each
constexpr int TC_STAGE_BYTES = TC_STAGE_A + TC_STAGE_B;
constexpr int TC_THREADS = 192;
constexpr int TC_TMEM_COLS = 256;  // two 128-column fp32 accumulators
constexpr int TC_SMEM_BYTES = 1024 /*align slack*/ + TC_STAGES * TC_STAGE_BYTES + 4 * 2 * TC_BN * 4 /*stats*/ + 256 /*barriers*/;

enum { MODE_NT = 0, MODE_TN = 1 };

struct TcParams {
  // common
  int nB, T, ntaps, shift0;
  int N, K;            // logical (unpadded) sizes of the output's two dims (NT: C cols = N; TN: dW is N x K)
  // NT
  int t_tiles, n_tiles, kblocks;   // tiles along T (128 frames), along N (128 cols), 64-wide k blocks per tap
  const float* bias;
  float* C;
  int ldc, accumulate;
  double* stats;
  // TN
  int k_tiles, splits, rblocks, rblocks_per_split, tb64;   // tb64 = ceil(T/64)
  float* part;
};

// butterfly transpose-reduce: on return lane l holds sum over the warp's 32 lanes of v[l]
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool upper = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = upper ? v[i] : v[i + s];
      const float keep = upper ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

// ---------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                 // SWIZZLE_128B atoms need 1024-byte alignment
  uint8_t* gen = smem_raw + (base - raw);
  const uint32_t stage0 = base;
  float* stat_s = reinterpret_cast<float*>(gen + TC_STAGES * TC_STAGE_BYTES);           // [4 warps][2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(gen + TC_STAGES * TC_STAGE_BYTES + 4 * 2 * TC_BN * 4);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (TC_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * TC_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * TC_STAGES + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), TC_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  int total_tiles, kiters;
  if (MODE == MODE_NT) {
    total_tiles = p.nB * p.t_tiles * p.n_tiles;
    kiters = p.ntaps * p.kblocks;
  } else {
    total_tiles = p.ntaps * p.n_tiles * p.k_tiles * p.splits;
    kiters = 0;  // per tile
  }

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        if (MODE == MODE_NT) {
          const int n_tile = tile % p.n_tiles, m_tile = tile / p.n_tiles;
          const int b = m_tile / p.t_tiles, t0 = (m_tile % p.t_tiles) * TC_BM;
          for (int it = 0; it < kiters; ++it) {
            const int tap = it / p.kblocks, kb = it - tap * p.kblocks;
            mbar_wait(empty_bar(stage), phase ^ 1);
            mbar_expect_tx(full_bar(stage), TC_STAGE_BYTES);
            const uint32_t sa = stage0 + stage * TC_STAGE_BYTES;
            tma_load_3d(sa, &mapA, full_bar(stage), kb * TC_BK, t0 + p.shift0 + tap, b);
            tma_load_3d(sa + TC_STAGE_A, &mapB, full_bar(stage), kb * TC_BK, n_tile * TC_BN, tap);
            if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
          }
        } else {
          int r = tile;
          const int split = r % p.splits; r /= p.splits;
          const int k_tile = r % p.k_tiles; r /= p.k_tiles;
          const int n_tile = r % p.n_tiles; r /= p.n_tiles;
          const int tap = r;
          const int rb0 = split * p.rblocks_per_split;
          const int rb1 = min(p.rblocks, rb0 + p.rblocks_per_split);
          for (int rb = rb0; rb < rb1; ++rb) {
            const int b = rb / p.tb64, t0 = (rb % p.tb64) * 64;
            mbar_wait(empty_bar(stage), phase ^ 1);
            mbar_expect_tx(full_bar(stage), TC_STAGE_BYTES);
            const uint32_t sa = stage0 + stage * TC_STAGE_BYTES;
            tma_load_3d(sa, &mapA, full_bar(stage), n_tile * TC_BM, t0, b);
            tma_load_3d(sa + TC_STAGE_A / 2, &mapA, full_bar(stage), n_tile * TC_BM + 64, t0, b);
            tma_load_3d(sa + TC_STAGE_A, &mapB, full_bar(stage), k_tile * TC_BN, t0 + p.shift0 + tap, b);
            tma_load_3d(sa + TC_STAGE_A + TC_STAGE_B / 2, &mapB, full_bar(stage), k_tile * TC_BN + 64, t0 + p.shift0 + tap, b);
            if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(TC_BM, TC_BN, MODE == MODE_TN, MODE == MODE_TN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int iters = kiters;
        if (MODE == MODE_TN) {
          const int split = tile % p.splits;
          const int rb0 = split * p.rblocks_per_split;
          iters = max(0, min(p.rblocks, rb0 + p.rblocks_per_split) - rb0);
        }
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);      // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * TC_BN;
        for (int it = 0; it < iters; ++it) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = stage0 + stage * TC_STAGE_BYTES, sb = sa + TC_STAGE_A;
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            uint64_t da, db;
            if (MODE == MODE_NT) {
              // K-major: 128-byte rows, 8-row groups 1024 B apart; a K=16 slice is 32 B further along the row
              da = make_desc(sa + k * 32, 16, 1024);
              db = make_desc(sb + k * 32, 16, 1024);
            } else {
              // MN-major: each frame is a 128-byte row of 64 channels; 8-frame groups 1024 B apart (SBO),
              // the second 64-channel half of the tile 8192 B further (LBO); a K=16 slice = 16 frames = 2048 B
              da = make_desc(sa + k * 2048, TC_STAGE_A / 2, 1024);
              db = make_desc(sb + k * 2048, TC_STAGE_B / 2, 1024);
            }
            umma_f16(d_tmem, da, db, idesc, (it > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));                // frees the smem slot when these MMAs retire
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar(acc));                    // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps (TMEM -> registers -> global) =====================
    const int q = warp & 3;                  // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;           // row of the 128-row tile
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      bool has_work = true;
      if (MODE == MODE_TN) {
        const int split = tile % p.splits;
        has_work = split * p.rblocks_per_split < p.rblocks;
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * TC_BN;
      if (MODE == MODE_NT) {
        const int n_tile = tile % p.n_tiles, m_tile = tile / p.n_tiles;
        const int b = m_tile / p.t_tiles, t = (m_tile % p.t_tiles) * TC_BM + row;
        const bool row_ok = t < p.T;
        float* crow = p.C + ((size_t)b * p.T + t) * p.ldc;
#pragma unroll 1
        for (int c = 0; c < TC_BN / 32; ++c) {
          float v[32];
          tmem_ld32(t_addr + c * 32, v);
          if (c == TC_BN / 32 - 1) {          // all TMEM reads of this accumulator are done
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
          }
          const int n0 = n_tile * TC_BN + c * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = n0 + j;
            float x = v[j] + ((p.bias != nullptr && n < p.N) ? __ldg(p.bias + n) : 0.f);
            v[j] = (row_ok && n < p.N) ? x : 0.f;
          }
          if (row_ok) {
            if (n0 + 32 <= p.N && (p.ldc & 3) == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                float4* dst = reinterpret_cast<float4*>(crow + n0 + j);
                if (p.accumulate) {
                  const float4 old = *dst;
                  o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                }
                *dst = o;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (n0 + j < p.N) crow[n0 + j] = p.accumulate ? crow[n0 + j] + v[j] : v[j];
            }
          }
          if (p.stats != nullptr) {
            float sq[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) sq[j] = v[j] * v[j];
            const float s1 = warp_colsum32(v, lane);
            const float s2 = warp_colsum32(sq, lane);
            stat_s[(q * 2 + 0) * TC_BN + c * 32 + lane] = s1;
            stat_s[(q * 2 + 1) * TC_BN + c * 32 + lane] = s2;
          }
        }
        if (p.stats != nullptr) {
          asm volatile("bar.sync 1, 128;" ::: "memory");      // the four epilogue warps
          const int col = threadIdx.x - 64;                   // 0..127
          const int n = n_tile * TC_BN + col;
          if (n < p.N) {
            double a = 0.0, bq = 0.0;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              a += (double)stat_s[(w * 2 + 0) * TC_BN + col];
              bq += (double)stat_s[(w * 2 + 1) * TC_BN + col];
            }
            atomicAdd(p.stats + n, a);
            atomicAdd(p.stats + p.N + n, bq);
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
      } else {
        int r = tile;
        const int split = r % p.splits; r /= p.splits;
        const int k_tile = r % p.k_tiles; r /= p.k_tiles;
        const int n_tile = r % p.n_tiles; r /= p.n_tiles;
        const int tap = r;
        const int n = n_tile * TC_BM + row;
        float* orow = p.part + (((size_t)split * p.ntaps + tap) * p.N + n) * p.K;
#pragma unroll 1
        for (int c = 0; c < TC_BN / 32; ++c) {
          float v[32];
          tmem_ld32(t_addr + c * 32, v);
          if (c == TC_BN / 32 - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
          }
          const int k0 = k_tile * TC_BN + c * 32;
          if (n < p.N) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (k0 + j < p.K) orow[k0 + j] = has_work ? v[j] : 0.f;
          }
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TC_TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------
// fp32 -> padded bf16 staging
// ---------------------------------------------------------------------------------------------------
// dst[r][c] (ld = Cp, bf16) = src[r][c] (ld = lds, fp32) for c < C, zero for C <= c < Cp; rows >= R_src are zero
__global__ void cvt_pad_bf16_kernel(const float* __restrict__ src, int lds, __nv_bfloat16* __restrict__ dst, int Cp,
                                    size_t R_dst, size_t R_src, int C) {
  const size_t total = R_dst * (size_t)(Cp / 2);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / (Cp / 2);
    const int c = (int)(i % (Cp / 2)) * 2;
    float a = 0.f, b = 0.f;
    if (r < R_src) {
      if (c < C) a = src[r * lds + c];
      if (c + 1 < C) b = src[r * lds + c + 1];
    }
    reinterpret_cast<__nv_bfloat162*>(dst)[i] = __floats2bfloat162_rn(a, b);
  }
}
// weights [ntaps][N][K] fp32 -> [ntaps][Np][Kp] bf16, zero padded
__global__ void cvt_pad_w_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int ntaps, int N, int K,
                                      int Np, int Kp) {
  const size_t total = (size_t)ntaps * Np * Kp;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % Kp);
    const int n = (int)((i / Kp) % Np);
    const int tap = (int)(i / ((size_t)Kp * Np));
    const float v = (n < N && k < K) ? src[((size_t)tap * N + n) * K + k] : 0.f;
    dst[i] = __float2bfloat16_rn(v);
  }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------

static int cvt_blocks(size_t total) { return (int)std::min<size_t>(ceil_div(total, (size_t)256), (size_t)num_sms() * 16); }

static int ensure_smem_attr() {
  static bool done = false;
  if (!done) {
    AVC_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<MODE_NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    AVC_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<MODE_TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    done = true;
  }
  return AVC_OK;
}

size_t gemm_nt_workspace_tc(int nB, int T, int N, int K, int ntaps) {
  const int Kp = round_up(K, TC_BK), Np = round_up(N, TC_BN);
  return align256((size_t)nB * T * Kp * 2) + align256((size_t)ntaps * Np * Kp * 2);
}

int gemm_nt_taps_tc(const float* A, int lda, const float* W, const float* bias, float* C, int ldc, int nB, int T, int N,
                    int K, int ntaps, int shift0, double* stats, int accumulate, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (stats && accumulate) {
    set_error("avc_gemm_nt_taps: chan_stats and accumulate are mutually exclusive");
    return AVC_ERR_UNSUPPORTED;
  }
  if (!ws || ws_bytes < gemm_nt_workspace_tc(nB, T, N, K, ntaps)) {
    set_error("avc_gemm_nt_taps(bf16): workspace %zu < %zu", ws_bytes, gemm_nt_workspace_tc(nB, T, N, K, ntaps));
    return AVC_ERR_WORKSPACE;
  }
  int rc = ensure_smem_attr();
  if (rc) return rc;
  const int Kp = round_up(K, TC_BK), Np = round_up(N, TC_BN);
  const size_t M = (size_t)nB * T;
  __nv_bfloat16* Ab = (__nv_bfloat16*)ws;
  __nv_bfloat16* Wb = (__nv_bfloat16*)((uint8_t*)ws + align256(M * Kp * 2));
  cvt_pad_bf16_kernel<<<cvt_blocks(M * (Kp / 2)), 256, 0, st>>>(A, lda, Ab, Kp, M, M, K);
  AVC_LAUNCHED();
  cvt_pad_w_bf16_kernel<<<cvt_blocks((size_t)ntaps * Np * Kp), 256, 0, st>>>(W, Wb, ntaps, N, K, Np, Kp);
  AVC_LAUNCHED();
  CUtensorMap mA, mB;
  rc = make_map3(&mA, Ab, Kp, T, nB, Kp, (uint64_t)T * Kp, TC_BK, TC_BM);
  if (rc) return rc;
  rc = make_map3(&mB, Wb, Kp, Np, ntaps, Kp, (uint64_t)Np * Kp, TC_BK, TC_BN);
  if (rc) return rc;
  TcParams p{};
  p.nB = nB; p.T = T; p.ntaps = ntaps; p.shift0 = shift0; p.N = N; p.K = K;
  p.t_tiles = ceil_div(T, TC_BM); p.n_tiles = Np / TC_BN; p.kblocks = Kp / TC_BK;
  p.bias = bias; p.C = C; p.ldc = ldc; p.accumulate = accumulate; p.stats = stats;
  const int tiles = nB * p.t_tiles * p.n_tiles;
  const int grid = std::min(tiles, num_sms());
  tc_gemm_kernel<MODE_NT><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(mA, mB, p);
  AVC_LAUNCHED();
  return AVC_OK;
}

static int tn_splits_tc(int rblocks, int tiles) {
  const int sms = num_sms();
  int best = 1;
  double best_eff = 0.0;
  for (int s = 1; s <= 16; ++s) {
    if (s > 1 && rblocks / s < 8) break;
    const int items = tiles * s;
    const double eff = (double)items / ((double)ceil_div(items, sms) * sms);
    if (eff > best_eff + 0.03) {
      best_eff = eff;
      best = s;
    }
  }
  return best;
}

struct TnPlan {
  int Np, Kp, rblocks, tiles, splits, rps;
  size_t off_y, off_x, off_part, total;
};
static TnPlan tn_plan(int nB, int T, int N, int K, int ntaps) {
  TnPlan pl;
  pl.Np = round_up(N, TC_BM);
  pl.Kp = round_up(K, TC_BN);
  pl.rblocks = nB * ceil_div(T, 64);
  pl.tiles = ntaps * (pl.Np / TC_BM) * (pl.Kp / TC_BN);
  pl.splits = tn_splits_tc(pl.rblocks, pl.tiles);
  pl.rps = ceil_div(pl.rblocks, pl.splits);
  const size_t M = (size_t)nB * T;
  pl.off_y = 0;
  pl.off_x = align256(M * pl.Np * 2);
  pl.off_part = pl.off_x + align256(M * pl.Kp * 2);
  pl.total = pl.off_part + align256((size_t)pl.splits * ntaps * N * K * 4);
  return pl;
}
size_t gemm_tn_workspace_tc(int nB, int T, int N, int K, int ntaps) { return tn_plan(nB, T, N, K, ntaps).total; }

int launch_wgrad_reduce(const float* part, float* dW, int N, int K, int ntaps, int splits, int out_mode, int accumulate,
                        cudaStream_t st);

int gemm_tn_taps_tc(const float* dY, int ldy, const float* X, int ldx, float* dW, int nB, int T, int N, int K, int ntaps,
                    int shift0, int out_mode, int accumulate, void* ws, size_t ws_bytes, cudaStream_t st) {
  const TnPlan pl = tn_plan(nB, T, N, K, ntaps);
  if (!ws || ws_bytes < pl.total) {
    set_error("avc_gemm_tn_taps(bf16): workspace %zu < %zu", ws_bytes, pl.total);
    return AVC_ERR_WORKSPACE;
  }
  int rc = ensure_smem_attr();
  if (rc) return rc;
  const size_t M = (size_t)nB * T;
  __nv_bfloat16* Yb = (__nv_bfloat16*)((uint8_t*)ws + pl.off_y);
  __nv_bfloat16* Xb = (__nv_bfloat16*)((uint8_t*)ws + pl.off_x);
  float* part = (float*)((uint8_t*)ws + pl.off_part);
  cvt_pad_bf16_kernel<<<cvt_blocks(M * (pl.Np / 2)), 256, 0, st>>>(dY, ldy, Yb, pl.Np, M, M, N);
  AVC_LAUNCHED();
  cvt_pad_bf16_kernel<<<cvt_blocks(M * (pl.Kp / 2)), 256, 0, st>>>(X, ldx, Xb, pl.Kp, M, M, K);
  AVC_LAUNCHED();
  CUtensorMap mA, mB;
  rc = make_map3(&mA, Yb, pl.Np, T, nB, pl.Np, (uint64_t)T * pl.Np, 64, 64);
  if (rc) return rc;
  rc = make_map3(&mB, Xb, pl.Kp, T, nB, pl.Kp, (uint64_t)T * pl.Kp, 64, 64);
  if (rc) return rc;
  TcParams p{};
  p.nB = nB; p.T = T; p.ntaps = ntaps; p.shift0 = shift0; p.N = N; p.K = K;
  p.n_tiles = pl.Np / TC_BM; p.k_tiles = pl.Kp / TC_BN; p.splits = pl.splits; p.rblocks = pl.rblocks;
  p.rblocks_per_split = pl.rps; p.tb64 = ceil_div(T, 64); p.part = part;
  const int items = pl.tiles * pl.splits;
  const int grid = std::min(items, num_sms());
  tc_gemm_kernel<MODE_TN><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(mA, mB, p);
  AVC_LAUNCHED();
  return launch_wgrad_reduce(part, dW, N, K, ntaps, pl.splits, out_mode, accumulate, st);
}

}  // namespace avc
