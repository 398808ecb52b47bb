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
#include <stdlib.h>

#include <cuda_fp16.h>

#include "tc_common.cuh"

namespace avc {

constexpr int TC_BM = 128;
constexpr int TC_STAGE_A = TC_BM * 128;   // 128 rows x one 128-byte swizzle row = 16 KB
// BN = tile width (output columns per CTA tile): 128 or 256.  Shared-memory bandwidth (operand reads by the MMA
// plus TMA fills) is what bounds a 128x128 tile (r01 ncu: tensor pipe 48%); 128x256 moves 25% fewer bytes per MAC.
template <int BN> struct TcCfg {
  static constexpr int STAGE_B = BN * 128;
  static constexpr int STAGE_BYTES = TC_STAGE_A + STAGE_B;
  static constexpr int STAGES = BN == 128 ? 6 : 4;
  static constexpr int TMEM_COLS = 2 * BN;          // two fp32 accumulators
  static constexpr int STAT_BYTES = 4 * 2 * BN * 4;
  static constexpr int SMEM_BYTES = 1024 /*align slack*/ + STAGES * STAGE_BYTES + STAT_BYTES + 256 /*barriers*/;
};
// EB = operand element bytes: 2 = bf16 (kind::f16), 4 = fp32 read as tf32 (kind::tf32).  One 128-byte row holds
// 128/EB elements of K (NT) or of channels (TN); one MMA consumes 32 bytes of K = 32/EB elements.
template <int EB> struct TcGeom {
  static constexpr int ROW = 128 / EB;         // elements per swizzle row
  static constexpr int KMMA = 32 / EB;         // K elements per tcgen05.mma
  static constexpr int RS = 4 * KMMA;          // TN: frames per pipeline stage (4 MMAs)
  static constexpr int NBOX = 128 / ROW;       // TN: 128-channel tile = NBOX TMA boxes of ROW channels
  static constexpr int BOX_BYTES = RS * 128;   // TN: bytes of one box (RS frames x 128 B)
};
constexpr int TC_THREADS = 192;

enum { MODE_NT = 0, MODE_TN = 1 };

struct TcParams {
  // common
  int nB, T, ntaps, shift0;
  int N, K;            // logical (unpadded) sizes of the output's two dims (NT: C cols = N; TN: dW is N x K)
  // NT
  int t_tiles, n_tiles, kblocks;   // tiles along T (128 frames), along N (128 cols), one-swizzle-row k blocks per tap
  int ksplit;                      // NT, chunked kernel only: the reduction of a tile is cut into ksplit work items; item s writes its
  size_t c_split_stride;           // partial result to C + s * c_split_stride (the caller sums them; bias goes with item 0)
  const float* bias;
  float* C;
  int ldc, accumulate;
  double* stats;
  // TN
  int k_tiles, splits, rblocks, rblocks_per_split, tbr;   // tbr = ceil(T / frames-per-stage)
  int grouped_a, grouped_b;   // operand fetched with one grouped 4-D box per stage (make_map4_grouped)
  uint32_t idesc;             // tcgen05 instruction descriptor (operand formats are chosen at run time)
  unsigned long long* trace;  // avc_debug_set_trace: %globaltimer stamps of pair 0's leader, 8 slots per tile (first 64 tiles)
  int tma_store;              // NT pair kernel: C leaves through shared memory + TMA stores (needs ldc % 4 == 0, 16-byte aligned C)
  float* part;
};

// TN work-item order: (k_tile, n_tile, tap) vary fastest and the row split slowest, so the CTAs that run
// concurrently (consecutive item indices) sweep the SAME row range of dY / X with different output tiles and
// re-use it from L2 instead of re-reading HBM (r01 ncu: 790 MB DRAM reads for 134 MB of operands before this).
struct TnItem { int tap, n_tile, k_tile, split; };
__device__ __forceinline__ TnItem tn_decode(int tile, int k_tiles, int n_tiles, int ntaps) {
  TnItem it;
  it.k_tile = tile % k_tiles; tile /= k_tiles;
  it.n_tile = tile % n_tiles; tile /= n_tiles;
  it.tap = tile % ntaps; tile /= ntaps;
  it.split = tile;
  return it;
}

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
// CH > 0 ("chunked accumulation", the fp32x3 mode): the reduction of a tile is cut into chunks of CH pipeline stages; every
// chunk is accumulated in TMEM from zero (the two accumulators alternate per CHUNK instead of per tile) and the epilogue
// warps add the finished chunks into fp32 registers with round-to-nearest while the next chunk's MMAs run.  tcgen05.mma
// adds into its accumulator with truncation: over the ~1000 dependent MMAs of a 3xTF32 conv tile that is a systematic
// 5e-5 relative shrink (measured: 4e-4..7e-4 max-abs on x_identic_psnt, four to seven times the 1e-4 gate); with 32-MMA
// chains summed in registers it is below the fp32 noise of the reference itself.  Needs BN == 128 (128 accumulator registers
// per epilogue thread).
// A chunked 256-wide tile runs EIGHT epilogue warps (320 threads): two per TMEM lane quadrant, each owning 128 of the 256 columns,
// so that a thread still keeps 128 accumulator registers.  (r02 ncu: the chunked 128 x 128 tf32 tile pulls 32 KB of operands per
// MFLOP and sits on the L2 -> SM delivery rate, 13.8 TB/s; 128 x 256 needs 23 KB per MFLOP.)
template <int BN, int CH>
struct TcThreads { static constexpr int value = (CH > 0 && BN == 256) ? 320 : TC_THREADS; };
template <int MODE, int EB, int BN, int CH = 0>
__global__ void __launch_bounds__((TcThreads<BN, CH>::value), 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const TcParams p) {
  constexpr int EPI_WARPS = (CH > 0 && BN == 256) ? 8 : 4;
  using Cf = TcCfg<BN>;
  constexpr int TC_STAGES = Cf::STAGES, TC_STAGE_BYTES = Cf::STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                 // SWIZZLE_128B atoms need 1024-byte alignment
  uint8_t* gen = smem_raw + (base - raw);
  const uint32_t stage0 = base;
  float* stat_s = reinterpret_cast<float*>(gen + TC_STAGES * TC_STAGE_BYTES);           // [4 warps][2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(gen + TC_STAGES * TC_STAGE_BYTES + Cf::STAT_BYTES);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (TC_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * TC_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * TC_STAGES + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  using Gm = TcGeom<EB>;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), Cf::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  int total_tiles, kiters;
  if (MODE == MODE_NT) {
    total_tiles = p.nB * p.t_tiles * p.n_tiles * (p.ksplit > 1 ? p.ksplit : 1);
    kiters = p.ntaps * p.kblocks;
  } else {
    total_tiles = p.ntaps * p.n_tiles * p.k_tiles * p.splits;
    kiters = 0;  // per tile
  }

  // NT with split reduction: work item -> (tile, first and one-past-last pipeline stage)
  const int ksplit = (MODE == MODE_NT && p.ksplit > 1) ? p.ksplit : 1;
  const int kper = (kiters + ksplit - 1) / ksplit;
  if (warp == 0) {
    // ===================== TMA producer (whole warp loops, one elected lane issues) =====================
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        if (MODE == MODE_NT) {
          const int ks = tile % ksplit, tl = tile / ksplit;
          const int n_tile = tl % p.n_tiles, m_tile = tl / p.n_tiles;
          const int b = m_tile / p.t_tiles, t0 = (m_tile % p.t_tiles) * TC_BM;
          for (int it = ks * kper; it < min(kiters, (ks + 1) * kper); ++it) {
            const int tap = it / p.kblocks, kb = it - tap * p.kblocks;
            mbar_wait(empty_bar(stage), phase ^ 1);
            if (elect_one()) {
              mbar_expect_tx(full_bar(stage), TC_STAGE_BYTES);
              const uint32_t sa = stage0 + stage * TC_STAGE_BYTES;
              tma_load_3d(sa, &mapA, full_bar(stage), kb * Gm::ROW, t0 + p.shift0 + tap, b);
              tma_load_3d(sa + TC_STAGE_A, &mapB, full_bar(stage), kb * Gm::ROW, n_tile * BN, tap);
            }
            __syncwarp();
            if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
          }
        } else {
          const TnItem wi = tn_decode(tile, p.k_tiles, p.n_tiles, p.ntaps);
          const int split = wi.split, k_tile = wi.k_tile, n_tile = wi.n_tile, tap = wi.tap;
          const int rb0 = split * p.rblocks_per_split;
          const int rb1 = min(p.rblocks, rb0 + p.rblocks_per_split);
          for (int rb = rb0; rb < rb1; ++rb) {
            const int b = rb / p.tbr, t0 = (rb % p.tbr) * Gm::RS;
            mbar_wait(empty_bar(stage), phase ^ 1);
            if (elect_one()) {
              mbar_expect_tx(full_bar(stage), TC_STAGE_BYTES);
              const uint32_t sa = stage0 + stage * TC_STAGE_BYTES;
              // one TMA instruction per operand when the channel count is a multiple of the 128-byte row (a single
              // thread issues ~1 TMA per 100 cycles: 12 small boxes per stage made the producer the bottleneck)
              if (p.grouped_a) {
                tma_load_4d(sa, &mapA, full_bar(stage), 0, t0, n_tile * Gm::NBOX, b);
              } else {
#pragma unroll
                for (int h = 0; h < Gm::NBOX; ++h)
                  tma_load_3d(sa + h * Gm::BOX_BYTES, &mapA, full_bar(stage), n_tile * TC_BM + h * Gm::ROW, t0, b);
              }
              if (p.grouped_b) {
                tma_load_4d(sa + TC_STAGE_A, &mapB, full_bar(stage), 0, t0 + p.shift0 + tap, k_tile * (BN / Gm::ROW), b);
              } else {
#pragma unroll
                for (int h = 0; h < BN / Gm::ROW; ++h)
                  tma_load_3d(sa + TC_STAGE_A + h * Gm::BOX_BYTES, &mapB, full_bar(stage), k_tile * BN + h * Gm::ROW,
                              t0 + p.shift0 + tap, b);
              }
            }
            __syncwarp();
            if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp loops, one elected lane issues) =====================
    {
      const uint32_t idesc = p.idesc;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int iters = kiters;
        if (MODE == MODE_TN) {
          const int split = tn_decode(tile, p.k_tiles, p.n_tiles, p.ntaps).split;
          const int rb0 = split * p.rblocks_per_split;
          iters = max(0, min(p.rblocks, rb0 + p.rblocks_per_split) - rb0);
        } else if (ksplit > 1) {
          const int ks = tile % ksplit;
          iters = max(0, min(kiters, (ks + 1) * kper) - ks * kper);
        }
        if (CH == 0) {
          mbar_wait(tempty_bar(acc), acc_phase ^ 1);      // epilogue has drained this accumulator
          tc_fence_after();
        }
        uint32_t d_tmem = tmem_base + acc * BN;
        for (int it = 0; it < iters; ++it) {
          const bool chunk_start = CH > 0 && (it % (CH > 0 ? CH : 1)) == 0;
          const bool chunk_end = CH > 0 && ((it % (CH > 0 ? CH : 1)) == (CH > 0 ? CH : 1) - 1 || it == iters - 1);
          if (chunk_start) {
            mbar_wait(tempty_bar(acc), acc_phase ^ 1);    // epilogue has added this accumulator's previous chunk into its registers
            tc_fence_after();
            d_tmem = tmem_base + acc * BN;
          }
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = stage0 + stage * TC_STAGE_BYTES, sb = sa + TC_STAGE_A;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              uint64_t da, db;
              if (MODE == MODE_NT) {
                // K-major: 128-byte rows, 8-row groups 1024 B apart; one MMA's K slice is 32 B further along the row
                da = make_desc(sa + k * 32, 16, 1024);
                db = make_desc(sb + k * 32, 16, 1024);
              } else {
                // MN-major: each frame is a 128-byte row of ROW channels; 8-frame groups 1024 B apart (SBO), the
                // next ROW-channel box of the tile BOX_BYTES further (LBO); one MMA's K slice = KMMA frames
                // (32-bit operands: the only MN-major layout is SWIZZLE_128B_BASE32B -- 32-byte swizzle atoms, 4-frame
                //  groups 512 B apart -- which the TMA map produces with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)
                da = make_desc(sa + k * Gm::KMMA * 128, Gm::BOX_BYTES, EB == 2 ? 1024 : 512, EB == 2 ? 2 : 1);
                db = make_desc(sb + k * Gm::KMMA * 128, Gm::BOX_BYTES, EB == 2 ? 1024 : 512, EB == 2 ? 2 : 1);
              }
              const bool first = CH > 0 ? (chunk_start && k == 0) : (it == 0 && k == 0);
              umma<EB>(d_tmem, da, db, idesc, first ? 0u : 1u);
            }
            umma_commit(empty_bar(stage));                // frees the smem slot when these MMAs retire
            if (CH > 0 ? chunk_end : (it == iters - 1)) umma_commit(tfull_bar(acc));   // accumulator (chunk) complete -> epilogue
          }
          __syncwarp();
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
          if (chunk_end) {
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
          }
        }
        if (iters == 0) {                                 // an empty split still hands its (untouched) accumulator on
          if (CH > 0) {
            mbar_wait(tempty_bar(acc), acc_phase ^ 1);
            tc_fence_after();
          }
          if (elect_one()) umma_commit(tfull_bar(acc));
          __syncwarp();
          if (CH > 0) {
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
          }
        }
        if (CH == 0) {
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
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
        const int split = tn_decode(tile, p.k_tiles, p.n_tiles, p.ntaps).split;
        has_work = split * p.rblocks_per_split < p.rblocks;
      }
      if constexpr (CH > 0) {
        // ---- chunked accumulation: add every finished chunk into this thread's 128 fp32 registers (its row, its column half) ----
        constexpr int EC = 128;                              // columns per epilogue thread
        const int ch0 = (BN == 256) ? ((warp - 2) >> 2) * EC : 0;      // first column of this warp's half of the tile
        int iters = kiters;
        const int ks = tile % ksplit, tl = tile / ksplit;
        if (MODE == MODE_TN) {
          const int split = tn_decode(tile, p.k_tiles, p.n_tiles, p.ntaps).split;
          const int rb0 = split * p.rblocks_per_split;
          iters = max(0, min(p.rblocks, rb0 + p.rblocks_per_split) - rb0);
        } else if (ksplit > 1) {
          iters = max(0, min(kiters, (ks + 1) * kper) - ks * kper);
        }
        float r[EC];
#pragma unroll
        for (int j = 0; j < EC; ++j) r[j] = 0.f;
        const int nck = iters > 0 ? (iters + CH - 1) / CH : 1;
#pragma unroll 1
        for (int ck = 0; ck < nck; ++ck) {
          mbar_wait(tfull_bar(acc), acc_phase);
          tc_fence_after();
          const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + ch0;
          if (iters > 0) {
#pragma unroll
            for (int c = 0; c < EC / 32; ++c) {
              float v[32];
              tmem_ld32(t_addr + c * 32, v);
#pragma unroll
              for (int j = 0; j < 32; ++j) r[c * 32 + j] += v[j];
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(acc));
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (MODE == MODE_NT) {
          const int n_tile = tl % p.n_tiles, m_tile = tl / p.n_tiles;
          const int b = m_tile / p.t_tiles, t = (m_tile % p.t_tiles) * TC_BM + row;
          const bool row_ok = t < p.T;
          float* crow = p.C + (size_t)ks * p.c_split_stride + ((size_t)b * p.T + t) * p.ldc;
          const bool add_bias = p.bias != nullptr && ks == 0;
#pragma unroll
          for (int c = 0; c < EC / 32; ++c) {
            const int n0 = n_tile * BN + ch0 + c * 32;
            if (n0 < p.N) {                      // warp-uniform
              float v[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int n = n0 + j;
                const float x = r[c * 32 + j] + ((add_bias && n < p.N) ? __ldg(p.bias + n) : 0.f);
                v[j] = (row_ok && n < p.N) ? x : 0.f;
              }
              if (row_ok) {
                if (n0 + 32 <= p.N && (p.ldc & 3) == 0 && (((uintptr_t)p.C & 15) == 0)) {
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
                stat_s[(q * 2 + 0) * BN + ch0 + c * 32 + lane] = s1;
                stat_s[(q * 2 + 1) * BN + ch0 + c * 32 + lane] = s2;
              }
            } else if (p.stats != nullptr) {
              stat_s[(q * 2 + 0) * BN + ch0 + c * 32 + lane] = 0.f;
              stat_s[(q * 2 + 1) * BN + ch0 + c * 32 + lane] = 0.f;
            }
          }
          if (p.stats != nullptr) {
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");      // the epilogue warps
            const int col = threadIdx.x - 64;                                     // EPI_WARPS * 32 threads sweep the BN columns once
            if (col < BN) {
              const int n = n_tile * BN + col;
              if (n < p.N) {
                double a = 0.0, bq = 0.0;
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                  a += (double)stat_s[(w * 2 + 0) * BN + col];
                  bq += (double)stat_s[(w * 2 + 1) * BN + col];
                }
                atomicAdd(p.stats + n, a);
                atomicAdd(p.stats + p.N + n, bq);
              }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
          }
        } else {
          const TnItem wi = tn_decode(tile, p.k_tiles, p.n_tiles, p.ntaps);
          const int n = wi.n_tile * TC_BM + row;
          float* orow = p.part + (((size_t)wi.split * p.ntaps + wi.tap) * p.N + n) * p.K;
          if (n < p.N) {
#pragma unroll
            for (int j = 0; j < EC; ++j) {
              const int k = wi.k_tile * BN + ch0 + j;
              if (k < p.K) orow[k] = r[j];
            }
          }
        }
        continue;
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      if (MODE == MODE_NT) {
        const int n_tile = tile % p.n_tiles, m_tile = tile / p.n_tiles;
        const int b = m_tile / p.t_tiles, t = (m_tile % p.t_tiles) * TC_BM + row;
        const bool row_ok = t < p.T;
        const bool rows_partial = (m_tile % p.t_tiles) * TC_BM + TC_BM > p.T;   // tile-uniform
        float* crow = p.C + ((size_t)b * p.T + t) * p.ldc;
        // chunks that lie entirely beyond column N are not even read (N = 64 or 80 in a 128-wide tile), and full chunks
        // skip the per-element masks: a lone epilogue warp per scheduler is issue-bound (see tc_gemm2_nt_kernel)
        const int nchunks = min(BN / 32, (p.N - n_tile * BN + 31) >> 5);
#pragma unroll 1
        for (int c = 0; c < nchunks; ++c) {
          float v[32];
          tmem_ld32(t_addr + c * 32, v);
          if (c == nchunks - 1) {          // all TMEM reads of this accumulator are done
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
          }
          const int n0 = n_tile * BN + c * 32;
          if (n0 + 32 <= p.N) {
            if (p.bias != nullptr) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] += __ldg(p.bias + n0 + j);
            }
            if (rows_partial && p.stats != nullptr && !row_ok) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = 0.f;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int n = n0 + j;
              float x = v[j] + ((p.bias != nullptr && n < p.N) ? __ldg(p.bias + n) : 0.f);
              v[j] = (row_ok && n < p.N) ? x : 0.f;
            }
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
            stat_s[(q * 2 + 0) * BN + c * 32 + lane] = s1;
            stat_s[(q * 2 + 1) * BN + c * 32 + lane] = s2;
          }
        }
        if (p.stats != nullptr) {
          asm volatile("bar.sync 1, 128;" ::: "memory");      // the four epilogue warps
#pragma unroll
          for (int cp = 0; cp < BN / 128; ++cp) {
            const int col = threadIdx.x - 64 + cp * 128;       // 128 epilogue threads sweep the BN columns
            const int n = n_tile * BN + col;
            if (n < p.N) {
              double a = 0.0, bq = 0.0;
#pragma unroll
              for (int w = 0; w < 4; ++w) {
                a += (double)stat_s[(w * 2 + 0) * BN + col];
                bq += (double)stat_s[(w * 2 + 1) * BN + col];
              }
              atomicAdd(p.stats + n, a);
              atomicAdd(p.stats + p.N + n, bq);
            }
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
      } else {
        const TnItem wi = tn_decode(tile, p.k_tiles, p.n_tiles, p.ntaps);
        const int split = wi.split, k_tile = wi.k_tile, n_tile = wi.n_tile, tap = wi.tap;
        const int n = n_tile * TC_BM + row;
        float* orow = p.part + (((size_t)split * p.ntaps + tap) * p.N + n) * p.K;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          float v[32];
          tmem_ld32(t_addr + c * 32, v);
          if (c == BN / 32 - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
          }
          const int k0 = k_tile * BN + c * 32;
          if (n < p.N && k0 < p.K) {
            if (!has_work) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = 0.f;
            }
            if ((p.K & 3) == 0 && k0 + 32 <= p.K) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(orow + k0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (k0 + j < p.K) orow[k0 + j] = v[j];
            }
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
    tmem_dealloc(tmem_base, Cf::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------
// MODE_NT on CTA pairs (tcgen05 cta_group::2): a cluster of two CTAs on neighbouring SMs owns a 256-row x 256-col
// output tile.  Each CTA stages ITS 128 rows of A and ITS 128-row half of the W tile (32 KB per k block instead
// of the 48 KB a lone CTA needs for 128x256), the leader's single thread issues M=256 MMAs that read both SMs'
// shared memory and write both SMs' TMEM, and each CTA drains its own 128 x 256 accumulator half.  Per-SM operand
// ingest -- what bounded the 1-CTA kernel (r01 ncu: tensor pipe 63%) -- drops by a third per MAC.
// Barrier protocol (P = peer, L = leader; "mc" = tcgen05.commit multicast to both CTAs):
//   full[s]  (L only, 1 arrival + 64 KB tx)  <- L's expect_tx, L's and P's TMA bytes
//   empty[s] (each CTA, 1 arrival)           <- L's MMA commit mc   -> each CTA's producer refills its own slot
//   tfull[a] (each CTA, 1 arrival)           <- L's MMA commit mc   -> each CTA's epilogue
//   tempty[a](L only, 8 arrivals)            <- 4 epilogue warps of L (local) and of P (remote arrive)
// ---------------------------------------------------------------------------------------------------
// trace slots per tile of pair 0 / leader: 0 producer slot free, 1 loads issued, 2 accumulator free (MMA), 3 operands landed,
// 4 MMAs issued, 5 epilogue woke, 6 first chunk done, 7 accumulator released
#define TC2_TRACE(slot)                                                                                          \
  do {                                                                                                           \
    if (p.trace != nullptr && blockIdx.x == 0 && iter < 64) {                                                    \
      unsigned long long t_;                                                                                     \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                                     \
      p.trace[iter * 8 + (slot)] = t_;                                                                           \
    }                                                                                                            \
  } while (0)

// chunk-level stamps of tile #4 (steady state): p.trace[512 + 4 * chunk + k]; k: 0 TMEM read, 1 bias/mask, 2 staged, 3 store issued
#define TC2_CTRACE(k)                                                                                            \
  do {                                                                                                           \
    if (p.trace != nullptr && blockIdx.x == 0 && iter == 4 && threadIdx.x == 64) {                               \
      unsigned long long t_;                                                                                     \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                                     \
      p.trace[512 + c * 4 + (k)] = t_;                                                                           \
    }                                                                                                            \
  } while (0)

struct Tc2Cfg {
  static constexpr int BN = 256;                      // pair tile width
  static constexpr int STAGE_B = 128 * 128;           // this CTA's half of the W tile: 128 rows x 128 B
  static constexpr int STAGE_BYTES = TC_STAGE_A + STAGE_B;
  static constexpr int STAGES = 5;                    // 6 stages with one staging tile per warp measured no faster
  static constexpr int TMEM_COLS = 512;               // two 256-column fp32 accumulators
  static constexpr int STAT_BYTES = 4 * 2 * BN * 4;
  static constexpr int EPI_BUFS = 2;                  // staging tiles per epilogue warp
  static constexpr int EPI_BYTES = 4 * EPI_BUFS * 4096;   // per epilogue warp: 32-row x 32-column fp32 staging tiles
  static constexpr int BIAS_BYTES = 4 * BN * 4;       // per epilogue warp: the tile's 256 bias values
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + STAT_BYTES + EPI_BYTES + BIAS_BYTES + 256;
};

template <int EB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
tc_gemm2_nt_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                   const __grid_constant__ CUtensorMap mapC, const TcParams p) {
  using Cf = Tc2Cfg;
  using Gm = TcGeom<EB>;
  constexpr int BN = Cf::BN, ST = Cf::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const uint32_t stage0 = base;
  float* stat_s = reinterpret_cast<float*>(gen + ST * Cf::STAGE_BYTES);
  const uint32_t epi0 = base + ST * Cf::STAGE_BYTES + Cf::STAT_BYTES;                   // 1024-byte aligned (128B swizzle atoms)
  const uint32_t bias0 = epi0 + Cf::EPI_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(gen + ST * Cf::STAGE_BYTES + Cf::STAT_BYTES + Cf::EPI_BYTES + Cf::BIAS_BYTES);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (ST + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * ST + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * ST + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * ST + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();           // 0 = leader (issues the MMAs)
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    if (p.tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"(&mapC) : "memory");
    for (int s = 0; s < ST; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc_2cta(smem_u32(tmem_slot), Cf::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                // both CTAs' barriers and TMEM exist before anyone signals
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int m_pairs = (p.nB * p.t_tiles + 1) >> 1;
  const int total = m_pairs * p.n_tiles;             // pair tiles
  const int kiters = p.ntaps * p.kblocks;

  if (warp == 0) {
    // ===================== TMA producer (whole warp loops, one elected lane issues) =====================
    {
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      for (int tile = pair; tile < total; tile += npairs, ++iter) {
        const int n_tile = tile % p.n_tiles, m_tile = (tile / p.n_tiles) * 2 + (int)rank;
        // an odd tile count leaves the last pair's second half empty: b == nB is out of range and zero-filled
        const int b = m_tile / p.t_tiles, t0 = (m_tile % p.t_tiles) * TC_BM;
        for (int it = 0; it < kiters; ++it) {
          const int tap = it / p.kblocks, kb = it - tap * p.kblocks;
          mbar_wait(empty_bar(stage), phase ^ 1);
          if (elect_one()) {
            if (it == 0) TC2_TRACE(0);
            if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * Cf::STAGE_BYTES);
            const uint32_t fb = mapa_u32(full_bar(stage), 0);
            const uint32_t sa = stage0 + stage * Cf::STAGE_BYTES;
            tma_load_3d_2cta(sa, &mapA, fb, kb * Gm::ROW, t0 + p.shift0 + tap, b);
            tma_load_3d_2cta(sa + TC_STAGE_A, &mapB, fb, kb * Gm::ROW, n_tile * BN + (int)rank * 128, tap);
            if (it == kiters - 1) TC2_TRACE(1);
          }
          __syncwarp();
          if (++stage == ST) { stage = 0; phase ^= 1; }
        }
      }
      // tail: every commit aimed at this CTA's empty barriers has landed before the CTA may exit
      for (int s = 0; s < ST; ++s) {
        mbar_wait(empty_bar(stage), phase ^ 1);
        if (++stage == ST) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA; whole warp loops, one elected lane issues) =====================
    if (rank == 0) {
      const uint32_t idesc = p.idesc;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int iter = 0;
      for (int tile = pair; tile < total; tile += npairs, ++iter) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);    // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int it = 0; it < kiters; ++it) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = stage0 + stage * Cf::STAGE_BYTES, sb = sa + TC_STAGE_A;
          if (elect_one()) {
            if (it == 0) TC2_TRACE(2);
            if (it == kiters - 1) TC2_TRACE(3);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_2cta<EB>(d_tmem, make_desc(sa + k * 32, 16, 1024), make_desc(sb + k * 32, 16, 1024), idesc,
                            (it > 0 || k > 0) ? 1u : 0u);
            umma_commit_2cta(empty_bar(stage), 3);
            if (it == kiters - 1) {
              umma_commit_2cta(tfull_bar(acc), 3);
              TC2_TRACE(4);
            }
          }
          __syncwarp();
          if (++stage == ST) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps: this CTA's 128 rows x 256 columns =====================
    const int q = warp & 3;                  // TMEM lane quadrant = rows [32q, 32q + 32) of this CTA's half tile
    const uint32_t tempty_leader0 = mapa_u32(tempty_bar(0), 0);
    const uint32_t epi_w = epi0 + (uint32_t)(warp - 2) * (Cf::EPI_BUFS * 4096u);   // this warp's staging tile(s)
    int acc = 0;
    uint32_t acc_phase = 0;
    int iter = 0;
    for (int tile = pair; tile < total; tile += npairs, ++iter) {
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      if (threadIdx.x == 64) TC2_TRACE(5);
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      const int n_tile = tile % p.n_tiles, m_tile = (tile / p.n_tiles) * 2 + (int)rank;
      const int b = m_tile / p.t_tiles, t_blk = (m_tile % p.t_tiles) * TC_BM + q * 32, t = t_blk + lane;
      const bool row_ok = t < p.T && b < p.nB;
      const bool rows_full = t_blk + 32 <= p.T && b < p.nB;       // warp-uniform: no row of this warp's block is clipped
      float* crow = p.C + ((size_t)b * p.T + t) * p.ldc;
      // A lone warp per scheduler is issue-bound (r01b trace: ~500 instructions and 0.9 us per 32-column chunk with per-element
      // bias loads and masks), so the common case -- full chunk, nothing to zero for the BN sums -- takes a short path: the
      // tile's bias values are parked in shared memory once and added with 8 vector loads per chunk.
      const uint32_t bias_w = bias0 + (uint32_t)(warp - 2) * (BN * 4);
      if (p.tma_store && p.bias != nullptr) {
        __syncwarp();
#pragma unroll
        for (int i = 0; i < BN / 32; ++i) {
          const int n = n_tile * BN + i * 32 + lane;
          st_shared_f32(bias_w + (i * 32 + lane) * 4, n < p.N ? __ldg(p.bias + n) : 0.f);
        }
        __syncwarp();
      }
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        float v[32];
        tmem_ld32(t_addr + c * 32, v);
        TC2_CTRACE(0);
        if (c == BN / 32 - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote_relaxed(tempty_leader0 + 8u * acc);   // the TMEM reads have completed (wait::ld)
          if (threadIdx.x == 64) TC2_TRACE(7);
        }
        if (c == 1 && threadIdx.x == 64) TC2_TRACE(6);
        const int n0 = n_tile * BN + c * 32;
        const bool fast = p.tma_store && n0 + 32 <= p.N && (rows_full || p.stats == nullptr);
        if (fast) {
          if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 bb = ld_shared_v4(bias_w + (c * 32 + j * 4) * 4);
              v[4 * j] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w;
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = n0 + j;
            float x = v[j] + ((p.bias != nullptr && n < p.N) ? __ldg(p.bias + n) : 0.f);
            v[j] = (row_ok && n < p.N) ? x : 0.f;
          }
        }
        TC2_CTRACE(1);
        if (p.tma_store) {
          // registers -> swizzled shared tile -> one TMA store (or reduce-add) of 32 rows x 128 B; rows >= T and columns
          // >= N are clipped by the TMA unit.  Thread = row: 16-byte chunk j of row r sits at chunk j ^ (r & 7).
          const uint32_t buf = epi_w + (uint32_t)(c & (Cf::EPI_BUFS - 1)) * 4096u;
          if (lane == 0) bulk_wait_read<Cf::EPI_BUFS - 1>();   // the store that last used this tile has drained it
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            st_shared_v4(buf + lane * 128 + ((j ^ (lane & 7)) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          fence_async_smem();
          __syncwarp();
          TC2_CTRACE(2);
          if (lane == 0) {
            if (p.accumulate) tma_reduce_add_3d(&mapC, buf, n0, t_blk, b);
            else tma_store_3d(&mapC, buf, n0, t_blk, b);
            bulk_commit();
          }
          TC2_CTRACE(3);
          if (p.stats != nullptr) {                      // lane = column: walk down the 32 rows of the staged tile
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int r = 0; r < 32; ++r) {
              const float x = ld_shared_f32(buf + r * 128 + ((((lane >> 2) ^ (r & 7)) << 4) | ((lane & 3) << 2)));
              s1 += x;
              s2 = fmaf(x, x, s2);
            }
            stat_s[(q * 2 + 0) * BN + c * 32 + lane] = s1;
            stat_s[(q * 2 + 1) * BN + c * 32 + lane] = s2;
          }
        } else {
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
            stat_s[(q * 2 + 0) * BN + c * 32 + lane] = s1;
            stat_s[(q * 2 + 1) * BN + c * 32 + lane] = s2;
          }
        }
      }
      if (p.stats != nullptr) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
        for (int cp = 0; cp < BN / 128; ++cp) {
          const int col = threadIdx.x - 64 + cp * 128;
          const int n = n_tile * BN + col;
          if (n < p.N) {
            double a = 0.0, bq = 0.0;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              a += (double)stat_s[(w * 2 + 0) * BN + col];
              bq += (double)stat_s[(w * 2 + 1) * BN + col];
            }
            atomicAdd(p.stats + n, a);
            atomicAdd(p.stats + p.N + n, bq);
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (p.tma_store && lane == 0) bulk_wait_all();       // the staged tiles must be read out before the CTA's smem goes away
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                // no CTA exits (or frees TMEM) while its peer may still touch it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, Cf::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------
// MODE_TN on CTA pairs: dW tile = 256 dY channels (two 128-channel tiles, one per CTA) x 256 X channels.  Each CTA stages
// ITS 128 dY channels and ITS 128-channel half of the X tile per 64-frame block: 32 KB per stage instead of 48 KB.  Needs
// whole grouped boxes (channel counts multiples of the tiles) and an even number of 128-channel dY tiles.  Barrier protocol
// as in tc_gemm2_nt_kernel; partials are written with direct stores (the epilogue is far from the critical path here).
// ---------------------------------------------------------------------------------------------------
template <int EB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
tc_gemm2_tn_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const TcParams p) {
  using Gm = TcGeom<EB>;
  constexpr int BN = 256, ST = 6, STAGE_BYTES = 32768;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const uint32_t stage0 = base;
  uint64_t* bars = reinterpret_cast<uint64_t*>(gen + ST * STAGE_BYTES);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (ST + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * ST + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * ST + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * ST + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    for (int s = 0; s < ST; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc_2cta(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_pairs = p.n_tiles >> 1;                         // pairs of 128-channel dY tiles
  const int total = p.ntaps * n_pairs * p.k_tiles * p.splits;

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int item = pair; item < total; item += npairs) {
      const TnItem wi = tn_decode(item, p.k_tiles, n_pairs, p.ntaps);
      const int n_tile = wi.n_tile * 2 + (int)rank;
      const int rb0 = wi.split * p.rblocks_per_split;
      const int rb1 = min(p.rblocks, rb0 + p.rblocks_per_split);
      for (int rb = rb0; rb < rb1; ++rb) {
        const int b = rb / p.tbr, t0 = (rb % p.tbr) * Gm::RS;
        mbar_wait(empty_bar(stage), phase ^ 1);
        if (elect_one()) {
          if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * STAGE_BYTES);
          const uint32_t fb = mapa_u32(full_bar(stage), 0);
          const uint32_t sa = stage0 + stage * STAGE_BYTES;
          tma_load_4d_2cta(sa, &mapA, fb, 0, t0, n_tile * Gm::NBOX, b);
          tma_load_4d_2cta(sa + TC_STAGE_A, &mapB, fb, 0, t0 + p.shift0 + wi.tap, wi.k_tile * (BN / Gm::ROW) + (int)rank * Gm::NBOX, b);
        }
        __syncwarp();
        if (++stage == ST) { stage = 0; phase ^= 1; }
      }
    }
    for (int s = 0; s < ST; ++s) {                             // tail: all commits to this CTA's empty barriers have landed
      mbar_wait(empty_bar(stage), phase ^ 1);
      if (++stage == ST) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      const uint32_t idesc = p.idesc;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int item = pair; item < total; item += npairs) {
        const int split = tn_decode(item, p.k_tiles, n_pairs, p.ntaps).split;
        const int rb0 = split * p.rblocks_per_split;
        const int iters = max(0, min(p.rblocks, rb0 + p.rblocks_per_split) - rb0);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int it = 0; it < iters; ++it) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = stage0 + stage * STAGE_BYTES, sb = sa + TC_STAGE_A;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_2cta<EB>(d_tmem, make_desc(sa + k * Gm::KMMA * 128, Gm::BOX_BYTES, EB == 2 ? 1024 : 512, EB == 2 ? 2 : 1),
                            make_desc(sb + k * Gm::KMMA * 128, Gm::BOX_BYTES, EB == 2 ? 1024 : 512, EB == 2 ? 2 : 1), idesc,
                            (it > 0 || k > 0) ? 1u : 0u);
            umma_commit_2cta(empty_bar(stage), 3);
            if (it == iters - 1) umma_commit_2cta(tfull_bar(acc), 3);
          }
          __syncwarp();
          if (++stage == ST) { stage = 0; phase ^= 1; }
        }
        if (iters == 0) {
          if (elect_one()) umma_commit_2cta(tfull_bar(acc), 3);
          __syncwarp();
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t tempty_leader0 = mapa_u32(tempty_bar(0), 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = pair; item < total; item += npairs) {
      const TnItem wi = tn_decode(item, p.k_tiles, n_pairs, p.ntaps);
      const bool has_work = wi.split * p.rblocks_per_split < p.rblocks;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      const int n = (wi.n_tile * 2 + (int)rank) * TC_BM + row;
      float* orow = p.part + (((size_t)wi.split * p.ntaps + wi.tap) * p.N + n) * p.K;
      const bool vec = (p.K & 3) == 0;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        float v[32];
        tmem_ld32(t_addr + c * 32, v);
        if (c == BN / 32 - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote_relaxed(tempty_leader0 + 8u * acc);
        }
        const int k0 = wi.k_tile * BN + c * 32;
        if (n < p.N) {
          if (!has_work) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
          }
          if (vec && k0 + 32 <= p.K) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(orow + k0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (k0 + j < p.K) orow[k0 + j] = v[j];
          }
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------
// fp32 -> padded bf16 staging
// ---------------------------------------------------------------------------------------------------
// dst[r][c] (ld = Cp, bf16) = src[r][c] (ld = lds, fp32) for c < C, zero for C <= c < Cp; rows >= R_src are zero
// 16-bit conversions of a pair of floats for the two half formats (1 = bf16, 2 = fp16)
__device__ __forceinline__ uint32_t pack16(float a, float b, int fmt) {
  if (fmt == 2) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ uint16_t to16(float a, int fmt) {
  if (fmt == 2) {
    const __half h = __float2half_rn(a);
    return *reinterpret_cast<const uint16_t*>(&h);
  }
  const __nv_bfloat16 h = __float2bfloat16_rn(a);
  return *reinterpret_cast<const uint16_t*>(&h);
}

__global__ void cvt_pad_bf16_kernel(const float* __restrict__ src, int lds, __nv_bfloat16* __restrict__ dst, int Cp,
                                    size_t R_dst, size_t R_src, int C, int fmt = 1) {
  const size_t total = R_dst * (size_t)(Cp / 2);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / (Cp / 2);
    const int c = (int)(i % (Cp / 2)) * 2;
    float a = 0.f, b = 0.f;
    if (r < R_src) {
      if (c < C) a = src[r * lds + c];
      if (c + 1 < C) b = src[r * lds + c + 1];
    }
    reinterpret_cast<uint32_t*>(dst)[i] = pack16(a, b, fmt);
  }
}
// weights [ntaps][N][K] fp32 -> [ntaps][Np][Kp] bf16, zero padded
__global__ void cvt_pad_w_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int ntaps, int N, int K,
                                      int Np, int Kp, int fmt = 1) {
  const size_t total = (size_t)ntaps * Np * Kp;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % Kp);
    const int n = (int)((i / Kp) % Np);
    const int tap = (int)(i / ((size_t)Kp * Np));
    const float v = (n < N && k < K) ? src[((size_t)tap * N + n) * K + k] : 0.f;
    reinterpret_cast<uint16_t*>(dst)[i] = to16(v, fmt);
  }
}

__global__ void cvt_pad_w_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int ntaps, int N, int K, int Np, int Kp) {
  const size_t total = (size_t)ntaps * Np * Kp;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % Kp);
    const int n = (int)((i / Kp) % Np);
    const int tap = (int)(i / ((size_t)Kp * Np));
    dst[i] = (n < N && k < K) ? src[((size_t)tap * N + n) * K + k] : 0.f;
  }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------

static int cvt_blocks(size_t total) { return (int)std::min<size_t>(ceil_div(total, (size_t)256), (size_t)num_sms() * 16); }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: remember per device (not per process) whether the
// kernel has been configured, so a second GPU driven from the same process configures its own copy of the function
struct AttrOnce {
  std::atomic<unsigned long long> done{0};        // bit d: set on device d (devices >= 64 set it on every call)
  template <class K>
  int ensure(K kern, int smem) {
    int dev = 0;
    AVC_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && (done.load(std::memory_order_acquire) >> dev) & 1ull) return AVC_OK;
    AVC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (dev < 64) done.fetch_or(1ull << dev, std::memory_order_release);
    return AVC_OK;
  }
};

constexpr int TC_X3_CHUNK = 8;      // pipeline stages (of 4 MMAs) per TMEM accumulation chunk of the fp32x3 mode
template <int MODE, int EB, int BN, int CH = 0>
static int tc_launch(const CUtensorMap& mA, const CUtensorMap& mB, const TcParams& p, int grid, cudaStream_t st) {
  static AttrOnce attr;
  auto kern = tc_gemm_kernel<MODE, EB, BN, CH>;
  if (int rc = attr.ensure(kern, TcCfg<BN>::SMEM_BYTES)) return rc;
  kern<<<grid, TcThreads<BN, CH>::value, TcCfg<BN>::SMEM_BYTES, st>>>(mA, mB, p);
  AVC_LAUNCHED();
  return AVC_OK;
}
static unsigned long long* g_gemm_trace = nullptr;
void tc_gemm_set_trace(unsigned long long* p) { g_gemm_trace = p; }

template <int EB>
static int tc2_launch(const CUtensorMap& mA, const CUtensorMap& mB, const CUtensorMap& mC, const TcParams& p, int grid, cudaStream_t st) {
  static AttrOnce attr;
  auto kern = tc_gemm2_nt_kernel<EB>;
  if (int rc = attr.ensure(kern, Tc2Cfg::SMEM_BYTES)) return rc;
  kern<<<grid, TC_THREADS, Tc2Cfg::SMEM_BYTES, st>>>(mA, mB, mC, p);
  AVC_LAUNCHED();
  return AVC_OK;
}
// AVC_GEMM_2CTA=0 keeps every NT GEMM on the single-CTA kernel
static bool use_cta_pairs() {
  const char* e = getenv("AVC_GEMM_2CTA");   // read per call so one process can compare both kernels
  return e ? atoi(e) != 0 : true;
}
template <int EB>
static int tc2_tn_launch(const CUtensorMap& mA, const CUtensorMap& mB, const TcParams& p, int grid, cudaStream_t st) {
  static AttrOnce attr;
  constexpr int SMEM = 1024 + 6 * 32768 + 256;
  auto kern = tc_gemm2_tn_kernel<EB>;
  if (int rc = attr.ensure(kern, SMEM)) return rc;
  kern<<<grid, TC_THREADS, SMEM, st>>>(mA, mB, p);
  AVC_LAUNCHED();
  return AVC_OK;
}
template <int MODE>
static int tc_dispatch(int eb, int bn, const CUtensorMap& mA, const CUtensorMap& mB, const TcParams& p, int grid, cudaStream_t st,
                       int chunk = 0) {
  if (chunk) {
    if (eb != 4) {
      set_error("tc_gemm: chunked accumulation is built for tf32 operands");
      return AVC_ERR_UNSUPPORTED;
    }
    return bn == 256 ? tc_launch<MODE, 4, 256, TC_X3_CHUNK>(mA, mB, p, grid, st) : tc_launch<MODE, 4, 128, TC_X3_CHUNK>(mA, mB, p, grid, st);
  }
  if (eb == 2) return bn == 256 ? tc_launch<MODE, 2, 256>(mA, mB, p, grid, st) : tc_launch<MODE, 2, 128>(mA, mB, p, grid, st);
  return bn == 256 ? tc_launch<MODE, 4, 256>(mA, mB, p, grid, st) : tc_launch<MODE, 4, 128>(mA, mB, p, grid, st);
}

// tile-width choice: fewest (rounds x tile cost); a 256-wide tile costs ~1.41x a 128-wide one (smem-bound model)
static int pick_bn(int row_tiles, int N) {
  if (N <= 128) return 128;
  const int sms = num_sms();
  const double c128 = (double)ceil_div(row_tiles * ceil_div(N, 128), sms) * 1.0;
  const double c256 = (double)ceil_div(row_tiles * ceil_div(N, 256), sms) * 1.41;
  return c256 < c128 ? 256 : 128;
}

// fp32 -> padded fp32 staging for the tf32 path when the caller's strides are not 16-byte multiples
__global__ void cvt_pad_f32_kernel(const float* __restrict__ src, int lds, float* __restrict__ dst, int Cp, size_t R, int C) {
  const size_t total = R * (size_t)Cp;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / Cp;
    const int c = (int)(i % Cp);
    dst[i] = c < C ? src[r * lds + c] : 0.f;
  }
}

static inline bool direct_ok(const void* p, int ld) { return ((uintptr_t)p & 15) == 0 && (ld & 3) == 0; }
static inline bool direct16_ok(const void* p, int ld) { return ((uintptr_t)p & 15) == 0 && (ld & 7) == 0; }
// operand formats: 0 = fp32 in memory, 1 = bf16, 2 = fp16.  instruction format codes: f16 = 0, bf16 = 1, tf32 = 2
static inline int ifmt_of(int ofmt16) { return ofmt16 == 2 ? 0 : 1; }

// ---- NT -------------------------------------------------------------------------------------------
// eb = 2: both operands are staged as zero-padded bf16.  eb = 4: operands are read in place by TMA when
// their row strides are 16-byte multiples (OOB zero fill covers K/N tails), otherwise staged as padded fp32.
struct NtPlan {
  int Kp, Np, bn;
  bool stageA, stageW;
  size_t offA, offW, total;
};
static NtPlan nt_plan(const void* A, int a_fmt, int lda, const void* W, int w_fmt, int nB, int T, int N, int K, int ntaps, int eb,
                      int chunk = 0) {
  NtPlan pl;
  const int row = 128 / eb;
  pl.bn = pick_bn(nB * ceil_div(T, TC_BM), N);
  pl.Kp = round_up(K, row);
  pl.Np = round_up(N, pl.bn);
  pl.stageA = a_fmt == 0 && (eb == 2 || !direct_ok(A, lda));     // 16-bit operands are always read in place
  pl.stageW = w_fmt == 0 && (eb == 2 || !direct_ok(W, K));
  pl.offA = 0;
  pl.offW = pl.stageA ? align256((size_t)nB * T * pl.Kp * eb) : 0;
  pl.total = pl.offW + (pl.stageW ? align256((size_t)ntaps * pl.Np * pl.Kp * eb) : 0);
  return pl;
}
size_t gemm_nt_workspace_tc(int nB, int T, int N, int K, int ntaps, int eb) {
  // worst case (pointers unknown): assume staging
  const int row = 128 / eb;
  const int Kp = round_up(K, row), Np = round_up(N, 256);
  return align256((size_t)nB * T * Kp * eb) + align256((size_t)ntaps * Np * Kp * eb);
}

// A: fp32 (a_fmt 0) or already 16-bit in HBM (a_fmt 1 = bf16, 2 = fp16; needs eb == 2).  fp32 operands of an
// eb == 2 GEMM are staged to `half_fmt`.
// W: fp32 [ntaps][N][K] (w_fmt 0; staged like A), or pre-packed 16-bit [ntaps][N][ldw] (w_fmt 1/2, ldw % 8 == 0, read in place;
// a 16-bit A must then have the same format, an fp32 A is staged to it).
int gemm_nt_taps_tc(const void* Av, int a_fmt, int lda, const void* Wv, int w_fmt, int ldw, const float* bias, float* C, int ldc,
                    int nB, int T, int N, int K, int ntaps, int shift0, double* stats, int accumulate, int eb, int half_fmt, void* ws,
                    size_t ws_bytes, cudaStream_t st, int chunk, int ksplit, size_t c_split_stride) {
  const float* A = (const float*)Av;
  const float* W = (const float*)Wv;
  if (a_fmt != 0 && (eb != 2 || !direct16_ok(Av, lda))) {
    set_error("avc_gemm_nt_taps_h: 16-bit A needs a 16-byte aligned pointer and lda %% 8 == 0 (lda=%d)", lda);
    return AVC_ERR_INVALID;
  }
  if (w_fmt != 0) {
    if (eb != 2 || !direct16_ok(Wv, ldw) || ldw < K) {
      set_error("avc_gemm_nt_taps_hw: 16-bit W needs a 16-byte aligned pointer and ldw %% 8 == 0, ldw >= K (ldw=%d)", ldw);
      return AVC_ERR_INVALID;
    }
    if (a_fmt != 0 && a_fmt != w_fmt) {   // tcgen05.mma kind::f16 traps on mixed bf16 x fp16 operands
      set_error("avc_gemm_nt_taps_hw: A and W must share one 16-bit format");
      return AVC_ERR_UNSUPPORTED;
    }
    half_fmt = w_fmt;
  }
  if (stats && accumulate) {
    set_error("avc_gemm_nt_taps: chan_stats and accumulate are mutually exclusive");
    return AVC_ERR_UNSUPPORTED;
  }
  const NtPlan pl = nt_plan(Av, a_fmt, lda, Wv, w_fmt, nB, T, N, K, ntaps, eb, chunk);
  if (pl.total > 0 && (!ws || ws_bytes < pl.total)) {
    set_error("avc_gemm_nt_taps(tensor): workspace %zu < %zu", ws_bytes, pl.total);
    return AVC_ERR_WORKSPACE;
  }
  int rc = 0;
  const size_t M = (size_t)nB * T;
  const int row = 128 / eb;
  const void* Aop = Av;
  const void* Wop = W;
  uint64_t a_ld = lda, a_k = K, w_k = K, w_n = N, w_ld = w_fmt != 0 ? ldw : K;
  const int fa = a_fmt != 0 ? a_fmt : half_fmt, fw = fa;            // weights are staged to A's 16-bit format (mixed formats trap)
  if (pl.stageA) {
    void* dst = (uint8_t*)ws + pl.offA;
    if (eb == 2) cvt_pad_bf16_kernel<<<cvt_blocks(M * (pl.Kp / 2)), 256, 0, st>>>(A, lda, (__nv_bfloat16*)dst, pl.Kp, M, M, K, fa);
    else cvt_pad_f32_kernel<<<cvt_blocks(M * pl.Kp), 256, 0, st>>>(A, lda, (float*)dst, pl.Kp, M, K);
    AVC_LAUNCHED();
    Aop = dst; a_ld = pl.Kp; a_k = pl.Kp;
  }
  if (pl.stageW) {
    void* dst = (uint8_t*)ws + pl.offW;
    if (eb == 2) cvt_pad_w_bf16_kernel<<<cvt_blocks((size_t)ntaps * pl.Np * pl.Kp), 256, 0, st>>>(W, (__nv_bfloat16*)dst, ntaps, N, K, pl.Np, pl.Kp, fw);
    else cvt_pad_w_f32_kernel<<<cvt_blocks((size_t)ntaps * pl.Np * pl.Kp), 256, 0, st>>>(W, (float*)dst, ntaps, N, K, pl.Np, pl.Kp);
    AVC_LAUNCHED();
    Wop = dst; w_k = pl.Kp; w_n = pl.Np; w_ld = pl.Kp;
  }
  CUtensorMap mA, mB;
  const bool pairs = !chunk && pl.bn == 256 && use_cta_pairs();
  rc = make_map3(&mA, Aop, a_k, T, nB, a_ld, (uint64_t)T * a_ld, row, TC_BM, eb, false, fa);
  if (rc) return rc;
  rc = make_map3(&mB, Wop, w_k, w_n, ntaps, w_ld, w_n * w_ld, row, pairs ? 128 : pl.bn, eb, false, fw);
  if (rc) return rc;
  TcParams p{};
  p.nB = nB; p.T = T; p.ntaps = ntaps; p.shift0 = shift0; p.N = N; p.K = K;
  p.t_tiles = ceil_div(T, TC_BM); p.n_tiles = pl.Np / pl.bn; p.kblocks = pl.Kp / row;
  p.bias = bias; p.C = C; p.ldc = ldc; p.accumulate = accumulate; p.stats = stats;
  if (pairs) {
    p.idesc = eb == 2 ? make_idesc(256, 256, 0, 0, ifmt_of(fa), ifmt_of(fw)) : make_idesc(256, 256, 0, 0, 2);
    const int pair_tiles = ceil_div(nB * p.t_tiles, 2) * p.n_tiles;
    const int grid2 = 2 * std::min(pair_tiles, num_sms() / 2);
    CUtensorMap mC = mA;                    // placeholder when the direct-store epilogue is used
    p.tma_store = (ldc % 4 == 0) && (((uintptr_t)C & 15) == 0);
    p.trace = g_gemm_trace;
    if (p.tma_store) {
      rc = make_map3_out_f32(&mC, C, N, T, nB, ldc, (uint64_t)T * ldc, 32);
      if (rc) return rc;
    }
    return eb == 2 ? tc2_launch<2>(mA, mB, mC, p, grid2, st) : tc2_launch<4>(mA, mB, mC, p, grid2, st);
  }
  p.idesc = eb == 2 ? make_idesc(TC_BM, pl.bn, 0, 0, ifmt_of(fa), ifmt_of(fw)) : make_idesc(TC_BM, pl.bn, 0, 0, 2);
  if (ksplit > 1 && (!chunk || stats || accumulate)) {
    set_error("tc_gemm: a split reduction needs the chunked kernel and neither chan_stats nor accumulate");
    return AVC_ERR_UNSUPPORTED;
  }
  p.ksplit = ksplit > 1 ? ksplit : 1;
  p.c_split_stride = c_split_stride;
  const int tiles = nB * p.t_tiles * p.n_tiles * p.ksplit;
  const int grid = std::min(tiles, num_sms());
  return tc_dispatch<MODE_NT>(eb, pl.bn, mA, mB, p, grid, st, chunk);
}

// ---- TN -------------------------------------------------------------------------------------------
static int tn_splits_tc(int rblocks, int tiles) {
  const int sms = num_sms();
  int best = 1;
  double best_eff = 0.0;
  // few output tiles (the encoder BiLSTM's 64 x 512 / 64 x 16 weight gradients are ONE tile over 32768 rows) need many row
  // splits to occupy the chip; large outputs stop early through the efficiency test below
  const int max_splits = tiles <= 4 ? 64 : 16;
  for (int s = 1; s <= max_splits; ++s) {
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
  int Np, Kp, rs, rblocks, tiles, splits, rps, bn;
  size_t off_y, off_x, off_part, total;
};
static TnPlan tn_plan(int nB, int T, int N, int K, int ntaps, int eb, bool stage_y, bool stage_x, int chunk = 0) {
  TnPlan pl;
  pl.bn = (K % 256 == 0 || K > 640) ? 256 : 128;
  pl.Np = round_up(N, TC_BM);
  pl.Kp = round_up(K, pl.bn);
  pl.rs = eb == 2 ? 64 : 32;
  pl.rblocks = nB * ceil_div(T, pl.rs);
  pl.tiles = ntaps * (pl.Np / TC_BM) * (pl.Kp / pl.bn);
  pl.splits = tn_splits_tc(pl.rblocks, pl.tiles);
  pl.rps = ceil_div(pl.rblocks, pl.splits);
  const size_t M = (size_t)nB * T;
  pl.off_y = 0;
  pl.off_x = stage_y ? align256(M * pl.Np * eb) : 0;
  pl.off_part = pl.off_x + (stage_x ? align256(M * pl.Kp * eb) : 0);
  pl.total = pl.off_part + align256((size_t)pl.splits * ntaps * N * K * 4);
  return pl;
}
size_t gemm_tn_workspace_tc(int nB, int T, int N, int K, int ntaps, int eb, int chunk) {
  return tn_plan(nB, T, N, K, ntaps, eb, true, true, chunk).total;
}
size_t gemm_tn_workspace_h(int nB, int T, int N, int K, int ntaps, int y_fmt, int x_fmt) {
  return tn_plan(nB, T, N, K, ntaps, 2, y_fmt == 0, x_fmt == 0).total;
}
size_t gemm_nt_workspace_h(int nB, int T, int N, int K, int ntaps, int a_fmt) {
  const int Kp = round_up(K, 64), Np = round_up(N, 256);
  return (a_fmt == 0 ? align256((size_t)nB * T * Kp * 2) : 0) + align256((size_t)ntaps * Np * Kp * 2);
}

int launch_wgrad_reduce(const float* part, float* dW, int N, int K, int ntaps, int splits, int out_mode, int accumulate,
                        cudaStream_t st);

int gemm_tn_taps_tc(const void* dYv, int y_fmt, int ldy, const void* Xv, int x_fmt, int ldx, float* dW, int nB, int T, int N, int K,
                    int ntaps, int shift0, int out_mode, int accumulate, int eb, int half_fmt, void* ws, size_t ws_bytes,
                    cudaStream_t st, int chunk) {
  const float* dY = (const float*)dYv;
  const float* X = (const float*)Xv;
  if ((y_fmt != 0 && (eb != 2 || !direct16_ok(dYv, ldy))) || (x_fmt != 0 && (eb != 2 || !direct16_ok(Xv, ldx)))) {
    set_error("avc_gemm_tn_taps_h: 16-bit operands need 16-byte aligned pointers and leading dimensions %% 8 == 0");
    return AVC_ERR_INVALID;
  }
  const bool stage_y = y_fmt == 0 && (eb == 2 || !direct_ok(dY, ldy));
  const bool stage_x = x_fmt == 0 && (eb == 2 || !direct_ok(X, ldx));
  const int fy = y_fmt != 0 ? y_fmt : half_fmt, fx = x_fmt != 0 ? x_fmt : half_fmt;
  if (eb == 2 && fy != fx) {   // tcgen05.mma kind::f16 traps (illegal instruction) on mixed bf16 x fp16 operands -- measured
    set_error("avc_gemm_tn_taps_h: both operands must have the same 16-bit format");
    return AVC_ERR_UNSUPPORTED;
  }
  const TnPlan pl = tn_plan(nB, T, N, K, ntaps, eb, stage_y, stage_x, chunk);
  if (!ws || ws_bytes < pl.total) {
    set_error("avc_gemm_tn_taps(tensor): workspace %zu < %zu", ws_bytes, pl.total);
    return AVC_ERR_WORKSPACE;
  }
  int rc = 0;
  const size_t M = (size_t)nB * T;
  const int row = 128 / eb;
  float* part = (float*)((uint8_t*)ws + pl.off_part);
  const void* Yop = dYv;
  const void* Xop = Xv;
  uint64_t y_ld = ldy, y_c = N, x_ld = ldx, x_c = K;
  if (stage_y) {
    void* dst = (uint8_t*)ws + pl.off_y;
    if (eb == 2) cvt_pad_bf16_kernel<<<cvt_blocks(M * (pl.Np / 2)), 256, 0, st>>>(dY, ldy, (__nv_bfloat16*)dst, pl.Np, M, M, N, fy);
    else cvt_pad_f32_kernel<<<cvt_blocks(M * pl.Np), 256, 0, st>>>(dY, ldy, (float*)dst, pl.Np, M, N);
    AVC_LAUNCHED();
    Yop = dst; y_ld = pl.Np; y_c = pl.Np;
  }
  if (stage_x) {
    void* dst = (uint8_t*)ws + pl.off_x;
    if (eb == 2) cvt_pad_bf16_kernel<<<cvt_blocks(M * (pl.Kp / 2)), 256, 0, st>>>(X, ldx, (__nv_bfloat16*)dst, pl.Kp, M, M, K, fx);
    else cvt_pad_f32_kernel<<<cvt_blocks(M * pl.Kp), 256, 0, st>>>(X, ldx, (float*)dst, pl.Kp, M, K);
    AVC_LAUNCHED();
    Xop = dst; x_ld = pl.Kp; x_c = pl.Kp;
  }
  CUtensorMap mA, mB;
  // grouped 4-D boxes need whole 128-byte channel groups and whole tiles (no partially out-of-range group)
  const bool ga = (y_c % TC_BM) == 0, gb = (x_c % pl.bn) == 0;
  if (ga) rc = make_map4_grouped(&mA, Yop, y_c, T, nB, y_ld, row, pl.rs, TC_BM / row, eb, eb == 4, fy);
  else rc = make_map3(&mA, Yop, y_c, T, nB, y_ld, (uint64_t)T * y_ld, row, pl.rs, eb, eb == 4, fy);
  if (rc) return rc;
  // CTA pairs (256 dY channels x 256 X channels per pair tile): whole grouped boxes and an even number of dY tiles
  const bool pairs = !chunk && use_cta_pairs() && pl.bn == 256 && ga && gb && ((pl.Np / TC_BM) % 2 == 0);
  if (gb) rc = make_map4_grouped(&mB, Xop, x_c, T, nB, x_ld, row, pl.rs, (pairs ? 128 : pl.bn) / row, eb, eb == 4, fx);
  else rc = make_map3(&mB, Xop, x_c, T, nB, x_ld, (uint64_t)T * x_ld, row, pl.rs, eb, eb == 4, fx);
  if (rc) return rc;
  TcParams p{};
  p.nB = nB; p.T = T; p.ntaps = ntaps; p.shift0 = shift0; p.N = N; p.K = K;
  p.n_tiles = pl.Np / TC_BM; p.k_tiles = pl.Kp / pl.bn; p.splits = pl.splits; p.rblocks = pl.rblocks;
  p.rblocks_per_split = pl.rps; p.tbr = ceil_div(T, pl.rs); p.part = part;
  p.grouped_a = ga; p.grouped_b = gb;
  p.idesc = eb == 2 ? make_idesc(TC_BM, pl.bn, 1, 1, ifmt_of(fy), ifmt_of(fx)) : make_idesc(TC_BM, pl.bn, 1, 1, 2);
  const int items = pl.tiles * pl.splits;
  if (pairs) {
    p.idesc = eb == 2 ? make_idesc(256, 256, 1, 1, ifmt_of(fy), ifmt_of(fx)) : make_idesc(256, 256, 1, 1, 2);
    const int grid2 = 2 * std::min(items / 2, num_sms() / 2);
    rc = eb == 2 ? tc2_tn_launch<2>(mA, mB, p, grid2, st) : tc2_tn_launch<4>(mA, mB, p, grid2, st);
  } else {
    const int grid = std::min(items, num_sms());
    rc = tc_dispatch<MODE_TN>(eb, pl.bn, mA, mB, p, grid, st, chunk);
  }
  if (rc) return rc;
  return launch_wgrad_reduce(part, dW, N, K, ntaps, pl.splits, out_mode, accumulate, st);
}

}  // namespace avc
