// Persistent tensor-core LSTM recurrences (AVC_PREC_BF16): one cooperative launch runs the whole
// sequence of a layer-direction, forward or BPTT.
//
// nn.LSTM (model_vc_mel.py:90/:111 lstm1, :104/:118 lstm2): per step  gates = P_t + h_{t-1} W_hh^T.
// Grid = (batch tiles of 128 utterances) x (column tiles); every CTA keeps ITS slice of W_hh
// resident in shared memory for the whole sequence (bf16, K-major, 128B swizzle -- loaded once by
// TMA), streams the step's activation tile (h_{t-1}, or dG_{t+1} for BPTT) through a TMA ring,
// multiplies with tcgen05.mma into a TMEM accumulator, and finishes the step in the epilogue warps:
// gate nonlinearities + cell update (cell state lives in registers for the whole sequence), or the
// BPTT gate-gradient algebra.  The new h_t / dG_t slice is published in a bf16 exchange buffer and a
// per-batch-tile release counter; consumers acquire it before their next TMA loads.  Only CTAs that
// share a batch tile synchronise with each other.
//
//   forward : CTA owns BN gate columns (BN/4 hidden units, gate-interleaved), K = H
//   backward: CTA owns 16 hidden units of dh, K = 4H (dh = dG_{t+1} W_hh)
#include <stdio.h>
#include <stdlib.h>

#include <cuda_fp16.h>

#include "tc_common.cuh"

namespace avc {

constexpr int LT_THREADS = 192;
constexpr int LT_STAGE = 128 * 64 * 2;   // one activation k-block: 128 utterances x 64 bf16

struct LstmTcParams {
  int nB, T, H, K, reverse;
  int MT, NT, nBpad, stages;
  const float* P;       // fwd: (nB,T,4H) pre-activations
  float* h_seq;         // fwd: out (nB,T,H) ld ldh
  int ldh;
  float* gates;         // fwd: out / bwd: in  (nB,T,4H) activated gates
  float* c_seq;         // fwd: out / bwd: in  (nB,T,H)
  const float* dH;      // bwd: (nB,T,H) ld lddh
  int lddh;
  float* dP;            // bwd: out (nB,T,4H)
  __nv_bfloat16* xbuf;  // exchange buffer [2][nBpad][K]
  unsigned int* counters;  // [MT][64], zero-initialised: [mt][0] = whole-tile counter, or one counter per 64-column k-block
  int kflags;              // 1: per-k-block release counters, consumers stream each k-block as soon as it is published
  void* h16;            // fwd: optional 16-bit copy of h_seq (nB,T,H) contiguous, format fmt16 (1 bf16 / 2 fp16)
  void* dP16;           // bwd: optional 16-bit copy of dP (nB,T,4H)
  void* h16b;           // fwd: optional bf16 copy of h_seq (nB,T,H) contiguous, next to h16 (the weight-gradient GEMMs' operand)
  int fmt16;
  int wide;                // every row-per-thread tensor is 32-byte aligned (pointer and row stride): use 256-bit accesses
  int out_tma;             // the saved tensors leave through shared staging tiles + TMA stores (maps in LtOutMaps)
  int exp_mode;            // experiment switch (AVC_LSTM_EXP): 0 none, 1 alternate two descriptors, 2 two half boxes
  unsigned long long* trace;  // optional per-step timestamps of CTA 0 (avc_debug_set_trace), else nullptr
};

// tensor maps of the tensors a recurrence saves / emits per step (TMA stores from a shared staging tile)
struct alignas(64) LtOutMaps {
  CUtensorMap gates;   // fwd: (G, T, nB) fp32, box (32, 1, 32), 128B swizzle     bwd: dP (same shape)
  CUtensorMap c;       // fwd: c_seq (H, T, nB) fp32, box (U, 1, 32)
  CUtensorMap h;       // fwd: h_seq (H, T, nB; row stride ldh) fp32, box (U, 1, 32)
  CUtensorMap h16;     // fwd: h16 (H, T, nB) 16-bit, box (U, 1, 32)              bwd: dP16 (G, T, nB), box (64, 1, 32)
  CUtensorMap h16b;    // fwd: h16b (H, T, nB) bf16, box (U, 1, 32)
};
constexpr int LT_OUT_STAGE = 16384;      // 4 epilogue warps x 4 KB

// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256).  The epilogues read and write row-per-thread tiles: every lane
// touches its own row, so the load/store path is paid per request, and 32-byte requests (one full sector) halve them.
__device__ __forceinline__ void ldg_nc_v8(const float* p, float* v) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void stg_v8(void* p, const uint32_t* w) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]),
               "r"(w[5]), "r"(w[6]), "r"(w[7])
               : "memory");
}
__device__ __forceinline__ void stg_v8f(float* p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]),
               "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }

template <int N>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, float* v);
template <>
__device__ __forceinline__ void tmem_ld_cols<32>(uint32_t taddr, float* v) {
  float t[32];
  tmem_ld32(taddr, t);
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = t[i];
}
template <>
__device__ __forceinline__ void tmem_ld_cols<16>(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// trace slot layout per step: 0 barrier passed, 1 first TMA issued, 2 last TMA issued, 3 first k-block landed,
// 4 last k-block landed, 5 all MMAs issued, 6 epilogue woke (tfull), 7 math done, 8 published
#define LT_TRACE(slot)                                                         \
  do {                                                                         \
    if (p.trace != nullptr && blockIdx.x == 0) p.trace[(size_t)s * 16 + (slot)] = gtime(); \
  } while (0)
// all-CTA stamps for steps 64..67 (skew analysis): p.trace[16*T + (cta*4 + (s-64))*2 + k], k = 0 barrier passed, 1 published;
// the SM id of each CTA goes to p.trace[16*T + 8*gridDim.x + cta]
#define LT_TRACE_ALL(k)                                                                        \
  do {                                                                                         \
    if (p.trace != nullptr && s >= 64 && s < 68)                                               \
      p.trace[(size_t)16 * T + ((size_t)blockIdx.x * 4 + (s - 64)) * 2 + (k)] = gtime();       \
  } while (0)
__device__ __forceinline__ uint4 ld_volatile_v4(const unsigned* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

// Streams the `nkb` (<= 16) k-blocks [kb_first, kb_first + nkb) of the step's activation tile in the order in which their
// producers publish them: one 64-byte poll reads all counters, every newly complete k-block is loaded at once and its index
// is left in slot_kb[stage] for the MMA thread.  Hides the skew between the publishing CTAs behind the streaming.
template <class IssueFn>
__device__ __forceinline__ void stream_kblocks_as_published(const unsigned* cnt, unsigned target, int nkb, volatile int* slot_kb,
                                                            int& stage, uint32_t& phase, int stages, uint32_t bar_full0,
                                                            uint32_t bar_empty0, const IssueFn& issue) {
  unsigned loaded = 0;
  int nloaded = 0;
  while (nloaded < nkb) {
    unsigned c[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (q * 4 < nkb) {
        const uint4 v = ld_volatile_v4(cnt + q * 4);
        c[q * 4] = v.x; c[q * 4 + 1] = v.y; c[q * 4 + 2] = v.z; c[q * 4 + 3] = v.w;
      } else {
        c[q * 4] = c[q * 4 + 1] = c[q * 4 + 2] = c[q * 4 + 3] = 0;
      }
    }
    unsigned ready = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k)
      if (k < nkb && c[k] >= target) ready |= 1u << k;
    ready &= ~loaded;
    if (ready == 0) continue;
    fence_acq_rel_gpu();                       // order the TMA reads after the counter observations
    while (ready) {
      const int k = __ffs(ready) - 1;
      ready &= ready - 1;
      mbar_wait(bar_empty0 + 8u * stage, phase ^ 1);
      slot_kb[stage] = k;
      mbar_expect_tx(bar_full0 + 8u * stage, LT_STAGE);
      issue(stage, k);
      loaded |= 1u << k;
      ++nloaded;
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
  }
}

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// BWD = false: BN = gate columns per CTA (64 or 32).  BWD = true: BN = hidden units per CTA (16).
// CL = thread-block cluster size along the column tiles of one batch tile: the activation k-block (identical for
// all of them) is fetched ONCE per cluster -- each CTA loads 128/CL of its rows and multicasts them to its peers.
template <bool BWD, int BN, int CL>
__global__ void __launch_bounds__(LT_THREADS, 1)
lstm_tc_kernel(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapX,
               const __grid_constant__ CUtensorMap mapX2, const __grid_constant__ LtOutMaps om, const LstmTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const int kblocks = p.K / 64;
  const uint32_t w_block = BN * 128;                       // bytes of one resident weight k-block
  const uint32_t w_base = base;
  const uint32_t ring = base + kblocks * w_block;          // multiples of 1024 (BN*128 with BN >= 16 and kblocks even)
  const uint32_t out_stage = ring + p.stages * LT_STAGE;   // staging tiles of the saved tensors (when p.out_tma), 1024-aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(gen + kblocks * w_block + p.stages * LT_STAGE + (p.out_tma ? LT_OUT_STAGE : 0));
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (8 + s); };
  const uint32_t w_bar = bar0 + 8u * 16, tfull = bar0 + 8u * 17, tempty = bar0 + 8u * 18;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);
  volatile int* slot_kb = reinterpret_cast<volatile int*>(bars + 20);     // [8] k-block index held by each ring slot
  const uint32_t go_bar = bar0 + 8u * 24;     // producer -> epilogue: "the loads of the next step are issued" (see the epilogue)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = blockIdx.x % p.NT, mt = blockIdx.x / p.NT;
  const int T = p.T, H = p.H, G = 4 * p.H;
  constexpr int TM_COLS = BN < 32 ? 32 : BN;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapX) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), CL);       // every CTA of the cluster must have consumed the slot
    }
    mbar_init(w_bar, 1);
    mbar_init(tfull, 1);
    mbar_init(tempty, 4);
    mbar_init(go_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), TM_COLS);
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();        // peers' barriers are initialised before anyone multicasts into them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  unsigned* counter = p.counters + mt * 64;
  const bool kflags = CL == 1 && !BWD && p.kflags != 0;
  const uint32_t crank = CL > 1 ? cluster_ctarank() : 0;
  constexpr uint16_t cmask = (uint16_t)((1u << CL) - 1);
  constexpr int SLICE_ROWS = 128 / CL;

  if (warp == 0) {
    // ===================== TMA producer =====================
    // The whole warp walks the loops (warp-uniform operands stay in uniform registers); one elected lane issues.  Issuing
    // from inside `if (lane == 0)` costs an ELECT/R2UR loop of ~20 dependent instructions per TMA / tcgen05.mma.
    if (kflags) {
      if (lane == 0) {
        mbar_expect_tx(w_bar, kblocks * w_block);
        for (int kb = 0; kb < kblocks; ++kb) tma_load_3d(w_base + kb * w_block, &mapW, w_bar, kb * 64, nt * BN, 0);
        int stage = 0;
        uint32_t phase = 0;
        for (int s = 1; s < T; ++s) {
          const int row0 = ((s - 1) & 1) * p.nBpad + mt * 128;
          constexpr int U_PUB = BN / 4;                            // units published per CTA -> 64 / U_PUB CTAs per k-block
          stream_kblocks_as_published(counter, (unsigned)s * (64 / U_PUB), kblocks, slot_kb, stage, phase, p.stages,
                                      full_bar(0), empty_bar(0), [&](int st, int kb) {
                                        tma_load_3d(ring + st * LT_STAGE, &mapX, full_bar(st), kb * 64, row0, 0);
                                      });
          mbar_arrive(go_bar);
        }
      }
    } else {
      if (elect_one()) {
        mbar_expect_tx(w_bar, kblocks * w_block);
        for (int kb = 0; kb < kblocks; ++kb) tma_load_3d(w_base + kb * w_block, &mapW, w_bar, kb * 64, nt * BN, 0);
      }
      __syncwarp();
      int stage = 0;
      uint32_t phase = 0;
      for (int s = 1; s < T; ++s) {
        const int row0 = ((s - 1) & 1) * p.nBpad + mt * 128;
        const unsigned target = (unsigned)s * (unsigned)p.NT;     // every column tile has published step s-1
        if (lane == 0) {
          while (ld_acquire(counter) < target) {
          }
          LT_TRACE(0);
          LT_TRACE_ALL(0);
          if (p.trace != nullptr && s == 64) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            p.trace[(size_t)16 * T + (size_t)8 * gridDim.x + blockIdx.x] = smid;
          }
          if (p.exp_mode == 4) fence_proxy_async();   // writer-side proxy fence + release/acquire order the TMA reads
        }
        __syncwarp();
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          if (elect_one()) {
            mbar_expect_tx(full_bar(stage), LT_STAGE);
            if (CL == 1) {
              tma_load_3d(ring + stage * LT_STAGE, &mapX, full_bar(stage), kb * 64, row0, 0);
            } else
              tma_load_3d_mc(ring + stage * LT_STAGE + crank * SLICE_ROWS * 128, &mapX, full_bar(stage), kb * 64,
                             row0 + crank * SLICE_ROWS, 0, cmask);
            if (kb == 0) LT_TRACE(1);
            if (kb == kblocks - 1) {
              LT_TRACE(2);
              mbar_arrive(go_bar);          // every load of this step is in flight: the epilogue may use the LSU again
            }
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc(128, BN, 0, 0);
    mbar_wait(w_bar, 0);
    int stage = 0;
    uint32_t phase = 0;
    for (int s = 0; s < T; ++s) {
      mbar_wait(tempty, (s & 1) ^ 1);
      tc_fence_after();
      if (s == 0) {
        if (elect_one()) mbar_arrive(tfull);               // h_{-1} = 0 / no later step: nothing to multiply
        __syncwarp();
        continue;
      }
      for (int kb0 = 0; kb0 < kblocks; ++kb0) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        int kb = kb0;                                      // k-blocks arrive in order unless they are streamed as published
        if (kflags) kb = slot_kb[stage];
        const uint32_t sa = ring + stage * LT_STAGE, sb = w_base + kb * w_block;
        if (elect_one()) {
          if (kb0 == 0) LT_TRACE(3);
          if (kb0 == kblocks - 1) LT_TRACE(4);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_base, make_desc(sa + k * 32, 16, 1024), make_desc(sb + k * 32, 16, 1024), idesc,
                     (kb0 > 0 || k > 0) ? 1u : 0u);
          if (CL == 1) umma_commit(empty_bar(stage));
          else umma_commit_mc(empty_bar(stage), cmask);
          if (kb0 == kblocks - 1) {
            umma_commit(tfull);
            LT_TRACE(5);
          }
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int b = mt * 128 + row;
    const bool live = b < p.nB;
    const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    if (!BWD) {
      constexpr int U = BN / 4;                 // hidden units owned per thread-row
      const int n0 = nt * BN, u0 = nt * U;
      float c[U];
#pragma unroll
      for (int i = 0; i < U; ++i) c[i] = 0.f;
      for (int s = 0; s < T; ++s) {
        const int t = p.reverse ? T - 1 - s : s;
        const size_t rowi = (size_t)b * T + t;
        float pre[BN];
        if ((p.exp_mode & 64) && s > 0) mbar_wait(go_bar, (s - 1) & 1);   // experiment: P loads only after the step's TMA loads are issued
        if (live) {
          if (p.wide) {
#pragma unroll
            for (int j = 0; j < BN; j += 8) ldg_nc_v8(p.P + rowi * G + n0 + j, &pre[j]);
          } else {
#pragma unroll
            for (int j = 0; j < BN; j += 4)
              *reinterpret_cast<float4*>(&pre[j]) = __ldg(reinterpret_cast<const float4*>(p.P + rowi * G + n0 + j));
          }
        } else {
#pragma unroll
          for (int j = 0; j < BN; ++j) pre[j] = 0.f;
        }
        mbar_wait(tfull, s & 1);
        if (threadIdx.x == 64) LT_TRACE(6);
        tc_fence_after();
        if (s > 0) {
          float d[BN];
#pragma unroll
          for (int cc = 0; cc < BN / 32; ++cc) tmem_ld_cols<32>(t_addr + cc * 32, d + cc * 32);
#pragma unroll
          for (int j = 0; j < BN; ++j) pre[j] += d[j];
        }
        tc_fence_before();
        if (!BWD) {
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty);       // the accumulator is in registers: the next step's MMAs may overwrite it
        }
        alignas(32) __nv_bfloat16 hb[U];
        float hf[U];
#pragma unroll
        for (int i = 0; i < U; ++i) {
          const float gi = sigmoid_fast(pre[4 * i + 0]);
          const float gf = sigmoid_fast(pre[4 * i + 1]);
          const float gg = tanh_fast(pre[4 * i + 2]);
          const float go = sigmoid_fast(pre[4 * i + 3]);
          c[i] = fmaf(gf, c[i], gi * gg);
          hf[i] = go * tanh_fast(c[i]);
          hb[i] = __float2bfloat16_rn(hf[i]);
          pre[4 * i + 0] = gi; pre[4 * i + 1] = gf; pre[4 * i + 2] = gg; pre[4 * i + 3] = go;
        }
        // publish h_t first: bf16 slice -> proxy fence -> gpu-scope fence -> CTA barrier -> one release add.
        // The (much larger) fp32 tensors saved for BPTT are stored after the release, off the critical path.
        if (live) {
          __nv_bfloat16* xb = p.xbuf + ((size_t)(s & 1) * p.nBpad + b) * p.K + u0;
          if (U == 16) {                                   // 32 bytes: one request
            stg_v8(xb, reinterpret_cast<const uint32_t*>(hb));
          } else {
#pragma unroll
            for (int i = 0; i < U; i += 8) *reinterpret_cast<uint4*>(xb + i) = *reinterpret_cast<const uint4*>(&hb[i]);
          }
        }
        if (threadIdx.x == 64) LT_TRACE(7);
        asm volatile("bar.sync 1, 128;" ::: "memory");     // all 128 rows' slices are written (CTA scope)
        if (threadIdx.x == 64) {                             // one cumulative gpu-scope release for the CTA
          fence_proxy_async();
          red_release_add(kflags ? counter + (u0 >> 6) : counter, 1u);   // red.release.gpu is the cumulative gpu-scope release
          LT_TRACE(8);
          LT_TRACE_ALL(1);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");     // bulk fp32 stores below must not queue ahead of that fence
        // The tensors saved for BPTT (48 KB per CTA and step, row-per-thread stores) occupy the SM's load/store path for
        // ~2-4 us.  Issued right after the publish they sit in front of the producer's ld.acquire polls: the all-CTA trace
        // (scripts/bench_lstm.py) showed the LAST publisher of a step noticing the completed barrier up to 4.6 us late, every
        // step, which set the period.  They are therefore held back until the producer has seen the barrier and issued the
        // next step's loads, and then overlap with the streaming.
        if ((p.exp_mode & 16) && s + 1 < T) mbar_wait(go_bar, s & 1);
        if (p.out_tma) {
          // Saved tensors: registers -> this warp's 4 KB staging tile -> TMA store, in four rounds.  As plain row-per-thread
          // stores (3300 16-byte requests per CTA and step) they held the SM's load/store path for ~4 us right when the
          // producer polls for the next step: the LAST publisher of every step noticed the completed barrier up to 4.6 us
          // late (all-CTA trace, scripts/bench_lstm.py) and 12.3 us steps ran in 9.2 us without them.  Rows of padded
          // utterances (b >= nB) are clipped by the TMA unit.
          const uint32_t buf = out_stage + (uint32_t)(warp - 2) * 4096u;
          const int brow = mt * 128 + q * 32;
          const bool save_bptt = p.gates != nullptr;                   // inference (no backward): only h leaves the kernel
#pragma unroll
          for (int half = 0; half < BN / 32; ++half) {                 // gates: 32 columns (one 128-byte swizzle row) per round
            if (!save_bptt) break;
            if (lane == 0) bulk_wait_read<0>();
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j)
              st_shared_v4(buf + lane * 128 + ((j ^ (lane & 7)) << 4), pre[half * 32 + 4 * j], pre[half * 32 + 4 * j + 1],
                           pre[half * 32 + 4 * j + 2], pre[half * 32 + 4 * j + 3]);
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_3d(&om.gates, buf, n0 + half * 32, t, brow);
              bulk_commit();
            }
          }
          if (lane == 0) bulk_wait_read<0>();                            // c_t and h_t: two [32][U] fp32 tiles
          __syncwarp();
#pragma unroll
          for (int i = 0; i < U; i += 4) {
            st_shared_v4(buf + lane * (U * 4) + i * 4, c[i], c[i + 1], c[i + 2], c[i + 3]);
            st_shared_v4(buf + 2048 + lane * (U * 4) + i * 4, hf[i], hf[i + 1], hf[i + 2], hf[i + 3]);
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (save_bptt) tma_store_3d(&om.c, buf, u0, t, brow);
            tma_store_3d(&om.h, buf + 2048, u0, t, brow);
            bulk_commit();
          }
          if (p.h16 != nullptr) {                                        // 16-bit copy of h_t: one [32][U] tile
            if (lane == 0) bulk_wait_read<0>();
            __syncwarp();
#pragma unroll
            for (int i = 0; i < U; i += 8) {
              uint32_t w[4];
#pragma unroll
              for (int k2 = 0; k2 < 4; ++k2) {
                if (p.fmt16 == 2) {
                  const __half2 v2 = __floats2half2_rn(hf[i + 2 * k2], hf[i + 2 * k2 + 1]);
                  w[k2] = *reinterpret_cast<const uint32_t*>(&v2);
                } else {
                  w[k2] = (uint32_t)*reinterpret_cast<const uint16_t*>(&hb[i + 2 * k2]) |
                          ((uint32_t)*reinterpret_cast<const uint16_t*>(&hb[i + 2 * k2 + 1]) << 16);
                }
              }
              st_shared_v4(buf + lane * (U * 2) + i * 2, __uint_as_float(w[0]), __uint_as_float(w[1]), __uint_as_float(w[2]),
                           __uint_as_float(w[3]));
              if (p.h16b != nullptr) {                                     // bf16 copy: a second [32][U] tile, same round
                const uint4 hv = *reinterpret_cast<const uint4*>(&hb[i]);
                st_shared_v4(buf + 2048 + lane * (U * 2) + i * 2, __uint_as_float(hv.x), __uint_as_float(hv.y), __uint_as_float(hv.z),
                             __uint_as_float(hv.w));
              }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_3d(&om.h16, buf, u0, t, brow);
              if (p.h16b != nullptr) tma_store_3d(&om.h16b, buf + 2048, u0, t, brow);
              bulk_commit();
            }
          }
        } else if (live && !(p.exp_mode & 8)) {
          if (p.gates != nullptr) {
#pragma unroll
            for (int j = 0; j < BN; j += 4)
              *reinterpret_cast<float4*>(p.gates + rowi * G + n0 + j) = *reinterpret_cast<const float4*>(&pre[j]);
#pragma unroll
            for (int i = 0; i < U; i += 4)
              *reinterpret_cast<float4*>(p.c_seq + rowi * H + u0 + i) = *reinterpret_cast<const float4*>(&c[i]);
          }
#pragma unroll
          for (int i = 0; i < U; i += 4)
            *reinterpret_cast<float4*>(p.h_seq + rowi * p.ldh + u0 + i) = *reinterpret_cast<const float4*>(&hf[i]);
          if (p.h16 != nullptr) {
            uint16_t* h16 = reinterpret_cast<uint16_t*>(p.h16) + rowi * H + u0;
            if (p.fmt16 == 2) {
#pragma unroll
              for (int i = 0; i < U; i += 4) {
                const __half2 lo = __floats2half2_rn(hf[i], hf[i + 1]), hi = __floats2half2_rn(hf[i + 2], hf[i + 3]);
                uint2 v;
                v.x = *reinterpret_cast<const uint32_t*>(&lo);
                v.y = *reinterpret_cast<const uint32_t*>(&hi);
                *reinterpret_cast<uint2*>(h16 + i) = v;
              }
            } else {
#pragma unroll
              for (int i = 0; i < U; i += 4) *reinterpret_cast<uint2*>(h16 + i) = *reinterpret_cast<const uint2*>(&hb[i]);
            }
          }
          if (p.h16b != nullptr) {
            uint16_t* h16b = reinterpret_cast<uint16_t*>(p.h16b) + rowi * H + u0;
#pragma unroll
            for (int i = 0; i < U; i += 4) *reinterpret_cast<uint2*>(h16b + i) = *reinterpret_cast<const uint2*>(&hb[i]);
          }
        }
      }
      if (p.out_tma && lane == 0) bulk_wait_all();      // the staged tiles are read out before the CTA's smem goes away
    } else {
      constexpr int U = BN;                      // 16 hidden units
      const int u0 = nt * U;
      float dc_rec[U];
#pragma unroll
      for (int i = 0; i < U; ++i) dc_rec[i] = 0.f;
      for (int s = 0; s < T; ++s) {
        const int t = p.reverse ? s : T - 1 - s;                 // BPTT walks against the forward direction
        const bool has_prev = s < T - 1;                         // a forward-earlier step exists
        const int t_prev = p.reverse ? t + 1 : t - 1;
        const size_t rowi = (size_t)b * T + t;
        float dh[U], ct[U], cp[U], g4[4 * U];
        if (live) {
#pragma unroll
          for (int i = 0; i < U; i += 4) {
            *reinterpret_cast<float4*>(&dh[i]) = __ldg(reinterpret_cast<const float4*>(p.dH + rowi * p.lddh + u0 + i));
            *reinterpret_cast<float4*>(&ct[i]) = __ldg(reinterpret_cast<const float4*>(p.c_seq + rowi * H + u0 + i));
            if (has_prev)
              *reinterpret_cast<float4*>(&cp[i]) =
                  __ldg(reinterpret_cast<const float4*>(p.c_seq + ((size_t)b * T + t_prev) * H + u0 + i));
            else
              cp[i] = cp[i + 1] = cp[i + 2] = cp[i + 3] = 0.f;
          }
#pragma unroll
          for (int j = 0; j < 4 * U; j += 4)
            *reinterpret_cast<float4*>(&g4[j]) = __ldg(reinterpret_cast<const float4*>(p.gates + rowi * G + 4 * u0 + j));
        } else {
#pragma unroll
          for (int i = 0; i < U; ++i) dh[i] = ct[i] = cp[i] = 0.f;
#pragma unroll
          for (int j = 0; j < 4 * U; ++j) g4[j] = 0.f;
        }
        mbar_wait(tfull, s & 1);
        if (threadIdx.x == 64) LT_TRACE(6);
        tc_fence_after();
        if (s > 0) {
          float d[U];
          tmem_ld_cols<16>(t_addr, d);
#pragma unroll
          for (int i = 0; i < U; ++i) dh[i] += d[i];
        }
        tc_fence_before();
        __nv_bfloat16 gb[4 * U];
#pragma unroll
        for (int i = 0; i < U; ++i) {
          const float gi = g4[4 * i], gf = g4[4 * i + 1], gg = g4[4 * i + 2], go = g4[4 * i + 3];
          const float tc = tanh_fast(ct[i]);
          const float dc = fmaf(dh[i] * go, 1.f - tc * tc, dc_rec[i]);
          const float di = dc * gg * gi * (1.f - gi);
          const float df = dc * cp[i] * gf * (1.f - gf);
          const float dg = dc * gi * (1.f - gg * gg);
          const float dO = dh[i] * tc * go * (1.f - go);
          dc_rec[i] = dc * gf;
          g4[4 * i] = di; g4[4 * i + 1] = df; g4[4 * i + 2] = dg; g4[4 * i + 3] = dO;
          gb[4 * i] = __float2bfloat16_rn(di); gb[4 * i + 1] = __float2bfloat16_rn(df);
          gb[4 * i + 2] = __float2bfloat16_rn(dg); gb[4 * i + 3] = __float2bfloat16_rn(dO);
        }
        if (live) {
          __nv_bfloat16* xb = p.xbuf + ((size_t)(s & 1) * p.nBpad + b) * p.K + 4 * u0;
#pragma unroll
          for (int j = 0; j < 4 * U; j += 8)
            *reinterpret_cast<uint4*>(xb + j) = *reinterpret_cast<const uint4*>(&gb[j]);
        }
        if (threadIdx.x == 64) LT_TRACE(7);
        asm volatile("bar.sync 1, 128;" ::: "memory");     // all 128 rows' slices are written (CTA scope)
        if (threadIdx.x == 64) {                             // one cumulative gpu-scope release for the CTA
          fence_proxy_async();
          red_release_add(counter, 1u);        // red.release.gpu is itself the cumulative gpu-scope release
          LT_TRACE(8);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");     // bulk fp32 stores below must not queue ahead of that fence
        if (live) {
#pragma unroll
          for (int j = 0; j < 4 * U; j += 4)
            *reinterpret_cast<float4*>(p.dP + rowi * G + 4 * u0 + j) = *reinterpret_cast<const float4*>(&g4[j]);
          if (p.dP16 != nullptr) {
            __nv_bfloat16* d16 = reinterpret_cast<__nv_bfloat16*>(p.dP16) + rowi * G + 4 * u0;
#pragma unroll
            for (int j = 0; j < 4 * U; j += 8) *reinterpret_cast<uint4*>(d16 + j) = *reinterpret_cast<const uint4*>(&gb[j]);
          }
        }
        if (lane == 0) mbar_arrive(tempty);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();        // no CTA leaves while peers may still signal its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------
// BPTT with the 4H-long reduction split over a cluster of 4 CTAs ("K-split").
// The plain BPTT kernel above gives each CTA 16 hidden units and the whole K = 4H reduction, so every SM has to
// ingest the full dG tile (128 x 4H bf16 = 1 MB at H=1024) every step -- r01 traces show that ingest (~60 GB/s per
// SM) is what bounds the step.  Here a cluster owns 64 hidden units of one batch tile; CTA rank r multiplies the
// r-th quarter of K (W slice 64 x H, resident; dG quarter 128 x H streamed: 256 KB), then the four partial
// 128 x 64 tiles are reduce-scattered through distributed shared memory: rank r receives the three foreign
// partials of ITS 16 units, adds its own, and runs the gate-gradient algebra for those units.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f4(uint32_t addr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t remote_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, 0x989680;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}

constexpr int KS_UNITS = 64;   // hidden units per cluster
constexpr int KS_CL = 4;       // cluster size = K split

// PAIRS = 2: the cluster holds TWO unit tiles (8 CTAs); the two CTAs with the same K quarter stream the same dG quarter, so each
// loads half of its rows and multicasts them to both (TMA multicast), halving the L2 requests per CTA.
template <int PAIRS>
__global__ void __launch_bounds__(LT_THREADS, 1)
lstm_tc_bwd_ks_kernel(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapX, const LstmTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const int T = p.T, H = p.H, G = 4 * p.H;
  const int kblocks = H / 64;                               // this CTA's quarter of K = 4H
  constexpr uint32_t w_block = KS_UNITS * 128;              // 8 KB
  const uint32_t w_base = base;
  const uint32_t ring = base + kblocks * w_block;
  const uint32_t red_off = kblocks * w_block + p.stages * LT_STAGE;      // [3][128][16] fp32 = 24 KB
  float* red = reinterpret_cast<float*>(gen + red_off);
  uint64_t* bars = reinterpret_cast<uint64_t*>(gen + red_off + 3 * 128 * 16 * 4);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (8 + s); };
  const uint32_t w_bar = bar0 + 8u * 16, tfull = bar0 + 8u * 17, tempty = bar0 + 8u * 18, red_full = bar0 + 8u * 19;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  volatile int* slot_kb = reinterpret_cast<volatile int*>(bars + 21);
  const uint32_t go_bar = bar0 + 8u * 26;      // producer -> epilogue: "the barrier of the next step has been seen"

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const uint32_t r = crank & 3;                             // K quarter and unit quarter owned by this CTA
  const uint32_t ug = crank >> 2;                           // which unit tile of the cluster (always 0 when PAIRS == 1)
  const int cluster_id = blockIdx.x / (KS_CL * PAIRS);
  const int UT = H / KS_UNITS;
  const int ut = (cluster_id % (UT / PAIRS)) * PAIRS + (int)ug, mt = cluster_id / (UT / PAIRS);
  constexpr int SLICE_ROWS = 128 / PAIRS;                   // rows of the activation tile this CTA loads (and multicasts)
  const uint16_t mc_mask = (uint16_t)((1u << r) | (PAIRS == 2 ? (1u << (4 + r)) : 0u));
  // reduce-scatter of the partial tiles by bulk DSMEM copies (see the epilogue); not with per-k-block release counters, whose
  // producers may refill the ring (= the send staging area) before every CTA has published
  const bool rs_bulk = !(PAIRS == 1 && p.kflags != 0) && !(p.exp_mode & 128);
  // The epilogue's row-per-thread traffic (dP16 stores of a step, dH / c / gates loads of the next: 16 + 48 KB per CTA) queues in
  // the SM's load/store path right where the producer's ld.acquire polls the step counter: the all-CTA trace showed CTAs noticing
  // the completed barrier up to 3.8 us late (median 0.5-1.5).  With defer_lsu (AVC_LSTM_EXP bit 256) the epilogue holds that
  // traffic back until the producer has seen the barrier.  Measured (r01e): detection becomes tight (0.3-0.8 us, publish spread
  // 3.6 -> 1.8 us) but the operand loads then compete with the TMA stream (first k-block lands after 1.5 us instead of 0.3) and
  // the step stays at 10.5-10.7 us: the memory system around the SM is the bound, not the ordering.  Off by default.
  const bool defer_lsu = !(PAIRS == 1 && p.kflags != 0) && (p.exp_mode & 256) != 0;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapX) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), PAIRS);     // both CTAs that receive a multicast slot must have consumed it
    }
    mbar_init(w_bar, 1);
    mbar_init(tfull, 1);
    mbar_init(tempty, 4);
    mbar_init(go_bar, 1);
    mbar_init(red_full, rs_bulk ? 1 : 3 * 4);               // bulk: my own arrive.expect_tx (+ 3 x 8 KB of complete_tx); else one arrive per epilogue warp of each of the 3 peers
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 64);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  unsigned* counter = p.counters + mt * 64;
  const unsigned per_step = (unsigned)(UT * KS_CL);
  const bool kflags = PAIRS == 1 && p.kflags != 0;     // counter[kb]: k-block kb of dG (64 gate columns) is published by exactly one CTA

  if (warp == 0) {
    // whole-warp loops, one elected lane issues (see lstm_tc_kernel)
    if (kflags) {
      if (lane == 0) {
        mbar_expect_tx(w_bar, kblocks * w_block);
        for (int kb = 0; kb < kblocks; ++kb)
          tma_load_3d(w_base + kb * w_block, &mapW, w_bar, (int)r * H + kb * 64, ut * KS_UNITS, 0);
        int stage = 0;
        uint32_t phase = 0;
        for (int s = 1; s < T; ++s) {
          const int row0 = ((s - 1) & 1) * p.nBpad + mt * 128;
          stream_kblocks_as_published(counter + r * kblocks, (unsigned)s, kblocks, slot_kb, stage, phase, p.stages, full_bar(0),
                                      empty_bar(0), [&](int st, int kb) {
                                        tma_load_3d(ring + st * LT_STAGE, &mapX, full_bar(st), (int)r * H + kb * 64, row0, 0);
                                      });
        }
      }
    } else {
      if (elect_one()) {
        mbar_expect_tx(w_bar, kblocks * w_block);
        for (int kb = 0; kb < kblocks; ++kb)
          tma_load_3d(w_base + kb * w_block, &mapW, w_bar, (int)r * H + kb * 64, ut * KS_UNITS, 0);
      }
      __syncwarp();
      int stage = 0;
      uint32_t phase = 0;
      for (int s = 1; s < T; ++s) {
        const int row0 = ((s - 1) & 1) * p.nBpad + mt * 128;
        if (lane == 0) {
          while (ld_acquire(counter) < (unsigned)s * per_step) {
          }
          if (defer_lsu) mbar_arrive(go_bar);
          LT_TRACE(0);
          LT_TRACE_ALL(0);
          if (p.trace != nullptr && s == 64) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            p.trace[(size_t)16 * T + (size_t)8 * gridDim.x + blockIdx.x] = smid;
          }
          if (p.exp_mode == 4) fence_proxy_async();
        }
        __syncwarp();
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          if (elect_one()) {
            mbar_expect_tx(full_bar(stage), LT_STAGE);
            if (PAIRS == 1)
              tma_load_3d(ring + stage * LT_STAGE, &mapX, full_bar(stage), (int)r * H + kb * 64, row0, 0);
            else
              tma_load_3d_mc(ring + stage * LT_STAGE + ug * SLICE_ROWS * 128, &mapX, full_bar(stage), (int)r * H + kb * 64,
                             row0 + (int)ug * SLICE_ROWS, 0, mc_mask);
            if (kb == 0) LT_TRACE(1);
            if (kb == kblocks - 1) LT_TRACE(2);
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc(128, KS_UNITS, 0, 0);
    mbar_wait(w_bar, 0);
    int stage = 0;
    uint32_t phase = 0;
    for (int s = 0; s < T; ++s) {
      mbar_wait(tempty, (s & 1) ^ 1);
      tc_fence_after();
      if (s == 0) {
        if (elect_one()) mbar_arrive(tfull);
        __syncwarp();
        continue;
      }
      for (int kb0 = 0; kb0 < kblocks; ++kb0) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        int kb = kb0;
        if (kflags) kb = slot_kb[stage];
        const uint32_t sa = ring + stage * LT_STAGE, sb = w_base + kb * w_block;
        if (elect_one()) {
          if (kb0 == 0) LT_TRACE(3);
          if (kb0 == kblocks - 1) LT_TRACE(4);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_base, make_desc(sa + k * 32, 16, 1024), make_desc(sb + k * 32, 16, 1024), idesc,
                     (kb0 > 0 || k > 0) ? 1u : 0u);
          if (PAIRS == 1) umma_commit(empty_bar(stage));
          else umma_commit_mc(empty_bar(stage), mc_mask);
          if (kb0 == kblocks - 1) {
            umma_commit(tfull);
            LT_TRACE(5);
          }
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int b = mt * 128 + row;
    const bool live = b < p.nB;
    const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    constexpr int U = 16;
    const int u0 = ut * KS_UNITS + (int)r * U;              // the 16 units this CTA finalises
    const uint32_t red_s = smem_u32(red);
    float dc_rec[U], c_carry[U];
#pragma unroll
    for (int i = 0; i < U; ++i) dc_rec[i] = c_carry[i] = 0.f;
    for (int s = 0; s < T; ++s) {
      const int t = p.reverse ? s : T - 1 - s;
      const bool has_prev = s < T - 1;
      const int t_prev = p.reverse ? t + 1 : t - 1;
      const size_t rowi = (size_t)b * T + t;
      float dh[U], ct[U], cp[U], g4[4 * U];
      if (live && !(p.exp_mode & 32)) {
        // c_t of this step is c_{t_prev} of the previous one (kept in registers): one c_seq read per step
        if (s == 0) {
#pragma unroll
          for (int i = 0; i < U; i += 4)
            *reinterpret_cast<float4*>(&ct[i]) = __ldg(reinterpret_cast<const float4*>(p.c_seq + rowi * H + u0 + i));
        } else {
#pragma unroll
          for (int i = 0; i < U; ++i) ct[i] = c_carry[i];
        }
        if (p.wide) {
#pragma unroll
          for (int i = 0; i < U; i += 8) {
            ldg_nc_v8(p.dH + rowi * p.lddh + u0 + i, &dh[i]);
            if (has_prev) ldg_nc_v8(p.c_seq + ((size_t)b * T + t_prev) * H + u0 + i, &cp[i]);
          }
#pragma unroll
          for (int j = 0; j < 4 * U; j += 8) ldg_nc_v8(p.gates + rowi * G + 4 * u0 + j, &g4[j]);
        } else {
#pragma unroll
          for (int i = 0; i < U; i += 4) {
            *reinterpret_cast<float4*>(&dh[i]) = __ldg(reinterpret_cast<const float4*>(p.dH + rowi * p.lddh + u0 + i));
            if (has_prev)
              *reinterpret_cast<float4*>(&cp[i]) =
                  __ldg(reinterpret_cast<const float4*>(p.c_seq + ((size_t)b * T + t_prev) * H + u0 + i));
          }
#pragma unroll
          for (int j = 0; j < 4 * U; j += 4)
            *reinterpret_cast<float4*>(&g4[j]) = __ldg(reinterpret_cast<const float4*>(p.gates + rowi * G + 4 * u0 + j));
        }
        if (!has_prev) {
#pragma unroll
          for (int i = 0; i < U; ++i) cp[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < U; ++i) c_carry[i] = cp[i];
      } else {
#pragma unroll
        for (int i = 0; i < U; ++i) dh[i] = ct[i] = cp[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 4 * U; ++j) g4[j] = 0.f;
      }
      // everything of the gate algebra that does not depend on dh is done NOW, while the MMAs of this step are still running:
      // g4 <- (d i, d f, d g, d o) per unit of dc resp. dh, ct <- d(dc)/d(dh), cp <- f (the carry factor of dc)
#pragma unroll
      for (int i = 0; i < U; ++i) {
        const float gi = g4[4 * i], gf = g4[4 * i + 1], gg = g4[4 * i + 2], go = g4[4 * i + 3];
        const float tc = tanh_fast(ct[i]);
        g4[4 * i] = gg * gi * (1.f - gi);
        g4[4 * i + 1] = cp[i] * gf * (1.f - gf);
        g4[4 * i + 2] = gi * (1.f - gg * gg);
        g4[4 * i + 3] = tc * go * (1.f - go);
        ct[i] = go * (1.f - tc * tc);
        cp[i] = gf;
      }
      if (rs_bulk && s > 0 && threadIdx.x == 64) mbar_expect_tx(red_full, 3 * 128 * 16 * 4);   // this step's three foreign partials
      mbar_wait(tfull, s & 1);
      if (threadIdx.x == 64) LT_TRACE(6);
      tc_fence_after();
      if (s > 0 && rs_bulk) {
        // my partial 128 x 64 tile: keep quarter r; quarter qq goes to cluster rank qq.  As per-thread st.shared::cluster stores
        // + release arrives this took 2.8 us of a 12.5 us step (trace): 12 remote 16-byte stores per thread, then an arrive
        // that has to wait for all of them.  Now the three foreign quarters are staged in LOCAL shared memory (the ring is
        // idle between the last MMA of a step and this CTA's publish, so its first 24 KB serve as staging) and one thread
        // hands each 8 KB tile to the copy engine, which signals the receiver's mbarrier with the byte count.
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float d[32];
          tmem_ld32(t_addr + half * 32, d);
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            const uint32_t qq = half * 2 + h2;
            if (qq == r) {
#pragma unroll
              for (int i = 0; i < U; ++i) dh[i] += d[h2 * 16 + i];
            } else {
              const uint32_t j = qq < r ? qq : qq - 1;           // staging slot of destination qq
              const uint32_t dst = ring + (j * 128 + row) * 64;
#pragma unroll
              for (int i = 0; i < U; i += 4) st_shared_v4(dst + i * 4, d[h2 * 16 + i], d[h2 * 16 + i + 1], d[h2 * 16 + i + 2], d[h2 * 16 + i + 3]);
            }
          }
        }
        tc_fence_before();
        fence_async_smem();                                       // my staged rows are visible to the copy engine
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 64) {
#pragma unroll
          for (uint32_t qq = 0; qq < KS_CL; ++qq) {
            if (qq == r) continue;
            const uint32_t j = qq < r ? qq : qq - 1;
            const uint32_t slot = r < qq ? r : r - 1;            // my index among qq's three senders
            const uint32_t peer = (ug << 2) | qq;
            asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             mapa_shared(red_s + slot * 8192u, peer)),
                         "r"(ring + j * 8192u), "r"(8192u), "r"(mapa_shared(red_full, peer))
                         : "memory");
          }
        }
      } else if (s > 0) {
        // my partial 128 x 64 tile: keep quarter r, ship quarter qq to cluster rank qq
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float d[32];
          tmem_ld32(t_addr + half * 32, d);
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            const uint32_t qq = half * 2 + h2;
            if (qq == r) {
#pragma unroll
              for (int i = 0; i < U; ++i) dh[i] += d[h2 * 16 + i];
            } else {
              const uint32_t slot = r < qq ? r : r - 1;        // my index among qq's three senders
              const uint32_t dst = mapa_shared(red_s + ((slot * 128 + row) * 16) * 4, (ug << 2) | qq);
#pragma unroll
              for (int i = 0; i < U; i += 4)
                st_cluster_f4(dst + i * 4, make_float4(d[h2 * 16 + i], d[h2 * 16 + i + 1], d[h2 * 16 + i + 2], d[h2 * 16 + i + 3]));
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
#pragma unroll
          for (uint32_t qq = 0; qq < KS_CL; ++qq)
            if (qq != r) mbar_arrive_cluster(mapa_shared(red_full, (ug << 2) | qq));
        }
      }
      if (s > 0) {
        if (threadIdx.x == 64) LT_TRACE(9);
        mbar_wait_cluster(red_full, (s - 1) & 1);
        if (threadIdx.x == 64) LT_TRACE(10);
#pragma unroll
        for (int src = 0; src < 3; ++src) {
#pragma unroll
          for (int i = 0; i < U; i += 4) {
            const float4 v = *reinterpret_cast<const float4*>(red + (src * 128 + row) * 16 + i);
            dh[i] += v.x; dh[i + 1] += v.y; dh[i + 2] += v.z; dh[i + 3] += v.w;
          }
        }
      } else {
        tc_fence_before();
      }
      alignas(32) __nv_bfloat16 gb[4 * U];
#pragma unroll
      for (int i = 0; i < U; ++i) {
        const float dc = fmaf(dh[i], ct[i], dc_rec[i]);
        const float di = dc * g4[4 * i];
        const float df = dc * g4[4 * i + 1];
        const float dg = dc * g4[4 * i + 2];
        const float dO = dh[i] * g4[4 * i + 3];
        dc_rec[i] = dc * cp[i];
        g4[4 * i] = di; g4[4 * i + 1] = df; g4[4 * i + 2] = dg; g4[4 * i + 3] = dO;
        gb[4 * i] = __float2bfloat16_rn(di); gb[4 * i + 1] = __float2bfloat16_rn(df);
        gb[4 * i + 2] = __float2bfloat16_rn(dg); gb[4 * i + 3] = __float2bfloat16_rn(dO);
      }
      if (live) {
        __nv_bfloat16* xb = p.xbuf + ((size_t)(s & 1) * p.nBpad + b) * (size_t)G + 4 * u0;
#pragma unroll
        for (int j = 0; j < 4 * U; j += 16) stg_v8(xb + j, reinterpret_cast<const uint32_t*>(&gb[j]));
      }
      if (threadIdx.x == 64) LT_TRACE(7);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 64) {
        LT_TRACE(11);
        fence_proxy_async();
        red_release_add(kflags ? counter + (u0 >> 4) : counter, 1u);   // dG columns 4*u0 .. 4*u0+63 = k-block u0/16
        LT_TRACE(8);
        LT_TRACE_ALL(1);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (defer_lsu && s + 1 < T) mbar_wait(go_bar, s & 1);
      if (live && !(p.exp_mode & 8)) {
        if (p.dP == nullptr) {
          // half mode: every consumer (dX / dW GEMMs, bias column sums) reads the bf16 copy
        } else if (p.wide) {
#pragma unroll
          for (int j = 0; j < 4 * U; j += 8) stg_v8f(p.dP + rowi * G + 4 * u0 + j, &g4[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 4 * U; j += 4)
            *reinterpret_cast<float4*>(p.dP + rowi * G + 4 * u0 + j) = *reinterpret_cast<const float4*>(&g4[j]);
        }
        if (p.dP16 != nullptr) {
          __nv_bfloat16* d16 = reinterpret_cast<__nv_bfloat16*>(p.dP16) + rowi * G + 4 * u0;
          if (p.wide) {
#pragma unroll
            for (int j = 0; j < 4 * U; j += 16) stg_v8(d16 + j, reinterpret_cast<const uint32_t*>(&gb[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 4 * U; j += 8) *reinterpret_cast<uint4*>(d16 + j) = *reinterpret_cast<const uint4*>(&gb[j]);
          }
        }
      }
      if (lane == 0) mbar_arrive(tempty);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

// fp32 -> bf16 copy of a weight matrix
__global__ void cvt_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16_rn(src[i]);
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static unsigned long long* g_trace = nullptr;
void lstm_tc_set_trace(unsigned long long* p) { g_trace = p; }

bool lstm_tc_supported(int H) { return H >= 128 && H <= 1024 && H % 64 == 0; }

static int fwd_bn(int H) { return H >= 1024 ? 64 : 32; }

struct LtPlan {
  int BN, NT, MTmax, K, stages;
  size_t smem, w_bytes, out_stage, off_w, off_x, off_cnt, total;
  int chunk;   // utterances per launch
};
static LtPlan lt_plan(int nB, int H, bool bwd) {
  LtPlan pl;
  pl.K = bwd ? 4 * H : H;
  pl.BN = bwd ? 16 : fwd_bn(H);
  pl.NT = bwd ? H / 16 : 4 * H / pl.BN;
  pl.MTmax = std::max(1, num_sms() / pl.NT);
  pl.w_bytes = (size_t)pl.BN * pl.K * 2;
  pl.out_stage = bwd ? 0 : LT_OUT_STAGE;          // forward: staging tiles for the TMA stores of the saved tensors
  const size_t budget = 227 * 1024;
  int stages = (int)((budget - 1024 - 256 - pl.w_bytes - pl.out_stage) / LT_STAGE);
  pl.stages = std::min(8, std::max(2, stages));
  pl.smem = 1024 + pl.w_bytes + (size_t)pl.stages * LT_STAGE + pl.out_stage + 256;
  pl.chunk = std::min(nB, pl.MTmax * 128);
  const int MT = ceil_div(pl.chunk, 128);
  pl.off_w = 0;
  pl.off_x = align256((size_t)4 * H * H * 2);
  pl.off_cnt = pl.off_x + align256((size_t)2 * MT * 128 * pl.K * 2);
  const int nchunks = ceil_div(nB, pl.chunk);
  pl.total = pl.off_cnt + align256((size_t)nchunks * 128 * sizeof(unsigned));
  return pl;
}
// CTA-pair recurrences (lstm_pair.cu)
bool lstm_pair_fwd_supported(int H);
size_t lstm_pair_fwd_workspace(int nB, int H);
int lstm_seq_pair_fwd(const __nv_bfloat16* Wb, const float* P, float* h_seq, int ldh, float* gates, float* c_seq, int nB, int T,
                      int H, int reverse, void* ws, size_t ws_bytes, cudaStream_t st, void* h16, int fmt16, void* h16b);

size_t lstm_tc_workspace(int nB, int T, int H, bool bwd) {
  (void)T;
  const LtPlan pl = lt_plan(nB, H, bwd);     // the K-split BPTT variant uses the same layout (xbuf [2][nBpad][4H], counters)
  size_t total = pl.total;
  if (!bwd && lstm_pair_fwd_supported(H)) total = std::max(total, pl.off_x + lstm_pair_fwd_workspace(nB, H));
  return total;
}

static bool bwd_ksplit_enabled(int H) {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("AVC_LSTM_BWD_KSPLIT");
    mode = e ? atoi(e) : 1;
  }
  return mode != 0 && H % 64 == 0;
}

// K-split BPTT launch: clusters of 4 along x
template <int PAIRS>
static int lt_launch_ks_t(const CUtensorMap& mW, const CUtensorMap& mX, LstmTcParams p, int H, size_t smem, cudaStream_t st) {
  auto kern = lstm_tc_bwd_ks_kernel<PAIRS>;
  AVC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(p.MT * (H / KS_UNITS) * KS_CL);
  cfg.blockDim = dim3(LT_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = KS_CL * PAIRS;
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  attrs[1].id = cudaLaunchAttributeCooperative;
  attrs[1].val.cooperative = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 2;
  if (PAIRS > 1) {      // all clusters must be co-resident (the CTAs spin on each other's counters)
    int max_clusters = 0;
    cfg.numAttrs = 1;
    const cudaError_t oe = cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg);
    if (oe != cudaSuccess || max_clusters * KS_CL * PAIRS < (int)cfg.gridDim.x) {
      if (getenv("AVC_DEBUG")) fprintf(stderr, "lstm bwd: clusters of 8 rejected (%s, max %d clusters, grid %u)\n", cudaGetErrorString(oe), max_clusters, cfg.gridDim.x);
      (void)cudaGetLastError();
      return AVC_ERR_UNSUPPORTED;
    }
    cfg.numAttrs = 2;
  }
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, mW, mX, p);
  if (e != cudaSuccess && PAIRS > 1) {
    if (getenv("AVC_DEBUG")) fprintf(stderr, "lstm bwd: cluster-8 launch failed: %s\n", cudaGetErrorString(e));
    (void)cudaGetLastError();
    return AVC_ERR_UNSUPPORTED;         // the caller retries with clusters of 4
  }
  AVC_CUDA(e);
  g_launches.fetch_add(1);
  return AVC_OK;
}

// K-split BPTT launch: clusters of 4 (or 8 = two unit tiles sharing the activation stream by multicast) along x
static int lt_launch_ks(const CUtensorMap& mW, const CUtensorMap& mX, const CUtensorMap& mXhalf, LstmTcParams p, int H, cudaStream_t st) {
  const size_t w_bytes = (size_t)KS_UNITS * H * 2;
  const size_t fixed = 1024 + w_bytes + 3 * 128 * 16 * 4 + 256;
  int stages = (int)((225 * 1024 - fixed) / LT_STAGE);
  stages = std::min(8, std::max(2, stages));
  const size_t smem = fixed + (size_t)stages * LT_STAGE;
  p.stages = stages;
  static int mc = -1;
  if (mc < 0) {
    const char* e = getenv("AVC_LSTM_BWD_MC");
    mc = e ? atoi(e) : 1;
  }
  // 16 clusters of 8 (H = 1024 at two batch tiles) do not fit the B200's GPCs at one CTA per SM (occupancy query: 15); smaller
  // grids do.  Remember per grid size whether the launch was refused.
  static bool refused[64] = {};
  const int gidx = std::min(63, p.MT * (H / KS_UNITS) / 2);
  if (mc != 0 && (H / KS_UNITS) % 2 == 0 && !refused[gidx]) {
    const int rc = lt_launch_ks_t<2>(mW, mXhalf, p, H, smem, st);
    if (rc != AVC_ERR_UNSUPPORTED) return rc;
    refused[gidx] = true;
  }
  return lt_launch_ks_t<1>(mW, mX, p, H, smem, st);
}

template <bool BWD, int BN, int CL>
static int lt_launch(const CUtensorMap& mW, const CUtensorMap& mX, const CUtensorMap& mX2, const LtOutMaps& om, const LstmTcParams& p,
                     const LtPlan& pl, cudaStream_t st) {
  auto kern = lstm_tc_kernel<BWD, BN, CL>;
  AVC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(p.MT * p.NT);
  cfg.blockDim = dim3(LT_THREADS);
  cfg.dynamicSmemBytes = pl.smem;
  cfg.stream = st;
  cudaLaunchAttribute attrs[2];
  int na = 0;
  if (CL > 1) {
    attrs[na].id = cudaLaunchAttributeClusterDimension;
    attrs[na].val.clusterDim.x = CL;
    attrs[na].val.clusterDim.y = 1;
    attrs[na].val.clusterDim.z = 1;
    ++na;
    cfg.attrs = attrs;
    cfg.numAttrs = na;
    // all CTAs must be co-resident (they spin on each other's counters): check that the clusters fit at once
    int max_clusters = 0;
    AVC_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg));
    if (max_clusters * CL < p.MT * p.NT) {
      set_error("lstm_tc: only %d clusters of %d CTAs can be co-resident, %d needed", max_clusters, CL, p.MT * p.NT / CL);
      return AVC_ERR_UNSUPPORTED;
    }
    // the CTAs spin on each other's global counters: the occupancy query above assumes an empty GPU, the cooperative
    // attribute makes the runtime hold the launch until the whole grid can be resident next to whatever else runs
    attrs[na].id = cudaLaunchAttributeCooperative;
    attrs[na].val.cooperative = 1;
    ++na;
    cfg.numAttrs = na;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, mW, mX, mX2, om, p);
    if (e != cudaSuccess) {                  // refused: the caller retries with a smaller cluster, down to the plain cooperative launch
      (void)cudaGetLastError();
      set_error("lstm_tc: cooperative launch of %d clusters of %d CTAs refused: %s", p.MT * p.NT / CL, CL, cudaGetErrorString(e));
      return AVC_ERR_UNSUPPORTED;
    }
    g_launches.fetch_add(1);
    return AVC_OK;
  } else {
    attrs[na].id = cudaLaunchAttributeCooperative;    // the runtime enforces co-residency
    attrs[na].val.cooperative = 1;
    ++na;
    cfg.attrs = attrs;
    cfg.numAttrs = na;
  }
  AVC_CUDA(cudaLaunchKernelEx(&cfg, kern, mW, mX, mX2, om, p));
  g_launches.fetch_add(1);
  return AVC_OK;
}

template <bool BWD, int BN>
static int lt_launch_cl(int cl, const CUtensorMap& mW, const CUtensorMap& mX, const CUtensorMap& mX2, const LtOutMaps& om,
                        const LstmTcParams& p, const LtPlan& pl, cudaStream_t st) {
  if (cl == 8) return lt_launch<BWD, BN, 8>(mW, mX, mX2, om, p, pl, st);
  if (cl == 4) return lt_launch<BWD, BN, 4>(mW, mX, mX2, om, p, pl, st);
  return lt_launch<BWD, BN, 1>(mW, mX, mX2, om, p, pl, st);
}

static int lt_cluster_size(int NT) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("AVC_LSTM_CLUSTER");
    forced = e ? atoi(e) : 0;
  }
  // default: clusters of 4 column tiles share each activation k-block by TMA multicast (every CTA loads a quarter).  Neutral
  // while the step was bound by the late barrier detection; 8.9 -> 7.9 us per step at H=1024 once that was fixed.
  int cl = forced > 0 ? forced : 4;
  while (cl > 1 && NT % cl != 0) cl >>= 1;
  return (cl == 8 || cl == 4) ? cl : 1;
}

// W: fwd -> Whh_p (4H, H);  bwd -> Whh_pT (H, 4H)  (fp32, packed/interleaved)
// w_fmt 0: W is fp32 and converted to bf16 here; 1: W is already bf16 (avc_pack_lstm_weight_h) and read in place
int lstm_seq_tc(bool bwd, const void* Wv, const float* P, float* h_seq, int ldh, float* gates, float* c_seq,
                const float* dH, int lddh, float* dP, int nB, int T, int H, int reverse, void* ws, size_t ws_bytes,
                cudaStream_t st, void* aux16, int fmt16, int w_fmt, void* aux16b) {
  const LtPlan pl = lt_plan(nB, H, bwd);
  if (!ws || ws_bytes < lstm_tc_workspace(nB, T, H, bwd)) {
    set_error("avc_lstm_seq_%s(bf16): workspace %zu < %zu", bwd ? "bwd" : "fwd", ws_bytes, lstm_tc_workspace(nB, T, H, bwd));
    return AVC_ERR_WORKSPACE;
  }
  if (pl.smem > 227 * 1024) {
    set_error("avc_lstm_seq(bf16): H=%d needs %zu bytes of shared memory", H, pl.smem);
    return AVC_ERR_UNSUPPORTED;
  }
  uint8_t* w8 = (uint8_t*)ws;
  const __nv_bfloat16* Wb = (const __nv_bfloat16*)(w8 + pl.off_w);
  __nv_bfloat16* xbuf = (__nv_bfloat16*)(w8 + pl.off_x);
  unsigned* counters = (unsigned*)(w8 + pl.off_cnt);
  if (w_fmt == 0) {
    const size_t wn = (size_t)4 * H * H;
    cvt_bf16_kernel<<<(int)std::min<size_t>(ceil_div(wn, (size_t)256), (size_t)num_sms() * 8), 256, 0, st>>>(
        (const float*)Wv, (__nv_bfloat16*)(w8 + pl.off_w), wn);
    AVC_LAUNCHED();
  } else {
    if (w_fmt != 1 || ((uintptr_t)Wv & 15) != 0) {
      set_error("avc_lstm_seq_*_h: a pre-packed W_hh must be bf16 (w_fmt 1) and 16-byte aligned");
      return AVC_ERR_INVALID;
    }
    Wb = (const __nv_bfloat16*)Wv;
  }
  if (!bwd && lstm_pair_fwd_supported(H))      // CTA pairs: one resident W_hh copy for both batch tiles, 64 SMs (lstm_pair.cu)
    return lstm_seq_pair_fwd(Wb, P, h_seq, ldh, gates, c_seq, nB, T, H, reverse, w8 + pl.off_x, ws_bytes - pl.off_x, st, aux16, fmt16,
                             aux16b);
  const int nchunks = ceil_div(nB, pl.chunk);
  AVC_CUDA(cudaMemsetAsync(counters, 0, (size_t)nchunks * 128 * sizeof(unsigned), st));
  static int kflags_mode = -1;
  if (kflags_mode < 0) {
    const char* e = getenv("AVC_LSTM_KFLAGS");
    kflags_mode = e ? atoi(e) : 0;   // measured neutral (r01): per-k-block counters stay opt-in
  }
  CUtensorMap mW, mX, mX2;
  const char* exp_env = getenv("AVC_LSTM_EXP");
  const int exp_mode = exp_env ? atoi(exp_env) : 0;
  const int w_rows = bwd ? H : 4 * H;
  int rc = make_map3(&mW, Wb, pl.K, w_rows, 1, pl.K, (uint64_t)w_rows * pl.K, 64, pl.BN);
  if (rc) return rc;
  for (int ch = 0; ch < nchunks; ++ch) {
    const int b0 = ch * pl.chunk;
    const int nb = std::min(pl.chunk, nB - b0);
    LstmTcParams p{};
    p.nB = nb; p.T = T; p.H = H; p.K = pl.K; p.reverse = reverse;
    p.MT = ceil_div(nb, 128); p.NT = pl.NT; p.nBpad = p.MT * 128; p.stages = pl.stages;
    const size_t G = 4 * (size_t)H;
    p.P = P ? P + (size_t)b0 * T * G : nullptr;
    p.h_seq = h_seq ? h_seq + (size_t)b0 * T * ldh : nullptr;
    p.ldh = ldh;
    p.gates = gates ? gates + (size_t)b0 * T * G : nullptr;       // forward without BPTT state (inference): both NULL
    p.c_seq = c_seq ? c_seq + (size_t)b0 * T * H : nullptr;
    p.dH = dH ? dH + (size_t)b0 * T * lddh : nullptr;
    p.lddh = lddh;
    p.dP = dP ? dP + (size_t)b0 * T * G : nullptr;
    p.xbuf = xbuf;
    p.counters = counters + ch * 128;
    p.kflags = (kflags_mode != 0 && H / 64 <= 16 && (H / 64) % 4 == 0 && pl.MTmax <= 2) ? 1 : 0;
    p.trace = (ch == 0) ? g_trace : nullptr;
    p.exp_mode = exp_mode;
    p.fmt16 = fmt16;
    {
      auto al32 = [](const void* q) { return ((uintptr_t)q & 31) == 0; };
      p.wide = al32(p.P) && al32(p.gates) && al32(p.c_seq) && al32(p.dH) && al32(p.dP) && (lddh % 8 == 0) && (H % 8 == 0);
    }
    p.h16 = (!bwd && aux16) ? (void*)((uint16_t*)aux16 + (size_t)b0 * T * H) : nullptr;
    p.dP16 = (bwd && aux16) ? (void*)((uint16_t*)aux16 + (size_t)b0 * T * G) : nullptr;
    p.h16b = (!bwd && aux16 && aux16b) ? (void*)((uint16_t*)aux16b + (size_t)b0 * T * H) : nullptr;
    if (bwd && bwd_ksplit_enabled(H)) {
      CUtensorMap mWk;
      rc = make_map3(&mWk, Wb, pl.K, w_rows, 1, pl.K, (uint64_t)w_rows * pl.K, 64, KS_UNITS);
      if (rc) return rc;
      rc = make_map3(&mX, xbuf, pl.K, (uint64_t)2 * p.nBpad, 1, pl.K, (uint64_t)2 * p.nBpad * pl.K, 64, 128);
      if (rc) return rc;
      rc = make_map3(&mX2, xbuf, pl.K, (uint64_t)2 * p.nBpad, 1, pl.K, (uint64_t)2 * p.nBpad * pl.K, 64, 64);
      if (rc) return rc;
      rc = lt_launch_ks(mWk, mX, mX2, p, H, st);
      if (rc) return rc;
      continue;
    }
    int cl = lt_cluster_size(p.NT);
    rc = make_map3(&mX, xbuf, pl.K, (uint64_t)2 * p.nBpad, 1, pl.K, (uint64_t)2 * p.nBpad * pl.K, 64, 128 / cl);
    if (rc) return rc;
    rc = make_map3(&mX2, xbuf, pl.K, (uint64_t)2 * p.nBpad, 1, pl.K, (uint64_t)2 * p.nBpad * pl.K, 64, exp_mode == 2 ? 64 : 128 / cl);
    if (rc) return rc;
    // forward: the saved tensors leave through TMA stores (needs 16-byte aligned tensors; ldh % 4 == 0 is an API precondition)
    LtOutMaps om{};
    om.gates = om.c = om.h = om.h16 = om.h16b = mW;
    static const bool out_tma_env = getenv("AVC_LSTM_OUT_TMA") ? atoi(getenv("AVC_LSTM_OUT_TMA")) != 0 : true;
    p.out_tma = 0;
    if (!bwd && out_tma_env && pl.out_stage > 0 && (((uintptr_t)p.gates | (uintptr_t)p.c_seq | (uintptr_t)p.h_seq | (uintptr_t)p.h16 | (uintptr_t)p.h16b) & 15) == 0) {
      const int U = pl.BN / 4;
      const uint64_t Tn = (uint64_t)T;
      if (p.gates) rc = make_map3_store(&om.gates, p.gates, 4, G, Tn, nb, G, Tn * G, 32, 1, 32, true);
      if (!rc && p.c_seq) rc = make_map3_store(&om.c, p.c_seq, 4, H, Tn, nb, H, Tn * H, U, 1, 32, false);
      if (!rc) rc = make_map3_store(&om.h, p.h_seq, 4, H, Tn, nb, ldh, Tn * ldh, U, 1, 32, false);
      if (!rc && p.h16) rc = make_map3_store(&om.h16, p.h16, 2, H, Tn, nb, H, Tn * H, U, 1, 32, false);
      if (!rc && p.h16b) rc = make_map3_store(&om.h16b, p.h16b, 2, H, Tn, nb, H, Tn * H, U, 1, 32, false);
      if (rc) return rc;
      p.out_tma = 1;
    }
    for (;;) {
      if (bwd) rc = lt_launch_cl<true, 16>(cl, mW, mX, mX2, om, p, pl, st);
      else if (pl.BN == 64) rc = lt_launch_cl<false, 64>(cl, mW, mX, mX2, om, p, pl, st);
      else rc = lt_launch_cl<false, 32>(cl, mW, mX, mX2, om, p, pl, st);
      if (rc != AVC_ERR_UNSUPPORTED || cl == 1) break;
      cl = cl == 8 ? 4 : 1;                             // clusters do not fit: retry with a smaller cluster
      rc = make_map3(&mX, xbuf, pl.K, (uint64_t)2 * p.nBpad, 1, pl.K, (uint64_t)2 * p.nBpad * pl.K, 64, 128 / cl);
      if (rc) return rc;
    }
    if (rc) return rc;
  }
  return AVC_OK;
}

}  // namespace avc
