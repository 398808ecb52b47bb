// Persistent tensor-core LSTM recurrences (bf16 operands, fp32 state): one cooperative launch runs the whole
// sequence of a layer-direction, forward or BPTT.
//
// nn.LSTM (model_vc_mel.py:90/:111 lstm1, :104/:118 lstm2): per step  gates = P_t + h_{t-1} W_hh^T.
//
// Two generations of kernels live here.
//
// (1) WEIGHT-STATIONARY kernels (lstm_tc_fwd_ws_kernel, lstm_tc_bwd_ws_kernel; r02c) -- the path every benched shape takes.
//     The W_hh slice of a CTA (128 gate rows forward; 128 units x one K quarter for BPTT) is the A operand of tcgen05.mma
//     and lives in TENSOR MEMORY for the whole sequence (896 of a row's k; the rest, 128 of 1024, in shared memory); the
//     batch is the N dimension, NB = 32 or 64 utterances per CTA.  A step ingests only NB x K bf16 per SM (128 KB at
//     H = 1024 instead of 256 KB), the whole tile is in flight at once (one or two TMA boxes, no ring), the epilogue
//     transposes gate quads across lanes with shuffles (forward) or drains unit quarters for the DSMEM reduce-scatter
//     (BPTT).  Per step at B = 256: forward 5.3 us (H = 1024) / 3.4 (H = 512), BPTT 6.2 / 3.8.
//
// (2) The ring kernel (lstm_tc_fwd_kernel) and the K-split kernel with the W slice in shared memory
//     (lstm_tc_bwd_ks_kernel): grid = (batch tiles of 128 utterances) x (column tiles); every CTA keeps ITS slice of W_hh
//     resident in shared memory (bf16, K-major, 128B swizzle -- loaded once by TMA) and streams the step's 128 x K
//     activation tile (h_{t-1}, or dG_{t+1} for BPTT) through a TMA ring into tcgen05.mma.  7.5 / 10.2 us per step at
//     H = 1024.  They remain the path for shapes (1) does not take and are what (1) is tested against: both generations
//     produce the same bits (tests/test_gpu_lstm_tc.py).  AVC_LSTM_FWD_WS=0 / AVC_LSTM_BWD_WS=0 select them.
//
// Common to both: the new h_t / dG_t slice is published in a bf16 exchange buffer and a release counter per batch tile /
// batch group; consumers acquire it before their next TMA loads.  Only CTAs that share a batch tile / group synchronise.
//
// What bounded (2), measured (profiles/r02_*): (a) ISSUING a TMA tensor load costs the producer thread 120-330 SM clocks
// whatever the box size (scripts/micro/tma_issue.cu), so tiles are fetched with few large boxes; (b) the publishers write
// with generic-proxy stores and release a counter, the CONSUMER executes the generic->async proxy fence between its acquire
// and its TMA reads; (c) SHARED-MEMORY BANDWIDTH and L2 -> SM delivery: per step and SM the TMA unit writes the 256 KB
// activation tile and tcgen05.mma reads it back plus the resident slice = 640 KB at 128 B/clk = 2.5 us, and the resident
// slice leaves room for only 5 ring stages.  Two cta_group::2 variants that keep ONE copy of W_hh for both batch tiles and
// a K-split over a CTA pair were built and measured slower (profiles/experiments/).  (1) is the decomposition with less
// ingest per SM that those measurements asked for.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include <cuda_fp16.h>

#include "tc_common.cuh"

namespace avc {

constexpr int LT_THREADS = 192;          // warp 0 TMA producer, warp 1 MMA issuer / TMEM owner, warps 2-5 epilogue
constexpr int LT_KB = 128 * 64 * 2;      // one activation k-block: 128 utterances x 64 bf16 (16 KB)
constexpr int LT_OUT_STAGE = 16384;      // forward: 4 epilogue warps x 4 KB staging for the TMA stores of the saved tensors

struct LstmTcParams {
  int nB, T, H, K, reverse;
  int MT, NT, nBpad, stages;
  int kbs;              // k-blocks per ring stage (BPTT: 2 where H % 128 == 0, else 1; forward: 1); weight-stationary forward: per TMA instruction
  int kt;               // weight-stationary forward: leading elements of every W_hh row that live in TMEM
  const float* P;       // fwd: (nB,T,4H) pre-activations
  float* h_seq;         // fwd: out (nB,T,H) ld ldh
  int ldh;
  float* gates;         // fwd: out / bwd: in  (nB,T,4H) activated gates
  float* c_seq;         // fwd: out / bwd: in  (nB,T,H)
  const float* dH;      // bwd: (nB,T,H) ld lddh
  int lddh;
  float* dP;            // bwd: optional fp32 out (nB,T,4H)
  __nv_bfloat16* xbuf;  // exchange buffer [2][nBpad][K]
  unsigned int* counters;  // one release counter per batch tile (64 words apart), zero-initialised
  void* h16;            // fwd: optional 16-bit copy of h_seq (nB,T,H) contiguous, format fmt16 (1 bf16 / 2 fp16)
  void* dP16;           // bwd: optional bf16 copy of dP (nB,T,4H)
  void* h16b;           // fwd: optional bf16 copy of h_seq (nB,T,H) contiguous, next to h16 (the weight-gradient GEMMs' operand)
  int fmt16;
  int wide;                // every row-per-thread tensor is 32-byte aligned (pointer and row stride): use 256-bit accesses
  unsigned long long* trace;  // optional per-step timestamps of CTA 0 (avc_debug_set_trace), else nullptr
};

// tensor maps of the tensors the forward recurrence saves / emits per step (TMA stores from a shared staging tile)
struct alignas(64) LtOutMaps {
  CUtensorMap gates;   // (G, T, nB) fp32, box (32, 1, 32), 128B swizzle
  CUtensorMap c;       // c_seq (H, T, nB) fp32, box (U, 1, 32)
  CUtensorMap h;       // h_seq (H, T, nB; row stride ldh) fp32, box (U, 1, 32)
  CUtensorMap h16;     // h16 (H, T, nB) 16-bit, box (U, 1, 32)
  CUtensorMap h16b;    // h16b (H, T, nB) bf16, box (U, 1, 32)
};

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

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// generic -> async proxy fence for global memory: the exchange buffer is written with generic-proxy stores by the
// publishers and read by this CTA's TMA (async proxy) loads
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// trace slot layout per step (CTA 0): 0 counter seen, 1 first TMA issued, 2 last TMA issued, 3 first stage landed,
// 4 last stage landed, 5 all MMAs issued, 6 epilogue woke (tfull), 7 math done, 8 published, 9 partials sent (BPTT),
// 10 partials received (BPTT), 11 publish barrier passed (BPTT)
#define LT_TRACE(slot)                                                         \
  do {                                                                         \
    if (p.trace != nullptr && blockIdx.x == 0) p.trace[(size_t)s * 16 + (slot)] = gtime(); \
  } while (0)

// ---------------------------------------------------------------------------------------------------
// forward.  BN = gate columns per CTA (64 or 32).  CL = thread-block cluster size along the column tiles of one
// batch tile.  The activation tile of a step is identical for all column tiles: the CTAs of a cluster take turns
// fetching it -- CTA `rank` loads the whole 16 KB k-blocks kb = rank (mod CL) and multicasts each to all CL ring slots.
// ---------------------------------------------------------------------------------------------------
template <int BN, int CL>
__global__ void __launch_bounds__(LT_THREADS, 1)
lstm_tc_fwd_kernel(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapX,
                   const __grid_constant__ LtOutMaps om, const LstmTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const int kblocks = p.K / 64;
  constexpr uint32_t w_block = BN * 128;                   // bytes of one resident weight k-block
  const uint32_t w_base = base;
  const uint32_t ring = base + kblocks * w_block;          // multiples of 1024
  const uint32_t out_stage = ring + p.stages * LT_KB;      // staging tiles of the saved tensors, 1024-aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(gen + kblocks * w_block + p.stages * LT_KB + LT_OUT_STAGE);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (8 + s); };
  const uint32_t w_bar = bar0 + 8u * 16, tfull = bar0 + 8u * 17, tempty = bar0 + 8u * 18;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);

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
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), TM_COLS);
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();        // peers' barriers are initialised before anyone multicasts into them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  unsigned* counter = p.counters + mt * 64;
  const int crank = CL > 1 ? (int)cluster_ctarank() : 0;
  constexpr uint16_t cmask = (uint16_t)((1u << CL) - 1);

  if (warp == 0) {
    // ===================== TMA producer =====================
    // The whole warp walks the loops (warp-uniform operands stay in uniform registers); one elected lane issues.
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
      }
      __syncwarp();
      fence_proxy_async_global();        // the rows were published through the generic proxy: order the TMA reads below after them
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(full_bar(stage), LT_KB);      // my slot receives the k-block, whoever fetches it
          if (CL == 1) {
            tma_load_3d(ring + stage * LT_KB, &mapX, full_bar(stage), kb * 64, row0, 0);
          } else if ((kb % CL) == crank) {
            // my own empty barrier has collected the commits of all CL consumers of this slot: every peer's slot is free
            tma_load_3d_mc(ring + stage * LT_KB, &mapX, full_bar(stage), kb * 64, row0, 0, cmask);
          }
          if (kb == 0) LT_TRACE(1);
          if (kb == kblocks - 1) LT_TRACE(2);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc(128, BN, 0, 0);
    mbar_wait(w_bar, 0);
    int stage = 0;
    uint32_t phase = 0;
    for (int s = 1; s < T; ++s) {
      mbar_wait(tempty, ((s - 1) & 1) ^ 1);
      tc_fence_after();
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sa = ring + stage * LT_KB, sb = w_base + kb * w_block;
        if (elect_one()) {
          if (kb == 0) LT_TRACE(3);
          if (kb == kblocks - 1) LT_TRACE(4);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_base, make_desc(sa + k * 32, 16, 1024), make_desc(sb + k * 32, 16, 1024), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          if (CL == 1) umma_commit(empty_bar(stage));
          else umma_commit_mc(empty_bar(stage), cmask);
          if (kb == kblocks - 1) {
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
    constexpr int U = BN / 4;                 // hidden units owned per thread-row
    const int n0 = nt * BN, u0 = nt * U;
    const uint32_t buf = out_stage + (uint32_t)(warp - 2) * 4096u;
    const int brow = mt * 128 + q * 32;
    const bool save_bptt = p.gates != nullptr;                   // inference (no backward): only h leaves the kernel
    float c[U];
#pragma unroll
    for (int i = 0; i < U; ++i) c[i] = 0.f;
    for (int s = 0; s < T; ++s) {
      const int t = p.reverse ? T - 1 - s : s;
      const size_t rowi = (size_t)b * T + t;
      float pre[BN];
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
      if (s > 0) {                                    // h_{-1} = 0: nothing to multiply at the first step
        mbar_wait(tfull, (s - 1) & 1);
        if (threadIdx.x == 64) LT_TRACE(6);
        tc_fence_after();
#pragma unroll
        for (int cc = 0; cc < BN / 32; ++cc) {
          float d[32];
          tmem_ld32(t_addr + cc * 32, d);
#pragma unroll
          for (int j = 0; j < 32; ++j) pre[cc * 32 + j] += d[j];
        }
        tc_fence_before();
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
      // publish h_t first: bf16 slice -> CTA barrier of the epilogue warps -> one cumulative gpu-scope release.
      // The (much larger) fp32 tensors saved for BPTT leave afterwards, off the critical path.
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
      if (threadIdx.x == 64) {
        red_release_add(counter, 1u);                    // red.release.gpu is the cumulative gpu-scope release
        LT_TRACE(8);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");     // the bulk stores below must not queue ahead of that release
      // Saved tensors: registers -> this warp's 4 KB staging tile -> TMA store, in rounds.  As plain row-per-thread stores
      // (3300 16-byte requests per CTA and step) they held the SM's load/store path for ~4 us right when the producer polls
      // for the next step (r01 all-CTA trace).  Rows of padded utterances (b >= nB) are clipped by the TMA unit.
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
      if (p.h16 != nullptr) {                                        // 16-bit copies of h_t: [32][U] tiles
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
    }
    if (lane == 0) bulk_wait_all();      // the staged tiles are read out before the CTA's smem goes away
  }

  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();        // no CTA leaves while peers may still signal its barriers or multicast into its ring
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------
// forward, weight-stationary form (r02c):  D^T[gate rows, batch] = W_hh[gate rows, K] . h_{t-1}[batch, K]^T.
//
// The kernel above makes every CTA take the whole 128 x K activation tile of its batch tile through shared memory every
// step (256 KB at H = 1024), because the resident W_hh slice (64 gate columns = 128 KB) is all the shared memory can hold.
// Here W_hh is the A operand and lives in TENSOR MEMORY: a CTA owns 128 gate rows (32 hidden units, gate-interleaved);
// the first KT <= 896 elements of every row are written once into TMEM (lane = gate row, two bf16 per 32-bit column:
// 448 of the 512 columns) and read from there by `tcgen05.mma` (A from TMEM), the remaining K - KT (128 at H = 1024) sit
// in shared memory as before.  The batch is the N dimension: a CTA multiplies only NB = 32 or 64 utterances (grid =
// 4H/128 row tiles x batch groups), so a step ingests NB x K bf16 = 128 KB at H = 1024 / NB = 64 and 32 KB at H = 512 /
// NB = 32 -- and the whole tile fits in shared memory at once: no ring, every TMA load of a step is in flight at the same
// time.  The accumulator is 128 lanes x NB columns: a thread of the epilogue holds ONE gate of one unit for NB utterances,
// applies its nonlinearity (one tanh.approx per element, sigmoid = 0.5 tanh(0.5 x) + 0.5 with per-lane constants: the same
// arithmetic as sigmoid_fast / tanh_fast above), a 4 x 4 register transpose across the lane quad turns that into (i, f, g, o)
// of one unit for NB / 4 utterances, and the cell update follows.  Outputs are staged in swizzled shared tiles and leave
// as 16-byte row pieces; the activated gates are float4 stores straight from registers (8 lanes = 128 contiguous bytes).
// Only the 4H/128 CTAs of a batch group synchronise (32 at H = 1024, was 64).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}
__device__ __forceinline__ uint4 ld_shared_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void st_shared_u16(uint32_t addr, uint16_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}

constexpr int WS_TM_A = 64;        // first TMEM column of the resident W_hh slice (the accumulator owns columns [0, NB))
constexpr int WS_KT_MAX = 896;     // (512 - 64) columns x 2 bf16
constexpr int WS_MAX_PARTS = 16;

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
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

// NB = utterances per batch group; EW = epilogue warps (4: one per TMEM lane quadrant; 8: two per quadrant, each taking half
// of the NB accumulator columns: 5.93 -> 5.49 us per step at H = 1024 / NB = 64).  Sharing the TMA parts of a step over a
// cluster of 2 or 4 row tiles by multicast was measured too and is 0.1-0.3 us per step SLOWER here (the L2 -> SM bytes per SM
// are what they are, and the cluster adds its own start-up skew): not kept.
template <int NB, int EW>
__global__ void __launch_bounds__(64 + EW * 32, 1)
lstm_tc_fwd_ws_kernel(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapX, const LstmTcParams p,
                      const __nv_bfloat16* __restrict__ Wg) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const int K = p.K, kblocks = K / 64, KT = p.kt, ksb = (K - KT) / 64;   // ksb: k-blocks of W_hh kept in shared memory
  constexpr uint32_t HB = NB * 128;                       // bytes of one activation k-block [NB][64] bf16
  constexpr int ET = EW * 32;                             // epilogue threads
  const uint32_t htile = base;                            // [kblocks][NB][64] bf16, 128B swizzle
  const uint32_t wsm = htile + kblocks * HB;              // [ksb][128][64] bf16, 128B swizzle
  const uint32_t st_xb = wsm + ksb * 16384;               // staging: h_t bf16 [NB][32]   (the exchange slice / bf16 copy)
  const uint32_t st_hh = st_xb + NB * 64;                 //          h_t 16-bit copy in fmt16 [NB][32]
  const uint32_t st_h = st_hh + NB * 64;                  //          h_t fp32 [NB][32]
  const uint32_t st_c = st_h + NB * 128;                  //          c_t fp32 [NB][32]
  const uint32_t bar_off = (st_c + NB * 128) - base;
  uint64_t* bars = reinterpret_cast<uint64_t*>(gen + bar_off);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int j) { return bar0 + 8u * j; };
  const uint32_t w_bar = bar0 + 8u * WS_MAX_PARTS, tfull = bar0 + 8u * (WS_MAX_PARTS + 1), wtm_bar = bar0 + 8u * (WS_MAX_PARTS + 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + WS_MAX_PARTS + 3);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = blockIdx.x % p.NT, grp = blockIdx.x / p.NT;      // gate-row tile, batch group
  const int T = p.T, H = p.H, G = 4 * p.H;
  const int kbp = p.kbs, parts = kblocks / kbp;                    // k-blocks per TMA instruction, instructions per step

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapX) : "memory");
    for (int j = 0; j < parts; ++j) mbar_init(full_bar(j), 1);
    mbar_init(w_bar, 1);
    mbar_init(tfull, 1);
    mbar_init(wtm_bar, EW);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  unsigned* counter = p.counters + grp * 64;
  const int b0 = grp * NB;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (ksb > 0 && elect_one()) {
      mbar_expect_tx(w_bar, ksb * 16384);
      for (int kb = 0; kb < ksb; ++kb) tma_load_3d(wsm + kb * 16384, &mapW, w_bar, kb * 64, nt * 128, 0);
    }
    __syncwarp();
    for (int s = 1; s < T; ++s) {
      const int row0 = ((s - 1) & 1) * p.nBpad + b0;
      const unsigned target = (unsigned)s * (unsigned)p.NT;     // every row tile of this batch group has published step s-1
      // arm this step's barriers while the group is still working: the previous phase of each has completed by the time
      // this CTA's MMAs of step s-1 were issued, long before its own publish
      if (s > 1)
        for (int j = 0; j < parts; ++j) mbar_wait(full_bar(j), s & 1);
      if (lane == 0) {
        for (int j = 0; j < parts; ++j) mbar_expect_tx(full_bar(j), kbp * HB);
        while (ld_acquire(counter) < target) {
        }
        LT_TRACE(0);
      }
      __syncwarp();
      // The tile is free: this CTA's own publish of step s-1 (counted in `target`) followed its epilogue's read of the
      // accumulator, which followed the completion of every MMA that read the tile.
      fence_proxy_async_global();
      if (elect_one()) {
        for (int j = 0; j < parts; ++j) {
          tma_load_4d(htile + j * kbp * HB, &mapX, full_bar(j), 0, row0, j * kbp, 0);
          if (j == 0) LT_TRACE(1);
        }
        LT_TRACE(2);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc(128, NB, 0, 0);
    if (ksb > 0) mbar_wait(w_bar, 0);
    mbar_wait(wtm_bar, 0);                 // the epilogue warps have written the W_hh slice into TMEM
    tc_fence_after();
    const uint32_t tmem_a = tmem_base + WS_TM_A;
    for (int s = 1; s < T; ++s) {
      for (int j = 0; j < parts; ++j) {
        mbar_wait(full_bar(j), (s - 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          if (j == 0) LT_TRACE(3);
          if (j == parts - 1) LT_TRACE(4);
          for (int kk = 0; kk < kbp; ++kk) {
            const int kb = j * kbp + kk;
            const uint32_t sb = htile + kb * HB;
            if (kb >= ksb) {        // the step's LAST k-blocks take A from TMEM (the faster form): shortest tail after the last part lands
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16_ts(tmem_base, tmem_a + (uint32_t)((kb - ksb) * 32 + k * 8), make_desc(sb + k * 32, 16, 1024), idesc, (kb > 0 || k > 0) ? 1u : 0u);
            } else {
              const uint32_t sa = wsm + kb * 16384;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16(tmem_base, make_desc(sa + k * 32, 16, 1024), make_desc(sb + k * 32, 16, 1024), idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
          }
          if (j == parts - 1) {
            umma_commit(tfull);
            LT_TRACE(5);
          }
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;                        // TMEM lane quadrant this warp may access
    const int ch = (warp - 2) >> 2;                // which share of the accumulator columns (EW = 8: two warps per quadrant)
    constexpr int NC = NB / (EW / 4);              // accumulator columns (utterances) per thread before the transpose
    constexpr int NQ = NC / 4;                     // utterances per thread after it: j = j0 + 4 m + gt
    const int j0 = ch * NC;
    const int lrow = q * 32 + lane;                // gate row inside the tile = TMEM lane
    const int ul = lrow >> 2, gt = lane & 3;       // hidden unit inside the tile, gate (i, f, g, o)
    const int et = threadIdx.x - 64;               // 0..ET-1
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const int u0 = nt * 32;
    const bool save_bptt = p.gates != nullptr;
    // ---- resident W_hh slice -> TMEM: lane = gate row, column c = elements (2c, 2c+1) of the row's k-range [K - KT, K) ----
    {
      const uint4* wrow = reinterpret_cast<const uint4*>(Wg + (size_t)(nt * 128 + lrow) * K + (K - KT));
      for (int kb = ch; kb < KT / 64; kb += EW / 4) {
        uint32_t r[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 v = __ldg(wrow + kb * 8 + i);
          r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
        }
        tmem_st32(t_lane + WS_TM_A + kb * 32, r);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(wtm_bar);
    }
    const float sc = gt == 2 ? 1.f : 0.5f, of = gt == 2 ? 0.f : 0.5f;
    float c[NQ];
#pragma unroll
    for (int m = 0; m < NQ; ++m) c[m] = 0.f;
    const int grow = nt * 128 + lrow;              // gate row in (4H)
    for (int s = 0; s < T; ++s) {
      const int t = p.reverse ? T - 1 - s : s;
      float pre[NC];
      {
        const float* pp = p.P + ((size_t)(b0 + j0) * T + t) * G + grow;
#pragma unroll
        for (int j = 0; j < NC; ++j) pre[j] = (b0 + j0 + j < p.nB) ? __ldg(pp + (size_t)j * T * G) : 0.f;
      }
      if (s > 0) {
        mbar_wait(tfull, (s - 1) & 1);
        if (et == 0) LT_TRACE(6);
        tc_fence_after();
        if constexpr (NC == 16) {
          float d[16];
          tmem_ld16(t_lane + j0, d);
#pragma unroll
          for (int j = 0; j < 16; ++j) pre[j] += d[j];
        } else {
#pragma unroll
          for (int cc = 0; cc < NC / 32; ++cc) {
            float d[32];
            tmem_ld32(t_lane + j0 + cc * 32, d);
#pragma unroll
            for (int j = 0; j < 32; ++j) pre[cc * 32 + j] += d[j];
          }
        }
        tc_fence_before();
      }
#pragma unroll
      for (int j = 0; j < NC; ++j) pre[j] = fmaf(sc, tanh_fast(sc * pre[j]), of);
      // 4 x 4 transpose across the lane quad: afterwards pre[4m + g] = gate g of (unit ul, utterance j0 + 4m + gt)
#pragma unroll
      for (int m = 0; m < NQ; ++m) {
#pragma unroll
        for (int pp2 = 0; pp2 < 4; pp2 += 2) {
          const float snd = (lane & 1) ? pre[4 * m + pp2] : pre[4 * m + pp2 + 1];
          const float rcv = __shfl_xor_sync(0xffffffffu, snd, 1);
          if (lane & 1) pre[4 * m + pp2] = rcv; else pre[4 * m + pp2 + 1] = rcv;
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float snd = (lane & 2) ? pre[4 * m + i] : pre[4 * m + 2 + i];
          const float rcv = __shfl_xor_sync(0xffffffffu, snd, 2);
          if (lane & 2) pre[4 * m + i] = rcv; else pre[4 * m + 2 + i] = rcv;
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(ET) : "memory");     // the previous step's staged tiles have been read out
      float hf[NQ];
#pragma unroll
      for (int m = 0; m < NQ; ++m) {
        const float gi = pre[4 * m], gf = pre[4 * m + 1], gg = pre[4 * m + 2], go = pre[4 * m + 3];
        c[m] = fmaf(gf, c[m], gi * gg);
        hf[m] = go * tanh_fast(c[m]);
        const int j = j0 + 4 * m + gt;
        const __nv_bfloat16 hb = __float2bfloat16_rn(hf[m]);
        // bf16 tiles: 64-byte rows, 16-byte piece p stored at (p + 2 ((j >> 1) & 1)) & 3; fp32 tiles: 128-byte rows, word w
        // stored at (w + 8 (j & 3)) & 31 -- both conflict-free for these writes and for the 16-byte row reads below
        const uint32_t o16 = (uint32_t)j * 64u + ((((uint32_t)(ul >> 3) + 2u * ((j >> 1) & 1)) & 3u) << 4) + (uint32_t)(ul & 7) * 2u;
        st_shared_u16(st_xb + o16, *reinterpret_cast<const uint16_t*>(&hb));
        if (p.h16 != nullptr) {
          uint16_t v16;
          if (p.fmt16 == 2) {
            const __half hh = __float2half_rn(hf[m]);
            v16 = *reinterpret_cast<const uint16_t*>(&hh);
          } else {
            v16 = *reinterpret_cast<const uint16_t*>(&hb);
          }
          st_shared_u16(st_hh + o16, v16);
        }
        const uint32_t o32 = (uint32_t)j * 128u + ((((uint32_t)ul + 8u * (uint32_t)gt) & 31u) << 2);
        st_shared_f32(st_h + o32, hf[m]);
        if (save_bptt) st_shared_f32(st_c + o32, c[m]);
      }
      if (et == 0) LT_TRACE(7);
      asm volatile("bar.sync 1, %0;" ::"n"(ET) : "memory");     // the staged tiles are complete
      // publish h_t: NB rows x 64 bytes of the exchange buffer in 16-byte pieces
      {
        __nv_bfloat16* xb = p.xbuf + ((size_t)(s & 1) * p.nBpad + b0) * K + u0;
#pragma unroll
        for (int i = 0; i < (NB * 4 + ET - 1) / ET; ++i) {
          const int idx = et + ET * i, j = idx >> 2, pc = idx & 3;
          if (NB * 4 % ET == 0 || idx < NB * 4) {
            const uint4 v = ld_shared_u4(st_xb + j * 64 + (((pc + 2 * ((j >> 1) & 1)) & 3) << 4));
            *reinterpret_cast<uint4*>(xb + (size_t)j * K + pc * 8) = v;
          }
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(ET) : "memory");     // all rows of the slice are written (CTA scope)
      if (et == 0) {
        red_release_add(counter, 1u);                    // cumulative gpu-scope release
        LT_TRACE(8);
      }
      // ---- tensors saved for BPTT / handed to the next layer: off the critical path ----
      if (save_bptt) {
#pragma unroll
        for (int m = 0; m < NQ; ++m) {
          const int b = b0 + j0 + 4 * m + gt;
          if (b < p.nB)
            *reinterpret_cast<float4*>(p.gates + ((size_t)b * T + t) * G + nt * 128 + ul * 4) =
                make_float4(pre[4 * m], pre[4 * m + 1], pre[4 * m + 2], pre[4 * m + 3]);
        }
      }
#pragma unroll
      for (int i = 0; i < NB * 8 / ET; ++i) {              // fp32 rows: 8 pieces of 16 bytes
        const int idx = et + ET * i, j = idx >> 3, pc = idx & 7;
        const int b = b0 + j;
        if (b < p.nB) {
          const uint32_t so = j * 128 + (((pc + 2 * (j & 3)) & 7) << 4);
          const size_t ro = (size_t)b * T + t;
          *reinterpret_cast<uint4*>(p.h_seq + ro * p.ldh + u0 + pc * 4) = ld_shared_u4(st_h + so);
          if (save_bptt) *reinterpret_cast<uint4*>(p.c_seq + ro * H + u0 + pc * 4) = ld_shared_u4(st_c + so);
        }
      }
      if (p.h16 != nullptr) {
#pragma unroll
        for (int i = 0; i < (NB * 4 + ET - 1) / ET; ++i) {
          const int idx = et + ET * i, j = idx >> 2, pc = idx & 3;
          const int b = b0 + j;
          if ((NB * 4 % ET == 0 || idx < NB * 4) && b < p.nB) {
            const uint32_t so = j * 64 + (((pc + 2 * ((j >> 1) & 1)) & 3) << 4);
            const size_t eo = ((size_t)b * T + t) * H + u0 + pc * 8;
            *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.h16) + eo) = ld_shared_u4(st_hh + so);
            if (p.h16b != nullptr) *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.h16b) + eo) = ld_shared_u4(st_xb + so);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------
// BPTT with the 4H-long reduction split over a cluster of 4 CTAs ("K-split").
// Giving each CTA 16 hidden units and the whole K = 4H reduction makes every SM ingest the full dG tile (128 x 4H bf16
// = 1 MB at H=1024) every step -- r01 traces showed that ingest to bound the step.  Here a cluster owns 64 hidden units
// of one batch tile; CTA rank r multiplies the r-th quarter of K (W slice 64 x H, resident; dG quarter 128 x H streamed:
// 256 KB), then the four partial 128 x 64 tiles are reduce-scattered through distributed shared memory: rank r receives
// the three foreign partials of ITS 16 units, adds its own, and runs the gate-gradient algebra for those units.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
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

// PAIRS = 2: the cluster holds TWO unit tiles (8 CTAs); the two CTAs with the same K quarter stream the same dG quarter and
// take turns fetching its ring stages, multicasting each to both (half of the L2 requests and TMA instructions per CTA).
template <int PAIRS>
__global__ void __launch_bounds__(LT_THREADS, 1)
lstm_tc_bwd_ks_kernel(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapX, const LstmTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const int T = p.T, H = p.H, G = 4 * p.H;
  const int kblocks = H / 64;                               // this CTA's quarter of K = 4H
  const int kbs = p.kbs;                                    // k-blocks per ring stage (one TMA box)
  const int nst = kblocks / kbs;                            // stages per step
  const uint32_t stage_bytes = (uint32_t)kbs * LT_KB;
  constexpr uint32_t w_block = KS_UNITS * 128;              // 8 KB
  const uint32_t w_base = base;
  const uint32_t ring = base + kblocks * w_block;
  const uint32_t red_off = kblocks * w_block + p.stages * stage_bytes;      // [3][128][16] fp32 = 24 KB
  float* red = reinterpret_cast<float*>(gen + red_off);
  uint64_t* bars = reinterpret_cast<uint64_t*>(gen + red_off + 3 * 128 * 16 * 4);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (8 + s); };
  const uint32_t w_bar = bar0 + 8u * 16, tfull = bar0 + 8u * 17, tempty = bar0 + 8u * 18, red_full = bar0 + 8u * 19;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const uint32_t r = crank & 3;                             // K quarter and unit quarter owned by this CTA
  const uint32_t ug = crank >> 2;                           // which unit tile of the cluster (always 0 when PAIRS == 1)
  const int cluster_id = blockIdx.x / (KS_CL * PAIRS);
  const int UT = H / KS_UNITS;
  const int ut = (cluster_id % (UT / PAIRS)) * PAIRS + (int)ug, mt = cluster_id / (UT / PAIRS);
  const uint16_t mc_mask = (uint16_t)((1u << r) | (PAIRS == 2 ? (1u << (4 + r)) : 0u));

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
    mbar_init(red_full, 1);               // my own arrive.expect_tx (+ 3 x 8 KB of complete_tx from the peers' bulk copies)
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

  if (warp == 0) {
    // whole-warp loops, one elected lane issues (see lstm_tc_fwd_kernel)
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
        LT_TRACE(0);
      }
      __syncwarp();
      fence_proxy_async_global();        // generic-proxy publishes -> my async-proxy reads
      for (int st = 0; st < nst; ++st) {
        mbar_wait(empty_bar(stage), phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(full_bar(stage), stage_bytes);
          const int g0 = ((int)r * H) / 64 + st * kbs;           // first 64-column group of the box
          if (PAIRS == 1) {
            tma_load_4d(ring + stage * stage_bytes, &mapX, full_bar(stage), 0, row0, g0, 0);
          } else if ((st & 1) == (int)ug) {
            tma_load_4d_mc(ring + stage * stage_bytes, &mapX, full_bar(stage), 0, row0, g0, 0, mc_mask);
          }
          if (st == 0) LT_TRACE(1);
          if (st == nst - 1) LT_TRACE(2);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc(128, KS_UNITS, 0, 0);
    mbar_wait(w_bar, 0);
    int stage = 0;
    uint32_t phase = 0;
    for (int s = 1; s < T; ++s) {
      mbar_wait(tempty, ((s - 1) & 1) ^ 1);
      tc_fence_after();
      for (int st = 0; st < nst; ++st) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sa0 = ring + stage * stage_bytes, sb0 = w_base + (uint32_t)(st * kbs) * w_block;
        if (elect_one()) {
          if (st == 0) LT_TRACE(3);
          if (st == nst - 1) LT_TRACE(4);
          for (int kbi = 0; kbi < kbs; ++kbi) {
            const uint32_t sa = sa0 + kbi * LT_KB, sb = sb0 + kbi * w_block;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(tmem_base, make_desc(sa + k * 32, 16, 1024), make_desc(sb + k * 32, 16, 1024), idesc,
                       (st > 0 || kbi > 0 || k > 0) ? 1u : 0u);
          }
          if (PAIRS == 1) umma_commit(empty_bar(stage));
          else umma_commit_mc(empty_bar(stage), mc_mask);
          if (st == nst - 1) {
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
      const int t = p.reverse ? s : T - 1 - s;               // BPTT walks against the forward direction
      const bool has_prev = s < T - 1;                       // a forward-earlier step exists
      const int t_prev = p.reverse ? t + 1 : t - 1;
      const size_t rowi = (size_t)b * T + t;
      float dh[U], ct[U], cp[U], g4[4 * U];
      if (live) {
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
      if (s > 0) {
        if (threadIdx.x == 64) mbar_expect_tx(red_full, 3 * 128 * 16 * 4);   // this step's three foreign partials
        mbar_wait(tfull, (s - 1) & 1);
        if (threadIdx.x == 64) LT_TRACE(6);
        tc_fence_after();
        // my partial 128 x 64 tile: keep quarter r; quarter qq goes to cluster rank qq.  The three foreign quarters are staged in
        // LOCAL shared memory (the ring is idle between the last MMA of a step and this CTA's publish, so its first 24 KB serve as
        // staging) and one thread hands each 8 KB tile to the copy engine, which signals the receiver's mbarrier with the byte count
        // (per-thread st.shared::cluster stores + release arrives took 2.8 us of a 12.5 us step in r01).
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
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty);                        // the accumulator is in registers / staged
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
          LT_TRACE(9);
        }
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
        red_release_add(counter, 1u);      // cumulative gpu-scope release of the CTA's slices (the consumers run the proxy fence)
        LT_TRACE(8);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (live) {
        if (p.dP != nullptr) {
          if (p.wide) {
#pragma unroll
            for (int j = 0; j < 4 * U; j += 8) stg_v8f(p.dP + rowi * G + 4 * u0 + j, &g4[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 4 * U; j += 4)
              *reinterpret_cast<float4*>(p.dP + rowi * G + 4 * u0 + j) = *reinterpret_cast<const float4*>(&g4[j]);
          }
        }
        if (p.dP16 != nullptr) {      // half mode: every consumer (dX / dW GEMMs, bias column sums) reads the bf16 copy
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

// ---------------------------------------------------------------------------------------------------
// BPTT, weight-stationary form (r02c):  dh^T[unit, utterance] = W_hh^T[unit, k] . dG_{t+1}[utterance, k]^T  with the
// K = 4H reduction split over a cluster of 4 CTAs as above -- but the W_hh^T slice of a CTA (128 units x its K quarter) is the
// A operand and lives in TENSOR MEMORY (plus the first 128 of 1024 k in shared memory at H = 1024), the batch is N: a CTA
// streams only NB utterances x H of dG per step (128 KB at H = 1024 / NB = 64 instead of 256 KB), all of it in flight at once.
// Grid = batch groups x (H/128 unit tiles) x 4 K ranks.  The accumulator (128 unit lanes x NB columns) is drained into four
// [NB][32] fp32 tiles, one per unit quarter; quarter qq goes to cluster rank qq with one cp.async.bulk over DSMEM (as above),
// rank r finalises units [32 r, 32 r + 32) of the tile: a thread owns 4 consecutive units of an utterance (float4 / 32-byte
// accesses of gates, dH, c and of the bf16 gate-gradient slice it publishes).  The four partial sums are added in the order
// the K-split kernel above uses for the same unit, so both kernels produce the same bits.
// ---------------------------------------------------------------------------------------------------
template <int NB, int EW>
__global__ void __launch_bounds__(64 + EW * 32, 1)
lstm_tc_bwd_ws_kernel(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapX, const LstmTcParams p,
                      const __nv_bfloat16* __restrict__ WTg) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const int T = p.T, H = p.H, G = 4 * p.H;
  const int kblocks = H / 64, KT = p.kt, ksb = (H - KT) / 64;      // this CTA's quarter of K = 4H; ksb k-blocks of W in shared memory
  constexpr uint32_t HB = NB * 128;                         // bytes of one dG k-block [NB][64] bf16
  constexpr uint32_t TILE = NB * 128;                       // bytes of one [NB][32] fp32 partial-sum tile
  constexpr int ET = EW * 32;
  const uint32_t dtile = base;                              // [kblocks][NB][64] bf16, 128B swizzle
  const uint32_t wsm = dtile + kblocks * HB;                // [ksb][128][64] bf16, 128B swizzle
  const uint32_t stg = wsm + ksb * 16384;                   // my partial sums by unit quarter: [4][NB][32] fp32
  const uint32_t red = stg + 4 * TILE;                      // the three foreign partial sums of MY unit quarter: [3][NB][32] fp32
  uint64_t* bars = reinterpret_cast<uint64_t*>(gen + (red + 3 * TILE - base));
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int j) { return bar0 + 8u * j; };
  const uint32_t w_bar = bar0 + 8u * WS_MAX_PARTS, tfull = bar0 + 8u * (WS_MAX_PARTS + 1), wtm_bar = bar0 + 8u * (WS_MAX_PARTS + 2),
                 red_full = bar0 + 8u * (WS_MAX_PARTS + 3);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + WS_MAX_PARTS + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t r = cluster_ctarank();                     // K quarter and unit quarter owned by this CTA
  const int cluster_id = blockIdx.x / KS_CL;
  const int UT = H / 128;
  const int ut = cluster_id % UT, grp = cluster_id / UT;
  const int kbp = p.kbs, parts = kblocks / kbp;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapX) : "memory");
    for (int j = 0; j < parts; ++j) mbar_init(full_bar(j), 1);
    mbar_init(w_bar, 1);
    mbar_init(tfull, 1);
    mbar_init(wtm_bar, EW);
    mbar_init(red_full, 1);               // my own arrive.expect_tx (+ 3 tiles of complete_tx from the peers' bulk copies)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  unsigned* counter = p.counters + grp * 64;
  const unsigned per_step = (unsigned)(UT * KS_CL);
  const int b0 = grp * NB;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (ksb > 0 && elect_one()) {
      mbar_expect_tx(w_bar, ksb * 16384);
      for (int kb = 0; kb < ksb; ++kb) tma_load_3d(wsm + kb * 16384, &mapW, w_bar, (int)r * H + kb * 64, ut * 128, 0);
    }
    __syncwarp();
    for (int s = 1; s < T; ++s) {
      const int row0 = ((s - 1) & 1) * p.nBpad + b0;
      if (s > 1)
        for (int j = 0; j < parts; ++j) mbar_wait(full_bar(j), s & 1);
      if (lane == 0) {
        for (int j = 0; j < parts; ++j) mbar_expect_tx(full_bar(j), kbp * HB);
        while (ld_acquire(counter) < (unsigned)s * per_step) {
        }
        LT_TRACE(0);
      }
      __syncwarp();
      fence_proxy_async_global();        // generic-proxy publishes -> my async-proxy reads
      if (elect_one()) {
        for (int j = 0; j < parts; ++j) {
          tma_load_4d(dtile + j * kbp * HB, &mapX, full_bar(j), 0, row0, (int)r * kblocks + j * kbp, 0);
          if (j == 0) LT_TRACE(1);
        }
        LT_TRACE(2);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc(128, NB, 0, 0);
    if (ksb > 0) mbar_wait(w_bar, 0);
    mbar_wait(wtm_bar, 0);
    tc_fence_after();
    const uint32_t tmem_a = tmem_base + WS_TM_A;
    for (int s = 1; s < T; ++s) {
      for (int j = 0; j < parts; ++j) {
        mbar_wait(full_bar(j), (s - 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          if (j == 0) LT_TRACE(3);
          if (j == parts - 1) LT_TRACE(4);
          for (int kk = 0; kk < kbp; ++kk) {
            const int kb = j * kbp + kk;
            const uint32_t sb = dtile + kb * HB;
            if (kb >= ksb) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16_ts(tmem_base, tmem_a + (uint32_t)((kb - ksb) * 32 + k * 8), make_desc(sb + k * 32, 16, 1024), idesc, (kb > 0 || k > 0) ? 1u : 0u);
            } else {
              const uint32_t sa = wsm + kb * 16384;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16(tmem_base, make_desc(sa + k * 32, 16, 1024), make_desc(sb + k * 32, 16, 1024), idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
          }
          if (j == parts - 1) {
            umma_commit(tfull);
            LT_TRACE(5);
          }
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;                        // TMEM lane quadrant = unit quarter this warp drains
    const int ch = (warp - 2) >> 2;
    constexpr int NC = NB / (EW / 4);              // accumulator columns drained per thread
    const int j0 = ch * NC;
    const int et = threadIdx.x - 64;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    // ---- resident W_hh^T slice -> TMEM: lane = unit row, k-range [r H + (H - KT), r H + H) of the row ----
    {
      const uint4* wrow = reinterpret_cast<const uint4*>(WTg + (size_t)(ut * 128 + q * 32 + lane) * G + (size_t)r * H + (H - KT));
      for (int kb = ch; kb < KT / 64; kb += EW / 4) {
        uint32_t w[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 v = __ldg(wrow + kb * 8 + i);
          w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
        }
        tmem_st32(t_lane + WS_TM_A + kb * 32, w);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(wtm_bar);
    }
    // ---- gate algebra: a thread owns units [4 i4, 4 i4 + 4) of this rank's quarter for EPT utterances ----
    constexpr int EPT = NB * 8 / ET;
    const int i4 = et & 7;
    const int u0 = ut * 128 + (int)r * 32 + 4 * i4;          // first of my 4 hidden units
    // order of the partial sums: the K-split kernel adds the partial of K quarter (u % 64) / 16 first, then the others ascending
    const uint32_t first = 2u * (r & 1u) + (uint32_t)(i4 >> 2);
    // where the partial sum of K quarter X for my unit quarter lives: my own drain tile, or the slot its sender copied into
    auto tile_of = [&](uint32_t X) { return (X == r) ? (stg + r * TILE) : (red + (X < r ? X : X - 1) * TILE); };
    float dc_rec[EPT][4], c_carry[EPT][4];
#pragma unroll
    for (int e = 0; e < EPT; ++e)
#pragma unroll
      for (int i = 0; i < 4; ++i) dc_rec[e][i] = c_carry[e][i] = 0.f;
    for (int s = 0; s < T; ++s) {
      const int t = p.reverse ? s : T - 1 - s;               // BPTT walks against the forward direction
      const bool has_prev = s < T - 1;
      const int t_prev = p.reverse ? t + 1 : t - 1;
      float dh[EPT][4], ct[EPT][4], cp[EPT][4], g4[EPT][16];
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        const int bl = (et >> 3) + e * (ET / 8);
        const int b = b0 + bl;
        if (b < p.nB) {
          const size_t rowi = (size_t)b * T + t;
          if (s == 0) {
            *reinterpret_cast<float4*>(ct[e]) = __ldg(reinterpret_cast<const float4*>(p.c_seq + rowi * H + u0));
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) ct[e][i] = c_carry[e][i];
          }
          *reinterpret_cast<float4*>(dh[e]) = __ldg(reinterpret_cast<const float4*>(p.dH + rowi * p.lddh + u0));
          if (has_prev) {
            *reinterpret_cast<float4*>(cp[e]) = __ldg(reinterpret_cast<const float4*>(p.c_seq + ((size_t)b * T + t_prev) * H + u0));
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) cp[e][i] = 0.f;
          }
          ldg_nc_v8(p.gates + rowi * G + 4 * u0, &g4[e][0]);
          ldg_nc_v8(p.gates + rowi * G + 4 * u0 + 8, &g4[e][8]);
#pragma unroll
          for (int i = 0; i < 4; ++i) c_carry[e][i] = cp[e][i];
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) dh[e][i] = ct[e][i] = cp[e][i] = 0.f;
#pragma unroll
          for (int j = 0; j < 16; ++j) g4[e][j] = 0.f;
        }
        // everything of the gate algebra that does not depend on dh, while the MMAs of this step are still running:
        // g4 <- (d i, d f, d g, d o) per unit of dc resp. dh, ct <- d(dc)/d(dh), cp <- f (the carry factor of dc)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float gi = g4[e][4 * i], gf = g4[e][4 * i + 1], gg = g4[e][4 * i + 2], go = g4[e][4 * i + 3];
          const float tc = tanh_fast(ct[e][i]);
          g4[e][4 * i] = gg * gi * (1.f - gi);
          g4[e][4 * i + 1] = cp[e][i] * gf * (1.f - gf);
          g4[e][4 * i + 2] = gi * (1.f - gg * gg);
          g4[e][4 * i + 3] = tc * go * (1.f - go);
          ct[e][i] = go * (1.f - tc * tc);
          cp[e][i] = gf;
        }
      }
      if (s > 0) {
        if (et == 0) mbar_expect_tx(red_full, 3 * TILE);       // this step's three foreign partial tiles
        mbar_wait(tfull, (s - 1) & 1);
        if (et == 0) LT_TRACE(6);
        tc_fence_after();
        // drain my quadrant: unit quarter q of the CTA's partial sums -> tile q, [utterance][unit] (lanes = consecutive words)
        {
          const uint32_t dst = stg + (uint32_t)q * TILE + (uint32_t)lane * 4u;
          if constexpr (NC == 16) {
            float d[16];
            tmem_ld16(t_lane + j0, d);
#pragma unroll
            for (int j = 0; j < 16; ++j) st_shared_f32(dst + (uint32_t)(j0 + j) * 128u, d[j]);
          } else {
#pragma unroll
            for (int cc = 0; cc < NC / 32; ++cc) {
              float d[32];
              tmem_ld32(t_lane + j0 + cc * 32, d);
#pragma unroll
              for (int j = 0; j < 32; ++j) st_shared_f32(dst + (uint32_t)(j0 + cc * 32 + j) * 128u, d[j]);
            }
          }
        }
        tc_fence_before();
        fence_async_smem();                                       // my staged rows are visible to the copy engine
        asm volatile("bar.sync 1, %0;" ::"n"(ET) : "memory");
        if (et == 0) {
#pragma unroll
          for (uint32_t qq = 0; qq < KS_CL; ++qq) {
            if (qq == r) continue;
            const uint32_t slot = r < qq ? r : r - 1;            // my index among qq's three senders
            asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             mapa_shared(red + slot * TILE, qq)),
                         "r"(stg + qq * TILE), "r"(TILE), "r"(mapa_shared(red_full, qq))
                         : "memory");
          }
          LT_TRACE(9);
        }
        mbar_wait_cluster(red_full, (s - 1) & 1);
        if (et == 0) LT_TRACE(10);
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
          const uint32_t ro = (uint32_t)((et >> 3) + e * (ET / 8)) * 128u + (uint32_t)i4 * 16u;
          {
            const float4 v = ld_shared_v4(tile_of(first) + ro);
            dh[e][0] += v.x; dh[e][1] += v.y; dh[e][2] += v.z; dh[e][3] += v.w;
          }
#pragma unroll
          for (uint32_t X = 0; X < 4; ++X) {
            if (X == first) continue;
            const float4 v = ld_shared_v4(tile_of(X) + ro);
            dh[e][0] += v.x; dh[e][1] += v.y; dh[e][2] += v.z; dh[e][3] += v.w;
          }
        }
      }
      alignas(32) __nv_bfloat16 gb[EPT][16];
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float dc = fmaf(dh[e][i], ct[e][i], dc_rec[e][i]);
          const float di = dc * g4[e][4 * i];
          const float df = dc * g4[e][4 * i + 1];
          const float dg = dc * g4[e][4 * i + 2];
          const float dO = dh[e][i] * g4[e][4 * i + 3];
          dc_rec[e][i] = dc * cp[e][i];
          g4[e][4 * i] = di; g4[e][4 * i + 1] = df; g4[e][4 * i + 2] = dg; g4[e][4 * i + 3] = dO;
          gb[e][4 * i] = __float2bfloat16_rn(di); gb[e][4 * i + 1] = __float2bfloat16_rn(df);
          gb[e][4 * i + 2] = __float2bfloat16_rn(dg); gb[e][4 * i + 3] = __float2bfloat16_rn(dO);
        }
        const int bl = (et >> 3) + e * (ET / 8);
        stg_v8(p.xbuf + ((size_t)(s & 1) * p.nBpad + b0 + bl) * (size_t)G + 4 * u0, reinterpret_cast<const uint32_t*>(gb[e]));
      }
      if (et == 0) LT_TRACE(7);
      asm volatile("bar.sync 1, %0;" ::"n"(ET) : "memory");
      if (et == 0) {
        LT_TRACE(11);
        red_release_add(counter, 1u);      // cumulative gpu-scope release of the CTA's slices (the consumers run the proxy fence)
        LT_TRACE(8);
      }
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        const int b = b0 + (et >> 3) + e * (ET / 8);
        if (b < p.nB) {
          const size_t rowi = (size_t)b * T + t;
          if (p.dP != nullptr) {
            stg_v8f(p.dP + rowi * G + 4 * u0, &g4[e][0]);
            stg_v8f(p.dP + rowi * G + 4 * u0 + 8, &g4[e][8]);
          }
          if (p.dP16 != nullptr) stg_v8(reinterpret_cast<__nv_bfloat16*>(p.dP16) + rowi * G + 4 * u0, reinterpret_cast<const uint32_t*>(gb[e]));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                      // no CTA leaves while peers may still copy into its tiles
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
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
  int BN, NT, MTmax, K, stages, kbs;
  size_t smem, off_w, off_x, off_cnt, total;
  int chunk;   // utterances per launch
};
static LtPlan lt_plan(int nB, int H, bool bwd) {
  LtPlan pl;
  pl.K = bwd ? 4 * H : H;
  const size_t budget = 227 * 1024;
  if (!bwd) {
    pl.BN = fwd_bn(H);
    pl.NT = 4 * H / pl.BN;
    pl.kbs = 1;
    const size_t w_bytes = (size_t)pl.BN * H * 2;
    const int stages = (int)((budget - 1024 - 256 - w_bytes - LT_OUT_STAGE) / LT_KB);
    pl.stages = std::min(8, std::max(2, stages));
    pl.smem = 1024 + w_bytes + (size_t)pl.stages * LT_KB + LT_OUT_STAGE + 256;
  } else {
    pl.BN = 16;
    pl.NT = H / 16;                                   // CTAs per batch tile: H/64 clusters of 4
    pl.kbs = (H / 64) % 2 == 0 ? 2 : 1;
    const size_t w_bytes = (size_t)KS_UNITS * H * 2;
    const size_t fixed = 1024 + w_bytes + 3 * 128 * 16 * 4 + 256;
    const int stages = (int)((225 * 1024 - fixed) / ((size_t)pl.kbs * LT_KB));
    pl.stages = std::min(8, std::max(2, stages));
    pl.smem = fixed + (size_t)pl.stages * pl.kbs * LT_KB;
  }
  pl.MTmax = std::min(2, std::max(1, num_sms() / pl.NT));      // one release counter line per batch tile, two per launch
  pl.chunk = std::min(nB, pl.MTmax * 128);
  const int MT = ceil_div(pl.chunk, 128);
  pl.off_w = 0;
  pl.off_x = align256((size_t)4 * H * H * 2);
  pl.off_cnt = pl.off_x + align256((size_t)2 * MT * 128 * pl.K * 2);
  const int nchunks = ceil_div(nB, pl.chunk);
  pl.total = pl.off_cnt + align256((size_t)nchunks * 128 * sizeof(unsigned));
  return pl;
}
// weight-stationary forward (lstm_tc_fwd_ws_kernel): NB utterances per batch group, as many groups per launch as fit next to
// the 4H/128 row tiles at one CTA per SM
struct LtPlanWs {
  bool ok;
  int NB, NT, NGmax, KT, kbp, chunk;
  size_t smem, off_x, off_cnt, total;
};
constexpr int WS_MAX_GROUPS = 16;
// A CTA of the weight-stationary kernels allocates all 512 TMEM columns and waits for its whole grid: two of them on one SM would
// deadlock in tcgen05.alloc.  Their register footprint (320 threads x ~150) already keeps a second CTA off the SM; requesting more
// than half of the SM's shared memory makes that a guarantee instead of a property of the compiler's register allocation.
constexpr size_t WS_MIN_SMEM = 120 * 1024;
static LtPlanWs lt_plan_ws(int nB, int H) {
  LtPlanWs pl{};
  pl.ok = false;
  if (H % 64 != 0 || H < 128 || H > 1024) return pl;
  pl.NT = 4 * H / 128;
  const int gmax = std::min(WS_MAX_GROUPS, num_sms() / pl.NT);
  if (gmax < 1) return pl;
  pl.NB = ceil_div(nB, 32) <= gmax ? 32 : 64;
  pl.NGmax = gmax;
  pl.chunk = std::min(nB, gmax * pl.NB);
  pl.KT = std::min(H, WS_KT_MAX) / 64 * 64;
  const int kblocks = H / 64;
  // Issuing a TMA instruction costs the producer 120-330 clocks whatever its box (scripts/micro/tma_issue.cu): the step's tile
  // arrives in ONE box, or in two halves where it is long enough for the first half's MMAs to hide under the second's flight
  // (H = 1024 / NB = 64, per step: 4 boxes 5.52 us, 2 boxes 5.25, 1 box 5.29; 16 boxes 7.3)
  pl.kbp = (kblocks >= 12 && kblocks % 2 == 0) ? kblocks / 2 : kblocks;
  pl.smem = 1024 + (size_t)kblocks * pl.NB * 128 + (size_t)(H - pl.KT) / 64 * 16384 + (size_t)pl.NB * (64 + 64 + 128 + 128) + 256;
  pl.smem = std::max(pl.smem, WS_MIN_SMEM);
  if (pl.smem > 227 * 1024) return pl;
  const int NG = ceil_div(pl.chunk, pl.NB);
  pl.off_x = align256((size_t)4 * H * H * 2);
  pl.off_cnt = pl.off_x + align256((size_t)2 * NG * pl.NB * H * 2);
  const int nchunks = ceil_div(nB, pl.chunk);
  pl.total = pl.off_cnt + align256((size_t)nchunks * WS_MAX_GROUPS * 64 * sizeof(unsigned));
  pl.ok = true;
  return pl;
}
// weight-stationary BPTT (lstm_tc_bwd_ws_kernel): clusters of 4 K ranks x H/128 unit tiles per batch group
static LtPlanWs lt_plan_ws_bwd(int nB, int H) {
  LtPlanWs pl{};
  pl.ok = false;
  if (H % 128 != 0 || H < 128 || H > 1024) return pl;
  pl.NT = (H / 128) * KS_CL;                              // CTAs per batch group
  const int gmax = std::min(WS_MAX_GROUPS, num_sms() / pl.NT);
  if (gmax < 1) return pl;
  pl.NB = ceil_div(nB, 32) <= gmax ? 32 : 64;
  pl.NGmax = gmax;
  pl.chunk = std::min(nB, gmax * pl.NB);
  pl.KT = std::min(H, WS_KT_MAX) / 64 * 64;
  const int kblocks = H / 64;
  // Issuing a TMA instruction costs the producer 120-330 clocks whatever its box (scripts/micro/tma_issue.cu): the step's tile
  // arrives in ONE box, or in two halves where it is long enough for the first half's MMAs to hide under the second's flight
  // (H = 1024 / NB = 64, per step: 4 boxes 5.52 us, 2 boxes 5.25, 1 box 5.29; 16 boxes 7.3)
  pl.kbp = (kblocks >= 12 && kblocks % 2 == 0) ? kblocks / 2 : kblocks;
  pl.smem = 1024 + (size_t)kblocks * pl.NB * 128 + (size_t)(H - pl.KT) / 64 * 16384 + (size_t)7 * pl.NB * 128 + 256;
  pl.smem = std::max(pl.smem, WS_MIN_SMEM);
  if (pl.smem > 227 * 1024) return pl;
  const int NG = ceil_div(pl.chunk, pl.NB);
  pl.off_x = align256((size_t)4 * H * H * 2);
  pl.off_cnt = pl.off_x + align256((size_t)2 * NG * pl.NB * 4 * H * 2);
  const int nchunks = ceil_div(nB, pl.chunk);
  pl.total = pl.off_cnt + align256((size_t)nchunks * WS_MAX_GROUPS * 64 * sizeof(unsigned));
  pl.ok = true;
  return pl;
}
// AVC_LSTM_BWD_WS=0 keeps BPTT on the K-split kernel with the W slice in shared memory (same rules as AVC_LSTM_FWD_WS)
static bool bwd_ws_enabled() {
  const char* e = getenv("AVC_LSTM_BWD_WS");
  return !(e && e[0] == '0');
}
// AVC_LSTM_FWD_WS=0 keeps the forward recurrence on the ring kernel (read per call so one process can compare both; the ring
// kernel stays the path for shapes the weight-stationary kernel does not take -- tests/test_gpu_lstm_tc.py runs both settings)
static bool fwd_ws_enabled() {
  const char* e = getenv("AVC_LSTM_FWD_WS");
  return !(e && e[0] == '0');
}

size_t lstm_tc_workspace(int nB, int T, int H, bool bwd) {
  (void)T;
  size_t n = lt_plan(nB, H, bwd).total;
  const LtPlanWs pw = bwd ? lt_plan_ws_bwd(nB, H) : lt_plan_ws(nB, H);
  if (pw.ok) n = std::max(n, pw.total);
  return n;
}

// K-split BPTT launch: clusters of 4 (PAIRS = 1) or 8 (PAIRS = 2) along x, cooperative (the CTAs spin on each other's counters)
template <int PAIRS>
static int lt_launch_ks_t(const CUtensorMap& mW, const CUtensorMap& mX, const LstmTcParams& p, int H, size_t smem, cudaStream_t st) {
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
  if (PAIRS > 1) {      // all clusters must be co-resident
    int max_clusters = 0;
    cfg.numAttrs = 1;
    const cudaError_t oe = cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg);
    if (oe != cudaSuccess || max_clusters * KS_CL * PAIRS < (int)cfg.gridDim.x) {
      (void)cudaGetLastError();
      return AVC_ERR_UNSUPPORTED;
    }
    cfg.numAttrs = 2;
  }
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, mW, mX, p);
  if (e != cudaSuccess && PAIRS > 1) {
    (void)cudaGetLastError();
    return AVC_ERR_UNSUPPORTED;         // the caller retries with clusters of 4
  }
  AVC_CUDA(e);
  g_launches.fetch_add(1);
  return AVC_OK;
}

// clusters of 8 = two unit tiles sharing the dG stream by multicast where they fit, else clusters of 4
static int lt_launch_ks(const CUtensorMap& mW, const CUtensorMap& mX, const LstmTcParams& p, int H, size_t smem, cudaStream_t st) {
  // 16 clusters of 8 (H = 1024 at two batch tiles) do not fit the B200's GPCs at one CTA per SM (occupancy query: 15); smaller
  // grids do.  Remember per grid size whether the launch was refused.
  static bool refused[64] = {};
  const int nst = (H / 64) / p.kbs;
  const int gidx = std::min(63, p.MT * (H / KS_UNITS) / 2);
  if ((H / KS_UNITS) % 2 == 0 && nst % 2 == 0 && !refused[gidx]) {
    const int rc = lt_launch_ks_t<2>(mW, mX, p, H, smem, st);
    if (rc != AVC_ERR_UNSUPPORTED) return rc;
    refused[gidx] = true;
  }
  return lt_launch_ks_t<1>(mW, mX, p, H, smem, st);
}

template <int BN, int CL>
static int lt_launch_fwd(const CUtensorMap& mW, const CUtensorMap& mX, const LtOutMaps& om, const LstmTcParams& p, const LtPlan& pl,
                         cudaStream_t st) {
  auto kern = lstm_tc_fwd_kernel<BN, CL>;
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
  }
  // the CTAs spin on each other's global counters: the cooperative attribute makes the runtime hold the launch until the
  // whole grid can be resident next to whatever else runs (NCCL kernels, side-stream GEMMs)
  attrs[na].id = cudaLaunchAttributeCooperative;
  attrs[na].val.cooperative = 1;
  ++na;
  cfg.attrs = attrs;
  cfg.numAttrs = na;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, mW, mX, om, p);
  if (e != cudaSuccess && CL > 1) {          // refused: the caller retries without clusters
    (void)cudaGetLastError();
    return AVC_ERR_UNSUPPORTED;
  }
  AVC_CUDA(e);
  g_launches.fetch_add(1);
  return AVC_OK;
}

template <int NB, int EW>
static int lt_launch_fwd_ws(const CUtensorMap& mW, const CUtensorMap& mX, const LstmTcParams& p, size_t smem, const __nv_bfloat16* Wb,
                            cudaStream_t st) {
  auto kern = lstm_tc_fwd_ws_kernel<NB, EW>;
  AVC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(p.MT * p.NT);
  cfg.blockDim = dim3(64 + EW * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeCooperative;      // the CTAs of a batch group spin on each other's release counter
  attrs[0].val.cooperative = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, kern, mW, mX, p, Wb) != cudaSuccess) {      // refused (co-residency): the caller falls back to the ring kernel
    (void)cudaGetLastError();
    return AVC_ERR_UNSUPPORTED;
  }
  g_launches.fetch_add(1);
  return AVC_OK;
}

// forward through the weight-stationary kernel; ws laid out by lt_plan_ws
static int lstm_seq_fwd_ws(const LtPlanWs& pl, const __nv_bfloat16* Wb, const float* P, float* h_seq, int ldh, float* gates, float* c_seq,
                           int nB, int T, int H, int reverse, uint8_t* w8, cudaStream_t st, void* aux16, int fmt16, void* aux16b) {
  __nv_bfloat16* xbuf = (__nv_bfloat16*)(w8 + pl.off_x);
  unsigned* counters = (unsigned*)(w8 + pl.off_cnt);
  const int nchunks = ceil_div(nB, pl.chunk);
  AVC_CUDA(cudaMemsetAsync(counters, 0, (size_t)nchunks * WS_MAX_GROUPS * 64 * sizeof(unsigned), st));
  CUtensorMap mW, mX;
  int rc = make_map3(&mW, Wb, H, 4 * H, 1, H, (uint64_t)4 * H * H, 64, 128);
  if (rc) return rc;
  const size_t G = 4 * (size_t)H;
  for (int ch = 0; ch < nchunks; ++ch) {
    const int b0 = ch * pl.chunk;
    const int nb = std::min(pl.chunk, nB - b0);
    LstmTcParams p{};
    p.nB = nb; p.T = T; p.H = H; p.K = H; p.reverse = reverse;
    p.MT = ceil_div(nb, pl.NB); p.NT = pl.NT; p.nBpad = p.MT * pl.NB; p.kbs = pl.kbp; p.kt = pl.KT;
    p.P = P + (size_t)b0 * T * G;
    p.h_seq = h_seq + (size_t)b0 * T * ldh;
    p.ldh = ldh;
    p.gates = gates ? gates + (size_t)b0 * T * G : nullptr;
    p.c_seq = c_seq ? c_seq + (size_t)b0 * T * H : nullptr;
    p.xbuf = xbuf;
    p.counters = counters + ch * WS_MAX_GROUPS * 64;
    p.trace = (ch == 0) ? g_trace : nullptr;
    p.fmt16 = fmt16;
    p.h16 = aux16 ? (void*)((uint16_t*)aux16 + (size_t)b0 * T * H) : nullptr;
    p.h16b = (aux16 && aux16b) ? (void*)((uint16_t*)aux16b + (size_t)b0 * T * H) : nullptr;
    rc = make_map4_grouped(&mX, xbuf, H, (uint64_t)2 * p.nBpad, 1, H, 64, pl.NB, pl.kbp, 2, false);
    if (rc) return rc;
    rc = pl.NB == 32 ? lt_launch_fwd_ws<32, 8>(mW, mX, p, pl.smem, Wb, st) : lt_launch_fwd_ws<64, 8>(mW, mX, p, pl.smem, Wb, st);
    if (rc) return rc;
  }
  return AVC_OK;
}

template <int NB, int EW>
static int lt_launch_bwd_ws(const CUtensorMap& mW, const CUtensorMap& mX, const LstmTcParams& p, size_t smem, const __nv_bfloat16* Wb,
                            cudaStream_t st) {
  auto kern = lstm_tc_bwd_ws_kernel<NB, EW>;
  AVC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(p.MT * p.NT);
  cfg.blockDim = dim3(64 + EW * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = KS_CL;
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  attrs[1].id = cudaLaunchAttributeCooperative;
  attrs[1].val.cooperative = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 2;
  if (cudaLaunchKernelEx(&cfg, kern, mW, mX, p, Wb) != cudaSuccess) {      // refused (cluster placement / co-residency): the caller falls back to the K-split kernel
    (void)cudaGetLastError();
    return AVC_ERR_UNSUPPORTED;
  }
  g_launches.fetch_add(1);
  return AVC_OK;
}

// BPTT through the weight-stationary kernel; ws laid out by lt_plan_ws_bwd.  Wb = Whh_pT (H, 4H) bf16
static int lstm_seq_bwd_ws(const LtPlanWs& pl, const __nv_bfloat16* Wb, float* gates, float* c_seq, const float* dH, int lddh, float* dP,
                           int nB, int T, int H, int reverse, uint8_t* w8, cudaStream_t st, void* aux16) {
  __nv_bfloat16* xbuf = (__nv_bfloat16*)(w8 + pl.off_x);
  unsigned* counters = (unsigned*)(w8 + pl.off_cnt);
  const int nchunks = ceil_div(nB, pl.chunk);
  AVC_CUDA(cudaMemsetAsync(counters, 0, (size_t)nchunks * WS_MAX_GROUPS * 64 * sizeof(unsigned), st));
  const size_t G = 4 * (size_t)H;
  CUtensorMap mW, mX;
  int rc = make_map3(&mW, Wb, G, H, 1, G, (uint64_t)H * G, 64, 128);
  if (rc) return rc;
  for (int ch = 0; ch < nchunks; ++ch) {
    const int b0 = ch * pl.chunk;
    const int nb = std::min(pl.chunk, nB - b0);
    LstmTcParams p{};
    p.nB = nb; p.T = T; p.H = H; p.K = 4 * H; p.reverse = reverse;
    p.MT = ceil_div(nb, pl.NB); p.NT = pl.NT; p.nBpad = p.MT * pl.NB; p.kbs = pl.kbp; p.kt = pl.KT;
    p.gates = gates + (size_t)b0 * T * G;
    p.c_seq = c_seq + (size_t)b0 * T * H;
    p.dH = dH + (size_t)b0 * T * lddh;
    p.lddh = lddh;
    p.dP = dP ? dP + (size_t)b0 * T * G : nullptr;
    p.dP16 = aux16 ? (void*)((uint16_t*)aux16 + (size_t)b0 * T * G) : nullptr;
    p.xbuf = xbuf;
    p.counters = counters + ch * WS_MAX_GROUPS * 64;
    p.trace = (ch == 0) ? g_trace : nullptr;
    p.wide = 1;
    rc = make_map4_grouped(&mX, xbuf, G, (uint64_t)2 * p.nBpad, 1, G, 64, pl.NB, pl.kbp, 2, false);
    if (rc) return rc;
    rc = pl.NB == 32 ? lt_launch_bwd_ws<32, 8>(mW, mX, p, pl.smem, Wb, st) : lt_launch_bwd_ws<64, 8>(mW, mX, p, pl.smem, Wb, st);
    if (rc) return rc;
  }
  return AVC_OK;
}

// W: fwd -> Whh_p (4H, H);  bwd -> Whh_pT (H, 4H)  (packed/interleaved)
// w_fmt 0: W is fp32 and converted to bf16 here; 1: W is already bf16 (avc_pack_lstm_weight_h) and read in place
int lstm_seq_tc(bool bwd, const void* Wv, const float* P, float* h_seq, int ldh, float* gates, float* c_seq,
                const float* dH, int lddh, float* dP, int nB, int T, int H, int reverse, void* ws, size_t ws_bytes,
                cudaStream_t st, void* aux16, int fmt16, int w_fmt, void* aux16b) {
  const LtPlan pl = lt_plan(nB, H, bwd);
  if (!ws || ws_bytes < pl.total) {
    set_error("avc_lstm_seq_%s(bf16): workspace %zu < %zu", bwd ? "bwd" : "fwd", ws_bytes, pl.total);
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
  if (!bwd && fwd_ws_enabled()) {
    const LtPlanWs pw = lt_plan_ws(nB, H);
    const bool al16 = ((((uintptr_t)P | (uintptr_t)gates | (uintptr_t)c_seq | (uintptr_t)h_seq | (uintptr_t)aux16 | (uintptr_t)aux16b) & 15) == 0) &&
                      ldh % 4 == 0;
    if (pw.ok && al16 && ws_bytes >= pw.total) {
      const int rc = lstm_seq_fwd_ws(pw, Wb, P, h_seq, ldh, gates, c_seq, nB, T, H, reverse, w8, st, aux16, fmt16, aux16b);
      if (rc != AVC_ERR_UNSUPPORTED) return rc;      // a refused launch: the ring kernel below recomputes the whole call
    }
  }
  if (bwd && bwd_ws_enabled()) {
    const LtPlanWs pw = lt_plan_ws_bwd(nB, H);
    auto al32 = [](const void* q) { return ((uintptr_t)q & 31) == 0; };
    const bool wide = al32(gates) && al32(c_seq) && al32(dH) && al32(dP) && al32(aux16) && lddh % 8 == 0;
    if (pw.ok && wide && ws_bytes >= pw.total) {
      const int rc = lstm_seq_bwd_ws(pw, Wb, gates, c_seq, dH, lddh, dP, nB, T, H, reverse, w8, st, aux16);
      if (rc != AVC_ERR_UNSUPPORTED) return rc;      // a refused launch: the K-split kernel below recomputes the whole call
    }
  }
  const int nchunks = ceil_div(nB, pl.chunk);
  AVC_CUDA(cudaMemsetAsync(counters, 0, (size_t)nchunks * 128 * sizeof(unsigned), st));
  CUtensorMap mW, mX;
  const int w_rows = bwd ? H : 4 * H;
  int rc = make_map3(&mW, Wb, pl.K, w_rows, 1, pl.K, (uint64_t)w_rows * pl.K, 64, bwd ? KS_UNITS : pl.BN);
  if (rc) return rc;
  const size_t G = 4 * (size_t)H;
  for (int ch = 0; ch < nchunks; ++ch) {
    const int b0 = ch * pl.chunk;
    const int nb = std::min(pl.chunk, nB - b0);
    LstmTcParams p{};
    p.nB = nb; p.T = T; p.H = H; p.K = pl.K; p.reverse = reverse;
    p.MT = ceil_div(nb, 128); p.NT = pl.NT; p.nBpad = p.MT * 128; p.stages = pl.stages; p.kbs = pl.kbs;
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
    p.trace = (ch == 0) ? g_trace : nullptr;
    p.fmt16 = fmt16;
    {
      auto al32 = [](const void* q) { return ((uintptr_t)q & 31) == 0; };
      p.wide = al32(p.P) && al32(p.gates) && al32(p.c_seq) && al32(p.dH) && al32(p.dP) && (lddh % 8 == 0) && (H % 8 == 0);
    }
    p.h16 = (!bwd && aux16) ? (void*)((uint16_t*)aux16 + (size_t)b0 * T * H) : nullptr;
    p.dP16 = (bwd && aux16) ? (void*)((uint16_t*)aux16 + (size_t)b0 * T * G) : nullptr;
    p.h16b = (!bwd && aux16 && aux16b) ? (void*)((uint16_t*)aux16b + (size_t)b0 * T * H) : nullptr;
    if (bwd) {
      // activation tile: 4-D box of kbs 64-column groups x 128 rows (one instruction per ring stage)
      rc = make_map4_grouped(&mX, xbuf, pl.K, (uint64_t)2 * p.nBpad, 1, pl.K, 64, 128, pl.kbs, 2, false);
      if (rc) return rc;
      rc = lt_launch_ks(mW, mX, p, H, pl.smem, st);
      if (rc) return rc;
      continue;
    }
    // forward: the saved tensors leave through TMA stores (16-byte aligned tensors; ldh % 4 == 0 is an API precondition)
    if ((((uintptr_t)p.gates | (uintptr_t)p.c_seq | (uintptr_t)p.h_seq | (uintptr_t)p.h16 | (uintptr_t)p.h16b) & 15) != 0 || ldh % 4 != 0) {
      set_error("avc_lstm_seq_fwd(bf16): h_seq / gates / c_seq / h16 must be 16-byte aligned and ldh a multiple of 4");
      return AVC_ERR_INVALID;
    }
    rc = make_map3(&mX, xbuf, pl.K, (uint64_t)2 * p.nBpad, 1, pl.K, (uint64_t)2 * p.nBpad * pl.K, 64, 128);
    if (rc) return rc;
    LtOutMaps om{};
    om.gates = om.c = om.h = om.h16 = om.h16b = mW;
    {
      const int U = pl.BN / 4;
      const uint64_t Tn = (uint64_t)T;
      if (p.gates) rc = make_map3_store(&om.gates, p.gates, 4, G, Tn, nb, G, Tn * G, 32, 1, 32, true);
      if (!rc && p.c_seq) rc = make_map3_store(&om.c, p.c_seq, 4, H, Tn, nb, H, Tn * H, U, 1, 32, false);
      if (!rc) rc = make_map3_store(&om.h, p.h_seq, 4, H, Tn, nb, ldh, Tn * ldh, U, 1, 32, false);
      if (!rc && p.h16) rc = make_map3_store(&om.h16, p.h16, 2, H, Tn, nb, H, Tn * H, U, 1, 32, false);
      if (!rc && p.h16b) rc = make_map3_store(&om.h16b, p.h16b, 2, H, Tn, nb, H, Tn * H, U, 1, 32, false);
      if (rc) return rc;
    }
    // clusters of 4 column tiles (NT is a multiple of 4 for every supported H); without clusters if the launch is refused
    rc = pl.BN == 64 ? lt_launch_fwd<64, 4>(mW, mX, om, p, pl, st) : lt_launch_fwd<32, 4>(mW, mX, om, p, pl, st);
    if (rc == AVC_ERR_UNSUPPORTED)
      rc = pl.BN == 64 ? lt_launch_fwd<64, 1>(mW, mX, om, p, pl, st) : lt_launch_fwd<32, 1>(mW, mX, om, p, pl, st);
    if (rc) return rc;
  }
  return AVC_OK;
}

}  // namespace avc
