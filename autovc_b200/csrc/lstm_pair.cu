// Persistent LSTM recurrences on CTA pairs (tcgen05 cta_group::2, M = 128): ONE resident copy of W_hh serves both
// 128-utterance batch tiles of a step, on 64 SMs.
//
// nn.LSTM (model_vc_mel.py:90/:111 lstm1, :104/:118 lstm2): per step  gates = P_t + h_{t-1} W_hh^T  (forward) and
// dh_t = dH_t + dG_{t+1} W_hh (BPTT).  lstm_tc.cu gives every 128-row batch tile its own set of CTAs, each with a resident
// W_hh slice and the WHOLE 128 x K activation tile streamed through its ring every step (256 KB per SM and step at
// H = 1024): the r01 traces showed a step to be that ingest (~5 of 7.8 us) plus a latency chain (publish -> detect ->
// epilogue), with two copies of W_hh resident on 128 SMs.
//
// Here a CTA PAIR shares one W_hh slice through cta_group::2: the pair's MMA has M = 128, each CTA supplies ITS 64
// utterances of the activation tile (A) and ITS half of the slice's gate columns (B), and holds the accumulator rows of its
// 64 utterances for all of the pair's columns.  Per SM and batch tile the ingest is 64 x K instead of 128 x K, and the
// freed half of the machine is not needed: the SAME pair runs the second batch tile of the step through a second TMEM
// accumulator, its own epilogue warps and its own release counters, so that one tile's latency chain (publish, detection,
// epilogue) hides under the other tile's loads and MMAs.  A recurrence occupies 64 SMs; the other 84 run the GEMMs of
// the neighbouring layers (weight gradients on the side stream).
//
//   forward : pair owns NP gate columns (NP/4 hidden units, gate-interleaved rows of W_hh), K = H
//   TMEM    : D[64 rows, NP] of a CTA sits in NP/2 columns: lanes 0-63 hold columns [0, NP/2), lanes 64-127 columns
//             [NP/2, NP) (the cta_group::2, M = 128 accumulator layout) -> epilogue warp q covers rows (q&1)*32.. and
//             gate columns (q>>1)*NP/2..
//   exchange: h_t in a bf16 double buffer [2][nBpad][H]; one release counter per (batch tile, row half): a CTA only waits
//             for the CTAs that publish ITS 64 rows (the same-rank CTAs of all pairs)
//   ring    : a stage holds KBS = 4 consecutive k-blocks of this CTA's 64 rows (32 KB), fetched by ONE 4-D TMA box.  A TMA tensor
//             load costs the issuing thread ~120-330 SM clocks whatever its size (scripts/micro/tma_issue.cu: 122 clocks per
//             instruction from 2 KB to 32 KB boxes; 320 in this kernel's producer loop), so 8 KB boxes bounded the first version
//             at 16 x 320 clocks = 2.7 us of issue per tile and step, while 32 KB boxes land 280 clocks apart (117 B/clk per SM).
//             A multicast variant (clusters of 4 pairs, 2 KB slices) was measured slower for the same reason.
//   proxies : the publishers write h_t with generic-proxy stores and release a counter; the consumer acquires the counter and
//             executes fence.proxy.async BEFORE its TMA (async-proxy) loads.  With the proxy fence on the writer side (r01) the
//             publish cost two serial drains of the CTA's outstanding stores (fence.proxy.async ~0.5-0.9 us, then
//             red.release.gpu ~0.5-1.0 us, SM-clock trace); on the consumer side it has nothing to wait for.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include <cuda_fp16.h>

#include "tc_common.cuh"

namespace avc {

constexpr int LP_THREADS = 320;            // warp 0 TMA producer, warp 1 MMA issuer / TMEM owner, warps 2-5 epilogue of batch tile 0, 6-9 of tile 1
constexpr int LP_STAGE = 64 * 64 * 2;      // one activation k-block of a CTA: 64 utterances x 64 bf16 (8 KB)
constexpr int LP_OUT_STAGE = 8 * 4096;     // 8 epilogue warps x 4 KB staging for the TMA stores
constexpr int LP_CNT_STRIDE = 32;          // words between release counters (one 128-byte line each)

struct LpParams {
  int nB, T, H, K, reverse;
  int MT;                 // batch tiles of 128 utterances every pair walks (1 or 2)
  int npairs, nBpad, stages;
  int kbs;                // k-blocks (64 columns of K) per ring stage
  const float* P;         // fwd: (nB,T,4H) pre-activations
  float* h_seq;           // fwd: out (nB,T,H), row stride ldh
  int ldh;
  float* gates;           // fwd: out / bwd: in (nB,T,4H) activated gates (NULL with c_seq: nothing saved for BPTT)
  float* c_seq;           // fwd: out / bwd: in (nB,T,H)
  const float* dH;        // bwd: (nB,T,H), row stride lddh
  int lddh;
  void* dP16;             // bwd: out (nB,T,4H) bf16 gate gradient
  float* dP;              // bwd: optional fp32 copy of the gate gradient
  __nv_bfloat16* xbuf;    // exchange buffer [2][nBpad][K]
  unsigned* counters;     // [MT][2] release counters, LP_CNT_STRIDE words apart, zero-initialised
  void* h16;              // fwd: 16-bit copy of h_seq (format fmt16: 1 bf16, 2 fp16) or NULL
  void* h16b;             // fwd: bf16 copy of h_seq or NULL
  int fmt16;
  int wide;               // row-per-thread tensors are 32-byte aligned: 256-bit global accesses
  int exp_mode;           // development switch (AVC_LP_EXP), 0 in production
  unsigned long long* trace;   // optional per-step stamps of pair 0 / leader (avc_debug_set_trace)
};

struct alignas(64) LpOutMaps {
  CUtensorMap gates;   // (G, T, nB) fp32, box (32, 1, 32), 128B swizzle
  CUtensorMap c;       // (H, T, nB) fp32, box (U, 1, 32)
  CUtensorMap h;       // (H, T, nB; row stride ldh) fp32, box (U, 1, 32)
  CUtensorMap h16;     // (H, T, nB) 16-bit, box (U, 1, 32)
  CUtensorMap h16b;    // (H, T, nB) bf16, box (U, 1, 32)
};

namespace {

__device__ __forceinline__ void lp_ldg_nc_v8(const float* p, float* v) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void lp_stg_v8(void* p, const uint32_t* w) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]),
               "r"(w[5]), "r"(w[6]), "r"(w[7])
               : "memory");
}
__device__ __forceinline__ float lp_tanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lp_sigmoid(float x) { return fmaf(0.5f, lp_tanh(0.5f * x), 0.5f); }
__device__ __forceinline__ unsigned lp_ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void lp_red_release_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void lp_fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void lp_fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ unsigned long long lp_gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void lp_named_bar(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

// per-step stamps of pair 0 / leader / batch tile 0: slot 0 counter seen, 1 last load issued, 2 first k-block landed,
// 3 last k-block landed (MMAs issued), 4 epilogue woke, 5 math done, 6 published
__device__ __forceinline__ unsigned long long lp_clock() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));
  return t;
}
// SM-clock stamps of pair 0 / leader / tile 0: slots 8.. of the step's row; and per-k-block issue / landing clocks of step 64
#define LP_CLK(slot)                                                                                     \
  do {                                                                                                   \
    if (p.trace != nullptr && blockIdx.x == 0) p.trace[(size_t)s * 16 + (slot)] = lp_clock();            \
  } while (0)
#define LP_CLK_KB(off, kb)                                                                               \
  do {                                                                                                   \
    if (p.trace != nullptr && blockIdx.x == 0 && s == 64 && m == 0) p.trace[(size_t)16 * T + (off) + (kb)] = lp_clock(); \
  } while (0)
#define LP_TRACE(slot)                                                                                   \
  do {                                                                                                   \
    if (p.trace != nullptr && blockIdx.x == 0) p.trace[(size_t)s * 16 + (slot)] = lp_gtime();            \
  } while (0)

}  // namespace

// NP = gate columns per pair (128: H = 1024; 64: smaller H, so that the recurrence still spreads over >= 32 pairs)
template <int NP>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(LP_THREADS, 1)
lstm_pair_fwd_kernel(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapX,
                     const __grid_constant__ LpOutMaps om, const LpParams p) {
  constexpr int BN = NP / 2;                 // gate columns per epilogue thread ( = TMEM columns per accumulator)
  constexpr int U = NP / 8;                  // hidden units per epilogue thread
  constexpr int TM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const int kblocks = p.K / 64;
  constexpr uint32_t w_block = BN * 128;                    // this CTA's half of the slice: BN rows x 64 bf16 per k-block
  const uint32_t w_base = base;
  const uint32_t ring = base + kblocks * w_block;
  const uint32_t out_stage = ring + p.stages * p.kbs * LP_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(gen + kblocks * w_block + p.stages * p.kbs * LP_STAGE + LP_OUT_STAGE);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (8 + s); };
  const uint32_t w_bar = bar0 + 8u * 16;
  auto tfull_bar = [&](int m) { return bar0 + 8u * (17 + m); };
  auto tempty_bar = [&](int m) { return bar0 + 8u * (19 + m); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();           // 0 = leader (issues the MMAs); also the row half this CTA owns
  const int pair = blockIdx.x >> 1;
  const int kbs = p.kbs;                             // k-blocks per ring stage
  const uint32_t stage_bytes = (uint32_t)kbs * LP_STAGE;
  const int nst = kblocks / kbs;                     // stages per tile and step
  const int T = p.T, H = p.H, G = 4 * p.H;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapX) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(w_bar, 1);
    for (int m = 0; m < 2; ++m) {
      mbar_init(tfull_bar(m), 1);
      mbar_init(tempty_bar(m), 8);         // 4 epilogue warps of each CTA of the pair
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc_2cta(smem_u32(tmem_slot), TM_COLS);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                      // both CTAs' barriers and TMEM exist before anyone signals
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (whole warp loops, one elected lane issues) =====================
    if (elect_one()) {
      if (rank == 0) mbar_expect_tx(w_bar, 2u * kblocks * w_block);
      const uint32_t wb = mapa_u32(w_bar, 0);
      for (int kb = 0; kb < kblocks; ++kb)
        tma_load_3d_2cta(w_base + kb * w_block, &mapW, wb, kb * 64, pair * NP + (int)rank * BN, 0);
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int s = 1; s < T; ++s) {
      for (int m = 0; m < p.MT; ++m) {
        const int row0 = ((s - 1) & 1) * p.nBpad + m * 128 + (int)rank * 64;
        const unsigned* cnt = p.counters + (m * 2 + (int)rank) * LP_CNT_STRIDE;
        const unsigned target = (unsigned)s * (unsigned)p.npairs;       // every pair has published step s-1 of my 64 rows
        if (lane == 0) {
          while (lp_ld_acquire(cnt) < target) {
          }
          if (m == 0) LP_TRACE(0);
        }
        __syncwarp();
        lp_fence_proxy_async_global();   // the published rows were written through the generic proxy: order my TMA reads after them
        for (int st = 0; st < nst; ++st) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          if (elect_one()) {
            if (rank == 0) mbar_expect_tx(full_bar(stage), 2u * stage_bytes);
            tma_load_4d_2cta(ring + stage * stage_bytes, &mapX, mapa_u32(full_bar(stage), 0), 0, row0, st * kbs, 0);
            LP_CLK_KB(0, st);
            if (m == 0 && st == nst - 1) LP_TRACE(1);
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
    // tail: every commit aimed at this CTA's empty barriers has landed before the CTA may exit
    for (int s = 0; s < p.stages; ++s) {
      mbar_wait(empty_bar(stage), phase ^ 1);
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc(128, NP, 0, 0);
      mbar_wait(w_bar, 0);
      tc_fence_after();
      int stage = 0;
      uint32_t phase = 0;
      for (int s = 1; s < T; ++s) {
        for (int m = 0; m < p.MT; ++m) {
          mbar_wait(tempty_bar(m), ((s - 1) & 1) ^ 1);       // both CTAs' epilogues have drained this accumulator
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + m * BN;
          for (int st = 0; st < nst; ++st) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t sa0 = ring + stage * stage_bytes, sb0 = w_base + (uint32_t)(st * kbs) * w_block;
            if (elect_one()) {
              LP_CLK_KB(64, st);
              if (m == 0 && st == 0) LP_TRACE(2);
              for (int kbi = 0; kbi < kbs; ++kbi) {
                const uint32_t sa = sa0 + kbi * LP_STAGE, sb = sb0 + kbi * w_block;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_2cta<2>(d_tmem, make_desc(sa + k * 32, 16, 1024), make_desc(sb + k * 32, 16, 1024), idesc,
                               (st > 0 || kbi > 0 || k > 0) ? 1u : 0u);
              }
              umma_commit_2cta(empty_bar(stage), 3);
              if (st == nst - 1) {
                umma_commit_2cta(tfull_bar(m), 3);
                if (m == 0) LP_TRACE(3);
              }
            }
            __syncwarp();
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if ((warp - 2) / 4 < p.MT) {
    // ===================== epilogue warps of batch tile m =====================
    const int m = (warp - 2) >> 2;
    const int q = warp & 3;                               // TMEM lane quadrant
    const int rl = (q & 1) * 32 + lane;                   // row within this CTA's 64 utterances
    const int ch = q >> 1;                                // which half of the pair's gate columns
    const int b = m * 128 + (int)rank * 64 + rl;
    const bool live = b < p.nB;
    const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(m * BN);
    const uint32_t tempty_leader = mapa_u32(tempty_bar(m), 0);
    const int n0 = pair * NP + ch * BN, u0 = n0 >> 2;
    unsigned* counter = p.counters + (m * 2 + (int)rank) * LP_CNT_STRIDE;
    const uint32_t buf = out_stage + (uint32_t)(warp - 2) * 4096u;
    const int brow = m * 128 + (int)rank * 64 + (q & 1) * 32;
    const bool save_bptt = p.gates != nullptr;            // inference (no backward): only h leaves the kernel
    const bool lead_thread = ((warp - 2) & 3) == 0 && lane == 0;
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
          for (int j = 0; j < BN; j += 8) lp_ldg_nc_v8(p.P + rowi * G + n0 + j, &pre[j]);
        } else {
#pragma unroll
          for (int j = 0; j < BN; j += 4)
            *reinterpret_cast<float4*>(&pre[j]) = __ldg(reinterpret_cast<const float4*>(p.P + rowi * G + n0 + j));
        }
      } else {
#pragma unroll
        for (int j = 0; j < BN; ++j) pre[j] = 0.f;
      }
      if (s > 0) {
        mbar_wait(tfull_bar(m), (s - 1) & 1);
        if (m == 0 && lead_thread) { LP_TRACE(4); LP_CLK(8); }
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
        if (lane == 0) mbar_arrive_remote_relaxed(tempty_leader);   // the accumulator is in registers (tcgen05.wait::ld done)
        if (m == 0 && lead_thread) LP_CLK(9);
      }
      alignas(32) __nv_bfloat16 hb[U];
      float hf[U];
#pragma unroll
      for (int i = 0; i < U; ++i) {
        const float gi = lp_sigmoid(pre[4 * i + 0]);
        const float gf = lp_sigmoid(pre[4 * i + 1]);
        const float gg = lp_tanh(pre[4 * i + 2]);
        const float go = lp_sigmoid(pre[4 * i + 3]);
        c[i] = fmaf(gf, c[i], gi * gg);
        hf[i] = go * lp_tanh(c[i]);
        hb[i] = __float2bfloat16_rn(hf[i]);
        pre[4 * i + 0] = gi; pre[4 * i + 1] = gf; pre[4 * i + 2] = gg; pre[4 * i + 3] = go;
      }
      // publish h_t first: bf16 slice -> CTA-scope barrier of the tile's 4 warps -> one thread: proxy fence + gpu-scope release
      if (live) {
        __nv_bfloat16* xb = p.xbuf + ((size_t)(s & 1) * p.nBpad + b) * p.K + u0;
        if (U == 16) {
          lp_stg_v8(xb, reinterpret_cast<const uint32_t*>(hb));
        } else {
#pragma unroll
          for (int i = 0; i < U; i += 8) *reinterpret_cast<uint4*>(xb + i) = *reinterpret_cast<const uint4*>(&hb[i]);
        }
      }
      if (m == 0 && lead_thread) { LP_TRACE(5); LP_CLK(10); }
      lp_named_bar(1 + m);
      if (lead_thread) {
        if (m == 0) LP_CLK(11);
        if (p.exp_mode & 2) lp_fence_proxy_async();
        if (m == 0) LP_CLK(12);
        lp_red_release_add(counter, 1u);
        if (m == 0) { LP_TRACE(6); LP_CLK(13); }
      }
      lp_named_bar(1 + m);             // the bulk stores below must not queue ahead of that release
      // saved tensors: registers -> this warp's 4 KB staging tile -> TMA store (rows of padded utterances are clipped by the TMA unit)
      if (p.exp_mode & 1) continue;
#pragma unroll
      for (int half = 0; half < BN / 32; ++half) {
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
          if (p.h16b != nullptr) {
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
    if (lane == 0) bulk_wait_all();        // the staged tiles are read out before the CTA's shared memory goes away
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                      // no CTA leaves while its peer may still signal its barriers / read its smem
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, TM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static unsigned long long* g_lp_trace = nullptr;
void lstm_pair_set_trace(unsigned long long* p) { g_lp_trace = p; }

static int lp_np(int H) { return H >= 1024 ? 128 : 64; }

bool lstm_pair_fwd_supported(int H) {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("AVC_LSTM_PAIR");
    on = e ? atoi(e) : 1;
  }
  return on != 0 && H >= 128 && H <= 1024 && H % 64 == 0;
}

struct LpPlan {
  int NP, npairs, stages, kbs, chunk;
  size_t smem, off_x, off_cnt, total;
};
static LpPlan lp_plan_fwd(int nB, int H) {
  LpPlan pl;
  pl.NP = lp_np(H);
  pl.npairs = 4 * H / pl.NP;
  const size_t w_bytes = (size_t)(pl.NP / 2) * H * 2;
  const size_t budget = 227 * 1024;
  const int kblocks = H / 64;
  pl.kbs = kblocks % 4 == 0 ? 4 : kblocks % 3 == 0 ? 3 : kblocks % 2 == 0 ? 2 : 1;
  int stages = (int)((budget - 1024 - 256 - w_bytes - LP_OUT_STAGE) / ((size_t)pl.kbs * LP_STAGE));
  pl.stages = std::min(8, std::max(2, stages));
  pl.smem = 1024 + w_bytes + (size_t)pl.stages * pl.kbs * LP_STAGE + LP_OUT_STAGE + 256;
  pl.chunk = std::min(nB, 256);
  const int MT = ceil_div(pl.chunk, 128);
  pl.off_x = 0;
  pl.off_cnt = align256((size_t)2 * MT * 128 * H * 2);
  const int nchunks = ceil_div(nB, pl.chunk);
  pl.total = pl.off_cnt + align256((size_t)nchunks * 4 * LP_CNT_STRIDE * sizeof(unsigned));
  return pl;
}
size_t lstm_pair_fwd_workspace(int nB, int H) { return lp_plan_fwd(nB, H).total; }

template <int NP>
static int lp_launch_fwd(const CUtensorMap& mW, const CUtensorMap& mX, const LpOutMaps& om, const LpParams& p, size_t smem,
                         cudaStream_t st) {
  auto kern = lstm_pair_fwd_kernel<NP>;
  AVC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * p.npairs);
  cfg.blockDim = dim3(LP_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeCooperative;      // the CTAs spin on each other's counters: all of them must be resident
  attrs[0].val.cooperative = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 1;
  AVC_CUDA(cudaLaunchKernelEx(&cfg, kern, mW, mX, om, p));
  g_launches.fetch_add(1);
  return AVC_OK;
}

// Forward recurrence of one layer-direction.  Wb: W_hh packed gate-interleaved (4H, H) bf16.
int lstm_seq_pair_fwd(const __nv_bfloat16* Wb, const float* P, float* h_seq, int ldh, float* gates, float* c_seq, int nB, int T,
                      int H, int reverse, void* ws, size_t ws_bytes, cudaStream_t st, void* h16, int fmt16, void* h16b) {
  const LpPlan pl = lp_plan_fwd(nB, H);
  if (!ws || ws_bytes < pl.total) {
    set_error("avc_lstm_seq_fwd(pair): workspace %zu < %zu", ws_bytes, pl.total);
    return AVC_ERR_WORKSPACE;
  }
  if (pl.smem > 227 * 1024) {
    set_error("avc_lstm_seq_fwd(pair): H=%d needs %zu bytes of shared memory", H, pl.smem);
    return AVC_ERR_UNSUPPORTED;
  }
  uint8_t* w8 = (uint8_t*)ws;
  __nv_bfloat16* xbuf = (__nv_bfloat16*)(w8 + pl.off_x);
  unsigned* counters = (unsigned*)(w8 + pl.off_cnt);
  const int nchunks = ceil_div(nB, pl.chunk);
  AVC_CUDA(cudaMemsetAsync(counters, 0, (size_t)nchunks * 4 * LP_CNT_STRIDE * sizeof(unsigned), st));
  CUtensorMap mW, mX;
  int rc = make_map3(&mW, Wb, H, (uint64_t)4 * H, 1, H, (uint64_t)4 * H * H, 64, pl.NP / 2);
  if (rc) return rc;
  const size_t G = 4 * (size_t)H;
  for (int ch = 0; ch < nchunks; ++ch) {
    const int b0 = ch * pl.chunk;
    const int nb = std::min(pl.chunk, nB - b0);
    LpParams p{};
    p.nB = nb; p.T = T; p.H = H; p.K = H; p.reverse = reverse;
    p.MT = ceil_div(nb, 128); p.npairs = pl.npairs; p.nBpad = p.MT * 128; p.stages = pl.stages; p.kbs = pl.kbs;
    p.P = P + (size_t)b0 * T * G;
    p.h_seq = h_seq + (size_t)b0 * T * ldh;
    p.ldh = ldh;
    p.gates = gates ? gates + (size_t)b0 * T * G : nullptr;
    p.c_seq = c_seq ? c_seq + (size_t)b0 * T * H : nullptr;
    p.xbuf = xbuf;
    p.counters = counters + ch * 4 * LP_CNT_STRIDE;
    p.fmt16 = fmt16;
    p.h16 = h16 ? (void*)((uint16_t*)h16 + (size_t)b0 * T * H) : nullptr;
    p.h16b = (h16 && h16b) ? (void*)((uint16_t*)h16b + (size_t)b0 * T * H) : nullptr;
    p.trace = ch == 0 ? g_lp_trace : nullptr;
    p.exp_mode = getenv("AVC_LP_EXP") ? atoi(getenv("AVC_LP_EXP")) : 0;
    p.wide = (((uintptr_t)p.P) & 31) == 0 && (H % 8 == 0);
    if ((((uintptr_t)p.gates | (uintptr_t)p.c_seq | (uintptr_t)p.h_seq | (uintptr_t)p.h16 | (uintptr_t)p.h16b) & 15) != 0 || ldh % 4 != 0) {
      set_error("avc_lstm_seq_fwd(pair): output tensors must be 16-byte aligned (ldh %% 4 == 0)");
      return AVC_ERR_INVALID;
    }
    rc = make_map4_grouped(&mX, xbuf, H, (uint64_t)2 * p.nBpad, 1, H, 64, 64, pl.kbs, 2, false);     // box: 64 columns x 64 rows x kbs k-blocks
    if (rc) return rc;
    LpOutMaps om{};
    om.gates = om.c = om.h = om.h16 = om.h16b = mW;
    const int U = pl.NP / 8;
    const uint64_t Tn = (uint64_t)T;
    if (p.gates) rc = make_map3_store(&om.gates, p.gates, 4, G, Tn, nb, G, Tn * G, 32, 1, 32, true);
    if (!rc && p.c_seq) rc = make_map3_store(&om.c, p.c_seq, 4, H, Tn, nb, H, Tn * H, U, 1, 32, false);
    if (!rc) rc = make_map3_store(&om.h, p.h_seq, 4, H, Tn, nb, ldh, Tn * ldh, U, 1, 32, false);
    if (!rc && p.h16) rc = make_map3_store(&om.h16, p.h16, 2, H, Tn, nb, H, Tn * H, U, 1, 32, false);
    if (!rc && p.h16b) rc = make_map3_store(&om.h16b, p.h16b, 2, H, Tn, nb, H, Tn * H, U, 1, 32, false);
    if (rc) return rc;
    rc = pl.NP == 128 ? lp_launch_fwd<128>(mW, mX, om, p, pl.smem, st) : lp_launch_fwd<64>(mW, mX, om, p, pl.smem, st);
    if (rc) return rc;
  }
  return AVC_OK;
}

}  // namespace avc
