// K2: stride-1 3x3 / 1x1 convolution as a tap-shifted GEMM on the 5th-gen tensor cores (sm_100a).
//
//   out[m][n] = act( sum_{tap} sum_{c} in[m + shift(tap)][c] * w[tap][c][n] + bias[n] (+ res[m][n]) )
//
// over the flat padded ("PS") activation matrix of pe_common.cuh: because every image carries a zero
// 1-pixel border, a 3x3/pad-1 convolution is nine GEMMs whose A operand is the SAME matrix shifted by
// (ky-1)*(W+2) + (kx-1) rows.  Replaces the cuDNN conv + BN + ReLU (+ residual) sequences that mmpose's
// HRNet.forward launches (reference call site pose_pipeline/wrappers/mmpose.py:75).
//
// Precision: 3xTF32.  Activations and weights are stored as (hi, lo) TF32 pairs; each logical MAC is
// hi*hi + hi*lo + lo*hi on tcgen05.mma kind::tf32 with FP32 accumulation in TMEM (error ~2^-21, the
// oracle's own fp32 noise level; plain TF32/BF16 cannot hold the 1e-3 px keypoint gate).
//
// One CTA: MT accumulators of 128 rows x NC channels in TMEM.  Per 16-channel chunk, TMA loads ONE halo
// window of (128*MT + 2*(W+3)) activation rows (SWIZZLE_128B, 128 B per row = hi16|lo16) that serves
// all nine taps: each tap's A operand is an UMMA shared-memory descriptor into the same window at a row
// offset (base_offset 0: the swizzle is a function of absolute smem address bits, measured).  Weights stream
// through a second TMA ring, one stage = TPS taps x NC rows x 128 B, and are amortised over MT tiles.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane), warps 2-5 = epilogue
// (tcgen05.ld -> bias / residual / ReLU -> tf32 split -> 128-byte row-chunk stores).
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "kernels.h"
#include "pe_common.cuh"

// layout constants of this build (pe_common.cuh): bytes per 16-channel chunk row, 16-byte units per row, offset of the
// lo / l half, MMA k-steps per chunk
#define CHB PS_CHUNK_BYTES
constexpr uint32_t ROW16 = CHB / 16;          // 8 (tf32) / 4 (fp16)
constexpr uint32_t LO16 = CHB / 32;           // 16-byte units from hi to lo: 4 / 2
constexpr int KSTEPS = PE_FP16 ? 1 : 2;
// warp roles: 0 = activation-window TMA producer, 1 = MMA issuer, 2..9 = epilogue (TMEM lane quarter = warp & 3, two warps per
// quarter splitting the 16-column groups), 10 = weight TMA producer, 11 = second MMA issuer (cross terms)
constexpr int TC_THREADS = 384;
constexpr int EPI_WARPS = 8;       // MMA k-steps per chunk half: 16 x f16 = one K=16 MMA; 16 x tf32 = two K=8 MMAs

struct TcParams {
  float* out;
  const float* res;
  const float* bias;
  const float* scale;   // [scale | inv scale]: per-output-channel powers of two un-/re-scaling the packed weights (engine.pack_tc_weights)
  int scale_pad;        // floats between the two vectors (Cout rounded up to 64)
  long long M;     // rows (padded positions) of this launch
  int H, W, Hp, Wp;
  int nchunk, ntaps, Cout, NC, MT, TPS, SA, SB;
  int Rpad, RB, nbA, halo;   // halo = window rows BEFORE the tile's first row
  int tapw;                  // taps per stencil row: 3 (3x3), 2 (2x2 over the space-to-depth tensor), 1 (1x1)
  int nsub, Nsub, nboxW, NCbox;
  int relu, tmem_cols, bo_mode;
  int tiles_m, total_work, dbuf;   // persistent schedule: work item w -> (tile = w % tiles_m, n-slice = w / tiles_m)
  long long* prof;                 // optional per-CTA cycle counters of the MMA warp (PE_TC_PROF=1): [wait_b, issue, wait_a, wait_main, total, stages]
  int dbg;                         // PE_TC_DBG experiment bits: 1 = issue no MMAs, 2 = no epilogue global traffic
  int SPD;                         // weight stages per drain group (accumulation length bound, see kernel comment)
};

// ------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Non-suspending poll.  Measured on B200: a thread parked in mbarrier.try_wait is NOT woken promptly by the
// completing arrive -- every wait that was not already satisfied cost a ~1300-cycle sleep quantum, which made the
// pipeline skeleton (not the MMAs, not the loads) 70 % of this kernel's time.  test_wait never parks the thread.
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug must trap (launch failure) instead of hanging the GPU box
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try(bar, parity)) {
    if (++spins > 200000000u) {
      printf("conv_tc: mbarrier timeout (block %d,%d thread %d bar %u parity %u)\n", blockIdx.x, blockIdx.y, threadIdx.x, bar, parity);
      __trap();
    }
  }
}
// same, for waits that are expected to be long (epilogue): back off so the polling does not crowd the LSU
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity, uint32_t ns) {
  uint32_t spins = 0;
  while (!mbar_try(bar, parity)) {
    if (ns) __nanosleep(ns);
    if (++spins > 100000000u) { printf("conv_tc: mbarrier timeout (relaxed)\n"); __trap(); }
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"((uint64_t)tm), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
#if PE_FP16
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
#else
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
#endif
      "}\n"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .b32 rx;\n"
      ".reg .pred px;\n"
      "elect.sync rx|px, 0xFFFFFFFF;\n"
      "@px mov.s32 %0, 1;\n"
      "}\n"
      : "+r"(pred));
  return pred;
}

__device__ __forceinline__ void tc_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// UMMA shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, int bo_mode) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;                         // leading-dim byte offset (unused for swizzled K-major) = 16 B
  d |= (uint64_t)((8 * CHB) >> 4) << 32;          // stride-dim byte offset: next 8-row group
  d |= (uint64_t)1 << 46;                         // descriptor version (sm_100)
  // Measured on B200 (tests/tc_bringup.py): the 128B swizzle XOR is applied to ABSOLUTE shared-memory address bits,
  // so a descriptor that starts at an arbitrary row of a TMA-written window needs base_offset = 0; setting the
  // "swizzle phase" there (bo_mode 0, kept for the experiment) reads the wrong 16-byte chunks.
  if (bo_mode == 0) d |= (uint64_t)((saddr >> 7) & 7u) << 49;
  d |= (uint64_t)(CHB == 128 ? 2 : 4) << 61;      // SWIZZLE_128B / SWIZZLE_64B
  return d;
}

// ------------------------------------------------------------------------------------------ kernel
// Persistent: gridDim.x CTAs walk the work list; TMA and MMA warps run ahead while the epilogue warps drain.
//
// Accumulation-length bound.  Measured on B200 (tests/tc_bringup.py): tcgen05 kind::tf32 accumulates into TMEM with
// truncation, a bias of ~1.2e-8 (relative) per MMA step that grows linearly with K -- 1.6e-5 at K=3456, too much for
// the 1e-3 px keypoint gate after ~100 layers (the bias is systematic, so it compounds through the residual stream).
// So no TMEM accumulator ever sees more than ~6 hi*hi steps: the MMA warp rotates through NMAIN `main` accumulators, one
// drain group (SPD weight stages) each,
// and the epilogue warps add every drained partial into FP32 registers (round-to-nearest).  The two cross terms
// hi*lo + lo*hi are 2^-11 smaller, so their truncation is harmless and they accumulate over the whole K in `corr`
// (double-buffered per tile).  TMEM columns: main0 | main1 | corr0 | corr1, MT*NC = NG*16 columns each.
__device__ __forceinline__ void st_shared_u4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, int c0, int c1, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
               ::"l"((uint64_t)tm), "r"(c0), "r"(c1), "r"(src)
               : "memory");
}

// Ring cursor without integer division (the issue loops are latency-critical: ONE warp's scalar instruction stream
// paces the tensor core; runtime div/mod per pipeline stage cost more than the MMAs themselves -- measured).
struct Ring {
  uint32_t idx = 0, phase = 0;
  __device__ __forceinline__ void advance(uint32_t n) { if (++idx == n) { idx = 0; phase ^= 1u; } }
};

__device__ __forceinline__ uint64_t desc64(uint32_t hi, uint32_t lo) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}

template <int NG, int MT, int TPS>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmO, const TcParams p) {
  constexpr int NC = NG * 16 / MT;                 // output channels per CTA
  constexpr uint32_t GC = NG * 16;                 // columns of one accumulator set (= MT*NC)
  constexpr uint32_t NMAIN = (NG <= 6) ? 3u : 2u;   // `main` accumulator buffers in flight (TMEM: (NMAIN+2)*GC <= 512 columns)
  constexpr uint32_t b_bytes = (uint32_t)TPS * NC * CHB;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t a_bytes = (uint32_t)p.Rpad * CHB;
  const uint32_t sA = base;
  const uint32_t sB = sA + p.SA * a_bytes;
  const uint32_t sStage = sB + p.SB * b_bytes;     // epilogue store staging: 8 warps x 2 buffers x (32 rows x CHB), hardware swizzle
  const uint32_t sBar = sStage + (uint32_t)EPI_WARPS * 2u * 32u * CHB;  // 8-byte barriers
  const uint32_t bar_a_full = sBar, bar_a_empty = sBar + 8 * p.SA;
  const uint32_t bar_b_full = bar_a_empty + 8 * p.SA, bar_b_empty = bar_b_full + 8 * p.SB;
  const uint32_t bar_main_full = bar_b_empty + 8 * p.SB;    // [NMAIN] (room for 4)
  const uint32_t bar_main_empty = bar_main_full + 32;       // [NMAIN]
  const uint32_t bar_corr_empty = bar_main_empty + 32;      // [2]
  const uint32_t bar_corr_full = bar_corr_empty + 16;       // [2]
  const uint32_t s_tmem = bar_corr_full + 16;
  uint8_t* gen = smem_raw + (base - raw);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen + (s_tmem - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.SA; ++i) { mbar_init(bar_a_full + 8 * i, 1); mbar_init(bar_a_empty + 8 * i, 2); }   // empty: both MMA warps
    for (int i = 0; i < p.SB; ++i) { mbar_init(bar_b_full + 8 * i, 1); mbar_init(bar_b_empty + 8 * i, 2); }
    for (int i = 0; i < (int)NMAIN; ++i) { mbar_init(bar_main_full + 8 * i, 1); mbar_init(bar_main_empty + 8 * i, EPI_WARPS); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_corr_empty + 8 * i, EPI_WARPS); mbar_init(bar_corr_full + 8 * i, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_tmem), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // broadcast through a shuffle so the compiler knows the value is warp-uniform (UTCHMMA operands live in uniform registers)
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  const int ngroups = p.ntaps / TPS;

  if (warp == 0) {
    // ===================== activation-window TMA producer (one lane) =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
      Ring ra;
      int tile = blockIdx.x % p.tiles_m;
      for (int w = blockIdx.x; w < p.total_work; w += gridDim.x) {
        const int m0 = tile * 128 * MT;            // row index fits 31 bits (asserted on the host)
        for (int j = 0; j < p.nchunk; ++j) {
          mbar_wait(bar_a_empty + 8 * ra.idx, ra.phase ^ 1u);
          const uint32_t full = bar_a_full + 8 * ra.idx, dst = sA + ra.idx * a_bytes;
          if (p.dbg & 8) { mbar_arrive(full); }
          else {
            mbar_expect_tx(full, a_bytes);
            for (int b = 0; b < p.nbA; ++b) tma_load_2d(dst + (uint32_t)b * p.RB * CHB, &tmA, j * (CHB / 4), m0 - p.halo + b * p.RB, full);
          }
          ra.advance(p.SA);
        }
        tile += gridDim.x;
        while (tile >= p.tiles_m) tile -= p.tiles_m;
      }
    }
  } else if (warp == 2 + EPI_WARPS) {
    // ===================== weight TMA producer (one lane): its own warp, so a full activation ring never delays weights
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmW) : "memory");
      Ring rb;
      int tile = blockIdx.x % p.tiles_m, nsl = blockIdx.x / p.tiles_m;
      const int tap_stride = p.nchunk * p.Cout;
      for (int w = blockIdx.x; w < p.total_work; w += gridDim.x) {
        int wrow = nsl * NC;                        // row of W tile (tap 0, chunk j): (tap*nchunk + j)*Cout + n0
        for (int j = 0; j < p.nchunk; ++j, wrow += p.Cout) {
          int wr = wrow;
          for (int g = 0; g < ngroups; ++g) {
            mbar_wait(bar_b_empty + 8 * rb.idx, rb.phase ^ 1u);
            const uint32_t full = bar_b_full + 8 * rb.idx, dst = sB + rb.idx * b_bytes;
            if (p.dbg & 4) mbar_arrive(full);
            else {
              mbar_expect_tx(full, b_bytes);
#pragma unroll
              for (int t = 0; t < TPS; ++t, wr += tap_stride) tma_load_2d(dst + (uint32_t)(t * NC) * CHB, &tmW, 0, wr, full);
            }
            rb.advance(p.SB);
          }
        }
        tile += gridDim.x;
        while (tile >= p.tiles_m) { tile -= p.tiles_m; ++nsl; }
      }
    }
  } else if (warp == 1 || warp == 3 + EPI_WARPS) {
    // ===================== MMA issuers: warp-uniform control flow, one elected lane issues =====================
    // Measured on B200 (tools/issue_bench.cu): for N <= 128 a tcgen05.mma blocks its issuing thread for the whole shared-
    // memory operand fetch (~51 clk at N = 48), so NOTHING the issuing warp does besides issuing -- barrier polls, commits,
    // address arithmetic -- overlaps with tensor work.  Two issuing warps interleave in the tensor pipe (44 clk per N = 48
    // MMA, each warp's non-MMA time hidden behind the other's MMAs), so the work is split by accumulator:
    //   warp X (1):  hi*hi  -> `main` accumulators, owns the drain-group protocol with the epilogue
    //   warp Y (11): hi*lo + lo*hi -> `corr` accumulator of the tile, commits corr_full at the end of a tile
    // Both wait on the same a_full / b_full barriers; a stage is free once BOTH have committed (a_empty / b_empty count 2).
    const bool roleX = (warp == 1);
    const bool no_mma = (p.dbg & 1) != 0;
    // instruction descriptor: D=F32, A=B=TF32, K-major both, N = NC, M = 128
    constexpr uint32_t FMT = PE_FP16 ? 0u : 2u;                       // A/B format: F16 = 0, TF32 = 2
    constexpr uint32_t idesc = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(NC >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t d0 = umma_desc(sA, p.bo_mode);
    const uint32_t desc_hi = (uint32_t)(d0 >> 32);                    // identical for A and B tiles
    const uint32_t a_lo0 = (uint32_t)d0, b_lo0 = (uint32_t)umma_desc(sB, p.bo_mode);
    const uint32_t a_step = a_bytes >> 4;
    const uint32_t wp8 = (uint32_t)p.Wp * ROW16;                      // one image row down, in 16-byte units of the window
    Ring ra, rb;
    uint32_t dg = 0, dgp = 0, tl = 0;       // drain-group buffer / phase, tile counter
    bool b_ready = false, a_ready = false, m_ready = false;   // results of early polls (latency hidden behind MMA issue)
    long long c_wb = 0, c_is = 0, c_wa = 0, c_wm = 0, c_st = 0, c_wc = 0;
    const bool prof = p.prof != nullptr;
    const long long c_t0 = prof ? clock64() : 0;
    for (int w = blockIdx.x; w < p.total_work; w += gridDim.x, ++tl) {
      const uint32_t cbuf = tl & 1u;
      if (!roleX) {
        const long long cc = prof ? clock64() : 0;
        mbar_wait(bar_corr_empty + 8 * cbuf, ((tl >> 1) & 1u) ^ 1u);   // epilogue has read this corr buffer
        if (prof) c_wc += clock64() - cc;
        tc_fence_after();
      }
      const uint32_t d_corr = tmem_base + (NMAIN + cbuf) * GC;
      uint32_t d_main = tmem_base;
      int sj = 0;                           // stages issued into the current drain group
      const int last_stage = p.nchunk * ngroups - 1;
      int stage_no = 0;
      for (int j = 0; j < p.nchunk; ++j) {
        long long c0 = prof ? clock64() : 0;
        if (!a_ready) mbar_wait(bar_a_full + 8 * ra.idx, ra.phase);
        tc_fence_after();
        if (prof) c_wa += clock64() - c0;
        const uint32_t a_slot = a_lo0 + ra.idx * a_step;
        uint32_t sh8 = 0;                   // window row shift of the group's first tap, in 16-byte units
        uint32_t kx = 0;
        for (int g = 0; g < ngroups; ++g, ++stage_no) {
          if (roleX && sj == 0) {
            const long long cm = prof ? clock64() : 0;
            if (!m_ready) mbar_wait(bar_main_empty + 8 * dg, dgp ^ 1u); // epilogue has drained this main buffer
            tc_fence_after();
            d_main = tmem_base + dg * GC;
            if (prof) c_wm += clock64() - cm;
          }
          long long c2 = prof ? clock64() : 0;
          if (!b_ready) mbar_wait(bar_b_full + 8 * rb.idx, rb.phase);
          tc_fence_after();
          if (prof) { const long long c3 = clock64(); c_wb += c3 - c2; c2 = c3; ++c_st; }
          const uint32_t b_slot = b_lo0 + rb.idx * (b_bytes >> 4);
          const uint32_t b_empty_bar = bar_b_empty + 8 * rb.idx;
          rb.advance(p.SB);
          b_ready = mbar_try(bar_b_full + 8 * rb.idx, rb.phase);   // poll the NEXT stage now; its latency hides behind the MMA issue
          const uint32_t acc_main = (sj == 0) ? 0u : 1u;
          const uint32_t acc_corr = (j == 0 && g == 0) ? 0u : 1u;
          if (no_mma) {
            if (lane == 0) mbar_arrive(b_empty_bar);
          } else if (roleX) {
            if (elect_one()) {
#pragma unroll
              for (int t = 0; t < TPS; ++t) {
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                  const uint32_t a = a_slot + sh8 + (uint32_t)(t * ROW16 + mt * 128 * ROW16);   // +1 row per tap, +128 rows per mt
                  const uint32_t b = b_slot + (uint32_t)(t * NC * ROW16);
                  const uint32_t dm = d_main + (uint32_t)(mt * NC);
#pragma unroll
                  for (int ks = 0; ks < KSTEPS; ++ks)
                    tc_mma_tf32(dm, desc64(desc_hi, a + 2 * ks), desc64(desc_hi, b + 2 * ks), idesc, (t == 0 && ks == 0) ? acc_main : 1u);   // hi * hi
                }
              }
              tc_commit(b_empty_bar);                         // weights stage free once these MMAs (and warp Y's) retire
            }
          } else {
            if (elect_one()) {
#pragma unroll
              for (int t = 0; t < TPS; ++t) {
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                  const uint32_t a = a_slot + sh8 + (uint32_t)(t * ROW16 + mt * 128 * ROW16);
                  const uint32_t b = b_slot + (uint32_t)(t * NC * ROW16);
                  const uint32_t dc = d_corr + (uint32_t)(mt * NC);
                  // hi/h half at +0, lo/l half at +LO16 (16-byte units); tf32: two K=8 steps of 32 B, fp16: one K=16 step
#pragma unroll
                  for (int ks = 0; ks < KSTEPS; ++ks) {
                    const uint32_t ak = a + 2 * ks, bk = b + 2 * ks;
                    tc_mma_tf32(dc, desc64(desc_hi, ak), desc64(desc_hi, bk + LO16), idesc, (t == 0 && ks == 0) ? acc_corr : 1u);   // hi * lo
                    tc_mma_tf32(dc, desc64(desc_hi, ak + LO16), desc64(desc_hi, bk), idesc, 1u);                                    // lo * hi
                  }
                }
              }
              tc_commit(b_empty_bar);
            }
          }
          if (prof) c_is += clock64() - c2;
          if (roleX && (++sj == p.SPD || stage_no == last_stage)) {
            if (no_mma) { if (lane == 0) mbar_arrive(bar_main_full + 8 * dg); }
            else if (elect_one()) tc_commit(bar_main_full + 8 * dg);     // this drain group's partial sums are complete
            sj = 0;
            if (++dg == NMAIN) { dg = 0; dgp ^= 1u; }
            m_ready = mbar_try(bar_main_empty + 8 * dg, dgp ^ 1u);
          }
          // next group's first tap: TPS>1 -> a group is one stencil row, go one image row down; TPS==1 -> next tap
          if (TPS > 1) sh8 += wp8;
          else if (p.ntaps > 1) { if (++kx == (uint32_t)p.tapw) { kx = 0; sh8 += wp8 - ROW16 * (p.tapw - 1); } else sh8 += ROW16; }
          __syncwarp();
        }
        if (no_mma) { if (lane == 0) mbar_arrive(bar_a_empty + 8 * ra.idx); }
        else if (elect_one()) tc_commit(bar_a_empty + 8 * ra.idx);     // activation window free (once both issuers committed)
        ra.advance(p.SA);
        a_ready = mbar_try(bar_a_full + 8 * ra.idx, ra.phase);         // early polls for the next chunk
        __syncwarp();
      }
      if (!roleX) {
        if (no_mma) { if (lane == 0) mbar_arrive(bar_corr_full + 8 * cbuf); }
        else if (elect_one()) tc_commit(bar_corr_full + 8 * cbuf);     // every cross-term MMA of this tile has retired
        __syncwarp();
      }
    }
    if (prof && lane == 0) {
      long long* o = p.prof + (size_t)blockIdx.x * 16 + (roleX ? 0 : 8);
      o[0] = c_wb; o[1] = c_is; o[2] = c_wa; o[3] = c_wm; o[4] = clock64() - c_t0; o[5] = c_st; o[6] = c_wc;
    }
  } else {
    // ===================== epilogue (warps 2..9; TMEM lane quarter = warp & 3; the two warps of a quarter take the even / odd
    // 16-column groups, so TMEM drains, residual loads, the fp16 split and the stores of one tile run on 8 warps) ==========
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int NGH = (NG + 1) / 2;                       // groups per warp (the odd warp of an odd NG has one fewer)
    const int rowF = ps_row_floats(p.Cout);
    constexpr int CF = PS_CHUNK_FLOATS;                     // floats per 16-channel chunk of a row
    constexpr int gpm = NC / 16;                            // 16-column groups per 128-row accumulator
    const int ndrain = (p.nchunk * ngroups + p.SPD - 1) / p.SPD;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const int hpwp = p.Hp * p.Wp;
    const uint32_t st_base = sStage + (uint32_t)(warp - 2) * 2u * 32u * CHB;
    uint32_t st_cnt = 0;
    uint32_t tl = 0, dg = 0, dgp = 0;
    int tile = blockIdx.x % p.tiles_m, nsl = blockIdx.x / p.tiles_m;
    for (int w = blockIdx.x; w < p.total_work; w += gridDim.x, ++tl) {
      const long long m0 = (long long)tile * 128 * MT;
      const int n0 = nsl * NC;
      float acc[NGH][16];
      // interior test of this lane's row in each of the MT 128-row accumulators (bit mt)
      uint32_t interior = 0;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const long long m = m0 + mt * 128 + q * 32 + lane;
        if (m < p.M) {
          const int r = (int)(m % hpwp);
          const int py = r / p.Wp, px = r - py * p.Wp;
          if (py >= 1 && py <= p.H && px >= 1 && px <= p.W) interior |= 1u << mt;
        }
      }
      if (p.res && !(p.dbg & 2)) {
        // pull this tile's residual rows towards L2 while the first chunk's MMAs run
#pragma unroll
        for (int gi = 0; gi < NGH; ++gi) {
          const int g = 2 * gi + half;
          if (g < NG && ((interior >> (g / gpm)) & 1u)) {
            const long long m = m0 + (g / gpm) * 128 + q * 32 + lane;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.res + m * rowF + (n0 / 16 + g % gpm) * CF));
          }
        }
      }
      for (int d = 0; d < ndrain; ++d) {
        mbar_wait_relaxed(bar_main_full + 8 * dg, dgp, (p.dbg >> 8) & 0xfff);
        tc_fence_after();
        if (!(p.dbg & 16)) {
          // two 16-column groups per TMEM wait (32 live registers)
#pragma unroll
          for (int g2 = 0; g2 < NGH; g2 += 2) {
            uint32_t r[2][16];
#pragma unroll
            for (int u = 0; u < 2; ++u)
              if (g2 + u < NGH && 2 * (g2 + u) + half < NG) tc_ld16_nowait(t_lane + dg * GC + (2 * (g2 + u) + half) * 16, r[u]);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            // accumulators stay in the SCALED domain (weights were multiplied by 2^k per channel); the residual is brought
            // into that domain when it is added and the final phase multiplies by 2^-k: all exact
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int gi = g2 + u;
              if (gi >= NGH || 2 * gi + half >= NG) continue;
              if (d == 0) {
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[gi][i] = __uint_as_float(r[u][i]);
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[gi][i] += __uint_as_float(r[u][i]);
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_main_empty + 8 * dg);
        if (++dg == NMAIN) { dg = 0; dgp ^= 1u; }
        if (d == 0 && p.res && !(p.dbg & 2)) {
          // Residual add, folded into the accumulators NOW: the loads' HBM/L2 latency hides behind the MMAs of the
          // remaining channel chunks instead of sitting in the store phase.  All loads are issued before the first use.
          constexpr int NV = CHB / 16;                       // 16-byte vectors per row chunk
#pragma unroll
          for (int g2 = 0; g2 < NGH; g2 += 2) {
            float4 rv[2][NV];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int g = 2 * (g2 + u) + half;
              const bool ok = g2 + u < NGH && g < NG && ((interior >> (g / gpm)) & 1u);
              if (ok) {
                const long long m = m0 + (g / gpm) * 128 + q * 32 + lane;
                const float4* rp = reinterpret_cast<const float4*>(p.res + m * rowF + (n0 / 16 + g % gpm) * CF);
#pragma unroll
                for (int i = 0; i < NV; ++i) rv[u][i] = __ldg(rp + i);
              } else {
#pragma unroll
                for (int i = 0; i < NV; ++i) rv[u][i] = make_float4(0.f, 0.f, 0.f, 0.f);
              }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int gi = g2 + u, g = 2 * gi + half;
              if (gi >= NGH || g >= NG) continue;
              float isc[16];                       // 2^k of this group's channels (exact)
              {
                const float4* ip = reinterpret_cast<const float4*>(p.scale + p.scale_pad + n0 + (g % gpm) * 16);
#pragma unroll
                for (int i = 0; i < 4; ++i) { const float4 t4 = __ldg(ip + i); isc[4 * i] = t4.x; isc[4 * i + 1] = t4.y; isc[4 * i + 2] = t4.z; isc[4 * i + 3] = t4.w; }
              }
#if PE_FP16
              // rv[0..1] = 16 h halfs, rv[2..3] = 16 l halfs
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                const uint32_t* hw = reinterpret_cast<const uint32_t*>(&rv[u][i]);
                const uint32_t* lw = reinterpret_cast<const uint32_t*>(&rv[u][2 + i]);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw[k]));
                  const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(&lw[k]));
                  acc[gi][8 * i + 2 * k + 0] = fmaf(hf.x + lf.x, isc[8 * i + 2 * k + 0], acc[gi][8 * i + 2 * k + 0]);
                  acc[gi][8 * i + 2 * k + 1] = fmaf(hf.y + lf.y, isc[8 * i + 2 * k + 1], acc[gi][8 * i + 2 * k + 1]);
                }
              }
#else
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                acc[gi][4 * i + 0] = fmaf(rv[u][i].x + rv[u][4 + i].x, isc[4 * i + 0], acc[gi][4 * i + 0]);
                acc[gi][4 * i + 1] = fmaf(rv[u][i].y + rv[u][4 + i].y, isc[4 * i + 1], acc[gi][4 * i + 1]);
                acc[gi][4 * i + 2] = fmaf(rv[u][i].z + rv[u][4 + i].z, isc[4 * i + 2], acc[gi][4 * i + 2]);
                acc[gi][4 * i + 3] = fmaf(rv[u][i].w + rv[u][4 + i].w, isc[4 * i + 3], acc[gi][4 * i + 3]);
              }
#endif
            }
          }
        }
      }
      // cross terms: committed by the second MMA warp at the end of the tile
      const uint32_t cbuf = tl & 1u;
      mbar_wait_relaxed(bar_corr_full + 8 * cbuf, (tl >> 1) & 1u, 0);
      tc_fence_after();
      if (!(p.dbg & 16)) {
#pragma unroll
        for (int g2 = 0; g2 < NGH; g2 += 2) {
          uint32_t r[2][16];
#pragma unroll
          for (int u = 0; u < 2; ++u)
            if (g2 + u < NGH && 2 * (g2 + u) + half < NG) tc_ld16_nowait(t_lane + (NMAIN + cbuf) * GC + (2 * (g2 + u) + half) * 16, r[u]);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int gi = g2 + u;
            if (gi >= NGH || 2 * gi + half >= NG) continue;
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[gi][i] += __uint_as_float(r[u][i]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_corr_empty + 8 * cbuf);
      // ---- bias / ReLU / split / store (the MMA warp is already on the next tile)
#pragma unroll
      for (int gi = 0; gi < NGH; ++gi) {
        const int g = 2 * gi + half;
        if (g >= NG || (p.dbg & 2)) continue;
        const int mt = g / gpm, c0 = (g % gpm) * 16;
        constexpr int NV = CHB / 16;                       // 16-byte vectors of one staged row: [hi.. | lo..]
        uint4 ov[NV];
        if (!((interior >> mt) & 1u)) {
#pragma unroll
          for (int i = 0; i < NV; ++i) ov[i] = make_uint4(0u, 0u, 0u, 0u);
        } else {
          const float4* bp = reinterpret_cast<const float4*>(p.bias + n0 + c0);
          const float4* sp = reinterpret_cast<const float4*>(p.scale + n0 + c0);
          float v[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 b4 = __ldg(bp + i), s4 = __ldg(sp + i);
            v[4 * i + 0] = fmaf(acc[gi][4 * i + 0], s4.x, b4.x); v[4 * i + 1] = fmaf(acc[gi][4 * i + 1], s4.y, b4.y);
            v[4 * i + 2] = fmaf(acc[gi][4 * i + 2], s4.z, b4.z); v[4 * i + 3] = fmaf(acc[gi][4 * i + 3], s4.w, b4.w);
          }
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
          }
#if PE_FP16
          uint2 h[4], l[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) split4_h(make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]), h[i], l[i]);
          ov[0] = make_uint4(h[0].x, h[0].y, h[1].x, h[1].y); ov[1] = make_uint4(h[2].x, h[2].y, h[3].x, h[3].y);
          ov[2] = make_uint4(l[0].x, l[0].y, l[1].x, l[1].y); ov[3] = make_uint4(l[2].x, l[2].y, l[3].x, l[3].y);
#else
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float4 hi, lo;
            split4(make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]), hi, lo);
            ov[i] = make_uint4(__float_as_uint(hi.x), __float_as_uint(hi.y), __float_as_uint(hi.z), __float_as_uint(hi.w));
            ov[4 + i] = make_uint4(__float_as_uint(lo.x), __float_as_uint(lo.y), __float_as_uint(lo.z), __float_as_uint(lo.w));
          }
#endif
        }
        // stage this warp's 32 rows x CHB bytes in shared memory (hardware swizzle pattern of the store tensor map:
        // conflict-free 16-byte stores), then ONE bulk tensor store writes them as full lines (a per-thread row store
        // would scatter 16-byte pieces over 32 lines per instruction).  Rows past the tensor end are clipped by TMA.
        const uint32_t sbuf = st_base + (st_cnt & 1u) * 32u * CHB;
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the buffer used two stores ago is free
        __syncwarp();
        const uint32_t srow = sbuf + (uint32_t)lane * CHB;
        // SWIZZLE_128B: 16-byte chunk index ^= row & 7;  SWIZZLE_64B: chunk index ^= (row >> 1) & 3
        const uint32_t sw = (CHB == 128) ? (lane & 7u) : ((lane >> 1) & 3u);
#pragma unroll
        for (int i = 0; i < NV; ++i) st_shared_u4(srow + (((uint32_t)i ^ sw) << 4), ov[i]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          const long long mrow = m0 + mt * 128 + q * 32;
          if (mrow < p.M) tma_store_2d(&tmO, (n0 + c0) / 16 * CF, (int)mrow, sbuf);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        ++st_cnt;
      }
      tile += gridDim.x;
      while (tile >= p.tiles_m) { tile -= p.tiles_m; ++nsl; }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // staged rows fully written before smem goes away
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

typedef void (*TcKernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const TcParams);
// (MT, NC, TPS) instantiations: NC*MT in {16..128} columns, TPS = 3 taps per weight stage for NC <= 64 3x3 convs
static TcKernelFn tc_kernel_for(int MT, int NC, int TPS) {
#define TCK(mt, nc, tps) if (MT == mt && NC == nc && TPS == tps) return conv_tc_kernel<(mt) * (nc) / 16, mt, tps>;
  TCK(1, 16, 1) TCK(1, 16, 3) TCK(1, 32, 1) TCK(1, 32, 3) TCK(2, 32, 1) TCK(2, 32, 3)
  TCK(1, 48, 1) TCK(1, 48, 3) TCK(2, 48, 1) TCK(2, 48, 3)
  TCK(1, 64, 1) TCK(1, 64, 3) TCK(2, 64, 1) TCK(2, 64, 3)
  TCK(1, 96, 1) TCK(1, 96, 3) TCK(1, 128, 1) TCK(1, 128, 3)
  TCK(1, 16, 2) TCK(1, 32, 2) TCK(2, 32, 2) TCK(1, 48, 2) TCK(2, 48, 2) TCK(1, 64, 2) TCK(2, 64, 2) TCK(1, 96, 2) TCK(1, 128, 2)
#undef TCK
  return nullptr;
}

// ------------------------------------------------------------------------------------------ host side
struct TcConvPlan {
  CUtensorMap tmA, tmW, tmO;
  TcParams p;
  int rows_per_img;
  size_t smem;
  int ns, num_sms;
  TcKernelFn kernel;
};

static int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return s ? atoi(s) : dflt;
}

// libcuda is resolved at run time through the runtime API so the library still loads (for its host-only entry points
// and the export check) on a box without a driver.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

static CUresult encode_2d(CUtensorMap* tm, const void* gptr, uint64_t dim0, uint64_t dim1, uint64_t stride1_bytes, uint32_t box0,
                          uint32_t box1) {
  EncodeTiledFn cuTensorMapEncodeTiled = get_encode_fn();
  if (!cuTensorMapEncodeTiled) return CUDA_ERROR_NOT_INITIALIZED;
  cuuint64_t gdim[2] = {dim0, dim1};
  cuuint64_t gstr[1] = {stride1_bytes};
  cuuint32_t box[2] = {box0, box1};
  cuuint32_t estr[2] = {1, 1};
  return cuTensorMapEncodeTiled(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(gptr), gdim, gstr, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CHB == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

cudaError_t tc_conv_plan_create(TcConvPlan** out, const float* in, float* outp, const float* res, const float* wtc,
                                const float* bias, int Cin, int Cout, int ks, int relu, int H, int W, int max_img) {
  if (env_int("PE_TC_DISABLE", 0)) return cudaErrorNotSupported;
  // ks = 3: 3x3 pad 1;  ks = 1: 1x1;  ks = 2: 2x2 stencil with taps at (+0,+1) rows/cols (no pad) -- the form a
  // stride-2 3x3 convolution takes over the space-to-depth repack of its input (kernels_simt.cu s2d_kernel)
  if ((ks != 1 && ks != 2 && ks != 3) || Cin % 16 || Cout % 16 || Cout > 512) return cudaErrorNotSupported;
  const int Hp = H + 2, Wp = W + 2, ntaps = ks * ks, nchunk = Cin / 16;
  const int halo = ks == 3 ? Wp + 1 : 0;
  const int halo_after = ks == 1 ? 0 : Wp + 1;
  const long long Mmax = (long long)max_img * Hp * Wp;
  if (Mmax + 1024 >= (1LL << 31)) return cudaErrorNotSupported;   // TMA row coordinates are int32
  const size_t stage_bytes = (size_t)EPI_WARPS * 2 * 32 * CHB;   // epilogue store staging
  const size_t smem_cap = 220 * 1024 - stage_bytes;
  double best = 1e30;
  TcParams bp{};
  size_t bsmem = 0;
  int bns = 0;
  const int force_mt = env_int("PE_TC_MT", 0), force_ns = env_int("PE_TC_NS", 0), force_dbuf = env_int("PE_TC_DBUF", -1);
  int num_sms = 148;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  for (int ns = 1; ns <= 8; ns *= 2) {
    if (Cout % ns) continue;
    const int NC = Cout / ns;
    if (NC % 16 || NC > 256) continue;
    if (force_ns && ns != force_ns) continue;
    for (int MT = 4; MT >= 1; --MT) {
     for (int tps = ks; tps >= 1; tps -= (ks > 1 ? ks - 1 : 1)) {
      if (!tc_kernel_for(MT, NC, tps)) continue;
      if (force_mt && MT != force_mt) continue;
      TcParams p{};
      p.NC = NC; p.MT = MT; p.nchunk = nchunk; p.ntaps = ntaps; p.Cout = Cout; p.halo = halo; p.dbuf = 1;
      p.TPS = tps;
      // accumulation-length bound: at most ~max_steps hi*hi MMA steps per TMEM accumulator before a drain
      const int max_steps = env_int("PE_TC_MAXSTEPS", 6);   // measured: 18 steps -> heatmap error 7e-5 (keypoint gate fails), 6 -> 2.3e-5
      p.SPD = std::max(1, max_steps / (KSTEPS * tps));
      p.tapw = ks;
      const int R = 128 * MT + halo + halo_after;
      p.nbA = (R + 255) / 256;
      p.Rpad = ((R + 8 * p.nbA - 1) / (8 * p.nbA)) * (8 * p.nbA);
      p.RB = p.Rpad / p.nbA;
      p.nsub = 1; p.Nsub = NC; p.nboxW = 1; p.NCbox = NC;
      const size_t a_bytes = (size_t)p.Rpad * CHB, b_bytes = (size_t)p.TPS * NC * CHB;
      // ring depths: three activation windows in flight when they fit next to >= 4 weight stages (the window loads come from
      // HBM: 2 stages left the load path latency-bound, measured), else two
      int SB = 6;
      p.SA = env_int("PE_TC_SA", 3);
      while (SB > 4 && p.SA * a_bytes + SB * b_bytes + 4096 > smem_cap) --SB;
      if (p.SA * a_bytes + SB * b_bytes + 4096 > smem_cap) {
        p.SA = 2; SB = 6;
        while (SB > 3 && p.SA * a_bytes + SB * b_bytes + 4096 > smem_cap) --SB;
        if (p.SA * a_bytes + SB * b_bytes + 4096 > smem_cap) continue;
      }
      p.SB = SB;
      int cols = 32;
      while (cols < ((MT * NC / 16 <= 6) ? 5 : 4) * MT * NC) cols <<= 1;
      p.tmem_cols = cols;
      p.tiles_m = (int)((Mmax + 128LL * MT - 1) / (128LL * MT));
      p.total_work = p.tiles_m * ns;
      const int ctas = p.total_work < num_sms ? p.total_work : num_sms;
      const double items = (double)((p.total_work + ctas - 1) / ctas);
      // clocks per work item: tf32 MMA at 2048 MAC/clk/SM, but never faster than the operands can be read from
      // shared memory (128 B/clk: A 4 KB + B NC*32 B per MMA) or fetched from L2 (~32 B/clk/SM)
      // measured (tools/mma_bench.cu): one M=128 SS tcgen05.mma takes max(N/2, 32 + N/4) clocks -- below N=128 the
      // 4 KB A-operand read from shared memory paces it -- plus ~300 clocks of barrier hand-off per pipeline stage
      const double n_mma = 3.0 * KSTEPS * ntaps * nchunk * MT;
      const double mma = n_mma * std::max(NC / 2.0, 32.0 + NC / 4.0) + 300.0 * (ntaps / tps) * nchunk;
      const double bytes = (double)nchunk * (a_bytes + (double)ntaps * NC * CHB);
      const double epi = (double)MT * (NC / 16) * 260.0 + 1500.0;
      const double item = std::max(std::max(mma, bytes / 32.0), epi);
      const double t = items * item + 4000.0;
      if (t < best) { best = t; bp = p; bsmem = p.SA * a_bytes + p.SB * b_bytes + stage_bytes + 4096; bns = ns; }
     }
    }
  }
  if (best >= 1e30) return cudaErrorNotSupported;
  TcConvPlan* pl = new TcConvPlan();
  pl->p = bp;
  // weight blob = [per-channel scale 2^-k: Cout floats padded to 64][inverse 2^k: same][packed operand]
  const float* wpack = wtc + 2 * (((Cout + 63) / 64) * 64);
  pl->p.scale_pad = ((Cout + 63) / 64) * 64;
  pl->p.out = outp; pl->p.res = res; pl->p.bias = bias; pl->p.scale = wtc;
  pl->p.H = H; pl->p.W = W; pl->p.Hp = Hp; pl->p.Wp = Wp; pl->p.relu = relu;
  pl->p.bo_mode = env_int("PE_TC_BO_MODE", 1);
  pl->p.dbg = env_int("PE_TC_DBG", 0);
  pl->p.prof = nullptr;
  if (env_int("PE_TC_PROF", 0)) { cudaMalloc(&pl->p.prof, 148 * 16 * sizeof(long long)); cudaMemset(pl->p.prof, 0, 148 * 16 * sizeof(long long)); }
  pl->rows_per_img = Hp * Wp;
  pl->smem = bsmem;
  pl->ns = bns;
  const uint32_t cf = PS_CHUNK_FLOATS;    // tensor maps address 4-byte words: one 16-channel chunk = cf words
  CUresult r1 = encode_2d(&pl->tmA, in, (uint64_t)ps_row_floats(Cin), (uint64_t)Mmax, (uint64_t)ps_row_floats(Cin) * 4, cf, (uint32_t)bp.RB);
  CUresult r2 = encode_2d(&pl->tmW, wpack, cf, (uint64_t)ntaps * nchunk * Cout, CHB, cf, (uint32_t)bp.NCbox);
  CUresult r3 = encode_2d(&pl->tmO, outp, (uint64_t)ps_row_floats(Cout), (uint64_t)Mmax, (uint64_t)ps_row_floats(Cout) * 4, cf, 32);
  if (r3 != CUDA_SUCCESS) r1 = r3;
  if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) {
    fprintf(stderr, "conv_tc: cuTensorMapEncodeTiled failed (%d, %d) Cin=%d Cout=%d RB=%d NCbox=%d\n", (int)r1, (int)r2, Cin, Cout, bp.RB, bp.NCbox);
    delete pl;
    return cudaErrorInvalidValue;
  }
  pl->kernel = tc_kernel_for(bp.MT, bp.NC, bp.TPS);
  {
    cudaError_t e = cudaFuncSetAttribute(pl->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
    if (e != cudaSuccess) { delete pl; return e; }
  }
  pl->num_sms = num_sms;
  if (env_int("PE_TC_VERBOSE", 0))
    fprintf(stderr, "conv_tc plan: Cin=%d Cout=%d ks=%d %dx%d  MT=%d NS=%d NC=%d SPD=%d TPS=%d SA=%d SB=%d Rpad=%d RB=%d smem=%zu tmem=%d work=%d\n", Cin, Cout, ks, H, W,
            bp.MT, bns, bp.NC, bp.SPD, bp.TPS, bp.SA, bp.SB, bp.Rpad, bp.RB, pl->smem, bp.tmem_cols, bp.total_work);
  *out = pl;
  return cudaSuccess;
}

void tc_conv_plan_destroy(TcConvPlan* plan) { if (plan && plan->p.prof) cudaFree(plan->p.prof); delete plan; }

cudaError_t tc_conv_launch(TcConvPlan* pl, int nimg, cudaStream_t st) {
  TcParams p = pl->p;
  p.M = (long long)nimg * pl->rows_per_img;
  p.tiles_m = (int)((p.M + 128LL * p.MT - 1) / (128LL * p.MT));
  p.total_work = p.tiles_m * pl->ns;
  const unsigned grid = (unsigned)(p.total_work < pl->num_sms ? p.total_work : pl->num_sms);
  pl->kernel<<<grid, TC_THREADS, pl->smem, st>>>(pl->tmA, pl->tmW, pl->tmO, p);
  if (p.prof) {
    long long h[148 * 16];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, p.prof, sizeof h, cudaMemcpyDeviceToHost);
    for (int role = 0; role < 2; ++role) {
      double a[7] = {0, 0, 0, 0, 0, 0, 0};
      for (unsigned b = 0; b < grid; ++b) for (int k = 0; k < 7; ++k) a[k] += (double)h[b * 16 + role * 8 + k] / grid;
      fprintf(stderr, "conv_tc prof %s (NC=%d MT=%d TPS=%d nchunk=%d ntaps=%d work=%d grid=%u): per-CTA cycles total %.0f | wait_b %.0f issue %.0f wait_a %.0f wait_main %.0f wait_corr %.0f res=%d | stages %.0f -> per stage: total %.0f wait_b %.0f issue %.0f\n",
              role ? "Y" : "X", p.NC, p.MT, p.TPS, p.nchunk, p.ntaps, p.total_work, grid, a[4], a[0], a[1], a[2], a[3], a[6], p.res ? 1 : 0, a[5], a[4] / a[5], a[0] / a[5], a[1] / a[5]);
    }
  }
  return cudaGetLastError();
}
