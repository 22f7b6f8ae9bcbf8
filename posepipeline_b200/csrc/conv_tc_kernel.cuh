// K2: stride-1 3x3 / 2x2 / 1x1 convolution as a tap-shifted GEMM on the 5th-gen tensor cores (sm_100a).
//
//   out[m][n] = act( sum_{tap} sum_{c} in[m + shift(tap)][c] * w[tap][c][n] + bias[n] (+ res[m][n]) )
//
// over the flat padded ("PS") activation matrix of pe_common.cuh: because every image carries a zero
// 1-pixel border, a 3x3/pad-1 convolution is nine GEMMs whose A operand is the SAME matrix shifted by
// (ky-1)*(W+2) + (kx-1) rows.  Replaces the cuDNN conv + BN + ReLU (+ residual) sequences that mmpose's
// HRNet.forward launches (reference call site pose_pipeline/wrappers/mmpose.py:75).
//
// Precision: split operands.  Activations and weights are stored as (hi, lo) pairs (fp16x2 or tf32x2, pe_common.cuh);
// each logical MAC is hi*hi + hi*lo + lo*hi on tcgen05.mma with FP32 accumulation in TMEM (error ~2^-21, the oracle's
// own fp32 noise level; plain TF32/BF16/FP16 cannot hold the 1e-3 px keypoint gate).
//
// One CTA (persistent, one per SM): MT accumulators of 128 rows x NC channels in TMEM.  One pipeline STAGE holds, for KC
// 16-channel chunks, (a) ONE halo window of (128*MT + 2*(W+3)) activation rows that serves all taps -- each tap's A operand
// is an UMMA shared-memory descriptor into the same window at a row offset (base_offset 0: the swizzle is a function of
// absolute smem address bits, measured) -- and (b) the weights of ALL taps of those chunks (one 3-D TMA box).  A stage is
// therefore 9*MT*3 MMAs for a 3x3 layer: the per-stage synchronisation cost is paid once per ~50 MMAs.
//
// Warp roles (16 warps, 128 registers per thread):
//   0      TMA producer (one lane): activation windows + weights of a stage, one full barrier
//   1      MMA issuer X: hi*hi  -> `main` accumulators; owns the drain-group protocol with the epilogue
//   14     MMA issuer Y: hi*lo + lo*hi -> `corr` accumulator of the tile
//   2..13  epilogue (TMEM lane quarter = warp & 3; the three warps of a quarter take the 16-column groups round-robin;
//          measured: 12 epilogue warps instead of 8 took 6 % off the residual layers, which are epilogue-bound)
//   15     idle
// Why two issuers -- measured (tools/issue_bench.cu): for N <= 128 a tcgen05.mma blocks its issuing thread for the whole
// shared-memory operand fetch (~51 clk at N = 48), so nothing else the issuing warp does overlaps with tensor work, and the
// pure barrier skeleton of the one-issuer kernel cost as much as the MMAs.  Two issuers interleave in the tensor pipe
// (44 clk per N = 48 MMA) and hide each other's bookkeeping.
#pragma once
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <type_traits>

#include "kernels.h"
#include "pe_common.cuh"

// layout constants of this build (pe_common.cuh): bytes per 16-channel chunk row, 16-byte units per row, offset of the
// lo / l half, MMA k-steps per chunk half
#define CHB PS_CHUNK_BYTES
constexpr uint32_t ROW16 = CHB / 16;          // 8 (tf32) / 4 (fp16)
constexpr uint32_t LO16 = CHB / 32;           // 16-byte units from hi to lo: 4 / 2
constexpr int KSTEPS = PE_FP16 ? 1 : 2;       // 16 x f16 = one K=16 MMA; 16 x tf32 = two K=8 MMAs
// Epilogue organisation (template parameter SETS):
//   SETS = 1: 12 epilogue warps (three per TMEM lane quarter, splitting the 16-column groups) work on every tile; 16 warps / CTA.
//   SETS = 2: two sets of 8 epilogue warps (two per quarter) take ALTERNATE tiles, 20 warps / CTA at 96 registers.  Measured with
//             the cycle counters (profiles/r02_conv_tc_notes.md): a tile's final phase (bias / activation / split / TMA stores,
//             ~3600-4500 clocks) kept the one set away from the next tile's drains, warp X ran out of `main` accumulators after two
//             drain groups (~2000 clocks) and stalled 30-40 % of the time; with two sets one finishes tile t while the other
//             drains tile t+1.
//   SETS = 3: split roles on 16 warps: 8 DRAIN warps (two per quarter) only move drain groups TMEM -> registers (batched loads), add
//             the cross terms and hand the finished sums back through the tile's corr accumulator columns (tcgen05.st); 4 FINALIZE
//             warps (one per quarter) take them from there a whole tile later: residual, bias, activation, split, TMA stores.
//             The drain path never waits for a final phase, and the finalize warps own all staging buffers (4 x nstg instead of
//             12 x nstg: room for a deeper ring).
//   SETS = 4: the same split with 8 finalize warps (two per quarter), 20 warps / CTA at 96 registers: layers with a short K (48 / 96
//             input channels) have too short a tile for ONE warp per quarter to finish 6 groups (measured: 96 -> 96 slower than
//             SETS = 2 with four finalize warps).
__host__ __device__ constexpr int tc_threads(int sets) { return (sets == 2 || sets == 4) ? 640 : 512; }
__host__ __device__ constexpr int epi_parts(int sets) { return sets == 1 ? 3 : 2; }      // drain warps per TMEM lane quarter within one set: they split the 16-column groups round-robin
__host__ __device__ constexpr int fin_parts(int sets) { return sets == 4 ? 2 : 1; }      // split epilogue: finalize warps per quarter
__host__ __device__ constexpr int epi_warps(int sets) { return sets == 3 ? 12 : 16 - (sets == 1 ? 4 : 0); }      // warps 2 .. 2 + epi_warps - 1 (12, 16, 12, 16)
__host__ __device__ constexpr int stage_warps(int sets) { return sets >= 3 ? 4 * fin_parts(sets) : epi_warps(sets); }   // warps that own store-staging buffers
constexpr int MAX_ACC_STEPS = 6;              // hi*hi MMA steps one TMEM accumulator may take before it is drained (see kernel)
#ifndef PE_TC_PROFILE
#define PE_TC_PROFILE 0                       // 1: per-CTA cycle counters of the two MMA warps (build flag; costs issue slots)
#endif

struct TcParams {
  const float* res;
  const float* bias;
  const float* scale;   // [scale | inv scale]: per-output-channel powers of two un-/re-scaling the packed weights (engine.pack_tc_weights)
  int scale_pad;        // floats between the two vectors (Cout rounded up to 64)
  long long M;          // rows (padded positions) of this launch
  int H, W, Hp, Wp;
  int nchunk, Cout;
  int Rpad, RB, nbA, halo;   // activation window: Rpad rows loaded as nbA boxes of RB rows; halo = window rows BEFORE the tile's first row
  // Segmented windows (wide images): contiguous mode loads ONE run of 128*MT + halo rows that serves every stencil row, which
  // costs 2*(W+3) extra rows per tile -- more than the tile itself once W > 128.  Segmented mode loads one run of
  // 128*MT + 2 rows per stencil row instead (nseg runs, seg_step rows apart in the tensor, Rseg rows apart in shared memory).
  int nseg, nb_seg, Rseg, seg_step;
  int row_step;              // shared-memory rows between two stencil rows of the window: Wp (contiguous), Rseg (segmented), dilation (1-D)
  int in_coff, out_coff, res_coff;   // first 16-channel chunk of this layer's view inside wider input / output / residual tensors
  int res_rowF;              // floats per row of the residual tensor
  int res_row_off;           // residual row = output row + res_row_off (1-D temporal layers read a centre-cropped residual)
  int res_post;              // 1 = residual added AFTER the activation (VideoPose3D blocks), 0 = before (HRNet / Darknet blocks)
  int no_border;             // 1 = every row < M is an output row (1-D / GEMM layers); 0 = rows on the zero border are written as zeros
  int act;                   // 0 none, 1 ReLU, 2 SiLU, 3 GELU
  int S;                     // pipeline stages
  int nstage;                // stages per tile = nchunk / KC
  int rpg;                   // stencil rows per drain group
  int ndrain;                // drain groups per tile
  int nstg;                  // epilogue store-staging buffers per warp (1..3)
  int gather;                // 2x2 layers: 1 = the activation windows are gathered by TMA from the ORIGINAL stride-2 input (no s2d copy)
  int cpp;                   // gather: 16-channel chunks per input parity (py,px)
  int tmem_cols;
  uint32_t div_hpwp_mul, div_hpwp_sh, div_wp_mul, div_wp_sh;   // exact n / (Hp*Wp) and n / Wp for n < 2^31: (n * mul) >> sh (64-bit product)
  int mma_wait_ns;           // MMA warps: 0 = spin on test_wait (default), > 0 = suspended try_wait with this time hint
  int poll_ns;               // producer / epilogue waits: > 0 nanosleep back-off between polls, < 0 suspended try_wait with that time hint, 0 spin
  int tiles_m, total_work;   // persistent schedule: work item w -> (tile = w % tiles_m, n-slice = w / tiles_m) when grp == 0
  int prod_par;              // 1 = the producer warp issues the TMA operations of a stage from one lane each (else lane 0 issues all)
  int grp, nsplit;           // grp > 0: groups of grp M tiles, all nsplit N slices of a group before the next group (tc_work_item)
  uint32_t a_bytes, stage_bytes;
  unsigned int* flag;        // range flag of the forward (kernels.h pe_range_flag), or nullptr
  long long* prof;
  // 2x2 layers (stride-2 3x3 convolutions in space-to-depth form): 7 of the 16 (tap, input parity) weight blocks are
  // structurally zero (W'[(py,px,c)][dy][dx] = w[c][2dy+py][2dx+px] exists only for 2dy+py < 3 and 2dx+px < 3).  zmask holds
  // 4 bits per 16-channel chunk, bit (dy*2+dx) set = that tap's weights of the chunk are all zero (found by READING the
  // packed weights at plan creation, so any channel order / pruned block is handled): the MMA warps skip those steps.
  // Adding an exact zero product leaves an accumulator unchanged, so the results are bit-identical.
  uint32_t zmask[32];
  // With skipped taps a drain group of p.rpg stencil rows holds fewer than MAX_ACC_STEPS MMA steps (3.4 on average instead of 6), so
  // the 2x2 layers drained 1.8x as often per useful MMA as the 3x3 layers.  closemask (bit r = stencil row r of a tile, r = 2 *
  // chunk + dy, closes its drain group) regroups the rows greedily by ISSUED steps; a function of the layer's weights only, so every
  // tiling of the layer still sums in the same order.
  uint32_t closemask[16];
  int use_closemask;
  // Direct output stores: the epilogue warp stages its 32 rows x one chunk in shared memory as before, reads them back
  // TRANSPOSED (CHB/16 lanes per row) and writes them with plain 16-byte global stores -- 32/(CHB/16) whole row chunks per
  // instruction -- instead of one TMA store per group.  Measured: issuing a bulk tensor store blocks the warp for 300-850 clocks
  // behind the producer's loads in the TMA queue and its staging buffer stays busy ~1000 clocks; for the finalize warps of the
  // split epilogue (all groups of a quarter in ONE warp) that was the critical path.
  int dstore;
  float* out;                // output tensor base, floats per output row
  int out_rowF;
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
// suspending wait: the thread is parked by the hardware until the phase completes or the time hint expires (no issue slots,
// no power while waiting); used for the long waits of the producer / epilogue warps when p.poll_ns < 0
__device__ __forceinline__ bool mbar_try_suspend(uint32_t bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(hint_ns)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity, int ns) {
  uint32_t spins = 0;
  if (ns < 0) {
    while (!mbar_try_suspend(bar, parity, (uint32_t)(-ns))) {
      if (++spins > 100000000u) { printf("conv_tc: mbarrier timeout (suspend)\n"); __trap(); }
    }
    return;
  }
  while (!mbar_try(bar, parity)) {
    if (ns) __nanosleep(ns);
    if (++spins > 100000000u) { printf("conv_tc: mbarrier timeout (relaxed)\n"); __trap(); }
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"((uint64_t)tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst), "l"((uint64_t)tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"((uint64_t)tm), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
// ---- CTA-pair (cta_group::2) variants.  A pair = a cluster of two CTAs on the two SMs of one TPC; the leader (cluster rank 0)
// issues every MMA for both, each CTA supplies its own 128 A rows and HALF of the B tile (the weights), so each SM fetches
// half of B per MMA; barrier signals that concern both CTAs are multicast by tcgen05.commit, and loads of either CTA report
// their bytes to the LEADER's full barrier (only the leader's MMA warps wait on it).
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {      // shared::cta address -> shared::cluster address in CTA `rank`
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t bar_cluster, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(bar_cluster), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
  // default semantics (as CUTLASS ClusterBarrier::arrive): an explicit .release.cluster cost ~900 clocks per arrive (measured:
  // drains 1306 instead of 415 clocks); the TMEM reads it orders are already complete (tcgen05.wait::ld + fence::before_thread_sync)
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void tma2_load_3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar_cluster) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"((uint64_t)tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar_cluster)
      : "memory");
}
__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint32_t bar_cluster) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst), "l"((uint64_t)tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar_cluster)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar_cluster) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"((uint64_t)tm), "r"(c0), "r"(c1), "r"(bar_cluster)
      : "memory");
}
__device__ __forceinline__ void tc_commit2(uint32_t bar) {      // arrives on the barrier at this offset in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
#if PE_FP16
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
#else
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
#endif
      "}\n"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
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

// MMA-warp wait: spin (lowest latency) or hardware-suspended with a short time hint
__device__ __forceinline__ void mbar_wait_mma(uint32_t bar, uint32_t parity, int hint_ns) {
  if (hint_ns <= 0) { mbar_wait(bar, parity); return; }
  uint32_t spins = 0;
  while (!mbar_try_suspend(bar, parity, (uint32_t)hint_ns)) {
    if (++spins > 100000000u) { printf("conv_tc: mbarrier timeout (mma)\n"); __trap(); }
  }
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

// registers -> TMEM (the drain warps of the split epilogue hand the finished sums to the finalize warps through the tile's
// corr accumulator columns)
__device__ __forceinline__ void tc_st16(uint32_t taddr, const float (&a)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "f"(a[0]), "f"(a[1]), "f"(a[2]), "f"(a[3]), "f"(a[4]), "f"(a[5]), "f"(a[6]), "f"(a[7]), "f"(a[8]), "f"(a[9]),
        "f"(a[10]), "f"(a[11]), "f"(a[12]), "f"(a[13]), "f"(a[14]), "f"(a[15])
      : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// UMMA shared-memory matrix descriptor: K-major, hardware swizzle of the chunk width, 8-row groups 8*CHB bytes apart.
// Measured on B200 (tests/tc_bringup.py): the swizzle XOR is applied to ABSOLUTE shared-memory address bits, so a descriptor
// that starts at an arbitrary row of a TMA-written window needs base_offset = 0.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;                         // leading-dim byte offset (unused for swizzled K-major) = 16 B
  d |= (uint64_t)((8 * CHB) >> 4) << 32;          // stride-dim byte offset: next 8-row group
  d |= (uint64_t)1 << 46;                         // descriptor version (sm_100)
  d |= (uint64_t)(CHB == 128 ? 2 : 4) << 61;      // SWIZZLE_128B / SWIZZLE_64B
  return d;
}

// ------------------------------------------------------------------------------------------ kernel
// Accumulation-length bound.  Measured on B200 (tests/tc_bringup.py): tcgen05 accumulates into TMEM with truncation, a bias
// of ~1.2e-8 (relative) per MMA step that grows linearly with K -- 1.6e-5 at K=3456, too much for the 1e-3 px keypoint gate
// after ~100 layers (the bias is systematic, so it compounds through the residual stream).  So no TMEM accumulator ever
// sees more than MAX_ACC_STEPS hi*hi steps: warp X rotates through NMAIN `main` accumulators, one drain group (p.rpg stencil
// rows) each, and the epilogue warps add every drained partial into FP32 registers (round-to-nearest).  The two cross
// terms hi*lo + lo*hi are 2^-11 smaller, so their truncation is harmless and they accumulate over the whole K in `corr`
// (double-buffered per tile).  TMEM columns: main[NMAIN] | corr0 | corr1, MT*NC = NG*16 columns each.
__device__ __forceinline__ void st_shared_u4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, int c0, int c1, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
               ::"l"((uint64_t)tm), "r"(c0), "r"(c1), "r"(src)
               : "memory");
}

// n / d for n < 2^31 with host-computed mul = ceil(2^sh / d), sh = 31 + ceil(log2 d): exact, two instructions instead of a
// ~100-clock runtime division (the epilogue's per-tile geometry cost 700 clocks with `/` and `%`)
__device__ __forceinline__ uint32_t fastdiv(uint32_t n, uint32_t mul, uint32_t sh) { return (uint32_t)(((uint64_t)n * mul) >> sh); }

// wait until at most n of this thread's bulk store groups still have to READ their shared-memory source
__device__ __forceinline__ void bulk_wait_read(int n) {
  switch (n) {
    case 0: asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory"); break;
    case 3: asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory"); break;
    case 4: asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory"); break;
    case 5: asm volatile("cp.async.bulk.wait_group.read 5;" ::: "memory"); break;
    case 6: asm volatile("cp.async.bulk.wait_group.read 6;" ::: "memory"); break;
    default: asm volatile("cp.async.bulk.wait_group.read 7;" ::: "memory"); break;
  }
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

// Persistent schedule.  n-major (grp == 0): the CTAs sweep all M tiles for one N slice, then the next slice -- every slice
// streams the whole activation matrix again, which is free while that matrix stays in L2 (HRNet: at most two slices) and was
// the bound of the ViT / lifter / detector GEMMs (fc1: 24 slices x 302 MB = 7.2 GB per launch, 213 TFLOP/s).  Grouped
// (grp > 0): groups of grp M tiles whose activation rows fit in L2, every N slice of a group before the next group; the
// weight matrix (a few MB) stays resident.  The per-tile arithmetic does not depend on the order: bit-identical results.
__host__ __device__ __forceinline__ void tc_work_item_map(int tiles_m, int nsplit, int grp, int w, int& tile, int& nsl) {
  if (grp <= 0) { tile = w % tiles_m; nsl = w / tiles_m; return; }
  const int span = grp * nsplit, g = w / span, base = g * grp;
  const int left = tiles_m - base, gc = grp < left ? grp : left, r = w - g * span;      // the last group may be partial
  nsl = r / gc;
  tile = base + r - nsl * gc;
}
__device__ __forceinline__ void tc_work_item(const TcParams& p, int w, int& tile, int& nsl) {
  tc_work_item_map(p.tiles_m, p.nsplit, p.grp, w, tile, nsl);
}

template <int NG, int MT, int TAPS, int KC, int CG, int SETS>
__global__ void __launch_bounds__(tc_threads(SETS), 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmR, const TcParams p) {
  // CG = 2: CTA pairs (launched as clusters of 2).  A work item is a PAIR of adjacent M tiles of one n-slice: CTA `rank` of the
  // pair owns tile 2*item + rank (rows, activation window, accumulators, epilogue all its own, exactly as in the CG = 1 form)
  // and loads the weight rows [rank*NC/2, (rank+1)*NC/2) of the n-slice; the leader's two MMA warps issue M = 256 MMAs.
  constexpr int NC = NG * 16 / MT;                 // output channels per CTA
  constexpr int NCB = NC / CG;                     // weight rows (B-operand rows) this CTA holds
  constexpr int EPI_WARPS = epi_warps(SETS), EPI_PARTS = epi_parts(SETS);
  constexpr bool SPLIT = SETS >= 3;                                // drain / finalize roles
  constexpr int SET_WARPS = SPLIT ? 8 : EPI_WARPS / SETS;          // warps that drain one tile
  constexpr int CORR_WARPS = SPLIT ? 4 * fin_parts(SETS) : SET_WARPS;   // warps that release a tile's corr accumulator (split: the finalize warps)
  constexpr int STG_WARPS = stage_warps(SETS);
  const uint32_t rank = (CG == 2) ? (blockIdx.x & 1u) : 0u;
  const int wfirst = (CG == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;      // first work item, stride of the persistent schedule
  const int wstep = (CG == 2) ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr uint32_t GC = NG * 16;                 // columns of one accumulator set (= MT*NC)
  constexpr uint32_t NMAIN = (NG <= 6) ? 3u : 2u;  // `main` accumulator buffers in flight (TMEM: (NMAIN+2)*GC <= 512 columns)
  constexpr int TAPW = TAPS == 9 ? 3 : (TAPS == 4 ? 2 : 1);   // taps per stencil row
  constexpr int ROWS = TAPS / TAPW;                            // stencil rows
  constexpr uint32_t b_chunk_bytes = (uint32_t)TAPS * NCB * CHB;  // weights of all taps of one 16-channel chunk (this CTA's rows)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t a_bytes = p.a_bytes, stage_bytes = p.stage_bytes;
  const uint32_t sRing = base;                                       // S stages: [KC activation windows][KC x TAPS weight tiles]
  const uint32_t sStage = sRing + (uint32_t)p.S * stage_bytes;       // epilogue store staging: 8 warps x nstg x (32 rows x CHB)
  const uint32_t sBar = sStage + (uint32_t)STG_WARPS * (uint32_t)p.nstg * 32u * CHB;   // 8-byte barriers
  const uint32_t bar_full = sBar, bar_empty = sBar + 8 * p.S;
  const uint32_t bar_main_full = bar_empty + 8 * p.S;       // [NMAIN] (room for 4)
  const uint32_t bar_main_empty = bar_main_full + 32;       // [NMAIN]
  const uint32_t bar_corr_empty = bar_main_empty + 32;      // [2]
  const uint32_t bar_corr_full = bar_corr_empty + 16;       // [2]
  const uint32_t bar_res = bar_corr_full + 16;              // [EPI_WARPS] residual chunks of a warp have landed in its staging buffers
  const uint32_t bar_turn = bar_res + 8 * EPI_WARPS;        // [2] SETS = 2: set s has finished the drains of its current tile
  const uint32_t bar_sum = bar_turn + 16;                   // [2] SETS = 3: the drain warps have stored a tile's sums into corr buffer i
  const uint32_t s_tmem = bar_sum + 16;
  uint8_t* gen = smem_raw + (base - raw);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen + (s_tmem - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    // full: one arrival (with its byte count) per producer -- in pair mode both CTAs' producers report to the leader's barrier;
    // empty: both MMA warps; main/corr empty: the epilogue warps of every CTA whose accumulators the MMA overwrites
    for (int i = 0; i < p.S; ++i) { mbar_init(bar_full + 8 * i, CG); mbar_init(bar_empty + 8 * i, 2); }
    for (int i = 0; i < (int)NMAIN; ++i) { mbar_init(bar_main_full + 8 * i, 1); mbar_init(bar_main_empty + 8 * i, SET_WARPS * CG); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_corr_empty + 8 * i, CORR_WARPS * CG); mbar_init(bar_corr_full + 8 * i, 1); }
    for (int i = 0; i < EPI_WARPS; ++i) mbar_init(bar_res + 8 * i, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(bar_turn + 8 * i, SET_WARPS); mbar_init(bar_sum + 8 * i, SET_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if (CG == 2) {      // the same warp of both CTAs, the same destination offset
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_tmem), "r"((uint32_t)p.tmem_cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_tmem), "r"((uint32_t)p.tmem_cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all();      // the peer's barriers are initialised before anything is signalled across the pair
  else __syncthreads();
  tc_fence_after();
  // broadcast through a shuffle so the compiler knows the value is warp-uniform (UTCHMMA operands live in uniform registers)
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (p.prod_par) {
      // Lane-parallel issue.  Measured (profiles/r02_producer_notes.md): the single-lane loop below spends ~100 clocks of dependent
      // address arithmetic + issue per TMA operation; layers with many small operations per stage and little MMA work per stage
      // (the stride-2 gather: 5-14 boxes + weights per ~400-800 clocks of MMAs) were paced by the producer THREAD, not by bytes.
      // Here lane 0 waits for the free stage and posts the byte count, then every operation of the stage is issued by its own
      // lane in one pass (per-lane box coordinates are fixed per tile, or for the whole kernel in the non-gather forms).
      if (lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmW) : "memory");
      }
      Ring r;
      int tile = 0, nsl = 0;
      if (wfirst < p.total_work) tc_work_item(p, wfirst, tile, nsl);
      const bool gather = TAPS == 4 && p.gather;
      const uint32_t full0 = (CG == 2) ? mapa_rank(bar_full, 0u) : bar_full;
      // non-gather forms: lane -> (chunk of the stage, operation of the chunk); the last operation of a chunk is its weight box
      const int opk = p.nseg * p.nb_seg + 1;
      const int my_kc = lane / opk, my_o = lane - my_kc * opk;
      const bool my_on = my_kc < KC, my_w = my_o == opk - 1;
      const int my_sg = my_w ? 0 : my_o / p.nb_seg, my_b = my_w ? 0 : my_o - my_sg * p.nb_seg;
      const uint32_t my_off = my_w ? (uint32_t)KC * a_bytes + (uint32_t)my_kc * b_chunk_bytes
                                   : (uint32_t)my_kc * a_bytes + (uint32_t)(my_sg * p.Rseg + my_b * p.RB) * CHB;
      const int my_row = my_sg * p.seg_step + my_b * p.RB - p.halo;
      for (int w = wfirst; w < p.total_work; w += wstep) {
        const int m0 = (tile * CG + (int)rank) * 128 * MT;
        const int n0 = nsl * NC + (int)rank * NCB;
        uint32_t tx = (uint32_t)KC * (a_bytes + b_chunk_bytes);
        int g_nbox = 0, my_q = 0, my_n = 0;
        if (gather) {                                                   // see the single-lane form below for the geometry
          const int g0 = m0 / p.Wp, g_n = g0 / p.Hp, g_q = (g0 - g_n * p.Hp) >> 1, hq = p.Hp >> 1;
          const int off = m0 - (g_n * p.Hp + 2 * g_q) * p.Wp;
          g_nbox = (off + 128 * MT + p.Wp + 1 + 2 * p.Wp - 1) / (2 * p.Wp);
          tx = (uint32_t)g_nbox * 2u * (uint32_t)p.Wp * CHB + b_chunk_bytes;
          my_q = g_q + lane; my_n = g_n;                                // box `lane`: S image-row pair my_q of image my_n
          while (my_q >= hq) { my_q -= hq; ++my_n; }
        }
        int j = 0, gpar = 0, gjj = 0;
        for (int st = 0; st < p.nstage; ++st) {
          const uint32_t full = full0 + 8 * r.idx, dst = sRing + r.idx * stage_bytes;
          if (lane == 0) {
            mbar_wait_relaxed(bar_empty + 8 * r.idx, r.phase ^ 1u, p.poll_ns);
            if (CG == 2) mbar_expect_tx_cluster(full, tx); else mbar_expect_tx(full, tx);
          }
          __syncwarp();
          if (gather) {                                                 // one chunk per stage (KC = 1)
            if (lane < g_nbox) {
              const uint32_t da = dst + (uint32_t)lane * 2u * (uint32_t)p.Wp * CHB;
              if (CG == 2) tma2_load_4d(da, &tmA, (p.in_coff + gjj) * (CHB / 4), (gpar & 1) - 2, 4 * my_q - 2 + (gpar >> 1), my_n, full);
              else tma_load_4d(da, &tmA, (p.in_coff + gjj) * (CHB / 4), (gpar & 1) - 2, 4 * my_q - 2 + (gpar >> 1), my_n, full);
            } else if (lane == g_nbox) {
              if (CG == 2) tma2_load_3d(dst + (uint32_t)KC * a_bytes, &tmW, 0, j * p.Cout + n0, 0, full);
              else tma_load_3d(dst + (uint32_t)KC * a_bytes, &tmW, 0, j * p.Cout + n0, 0, full);
            }
            ++j;
            if (++gjj == p.cpp) { gjj = 0; ++gpar; }
          } else {
            if (my_on) {
              const int jc = j + my_kc;
              if (my_w) {
                if (CG == 2) tma2_load_3d(dst + my_off, &tmW, 0, jc * p.Cout + n0, 0, full);
                else tma_load_3d(dst + my_off, &tmW, 0, jc * p.Cout + n0, 0, full);
              } else {
                if (CG == 2) tma2_load_2d(dst + my_off, &tmA, (p.in_coff + jc) * (CHB / 4), m0 + my_row, full);
                else tma_load_2d(dst + my_off, &tmA, (p.in_coff + jc) * (CHB / 4), m0 + my_row, full);
              }
            }
            j += KC;
          }
          r.advance(p.S);
        }
        if (p.grp > 0) { if (w + wstep < p.total_work) tc_work_item(p, w + wstep, tile, nsl); }
        else { tile += wstep; while (tile >= p.tiles_m) { tile -= p.tiles_m; ++nsl; } }
      }
    } else if (lane == 0) {                 // single-lane form (PE_TC_PLANES=1, or more than 32 operations per stage)
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmW) : "memory");
      Ring r;
      int tile = 0, nsl = 0;
      if (wfirst < p.total_work) tc_work_item(p, wfirst, tile, nsl);
      uint32_t tx = (uint32_t)KC * (a_bytes + b_chunk_bytes);
      const bool gather = TAPS == 4 && p.gather;
      const uint32_t full0 = (CG == 2) ? mapa_rank(bar_full, 0u) : bar_full;    // pair mode: the LEADER's full barriers (cluster address)
      for (int w = wfirst; w < p.total_work; w += wstep) {
        const int m0 = (tile * CG + (int)rank) * 128 * MT;            // row index fits 31 bits (asserted on the host)
        const int n0 = nsl * NC + (int)rank * NCB;                    // first weight row this CTA loads
        // Gather mode (stride-2 3x3 convolution in its 2x2 space-to-depth form): the rows [m0, m0 + 128*MT + Wp + 1) of the
        // space-to-depth tensor S[a'][b'][(py,px,c)] = in[2(a'-1)+py][2(b'-1)+px][c] are never materialised; TMA gathers them
        // from the original tensor with element strides (2, 2), one box = two whole S image rows (a' = 2q, 2q+1) of one
        // parity (py,px) and one 16-channel chunk.  Out-of-image coordinates (a' = 0, b' = 0) are zero-filled by TMA.
        int g_n = 0, g_q = 0, g_nbox = 0;
        if (gather) {
          const int g0 = m0 / p.Wp;
          g_n = g0 / p.Hp;
          g_q = (g0 - g_n * p.Hp) >> 1;
          const int off = m0 - (g_n * p.Hp + 2 * g_q) * p.Wp;
          g_nbox = (off + 128 * MT + p.Wp + 1 + 2 * p.Wp - 1) / (2 * p.Wp);
          tx = (uint32_t)g_nbox * 2u * (uint32_t)p.Wp * CHB + b_chunk_bytes;
        }
        int j = 0;
        for (int st = 0; st < p.nstage; ++st) {
          mbar_wait_relaxed(bar_empty + 8 * r.idx, r.phase ^ 1u, p.poll_ns);
          const uint32_t full = full0 + 8 * r.idx, dst = sRing + r.idx * stage_bytes;
          if (CG == 2) mbar_expect_tx_cluster(full, tx); else mbar_expect_tx(full, tx);
#pragma unroll
          for (int kc = 0; kc < KC; ++kc, ++j) {
            if (gather) {
              const int par = j / p.cpp, jj = j - par * p.cpp;
              int qq = g_q, nn = g_n;
              for (int b = 0; b < g_nbox; ++b) {
                const uint32_t da = dst + (uint32_t)b * 2u * (uint32_t)p.Wp * CHB;
                if (CG == 2) tma2_load_4d(da, &tmA, (p.in_coff + jj) * (CHB / 4), (par & 1) - 2, 4 * qq - 2 + (par >> 1), nn, full);
                else tma_load_4d(da, &tmA, (p.in_coff + jj) * (CHB / 4), (par & 1) - 2, 4 * qq - 2 + (par >> 1), nn, full);
                if (++qq == (p.Hp >> 1)) { qq = 0; ++nn; }
              }
            } else
            for (int sg = 0; sg < p.nseg; ++sg)
              for (int b = 0; b < p.nb_seg; ++b) {
                const uint32_t da = dst + (uint32_t)kc * a_bytes + (uint32_t)(sg * p.Rseg + b * p.RB) * CHB;
                if (CG == 2) tma2_load_2d(da, &tmA, (p.in_coff + j) * (CHB / 4), m0 - p.halo + sg * p.seg_step + b * p.RB, full);
                else tma_load_2d(da, &tmA, (p.in_coff + j) * (CHB / 4), m0 - p.halo + sg * p.seg_step + b * p.RB, full);
              }
            // weights: box {one chunk row, NCB output channels, TAPS taps} of the [tap][chunk*Cout + n][CHB] tensor
            if (CG == 2) tma2_load_3d(dst + (uint32_t)KC * a_bytes + (uint32_t)kc * b_chunk_bytes, &tmW, 0, j * p.Cout + n0, 0, full);
            else tma_load_3d(dst + (uint32_t)KC * a_bytes + (uint32_t)kc * b_chunk_bytes, &tmW, 0, j * p.Cout + n0, 0, full);
          }
          r.advance(p.S);
        }
        if (p.grp > 0) { if (w + wstep < p.total_work) tc_work_item(p, w + wstep, tile, nsl); }
        else { tile += wstep; while (tile >= p.tiles_m) { tile -= p.tiles_m; ++nsl; } }
      }
    }
  } else if ((warp == 1 || warp == 2 + EPI_WARPS) && rank == 0) {
    // ===================== MMA issuers: warp-uniform control flow, one elected lane issues (pair mode: the leader's only) =====
    const bool roleX = (warp == 1);
    // instruction descriptor: D=F32, A/B format, K-major both, N = NC, M = 128 per CTA (256 over a pair)
    constexpr uint32_t FMT = PE_FP16 ? 0u : 2u;                       // A/B format: F16 = 0, TF32 = 2
    constexpr uint32_t idesc = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(NC >> 3) << 17) | (((128u * CG) >> 4) << 24);
    auto mma = [](uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) { if (CG == 2) tc_mma2(d, a, b, id, acc); else tc_mma_tf32(d, a, b, id, acc); };
    auto commit = [](uint32_t bar) { if (CG == 2) tc_commit2(bar); else tc_commit(bar); };
    const uint64_t d0 = umma_desc(sRing);
    const uint32_t desc_hi = (uint32_t)(d0 >> 32);                    // identical for A and B tiles
    const uint32_t ring_lo0 = (uint32_t)d0;
    const uint32_t stage16 = stage_bytes >> 4, a16 = a_bytes >> 4;
    const uint32_t wp16 = (uint32_t)p.row_step * ROW16;               // one stencil row down, in 16-byte units of the window
    Ring r;
#if PE_TC_PROFILE
    long long c_wf = 0, c_is = 0, c_wm = 0, c_wc = 0, c_st = 0;
    const long long c_t0 = clock64();
#endif
    if (roleX) {
      // ---------------- X: hi*hi into the rotating `main` accumulators
      uint32_t dg = 0, dgp = 0;             // drain-group buffer / phase
      uint32_t fresh = 1u;                  // the current drain group's accumulator has not been written yet
      const int total_rows = p.nstage * KC * ROWS;
      int tile = wfirst % p.tiles_m;
      for (int w = wfirst; w < p.total_work; w += wstep) {
        uint32_t a_off16 = 0;               // gather mode: the tile's first row inside the gathered window (whole S image rows)
        if (TAPS == 4 && p.gather) {
          const int m0 = tile * CG * 128 * MT, g0 = m0 / p.Wp, gn = g0 / p.Hp, gq = (g0 - gn * p.Hp) >> 1;
          a_off16 = (uint32_t)(m0 - (gn * p.Hp + 2 * gq) * p.Wp) * ROW16;
          tile += wstep;
          while (tile >= p.tiles_m) tile -= p.tiles_m;
        }
        int rig = 0, row_no = 0;            // stencil rows issued into the current drain group / in this tile
        for (int st = 0; st < p.nstage; ++st) {
#if PE_TC_PROFILE
          long long c0 = clock64();
#endif
          mbar_wait_mma(bar_full + 8 * r.idx, r.phase, p.mma_wait_ns);
          tc_fence_after();
#if PE_TC_PROFILE
          { const long long c1 = clock64(); c_wf += c1 - c0; c0 = c1; ++c_st; }
#endif
          const uint32_t a_st = ring_lo0 + r.idx * stage16 + a_off16;
          const uint32_t b_st = ring_lo0 + r.idx * stage16 + (uint32_t)KC * a16;
          const uint32_t empty_bar = bar_empty + 8 * r.idx;
#pragma unroll
          for (int kc = 0; kc < KC; ++kc) {
            const int jz = st * KC + kc;
            const uint32_t zm = (TAPS == 4) ? ((p.zmask[(jz >> 3) & 31] >> ((jz & 7) * 4)) & 15u) : 0u;
#pragma unroll
            for (int row = 0; row < ROWS; ++row) {
              if (rig == 0) {
#if PE_TC_PROFILE
                const long long cm = clock64();
#endif
                mbar_wait_mma(bar_main_empty + 8 * dg, dgp ^ 1u, p.mma_wait_ns);   // epilogue has drained this main buffer
                tc_fence_after();
#if PE_TC_PROFILE
                { const long long c1 = clock64(); c_wm += c1 - cm; c0 += c1 - cm; }
#endif
              }
              const uint32_t d_main = tmem_base + dg * GC;
              if (rig == 0) fresh = 1u;         // no MMA has written this drain group's accumulator yet
              ++row_no;
              ++rig;
              const bool close = (TAPS == 4 && p.use_closemask) ? (((p.closemask[((row_no - 1) >> 5) & 15] >> ((row_no - 1) & 31)) & 1u) != 0u)
                                                                 : ((rig == p.rpg) || (row_no == total_rows));
              // 2x2 form: taps of this stencil row whose weights are all zero for this chunk are skipped -- except that a
              // drain group must not close without a single MMA (its accumulator would hold stale sums)
              uint32_t skip = 0u;
              if (TAPS == 4) {
                skip = (zm >> (row * TAPW)) & ((1u << TAPW) - 1u);
                if (close && fresh && skip == ((1u << TAPW) - 1u)) skip &= ~1u;
              }
              if (elect_one()) {
                uint32_t fr = fresh;
#pragma unroll
                for (int t = 0; t < TAPW; ++t) {
                  if (TAPS == 4 && ((skip >> t) & 1u)) continue;
#pragma unroll
                  for (int mt = 0; mt < MT; ++mt) {
                    const uint32_t a = a_st + (uint32_t)kc * a16 + (uint32_t)row * wp16 + (uint32_t)(t * ROW16 + mt * 128 * ROW16);
                    const uint32_t b = b_st + (uint32_t)((kc * TAPS + row * TAPW + t) * NCB * ROW16);
#pragma unroll
                    for (int ks = 0; ks < KSTEPS; ++ks)
                      mma(d_main + (uint32_t)(mt * NC), desc64(desc_hi, a + 2 * ks), desc64(desc_hi, b + 2 * ks), idesc,
                                  (fr && ks == 0) ? 0u : 1u);   // hi * hi
                  }
                  fr = 0u;
                }
                if (close) commit(bar_main_full + 8 * dg);          // this drain group's partial sums are complete
                if (kc == KC - 1 && row == ROWS - 1) commit(empty_bar);   // stage free once these MMAs (and warp Y's) retire
              }
              __syncwarp();
              if (skip != ((1u << TAPW) - 1u)) fresh = 0u;
              if (close) { rig = 0; if (++dg == NMAIN) { dg = 0; dgp ^= 1u; } }
            }
          }
          r.advance(p.S);
#if PE_TC_PROFILE
          c_is += clock64() - c0;
#endif
        }
      }
    } else {
      // ---------------- Y: hi*lo + lo*hi into the tile's `corr` accumulator
      uint32_t tl = 0;
      int tile = wfirst % p.tiles_m;
      for (int w = wfirst; w < p.total_work; w += wstep, ++tl) {
        uint32_t a_off16 = 0;               // gather mode: see warp X
        if (TAPS == 4 && p.gather) {
          const int m0 = tile * CG * 128 * MT, g0 = m0 / p.Wp, gn = g0 / p.Hp, gq = (g0 - gn * p.Hp) >> 1;
          a_off16 = (uint32_t)(m0 - (gn * p.Hp + 2 * gq) * p.Wp) * ROW16;
          tile += wstep;
          while (tile >= p.tiles_m) tile -= p.tiles_m;
        }
        const uint32_t cbuf = tl & 1u;
        uint32_t freshY = 1u;               // nothing has been written to this tile's corr accumulator yet
#if PE_TC_PROFILE
        { const long long cc = clock64();
#endif
        mbar_wait(bar_corr_empty + 8 * cbuf, ((tl >> 1) & 1u) ^ 1u);   // epilogue has read this corr buffer
        tc_fence_after();
#if PE_TC_PROFILE
          c_wc += clock64() - cc; }
#endif
        const uint32_t d_corr = tmem_base + (NMAIN + cbuf) * GC;
        for (int st = 0; st < p.nstage; ++st) {
#if PE_TC_PROFILE
          long long c0 = clock64();
#endif
          mbar_wait_mma(bar_full + 8 * r.idx, r.phase, p.mma_wait_ns);
          tc_fence_after();
#if PE_TC_PROFILE
          { const long long c1 = clock64(); c_wf += c1 - c0; c0 = c1; ++c_st; }
#endif
          const uint32_t a_st = ring_lo0 + r.idx * stage16 + a_off16;
          const uint32_t b_st = ring_lo0 + r.idx * stage16 + (uint32_t)KC * a16;
          // 2x2 form: all-zero (tap, chunk) weight blocks are skipped (see TcParams::zmask); the first MMA that is issued for
          // a tile overwrites the accumulator, and the tile's last step is never skipped while nothing has been issued
          uint32_t zmst = 0u;
          if (TAPS == 4) {
#pragma unroll
            for (int kc = 0; kc < KC; ++kc) {
              const int jz = st * KC + kc;
              zmst |= ((p.zmask[(jz >> 3) & 31] >> ((jz & 7) * 4)) & 15u) << (4 * kc);
            }
            if (st == p.nstage - 1 && freshY) zmst &= ~(1u << (4 * (KC - 1) + TAPS - 1));
          }
          if (elect_one()) {
            uint32_t fr = freshY;
#pragma unroll
            for (int kc = 0; kc < KC; ++kc) {
#pragma unroll
              for (int row = 0; row < ROWS; ++row) {
#pragma unroll
                for (int t = 0; t < TAPW; ++t) {
                  if (TAPS == 4 && ((zmst >> (4 * kc + row * TAPW + t)) & 1u)) continue;
#pragma unroll
                  for (int mt = 0; mt < MT; ++mt) {
                    const uint32_t a = a_st + (uint32_t)kc * a16 + (uint32_t)row * wp16 + (uint32_t)(t * ROW16 + mt * 128 * ROW16);
                    const uint32_t b = b_st + (uint32_t)((kc * TAPS + row * TAPW + t) * NCB * ROW16);
                    // hi/h half at +0, lo/l half at +LO16 (16-byte units); tf32: two K=8 steps of 32 B, fp16: one K=16 step
#pragma unroll
                    for (int ks = 0; ks < KSTEPS; ++ks) {
                      const uint32_t ak = a + 2 * ks, bk = b + 2 * ks;
                      mma(d_corr + (uint32_t)(mt * NC), desc64(desc_hi, ak), desc64(desc_hi, bk + LO16), idesc,
                                  (fr && ks == 0) ? 0u : 1u);                                                                          // hi * lo
                      mma(d_corr + (uint32_t)(mt * NC), desc64(desc_hi, ak + LO16), desc64(desc_hi, bk), idesc, 1u);            // lo * hi
                    }
                  }
                  fr = 0u;
                }
              }
            }
            commit(bar_empty + 8 * r.idx);
            if (st == p.nstage - 1) commit(bar_corr_full + 8 * cbuf);   // every cross-term MMA of this tile has retired
          }
          __syncwarp();
          if (TAPS != 4 || zmst != ((KC == 1) ? 0xFu : (KC == 2) ? 0xFFu : 0xFFFFu)) freshY = 0u;
          r.advance(p.S);
#if PE_TC_PROFILE
          c_is += clock64() - c0;
#endif
        }
      }
    }
#if PE_TC_PROFILE
    if (p.prof && lane == 0) {
      long long* o = p.prof + (size_t)blockIdx.x * 32 + (roleX ? 0 : 8);
      o[0] = c_wf; o[1] = c_is; o[2] = c_wm; o[3] = c_wc; o[4] = clock64() - c_t0; o[5] = c_st;
    }
#endif
  } else if (warp >= 2 && warp < 2 + EPI_WARPS) {
    // ===================== epilogue (TMEM lane quarter = warp & 3).  ROLE 0 (SETS 1 / 2): the PARTS warps of a quarter take the
    // 16-column groups round-robin and do everything for their groups (drains, cross terms, residual, final phase).  SETS 3:
    // ROLE 1 = drain warps 2..9 (drains + cross terms, sums handed over through TMEM), ROLE 2 = finalize warps 10..13 (all
    // groups of their quarter: residual + final phase).  One body, instantiated per role.
    auto epilogue = [&](auto role_tag) {
    constexpr int ROLE = decltype(role_tag)::value;
    constexpr int PARTS = ROLE == 0 ? EPI_PARTS : (ROLE == 1 ? 2 : fin_parts(SETS));
    constexpr int ESETS = SETS == 2 ? 2 : 1;                // sets that alternate tiles
    const int q = warp & 3;
    const int eset = (SETS == 2) ? ((warp - 2) >> 3) : 0;   // SETS = 2: this warp's set takes the CTA's work items eset, eset + 2, ...
    const int half = ROLE == 2 ? ((warp - 10) >> 2) : ((warp - 2) >> 2) % PARTS;   // which of the PARTS warps of this quarter (within the set)
    const int sidx = ROLE == 2 ? warp - 10 : warp - 2;      // owner index of this warp's staging buffers / residual barrier
    const int ewfirst = wfirst + eset * wstep, ewstep = wstep * ESETS;
    constexpr int NGH = (NG + PARTS - 1) / PARTS;           // groups per warp (the odd warp of an odd NG has one fewer)
    constexpr int NACC = ROLE == 2 ? 1 : NGH;               // finalize warps hold one group at a time
    const int rowF = p.res_rowF;                            // floats per row of the residual tensor (plain-load fallback)
    constexpr int CF = PS_CHUNK_FLOATS;                     // floats per 16-channel chunk of a row
    constexpr int gpm = NC / 16;                            // 16-column groups per 128-row accumulator
    const int ndrain = p.ndrain;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const int hpwp = p.Hp * p.Wp;
    const uint32_t st_base = sStage + (uint32_t)sidx * (uint32_t)p.nstg * 32u * CHB;
    uint32_t st_slot = 0;
    // tl = ordinal of the tile among the CTA's work items (both sets count all of them: it selects the corr buffer);
    // (dg, dgp) = `main` buffer / phase of the tile's first drain group: the other set's tiles advance it by ndrain each
    uint32_t tl = (uint32_t)eset;
    uint32_t dg = ((uint32_t)eset * (uint32_t)ndrain) % NMAIN, dgp = (((uint32_t)eset * (uint32_t)ndrain) / NMAIN) & 1u;
    int tile = 0, nsl = 0;
    if (ewfirst < p.total_work) tc_work_item(p, ewfirst, tile, nsl);
    constexpr int NV = CHB / 16;                         // 16-byte vectors per row chunk
    // "accumulator drained" signals go to the CTA whose warps issue the MMAs: the leader of the pair
    const uint32_t main_empty0 = (CG == 2) ? mapa_rank(bar_main_empty, 0u) : bar_main_empty;
    const uint32_t corr_empty0 = (CG == 2) ? mapa_rank(bar_corr_empty, 0u) : bar_corr_empty;
    // Residual rows.  Measured: per-lane row loads (lane = accumulator row, 64 B each, 192+ B apart) cost the SM's load/store
    // unit 32 line lookups per instruction -- ~3000 clocks per 256-row tile, which made every residual layer epilogue-bound
    // (warp X waited ~2000 clk per stage for a free `main` buffer).  So the rows come through TMA instead: each epilogue warp
    // loads its 32-row x 16-channel chunks of the NEXT tile into its own (idle) store-staging buffers as soon as the current
    // tile's stores have read them, and reads them back conflict-free (the tensor map's swizzle) after the first drain.
    const uint32_t res_bar = bar_res + 8 * (uint32_t)sidx;
    auto res_issue = [&](int tile_, int nsl_) {
      if (lane == 0) {
        bulk_wait_read(0);                                                   // this warp's output stores have read the buffers
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        uint32_t nb = 0;
#pragma unroll
        for (int gi = 0; gi < NGH; ++gi) nb += (PARTS * gi + half < NG && gi < p.nstg) ? 1u : 0u;
        mbar_expect_tx(res_bar, nb * 32u * CHB);
#pragma unroll
        for (int gi = 0; gi < NGH; ++gi) {
          const int g = PARTS * gi + half;
          if (g < NG && gi < p.nstg)
            tma_load_2d(st_base + (uint32_t)gi * 32u * CHB, &tmR, (p.res_coff + (nsl_ * NC + (g % gpm) * 16) / 16) * CF,
                        (tile_ * CG + (int)rank) * 128 * MT + (g / gpm) * 128 + q * 32 + p.res_row_off, res_bar);
        }
      }
    };
    // When: measured, a TMA store takes ~1000 clocks to read its 2 KB source under load, so waiting for the stores at the end
    // of a tile stalled every epilogue warp that long.  Layers with >= 3 drain groups request a tile's residual after its
    // FIRST drain (the previous tile's stores are long done) and add it after the LAST (an epilogue that is behind runs its
    // drains back to back, so the distance must be several drains); layers with fewer (1x1) request the next tile's at the end
    // of the current one.
    const bool res_lazy = ROLE == 0 && p.res && (ndrain >= 3 || p.res_post);
    if (ROLE == 0 && p.res && !res_lazy && ewfirst < p.total_work) res_issue(tile, nsl);
#if PE_TC_PROFILE
    long long e_wm = 0, e_dr = 0, e_rs = 0, e_wc = 0, e_co = 0, e_fi = 0, e_ri = 0, e_tc = 0, e_f1 = 0, e_f2 = 0, e_f3 = 0, e_f4 = 0, e_t = clock64();
    const long long e_t0 = e_t;
#define EPI_TICK(v) { const long long _c = clock64(); v += _c - e_t; e_t = _c; }
#else
#define EPI_TICK(v)
#endif
    for (int w = ewfirst; w < p.total_work; w += ewstep, tl += ESETS) {
      const long long m0 = (long long)(tile * CG + (int)rank) * 128 * MT;
      const int n0 = nsl * NC;
      const uint32_t mytl = tl / ESETS;                    // ordinal among this set's own tiles (phase of the per-warp residual barrier)
      float acc[NACC][16];
      if (ROLE == 2 && p.res) { res_issue(tile, nsl); __syncwarp(); }      // finalize warps: this tile's residual chunks, needed a whole tile of drains from now
      // interior test of this lane's row in each of the MT 128-row accumulators (bit mt)
      uint32_t interior = 0;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const long long m = m0 + mt * 128 + q * 32 + lane;
        if (m < p.M && p.no_border) interior |= 1u << mt;
        else if (m < p.M) {                                              // M < 2^31 (asserted on the host)
          const uint32_t r = (uint32_t)m - fastdiv((uint32_t)m, p.div_hpwp_mul, p.div_hpwp_sh) * (uint32_t)hpwp;
          const int py = (int)fastdiv(r, p.div_wp_mul, p.div_wp_sh), px = (int)r - py * p.Wp;
          if (py >= 1 && py <= p.H && px >= 1 && px <= p.W) interior |= 1u << mt;
        }
      }
      // add 16 residual channels (raw chunk v) of group g into accumulator a, in the scaled domain (x 2^k per channel, exact)
      auto res_add16 = [&](float (&a)[16], const uint4 (&v)[NV], int g) {
        const float4* ip = reinterpret_cast<const float4*>(p.scale + p.scale_pad + n0 + (g % gpm) * 16);
#if PE_FP16
        // v[0..1] = 16 h halfs, v[2..3] = 16 l halfs
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const uint32_t* hw = reinterpret_cast<const uint32_t*>(&v[i]);
          const uint32_t* lw = reinterpret_cast<const uint32_t*>(&v[2 + i]);
#pragma unroll
          for (int k2 = 0; k2 < 2; ++k2) {
            const float4 sc = __ldg(ip + 2 * i + k2);
            const float2 h0 = __half22float2(*reinterpret_cast<const __half2*>(&hw[2 * k2])), h1 = __half22float2(*reinterpret_cast<const __half2*>(&hw[2 * k2 + 1]));
            const float2 l0 = __half22float2(*reinterpret_cast<const __half2*>(&lw[2 * k2])), l1 = __half22float2(*reinterpret_cast<const __half2*>(&lw[2 * k2 + 1]));
            float* ac = &a[8 * i + 4 * k2];
            ac[0] = fmaf(fmaf(l0.x, PS_LO_INV, h0.x), sc.x, ac[0]); ac[1] = fmaf(fmaf(l0.y, PS_LO_INV, h0.y), sc.y, ac[1]);
            ac[2] = fmaf(fmaf(l1.x, PS_LO_INV, h1.x), sc.z, ac[2]); ac[3] = fmaf(fmaf(l1.y, PS_LO_INV, h1.y), sc.w, ac[3]);
          }
        }
#else
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 hv = *reinterpret_cast<const float4*>(&v[i]), lv = *reinterpret_cast<const float4*>(&v[4 + i]);
          const float4 sc = __ldg(ip + i);
          a[4 * i + 0] = fmaf(hv.x + lv.x, sc.x, a[4 * i + 0]);
          a[4 * i + 1] = fmaf(hv.y + lv.y, sc.y, a[4 * i + 1]);
          a[4 * i + 2] = fmaf(hv.z + lv.z, sc.z, a[4 * i + 2]);
          a[4 * i + 3] = fmaf(hv.w + lv.w, sc.w, a[4 * i + 3]);
        }
#endif
      };
      // this lane's residual row chunk of group g: from the warp's staging buffer gi (TMA-loaded), or a plain load when the
      // warp owns more groups than staging buffers; rows on the zero border read as zeros
      auto res_fetch = [&](int gi, int g, uint4 (&v)[NV]) {
        const uint32_t sw = (CHB == 128) ? (lane & 7u) : ((lane >> 1) & 3u);   // swizzle of this lane's staged row
        if (gi < p.nstg) {
          const uint32_t srow = st_base + (uint32_t)gi * 32u * CHB + (uint32_t)lane * CHB;
#pragma unroll
          for (int i = 0; i < NV; ++i)
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[i].x), "=r"(v[i].y), "=r"(v[i].z), "=r"(v[i].w)
                         : "r"(srow + (((uint32_t)i ^ sw) << 4)) : "memory");
        } else if ((interior >> (g / gpm)) & 1u) {       // more groups than staging buffers: plain loads
          const long long m = m0 + (g / gpm) * 128 + q * 32 + lane + p.res_row_off;
          const uint4* rp = reinterpret_cast<const uint4*>(p.res + m * rowF + (p.res_coff + n0 / 16 + g % gpm) * CF);
#pragma unroll
          for (int i = 0; i < NV; ++i) v[i] = __ldg(rp + i);
        } else {
#pragma unroll
          for (int i = 0; i < NV; ++i) v[i] = make_uint4(0u, 0u, 0u, 0u);
        }
      };
      // same without the per-channel scaling: a += float(chunk)
      auto res_add16_unit = [&](float (&a)[16], const uint4 (&v)[NV]) {
#if PE_FP16
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const uint32_t* hw = reinterpret_cast<const uint32_t*>(&v[i]);
          const uint32_t* lw = reinterpret_cast<const uint32_t*>(&v[2 + i]);
#pragma unroll
          for (int k2 = 0; k2 < 4; ++k2) {
            const float2 h0 = __half22float2(*reinterpret_cast<const __half2*>(&hw[k2])), l0 = __half22float2(*reinterpret_cast<const __half2*>(&lw[k2]));
            a[8 * i + 2 * k2] += fmaf(l0.x, PS_LO_INV, h0.x); a[8 * i + 2 * k2 + 1] += fmaf(l0.y, PS_LO_INV, h0.y);
          }
        }
#else
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 hv = *reinterpret_cast<const float4*>(&v[i]), lv = *reinterpret_cast<const float4*>(&v[4 + i]);
          a[4 * i + 0] += hv.x + lv.x; a[4 * i + 1] += hv.y + lv.y; a[4 * i + 2] += hv.z + lv.z; a[4 * i + 3] += hv.w + lv.w;
        }
#endif
      };
      EPI_TICK(e_tc)
      // SETS = 2: the drains of a tile start only after the other set has finished those of the previous tile.  A parity wait is
      // only meaningful for a waiter that is at most one phase behind: without this hand-over a set would test a `main` barrier
      // whose earlier phases (drained by the other set) it never observed and could pass on a stale phase.
      if (SETS == 2) {
        if (eset == 1) mbar_wait_relaxed(bar_turn, mytl & 1u, p.poll_ns);
        else if (mytl > 0) mbar_wait_relaxed(bar_turn + 8, (mytl - 1u) & 1u, p.poll_ns);
      }
      for (int d = 0; ROLE != 2 && d < ndrain; ++d) {
        mbar_wait_relaxed(bar_main_full + 8 * dg, dgp, p.poll_ns);
        tc_fence_after();
        EPI_TICK(e_wm)
        // one 16-column group per TMEM wait: tcgen05.ld is throughput-bound (64 clk per x16 load, tools/tmem_bench.cu), so
        // batching the waits buys nothing and costs 16 live registers.  Accumulators stay in the SCALED domain (weights
        // were multiplied by 2^k per channel); the residual is brought into that domain when it is added and the final
        // phase multiplies by 2^-k: all exact
        // (drain warps of the split epilogue, two per quarter: the loads of up to three groups are issued back to back and
        // waited for once -- with only two warps per quarter the per-load latency, not the TMEM read port, paces a drain)
        constexpr int LB = ROLE == 1 ? ((NGH <= 3 && SETS == 3) ? NGH : 2) : 1;      // (96 registers with 20 warps: two at a time)
#pragma unroll
        for (int g0 = 0; g0 < NGH; g0 += LB) {
          uint32_t r[LB][16];
#pragma unroll
          for (int b = 0; b < LB; ++b)
            if (g0 + b < NGH && PARTS * (g0 + b) + half < NG) tc_ld16_nowait(t_lane + dg * GC + (PARTS * (g0 + b) + half) * 16, r[b]);
          tc_wait_ld();
#pragma unroll
          for (int b = 0; b < LB; ++b) {
            const int gi = g0 + b;
            if (gi >= NGH || PARTS * gi + half >= NG) continue;
            if (d == 0) {
#pragma unroll
              for (int i = 0; i < 16; ++i) acc[gi % NACC][i] = __uint_as_float(r[b][i]);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) acc[gi % NACC][i] += __uint_as_float(r[b][i]);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (CG == 2) mbar_arrive_cluster(main_empty0 + 8 * dg); else mbar_arrive(bar_main_empty + 8 * dg); }
        if (++dg == NMAIN) { dg = 0; dgp ^= 1u; }
        EPI_TICK(e_dr)
        if (d == 0 && res_lazy) { res_issue(tile, nsl); __syncwarp(); }
      }
      if (SETS == 2) {                                     // the other set drains the next tile's groups
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_turn + 8 * (uint32_t)eset);
        const uint32_t t = dg + (uint32_t)ndrain;
        dgp ^= (t / NMAIN) & 1u;
        dg = t % NMAIN;
      }
      const uint32_t interior_cur = interior;
      int tile_n = tile + ewstep, nsl_n = nsl;
      if (p.grp > 0) { if (w + ewstep < p.total_work) tc_work_item(p, w + ewstep, tile_n, nsl_n); }
      else { while (tile_n >= p.tiles_m) { tile_n -= p.tiles_m; ++nsl_n; } }
      // cross terms: committed by MMA warp Y at the end of the tile
      const uint32_t cbuf = tl & 1u;
      if (ROLE == 2) mbar_wait_relaxed(bar_sum + 8 * cbuf, (tl >> 1) & 1u, p.poll_ns);       // the drain warps' sums are in corr buffer cbuf
      else mbar_wait_relaxed(bar_corr_full + 8 * cbuf, (tl >> 1) & 1u, p.poll_ns);
      tc_fence_after();
      EPI_TICK(e_wc)
      if (ROLE != 2) {
#pragma unroll
        for (int gi = 0; gi < NGH; ++gi) {
          if (PARTS * gi + half >= NG) continue;
          uint32_t r[16];
          tc_ld16(t_lane + (NMAIN + cbuf) * GC + (PARTS * gi + half) * 16, r);
          // the low operand halves carry a factor PS_LO_SCALE (fp16x2 build: 2^11, pe_common.cuh), hence so do the cross terms
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[gi % NACC][i] = fmaf(__uint_as_float(r[i]), PS_LO_INV, acc[gi % NACC][i]);
        }
      }
      if (ROLE == 1) {
        // hand-over: sums (drains + cross terms) back into the columns the cross terms came from; each warp rewrites only the
        // lanes / columns it has just read
#pragma unroll
        for (int gi = 0; gi < NGH; ++gi) {
          if (PARTS * gi + half >= NG) continue;
          tc_st16(t_lane + (NMAIN + cbuf) * GC + (PARTS * gi + half) * 16, acc[gi % NACC]);
        }
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_sum + 8 * cbuf);
        tile = tile_n; nsl = nsl_n;
        continue;
      }
      if (ROLE == 0) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (CG == 2) mbar_arrive_cluster(corr_empty0 + 8 * cbuf); else mbar_arrive(bar_corr_empty + 8 * cbuf); }
      }
      EPI_TICK(e_co)
      // residual (pre-activation form) LAST: every form of this kernel sums in the order drains, cross terms, residual, so that
      // all tilings / epilogue organisations give identical bits
      if (ROLE == 0 && p.res && !p.res_post) {
        mbar_wait_relaxed(res_bar, mytl & 1u, p.poll_ns);
#pragma unroll
        for (int gi = 0; gi < NGH; ++gi) {
          const int g = PARTS * gi + half;
          if (g >= NG) continue;
          uint4 v[NV];
          res_fetch(gi, g, v);
          res_add16(acc[gi % NACC], v, g);
        }
        __syncwarp();
        EPI_TICK(e_rs)
      }
      if (ROLE == 2 && half >= NG) {                       // (a finalize warp without a group still releases the corr buffer)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (CG == 2) mbar_arrive_cluster(corr_empty0 + 8 * cbuf); else mbar_arrive(bar_corr_empty + 8 * cbuf); }
      }
      if (p.res && (p.res_post || ROLE == 2)) mbar_wait_relaxed(res_bar, mytl & 1u, p.poll_ns);
      // ---- bias / activation / split / store (the MMA warps are already on the next tile)
#pragma unroll
      for (int gi = 0; gi < NGH; ++gi) {
        const int g = PARTS * gi + half;
        if (g >= NG) continue;
        const int mt = g / gpm, c0 = (g % gpm) * 16;
        float (&ag)[16] = acc[gi % NACC];
        if (ROLE == 2) {                                   // finalize warp: this group's sums from the hand-over columns, then the residual
          uint32_t r[16];
          tc_ld16(t_lane + (NMAIN + cbuf) * GC + g * 16, r);
#pragma unroll
          for (int i = 0; i < 16; ++i) ag[i] = __uint_as_float(r[i]);
          if (gi == (NG - 1 - half) / PARTS) {             // this warp's last read of the corr buffer: warp Y may start the tile after next in it
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (CG == 2) mbar_arrive_cluster(corr_empty0 + 8 * cbuf); else mbar_arrive(bar_corr_empty + 8 * cbuf); }
          }
          if (p.res && !p.res_post) {
            uint4 rv[NV];
            res_fetch(gi, g, rv);
            res_add16(ag, rv, g);
          }
        }
        uint4 ov[NV];                                      // one staged row: [hi.. | lo..]
        if (!((interior_cur >> mt) & 1u)) {
#pragma unroll
          for (int i = 0; i < NV; ++i) ov[i] = make_uint4(0u, 0u, 0u, 0u);
        } else {
          const float4* bp = reinterpret_cast<const float4*>(p.bias + n0 + c0);
          const float4* sp = reinterpret_cast<const float4*>(p.scale + n0 + c0);
          float v[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 b4 = __ldg(bp + i), s4 = __ldg(sp + i);
            v[4 * i + 0] = fmaf(ag[4 * i + 0], s4.x, b4.x); v[4 * i + 1] = fmaf(ag[4 * i + 1], s4.y, b4.y);
            v[4 * i + 2] = fmaf(ag[4 * i + 2], s4.z, b4.z); v[4 * i + 3] = fmaf(ag[4 * i + 3], s4.w, b4.w);
          }
          if (p.act == 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
          }
#ifndef PE_TC_NO_SILU
          else if (p.act == 2) {                           // SiLU (YOLOX ConvModule): x * sigmoid(x) = x / (1 + exp(-x))
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __fdiv_rn(v[i], 1.0f + expf(-v[i]));
          } else if (p.act == 3) {                         // GELU (ViT MLP): x * 0.5 * (1 + erf(x / sqrt(2)))
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = v[i] * 0.5f * (1.0f + erff(v[i] * 0.70710678118654752440f));
          }
#endif
#ifndef PE_TC_NO_RESPOST
          if (p.res && p.res_post) {                         // residual after the activation (VideoPose3D: x = res + ReLU(bn(conv)))
            uint4 rv[NV];
            res_fetch(gi, g, rv);
            float one[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) one[i] = 0.f;
            res_add16_unit(one, rv);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += one[i];
          }
#endif
#if PE_FP16
          // split WITHOUT the saturating clamp of split4_h: a value beyond +-65504 converts to an fp16 infinity, which the integer
          // test below detects on the packed words (exponent field all ones <=> adding 0x0400 carries into bit 15) -- an
          // out-of-range activation raises the model's range flag instead of being clamped silently, and the test costs less than
          // the clamp did (measured: an fmax-based check took 2.5 % of the HRNet forward)
          uint2 h[4], l[4];
          uint32_t infbits = 0u;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const __half2 h0 = __floats2half2_rn(v[4 * i], v[4 * i + 1]), h1 = __floats2half2_rn(v[4 * i + 2], v[4 * i + 3]);
            const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
            const __half2 l0 = __floats2half2_rn((v[4 * i] - f0.x) * PS_LO_SCALE, (v[4 * i + 1] - f0.y) * PS_LO_SCALE);
            const __half2 l1 = __floats2half2_rn((v[4 * i + 2] - f1.x) * PS_LO_SCALE, (v[4 * i + 3] - f1.y) * PS_LO_SCALE);
            h[i].x = *reinterpret_cast<const uint32_t*>(&h0); h[i].y = *reinterpret_cast<const uint32_t*>(&h1);
            l[i].x = *reinterpret_cast<const uint32_t*>(&l0); l[i].y = *reinterpret_cast<const uint32_t*>(&l1);
            infbits |= ((h[i].x & 0x7C007C00u) + 0x04000400u) | ((h[i].y & 0x7C007C00u) + 0x04000400u);
          }
#ifndef PE_TC_NO_RANGECHECK
          if ((infbits & 0x80008000u) && p.flag) atomicOr(p.flag, 1u);
#endif
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
        EPI_TICK(e_f1)
        // the nstg buffers rotate: a store has nstg-1 group times to read its source.  Post-activation residual layers keep
        // group gi's residual in buffer gi until this point, so its output goes to that same buffer
        // (finalize warps consume group gi's residual from buffer gi right before this point as well)
        if (p.res && (p.res_post || ROLE == 2)) st_slot = (uint32_t)gi % (uint32_t)p.nstg;
        const uint32_t sbuf = st_base + st_slot * 32u * CHB;
        if (lane == 0 && !p.dstore) bulk_wait_read((p.res && (p.res_post || ROLE == 2)) ? (gi < p.nstg ? 7 : (ROLE == 2 ? p.nstg - 1 : 0)) : p.nstg - 1);   // the buffer about to be overwritten has been read by its store
        __syncwarp();
        EPI_TICK(e_f2)
        const uint32_t srow = sbuf + (uint32_t)lane * CHB;
        // SWIZZLE_128B: 16-byte chunk index ^= row & 7;  SWIZZLE_64B: chunk index ^= (row >> 1) & 3
        const uint32_t sw = (CHB == 128) ? (lane & 7u) : ((lane >> 1) & 3u);
#pragma unroll
        for (int i = 0; i < NV; ++i) st_shared_u4(srow + (((uint32_t)i ^ sw) << 4), ov[i]);
        if (p.dstore) {
          __syncwarp();
          EPI_TICK(e_f3)
          const long long mrow = m0 + mt * 128 + q * 32;
          char* const obase = reinterpret_cast<char*>(p.out) + (size_t)(p.out_coff + (n0 + c0) / 16) * CHB;
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            const uint32_t row = (uint32_t)i * (32u / NV) + (uint32_t)lane / NV, piece = (uint32_t)lane % NV;
            const uint32_t swr = (CHB == 128) ? (row & 7u) : ((row >> 1) & 3u);
            uint4 t;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w)
                         : "r"(sbuf + row * CHB + ((piece ^ swr) << 4)) : "memory");
            if (mrow + row < p.M) *reinterpret_cast<uint4*>(obase + (size_t)(mrow + row) * (size_t)p.out_rowF * 4 + piece * 16) = t;
          }
        } else {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        EPI_TICK(e_f3)
        if (lane == 0) {
          const long long mrow = m0 + mt * 128 + q * 32;
          if (mrow < p.M) tma_store_2d(&tmO, (p.out_coff + (n0 + c0) / 16) * CF, (int)mrow, sbuf);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        }
        if (++st_slot == (uint32_t)p.nstg) st_slot = 0;
        __syncwarp();
        EPI_TICK(e_f4)
      }
      tile = tile_n; nsl = nsl_n;
      EPI_TICK(e_fi)
      if (ROLE == 0 && p.res && !res_lazy && w + ewstep < p.total_work) res_issue(tile, nsl);   // next tile's residual chunks, one tile ahead
      __syncwarp();
      EPI_TICK(e_ri)
    }
#if PE_TC_PROFILE
    if (p.prof && (warp == 2 || (ROLE == 2 && warp == 10)) && lane == 0) {
      long long* o = p.prof + (size_t)blockIdx.x * 32 + (ROLE == 2 ? 24 : 16);
      if (ROLE == 2) { o[0] = e_wc; o[1] = e_f1; o[2] = e_f2; o[3] = e_f3; o[4] = e_f4; o[5] = clock64() - e_t0; o[6] = tl; o[7] = e_tc; } else
      o[0] = e_wm; o[1] = e_dr; o[2] = e_rs; o[3] = e_wc; o[4] = e_co; o[5] = e_fi; o[6] = e_ri; o[7] = e_tc; o[8] = clock64() - e_t0; o[9] = tl / ESETS; o[10] = e_f1; o[11] = e_f2; o[12] = e_f3; o[13] = e_f4;
    }
#endif
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // staged rows fully written before smem goes away
    };
    if constexpr (SPLIT) {
      if (warp < 10) epilogue(std::integral_constant<int, 1>{}); else epilogue(std::integral_constant<int, 2>{});
    } else {
      epilogue(std::integral_constant<int, 0>{});
    }
  }
  tc_fence_before();
  __syncwarp();
  if (CG == 2) cluster_sync_all();      // neither CTA may leave (or free TMEM) while the pair's MMAs / signals still target it
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

typedef void (*TcKernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const TcParams);
