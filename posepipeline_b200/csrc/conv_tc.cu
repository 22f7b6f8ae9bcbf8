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

struct TcParams {
  float* out;
  const float* res;
  const float* bias;
  long long M;     // rows (padded positions) of this launch
  int H, W, Hp, Wp;
  int nchunk, ntaps, Cout, NC, MT, TPS, SA, SB;
  int Rpad, RB, nbA, halo;
  int nsub, Nsub, nboxW, NCbox;
  int relu, tmem_cols, bo_mode;
  int tiles_m, total_work, dbuf;   // persistent schedule: work item w -> (tile = w % tiles_m, n-slice = w / tiles_m)
};

// ------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug must trap (launch failure) instead of hanging the GPU box
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("conv_tc: mbarrier timeout (block %d,%d thread %d bar %u parity %u)\n", blockIdx.x, blockIdx.y, threadIdx.x, bar, parity);
      __trap();
    }
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
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
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
  d |= (uint64_t)(1024 >> 4) << 32;               // stride-dim byte offset: next 8-row group
  d |= (uint64_t)1 << 46;                         // descriptor version (sm_100)
  // Measured on B200 (tests/tc_bringup.py): the 128B swizzle XOR is applied to ABSOLUTE shared-memory address bits,
  // so a descriptor that starts at an arbitrary row of a TMA-written window needs base_offset = 0; setting the
  // "swizzle phase" there (bo_mode 0, kept for the experiment) reads the wrong 16-byte chunks.
  if (bo_mode == 0) d |= (uint64_t)((saddr >> 7) & 7u) << 49;
  d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
  return d;
}

// ------------------------------------------------------------------------------------------ kernel
// Persistent: gridDim.x CTAs walk the work list; the TMA and MMA warps run ahead into the next tile while the
// epilogue warps drain the previous accumulators (double-buffered in TMEM when 4*MT*NC <= 512 columns).
// TMEM columns of buffer b: [main(mt=0..MT-1) | corr(mt=0..MT-1)], NC columns each.  hi*hi accumulates in `main`;
// the two small cross terms hi*lo + lo*hi accumulate in `corr`, so the tensor core's truncating FP32 accumulation
// only sees K/8 steps per accumulator (measured: error grows ~2^-24.7 per step) and the epilogue adds them in FP32.
__global__ void __launch_bounds__(192, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t a_bytes = (uint32_t)p.Rpad * 128u;
  const uint32_t b_bytes = (uint32_t)p.TPS * p.NC * 128u;
  const uint32_t sA = base;
  const uint32_t sB = sA + p.SA * a_bytes;
  const uint32_t sBar = sB + p.SB * b_bytes;       // 8-byte barriers
  const uint32_t bar_a_full = sBar, bar_a_empty = sBar + 8 * p.SA;
  const uint32_t bar_b_full = bar_a_empty + 8 * p.SA, bar_b_empty = bar_b_full + 8 * p.SB;
  const uint32_t bar_acc_full = bar_b_empty + 8 * p.SB;   // [2]
  const uint32_t bar_acc_empty = bar_acc_full + 16;       // [2]
  const uint32_t s_tmem = bar_acc_empty + 16;
  uint8_t* gen = smem_raw + (base - raw);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen + (s_tmem - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.SA; ++i) { mbar_init(bar_a_full + 8 * i, 1); mbar_init(bar_a_empty + 8 * i, 1); }
    for (int i = 0; i < p.SB; ++i) { mbar_init(bar_b_full + 8 * i, 1); mbar_init(bar_b_empty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_acc_full + 8 * i, 1); mbar_init(bar_acc_empty + 8 * i, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_tmem), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int ngroups = p.ntaps / p.TPS;
  const uint32_t buf_cols = 2u * p.MT * p.NC;

  if (warp == 0) {
    // ===================== TMA producer (one lane) =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmW) : "memory");
      uint32_t ita = 0, itb = 0;
      for (int w = blockIdx.x; w < p.total_work; w += gridDim.x) {
        const long long m0 = (long long)(w % p.tiles_m) * 128 * p.MT;
        const int n0 = (w / p.tiles_m) * p.NC;
        auto load_a = [&](int j) {
          const uint32_t sa = ita % p.SA, ph = (ita / p.SA) & 1u;
          ++ita;
          mbar_wait(bar_a_empty + 8 * sa, ph ^ 1u);
          mbar_expect_tx(bar_a_full + 8 * sa, a_bytes);
          for (int b = 0; b < p.nbA; ++b)
            tma_load_2d(sA + sa * a_bytes + (uint32_t)b * p.RB * 128u, &tmA, j * 32, (int)(m0 - p.halo + (long long)b * p.RB),
                        bar_a_full + 8 * sa);
        };
        load_a(0);
        for (int j = 0; j < p.nchunk; ++j) {
          for (int g = 0; g < ngroups; ++g) {
            const uint32_t sb = itb % p.SB, ph = (itb / p.SB) & 1u;
            ++itb;
            mbar_wait(bar_b_empty + 8 * sb, ph ^ 1u);
            mbar_expect_tx(bar_b_full + 8 * sb, b_bytes);
            for (int t = 0; t < p.TPS; ++t) {
              const int tap = g * p.TPS + t;
              tma_load_2d(sB + sb * b_bytes + (uint32_t)(t * p.NC) * 128u, &tmW, 0, (tap * p.nchunk + j) * p.Cout + n0,
                          bar_b_full + 8 * sb);
            }
            if (g == 0 && j + 1 < p.nchunk) load_a(j + 1);   // next activation window right behind the first weights
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: warp-uniform control flow, one elected lane issues =====================
    const uint32_t leader = elect_one();
    // instruction descriptor: D=F32, A=B=TF32, K-major both, N = NC, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.NC >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t a_desc0 = umma_desc(sA, p.bo_mode), b_desc0 = umma_desc(sB, p.bo_mode);
    uint32_t ita = 0, itb = 0, tl = 0;
    for (int w = blockIdx.x; w < p.total_work; w += gridDim.x, ++tl) {
      const uint32_t buf = p.dbuf ? (tl & 1u) : 0u;
      const uint32_t use = p.dbuf ? (tl >> 1) : tl;
      mbar_wait(bar_acc_empty + 8 * buf, (use & 1u) ^ 1u);        // epilogue has drained this accumulator buffer
      tc_fence_after();
      const uint32_t d_main = tmem_base + buf * buf_cols;
      const uint32_t d_corr = d_main + (uint32_t)(p.MT * p.NC);
      for (int j = 0; j < p.nchunk; ++j) {
        const uint32_t sa = ita % p.SA;
        mbar_wait(bar_a_full + 8 * sa, (ita / p.SA) & 1u);
        ++ita;
        tc_fence_after();
        const uint64_t a_slot = a_desc0 + (uint64_t)((sa * a_bytes) >> 4);
        for (int g = 0; g < ngroups; ++g) {
          const uint32_t sb = itb % p.SB;
          mbar_wait(bar_b_full + 8 * sb, (itb / p.SB) & 1u);
          ++itb;
          tc_fence_after();
          const uint64_t b_slot = b_desc0 + (uint64_t)((sb * b_bytes) >> 4);
          for (int t = 0; t < p.TPS; ++t) {
            const int tap = g * p.TPS + t;
            const uint32_t shift = (p.ntaps == 9) ? (uint32_t)((tap / 3) * p.Wp + (tap % 3)) : 0u;
            const uint64_t b_tap = b_slot + (uint64_t)(t * p.NC * 8);          // 128 B per row = 8 x 16 B
            const uint32_t fresh = (j == 0 && tap == 0) ? 0u : 1u;
            for (int mt = 0; mt < p.MT; ++mt) {
              const uint64_t a_mt = a_slot + (uint64_t)((mt * 128 + shift) * 8);
              const uint32_t dm = d_main + (uint32_t)(mt * p.NC), dc = d_corr + (uint32_t)(mt * p.NC);
              if (leader) {
                // k-step 0: floats 0..7 of hi (bytes 0..31) and of lo (bytes 64..95); k-step 1: +32 B
                tc_mma_tf32(dm, a_mt, b_tap, idesc, fresh);              // hi * hi
                tc_mma_tf32(dc, a_mt, b_tap + 4, idesc, fresh);          // hi * lo
                tc_mma_tf32(dc, a_mt + 4, b_tap, idesc, 1u);             // lo * hi
                tc_mma_tf32(dm, a_mt + 2, b_tap + 2, idesc, 1u);
                tc_mma_tf32(dc, a_mt + 2, b_tap + 6, idesc, 1u);
                tc_mma_tf32(dc, a_mt + 6, b_tap + 2, idesc, 1u);
              }
            }
          }
          if (leader) tc_commit(bar_b_empty + 8 * sb);     // weights stage free once these MMAs retire
          __syncwarp();
        }
        if (leader) tc_commit(bar_a_empty + 8 * sa);       // activation window free
        __syncwarp();
      }
      if (leader) tc_commit(bar_acc_full + 8 * buf);       // accumulators of this tile complete
      __syncwarp();
    }
  } else {
    // ===================== epilogue (warps 2..5; TMEM lane quarter = warp % 4) =====================
    const int q = warp & 3;
    const int rowF = 2 * p.Cout;
    uint32_t tl = 0;
    for (int w = blockIdx.x; w < p.total_work; w += gridDim.x, ++tl) {
      const long long m0 = (long long)(w % p.tiles_m) * 128 * p.MT;
      const int n0 = (w / p.tiles_m) * p.NC;
      const uint32_t buf = p.dbuf ? (tl & 1u) : 0u;
      const uint32_t use = p.dbuf ? (tl >> 1) : tl;
      mbar_wait(bar_acc_full + 8 * buf, use & 1u);
      tc_fence_after();
      const uint32_t t_main = tmem_base + buf * buf_cols + ((uint32_t)(q * 32) << 16);
      const uint32_t t_corr = t_main + (uint32_t)(p.MT * p.NC);
      for (int mt = 0; mt < p.MT; ++mt) {
        const long long m = m0 + mt * 128 + q * 32 + lane;
        const bool valid = m < p.M;
        bool interior = false;
        if (valid) {
          const int r = (int)(m % ((long long)p.Hp * p.Wp));
          const int py = r / p.Wp, px = r % p.Wp;
          interior = py >= 1 && py <= p.H && px >= 1 && px <= p.W;
        }
        float* orow = p.out + m * rowF + 2 * n0;
        const float* rrow = p.res ? p.res + m * rowF + 2 * n0 : nullptr;
        for (int c0 = 0; c0 < p.NC; c0 += 16) {
          uint32_t r[16], rc[16];
          tc_ld16_nowait(t_main + (uint32_t)(mt * p.NC + c0), r);
          tc_ld16_nowait(t_corr + (uint32_t)(mt * p.NC + c0), rc);
          float4 rh[4], rl[4];
          const bool do_res = rrow && valid && interior;
          if (do_res) {
            const float4* rp = reinterpret_cast<const float4*>(rrow + 2 * c0);
#pragma unroll
            for (int i = 0; i < 4; ++i) { rh[i] = __ldg(rp + i); rl[i] = __ldg(rp + 4 + i); }
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (!valid) continue;
          float4* o = reinterpret_cast<float4*>(orow + 2 * c0);
          if (!interior) {
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = z;
            continue;
          }
          const float4* bp = reinterpret_cast<const float4*>(p.bias + n0 + c0);
          float v[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 b4 = __ldg(bp + i);
            v[4 * i + 0] = (__uint_as_float(r[4 * i + 0]) + __uint_as_float(rc[4 * i + 0])) + b4.x;
            v[4 * i + 1] = (__uint_as_float(r[4 * i + 1]) + __uint_as_float(rc[4 * i + 1])) + b4.y;
            v[4 * i + 2] = (__uint_as_float(r[4 * i + 2]) + __uint_as_float(rc[4 * i + 2])) + b4.z;
            v[4 * i + 3] = (__uint_as_float(r[4 * i + 3]) + __uint_as_float(rc[4 * i + 3])) + b4.w;
          }
          if (do_res) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              v[4 * i + 0] += rh[i].x + rl[i].x; v[4 * i + 1] += rh[i].y + rl[i].y;
              v[4 * i + 2] += rh[i].z + rl[i].z; v[4 * i + 3] += rh[i].w + rl[i].w;
            }
          }
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float4 hi, lo;
            split4(make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]), hi, lo);
            o[i] = hi;
            o[4 + i] = lo;
          }
        }
      }
      // all tcgen05.ld of this buffer have completed (wait::ld above): hand the accumulators back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc_empty + 8 * buf);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------ host side
struct TcConvPlan {
  CUtensorMap tmA, tmW;
  TcParams p;
  int rows_per_img;
  size_t smem;
  int ns, num_sms;
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
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

cudaError_t tc_conv_plan_create(TcConvPlan** out, const float* in, float* outp, const float* res, const float* wtc,
                                const float* bias, int Cin, int Cout, int ks, int relu, int H, int W, int max_img) {
  if (env_int("PE_TC_DISABLE", 0)) return cudaErrorNotSupported;
  if ((ks != 1 && ks != 3) || Cin % 16 || Cout % 16 || Cout > 512) return cudaErrorNotSupported;
  const int Hp = H + 2, Wp = W + 2, ntaps = ks * ks, nchunk = Cin / 16;
  const int halo = ks == 3 ? Wp + 1 : 0;
  const long long Mmax = (long long)max_img * Hp * Wp;
  const size_t smem_cap = 200 * 1024;
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
      if (2 * MT * NC > 512) continue;
      if (force_mt && MT != force_mt) continue;
      for (int dbuf = 1; dbuf >= 0; --dbuf) {
        if (dbuf && 4 * MT * NC > 512) continue;
        if (force_dbuf >= 0 && dbuf != force_dbuf) continue;
        TcParams p{};
        p.NC = NC; p.MT = MT; p.nchunk = nchunk; p.ntaps = ntaps; p.Cout = Cout; p.halo = halo; p.dbuf = dbuf;
        p.TPS = (ntaps == 9 && NC <= 64) ? 3 : 1;
        const int R = 128 * MT + 2 * halo;
        p.nbA = (R + 255) / 256;
        p.Rpad = ((R + 8 * p.nbA - 1) / (8 * p.nbA)) * (8 * p.nbA);
        p.RB = p.Rpad / p.nbA;
        p.SA = 2;
        p.nsub = 1; p.Nsub = NC; p.nboxW = 1; p.NCbox = NC;
        const size_t a_bytes = (size_t)p.Rpad * 128, b_bytes = (size_t)p.TPS * NC * 128;
        int SB = 6;
        while (SB > 2 && p.SA * a_bytes + SB * b_bytes + 4096 > smem_cap) --SB;
        if (p.SA * a_bytes + SB * b_bytes + 4096 > smem_cap) continue;
        p.SB = SB;
        int cols = 32;
        while (cols < (dbuf ? 2 : 1) * 2 * MT * NC) cols <<= 1;
        p.tmem_cols = cols;
        p.tiles_m = (int)((Mmax + 128LL * MT - 1) / (128LL * MT));
        p.total_work = p.tiles_m * ns;
        const int ctas = p.total_work < num_sms ? p.total_work : num_sms;
        const double items = (double)((p.total_work + ctas - 1) / ctas);
        // clocks per work item: tf32 MMA at 2048 MAC/clk/SM, but never faster than the operands can be read from
        // shared memory (128 B/clk: A 4 KB + B NC*32 B per MMA) or fetched from L2 (~32 B/clk/SM)
        const double n_mma = 6.0 * ntaps * nchunk * MT;
        const double mma = n_mma * std::max(NC / 2.0, (4096.0 + NC * 32.0) / 128.0);
        const double bytes = (double)nchunk * (a_bytes + (double)ntaps * NC * 128);
        const double epi = (double)MT * (NC / 16) * 260.0 + 1500.0;
        const double item = std::max(mma, bytes / 32.0) + (dbuf ? 0.0 : epi);
        const double t = items * std::max(item, dbuf ? epi : 0.0) + 4000.0;
        if (t < best) { best = t; bp = p; bsmem = p.SA * a_bytes + p.SB * b_bytes + 4096; bns = ns; }
      }
    }
  }
  if (best >= 1e30) return cudaErrorNotSupported;
  TcConvPlan* pl = new TcConvPlan();
  pl->p = bp;
  pl->p.out = outp; pl->p.res = res; pl->p.bias = bias;
  pl->p.H = H; pl->p.W = W; pl->p.Hp = Hp; pl->p.Wp = Wp; pl->p.relu = relu;
  pl->p.bo_mode = env_int("PE_TC_BO_MODE", 1);
  pl->rows_per_img = Hp * Wp;
  pl->smem = bsmem;
  pl->ns = bns;
  CUresult r1 = encode_2d(&pl->tmA, in, (uint64_t)2 * Cin, (uint64_t)Mmax, (uint64_t)2 * Cin * 4, 32, (uint32_t)bp.RB);
  CUresult r2 = encode_2d(&pl->tmW, wtc, 32, (uint64_t)ntaps * nchunk * Cout, 128, 32, (uint32_t)bp.NCbox);
  if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) {
    fprintf(stderr, "conv_tc: cuTensorMapEncodeTiled failed (%d, %d) Cin=%d Cout=%d RB=%d NCbox=%d\n", (int)r1, (int)r2, Cin, Cout, bp.RB, bp.NCbox);
    delete pl;
    return cudaErrorInvalidValue;
  }
  static size_t attr_set = 0;
  if (pl->smem > attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(210 * 1024));
    if (e != cudaSuccess) { delete pl; return e; }
    attr_set = 210 * 1024;
  }
  pl->num_sms = num_sms;
  if (env_int("PE_TC_VERBOSE", 0))
    fprintf(stderr, "conv_tc plan: Cin=%d Cout=%d ks=%d %dx%d  MT=%d NS=%d NC=%d dbuf=%d TPS=%d SA=%d SB=%d Rpad=%d RB=%d smem=%zu tmem=%d work=%d\n", Cin, Cout, ks, H, W,
            bp.MT, bns, bp.NC, bp.dbuf, bp.TPS, bp.SA, bp.SB, bp.Rpad, bp.RB, pl->smem, bp.tmem_cols, bp.total_work);
  *out = pl;
  return cudaSuccess;
}

void tc_conv_plan_destroy(TcConvPlan* plan) { delete plan; }

cudaError_t tc_conv_launch(TcConvPlan* pl, int nimg, cudaStream_t st) {
  TcParams p = pl->p;
  p.M = (long long)nimg * pl->rows_per_img;
  p.tiles_m = (int)((p.M + 128LL * p.MT - 1) / (128LL * p.MT));
  p.total_work = p.tiles_m * pl->ns;
  const unsigned grid = (unsigned)(p.total_work < pl->num_sms ? p.total_work : pl->num_sms);
  conv_tc_kernel<<<grid, 192, pl->smem, st>>>(pl->tmA, pl->tmW, p);
  return cudaGetLastError();
}
