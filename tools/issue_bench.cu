// Microbenchmark: how deep is the tcgen05.mma queue?  One warp issues a burst of B MMAs (N=48, kind::f16, SWIZZLE_64B operands),
// then idles for G clocks (a clock64 spin), repeatedly.  If MMAs queue deeply, time/iter = max(B*T_mma, issue + G); if the queue
// is shallow (issue blocks until the pipe accepts), time/iter = B*T_mma + G.  Also times fence / elect / mbarrier test_wait.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n.reg .b32 rx;\n.reg .pred px;\nelect.sync rx|px, 0xFFFFFFFF;\n@px mov.s32 %0, 1;\n}\n" : "+r"(pred));
  return pred;
}
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t row_bytes, uint32_t layout) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16; d |= (uint64_t)((8 * row_bytes) >> 4) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)layout << 61;
  return d;
}
__device__ __forceinline__ void mma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
// collector usage for the A operand: 1 = fill (keep A in the collector buffer), 2 = lastuse (reuse it), 0 = plain
template <int CU>
__device__ __forceinline__ void mma_f16_cu(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (CU == 1)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else if (CU == 2)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__global__ void __launch_bounds__(128, 1) issue_bench(int N, int burst, int gap, int iters, int mode, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw), base = (raw + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 64 * 1024, sBar = sB + 64 * 1024, sT = sBar + 64;
  volatile uint32_t* tslot = (volatile uint32_t*)(smem_raw + (sT - raw));
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 32 * 1024; i += 128) ((float*)(smem_raw + (base - raw)))[i] = 0.f;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sBar));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sBar + 8));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sT), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tslot, 0);
  if (mode == 7 && (warp == 1 || (warp == 2 && burst < 0))) {
    // triples (A_hi x W_hi, A_hi x W_lo, A_lo x W_hi) as the conv kernel issues them per tap; gap != 0: the first two share A through
    // the collector buffer (fill / lastuse).  burst < 0: two warps, each its own accumulators (split by 128-row sub-tile).
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t a0 = desc(sA, 64, 4), b0 = desc(sB, 64, 4);
    const uint32_t bar = sBar + 16 + 8 * (warp - 1);
    if ((threadIdx.x & 31) == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    const uint32_t dm = tmem + (warp - 1) * 256, dc = dm + 128;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int u = 0; u < 9; ++u) {
          const uint64_t a = a0 + (uint64_t)(u * 4 + (warp - 1) * 512), b = b0 + (uint64_t)(u * 192);
          if (gap) {
            mma_f16_cu<1>(dm, a, b, idesc, 1u);
            mma_f16_cu<2>(dc, a, b + 2, idesc, 1u);
          } else {
            mma_f16_cu<0>(dm, a, b, idesc, 1u);
            mma_f16_cu<0>(dc, a, b + 2, idesc, 1u);
          }
          mma_f16_cu<0>(dc, a + 2, b, idesc, 1u);
        }
      }
      __syncwarp();
    }
    if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    while (!mbar_try(bar, 0u)) {}
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[(warp - 1)] = t1 - t0;
  } else if (mode == 6 && (warp == 1 || warp == 2)) {
    // two issuing warps: disjoint accumulators, same operands; each: burst then gap
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t a0 = desc(sA, 64, 4), b0 = desc(sB, 64, 4);
    const uint32_t bar = sBar + 16 + 8 * (warp - 1);
    if ((threadIdx.x & 31) == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
        for (int u = 0; u < burst; ++u) mma_f16(tmem + (warp - 1) * 256 + (u & 1) * 128, a0 + (uint64_t)((u % 9) * 4), b0 + (uint64_t)((u % 3) * 2), idesc, 1u);
      }
      __syncwarp();
      const long long c1 = clock64();
      while (clock64() - c1 < gap) {}
    }
    if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    while (!mbar_try(bar, 0u)) {}
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[(warp - 1)] = t1 - t0;
  } else if (warp == 1) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t a0 = desc(sA, 64, 4), b0 = desc(sB, 64, 4);
    long long t0 = clock64();
    long long c_issue = 0;
    if (mode == 0) {
      for (int it = 0; it < iters; ++it) {
        const long long c0 = clock64();
        if (elect_one()) {
          for (int u = 0; u < burst; ++u) mma_f16(tmem + (u & 1) * 256, a0 + (uint64_t)((u % 9) * 4), b0 + (uint64_t)((u % 3) * 2), idesc, 1u);
        }
        __syncwarp();
        const long long c1 = clock64();
        c_issue += c1 - c0;
        while (clock64() - c1 < gap) {}
      }
    } else if (mode == 1) {          // cost of tcgen05.fence::after_thread_sync
      for (int it = 0; it < iters; ++it) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    } else if (mode == 2) {          // elect.sync
      uint32_t s = 0;
      for (int it = 0; it < iters; ++it) { s += elect_one(); __syncwarp(); }
      if (s == 0xffffffffu) out[1] = s;
    } else if (mode == 3) {          // mbarrier test_wait (not satisfied)
      uint32_t s = 0;
      for (int it = 0; it < iters; ++it) s += mbar_try(sBar + 8, 0u);
      if (s == 0xffffffffu) out[1] = s;
    } else if (mode == 4) {          // tcgen05.commit
      for (int it = 0; it < iters; ++it) {
        if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(sBar + 8) : "memory");
        __syncwarp();
      }
    } else if (mode == 5) {          // clock64 pair
      long long s = 0;
      for (int it = 0; it < iters; ++it) { const long long c = clock64(); s += clock64() - c; }
      if (s == 1) out[1] = s;
    }
    if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(sBar) : "memory");
    while (!mbar_try(sBar, 0u)) {}
    long long t1 = clock64();
    if (threadIdx.x == 32) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = c_issue; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}
int main() {
  long long* d_out; cudaMalloc(&d_out, 148 * 2 * sizeof(long long));
  long long h[2];
  const int smem = 64 * 1024 * 2 + 2048;
  cudaFuncSetAttribute(issue_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 512;
  printf("N burst gap : cycles/iter  issue-cycles/iter\n");
  for (int N : {48, 96, 256})
    for (int burst : {6, 18, 54})
      for (int gap : {0, 100, 200, 400, 800, 1600}) {
        issue_bench<<<1, 128, smem>>>(N, burst, gap, iters, 0, d_out);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
        cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost);
        printf("%3d %2d %4d : %8.1f %8.1f\n", N, burst, gap, (double)h[0] / iters, (double)h[1] / iters);
      }
  printf("two issuing warps -- N burst(per warp) gap : cycles/iter warp1 warp2\n");
  for (int N : {48, 96})
    for (int burst : {3, 6, 9, 18})
      for (int gap : {0, 200, 400, 800}) {
        issue_bench<<<1, 128, smem>>>(N, burst, gap, iters, 6, d_out);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
        cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost);
        printf("%3d %2d %4d : %8.1f %8.1f\n", N, burst, gap, (double)h[0] / iters, (double)h[1] / iters);
      }
  printf("collector reuse of A (9 taps x [hi*hi, hi*lo, lo*hi]) -- N warps reuse : cycles per tap-triple (per warp)\n");
  for (int N : {48, 64, 96, 128})
    for (int nw : {1, 2})
      for (int reuse : {0, 1}) {
        issue_bench<<<1, 128, smem>>>(N, nw == 2 ? -1 : 1, reuse, iters, 7, d_out);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost);
        printf("%3d %d %d : %8.1f %8.1f\n", N, nw, reuse, (double)h[0] / iters / 9, nw == 2 ? (double)h[1] / iters / 9 : 0.0);
      }
  const char* names[] = {"", "tcgen05.fence::after", "elect.sync+syncwarp", "mbarrier.test_wait", "elect+tcgen05.commit", "clock64 pair"};
  for (int mode = 1; mode <= 5; ++mode) {
    issue_bench<<<1, 128, smem>>>(48, 0, 0, 4096, mode, d_out);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
    cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost);
    printf("%-24s %8.1f cycles each\n", names[mode], (double)h[0] / 4096);
  }
  return 0;
}
