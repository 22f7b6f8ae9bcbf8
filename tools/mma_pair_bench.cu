// Microbenchmark: tcgen05.mma.cta_group::2 (CTA pair, M = 256 over the two SMs of a TPC, each CTA holds half of B) issue/execute
// rate on B200 for kind::tf32 and kind::f16, various N -- against tools/mma_bench.cu (cta_group::1, M = 128).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_pair_bench tools/mma_pair_bench.cu && ./mma_pair_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n.reg .b32 rx;\n.reg .pred px;\nelect.sync rx|px, 0xFFFFFFFF;\n@px mov.s32 %0, 1;\n}\n" : "+r"(pred));
  return pred;
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16; d |= (uint64_t)(1024 >> 4) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61;
  return d;
}
template <int KIND>  // 0 tf32, 1 bf16
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (KIND == 0)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int KIND>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) bench(int N, int iters, int same_acc, int a_shift_rows, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw), base = (raw + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 64 * 1024, sBar = sB + 64 * 1024, sT = sBar + 16;
  volatile uint32_t* tslot = (volatile uint32_t*)(smem_raw + (sT - raw));
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 32 * 1024; i += 128) ((float*)(smem_raw + (base - raw)))[i] = 0.f;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sBar)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sT), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tslot, 0);
  if (warp == 1 && (blockIdx.x & 1) == 0) {
    const uint32_t fmt = KIND == 0 ? 2u : 1u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((256u >> 4) << 24);
    const uint64_t a0 = umma_desc(sA), b0 = umma_desc(sB);
    long long t0 = clock64();
    for (int it = 0; it < iters; it += 8) {
      if (elect_one()) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint32_t d = tmem + (same_acc ? 0u : (uint32_t)((u & 1) * 256));
          mma<KIND>(d, a0 + (uint64_t)((u & 3) * 2 + a_shift_rows * 8 * (u >> 2)), b0 + (uint64_t)((u & 3) * 2), idesc, 1u);
        }
      }
      __syncwarp();
    }
    if (elect_one()) asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(sBar), "h"((uint16_t)1) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(sBar), "r"(0u) : "memory");
    long long t1 = clock64();
    if (threadIdx.x == 32) out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncwarp();
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  long long* d_out; cudaMalloc(&d_out, 148 * sizeof(long long));
  long long h[148];
  const int smem = 64 * 1024 * 2 + 2048;
  cudaFuncSetAttribute(bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 4096;
  printf("cta_group::2, M=256\nkind  N  same_acc a_shift grid  cycles/MMA  MAC/clk/SM\n");
  for (int kind = 0; kind < 2; ++kind)
    for (int N : {32, 48, 64, 96, 128, 192, 256})
      for (int same : {1, 0})
        for (int shift : {0, 3})
          for (int grid : {2, 148}) {
            if ((shift || !same) && grid == 2) continue;
            if (kind == 0) bench<0><<<grid, 128, smem>>>(N, iters, same, shift, d_out);
            else bench<1><<<grid, 128, smem>>>(N, iters, same, shift, d_out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(h, d_out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
            double mx = 0; for (int i = 0; i < grid; i += 2) mx = h[i] > mx ? h[i] : mx;
            const double cyc = mx / iters, K = kind == 0 ? 8 : 16;
            printf("%s %4d %d %d %4d  %8.1f  %8.0f\n", kind ? "bf16" : "tf32", N, same, shift, grid, cyc, 128.0 * N * K / cyc);   // MAC/clk per SM (each SM does 128 rows)
          }
  return 0;
}
