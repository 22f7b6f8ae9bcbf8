// Microbenchmark: tcgen05.ld / tcgen05.st throughput on B200 (4 warps, one per TMEM lane quarter), and the
// f16 MMA rate with a SWIZZLE_32B "stacked" B tile.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bench tools/tmem_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define LD16(r, a) asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
  : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(a) : "memory")
#define LD32(r, a) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
  : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), \
    "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(a) : "memory")
#define ST16(a, z) asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(a), "r"(z) : "memory")

// mode 0: x16 ld + wait each; 1: 3 x x16 ld then one wait; 2: x32 ld + wait; 3: x16 st + wait::st; 4: x16 ld+wait + 16 FADD
__global__ void __launch_bounds__(128, 1) tmem_bench(int mode, int iters, int nwarps, long long* out, float* sink) {
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tslot + ((uint32_t)(warp * 32) << 16);
  float acc[48];
  for (int i = 0; i < 48; ++i) acc[i] = 0.f;
  // initialise the columns we read
  for (int c = 0; c < 512; c += 16) ST16(tmem + c, 0u);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  __syncthreads();
  long long t0 = clock64();
  if (warp < nwarps) {
    for (int it = 0; it < iters; ++it) {
      const uint32_t a = tmem + (uint32_t)((it & 7) * 48);
      if (mode == 0 || mode == 4) {
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          uint32_t r[16];
          LD16(r, a + g * 16);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[g * 16 + i] += __uint_as_float(r[i]);
        }
      } else if (mode == 1) {
        uint32_t r0[16], r1[16], r2[16];
        LD16(r0, a); LD16(r1, a + 16); LD16(r2, a + 32);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 16; ++i) { acc[i] += __uint_as_float(r0[i]); acc[16 + i] += __uint_as_float(r1[i]); acc[32 + i] += __uint_as_float(r2[i]); }
      } else if (mode == 2) {
        uint32_t r0[32], r1[16];
        LD32(r0, a); LD16(r1, a + 32);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] += __uint_as_float(r0[i]);
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[32 + i] += __uint_as_float(r1[i]);
      } else if (mode == 3) {
        ST16(a, 0u); ST16(a + 16, 0u); ST16(a + 32, 0u);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
    }
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < 48; ++i) s += acc[i];
  if (s == 12345.f) sink[0] = s;
  if ((threadIdx.x & 31) == 0) out[blockIdx.x * 4 + warp] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tslot), "r"(512u) : "memory");
}

// ---- f16 MMA with a stacked B tile (SWIZZLE_32B rows of 32 B, N rows) and the usual SWIZZLE_64B A window
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
// pattern 0: three N=NC MMAs (B rows 64 B, SWIZZLE_64B) -- today's kernel;  1: one N=2NC (B rows 32 B, SWIZZLE_32B) + one N=NC
__global__ void __launch_bounds__(128, 1) mma_pair_bench(int NC, int pattern, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw), base = (raw + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 64 * 1024, sBar = sB + 64 * 1024, sT = sBar + 16;
  volatile uint32_t* tslot = (volatile uint32_t*)(smem_raw + (sT - raw));
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 32 * 1024; i += 128) ((float*)(smem_raw + (base - raw)))[i] = 0.f;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sBar)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sT), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tslot, 0);
  if (warp == 1) {
    const uint32_t id1 = (1u << 4) | ((uint32_t)(NC >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t id2 = (1u << 4) | ((uint32_t)((2 * NC) >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t a0 = desc(sA, 64, 4), b64 = desc(sB, 64, 4), b32 = desc(sB, 32, 6);
    long long t0 = clock64();
    for (int it = 0; it < iters; it += 4) {
      if (elect_one()) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint64_t a = a0 + (uint64_t)(u * 4 * 3);   // shifted window rows
          if (pattern == 0) {
            mma_f16(tmem, a, b64, id1, 1u);
            mma_f16(tmem + 256, a, b64 + 2, id1, 1u);
            mma_f16(tmem + 256, a + 2, b64, id1, 1u);
          } else {
            mma_f16(tmem, a, b32, id2, 1u);
            mma_f16(tmem + NC, a + 2, b32, id1, 1u);
          }
        }
      }
      __syncwarp();
    }
    if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(sBar) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(sBar), "r"(0u) : "memory");
    long long t1 = clock64();
    if (threadIdx.x == 32) out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  long long* d_out; cudaMalloc(&d_out, 148 * 4 * sizeof(long long));
  float* sink; cudaMalloc(&sink, 4);
  long long h[148 * 4];
  const int iters = 4096;
  printf("tmem: mode warps  cycles/iter (48 cols x 32 lanes x 4 B = 6144 B per warp-iter)  B/clk/SM\n");
  const char* names[] = {"3x(ld16+wait)", "3xld16,wait", "ld32+ld16,wait", "3xst16,wait", "3x(ld16+wait)+fadd"};
  for (int mode = 0; mode < 5; ++mode)
    for (int nw : {1, 2, 4}) {
      tmem_bench<<<148, 128>>>(mode, iters, nw, d_out, sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost);
      double mx = 0; for (int i = 0; i < 148 * 4; ++i) if (i % 4 < nw) mx = h[i] > mx ? h[i] : mx;
      printf("%-20s %d  %8.1f  %8.1f\n", names[mode], nw, mx / iters, 6144.0 * nw / (mx / iters));
    }
  const int smem = 64 * 1024 * 2 + 2048;
  cudaFuncSetAttribute(mma_pair_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  printf("mma pair: NC pattern  cycles per (k16 x tap) step\n");
  for (int NC : {48, 64, 96, 128})
    for (int pat : {0, 1}) {
      mma_pair_bench<<<148, 128, smem>>>(NC, pat, iters, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d_out, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
      double mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("%4d %d  %8.1f\n", NC, pat, mx / iters);
    }
  return 0;
}
