// Microbenchmark: TMA load throughput per SM for the box shapes conv_tc uses (L2-resident source).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_bench tools/tma_bench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// mode 0: 2D tensor loads (box = inner x rows); mode 1: 1D bulk copies of `bytes` contiguous bytes
__global__ void __launch_bounds__(32, 1) tma_bench(const __grid_constant__ CUtensorMap tm, const uint8_t* src, int mode, int box_bytes, int rows_per_box,
                                                    int nboxes_total, int depth, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw), base = (raw + 1023u) & ~1023u;
  const uint32_t sBar = base + 160 * 1024;
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sBar + 8 * i));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    long long t0 = clock64();
    int box = blockIdx.x * 7;
    for (int it = 0; it < iters + depth; ++it) {
      const int s = it % depth;
      const uint32_t bar = sBar + 8 * s, dst = base + (uint32_t)s * ((box_bytes + 1023) & ~1023);
      if (it >= depth) { const uint32_t par = ((it / depth) - 1) & 1; while (!mbar_try(bar, par)) {} }
      if (it < iters) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(box_bytes) : "memory");
        box = (box + 1) % nboxes_total;
        if (mode == 0)
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                       ::"r"(dst), "l"((uint64_t)&tm), "r"(0), "r"(box * rows_per_box), "r"(bar) : "memory");
        else
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(dst), "l"(src + (size_t)box * box_bytes), "r"(box_bytes), "r"(bar) : "memory");
      }
    }
    out[blockIdx.x] = clock64() - t0;
  }
}
// burst mode: issue D loads back to back onto one barrier, wait for all; nwarps issuing threads (one per warp), each its own barrier
__global__ void __launch_bounds__(128, 1) tma_burst(const __grid_constant__ CUtensorMap tm, const uint8_t* src, int mode, int box_bytes, int rows_per_box,
                                                     int nboxes_total, int D, int nwarps, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw), base = (raw + 1023u) & ~1023u;
  const uint32_t sBar = base + 192 * 1024;
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && warp < nwarps) {
    const uint32_t bar = sBar + 8 * warp;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t slot = (box_bytes + 1023) & ~1023;
    long long t0 = clock64();
    int box = (blockIdx.x * 4 + warp) * 11;
    for (int it = 0; it < iters; ++it) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(box_bytes * D) : "memory");
      for (int d = 0; d < D; ++d) {
        box = (box + 1) % nboxes_total;
        const uint32_t dst = base + (uint32_t)(warp * D + d) * slot;
        if (mode == 0)
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                       ::"r"(dst), "l"((uint64_t)&tm), "r"(0), "r"(box * rows_per_box), "r"(bar) : "memory");
        else
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(dst), "l"(src + (size_t)box * box_bytes), "r"(box_bytes), "r"(bar) : "memory");
      }
      while (!mbar_try(bar, it & 1)) {}
    }
    out[blockIdx.x * 4 + warp] = clock64() - t0;
  }
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  void* fp = nullptr; cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr);
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  const size_t bytes = 48ull << 20;     // 48 MB source: L2-resident after the first pass
  uint8_t* src; cudaMalloc(&src, bytes); cudaMemset(src, 1, bytes);
  long long* d_out; cudaMalloc(&d_out, 148 * 4 * 8); long long h[148 * 4];
  cudaFuncSetAttribute(tma_burst, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(tma_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct Cfg { const char* name; int inner_bytes, pitch_bytes, rows, mode; CUtensorMapSwizzle sw; };
  Cfg cfgs[] = {
    {"2D 64B x 256 rows, pitch 192 (A window, C=48)", 64, 192, 256, 0, CU_TENSOR_MAP_SWIZZLE_64B},
    {"2D 64B x 256 rows, pitch 384 (A window, C=96)", 64, 384, 256, 0, CU_TENSOR_MAP_SWIZZLE_64B},
    {"2D 64B x 144 rows, pitch 64 (W stage, contiguous)", 64, 64, 144, 0, CU_TENSOR_MAP_SWIZZLE_64B},
    {"2D 64B x 256 rows, pitch 64 (chunk-plane A)", 64, 64, 256, 0, CU_TENSOR_MAP_SWIZZLE_64B},
    {"2D 128B x 128 rows, pitch 128 (contiguous)", 128, 128, 128, 0, CU_TENSOR_MAP_SWIZZLE_128B},
    {"2D 128B x 128 rows, pitch 384", 128, 384, 128, 0, CU_TENSOR_MAP_SWIZZLE_128B},
    {"2D 128B x 256 rows, pitch 384", 128, 384, 256, 0, CU_TENSOR_MAP_SWIZZLE_128B},
    {"1D bulk 9216 B", 9216, 9216, 1, 1, CU_TENSOR_MAP_SWIZZLE_NONE},
    {"1D bulk 16384 B", 16384, 16384, 1, 1, CU_TENSOR_MAP_SWIZZLE_NONE},
    {"1D bulk 26112 B", 26112, 26112, 1, 1, CU_TENSOR_MAP_SWIZZLE_NONE},
  };
  printf("%-52s depth  B/clk/SM  (GB/s chip @1.9GHz)\n", "config");
  for (auto& c : cfgs) {
    CUtensorMap tm{};
    const int box_bytes = c.inner_bytes * c.rows;
    const uint64_t nrows = bytes / c.pitch_bytes;
    int nboxes = c.mode == 0 ? (int)(nrows / c.rows) : (int)(bytes / box_bytes);
    if (c.mode == 0) {
      cuuint64_t gdim[2] = {(cuuint64_t)c.inner_bytes / 4, nrows}; cuuint64_t gstr[1] = {(cuuint64_t)c.pitch_bytes};
      cuuint32_t box[2] = {(cuuint32_t)c.inner_bytes / 4, (cuuint32_t)c.rows}; cuuint32_t es[2] = {1, 1};
      CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, src, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, c.sw,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d for %s\n", (int)r, c.name); continue; }
    }
    for (int nw : {1, 2, 4})
      for (int D : {1, 2, 4, 8}) {
        if ((size_t)nw * D * ((box_bytes + 1023) & ~1023) > 192 * 1024) continue;
        for (int rep = 0; rep < 2; ++rep) {
          tma_burst<<<148, 128, 200 * 1024>>>(tm, src, c.mode, box_bytes, c.rows, nboxes, D, nw, 500, d_out);
          if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        }
        cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost);
        double mx = 0; for (int i = 0; i < 148 * 4; ++i) if (i % 4 < nw) mx = h[i] > mx ? h[i] : mx;
        printf("  burst: warps %d x D %d : %7.0f clk per round, %6.1f B/clk/SM\n", nw, D, mx / 500, 500.0 * nw * D * box_bytes / mx);
      }
    // restrict the working set to ~40 MB so it stays in L2
    for (int depth : {2, 4, 6}) {
      if ((size_t)depth * ((box_bytes + 1023) & ~1023) > 160 * 1024) continue;
      for (int rep = 0; rep < 2; ++rep) {
        tma_bench<<<148, 32, 200 * 1024>>>(tm, src, c.mode, box_bytes, c.rows, nboxes, depth, 2000, d_out);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
      }
      cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost);
      double mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
      const double bpc = 2000.0 * box_bytes / mx;
      printf("%-52s %d  %8.1f  %8.0f\n", c.name, depth, bpc, bpc * 148 * 1.9);
    }
  }
  return 0;
}
