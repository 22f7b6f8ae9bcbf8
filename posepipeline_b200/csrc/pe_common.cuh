// Shared device/host helpers for libposeengine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// ---------------------------------------------------------------------------------------------
// Activation layout in HBM ("PS" = padded, split):
//   tensor (C,H,W) of image n is stored over a (H+2)x(W+2) grid whose 1-pixel border is zero, so a
//   3x3/pad-1 convolution is nine row-shifted GEMMs over one flat [rows][channels] matrix and no
//   kernel ever needs a bounds check.  Row p = (n*(H+2) + y+1)*(W+2) + x+1.
//   Each row holds C/16 chunks of 32 floats: [hi(16) | lo(16)], hi = value rounded to TF32 (10-bit
//   mantissa, round-to-nearest), lo = value - hi (exact in fp32).  The tensor-core kernels feed hi and
//   lo to tcgen05.mma kind::tf32 directly (3xTF32 split precision); SIMT kernels read hi+lo (= the fp32
//   value, exactly).  One 128-byte chunk = one TMA/UMMA SWIZZLE_128B row.
// ---------------------------------------------------------------------------------------------
struct ActView {
  float* base;       // device pointer to row 0
  int C, H, W;       // logical dims per image
  __host__ __device__ int Hp() const { return H + 2; }
  __host__ __device__ int Wp() const { return W + 2; }
  __host__ __device__ int rowFloats() const { return 2 * C; }
  __host__ __device__ long long rowsPerImage() const { return (long long)(H + 2) * (W + 2); }
};

__device__ __forceinline__ float tf32_round(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return __uint_as_float(u);
}

__device__ __forceinline__ void split4(const float4 v, float4& hi, float4& lo) {
  hi.x = tf32_round(v.x); hi.y = tf32_round(v.y); hi.z = tf32_round(v.z); hi.w = tf32_round(v.w);
  lo.x = v.x - hi.x; lo.y = v.y - hi.y; lo.z = v.z - hi.z; lo.w = v.w - hi.w;
}

// offset (in floats) of channel c (multiple of 4) inside a PS row
__device__ __forceinline__ int ps_chan_off(int c) { return ((c >> 4) << 5) + (c & 15); }

__device__ __forceinline__ float4 ps_load4(const float* row, int c) {
  const float* p = row + ps_chan_off(c);
  float4 h = *reinterpret_cast<const float4*>(p);
  float4 l = *reinterpret_cast<const float4*>(p + 16);
  return make_float4(h.x + l.x, h.y + l.y, h.z + l.z, h.w + l.w);
}

__device__ __forceinline__ void ps_store4(float* row, int c, float4 v) {
  float4 hi, lo;
  split4(v, hi, lo);
  float* p = row + ps_chan_off(c);
  *reinterpret_cast<float4*>(p) = hi;
  *reinterpret_cast<float4*>(p + 16) = lo;
}
