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
// Two builds of the same layout idea (compile-time switch PE_FP16, see csrc/build.py):
//   PE_FP16=0  "tf32x3": chunk = [hi(16 x f32) | lo(16 x f32)] = 128 B; hi = value rounded to TF32, lo = value - hi (exact).
//                        8 bytes per element, no range limit; tcgen05 kind::tf32, K = 8 per MMA.
//   PE_FP16=1  "fp16x2": chunk = [h(16 x f16) | l(16 x f16)]   =  64 B; h = fp16(value), l = fp16((value - h) * 2^11): 22
//                        significant bits, 4 bytes per element; tcgen05 kind::f16, K = 16 per MMA: half the MMAs and half
//                        the bytes.  The 2^11 on the low half (PS_LO_SCALE) keeps it a NORMAL fp16 number whenever the value
//                        itself is one (|l| <= |value|): without it the low half of any |value| < 2^-3 is an fp16
//                        subnormal with an absolute step of 6e-8, i.e. a tensor living around 1e-3 would carry only ~15
//                        bits.  The cross terms hi*lo + lo*hi accumulate in their own TMEM accumulator, so the 2^-11 is
//                        applied once, in the epilogue (exact).  Full precision range: 6.1e-5 <= |value| <= 65504;
//                        smaller values degrade gracefully (absolute error <= 3e-11); larger ones saturate AND raise the
//                        model's range flag (pe_topdown then fails with PE_ERR_RANGE instead of returning clamped results).
#ifndef PE_FP16
#define PE_FP16 0
#endif
#include <cuda_fp16.h>

#if PE_FP16
#define PS_CHUNK_BYTES 64
#else
#define PS_CHUNK_BYTES 128
#endif
#define PS_CHUNK_FLOATS (PS_CHUNK_BYTES / 4)
#if PE_FP16
#define PS_LO_SCALE 2048.0f
#define PS_LO_INV (1.0f / 2048.0f)
#define PS_ABS_MAX 65504.0f
#else
#define PS_LO_SCALE 1.0f
#define PS_LO_INV 1.0f
#define PS_ABS_MAX 3.0e38f
#endif

// floats (4-byte units) per PS row of a C-channel tensor
__host__ __device__ __forceinline__ int ps_row_floats(int C) { return (C >> 4) * PS_CHUNK_FLOATS; }

__device__ __forceinline__ float tf32_round(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return __uint_as_float(u);
}

__device__ __forceinline__ void split4(const float4 v, float4& hi, float4& lo) {
  hi.x = tf32_round(v.x); hi.y = tf32_round(v.y); hi.z = tf32_round(v.z); hi.w = tf32_round(v.w);
  lo.x = v.x - hi.x; lo.y = v.y - hi.y; lo.z = v.z - hi.z; lo.w = v.w - hi.w;
}

// fp16 split of 4 values: h (4 halfs in a uint2) and l
__device__ __forceinline__ void split4_h(const float4 v, uint2& h, uint2& l) {
  const float lim = 65504.f;
  const float x = fminf(fmaxf(v.x, -lim), lim), y = fminf(fmaxf(v.y, -lim), lim);
  const float z = fminf(fmaxf(v.z, -lim), lim), w = fminf(fmaxf(v.w, -lim), lim);
  const __half2 h0 = __floats2half2_rn(x, y), h1 = __floats2half2_rn(z, w);
  const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
  const __half2 l0 = __floats2half2_rn((x - f0.x) * PS_LO_SCALE, (y - f0.y) * PS_LO_SCALE);     // exact scaling
  const __half2 l1 = __floats2half2_rn((z - f1.x) * PS_LO_SCALE, (w - f1.y) * PS_LO_SCALE);
  h.x = *reinterpret_cast<const uint32_t*>(&h0); h.y = *reinterpret_cast<const uint32_t*>(&h1);
  l.x = *reinterpret_cast<const uint32_t*>(&l0); l.y = *reinterpret_cast<const uint32_t*>(&l1);
}
__device__ __forceinline__ float4 join4_h(const uint2 h, const uint2 l) {
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&h.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
  const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&l.x)), d = __half22float2(*reinterpret_cast<const __half2*>(&l.y));
  return make_float4(fmaf(c.x, PS_LO_INV, a.x), fmaf(c.y, PS_LO_INV, a.y), fmaf(d.x, PS_LO_INV, b.x), fmaf(d.y, PS_LO_INV, b.y));
}

// range flag: set when a value about to be stored does not fit the activation format (fp16x2: |v| > 65504)
__device__ __forceinline__ void ps_range_check4(const float4 v, unsigned int* flag) {
#if PE_FP16
  if (flag && fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))) > PS_ABS_MAX) atomicOr(flag, 1u);
#else
  (void)v; (void)flag;
#endif
}

// byte offset of channel c (multiple of 4) inside a PS row (the hi / h part; the lo / l part is PS_CHUNK_BYTES/2 further)
__device__ __forceinline__ int ps_chan_byte(int c) {
#if PE_FP16
  return ((c >> 4) << 6) + ((c & 15) << 1);
#else
  return ((c >> 4) << 7) + ((c & 15) << 2);
#endif
}

// the fp32 values of channels c..c+3 of a row
__device__ __forceinline__ float4 ps_load4(const float* row, int c) {
  const char* p = reinterpret_cast<const char*>(row) + ps_chan_byte(c);
#if PE_FP16
  return join4_h(*reinterpret_cast<const uint2*>(p), *reinterpret_cast<const uint2*>(p + 32));
#else
  const float4 h = *reinterpret_cast<const float4*>(p);
  const float4 l = *reinterpret_cast<const float4*>(p + 64);
  return make_float4(h.x + l.x, h.y + l.y, h.z + l.z, h.w + l.w);
#endif
}

__device__ __forceinline__ void ps_store4(float* row, int c, float4 v) {
  char* p = reinterpret_cast<char*>(row) + ps_chan_byte(c);
#if PE_FP16
  uint2 h, l;
  split4_h(v, h, l);
  *reinterpret_cast<uint2*>(p) = h;
  *reinterpret_cast<uint2*>(p + 32) = l;
#else
  float4 hi, lo;
  split4(v, hi, lo);
  *reinterpret_cast<float4*>(p) = hi;
  *reinterpret_cast<float4*>(p + 64) = lo;
#endif
}

__device__ __forceinline__ void ps_zero4(float* row, int c) {
  char* p = reinterpret_cast<char*>(row) + ps_chan_byte(c);
#if PE_FP16
  *reinterpret_cast<uint2*>(p) = make_uint2(0u, 0u);
  *reinterpret_cast<uint2*>(p + 32) = make_uint2(0u, 0u);
#else
  *reinterpret_cast<float4*>(p) = make_float4(0.f, 0.f, 0.f, 0.f);
  *reinterpret_cast<float4*>(p + 64) = make_float4(0.f, 0.f, 0.f, 0.f);
#endif
}

// the fp32 values of one whole 16-channel chunk of a row (16-byte loads) / the split store of 16 values into a chunk
__device__ __forceinline__ void chunk_load16(const float* row, int chunk, float (&v)[16]) {
  const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(row) + (size_t)chunk * PS_CHUNK_BYTES);
#if PE_FP16
  const uint4 h0 = __ldg(p), h1 = __ldg(p + 1), l0 = __ldg(p + 2), l1 = __ldg(p + 3);
  const uint32_t hw[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
  const uint32_t lw[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw[i]));
    const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(&lw[i]));
    v[2 * i] = fmaf(lf.x, PS_LO_INV, hf.x); v[2 * i + 1] = fmaf(lf.y, PS_LO_INV, hf.y);
  }
#else
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint4 h = __ldg(p + i), l = __ldg(p + 4 + i);
    v[4 * i + 0] = __uint_as_float(h.x) + __uint_as_float(l.x); v[4 * i + 1] = __uint_as_float(h.y) + __uint_as_float(l.y);
    v[4 * i + 2] = __uint_as_float(h.z) + __uint_as_float(l.z); v[4 * i + 3] = __uint_as_float(h.w) + __uint_as_float(l.w);
  }
#endif
}

__device__ __forceinline__ void chunk_store16(float* row, int chunk, const float (&v)[16]) {
  uint4* p = reinterpret_cast<uint4*>(reinterpret_cast<char*>(row) + (size_t)chunk * PS_CHUNK_BYTES);
#if PE_FP16
  uint2 h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split4_h(make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]), h[i], l[i]);
  p[0] = make_uint4(h[0].x, h[0].y, h[1].x, h[1].y); p[1] = make_uint4(h[2].x, h[2].y, h[3].x, h[3].y);
  p[2] = make_uint4(l[0].x, l[0].y, l[1].x, l[1].y); p[3] = make_uint4(l[2].x, l[2].y, l[3].x, l[3].y);
#else
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float4 hi, lo;
    split4(make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]), hi, lo);
    p[i] = make_uint4(__float_as_uint(hi.x), __float_as_uint(hi.y), __float_as_uint(hi.z), __float_as_uint(hi.w));
    p[4 + i] = make_uint4(__float_as_uint(lo.x), __float_as_uint(lo.y), __float_as_uint(lo.z), __float_as_uint(lo.w));
  }
#endif
}

// raw copy of 4 channels (both halves) between rows: no arithmetic, used by the space-to-depth repack
__device__ __forceinline__ void ps_copy4(float* drow, int cd, const float* srow, int cs) {
  char* d = reinterpret_cast<char*>(drow) + ps_chan_byte(cd);
  const char* s = reinterpret_cast<const char*>(srow) + ps_chan_byte(cs);
#if PE_FP16
  *reinterpret_cast<uint2*>(d) = *reinterpret_cast<const uint2*>(s);
  *reinterpret_cast<uint2*>(d + 32) = *reinterpret_cast<const uint2*>(s + 32);
#else
  *reinterpret_cast<float4*>(d) = *reinterpret_cast<const float4*>(s);
  *reinterpret_cast<float4*>(d + 64) = *reinterpret_cast<const float4*>(s + 64);
#endif
}
