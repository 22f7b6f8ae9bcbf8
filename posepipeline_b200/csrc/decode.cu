// K6: flip-merge + keypoints_from_heatmaps, one CTA per (crop, joint).
// Replaces (reference call site pose_pipeline/wrappers/mmpose.py:75):
//   flip_back + 1-px shift + average          (TopDown.forward_test, cfg :82,:84)
//   _get_max_preds                            (argmax, score = max of the un-blurred map)
//   _gaussian_blur(kernel) -> log(max(.,1e-10)) -> _taylor     (post_process='unbiased', cfg :83,:85)
//   or the +-0.25 px sign shift               (post_process='default')
//   transform_preds                           (back-projection to image pixels)
// The maths is also stated by the reference's in-tree DarkPose copy pose_pipeline/utils/inference.py:27-92.
// HBM-bound: reads 2*H*W*4 bytes per (crop, joint), writes 12 bytes.  The whole map lives in shared memory.
#include "pe_common.cuh"
#include "kernels.h"

// 1-D Gaussian taps (cv2.getGaussianKernel(k, 0.3*((k-1)*0.5-1)+0.8), float32) come per model through DecodeArgs::gauss
// (device buffer of 64 floats) and are copied to shared memory: two models with different modulate_kernel values coexist.

struct DecodeArgs {
  const float* hm;       // [n][K][H][W]
  const float* hm_flip;  // raw flipped-pass output or nullptr
  const int* flip_perm;  // [K]
  const float* center;   // [n][2]
  const float* scale;    // [n][2]
  float* out;            // [n][K][3]
  const float* gauss;    // [64] taps (post == 2)
  int K, H, W, shift, post, ksize;
};

__device__ __forceinline__ void block_argmax(float& v, int& idx, float* s_v, int* s_i) {
  // max value, smallest index on ties (np.argmax returns the first maximum)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_v[warp] = v; s_i[warp] = idx; }
  __syncthreads();
  if (warp == 0) {
    v = lane < (blockDim.x >> 5) ? s_v[lane] : -INFINITY;
    idx = lane < (blockDim.x >> 5) ? s_i[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    if (lane == 0) { s_v[0] = v; s_i[0] = idx; }
  }
  __syncthreads();
  v = s_v[0];
  idx = s_i[0];
  __syncthreads();
}

__global__ void __launch_bounds__(256) decode_kernel(DecodeArgs a) {
  extern __shared__ float smem[];
  const int HW = a.H * a.W;
  float* s_m = smem;        // merged heatmap
  float* s_t = smem + HW;   // row-filtered
  float* s_b = smem + 2 * HW;  // blurred
  __shared__ float s_v[8];
  __shared__ int s_i[8];
  __shared__ float c_gauss[64];
  if (threadIdx.x < 64) c_gauss[threadIdx.x] = a.gauss ? a.gauss[threadIdx.x] : 0.f;   // visible after block_argmax's barrier

  const int k = blockIdx.x, n = blockIdx.y;
  const float* src = a.hm + ((size_t)n * a.K + k) * HW;
  const float* srcf = a.hm_flip ? a.hm_flip + ((size_t)n * a.K + a.flip_perm[k]) * HW : nullptr;

  float best = -INFINITY;
  int besti = 0x7fffffff;
  for (int i = threadIdx.x; i < HW; i += 256) {
    float v = src[i];
    if (srcf) {
      const int y = i / a.W, x = i % a.W;
      const int xs = (a.shift && x > 0) ? x - 1 : x;
      v = __fmul_rn(__fadd_rn(v, srcf[y * a.W + (a.W - 1 - xs)]), 0.5f);
    }
    s_m[i] = v;
    if (v > best) { best = v; besti = i; }
  }
  block_argmax(best, besti, s_v, s_i);  // includes __syncthreads: s_m complete
  const float maxval = best;
  float cx = (float)(besti % a.W), cy = (float)(besti / a.W);
  if (!(maxval > 0.0f)) { cx = -1.f; cy = -1.f; }
  const int px = (int)cx, py = (int)cy;

  double offx = 0.0, offy = 0.0;
  float shx = 0.f, shy = 0.f;
  if (a.post == 2) {
    const int R = (a.ksize - 1) / 2;
    // row filter (zero padded), symmetric-pair form
    for (int i = threadIdx.x; i < HW; i += 256) {
      const int y = i / a.W, x = i % a.W;
      const float* row = s_m + y * a.W;
      float s = c_gauss[R] * row[x];
      for (int d = 1; d <= R; ++d) {
        const float l = (x - d >= 0) ? row[x - d] : 0.f;
        const float r = (x + d < a.W) ? row[x + d] : 0.f;
        s = fmaf(c_gauss[R + d], l + r, s);
      }
      s_t[i] = s;
    }
    __syncthreads();
    // column filter -> s_b
    float bmax = -INFINITY;
    int dummy = 0;
    for (int i = threadIdx.x; i < HW; i += 256) {
      const int y = i / a.W, x = i % a.W;
      float s = c_gauss[R] * s_t[i];
      for (int d = 1; d <= R; ++d) {
        const float u = (y - d >= 0) ? s_t[(y - d) * a.W + x] : 0.f;
        const float w = (y + d < a.H) ? s_t[(y + d) * a.W + x] : 0.f;
        s = fmaf(c_gauss[R + d], u + w, s);
      }
      s_b[i] = s;
      bmax = fmaxf(bmax, s);
    }
    block_argmax(bmax, dummy, s_v, s_i);
    if (threadIdx.x == 0 && 1 < px && px < a.W - 2 && 1 < py && py < a.H - 2) {
      const float sc = __fdiv_rn(maxval, bmax);
      auto L = [&](int yy, int xx) -> float { return logf(fmaxf(__fmul_rn(s_b[yy * a.W + xx], sc), 1e-10f)); };
      const float c00 = L(py, px);
      const double dx = 0.5 * (double)__fsub_rn(L(py, px + 1), L(py, px - 1));
      const double dy = 0.5 * (double)__fsub_rn(L(py + 1, px), L(py - 1, px));
      const double dxx = 0.25 * ((double)L(py, px + 2) - 2.0 * (double)c00 + (double)L(py, px - 2));
      const float m4 = __fadd_rn(__fsub_rn(__fsub_rn(L(py + 1, px + 1), L(py - 1, px + 1)), L(py + 1, px - 1)), L(py - 1, px - 1));
      const double dxy = 0.25 * (double)m4;
      const double dyy = 0.25 * ((double)L(py + 2, px) - 2.0 * (double)c00 + (double)L(py - 2, px));
      const double det = dxx * dyy - dxy * dxy;
      if (det != 0.0) {
        // offset = -H^-1 g
        offx = -(dyy * dx - dxy * dy) / det;
        offy = -(-dxy * dx + dxx * dy) / det;
      }
    }
  } else if (a.post == 3) {
    // mmpose post_dark_udp (use_udp + GaussianHeatmap, ViTPose configs): cv2.GaussianBlur in place with the default
    // BORDER_REFLECT_101, clip to [0.001, 50], log, second-order step on edge-replicated neighbours; the 2x2 system is solved
    // in float64 (hessian + float64 eps * I), the rest is float32 like the numpy arrays
    const int R = (a.ksize - 1) / 2;
    auto refl = [](int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); };
    for (int i = threadIdx.x; i < HW; i += 256) {
      const int y = i / a.W, x = i % a.W;
      const float* row = s_m + y * a.W;
      float s = c_gauss[R] * row[x];
      for (int d = 1; d <= R; ++d) s = fmaf(c_gauss[R + d], row[refl(x - d, a.W)] + row[refl(x + d, a.W)], s);
      s_t[i] = s;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < HW; i += 256) {
      const int y = i / a.W, x = i % a.W;
      float s = c_gauss[R] * s_t[i];
      for (int d = 1; d <= R; ++d) s = fmaf(c_gauss[R + d], s_t[refl(y - d, a.H) * a.W + x] + s_t[refl(y + d, a.H) * a.W + x], s);
      s_b[i] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0 && px >= 0 && py >= 0) {
      auto L = [&](int yy, int xx) -> float {
        yy = min(max(yy, 0), a.H - 1); xx = min(max(xx, 0), a.W - 1);           // np.pad(mode='edge')
        return logf(fminf(fmaxf(s_b[yy * a.W + xx], 0.001f), 50.f));
      };
      const float i_ = L(py, px), ix1 = L(py, px + 1), iy1 = L(py + 1, px), ix1y1 = L(py + 1, px + 1);
      const float ix1_y1_ = L(py - 1, px - 1), ix1_ = L(py, px - 1), iy1_ = L(py - 1, px);
      const float dx = __fmul_rn(0.5f, __fsub_rn(ix1, ix1_)), dy = __fmul_rn(0.5f, __fsub_rn(iy1, iy1_));
      const float dxx = __fadd_rn(__fsub_rn(ix1, __fmul_rn(2.f, i_)), ix1_), dyy = __fadd_rn(__fsub_rn(iy1, __fmul_rn(2.f, i_)), iy1_);
      float t = __fsub_rn(ix1y1, ix1);
      t = __fsub_rn(t, iy1); t = __fadd_rn(t, i_); t = __fadd_rn(t, i_); t = __fsub_rn(t, ix1_); t = __fsub_rn(t, iy1_); t = __fadd_rn(t, ix1_y1_);
      const float dxy = __fmul_rn(0.5f, t);
      const double eps = 1.1920928955078125e-07;
      const double h00 = (double)dxx + eps, h01 = (double)dxy, h11 = (double)dyy + eps;
      const double det = h00 * h11 - h01 * h01;
      if (det != 0.0) {
        offx = -((h11 * (double)dx - h01 * (double)dy) / det);      // coords -= H^-1 g
        offy = -((-h01 * (double)dx + h00 * (double)dy) / det);
      }
    }
  } else if (a.post == 1) {
    if (threadIdx.x == 0 && 1 < px && px < a.W - 1 && 1 < py && py < a.H - 1) {
      const float ddx = __fsub_rn(s_m[py * a.W + px + 1], s_m[py * a.W + px - 1]);
      const float ddy = __fsub_rn(s_m[(py + 1) * a.W + px], s_m[(py - 1) * a.W + px]);
      shx = (ddx > 0.f) ? 0.25f : (ddx < 0.f ? -0.25f : 0.f);
      shy = (ddy > 0.f) ? 0.25f : (ddy < 0.f ? -0.25f : 0.f);
    }
  }
  if (threadIdx.x == 0) {
    float fx = cx, fy = cy;
    if (a.post == 2 || a.post == 3) {
      fx = (float)((double)cx + offx);
      fy = (float)((double)cy + offy);
    } else if (a.post == 1) {
      fx = __fadd_rn(cx, shx);
      fy = __fadd_rn(cy, shy);
    }
    // transform_preds: float32 arithmetic, one rounding per numpy operation
    const float sx = __fmul_rn(a.scale[n * 2 + 0], 200.0f), sy = __fmul_rn(a.scale[n * 2 + 1], 200.0f);
    // transform_preds: scale / output_size, or scale / (output_size - 1) with use_udp
    const float scx = __fdiv_rn(sx, (float)(a.post == 3 ? a.W - 1 : a.W)), scy = __fdiv_rn(sy, (float)(a.post == 3 ? a.H - 1 : a.H));
    const float ox = __fsub_rn(__fadd_rn(__fmul_rn(fx, scx), a.center[n * 2 + 0]), __fmul_rn(sx, 0.5f));
    const float oy = __fsub_rn(__fadd_rn(__fmul_rn(fy, scy), a.center[n * 2 + 1]), __fmul_rn(sy, 0.5f));
    float* o = a.out + ((size_t)n * a.K + k) * 3;
    o[0] = ox; o[1] = oy; o[2] = maxval;
  }
}

int decode_smem_bytes(int H, int W) { return 3 * H * W * (int)sizeof(float); }

cudaError_t launch_decode(const float* hm, const float* hm_flip, const int* flip_perm, const float* center,
                          const float* scale, float* out, int n, int K, int H, int W, int shift, int post, int ksize,
                          const float* gauss, cudaStream_t st) {
  const int smem = decode_smem_bytes(H, W);
  if (smem > 220 * 1024) return cudaErrorInvalidValue;
  if ((post == 2 || post == 3) && !gauss) return cudaErrorInvalidValue;
  cudaError_t e = pe_smem_optin((const void*)decode_kernel, smem);
  if (e != cudaSuccess) return e;
  DecodeArgs a{hm, hm_flip, flip_perm, center, scale, out, gauss, K, H, W, shift, post, ksize};
  decode_kernel<<<dim3(K, n), 256, smem, st>>>(a);
  return cudaGetLastError();
}
