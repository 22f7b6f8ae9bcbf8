// fp32 SIMT kernels of the top-down path: affine crop, stem, generic conv (bring-up / odd shapes),
// HRModule fuse, head.  The dense 3x3 / 1x1 stride-1 convolutions run on tcgen05 (conv_tc.cu) when
// the model is created with use_tensor_cores=1; this file is the fp32 path they are checked against
// and the path for stride-2 / tiny-channel layers.
#include "pe_common.cuh"
#include "kernels.h"

// =============================================================================================
// K1: cv2.warpAffine(INTER_LINEAR, BORDER_CONSTANT=0) -- bit-exact fixed-point restatement.
// Replaces mmpose TopDownAffine (reference call site pose_pipeline/wrappers/mmpose.py:75; SURVEY A.1
// step 4, App. B.1).  OpenCV's algorithm: inverse matrix in double; per column
// adelta/bdelta = round(M*x*1024); per row X0/Y0 = round((M*y+b)*1024) + 16; coordinates in 1/32 px;
// bilinear weights are the exact integers (32-ax)(32-ay)...; result = (sum + 512) >> 10.
// =============================================================================================
__device__ __forceinline__ int sat_short(int v) { return max(-32768, min(32767, v)); }

__global__ void __launch_bounds__(256) warp_crop_kernel(const uint8_t* __restrict__ frames, int fh, int fw,
                                                        const int32_t* __restrict__ frame_idx,
                                                        const double* __restrict__ minv,  // n*6 inverse maps
                                                        uint8_t* __restrict__ crops, int ch, int cw, int swap_rb) {
  const int crop = blockIdx.z;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= cw || y >= ch) return;
  const double* M = minv + crop * 6;
  const int adelta = __double2int_rn(__dmul_rn(__dmul_rn(M[0], (double)x), 1024.0));
  const int bdelta = __double2int_rn(__dmul_rn(__dmul_rn(M[3], (double)x), 1024.0));
  const int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(M[1], (double)y), M[2]), 1024.0)) + 16;
  const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(M[4], (double)y), M[5]), 1024.0)) + 16;
  const int X = (X0 + adelta) >> 5, Y = (Y0 + bdelta) >> 5;
  const int sx = sat_short(X >> 5), sy = sat_short(Y >> 5);
  const int ax = X & 31, ay = Y & 31;
  const uint8_t* f = frames + (size_t)frame_idx[crop] * fh * fw * 3;
  const bool x0 = (unsigned)sx < (unsigned)fw, x1 = (unsigned)(sx + 1) < (unsigned)fw;
  const bool y0 = (unsigned)sy < (unsigned)fh, y1 = (unsigned)(sy + 1) < (unsigned)fh;
  const int w00 = (32 - ax) * (32 - ay), w01 = ax * (32 - ay), w10 = (32 - ax) * ay, w11 = ax * ay;
  uint8_t* o = crops + (((size_t)crop * ch + y) * cw + x) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    int p00 = (x0 && y0) ? f[((size_t)sy * fw + sx) * 3 + c] : 0;
    int p01 = (x1 && y0) ? f[((size_t)sy * fw + sx + 1) * 3 + c] : 0;
    int p10 = (x0 && y1) ? f[((size_t)(sy + 1) * fw + sx) * 3 + c] : 0;
    int p11 = (x1 && y1) ? f[((size_t)(sy + 1) * fw + sx + 1) * 3 + c] : 0;
    int v = (w00 * p00 + w01 * p01 + w10 * p10 + w11 * p11 + 512) >> 10;
    o[swap_rb ? 2 - c : c] = (uint8_t)v;
  }
}

void launch_warp_crop(const uint8_t* frames, int fh, int fw, const int32_t* frame_idx, const double* minv,
                      uint8_t* crops, int n, int ch, int cw, int swap_rb, cudaStream_t st) {
  dim3 block(32, 8), grid((cw + 31) / 32, (ch + 7) / 8, n);
  warp_crop_kernel<<<grid, block, 0, st>>>(frames, fh, fw, frame_idx, minv, crops, ch, cw, swap_rb);
}

// =============================================================================================
// Stem: conv1 3->64 3x3 s2 p1 + folded BN + ReLU, reading the uint8 crop through the normalisation
// LUT (ToTensor + NormalizeTensor, cfg :132-136, fused).  Images [ncrop, 2*ncrop) are the flip-test
// pass: they read the crop mirrored in x (== img.flip(3)).
// =============================================================================================
__global__ void __launch_bounds__(256) stem_kernel(const uint8_t* __restrict__ crops, int ncrop, int nimg, int ih, int iw,
                                                   const float* __restrict__ lut,   // [3][256]
                                                   const float* __restrict__ w,     // [27][64]  (tap-major: (ky*3+kx)*3+c)
                                                   const float* __restrict__ bias,  // [64]
                                                   float* __restrict__ out, int oh, int ow) {
  __shared__ float s_w[27 * 64];
  __shared__ float s_lut[768];
  __shared__ float s_b[64];
  for (int i = threadIdx.x; i < 27 * 64; i += 256) s_w[i] = w[i];
  for (int i = threadIdx.x; i < 768; i += 256) s_lut[i] = lut[i];
  if (threadIdx.x < 64) s_b[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const int Hp = oh + 2, Wp = ow + 2;
  const long long M = (long long)nimg * Hp * Wp;
  const long long m = (long long)blockIdx.x * 64 + (threadIdx.x >> 2);
  if (m >= M) return;
  const int chunk = threadIdx.x & 3;  // 16 output channels
  const int img = (int)(m / (Hp * Wp));
  const int r = (int)(m % (Hp * Wp));
  const int py = r / Wp, px = r % Wp;
  float* orow = out + m * ps_row_floats(64);
  if (py < 1 || py > oh || px < 1 || px > ow) {
#pragma unroll
    for (int i = 0; i < 16; i += 4) ps_zero4(orow, chunk * 16 + i);
    return;
  }
  const bool flip = img >= ncrop;
  const uint8_t* c0 = crops + (size_t)(flip ? img - ncrop : img) * ih * iw * 3;
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = s_b[chunk * 16 + i];
  const int oy = py - 1, ox = px - 1;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = oy * 2 - 1 + ky;
    if (iy < 0 || iy >= ih) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = ox * 2 - 1 + kx;
      if (ix < 0 || ix >= iw) continue;
      const int sxx = flip ? iw - 1 - ix : ix;
      const uint8_t* pix = c0 + ((size_t)iy * iw + sxx) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float v = s_lut[c * 256 + pix[c]];
        const float* wr = s_w + ((ky * 3 + kx) * 3 + c) * 64 + chunk * 16;
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fmaf(v, wr[i], acc[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 16; i += 4) {
    ps_store4(orow, chunk * 16 + i, make_float4(fmaxf(acc[i], 0.f), fmaxf(acc[i + 1], 0.f), fmaxf(acc[i + 2], 0.f), fmaxf(acc[i + 3], 0.f)));
  }
}

void launch_stem(const uint8_t* crops, int ncrop, int nimg, int ih, int iw, const float* lut, const float* w,
                 const float* bias, float* out, int oh, int ow, cudaStream_t st) {
  long long M = (long long)nimg * (oh + 2) * (ow + 2);
  stem_kernel<<<(unsigned)((M + 63) / 64), 256, 0, st>>>(crops, ncrop, nimg, ih, iw, lut, w, bias, out, oh, ow);
}

// =============================================================================================
// Generic fp32 implicit-GEMM convolution over PS tensors (k in {1,3}, stride in {1,2}, pad k/2):
//   out[m][n] = act( sum_{tap,ci} in[row(m)+shift(tap)][ci] * w[tap][ci][n] + bias[n] (+ res[m][n]) )
// Block tile 64 output positions x BN channels, K step = 16 input channels of one tap, 4x4 register
// tile per thread, register prefetch of the next K step.  Output positions run over the PADDED grid;
// border positions are written as zeros so every produced tensor keeps the zero-halo invariant.
// =============================================================================================
struct ConvArgs {
  const float* in;
  float* out;
  const float* res;
  const float* w;     // [ks*ks][Cin][Cout]
  const float* bias;  // [Cout]
  int Cin, Cout, ks, stride, relu;
  int Hin, Win, Hout, Wout, nimg;
  // linear (1-D, dilated) mode used by the VideoPose3D lifter: rows are time steps, tap t reads row m + t*dil
  int linear, ntaps, dil, res_off, plain_out, cout_real;
  long long M_lin;
};

template <int BN>
__global__ void __launch_bounds__(BN * 4) conv_simt_kernel(ConvArgs a) {
  constexpr int NT = BN * 4;
  constexpr int TN = BN / 4;  // threads along N
  constexpr int AU = (256 + NT - 1) / NT;
  __shared__ __align__(16) float As[16][64 + 4];
  __shared__ __align__(16) float Bs[16][BN];
  __shared__ long long s_inrow[64];  // input row of tap (0,0) per tile position, or -1

  const int tid = threadIdx.x;
  const int tn = tid % TN, tm = tid / TN;
  const int HpO = a.Hout + 2, WpO = a.Wout + 2, WpI = a.Win + 2, HpI = a.Hin + 2;
  const long long M = a.linear ? a.M_lin : (long long)a.nimg * HpO * WpO;
  const long long m0 = (long long)blockIdx.x * 64;
  const int n0 = blockIdx.y * BN;
  const int off = (a.ks == 3) ? 0 : 1;

  if (tid < 64) {
    long long m = m0 + tid;
    long long row = -1;
    if (m < M && a.linear) {
      row = m;
    } else if (m < M) {
      int img = (int)(m / (HpO * WpO));
      int r = (int)(m % (HpO * WpO));
      int py = r / WpO, px = r % WpO;
      if (py >= 1 && py <= a.Hout && px >= 1 && px <= a.Wout)
        row = ((long long)img * HpI + (py - 1) * a.stride + off) * WpI + (px - 1) * a.stride + off;
    }
    s_inrow[tid] = row;
  }
  __syncthreads();

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nchunk = a.Cin >> 4;
  const int ksteps = a.ntaps * nchunk;
  const int inRowF = ps_row_floats(a.Cin), outRowF = ps_row_floats(a.Cout);

  float4 pa[AU];
  float4 pb;
  auto fetch = [&](int step) {
    const int tap = step / nchunk, ch = step % nchunk;
    const int shift = a.linear ? tap * a.dil : (tap / a.ks) * WpI + (tap % a.ks);
#pragma unroll
    for (int u = 0; u < AU; ++u) {
      const int idx = tid + u * NT;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < 256) {
        const long long row = s_inrow[idx >> 2];
        if (row >= 0) v = ps_load4(a.in + (row + shift) * inRowF, ch * 16 + (idx & 3) * 4);
      }
      pa[u] = v;
    }
    {
      const int kk = tid / TN, nn = (tid % TN) * 4;
      pb = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n0 + nn < a.Cout)
        pb = *reinterpret_cast<const float4*>(a.w + ((size_t)tap * a.Cin + ch * 16 + kk) * a.Cout + n0 + nn);
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int u = 0; u < AU; ++u) {
      const int idx = tid + u * NT;
      if (idx < 256) {
        const int px = idx >> 2, q = (idx & 3) * 4;
        As[q + 0][px] = pa[u].x; As[q + 1][px] = pa[u].y; As[q + 2][px] = pa[u].z; As[q + 3][px] = pa[u].w;
      }
    }
    *reinterpret_cast<float4*>(&Bs[tid / TN][(tid % TN) * 4]) = pb;
  };

  fetch(0);
  for (int step = 0; step < ksteps; ++step) {
    stash();
    __syncthreads();
    if (step + 1 < ksteps) fetch(step + 1);
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][tm * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tn * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w};
      const float br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    __syncthreads();
  }

  const int c = n0 + tn * 4;
  if (c >= a.Cout) return;
  const float4 bz = *reinterpret_cast<const float4*>(a.bias + c);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + tm * 4 + i;
    if (m >= M) break;
    if (a.plain_out) {
      const float vv[4] = {acc[i][0] + bz.x, acc[i][1] + bz.y, acc[i][2] + bz.z, acc[i][3] + bz.w};
      for (int j = 0; j < 4; ++j)
        if (c + j < a.cout_real) a.out[m * a.cout_real + c + j] = a.relu ? fmaxf(vv[j], 0.f) : vv[j];
      continue;
    }
    float* orow = a.out + m * outRowF;
    if (s_inrow[tm * 4 + i] < 0) {
      ps_zero4(orow, c);
      continue;
    }
    float4 v = make_float4(acc[i][0] + bz.x, acc[i][1] + bz.y, acc[i][2] + bz.z, acc[i][3] + bz.w);
    // res_off >= 0: residual before the ReLU (HRNet blocks); res_off < 0: after it, at row m - res_off (lifter)
    if (a.res && a.res_off >= 0) {
      const float4 rr = ps_load4(a.res + (m + a.res_off) * outRowF, c);
      v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w;
    }
    if (a.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    if (a.res && a.res_off < 0) {
      const float4 rr = ps_load4(a.res + (m - a.res_off) * outRowF, c);
      v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w;
    }
    ps_store4(orow, c, v);
  }
}

void launch_conv_simt(const float* in, float* out, const float* res, const float* w, const float* bias, int Cin,
                      int Cout, int ks, int stride, int relu, int Hin, int Win, int Hout, int Wout, int nimg,
                      cudaStream_t st) {
  ConvArgs a{in, out, res, w, bias, Cin, Cout, ks, stride, relu, Hin, Win, Hout, Wout, nimg, 0, ks * ks, 0, 0, 0, Cout, 0};
  long long M = (long long)nimg * (Hout + 2) * (Wout + 2);
  unsigned gx = (unsigned)((M + 63) / 64);
  if (Cout % 48 == 0) {
    conv_simt_kernel<48><<<dim3(gx, Cout / 48), 192, 0, st>>>(a);
  } else {
    conv_simt_kernel<64><<<dim3(gx, (Cout + 63) / 64), 256, 0, st>>>(a);
  }
}

// 1-D dilated convolution over PS rows (VideoPose3D temporal model): out[m] = sum_t in[m + t*dil] * w[t]
void launch_conv_linear(const float* in, float* out, const float* res, const float* w, const float* bias, int Cin,
                        int Cout, int ntaps, int dil, int relu, long long M, int res_off, int plain_out, int cout_real,
                        cudaStream_t st) {
  ConvArgs a{in, out, res, w, bias, Cin, Cout, 1, 1, relu, 0, 0, 0, 0, 1, 1, ntaps, dil, res_off, plain_out, cout_real, M};
  unsigned gx = (unsigned)((M + 63) / 64);
  if (Cout % 48 == 0) conv_simt_kernel<48><<<dim3(gx, Cout / 48), 192, 0, st>>>(a);
  else conv_simt_kernel<64><<<dim3(gx, (Cout + 63) / 64), 256, 0, st>>>(a);
}

// =============================================================================================
// HRModule fuse: out = ReLU(sum_j nearest_upsample_{up_j}(in_j)), summed in branch order.
// =============================================================================================
struct FuseArgs {
  const float* in[4];
  int up[4];
  int n_in;
  float* out;
  int C, H, W, nimg, relu;
};

__global__ void __launch_bounds__(256) fuse_kernel(FuseArgs a) {
  const int c4 = a.C >> 2;
  const long long total = (long long)a.nimg * (a.H + 2) * (a.W + 2) * c4;
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t >= total) return;
  const int c = (int)(t % c4) * 4;
  const long long m = t / c4;
  const int Hp = a.H + 2, Wp = a.W + 2;
  const int img = (int)(m / (Hp * Wp));
  const int r = (int)(m % (Hp * Wp));
  const int py = r / Wp, px = r % Wp;
  const int rowF = ps_row_floats(a.C);
  float* orow = a.out + m * rowF;
  if (py < 1 || py > a.H || px < 1 || px > a.W) {
    ps_zero4(orow, c);
    return;
  }
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int j = 0; j < a.n_in; ++j) {
    const int u = a.up[j];
    const int h = a.H / u, w = a.W / u;
    const long long row = ((long long)img * (h + 2) + (py - 1) / u + 1) * (w + 2) + (px - 1) / u + 1;
    const float4 v = ps_load4(a.in[j] + row * rowF, c);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  if (a.relu) { s.x = fmaxf(s.x, 0.f); s.y = fmaxf(s.y, 0.f); s.z = fmaxf(s.z, 0.f); s.w = fmaxf(s.w, 0.f); }
  ps_store4(orow, c, s);
}

void launch_fuse(const float* const* in, const int* up, int n_in, float* out, int C, int H, int W, int nimg, int relu,
                 cudaStream_t st) {
  FuseArgs a;
  for (int j = 0; j < 4; ++j) { a.in[j] = j < n_in ? in[j] : nullptr; a.up[j] = j < n_in ? up[j] : 1; }
  a.n_in = n_in; a.out = out; a.C = C; a.H = H; a.W = W; a.nimg = nimg; a.relu = relu;
  long long total = (long long)nimg * (H + 2) * (W + 2) * (C / 4);
  fuse_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a);
}

// =============================================================================================
// Head: final_layer 1x1 conv Cin->K with bias (cfg :73-79) -> planar fp32 heatmaps [img][K][H][W].
// =============================================================================================
__global__ void __launch_bounds__(128) head_kernel(const float* __restrict__ in, int Cin, int H, int W, int nimg,
                                                   const float* __restrict__ w,  // [Cin][K]
                                                   const float* __restrict__ bias, int K, float* __restrict__ out) {
  extern __shared__ float s_w[];  // Cin*K + K
  for (int i = threadIdx.x; i < Cin * K; i += 128) s_w[i] = w[i];
  for (int i = threadIdx.x; i < K; i += 128) s_w[Cin * K + i] = bias[i];
  __syncthreads();
  const long long t = (long long)blockIdx.x * 128 + threadIdx.x;
  if (t >= (long long)nimg * H * W) return;
  const int img = (int)(t / (H * W));
  const int r = (int)(t % (H * W));
  const int y = r / W, x = r % W;
  const float* row = in + (((long long)img * (H + 2) + y + 1) * (W + 2) + x + 1) * ps_row_floats(Cin);
  for (int k0 = 0; k0 < K; k0 += 32) {
    float acc[32];
    const int kn = min(32, K - k0);
#pragma unroll
    for (int k = 0; k < 32; ++k) acc[k] = 0.f;
    for (int c = 0; c < Cin; c += 4) {
      const float4 v = ps_load4(row, c);
      const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float* wr = s_w + (c + i) * K + k0;
#pragma unroll
        for (int k = 0; k < 32; ++k)
          if (k < kn) acc[k] = fmaf(vv[i], wr[k], acc[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < 32; ++k)
      if (k < kn) out[(((long long)img * K + k0 + k) * H + y) * W + x] = acc[k] + s_w[Cin * K + k0 + k];
  }
}

void launch_head(const float* in, int Cin, int H, int W, int nimg, const float* w, const float* bias, int K, float* out,
                 cudaStream_t st) {
  long long total = (long long)nimg * H * W;
  size_t smem = (size_t)(Cin * K + K) * sizeof(float);
  head_kernel<<<(unsigned)((total + 127) / 128), 128, smem, st>>>(in, Cin, H, W, nimg, w, bias, K, out);
}

// debug: PS tensor image -> dense CHW fp32
__global__ void ps_to_chw_kernel(const float* __restrict__ in, int C, int H, int W, int img, float* __restrict__ out) {
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t >= (long long)C * H * W) return;
  const int c = (int)(t / (H * W));
  const int r = (int)(t % (H * W));
  const int y = r / W, x = r % W;
  const float* row = in + (((long long)img * (H + 2) + y + 1) * (W + 2) + x + 1) * ps_row_floats(C);
  const float4 v = ps_load4(row, c & ~3);
  const float vv[4] = {v.x, v.y, v.z, v.w};
  out[t] = vv[c & 3];
}

void launch_ps_to_chw(const float* in, int C, int H, int W, int img, float* out, cudaStream_t st) {
  long long total = (long long)C * H * W;
  ps_to_chw_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, C, H, W, img, out);
}

// debug: dense NCHW fp32 -> PS tensor (zero halo, split)
__global__ void chw_to_ps_kernel(const float* __restrict__ in, int C, int H, int W, int nimg, float* __restrict__ out) {
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  const int Hp = H + 2, Wp = W + 2, c4 = C >> 2;
  if (t >= (long long)nimg * Hp * Wp * c4) return;
  const int c = (int)(t % c4) * 4;
  const long long m = t / c4;
  const int img = (int)(m / (Hp * Wp));
  const int r = (int)(m % (Hp * Wp));
  const int py = r / Wp, px = r % Wp;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (py >= 1 && py <= H && px >= 1 && px <= W)
    for (int i = 0; i < 4; ++i) v[i] = in[(((long long)img * C + c + i) * H + py - 1) * W + px - 1];
  ps_store4(out + m * ps_row_floats(C), c, make_float4(v[0], v[1], v[2], v[3]));
}

void launch_chw_to_ps(const float* in, int C, int H, int W, int nimg, float* out, cudaStream_t st) {
  long long total = (long long)nimg * (H + 2) * (W + 2) * (C / 4);
  chw_to_ps_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, C, H, W, nimg, out);
}

// =============================================================================================
// Space-to-depth repack for stride-2 3x3 convolutions (HRNet transitions / fuse downsample paths):
//   S[a'][b'][(py,px,c)] = in_padded[2(a'-1)+py][2(b'-1)+px][c]   on the OUTPUT's padded grid,
// so that  out[oy][ox] = sum_{dy,dx in {0,1}} S[oy+1+dy][ox+1+dx] . W'[dy][dx]  with
// W'[dy][dx][(py,px,c)] = w[2dy+py][2dx+px][c] (zero when the tap index exceeds 2): a 2x2, stride-1, tap-shifted
// GEMM over 4*C channels that runs on the tensor-core kernel.  Pure data movement: hi/lo pairs are copied.
// =============================================================================================
__global__ void __launch_bounds__(256) s2d_kernel(const float* __restrict__ in, int C, int Hin, int Win, int nimg,
                                                  float* __restrict__ out, int Hout, int Wout) {
  const int C4 = 4 * C, g4 = C4 >> 2;
  const int Hp = Hout + 2, Wp = Wout + 2, HpI = Hin + 2, WpI = Win + 2;
  const long long total = (long long)nimg * Hp * Wp * g4;
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t >= total) return;
  const int ce = (int)(t % g4) * 4;            // channel of the 4C-channel tensor
  const long long m = t / g4;
  const int img = (int)(m / (Hp * Wp));
  const int r = (int)(m % (Hp * Wp));
  const int ap = r / Wp, bp = r % Wp;
  const int par = ce / C, c = ce % C;
  const int py = par >> 1, px = par & 1;
  const int sy = 2 * (ap - 1) + py, sx = 2 * (bp - 1) + px;
  float* drow = out + m * ps_row_floats(C4);
  if (ap >= 1 && bp >= 1 && sy < HpI && sx < WpI)
    ps_copy4(drow, ce, in + (((long long)img * HpI + sy) * WpI + sx) * ps_row_floats(C), c);
  else
    ps_zero4(drow, ce);
}

void launch_s2d(const float* in, int C, int Hin, int Win, int nimg, float* out, int Hout, int Wout, cudaStream_t st) {
  long long total = (long long)nimg * (Hout + 2) * (Wout + 2) * C;
  s2d_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, C, Hin, Win, nimg, out, Hout, Wout);
}
