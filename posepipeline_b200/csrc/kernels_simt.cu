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
// One CTA = a 16-row x 32-column tile of the PADDED output grid of one image; 256 threads, TWO positions per thread (columns
// lx and lx + 16 of one tile row), all 64 output channels in four passes of 16.
//   1. every global load of the CTA (weights, LUT, bias, the thread's <= 9 patch pixels) is issued before the first use; the
//      33 x 65 x 3 input patch goes through the normalisation LUT into shared memory once,
//   2. 27 x 64 FMAs per position, weights read from shared memory as float4 (same order ky,kx,c as the oracle's accumulation;
//      out-of-image taps contribute fmaf(0, w, acc) = acc, so padding is exact); every weight load feeds both positions.
//      Measured (profiles/r02_ncu_full_simt.md): the one-position form ran 618 M shared-memory wavefronts per launch (82 % of
//      the LSU cycles) in 2.62 ms; halving the weight loads this way did NOT shorten it (2.72 -> 2.78 ms per 512 images, 128
//      registers with 44 bytes of spills, still two CTAs per SM) -- the kernel stays at ~18 TFLOP/s fp32, 2 % of a forward,
//   3. each pass's 16-channel chunks are staged in shared memory (XOR-swizzled 16-byte units) and leave as whole chunks
//      (two positions = 128 contiguous staged bytes per quarter warp; 32 KB of staging instead of 64 KB: three CTAs per SM).
constexpr int STEM_TH = 16, STEM_TW = 32;                    // tile height / width (positions)
constexpr int STEM_PH = 2 * STEM_TH + 1, STEM_PW = 2 * STEM_TW + 1;   // input patch
constexpr int STEM_PWP = STEM_PW + 1;                        // padded patch row
constexpr int STEM_NPOS = STEM_TH * STEM_TW;
constexpr int STEM_IN_FLOATS = (3 * STEM_PH * STEM_PWP + 3) & ~3;      // padded so the staging area behind it stays 16-byte aligned
constexpr size_t STEM_SMEM = sizeof(float) * (27 * 64 + 768 + 64 + STEM_IN_FLOATS) + (size_t)STEM_NPOS * PS_CHUNK_BYTES;

__global__ void __launch_bounds__(256, 2) stem_kernel(const uint8_t* __restrict__ crops, int ncrop, int nimg, int ih, int iw,
                                                      const float* __restrict__ lut,   // [3][256]
                                                      const float* __restrict__ w,     // [27][64]  (tap-major: (ky*3+kx)*3+c)
                                                      const float* __restrict__ bias,  // [64]
                                                      unsigned int* flag,
                                                      float* __restrict__ out, int oh, int ow) {
  extern __shared__ __align__(16) uint8_t stem_smem[];
  float* s_w = reinterpret_cast<float*>(stem_smem);           // [27][64]
  float* s_lut = s_w + 27 * 64;                               // [768]
  float* s_b = s_lut + 768;                                   // [64]
  float* s_in = s_b + 64;                                     // [3][STEM_PH][STEM_PWP]
  uint4* s_st = reinterpret_cast<uint4*>(s_in + STEM_IN_FLOATS);               // [STEM_NPOS positions][UPC units], swizzled
  constexpr int UPC = PS_CHUNK_BYTES / 16;                     // 16-byte units per 16-channel chunk
  const int tid = threadIdx.x;
  const int Hp = oh + 2, Wp = ow + 2;
  const int img = blockIdx.z;
  const int ty0 = blockIdx.y * STEM_TH, tx0 = blockIdx.x * STEM_TW;    // tile origin on the padded grid
  const bool flip = img >= ncrop;
  const uint8_t* c0 = crops + (size_t)(flip ? img - ncrop : img) * ih * iw * 3;
  // input patch: rows iy0 .. iy0 + 32, cols ix0 .. ix0 + 64, where output (py, px) reads input (2(py-1)-1+ky, 2(px-1)-1+kx)
  const int iy0 = 2 * (ty0 - 1) - 1, ix0 = 2 * (tx0 - 1) - 1;
  constexpr int NW = (27 * 64 + 255) / 256, NP = (STEM_PH * STEM_PW + 255) / 256;
  float wv[NW], lv[3], bv = 0.f;
  uint8_t pv[NP][3];
  bool pin[NP];
#pragma unroll
  for (int k = 0; k < NW; ++k) { const int i = tid + 256 * k; wv[k] = i < 27 * 64 ? __ldg(w + i) : 0.f; }
#pragma unroll
  for (int k = 0; k < 3; ++k) lv[k] = __ldg(lut + tid + 256 * k);
  if (tid < 64) bv = __ldg(bias + tid);
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    const int i = tid + 256 * k;
    const int ry = i / STEM_PW, rx = i - ry * STEM_PW;
    const int iy = iy0 + ry, ix = ix0 + rx;
    pin[k] = i < STEM_PH * STEM_PW && iy >= 0 && iy < ih && ix >= 0 && ix < iw;
    pv[k][0] = pv[k][1] = pv[k][2] = 0;
    if (pin[k]) {
      const uint8_t* pix = c0 + ((size_t)iy * iw + (flip ? iw - 1 - ix : ix)) * 3;
      pv[k][0] = __ldg(pix); pv[k][1] = __ldg(pix + 1); pv[k][2] = __ldg(pix + 2);
    }
  }
#pragma unroll
  for (int k = 0; k < NW; ++k) { const int i = tid + 256 * k; if (i < 27 * 64) s_w[i] = wv[k]; }
#pragma unroll
  for (int k = 0; k < 3; ++k) s_lut[tid + 256 * k] = lv[k];
  if (tid < 64) s_b[tid] = bv;
  __syncthreads();                                                      // LUT ready
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    const int i = tid + 256 * k;
    if (i >= STEM_PH * STEM_PW) continue;
    const int ry = i / STEM_PW, rx = i - ry * STEM_PW;
    float v0 = 0.f, v1 = 0.f, v2 = 0.f;
    if (pin[k]) { v0 = s_lut[pv[k][0]]; v1 = s_lut[256 + pv[k][1]]; v2 = s_lut[512 + pv[k][2]]; }
    s_in[(0 * STEM_PH + ry) * STEM_PWP + rx] = v0;
    s_in[(1 * STEM_PH + ry) * STEM_PWP + rx] = v1;
    s_in[(2 * STEM_PH + ry) * STEM_PWP + rx] = v2;
  }
  __syncthreads();
  const int ly = tid >> 4, lx = tid & 15;                               // positions (ly, lx) and (ly, lx + 16) of the tile
  const int py = ty0 + ly;
  bool interior[2];
  int pidx[2];
  float vin[2][27];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int lxx = lx + 16 * h, px = tx0 + lxx;
    interior[h] = py >= 1 && py <= oh && px >= 1 && px <= ow;
    pidx[h] = ly * STEM_TW + lxx;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx)
#pragma unroll
        for (int c = 0; c < 3; ++c) vin[h][(ky * 3 + kx) * 3 + c] = s_in[(c * STEM_PH + 2 * ly + ky) * STEM_PWP + 2 * lxx + kx];
  }
  const size_t img_row0 = (size_t)img * Hp * Wp;
#pragma unroll 1
  for (int q = 0; q < 4; ++q) {                                         // 16 output channels = one PS chunk per pass
    float acc[2][16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { acc[0][i] = s_b[q * 16 + i]; acc[1][i] = acc[0][i]; }
#pragma unroll
    for (int t = 0; t < 27; ++t) {
      const float4* wr = reinterpret_cast<const float4*>(s_w + t * 64 + q * 16);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 w4 = wr[i];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          acc[h][4 * i + 0] = fmaf(vin[h][t], w4.x, acc[h][4 * i + 0]); acc[h][4 * i + 1] = fmaf(vin[h][t], w4.y, acc[h][4 * i + 1]);
          acc[h][4 * i + 2] = fmaf(vin[h][t], w4.z, acc[h][4 * i + 2]); acc[h][4 * i + 3] = fmaf(vin[h][t], w4.w, acc[h][4 * i + 3]);
        }
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint4 o[UPC];
      if (!interior[h]) {
#pragma unroll
        for (int i = 0; i < UPC; ++i) o[i] = make_uint4(0u, 0u, 0u, 0u);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[h][i] = fmaxf(acc[h][i], 0.f);
#pragma unroll
        for (int i = 0; i < 4; ++i) ps_range_check4(make_float4(acc[h][4 * i], acc[h][4 * i + 1], acc[h][4 * i + 2], acc[h][4 * i + 3]), flag);
#if PE_FP16
        uint2 hh[4], ll[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) split4_h(make_float4(acc[h][4 * i], acc[h][4 * i + 1], acc[h][4 * i + 2], acc[h][4 * i + 3]), hh[i], ll[i]);
        o[0] = make_uint4(hh[0].x, hh[0].y, hh[1].x, hh[1].y); o[1] = make_uint4(hh[2].x, hh[2].y, hh[3].x, hh[3].y);
        o[2] = make_uint4(ll[0].x, ll[0].y, ll[1].x, ll[1].y); o[3] = make_uint4(ll[2].x, ll[2].y, ll[3].x, ll[3].y);
#else
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float4 hi, lo;
          split4(make_float4(acc[h][4 * i], acc[h][4 * i + 1], acc[h][4 * i + 2], acc[h][4 * i + 3]), hi, lo);
          o[i] = make_uint4(__float_as_uint(hi.x), __float_as_uint(hi.y), __float_as_uint(hi.z), __float_as_uint(hi.w));
          o[4 + i] = make_uint4(__float_as_uint(lo.x), __float_as_uint(lo.y), __float_as_uint(lo.z), __float_as_uint(lo.w));
        }
#endif
      }
      // unit u of position p sits at p * UPC + (u ^ swizzle(p)): conflict-free for the per-position stores here (8 lanes = 8
      // consecutive positions) and for the unit-major reads of the copy-out below
      const int sw = (UPC == 4) ? ((pidx[h] >> 1) & 3) : (pidx[h] & 7);
#pragma unroll
      for (int i = 0; i < UPC; ++i) s_st[pidx[h] * UPC + (i ^ sw)] = o[i];
    }
    __syncthreads();
    // copy out: chunk q of every position of the tile, 16 bytes per thread and step, consecutive threads = consecutive units
#pragma unroll
    for (int k = 0; k < 2 * UPC; ++k) {
      const int e = tid + 256 * k;
      const int p_ = e / UPC, u = e - p_ * UPC;
      const int y = ty0 + p_ / STEM_TW, x = tx0 + (p_ % STEM_TW);
      if (y >= Hp || x >= Wp) continue;
      const int sw = (UPC == 4) ? ((p_ >> 1) & 3) : (p_ & 7);
      const uint4 v = s_st[p_ * UPC + (u ^ sw)];
      reinterpret_cast<uint4*>(reinterpret_cast<char*>(out) + (img_row0 + (size_t)y * Wp + x) * (4 * PS_CHUNK_BYTES) + (size_t)q * PS_CHUNK_BYTES)[u] = v;
    }
    __syncthreads();                                                    // the staging area is free for the next pass
  }
  (void)nimg;
}

void launch_stem(const uint8_t* crops, int ncrop, int nimg, int ih, int iw, const float* lut, const float* w,
                 const float* bias, float* out, int oh, int ow, cudaStream_t st) {
  if (pe_smem_optin((const void*)stem_kernel, (int)STEM_SMEM) != cudaSuccess) return;   // the launch below then fails and is reported
  dim3 grid((ow + 2 + STEM_TW - 1) / STEM_TW, (oh + 2 + STEM_TH - 1) / STEM_TH, nimg);
  stem_kernel<<<grid, 256, STEM_SMEM, st>>>(crops, ncrop, nimg, ih, iw, lut, w, bias, pe_range_flag(), out, oh, ow);
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
  unsigned int* flag;   // range flag (pe_common.cuh ps_range_check4)
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
    ps_range_check4(v, a.flag);
    ps_store4(orow, c, v);
  }
}

void launch_conv_simt(const float* in, float* out, const float* res, const float* w, const float* bias, int Cin,
                      int Cout, int ks, int stride, int relu, int Hin, int Win, int Hout, int Wout, int nimg,
                      cudaStream_t st) {
  ConvArgs a{in, out, res, w, bias, Cin, Cout, ks, stride, relu, Hin, Win, Hout, Wout, nimg, 0, ks * ks, 0, 0, 0, Cout, 0, pe_range_flag()};
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
  ConvArgs a{in, out, res, w, bias, Cin, Cout, 1, 1, relu, 0, 0, 0, 0, 1, 1, ntaps, dil, res_off, plain_out, cout_real, M, pe_range_flag()};
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
  unsigned int* flag;
};

// Work unit = (padded position, 16-channel chunk): a chunk is PS_CHUNK_BYTES contiguous bytes, so every access is a run of
// 16-byte vectors and consecutive threads touch consecutive chunks (4-channel threads with 8-byte accesses reached 27 % of
// the copy bandwidth).  One CTA per padded output row (img, py), threads over (px, chunk); N_IN is a template parameter so that the loads of all
// branches are in flight before the first addition (the round-1 kernel looped over the branches at run time behind three
// 64-bit divisions per thread: 3.1 TB/s at 96x72).  The sum runs in branch order (the reference's fp32 addition order).
template <int N_IN>
__global__ void __launch_bounds__(256) fuse_kernel(FuseArgs a) {
  const int nch = a.C >> 4;
  const int Hp = a.H + 2, Wp = a.W + 2;
  const int img = blockIdx.x / Hp, py = blockIdx.x - img * Hp;
  const size_t rowB = (size_t)nch * PS_CHUNK_BYTES;
  char* orow0 = reinterpret_cast<char*>(a.out) + (size_t)blockIdx.x * Wp * rowB;
  const int n = Wp * nch;
  const bool yin = py >= 1 && py <= a.H;
  const char* irow0[N_IN];
  int sh[N_IN];
#pragma unroll
  for (int j = 0; j < N_IN; ++j) {                     // first interior position of the source row that feeds output row py
    sh[j] = 31 - __clz(a.up[j]);                       // up in {1, 2, 4, 8}
    const int h = a.H >> sh[j], w = a.W >> sh[j];
    irow0[j] = reinterpret_cast<const char*>(a.in[j]) + (((size_t)img * (h + 2) + ((py - 1) >> sh[j]) + 1) * (w + 2) + 1) * rowB;
  }
  for (int i = threadIdx.x; i < n; i += 256) {
    const int px = i / nch, chunk = i - px * nch;
    float* orow = reinterpret_cast<float*>(orow0 + (size_t)px * rowB);
    if (!yin || px < 1 || px > a.W) {
      uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<char*>(orow) + (size_t)chunk * PS_CHUNK_BYTES);
#pragma unroll
      for (int k = 0; k < PS_CHUNK_BYTES / 16; ++k) op[k] = make_uint4(0u, 0u, 0u, 0u);
      continue;
    }
    float v[N_IN][16];
#pragma unroll
    for (int j = 0; j < N_IN; ++j)
      chunk_load16(reinterpret_cast<const float*>(irow0[j] + (size_t)((px - 1) >> sh[j]) * rowB), chunk, v[j]);
    float s[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) s[k] = 0.f;
#pragma unroll
    for (int j = 0; j < N_IN; ++j)
#pragma unroll
      for (int k = 0; k < 16; ++k) s[k] += v[j][k];
    if (a.relu) {
#pragma unroll
      for (int k = 0; k < 16; ++k) s[k] = fmaxf(s[k], 0.f);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) ps_range_check4(make_float4(s[4 * k], s[4 * k + 1], s[4 * k + 2], s[4 * k + 3]), a.flag);
    chunk_store16(orow, chunk, s);
  }
}

void launch_fuse(const float* const* in, const int* up, int n_in, float* out, int C, int H, int W, int nimg, int relu,
                 cudaStream_t st) {
  FuseArgs a;
  for (int j = 0; j < 4; ++j) { a.in[j] = j < n_in ? in[j] : nullptr; a.up[j] = j < n_in ? up[j] : 1; }
  a.n_in = n_in; a.out = out; a.C = C; a.H = H; a.W = W; a.nimg = nimg; a.relu = relu; a.flag = pe_range_flag();
  const unsigned grid = (unsigned)(nimg * (H + 2));
  switch (n_in) {
    case 1: fuse_kernel<1><<<grid, 256, 0, st>>>(a); break;
    case 2: fuse_kernel<2><<<grid, 256, 0, st>>>(a); break;
    case 3: fuse_kernel<3><<<grid, 256, 0, st>>>(a); break;
    default: fuse_kernel<4><<<grid, 256, 0, st>>>(a); break;
  }
}

// =============================================================================================
// Head: final_layer 1x1 conv Cin->K with bias (cfg :73-79) -> planar fp32 heatmaps [img][K][H][W].
// =============================================================================================
// One thread = two horizontally adjacent positions x KB output channels: every weight read from shared memory (a broadcast
// LDS) feeds two FMAs, the input rows arrive as whole 16-channel chunks (16-byte loads), and KB = 17 has no predicated-off
// lanes for the COCO head (the round-1 kernel issued 32 FMAs + 32 LDS per input channel for 17 joints: 1.3 ms per forward).
// Accumulation order over the input channels is unchanged (c ascending, bias added last).
template <int KB>
__global__ void __launch_bounds__(128) head_kernel(const float* __restrict__ in, int Cin, int H, int W, int nimg,
                                                   const float* __restrict__ w,  // [Cin][K]
                                                   const float* __restrict__ bias, int K, float* __restrict__ out) {
  extern __shared__ float s_w[];  // Cin*K + K
  for (int i = threadIdx.x; i < Cin * K; i += 128) s_w[i] = w[i];
  for (int i = threadIdx.x; i < K; i += 128) s_w[Cin * K + i] = bias[i];
  __syncthreads();
  const int HW = H * W;                                   // even (W is even: checked by the launcher)
  const long long t = ((long long)blockIdx.x * 128 + threadIdx.x) * 2;
  if (t >= (long long)nimg * HW) return;
  const int img = (int)(t / HW);
  const int r = (int)(t - (long long)img * HW);
  const int y = r / W, x = r - y * W;                     // x even: x + 1 is in the same image row
  const float* row = in + (((long long)img * (H + 2) + y + 1) * (W + 2) + x + 1) * ps_row_floats(Cin);
  const int rowF = ps_row_floats(Cin);
  for (int k0 = 0; k0 < K; k0 += KB) {
    float acc0[KB], acc1[KB];
    const int kn = min(KB, K - k0);
#pragma unroll
    for (int k = 0; k < KB; ++k) { acc0[k] = 0.f; acc1[k] = 0.f; }
    for (int ch = 0; ch < (Cin >> 4); ++ch) {
      float v0[16], v1[16];
      chunk_load16(row, ch, v0);
      chunk_load16(row + rowF, ch, v1);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float* wr = s_w + (ch * 16 + i) * K + k0;
#pragma unroll
        for (int k = 0; k < KB; ++k)
          if (KB == 17 || k < kn) { const float wk = wr[k]; acc0[k] = fmaf(v0[i], wk, acc0[k]); acc1[k] = fmaf(v1[i], wk, acc1[k]); }
      }
    }
#pragma unroll
    for (int k = 0; k < KB; ++k)
      if (KB == 17 || k < kn) {
        const float b = s_w[Cin * K + k0 + k];
        *reinterpret_cast<float2*>(out + (((long long)img * K + k0 + k) * H + y) * W + x) = make_float2(acc0[k] + b, acc1[k] + b);
      }
  }
}

// one position per thread: odd widths (not used by the shipped models)
__global__ void __launch_bounds__(128) head_kernel_1(const float* __restrict__ in, int Cin, int H, int W, int nimg,
                                                     const float* __restrict__ w, const float* __restrict__ bias, int K, float* __restrict__ out) {
  extern __shared__ float s_w[];
  for (int i = threadIdx.x; i < Cin * K; i += 128) s_w[i] = w[i];
  for (int i = threadIdx.x; i < K; i += 128) s_w[Cin * K + i] = bias[i];
  __syncthreads();
  const long long t = (long long)blockIdx.x * 128 + threadIdx.x;
  if (t >= (long long)nimg * H * W) return;
  const int img = (int)(t / (H * W));
  const int r = (int)(t % (H * W));
  const int y = r / W, x = r % W;
  const float* row = in + (((long long)img * (H + 2) + y + 1) * (W + 2) + x + 1) * ps_row_floats(Cin);
  for (int k = 0; k < K; ++k) {
    float acc = 0.f;
    for (int c = 0; c < Cin; c += 4) {
      const float4 v = ps_load4(row, c);
      acc = fmaf(v.x, s_w[c * K + k], acc); acc = fmaf(v.y, s_w[(c + 1) * K + k], acc);
      acc = fmaf(v.z, s_w[(c + 2) * K + k], acc); acc = fmaf(v.w, s_w[(c + 3) * K + k], acc);
    }
    out[(((long long)img * K + k) * H + y) * W + x] = acc + s_w[Cin * K + k];
  }
}

void launch_head(const float* in, int Cin, int H, int W, int nimg, const float* w, const float* bias, int K, float* out,
                 cudaStream_t st) {
  const long long total = (long long)nimg * H * W;
  const size_t smem = (size_t)(Cin * K + K) * sizeof(float);
  if (W % 2) { head_kernel_1<<<(unsigned)((total + 127) / 128), 128, smem, st>>>(in, Cin, H, W, nimg, w, bias, K, out); return; }
  const unsigned grid = (unsigned)((total / 2 + 127) / 128);
  if (K == 17) head_kernel<17><<<grid, 128, smem, st>>>(in, Cin, H, W, nimg, w, bias, K, out);
  else head_kernel<32><<<grid, 128, smem, st>>>(in, Cin, H, W, nimg, w, bias, K, out);
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
// One CTA per padded output row (img, a'), threads over (b', 16-byte unit of the 4C-channel row): consecutive threads copy
// consecutive 16-byte units, so a warp writes 512 contiguous bytes and reads runs of C/16 chunks from four source rows.
// (The first version moved 8 bytes per thread behind three 64-bit divisions: 0.33 ms for a 0.75 GB copy.)
__global__ void __launch_bounds__(256) s2d_kernel(const float* __restrict__ in, int C, int Hin, int Win, int nimg,
                                                  float* __restrict__ out, int Hout, int Wout) {
  constexpr int UPC = PS_CHUNK_BYTES / 16;                     // 16-byte units per 16-channel chunk
  const int Hp = Hout + 2, Wp = Wout + 2, HpI = Hin + 2, WpI = Win + 2;
  const int upp = (C >> 4) * UPC;                              // units per parity block = units of a source row
  const int upr = 4 * upp;                                     // units per output row
  const int img = blockIdx.x / Hp, ap = blockIdx.x - img * Hp;
  const uint4* __restrict__ src = reinterpret_cast<const uint4*>(in) + (size_t)img * HpI * WpI * upp;
  uint4* __restrict__ dst = reinterpret_cast<uint4*>(out) + (size_t)blockIdx.x * Wp * upr;
  const int n = Wp * upr;
  for (int i = threadIdx.x; i < n; i += 256) {
    const int bp = i / upr, u = i - bp * upr;
    const int par = u / upp, k = u - par * upp;
    const int sy = 2 * (ap - 1) + (par >> 1), sx = 2 * (bp - 1) + (par & 1);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (ap >= 1 && bp >= 1 && sy < HpI && sx < WpI) v = __ldg(src + ((size_t)sy * WpI + sx) * upp + k);
    dst[i] = v;
  }
  (void)nimg;
}

void launch_s2d(const float* in, int C, int Hin, int Win, int nimg, float* out, int Hout, int Wout, cudaStream_t st) {
  s2d_kernel<<<(unsigned)(nimg * (Hout + 2)), 256, 0, st>>>(in, C, Hin, Win, nimg, out, Hout, Wout);
}

// =============================================================================================
// Detector (YOLOX) data-movement kernels.  All work on 16-channel chunks of PS rows; `Ctot` / `coff` select a channel
// slice of a wider tensor (concatenations are never materialised: producers write into slices of the concat tensor).
// =============================================================================================
// space-to-depth of a channel slice (stride-2 3x3 convolutions whose input cannot be TMA-gathered)
__global__ void __launch_bounds__(256) s2d_slice_kernel(const float* __restrict__ in, int C, int Ctot, int coff, int Hin, int Win, int nimg,
                                                        float* __restrict__ out, int Hout, int Wout) {
  const int C4 = 4 * C, g4 = C4 >> 2;
  const int Hp = Hout + 2, Wp = Wout + 2, HpI = Hin + 2, WpI = Win + 2;
  const long long total = (long long)nimg * Hp * Wp * g4;
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t >= total) return;
  const int ce = (int)(t % g4) * 4;
  const long long m = t / g4;
  const int img = (int)(m / (Hp * Wp));
  const int r = (int)(m % (Hp * Wp));
  const int ap = r / Wp, bp = r % Wp;
  const int par = ce / C, c = ce % C;
  const int py = par >> 1, px = par & 1;
  const int sy = 2 * (ap - 1) + py, sx = 2 * (bp - 1) + px;
  float* drow = out + m * ps_row_floats(C4);
  if (ap >= 1 && bp >= 1 && sy < HpI && sx < WpI)
    ps_copy4(drow, ce, in + (((long long)img * HpI + sy) * WpI + sx) * ps_row_floats(Ctot), coff + c);
  else
    ps_zero4(drow, ce);
}

void launch_s2d_slice(const float* in, int C, int Ctot, int coff, int Hin, int Win, int nimg, float* out, int Hout, int Wout, cudaStream_t st) {
  long long total = (long long)nimg * (Hout + 2) * (Wout + 2) * C;
  s2d_slice_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, C, Ctot, coff, Hin, Win, nimg, out, Hout, Wout);
}

// nearest-neighbour x2 upsampling (F.interpolate(scale_factor=2, mode='nearest')): pure copy of chunks, slice -> slice
__global__ void __launch_bounds__(256) upsample2_kernel(const float* __restrict__ in, int C, int in_tot, int in_coff, int H, int W, int nimg,
                                                        float* __restrict__ out, int out_tot, int out_coff) {
  const int nch = C >> 4, Ho = 2 * H, Wo = 2 * W, Hp = Ho + 2, Wp = Wo + 2;
  const long long total = (long long)nimg * Hp * Wp * nch;
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t >= total) return;
  const int chunk = (int)(t % nch);
  const long long m = t / nch;
  const int img = (int)(m / (Hp * Wp)), r = (int)(m % (Hp * Wp)), py = r / Wp, px = r % Wp;
  uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<char*>(out + m * ps_row_floats(out_tot)) + (size_t)(out_coff / 16 + chunk) * PS_CHUNK_BYTES);
  if (py < 1 || py > Ho || px < 1 || px > Wo) {
#pragma unroll
    for (int i = 0; i < PS_CHUNK_BYTES / 16; ++i) op[i] = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  const long long srow = ((long long)img * (H + 2) + (py - 1) / 2 + 1) * (W + 2) + (px - 1) / 2 + 1;
  const uint4* ip = reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(in + srow * ps_row_floats(in_tot)) + (size_t)(in_coff / 16 + chunk) * PS_CHUNK_BYTES);
#pragma unroll
  for (int i = 0; i < PS_CHUNK_BYTES / 16; ++i) op[i] = __ldg(ip + i);
}

void launch_upsample2(const float* in, int C, int in_tot, int in_coff, int H, int W, int nimg, float* out, int out_tot, int out_coff, cudaStream_t st) {
  const long long total = (long long)nimg * (2 * H + 2) * (2 * W + 2) * (C / 16);
  upsample2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, C, in_tot, in_coff, H, W, nimg, out, out_tot, out_coff);
}

// MaxPool2d(k, stride 1, padding k/2) (SPPBottleneck): out-of-image taps are -inf, i.e. ignored; slice -> slice
__global__ void __launch_bounds__(256) maxpool_kernel(const float* __restrict__ in, int C, int tot, int in_coff, int H, int W, int nimg, int k,
                                                      float* __restrict__ out, int out_coff) {
  const int nch = C >> 4, Hp = H + 2, Wp = W + 2;
  const long long total = (long long)nimg * Hp * Wp * nch;
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t >= total) return;
  const int chunk = (int)(t % nch);
  const long long m = t / nch;
  const int img = (int)(m / (Hp * Wp)), r = (int)(m % (Hp * Wp)), py = r / Wp, px = r % Wp;
  float* orow = out + m * ps_row_floats(tot);
  if (py < 1 || py > H || px < 1 || px > W) {
    uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<char*>(orow) + (size_t)(out_coff / 16 + chunk) * PS_CHUNK_BYTES);
#pragma unroll
    for (int i = 0; i < PS_CHUNK_BYTES / 16; ++i) op[i] = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  float best[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) best[i] = -INFINITY;
  const int R = k / 2;
  for (int dy = -R; dy <= R; ++dy) {
    const int y = py + dy;
    if (y < 1 || y > H) continue;
    for (int dx = -R; dx <= R; ++dx) {
      const int x = px + dx;
      if (x < 1 || x > W) continue;
      float v[16];
      chunk_load16(in + (((long long)img * Hp + y) * Wp + x) * ps_row_floats(tot), in_coff / 16 + chunk, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) best[i] = fmaxf(best[i], v[i]);
    }
  }
  chunk_store16(orow, out_coff / 16 + chunk, best);      // re-splitting an exactly representable value reproduces it
}

void launch_maxpool(const float* in, int C, int tot, int in_coff, int H, int W, int nimg, int k, float* out, int out_coff, cudaStream_t st) {
  const long long total = (long long)nimg * (H + 2) * (W + 2) * (C / 16);
  maxpool_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, C, tot, in_coff, H, W, nimg, k, out, out_coff);
}

// ---------------------------------------------------------------------------------------------
// Detector input: cv2.resize(INTER_LINEAR, uint8 fixed point) + Pad(114) + Normalize(0,1) + Focus space-to-depth, fused.
// One thread = one position of the padded (H/2+2) x (W/2+2) Focus grid = 12 values (4 pixels x 3 channels) -> one
// 16-channel PS chunk (channels 12..15 zero).  Channel order of the Focus concat: (top-left, bottom-left, top-right,
// bottom-right) x (R, G, B) where R,G,B are the planes the reference's wrapper hands over (it swaps the BGR frame once,
// pose_pipeline/wrappers/mmtrack.py:43), i.e. BGR frame channel 2 - c.
// xofs/yofs: source index pairs, alpha/beta: 11-bit coefficient pairs (host tables, same arithmetic as OpenCV).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) det_input_kernel(const uint8_t* __restrict__ frames, const int32_t* __restrict__ frame_idx, int fh, int fw,
                                                        int rh, int rw,   // resized size (un-padded)
                                                        int H2, int W2,   // Focus grid = padded size / 2
                                                        const int32_t* __restrict__ xofs, const int16_t* __restrict__ alpha,
                                                        const int32_t* __restrict__ yofs, const int16_t* __restrict__ beta, float pad_val,
                                                        int nimg, float* __restrict__ out) {
  const int Hp = H2 + 2, Wp = W2 + 2;
  const long long total = (long long)nimg * Hp * Wp;
  const long long m = (long long)blockIdx.x * 256 + threadIdx.x;
  if (m >= total) return;
  const int img = (int)(m / (Hp * Wp)), r = (int)(m % (Hp * Wp)), py = r / Wp, px = r % Wp;
  float* orow = out + m * ps_row_floats(16);
  if (py < 1 || py > H2 || px < 1 || px > W2) {
#pragma unroll
    for (int c = 0; c < 16; c += 4) ps_zero4(orow, c);
    return;
  }
  const uint8_t* f = frames + (size_t)frame_idx[img] * fh * fw * 3;
  float v[16];
#pragma unroll
  for (int i = 12; i < 16; ++i) v[i] = 0.f;
#pragma unroll
  for (int q = 0; q < 4; ++q) {                       // Focus order: top-left, bottom-left, top-right, bottom-right
    const int dy = q & 1, dx = q >> 1;
    const int y = 2 * (py - 1) + dy, x = 2 * (px - 1) + dx;
    if (y >= rh || x >= rw) {
      v[3 * q] = v[3 * q + 1] = v[3 * q + 2] = pad_val;
      continue;
    }
    const int sy0 = yofs[2 * y], sy1 = yofs[2 * y + 1], sx0 = xofs[2 * x], sx1 = xofs[2 * x + 1];
    const int b0 = beta[2 * y], b1 = beta[2 * y + 1], a0 = alpha[2 * x], a1 = alpha[2 * x + 1];
    const uint8_t* r0 = f + (size_t)sy0 * fw * 3;
    const uint8_t* r1 = f + (size_t)sy1 * fw * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int cs = 2 - c;                             // R,G,B planes of the BGR frame
      const int S0 = r0[sx0 * 3 + cs] * a0 + r0[sx1 * 3 + cs] * a1;
      const int S1 = r1[sx0 * 3 + cs] * a0 + r1[sx1 * 3 + cs] * a1;
      int pix = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2;
      pix = pix < 0 ? 0 : (pix > 255 ? 255 : pix);
      v[3 * q + c] = (float)pix;
    }
  }
#pragma unroll
  for (int c = 0; c < 16; c += 4) ps_store4(orow, c, make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]));
}

void launch_det_input(const uint8_t* frames, const int32_t* frame_idx, int fh, int fw, int rh, int rw, int H2, int W2, const int32_t* xofs,
                      const int16_t* alpha, const int32_t* yofs, const int16_t* beta, float pad_val, int nimg, float* out, cudaStream_t st) {
  const long long total = (long long)nimg * (H2 + 2) * (W2 + 2);
  det_input_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(frames, frame_idx, fh, fw, rh, rw, H2, W2, xofs, alpha, yofs, beta, pad_val, nimg, out);
}

// ---------------------------------------------------------------------------------------------
// YOLOXHead output convolutions (1x1: cls 1, reg 4, obj 1 channel) + get_bboxes decode of one level, fused:
//   box = ((reg_xy * stride + prior) -/+ exp(reg_wh) * stride / 2) / scale_factor,  score = sigmoid(cls) * sigmoid(obj)
// and every candidate with score >= score_thr is appended to the image's candidate list (prior index kept for a
// deterministic order).  One warp per position: lanes split the input channels, shuffle-reduce the 6 dot products.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) det_head_kernel(const float* __restrict__ cls_feat, const float* __restrict__ reg_feat, int C, int H, int W,
                                                       int nimg, const float* __restrict__ w /*[6][C]: cls, reg x4, obj*/,
                                                       const float* __restrict__ b /*[6]*/, float stride, float4 scale_factor, float score_thr,
                                                       int prior_base, float* __restrict__ cand /*[nimg][cap][6]*/, int* __restrict__ count, int cap,
                                                       float* __restrict__ raw /*optional [nimg][H*W][6] logits*/) {
  const int warp = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const long long total = (long long)nimg * H * W;
  if (warp >= total) return;
  const int img = warp / (H * W), pos = warp % (H * W), y = pos / W, x = pos % W;
  const long long row = ((long long)img * (H + 2) + y + 1) * (W + 2) + x + 1;
  const float* cr = cls_feat + row * ps_row_floats(C);
  const float* rr = reg_feat + row * ps_row_floats(C);
  float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int c = lane * 4; c < C; c += 128) {
    const float4 cv = ps_load4(cr, c), rv = ps_load4(rr, c);
    const float cvv[4] = {cv.x, cv.y, cv.z, cv.w}, rvv[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      acc[0] = fmaf(cvv[i], w[c + i], acc[0]);
#pragma unroll
      for (int k = 1; k < 6; ++k) acc[k] = fmaf(rvv[i], w[k * C + c + i], acc[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < 6; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
  if (lane != 0) return;
#pragma unroll
  for (int k = 0; k < 6; ++k) acc[k] += b[k];
  if (raw) {
    float* o = raw + ((size_t)img * H * W + pos) * 6;
#pragma unroll
    for (int k = 0; k < 6; ++k) o[k] = acc[k];
  }
  const float sc = __fmul_rn(__fdiv_rn(1.0f, 1.0f + expf(-acc[0])), __fdiv_rn(1.0f, 1.0f + expf(-acc[5])));
  if (!(sc >= score_thr)) return;
  const float cx = __fadd_rn(__fmul_rn(acc[1], stride), (float)x * stride), cy = __fadd_rn(__fmul_rn(acc[2], stride), (float)y * stride);
  const float bw = __fmul_rn(expf(acc[3]), stride), bh = __fmul_rn(expf(acc[4]), stride);
  const float hw = __fdiv_rn(bw, 2.0f), hh = __fdiv_rn(bh, 2.0f);
  const int slot = atomicAdd(count + img, 1);
  if (slot >= cap) return;
  float* o = cand + ((size_t)img * cap + slot) * 6;
  o[0] = __fdiv_rn(__fsub_rn(cx, hw), scale_factor.x); o[1] = __fdiv_rn(__fsub_rn(cy, hh), scale_factor.y);
  o[2] = __fdiv_rn(__fadd_rn(cx, hw), scale_factor.z); o[3] = __fdiv_rn(__fadd_rn(cy, hh), scale_factor.w);
  o[4] = sc; o[5] = __int_as_float(prior_base + pos);
}

void launch_det_head(const float* cls_feat, const float* reg_feat, int C, int H, int W, int nimg, const float* w, const float* b, float stride,
                     const float* scale_factor4, float score_thr, int prior_base, float* cand, int* count, int cap, float* raw, cudaStream_t st) {
  const long long warps = (long long)nimg * H * W;
  det_head_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(cls_feat, reg_feat, C, H, W, nimg, w, b, stride,
                                                                         make_float4(scale_factor4[0], scale_factor4[1], scale_factor4[2], scale_factor4[3]),
                                                                         score_thr, prior_base, cand, count, cap, raw);
}
