// ViTPose-B kernels around the tensor-core GEMMs (north_star "HRNet/ViTPose backbone ... warp-shuffle reductions for
// BN/LayerNorm"; BASELINE configs[2]; SURVEY kernels K4/K5, App. A.4 -- the model is upstream ViTPose, not in the reference
// tree).  Token matrices are flat PS rows [image * tokens + t][768] (pe_common.cuh split format, no zero border); all the
// Linear layers (patch embedding, qkv, proj, fc1, fc2) run on conv_tc.cu as TC_KIND_LIN1 GEMMs with bias / GELU / residual
// epilogues.  Here: the patch gather, LayerNorm, the 192-token attention, and the depth-to-space of the deconvolution head.
#include <cmath>

#include "kernels.h"
#include "pe_common.cuh"

// ---------------------------------------------------------------------------------------------
// PatchEmbed input: Conv2d(3, 768, k16, s16, p2) == GEMM over rows of 768 = (c, ky, kx) values.  One thread writes 4
// consecutive k of one token row: pixels (ty*16 - pad + ky, tx*16 - pad + kx) of the crop through the normalisation LUT,
// zero outside the image (the padding applies to the normalised tensor).  Images [ncrop, nimg) are the horizontally
// flipped twins of crops [0, ncrop) (flip test).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) patchify_kernel(const uint8_t* __restrict__ crops, int ncrop, int nimg, int ih, int iw,
                                                       const float* __restrict__ lut, int patch, int pad, int th, int tw,
                                                       float* __restrict__ out) {
  const int K = 3 * patch * patch, K4 = K >> 2;
  const long long total = (long long)nimg * th * tw * K4;
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t >= total) return;
  const int k0 = (int)(t % K4) * 4;
  const long long row = t / K4;
  const int img = (int)(row / (th * tw)), tok = (int)(row % (th * tw)), ty = tok / tw, tx = tok % tw;
  const bool flip = img >= ncrop;
  const uint8_t* src = crops + (size_t)(flip ? img - ncrop : img) * ih * iw * 3;
  float v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = k0 + i, c = k / (patch * patch), r = k % (patch * patch), ky = r / patch, kx = r % patch;
    const int y = ty * patch - pad + ky;
    int x = tx * patch - pad + kx;
    float val = 0.f;
    if (y >= 0 && y < ih && x >= 0 && x < iw) {
      if (flip) x = iw - 1 - x;
      val = lut[c * 256 + src[((size_t)y * iw + x) * 3 + c]];
    }
    v[i] = val;
  }
  ps_store4(out + row * ps_row_floats(K), k0, make_float4(v[0], v[1], v[2], v[3]));
}

void launch_patchify(const uint8_t* crops, int ncrop, int nimg, int ih, int iw, const float* lut, int patch, int pad, int th, int tw, float* out,
                     cudaStream_t st) {
  const long long total = (long long)nimg * th * tw * (3 * patch * patch / 4);
  patchify_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(crops, ncrop, nimg, ih, iw, lut, patch, pad, th, tw, out);
}

// tile a [tokens][C] fp32 table over nimg images as PS rows (the position embedding, added through the GEMM's residual path)
__global__ void __launch_bounds__(256) tile_rows_kernel(const float* __restrict__ table, int tokens, int C, int nimg, float* __restrict__ out) {
  const int C4 = C >> 2;
  const long long total = (long long)nimg * tokens * C4;
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t >= total) return;
  const int c = (int)(t % C4) * 4;
  const long long row = t / C4;
  const float* s = table + (size_t)(row % tokens) * C + c;
  ps_store4(out + row * ps_row_floats(C), c, make_float4(s[0], s[1], s[2], s[3]));
}

void launch_tile_rows(const float* table, int tokens, int C, int nimg, float* out, cudaStream_t st) {
  const long long total = (long long)nimg * tokens * (C / 4);
  tile_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(table, tokens, C, nimg, out);
}

// ---------------------------------------------------------------------------------------------
// LayerNorm(C, eps): one warp per token row, warp-shuffle reductions (two-pass mean / variance in fp32 like torch's CPU
// kernel), gamma / beta, output as PS rows.  grid2d = 0: output row = input row.  grid2d = 1 (the final norm feeding the
// convolutional head): the launch runs over the padded (H+2)x(W+2) grid of each image, interior positions take token
// (y*W + x), border positions are written as zeros.
// ---------------------------------------------------------------------------------------------
template <int CPL>   // channels per lane = C / 32
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ in, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float eps, long long rows_out, int tokens, int H, int W, int grid2d, float* __restrict__ out) {
  constexpr int C = CPL * 32;
  const long long orow = ((long long)blockIdx.x * 256 + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (orow >= rows_out) return;
  float* o = out + orow * ps_row_floats(C);
  long long irow = orow;
  if (grid2d) {
    const int Hp = H + 2, Wp = W + 2;
    const int img = (int)(orow / (Hp * Wp)), r = (int)(orow % (Hp * Wp)), py = r / Wp, px = r % Wp;
    if (py < 1 || py > H || px < 1 || px > W) {
      for (int c = lane * 4; c < C; c += 128) ps_zero4(o, c);
      return;
    }
    irow = (long long)img * tokens + (py - 1) * W + (px - 1);
  }
  const float* x = in + irow * ps_row_floats(C);
  float v[CPL];
#pragma unroll
  for (int i = 0; i < CPL / 4; ++i) {
    const float4 a = ps_load4(x, lane * 4 + i * 128);
    v[4 * i] = a.x; v[4 * i + 1] = a.y; v[4 * i + 2] = a.z; v[4 * i + 3] = a.w;
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CPL; ++i) s += v[i];
#pragma unroll
  for (int o2 = 16; o2 > 0; o2 >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o2);
  const float mean = s * (1.0f / C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < CPL; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
#pragma unroll
  for (int o2 = 16; o2 > 0; o2 >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o2);
  const float rstd = 1.0f / sqrtf(q * (1.0f / C) + eps);
#pragma unroll
  for (int i = 0; i < CPL / 4; ++i) {
    const int c = lane * 4 + i * 128;
    const float4 g = *reinterpret_cast<const float4*>(gamma + c), b = *reinterpret_cast<const float4*>(beta + c);
    ps_store4(o, c, make_float4(fmaf((v[4 * i] - mean) * rstd, g.x, b.x), fmaf((v[4 * i + 1] - mean) * rstd, g.y, b.y),
                                fmaf((v[4 * i + 2] - mean) * rstd, g.z, b.z), fmaf((v[4 * i + 3] - mean) * rstd, g.w, b.w)));
  }
}

cudaError_t launch_layernorm(const float* in, const float* gamma, const float* beta, float eps, int C, int nimg, int tokens, int H, int W, int grid2d,
                             float* out, cudaStream_t st) {
  const long long rows = grid2d ? (long long)nimg * (H + 2) * (W + 2) : (long long)nimg * tokens;
  const unsigned blocks = (unsigned)((rows * 32 + 255) / 256);
  if (C == 768) layernorm_kernel<24><<<blocks, 256, 0, st>>>(in, gamma, beta, eps, rows, tokens, H, W, grid2d, out);
  else if (C == 1024) layernorm_kernel<32><<<blocks, 256, 0, st>>>(in, gamma, beta, eps, rows, tokens, H, W, grid2d, out);
  else if (C == 384) layernorm_kernel<12><<<blocks, 256, 0, st>>>(in, gamma, beta, eps, rows, tokens, H, W, grid2d, out);
  else return cudaErrorNotSupported;
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Multi-head self-attention over the T tokens of one image (T = 192, head dim 64): softmax((q * scale) k^T) v in fp32.
// One CTA per (image, head): K and V of the head live in shared memory; each warp takes query rows round-robin, lanes split
// the keys for the scores (shuffle max / sum) and the head dimension for the output.  0.6 % of the model's FLOPs.
// qkv rows: [3][heads][64] channels (reshape(B, N, 3, heads, 64) of the qkv Linear); out rows: [heads][64].
// ---------------------------------------------------------------------------------------------
constexpr int ATT_D = 64;
__global__ void __launch_bounds__(256) attention_kernel(const float* __restrict__ qkv, int T, int heads, float scale, float* __restrict__ out) {
  extern __shared__ float sm[];
  const int C = heads * ATT_D;
  float* sK = sm;                          // [T][65]
  float* sV = sm + (size_t)T * 65;         // [T][64]
  float* sP = sV + (size_t)T * 64;         // [8 warps][T]
  float* sQ = sP + 8 * T;                  // [8 warps][64]
  const int img = blockIdx.x / heads, head = blockIdx.x % heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row0 = (long long)img * T;
  const int rowF = ps_row_floats(3 * C);
  for (int i = threadIdx.x; i < T * (ATT_D / 4); i += 256) {
    const int t = i / (ATT_D / 4), d = (i % (ATT_D / 4)) * 4;
    const float* r = qkv + (row0 + t) * rowF;
    const float4 k4 = ps_load4(r, C + head * ATT_D + d), v4 = ps_load4(r, 2 * C + head * ATT_D + d);
    sK[t * 65 + d] = k4.x; sK[t * 65 + d + 1] = k4.y; sK[t * 65 + d + 2] = k4.z; sK[t * 65 + d + 3] = k4.w;
    *reinterpret_cast<float4*>(sV + t * 64 + d) = v4;
  }
  __syncthreads();
  float* myP = sP + warp * T;
  float* myQ = sQ + warp * ATT_D;
  for (int t = warp; t < T; t += 8) {
    const float* r = qkv + (row0 + t) * rowF;
    if (lane < 16) {
      const float4 q4 = ps_load4(r, head * ATT_D + lane * 4);
      myQ[lane * 4] = q4.x * scale; myQ[lane * 4 + 1] = q4.y * scale; myQ[lane * 4 + 2] = q4.z * scale; myQ[lane * 4 + 3] = q4.w * scale;
    }
    __syncwarp();
    float mx = -INFINITY;
    for (int j = lane; j < T; j += 32) {
      float s = 0.f;
#pragma unroll 16
      for (int d = 0; d < ATT_D; ++d) s = fmaf(myQ[d], sK[j * 65 + d], s);
      myP[j] = s;
      mx = fmaxf(mx, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int j = lane; j < T; j += 32) {
      const float e = expf(myP[j] - mx);
      myP[j] = e;
      sum += e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncwarp();
    const float inv = 1.0f / sum;
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < T; ++j) {
      const float p = myP[j];
      o0 = fmaf(p, sV[j * 64 + lane], o0);
      o1 = fmaf(p, sV[j * 64 + 32 + lane], o1);
    }
    o0 *= inv; o1 *= inv;
    // lanes hold dims (lane, lane+32): regroup to 4 consecutive channels per lane for the split store
    const float a0 = __shfl_sync(0xffffffffu, o0, (lane & 7) * 4 + 0), a1 = __shfl_sync(0xffffffffu, o0, (lane & 7) * 4 + 1);
    const float a2 = __shfl_sync(0xffffffffu, o0, (lane & 7) * 4 + 2), a3 = __shfl_sync(0xffffffffu, o0, (lane & 7) * 4 + 3);
    const float b0 = __shfl_sync(0xffffffffu, o1, (lane & 7) * 4 + 0), b1 = __shfl_sync(0xffffffffu, o1, (lane & 7) * 4 + 1);
    const float b2 = __shfl_sync(0xffffffffu, o1, (lane & 7) * 4 + 2), b3 = __shfl_sync(0xffffffffu, o1, (lane & 7) * 4 + 3);
    float* orow = out + (row0 + t) * ps_row_floats(C);
    if (lane < 8) ps_store4(orow, head * ATT_D + lane * 4, make_float4(a0, a1, a2, a3));
    else if (lane < 16) ps_store4(orow, head * ATT_D + 32 + (lane - 8) * 4, make_float4(b0, b1, b2, b3));
    __syncwarp();
  }
}

// Register-tiled form (T = 64 * NI tokens, NI <= 3; head dim 64): one CTA per (image, head), 256 threads.  K (transposed),
// V and a 64-row query block live in shared memory; per query block
//   A. S = (q * scale) k^T : each thread a 4-row x (4 * NI)-column tile, per head dimension one float4 of q (broadcast) and NI
//      float4 of k feed 16 * NI FMAs (the row-wise kernel above issued two LDS per FMA: 83 ms of the 160 ms ViTPose-B step),
//   B. row softmax numerators exp(s - max) in place (one warp per row, shuffle max / sum), 1 / sum kept per row,
//   C. O = P v : each thread 4 rows x 4 dims, four keys per step (float4 of p per row, float4 of v per key), scaled by 1 / sum.
// fp32 throughout; only the summation order differs from the row-wise kernel.
template <int NI, int QB>
__global__ void __launch_bounds__(4 * QB, 1) attention_tiled_kernel(const float* __restrict__ qkv, int heads, float scale, float* __restrict__ out) {
  constexpr int T = 64 * NI, TS = T + 4, NT = 4 * QB;         // NT threads = QB / 4 row groups x 16 column groups
  extern __shared__ __align__(16) float sm[];
  float* sKt = sm;                         // [64][T]   k transposed: [d][token]
  float* sV = sKt + 64 * T;                // [T][64]
  float* sQt = sV + T * 64;                // [64][QB]  q * scale transposed: [d][row]
  float* sS = sQt + 64 * QB;               // [QB][TS]  scores, then softmax numerators
  float* sInv = sS + QB * TS;              // [QB]
  const int C = heads * ATT_D;
  const int img = blockIdx.x / heads, head = blockIdx.x - img * heads;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long row0 = (long long)img * T;
  const int rowF = ps_row_floats(3 * C);
  const int kch = (C + head * ATT_D) >> 4, vch = (2 * C + head * ATT_D) >> 4, qch = (head * ATT_D) >> 4;
  for (int i = tid; i < 4 * T; i += NT) {                      // K: consecutive lanes = consecutive tokens (conflict-free transposed stores)
    const int c = i / T, t = i - c * T;
    float v[16];
    chunk_load16(qkv + (row0 + t) * rowF, kch + c, v);
#pragma unroll
    for (int k = 0; k < 16; ++k) sKt[(c * 16 + k) * T + t] = v[k];
  }
  for (int i = tid; i < 4 * T; i += NT) {                      // V: row-major
    const int t = i >> 2, c = i & 3;
    float v[16];
    chunk_load16(qkv + (row0 + t) * rowF, vch + c, v);
    float4* d = reinterpret_cast<float4*>(sV + t * 64 + c * 16);
#pragma unroll
    for (int k = 0; k < 4; ++k) { const int kk = (k + lane) & 3; d[kk] = make_float4(v[4 * kk], v[4 * kk + 1], v[4 * kk + 2], v[4 * kk + 3]); }
  }
  const int rg = tid >> 4, cg = tid & 15;
  for (int qb = 0; qb < T / QB; ++qb) {
    __syncthreads();                                           // K / V visible; previous block's sS / sQt no longer read
    {
      const int i = tid, c = i / QB, r = i - c * QB;           // 4 * QB chunks = one per thread
      float v[16];
      chunk_load16(qkv + (row0 + qb * QB + r) * rowF, qch + c, v);
#pragma unroll
      for (int k = 0; k < 16; ++k) sQt[(c * 16 + k) * QB + r] = v[k] * scale;
    }
    __syncthreads();
    {                                                          // A: scores
      float acc[4][4 * NI];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4 * NI; ++c) acc[r][c] = 0.f;
#pragma unroll 4
      for (int d = 0; d < 64; ++d) {
        const float4 q4 = *reinterpret_cast<const float4*>(sQt + d * QB + 4 * rg);
        const float q[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          const float4 k4 = *reinterpret_cast<const float4*>(sKt + d * T + 64 * i + 4 * cg);
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            acc[r][4 * i + 0] = fmaf(q[r], k4.x, acc[r][4 * i + 0]); acc[r][4 * i + 1] = fmaf(q[r], k4.y, acc[r][4 * i + 1]);
            acc[r][4 * i + 2] = fmaf(q[r], k4.z, acc[r][4 * i + 2]); acc[r][4 * i + 3] = fmaf(q[r], k4.w, acc[r][4 * i + 3]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < NI; ++i)
          *reinterpret_cast<float4*>(sS + (4 * rg + r) * TS + 64 * i + 4 * cg) = make_float4(acc[r][4 * i], acc[r][4 * i + 1], acc[r][4 * i + 2], acc[r][4 * i + 3]);
    }
    __syncthreads();
    for (int r = warp * 8; r < warp * 8 + 8; ++r) {            // B: softmax numerators
      float* srow = sS + r * TS;
      float v[2 * NI];
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < 2 * NI; ++i) { v[i] = srow[lane + 32 * i]; mx = fmaxf(mx, v[i]); }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 2 * NI; ++i) { const float e = expf(v[i] - mx); srow[lane + 32 * i] = e; sum += e; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (lane == 0) sInv[r] = 1.0f / sum;
    }
    __syncthreads();
    {                                                          // C: O = P v
      float acc[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r) { acc[r][0] = 0.f; acc[r][1] = 0.f; acc[r][2] = 0.f; acc[r][3] = 0.f; }
#pragma unroll 2
      for (int j = 0; j < T; j += 4) {
        float4 p4[4], v4[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) p4[r] = *reinterpret_cast<const float4*>(sS + (4 * rg + r) * TS + j);
#pragma unroll
        for (int k = 0; k < 4; ++k) v4[k] = *reinterpret_cast<const float4*>(sV + (j + k) * 64 + 4 * cg);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float p[4] = {p4[r].x, p4[r].y, p4[r].z, p4[r].w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            acc[r][0] = fmaf(p[k], v4[k].x, acc[r][0]); acc[r][1] = fmaf(p[k], v4[k].y, acc[r][1]);
            acc[r][2] = fmaf(p[k], v4[k].z, acc[r][2]); acc[r][3] = fmaf(p[k], v4[k].w, acc[r][3]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float inv = sInv[4 * rg + r];
        ps_store4(out + (row0 + qb * QB + 4 * rg + r) * ps_row_floats(C), head * ATT_D + 4 * cg,
                  make_float4(acc[r][0] * inv, acc[r][1] * inv, acc[r][2] * inv, acc[r][3] * inv));
      }
    }
  }
}

template <int NI, int QB>
static cudaError_t launch_attention_tiled(const float* qkv, int nimg, int heads, float scale, float* out, cudaStream_t st) {
  constexpr int T = 64 * NI;
  static_assert(T % QB == 0 && QB % 32 == 0, "query blocks tile the tokens; one warp per 8 rows");
  const size_t smem = sizeof(float) * (64 * T + T * 64 + 64 * QB + QB * (T + 4) + QB);
  cudaError_t e = pe_smem_optin((const void*)attention_tiled_kernel<NI, QB>, (int)smem);
  if (e != cudaSuccess) return e;
  attention_tiled_kernel<NI, QB><<<(unsigned)(nimg * heads), 4 * QB, smem, st>>>(qkv, heads, scale, out);
  return cudaGetLastError();
}

cudaError_t launch_attention(const float* qkv, int nimg, int T, int heads, float scale, float* out, cudaStream_t st) {
  static const bool rowwise = getenv("PE_ATT_ROWWISE") && atoi(getenv("PE_ATT_ROWWISE"));      // A/B knob: the round-2 first version
  if (!rowwise) {
    static const int qb = getenv("PE_ATT_QB") ? atoi(getenv("PE_ATT_QB")) : 96;                // 96: 12 warps per SM; 64: the first tiled form (8 warps)
    if (T == 192) return qb == 64 ? launch_attention_tiled<3, 64>(qkv, nimg, heads, scale, out, st) : launch_attention_tiled<3, 96>(qkv, nimg, heads, scale, out, st);
    if (T == 128) return launch_attention_tiled<2, 64>(qkv, nimg, heads, scale, out, st);
    if (T == 64) return launch_attention_tiled<1, 64>(qkv, nimg, heads, scale, out, st);
  }
  const size_t smem = sizeof(float) * ((size_t)T * 65 + (size_t)T * 64 + 8 * (size_t)T + 8 * ATT_D);
  cudaError_t e = pe_smem_optin((const void*)attention_kernel, (int)smem);
  if (e != cudaSuccess) return e;
  attention_kernel<<<(unsigned)(nimg * heads), 256, smem, st>>>(qkv, T, heads, scale, out);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Depth-to-space of the deconvolution head: ConvTranspose2d(k4, s2, p1) is computed as ONE 3x3 convolution producing the
// four output parities as channel blocks [(py*2+px)*C + c] at input resolution (engine.deconv_as_conv_weights); this kernel
// interleaves them into the (2H)x(2W) padded grid: out[2a+py][2b+px][c] = in[a][b][(py*2+px)*C + c].  Pure chunk copies.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) d2s_kernel(const float* __restrict__ in, int C, int H, int W, int nimg, float* __restrict__ out) {
  const int nch = C >> 4, Ho = 2 * H, Wo = 2 * W, Hp = Ho + 2, Wp = Wo + 2;
  const long long total = (long long)nimg * Hp * Wp * nch;
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t >= total) return;
  const int chunk = (int)(t % nch);
  const long long m = t / nch;
  const int img = (int)(m / (Hp * Wp)), r = (int)(m % (Hp * Wp)), py = r / Wp, px = r % Wp;
  uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<char*>(out + m * ps_row_floats(C)) + (size_t)chunk * PS_CHUNK_BYTES);
  if (py < 1 || py > Ho || px < 1 || px > Wo) {
#pragma unroll
    for (int i = 0; i < PS_CHUNK_BYTES / 16; ++i) op[i] = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  const int oy = py - 1, ox = px - 1, par = (oy & 1) * 2 + (ox & 1);
  const long long srow = ((long long)img * (H + 2) + (oy >> 1) + 1) * (W + 2) + (ox >> 1) + 1;
  const uint4* ip = reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(in + srow * ps_row_floats(4 * C)) + (size_t)(par * nch + chunk) * PS_CHUNK_BYTES);
#pragma unroll
  for (int i = 0; i < PS_CHUNK_BYTES / 16; ++i) op[i] = __ldg(ip + i);
}

void launch_d2s(const float* in, int C, int H, int W, int nimg, float* out, cudaStream_t st) {
  const long long total = (long long)nimg * (2 * H + 2) * (2 * W + 2) * (C / 16);
  d2s_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, C, H, W, nimg, out);
}
