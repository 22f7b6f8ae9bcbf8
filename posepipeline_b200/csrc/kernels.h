// Launchers shared between the kernel translation units and the C-ABI host code.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

void launch_warp_crop(const uint8_t* frames, int fh, int fw, const int32_t* frame_idx, const double* minv,
                      uint8_t* crops, int n, int ch, int cw, int swap_rb, cudaStream_t st);
void launch_stem(const uint8_t* crops, int ncrop, int nimg, int ih, int iw, const float* lut, const float* w,
                 const float* bias, float* out, int oh, int ow, cudaStream_t st);
void launch_conv_simt(const float* in, float* out, const float* res, const float* w, const float* bias, int Cin,
                      int Cout, int ks, int stride, int relu, int Hin, int Win, int Hout, int Wout, int nimg,
                      cudaStream_t st);
void launch_conv_linear(const float* in, float* out, const float* res, const float* w, const float* bias, int Cin,
                        int Cout, int ntaps, int dil, int relu, long long M, int res_off, int plain_out, int cout_real,
                        cudaStream_t st);
void launch_fuse(const float* const* in, const int* up, int n_in, float* out, int C, int H, int W, int nimg, int relu,
                 cudaStream_t st);
void launch_head(const float* in, int Cin, int H, int W, int nimg, const float* w, const float* bias, int K, float* out,
                 cudaStream_t st);
void launch_ps_to_chw(const float* in, int C, int H, int W, int img, float* out, cudaStream_t st);

void launch_s2d(const float* in, int C, int Hin, int Win, int nimg, float* out, int Hout, int Wout, cudaStream_t st);
// detector (YOLOX) kernels: slices of wider tensors are given as (total channels, first channel)
void launch_s2d_slice(const float* in, int C, int Ctot, int coff, int Hin, int Win, int nimg, float* out, int Hout, int Wout, cudaStream_t st);
void launch_upsample2(const float* in, int C, int in_tot, int in_coff, int H, int W, int nimg, float* out, int out_tot, int out_coff, cudaStream_t st);
void launch_maxpool(const float* in, int C, int tot, int in_coff, int H, int W, int nimg, int k, float* out, int out_coff, cudaStream_t st);
void launch_det_input(const uint8_t* frames, const int32_t* frame_idx, int fh, int fw, int rh, int rw, int H2, int W2, const int32_t* xofs,
                      const int16_t* alpha, const int32_t* yofs, const int16_t* beta, float pad_val, int nimg, float* out, cudaStream_t st);
void launch_det_head(const float* cls_feat, const float* reg_feat, int C, int H, int W, int nimg, const float* w, const float* b, float stride,
                     const float* scale_factor4, float score_thr, int prior_base, float* cand, int* count, int cap, float* raw, cudaStream_t st);
// ViTPose (vit.cu)
void launch_patchify(const uint8_t* crops, int ncrop, int nimg, int ih, int iw, const float* lut, int patch, int pad, int th, int tw, float* out,
                     cudaStream_t st);
void launch_tile_rows(const float* table, int tokens, int C, int nimg, float* out, cudaStream_t st);
cudaError_t launch_layernorm(const float* in, const float* gamma, const float* beta, float eps, int C, int nimg, int tokens, int H, int W, int grid2d,
                             float* out, cudaStream_t st);
cudaError_t launch_attention(const float* qkv, int nimg, int T, int heads, float scale, float* out, cudaStream_t st);
void launch_d2s(const float* in, int C, int H, int W, int nimg, float* out, cudaStream_t st);
void launch_chw_to_ps(const float* in, int C, int H, int W, int nimg, float* out, cudaStream_t st);

// Range flag of the forward being launched on this thread (device word; nullptr = no checking).  The launchers below pass
// it to their kernels, which OR 1 into it when a value does not fit the activation format (fp16x2: |v| > 65504).
unsigned int*& pe_range_flag();

// opt a kernel into `bytes` of dynamic shared memory on the CURRENT device (the attribute is per device; cached per
// (function, device) so the steady state costs one map lookup)
cudaError_t pe_smem_optin(const void* func, int bytes);
cudaError_t launch_decode(const float* hm, const float* hm_flip, const int* flip_perm, const float* center,
                          const float* scale, float* out, int n, int K, int H, int W, int shift, int post, int ksize,
                          const float* gauss /*64 taps on the device, post == 2*/, cudaStream_t st);

// Shifted-row GEMM on tensor cores (conv_tc.cu).  Returns cudaErrorNotSupported for shapes it does not cover.
struct TcConvPlan;
// gather_src != nullptr (ks = 2 only): the 2x2 layer is a stride-2 3x3 convolution whose space-to-depth input is gathered by
// TMA from gather_src (the original [img][2H+2][2W+2][Cin/4] tensor); `in` is then unused and no s2d copy is needed.
cudaError_t tc_conv_plan_create(TcConvPlan** plan, const float* in, float* out, const float* res, const float* wtc,
                                const float* bias, int Cin, int Cout, int ks, int relu, int H, int W, int max_img,
                                const float* gather_src = nullptr);
// General form: any layer kind, operands may be 16-channel-aligned slices of wider tensors.
enum { TC_KIND_1x1 = 1, TC_KIND_2x2 = 2, TC_KIND_3x3 = 3, TC_KIND_LIN1 = 11, TC_KIND_LIN3 = 13 };
struct TcConvDesc {
  int kind;                // TC_KIND_*: 3x3 pad 1 / 1x1 / 2x2 (stride-2 3x3 in space-to-depth form) over padded 2-D tensors;
                           // LIN3 = 3 taps at rows m, m+dil, m+2*dil and LIN1 = plain GEMM over flat row matrices (no zero border)
  int Cin, Cout, act, dil; // act: 0 none, 1 ReLU, 2 SiLU, 3 GELU (erf)
  int H, W;                // 2-D kinds: output (= input) height / width
  long long max_rows;      // rows of the flat output matrix (2-D kinds: max_img * (H+2) * (W+2))
  const float* in; int in_total, in_coff;           // input tensor, its channel count, first channel of the view
  float* out; int out_total, out_coff;
  const float* res; int res_total, res_coff, res_row_off, res_post;   // residual view; row m + res_row_off; added after the activation if res_post
  const float* wtc; const float* bias;
  const float* gather_src; int gather_total, gather_coff;             // 2x2 only: TMA-gather the space-to-depth rows from this tensor
};
cudaError_t tc_conv_plan_create_ex(TcConvPlan** plan, const TcConvDesc* desc);
cudaError_t tc_conv_launch_rows(TcConvPlan* plan, long long rows, cudaStream_t st);
void tc_conv_plan_destroy(TcConvPlan* plan, bool cuda_ok = true);
int tc_plan_candidates(int Cin, int Cout, int ks, int has_res, int H, int W, int max_img, int gather, int32_t* out, int cap);
cudaError_t tc_conv_launch(TcConvPlan* plan, int nimg, cudaStream_t st);
// host views of the persistent schedule (conv_tc_kernel.cuh tc_work_item): group size rule and work item -> (M tile, N slice)
int tc_group_size(int Cin, int Cout, int ns, int tile_rows, int tiles_m, int mode, int budget_kb);
void tc_work_item_host(int tiles_m, int nsplit, int grp, int w, int* tile, int* nsl);
