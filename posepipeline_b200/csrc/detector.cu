// Person detector behind `mmtrack_bounding_boxes(file, "bytetrack")`: YOLOX-X at 800x1440
// (reference config 3rdparty/mmtracking/mot/bytetrack/bytetrack_yolox_x_crowdhuman_mot17-private-half.py:6,9-20,60-81 and
// _base_/models/yolox_x_8x8.py:5-26; reached through mmtrack.apis.inference_mot at pose_pipeline/wrappers/mmtrack.py:45).
//
// The layer program is built on the host (posepipeline_b200/yolox_spec.py) and executed here for a block of staged frames:
//   det_input_kernel   cv2-exact bilinear resize + Pad(114) + Focus space-to-depth, straight from the staged BGR frames
//   conv_tc_kernel     every convolution (3x3 / 1x1 / stride-2 3x3 in 2x2 form) with folded BN + SiLU on tcgen05; channel
//                      concatenations (CSP layers, SPP, PAFPN) are slices of wider tensors, never copies
//   maxpool / upsample2 kernels for the SPP bottleneck and the top-down path
//   det_head_kernel    the 1x1 output convolutions + box decode + score threshold + candidate compaction, per level
// then one D2H copy of the candidate lists and the score-ordered greedy NMS (IoU 0.7) on the host, next to the tracker.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/poseengine.h"
#include "engine_internal.h"
#include "kernels.h"
#include "pe_common.cuh"

struct pe_detector {
  pe_engine* e = nullptr;
  pe_det_desc d{};
  std::vector<pe_gop_desc> ops;
  std::vector<pe_tensor_desc> tensors;
  std::vector<float*> slots;
  std::vector<TcConvPlan*> tc;     // per op
  std::vector<char> tc_s2d;        // per op: the stride-2 plan reads the space-to-depth scratch
  float* d_w = nullptr;
  float* d_s2d = nullptr;
  int32_t *d_xofs = nullptr, *d_yofs = nullptr, *d_fidx = nullptr;
  int16_t *d_alpha = nullptr, *d_beta = nullptr;
  float* d_cand = nullptr; int* d_count = nullptr;
  float* h_cand = nullptr; int* h_count = nullptr; int32_t* h_fidx = nullptr;
  int cap = 0;
  float scale_factor[4];
  int rh = 0, rw = 0;              // resized (un-padded) image size
  int64_t launches = 0;
  std::vector<cudaGraphExec_t> graphs;   // per nimg (index nimg), nullptr = not captured yet
  std::vector<int> graph_seen;
  std::vector<int64_t> graph_launches;
};

#define CUD(x)                                                                                                         \
  do {                                                                                                                 \
    cudaError_t _e = (x);                                                                                              \
    if (_e != cudaSuccess) return pe_fail(PE_ERR_CUDA, "%s failed: %s (%s:%d)", #x, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

static float* act(pe_detector* d, int tid) { return d->slots[d->tensors[tid].slot]; }

// mmcv.imrescale(img, (800, 1440)) target size + cv::resize coefficient tables (see oracle/yolox.py resize_linear_u8)
static void resize_tables(int sn, int dn, bool clamp_weights, std::vector<int32_t>& ofs, std::vector<int16_t>& coef) {
  ofs.resize(2 * dn);
  coef.resize(2 * dn);
  const double inv_scale = (double)dn / sn, scale = 1.0 / inv_scale;
  for (int d = 0; d < dn; ++d) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)std::floor(f);
    f -= s;
    if (clamp_weights) {
      if (s < 0) { f = 0; s = 0; }
      if (s >= sn - 1) { f = 0; s = sn - 1; }
    }
    const float c0 = (1.f - f) * 2048.f, c1 = f * 2048.f;
    coef[2 * d] = (int16_t)std::nearbyint(c0);          // cvRound: to nearest, ties to even (default rounding mode)
    coef[2 * d + 1] = (int16_t)std::nearbyint(c1);
    ofs[2 * d] = std::min(std::max(s, 0), sn - 1);
    ofs[2 * d + 1] = std::min(std::max(s + 1, 0), sn - 1);
  }
}

static void detector_free(pe_detector* d, bool cuda_ok) {
  if (cuda_ok) {
    cudaStreamSynchronize(d->e->stream);
    for (auto g : d->graphs) if (g) cudaGraphExecDestroy(g);
    for (auto* p : d->slots) if (p) cudaFree(p);
    cudaFree(d->d_w); cudaFree(d->d_s2d); cudaFree(d->d_xofs); cudaFree(d->d_yofs); cudaFree(d->d_alpha); cudaFree(d->d_beta);
    cudaFree(d->d_fidx); cudaFree(d->d_cand); cudaFree(d->d_count);
    cudaFreeHost(d->h_cand); cudaFreeHost(d->h_count); cudaFreeHost(d->h_fidx);
    cudaGetLastError();
  }
  for (auto* p : d->tc) if (p) tc_conv_plan_destroy(p, cuda_ok);
  delete d;
}

extern "C" int pe_detector_destroy(pe_detector* d) {
  if (!d || !pe_handle_release(PE_H_DETECTOR, d)) return PE_OK;
  pe_engine* e = d->e;
  e->detectors.erase(std::remove(e->detectors.begin(), e->detectors.end(), d), e->detectors.end());
  detector_free(d, pe_cuda_usable(e->device));
  return PE_OK;
}

extern "C" int pe_detector_create(pe_engine* e, const pe_det_desc* desc, const pe_gop_desc* ops, const pe_tensor_desc* tensors,
                                  const int64_t* slot_elems, const float* weights, int64_t n_weight_floats, pe_detector** out) {
  if (!e || !desc || !ops || !tensors || !slot_elems || !weights || !out) return pe_fail(PE_ERR_INVALID, "NULL argument to pe_detector_create");
  if (!pe_handle_alive(PE_H_ENGINE, e)) return pe_fail(PE_ERR_STATE, "engine handle is NULL or was destroyed");
  if (desc->max_frames <= 0 || desc->n_ops <= 0 || desc->frame_h <= 1 || desc->frame_w <= 1 || desc->net_h % 32 || desc->net_w % 32 ||
      desc->resized_h > desc->net_h || desc->resized_w > desc->net_w)
    return pe_fail(PE_ERR_INVALID, "bad detector description");
  CUD(cudaSetDevice(e->device));
  pe_range_flag() = nullptr;
  pe_detector* d = new pe_detector();
  d->e = e; d->d = *desc;
  d->ops.assign(ops, ops + desc->n_ops);
  d->tensors.assign(tensors, tensors + desc->n_tensors);
  const int maximg = desc->max_frames;
  cudaStream_t st = e->stream;
#define CUM(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { int rc = pe_fail(PE_ERR_CUDA, "%s failed: %s", #x, cudaGetErrorString(_e)); detector_free(d, true); return rc; } } while (0)
  d->slots.assign(desc->n_slots, nullptr);
  for (int s = 0; s < desc->n_slots; ++s) {
    const size_t bytes = (size_t)slot_elems[s] / 16 * PS_CHUNK_BYTES * maximg + 64 * PS_CHUNK_BYTES;
    CUM(cudaMalloc(&d->slots[s], bytes));
    CUM(cudaMemsetAsync(d->slots[s], 0, bytes, st));
  }
  CUM(cudaMalloc(&d->d_w, sizeof(float) * n_weight_floats));
  CUM(cudaMemcpyAsync(d->d_w, weights, sizeof(float) * n_weight_floats, cudaMemcpyHostToDevice, st));
  // resize tables: frame (frame_h x frame_w) -> resized (resized_h x resized_w); x weights clamped, y indices only
  d->rh = desc->resized_h; d->rw = desc->resized_w;
  {
    std::vector<int32_t> xo, yo;
    std::vector<int16_t> xa, ya;
    resize_tables(desc->frame_w, d->rw, true, xo, xa);
    resize_tables(desc->frame_h, d->rh, false, yo, ya);
    CUM(cudaMalloc(&d->d_xofs, xo.size() * 4)); CUM(cudaMalloc(&d->d_alpha, xa.size() * 2));
    CUM(cudaMalloc(&d->d_yofs, yo.size() * 4)); CUM(cudaMalloc(&d->d_beta, ya.size() * 2));
    CUM(cudaMemcpy(d->d_xofs, xo.data(), xo.size() * 4, cudaMemcpyHostToDevice)); CUM(cudaMemcpy(d->d_alpha, xa.data(), xa.size() * 2, cudaMemcpyHostToDevice));
    CUM(cudaMemcpy(d->d_yofs, yo.data(), yo.size() * 4, cudaMemcpyHostToDevice)); CUM(cudaMemcpy(d->d_beta, ya.data(), ya.size() * 2, cudaMemcpyHostToDevice));
    // scale_factor = [w_scale, h_scale, w_scale, h_scale] as float32 (mmdet Resize)
    d->scale_factor[0] = d->scale_factor[2] = (float)((double)d->rw / desc->frame_w);
    d->scale_factor[1] = d->scale_factor[3] = (float)((double)d->rh / desc->frame_h);
  }
  d->cap = desc->max_candidates > 0 ? desc->max_candidates : 4096;
  CUM(cudaMalloc(&d->d_fidx, sizeof(int32_t) * maximg));
  CUM(cudaMalloc(&d->d_cand, sizeof(float) * 6 * (size_t)d->cap * maximg));
  CUM(cudaMalloc(&d->d_count, sizeof(int) * maximg));
  CUM(cudaMallocHost(&d->h_cand, sizeof(float) * 6 * (size_t)d->cap * maximg));
  CUM(cudaMallocHost(&d->h_count, sizeof(int) * maximg));
  CUM(cudaMallocHost(&d->h_fidx, sizeof(int32_t) * maximg));
  // space-to-depth scratch for stride-2 layers the TMA gather does not cover
  size_t s2d_floats = 0;
  for (const pe_gop_desc& op : d->ops)
    if (op.kind == PE_GOP_CONV && op.stride == 2) {
      const pe_tensor_desc& to = d->tensors[op.out];
      s2d_floats = std::max(s2d_floats, (size_t)(to.H + 2) * (to.W + 2) * ps_row_floats(4 * op.cin) * maximg);
    }
  if (s2d_floats) { CUM(cudaMalloc(&d->d_s2d, s2d_floats * sizeof(float) + 64 * PS_CHUNK_BYTES)); CUM(cudaMemsetAsync(d->d_s2d, 0, s2d_floats * sizeof(float), st)); }
  CUM(cudaStreamSynchronize(st));
  // tensor-core plans: every convolution of the detector runs on conv_tc (no SIMT fallback: it would be 10x slower)
  d->tc.assign(desc->n_ops, nullptr);
  d->tc_s2d.assign(desc->n_ops, 0);
  for (int i = 0; i < desc->n_ops; ++i) {
    const pe_gop_desc& op = d->ops[i];
    if (op.kind != PE_GOP_CONV) continue;
    const pe_tensor_desc& to = d->tensors[op.out];
    const pe_tensor_desc& ti = d->tensors[op.in];
    TcConvDesc c{};
    const bool s2 = op.stride == 2;
    c.kind = s2 ? TC_KIND_2x2 : (op.ksize == 3 ? TC_KIND_3x3 : TC_KIND_1x1);
    c.Cin = s2 ? 4 * op.cin : op.cin; c.Cout = op.cout; c.act = op.act; c.dil = 0; c.H = to.H; c.W = to.W;
    c.max_rows = (long long)maximg * (to.H + 2) * (to.W + 2);
    c.out = act(d, op.out); c.out_total = to.C; c.out_coff = op.out_coff;
    // DarknetBottleneck: out = SiLU(bn(conv2(.))) + identity -- the residual joins AFTER the activation
    if (op.res >= 0) { c.res = act(d, op.res); c.res_total = d->tensors[op.res].C; c.res_coff = op.res_coff; c.res_post = 1; }
    c.wtc = d->d_w + op.wtc_off; c.bias = d->d_w + op.b_off;
    cudaError_t ce = cudaErrorNotSupported;
    if (s2) {
      c.in = nullptr; c.in_total = 4 * op.cin; c.in_coff = 0;
      c.gather_src = act(d, op.in); c.gather_total = ti.C; c.gather_coff = op.in_coff;
      ce = tc_conv_plan_create_ex(&d->tc[i], &c);
      if (ce == cudaErrorNotSupported) {
        c.gather_src = nullptr; c.in = d->d_s2d;
        ce = tc_conv_plan_create_ex(&d->tc[i], &c);
        if (ce == cudaSuccess) d->tc_s2d[i] = 1;
      }
    } else {
      c.in = act(d, op.in); c.in_total = ti.C; c.in_coff = op.in_coff;
      ce = tc_conv_plan_create_ex(&d->tc[i], &c);
    }
    if (ce != cudaSuccess) {
      int rc = pe_fail(PE_ERR_CUDA, "detector: no tensor-core plan for op %d (k=%d s=%d %d->%d @%dx%d): %s", i, op.ksize, op.stride, op.cin, op.cout, to.H, to.W,
                       cudaGetErrorString(ce));
      cudaGetLastError();
      detector_free(d, true);
      return rc;
    }
  }
  CUM(cudaStreamSynchronize(st));
#undef CUM
  d->graphs.assign(maximg + 1, nullptr);
  d->graph_seen.assign(maximg + 1, 0);
  d->graph_launches.assign(maximg + 1, 0);
  e->detectors.push_back(d);
  pe_handle_register(PE_H_DETECTOR, d);
  *out = d;
  return PE_OK;
}

// the layer program for nimg frames whose staged-frame indices are in d_fidx; candidates land in d_cand / d_count.
// part: DET_INPUT = only the input ops (resize / pad / Focus from the STAGED FRAMES), DET_NET = everything else, DET_ALL = both.
// The input ops take the engine's current staged-frames pointer as a kernel argument, and that pointer changes from block to
// block (upload slots, resident frame cache): they must never be part of a captured graph -- a replay would read the frames
// that were staged when the graph was captured.
enum { DET_INPUT = 1, DET_NET = 2, DET_ALL = 3 };
static int det_forward_eager(pe_detector* d, int nimg, int part = DET_ALL) {
  cudaStream_t st = d->e->stream;
  const pe_det_desc& dd = d->d;
  if (part & DET_NET) CUD(cudaMemsetAsync(d->d_count, 0, sizeof(int) * nimg, st));
  for (size_t i = 0; i < d->ops.size(); ++i) {
    const pe_gop_desc& op = d->ops[i];
    if (!(part & (op.kind == PE_GOP_INPUT ? DET_INPUT : DET_NET))) continue;
    switch (op.kind) {
      case PE_GOP_INPUT: {
        const pe_tensor_desc& to = d->tensors[op.out];
        launch_det_input(d->e->frames, d->d_fidx, dd.frame_h, dd.frame_w, d->rh, d->rw, to.H, to.W, d->d_xofs, d->d_alpha, d->d_yofs, d->d_beta, dd.pad_val, nimg,
                         act(d, op.out), st);
        break;
      }
      case PE_GOP_CONV: {
        if (d->tc_s2d[i]) {
          const pe_tensor_desc& ti = d->tensors[op.in];
          const pe_tensor_desc& to = d->tensors[op.out];
          launch_s2d_slice(act(d, op.in), op.cin, ti.C, op.in_coff, ti.H, ti.W, nimg, d->d_s2d, to.H, to.W, st);
          ++d->launches;
        }
        cudaError_t ce = tc_conv_launch(d->tc[i], nimg, st);
        if (ce != cudaSuccess) return pe_fail(PE_ERR_CUDA, "detector conv op %zu: %s", i, cudaGetErrorString(ce));
        break;
      }
      case PE_GOP_MAXPOOL: {
        const pe_tensor_desc& ti = d->tensors[op.in];
        launch_maxpool(act(d, op.in), op.cin, ti.C, op.in_coff, ti.H, ti.W, nimg, op.ksize, act(d, op.out), op.out_coff, st);
        break;
      }
      case PE_GOP_UPSAMPLE: {
        const pe_tensor_desc& ti = d->tensors[op.in];
        launch_upsample2(act(d, op.in), op.cin, ti.C, op.in_coff, ti.H, ti.W, nimg, act(d, op.out), d->tensors[op.out].C, op.out_coff, st);
        break;
      }
      case PE_GOP_DETHEAD: {
        const pe_tensor_desc& tc_ = d->tensors[op.in];     // classification tower output; op.res = regression tower output
        launch_det_head(act(d, op.in), act(d, op.res), op.cin, tc_.H, tc_.W, nimg, d->d_w + op.w_off, d->d_w + op.b_off, (float)op.stride, d->scale_factor,
                        dd.score_thr, op.out_coff /*first prior index of this level*/, d->d_cand, d->d_count, d->cap, nullptr, st);
        break;
      }
      default:
        return pe_fail(PE_ERR_INVALID, "detector: unknown op kind %d", op.kind);
    }
    ++d->launches;
  }
  CUD(cudaGetLastError());
  return PE_OK;
}

static int det_forward(pe_detector* d, int nimg) {
  static const bool use_graph = !(getenv("PE_GRAPH") && atoi(getenv("PE_GRAPH")) == 0);
  if (!use_graph) return det_forward_eager(d, nimg);
  cudaStream_t st = d->e->stream;
  if (d->graphs[nimg] || d->graph_seen[nimg] != 0) {        // graph path: the input ops run eagerly in front of the (replayed / captured) network
    const int rci = det_forward_eager(d, nimg, DET_INPUT);
    if (rci != PE_OK) return rci;
  }
  if (d->graphs[nimg]) {
    CUD(cudaGraphLaunch(d->graphs[nimg], st));
    d->launches += d->graph_launches[nimg];
    return PE_OK;
  }
  if (d->graph_seen[nimg]++ == 0) return det_forward_eager(d, nimg);      // first time eager: one-time attribute calls, warm-up
  if (d->graph_seen[nimg] < 0) return det_forward_eager(d, nimg, DET_NET);
  const int64_t l0 = d->launches;
  cudaGraph_t graph = nullptr;
  CUD(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  const int rc = det_forward_eager(d, nimg, DET_NET);
  const cudaError_t ce = cudaStreamEndCapture(st, &graph);
  if (rc != PE_OK || ce != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    d->graph_seen[nimg] = -1000000;
    return rc != PE_OK ? rc : det_forward_eager(d, nimg, DET_NET);
  }
  d->graph_launches[nimg] = d->launches - l0;
  const cudaError_t ci = cudaGraphInstantiate(&d->graphs[nimg], graph, 0);
  cudaGraphDestroy(graph);
  if (ci != cudaSuccess) { d->graphs[nimg] = nullptr; d->graph_seen[nimg] = -1000000; cudaGetLastError(); return det_forward_eager(d, nimg, DET_NET); }
  CUD(cudaGraphLaunch(d->graphs[nimg], st));
  return PE_OK;
}

// mmcv.ops.nms (offset 0) on candidates sorted by (score desc, prior index asc): keep a box unless its IoU with an already
// kept box exceeds iou_thr.  float32 arithmetic like the reference's CPU / CUDA op.
static int nms_host(std::vector<const float*>& c, float iou_thr, float* out, int max_det) {
  std::stable_sort(c.begin(), c.end(), [](const float* a, const float* b) {
    if (a[4] != b[4]) return a[4] > b[4];
    int ia, ib;
    std::memcpy(&ia, a + 5, 4); std::memcpy(&ib, b + 5, 4);
    return ia < ib;
  });
  std::vector<char> dead(c.size(), 0);
  int n = 0;
  for (size_t i = 0; i < c.size(); ++i) {
    if (dead[i]) continue;
    if (n < max_det) { std::memcpy(out + 5 * n, c[i], 5 * sizeof(float)); }
    ++n;
    const float* a = c[i];
    const float area_a = (a[2] - a[0]) * (a[3] - a[1]);
    for (size_t j = i + 1; j < c.size(); ++j) {
      if (dead[j]) continue;
      const float* b = c[j];
      const float iw = std::max(0.0f, std::min(a[2], b[2]) - std::max(a[0], b[0]));
      const float ih = std::max(0.0f, std::min(a[3], b[3]) - std::max(a[1], b[1]));
      const float inter = iw * ih;
      const float iou = inter / (area_a + (b[2] - b[0]) * (b[3] - b[1]) - inter);
      if (iou > iou_thr) dead[j] = 1;
    }
  }
  return n;
}

extern "C" int pe_detect(pe_detector* d, const int32_t* frame_idx, int32_t n_frames, float* out_dets, int32_t* out_counts, int32_t max_det) {
  PeRange whole("pe_detect");
  if (!d || !pe_handle_alive(PE_H_DETECTOR, d)) return pe_fail(PE_ERR_STATE, "detector handle is NULL or was destroyed (with its engine?)");
  if (!frame_idx || n_frames < 0 || !out_dets || !out_counts || max_det <= 0) return pe_fail(PE_ERR_INVALID, "bad argument to pe_detect");
  pe_engine* e = d->e;
  if (!e->frames) return pe_fail(PE_ERR_STATE, "no frames staged: call pe_stage_frames first");
  if (e->fh != d->d.frame_h || e->fw != d->d.frame_w)
    return pe_fail(PE_ERR_STATE, "staged frames are %dx%d but this detector was built for %dx%d", e->fh, e->fw, d->d.frame_h, d->d.frame_w);
  for (int i = 0; i < n_frames; ++i)
    if (frame_idx[i] < 0 || frame_idx[i] >= e->n_frames) return pe_fail(PE_ERR_STATE, "frame_idx[%d]=%d is not a staged frame (%d staged)", i, frame_idx[i], e->n_frames);
  CUD(cudaSetDevice(e->device));
  cudaStream_t st = e->stream;
  pe_range_flag() = nullptr;
  const int maxf = d->d.max_frames;
  for (int i0 = 0; i0 < n_frames; i0 += maxf) {
    const int nimg = std::min(maxf, n_frames - i0);
    CUD(cudaStreamSynchronize(st));                        // pinned staging is single-buffered
    for (int i = 0; i < nimg; ++i) d->h_fidx[i] = frame_idx[i0 + i];
    CUD(cudaMemcpyAsync(d->d_fidx, d->h_fidx, sizeof(int32_t) * nimg, cudaMemcpyHostToDevice, st));
    int rc = det_forward(d, nimg);
    if (rc) return rc;
    CUD(cudaMemcpyAsync(d->h_count, d->d_count, sizeof(int) * nimg, cudaMemcpyDeviceToHost, st));
    CUD(cudaStreamSynchronize(st));
    int maxc = 0;
    for (int i = 0; i < nimg; ++i) {
      if (d->h_count[i] > d->cap)
        return pe_fail(PE_ERR_STATE, "frame %d has %d candidates above score_thr (capacity %d): raise max_candidates", i0 + i, d->h_count[i], d->cap);
      maxc = std::max(maxc, d->h_count[i]);
    }
    if (maxc) {
      CUD(cudaMemcpy2DAsync(d->h_cand, sizeof(float) * 6 * d->cap, d->d_cand, sizeof(float) * 6 * d->cap, sizeof(float) * 6 * maxc, nimg, cudaMemcpyDeviceToHost, st));
      CUD(cudaStreamSynchronize(st));
    }
    for (int i = 0; i < nimg; ++i) {
      std::vector<const float*> c(d->h_count[i]);
      for (int k = 0; k < d->h_count[i]; ++k) c[k] = d->h_cand + ((size_t)i * d->cap + k) * 6;
      const int n = nms_host(c, d->d.nms_iou, out_dets + (size_t)(i0 + i) * max_det * 5, max_det);
      if (n > max_det) return pe_fail(PE_ERR_INVALID, "frame %d: %d detections after NMS exceed max_det=%d", i0 + i, n, max_det);
      out_counts[i0 + i] = n;
    }
  }
  return PE_OK;
}

// parity hook: copy one activation tensor (channel slice) of image `img` of the last forward as dense CHW fp32
extern "C" int pe_detector_debug_tensor(pe_detector* d, int32_t tensor_id, int32_t coff, int32_t C, int32_t img, float* out_chw) {
  if (!d || !pe_handle_alive(PE_H_DETECTOR, d)) return pe_fail(PE_ERR_STATE, "detector handle is NULL or was destroyed");
  if (!out_chw || tensor_id < 0 || tensor_id >= (int)d->tensors.size() || coff % 16 || C % 16) return pe_fail(PE_ERR_INVALID, "bad argument to pe_detector_debug_tensor");
  const pe_tensor_desc& t = d->tensors[tensor_id];
  if (coff + C > t.C || img < 0 || img >= d->d.max_frames) return pe_fail(PE_ERR_INVALID, "slice out of range");
  CUD(cudaSetDevice(d->e->device));
  cudaStream_t st = d->e->stream;
  const size_t rows = (size_t)(t.H + 2) * (t.W + 2);
  std::vector<float> host(rows * ps_row_floats(t.C));
  CUD(cudaMemcpyAsync(host.data(), act(d, tensor_id) + (size_t)img * rows * ps_row_floats(t.C), host.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
  CUD(cudaStreamSynchronize(st));
  const int rowF = ps_row_floats(t.C);
  for (int c = 0; c < C; ++c)
    for (int y = 0; y < t.H; ++y)
      for (int x = 0; x < t.W; ++x) {
        const float* row = host.data() + ((size_t)(y + 1) * (t.W + 2) + x + 1) * rowF;
        const int cc = coff + c;
#if PE_FP16
        const __half* hp = reinterpret_cast<const __half*>(reinterpret_cast<const char*>(row) + (cc >> 4) * 64);
        out_chw[((size_t)c * t.H + y) * t.W + x] = __half2float(hp[cc & 15]) + __half2float(hp[16 + (cc & 15)]) * PS_LO_INV;
#else
        const float* fp = row + (cc >> 4) * 32;
        out_chw[((size_t)c * t.H + y) * t.W + x] = fp[cc & 15] + fp[16 + (cc & 15)];
#endif
      }
  return PE_OK;
}

extern "C" int pe_detector_launch_count(pe_detector* d, int64_t* count) {
  if (!d || !count || !pe_handle_alive(PE_H_DETECTOR, d)) return pe_fail(PE_ERR_INVALID, "bad argument");
  *count = d->launches;
  return PE_OK;
}
