/*
 * poseengine.h -- C ABI of libposeengine.so, the B200 (sm_100a) engine behind the reference's
 * top-down pose path.  Plain pointers and sizes only; no C++/torch types cross this boundary.
 *
 * The reference (peabody124/PosePipeline) has no native FFI: its seam is the Python import inside
 * each DataJoint make() (pose_pipeline/pipeline.py:526,1021,1271).  Every entry point below names
 * the reference function (file:line under the reference tree) whose arithmetic it replaces; the
 * Python shims in posepipeline_b200/wrappers/ keep the reference signatures and call these through
 * ctypes (see INTEGRATION.md for the binding a maintainer adds on the reference side).
 *
 * Conventions: every function returns 0 (PE_OK) or a negative error code; the message is available
 * from pe_last_error() (thread-local).  Handles are opaque.  One engine per process per GPU; calls
 * on one engine are serialised by the caller.  All host buffers are caller-owned.  "n" counts
 * person crops (one bbox on one staged frame).
 */
#ifndef POSEENGINE_H
#define POSEENGINE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PE_ABI_VERSION 1

#define PE_OK 0
#define PE_ERR_INVALID (-1)   /* bad argument */
#define PE_ERR_CUDA (-2)      /* CUDA runtime/driver error */
#define PE_ERR_STATE (-3)     /* call sequence error (e.g. frame index not staged) */
#define PE_ERR_NOGPU (-4)     /* no CUDA device: the product path has no CPU fallback */
#define PE_ERR_RANGE (-5)     /* an activation left the range of the library's operand format (fp16x2 build: |v| > 65504);
                                 results are NOT returned clamped -- use the tf32x3 build (PE_PRECISION=tf32) for such weights */

typedef struct pe_engine pe_engine;
typedef struct pe_model pe_model;
typedef struct pe_lifter pe_lifter;
typedef struct pe_bytetrack pe_bytetrack;
typedef struct pe_detector pe_detector;

/* layer program handed over by the host graph builder (posepipeline_b200/hrnet_spec.py) */
enum { PE_OP_STEM = 0, PE_OP_CONV = 1, PE_OP_FUSE = 2, PE_OP_HEAD = 3,
       /* ViTPose (posepipeline_b200/vit_spec.py): token tensors are flat rows, described with W = 0 and H = tokens per image */
       PE_OP_PATCH = 4,   /* crop -> patch rows [tokens][3*ksize*ksize]; ksize = patch size, stride = padding */
       PE_OP_GEMM = 5,    /* Linear: in[0] -> out, bias b_off, weights wtc_off; relu = activation (0 none, 1 ReLU, 3 GELU);
                             residual = tensor id, or reserved = 1: a [tokens][cout] fp32 table at w_off added to every image */
       PE_OP_LN = 6,      /* LayerNorm over cout channels: gamma w_off, beta b_off, eps = 10^-up[1]; up[0] = 1 writes the padded
                             2-D grid of the output tensor (tokens -> H x W) instead of flat rows */
       PE_OP_ATTN = 7,    /* multi-head self-attention: in[0] = qkv rows [3][cin heads][64], out rows [heads][64] */
       PE_OP_D2S = 8 };   /* depth-to-space x2: in[0] [H x W][4*cout] -> out [2H x 2W][cout] (deconvolution head) */
enum { PE_POST_NONE = 0, PE_POST_DEFAULT = 1, PE_POST_UNBIASED = 2, PE_POST_UDP = 3 /* mmpose post_dark_udp + UDP back-projection */ };
#define PE_MODEL_FLAG_UDP 1   /* pe_model_desc.reserved bit 0: TopDownAffine(use_udp=True) warp matrix */

typedef struct pe_op_desc {
  int32_t kind;        /* PE_OP_* */
  int32_t out;         /* output tensor id */
  int32_t in[4];       /* input tensor ids (-1 = unused; STEM reads the uint8 crop) */
  int32_t up[4];       /* FUSE: nearest-upsample factor of each input */
  int32_t n_in;
  int32_t ksize, stride, cin, cout;
  int32_t relu;
  int32_t residual;    /* tensor id added before the ReLU, or -1 */
  int32_t reserved;
  int64_t w_off;       /* float offset of the packed weights [k*k][cin][cout] in the weight blob */
  int64_t b_off;       /* float offset of the folded bias [cout] */
  int64_t wtc_off;     /* float offset of the tensor-core packing, or -1 */
} pe_op_desc;

typedef struct pe_tensor_desc {
  int32_t C, H, W;     /* per-image dims (activations live in HBM as padded, tf32 hi/lo split NHWC) */
  int32_t slot;        /* buffer slot (tensors with disjoint live ranges share one) */
} pe_tensor_desc;

typedef struct pe_model_desc {
  int32_t in_h, in_w;          /* network input (crop) size: 384x288, cfg data_cfg.image_size */
  int32_t hm_h, hm_w;          /* heatmap size: 96x72 */
  int32_t num_joints;
  int32_t n_ops, n_tensors, n_slots;
  int32_t max_crops;           /* crops per internal batch (each crop = 2 images when flip_test) */
  int32_t flip_test;           /* cfg test_cfg.flip_test */
  int32_t shift_heatmap;       /* cfg test_cfg.shift_heatmap */
  int32_t post_process;        /* PE_POST_* : cfg test_cfg.post_process */
  int32_t blur_kernel;         /* cfg test_cfg.modulate_kernel (17) */
  int32_t swap_rb;             /* 0 = reproduce the wrapper's double BGR<->RGB swap */
  int32_t use_tensor_cores;    /* 1 = tcgen05 path for eligible convs, 0 = fp32 SIMT everywhere */
  int32_t reserved;            /* PE_MODEL_FLAG_* */
  float padding;               /* bbox padding 1.25 */
  float pixel_std;             /* 200 */
} pe_model_desc;

/* ---- library ---- */
int pe_abi_version(void);
/* activation / tensor-core operand format this library was built for: 0 = tf32x3 (8 B per element, hi/lo TF32 pairs),
 * 1 = fp16x2 (4 B per element, h/l FP16 pairs).  Decides how the host packs the tensor-core weights. */
int pe_precision_mode(void);
const char* pe_last_error(void);
int pe_device_count(int* count);

/* ---- engine: one per GPU.  `cuda_stream` may be NULL (engine creates its own) or a cudaStream_t
 * the caller times with its own events (bench.py passes torch's current stream). ---- */
int pe_engine_create(int device, void* cuda_stream, pe_engine** out);
/* Destroys the engine AND every model / lifter still alive on it.  All pe_*_destroy calls are idempotent: a handle that
 * is unknown or already destroyed (e.g. a model released after its engine, as a garbage collector may do) returns PE_OK
 * without being touched, and during process teardown (CUDA runtime unloading) only host memory is released.  This is the
 * contract that lets a populate() worker (pose_pipeline/utils/standard_pipelines.py:100) return normally. */
int pe_engine_destroy(pe_engine* e);
int pe_engine_sync(pe_engine* e);
/* Destroys every live engine of the process (the Python host registers it with atexit). */
int pe_shutdown(void);

/* Frame staging: replaces the per-frame host->device copy buried in
 * inference_top_down_pose_model (pose_pipeline/wrappers/mmpose.py:75) after cap.read() (:63).
 * `frames` = n contiguous-or-strided HWC uint8 BGR images exactly as cv2.VideoCapture.read()
 * returns them (pinned host memory recommended).  Copies asynchronously on the engine stream into
 * the device frame store; slot i holds frame i until the next call. */
int pe_stage_frames(pe_engine* e, const uint8_t* frames, int32_t n, int32_t height, int32_t width,
                    int64_t frame_stride_bytes);
/* As above but the frames are already in device memory (benchmark "resident" leg). */
int pe_stage_frames_device(pe_engine* e, const uint8_t* d_frames, int32_t n, int32_t height, int32_t width);

/* Frame source (replaces the synchronous cap.read() -> H2D sequence of wrappers/mmpose.py:60-76 and
 * wrappers/mmtrack.py:37-45): two device slots.  pe_frames_upload copies a block on the engine's own copy stream and may be
 * called from a decode thread while the engine stream computes on the other slot; pe_frames_select makes a slot the staged
 * frames (the engine stream waits for that slot's upload only).  The caller must not upload into a slot while a compute call
 * that selected it is still running (the compute entry points are synchronous, so "it has returned" suffices). */
int pe_frames_upload(pe_engine* e, int32_t slot, const uint8_t* frames, int32_t n, int32_t height, int32_t width,
                     int64_t frame_stride_bytes);
/* same, but the block lands in caller-owned device memory (e.g. a frame cache resident in HBM) instead of the engine's buffer */
int pe_frames_upload_to(pe_engine* e, int32_t slot, void* d_dst, const uint8_t* frames, int32_t n, int32_t height, int32_t width,
                        int64_t frame_stride_bytes);
int pe_frames_select(pe_engine* e, int32_t slot);
int pe_frames_slot_ptr(pe_engine* e, int32_t slot, void** out);

/* cv2.warpAffine(frame, trans, (out_w, out_h), INTER_LINEAR, border 0) for n (staged frame, forward 2x3 matrix) pairs ->
 * uint8 HWC crops n*out_h*out_w*3 on the host: the crop primitive of pose_pipeline/utils/bounding_box.py:32-53
 * (crop_image_bbox / get_person_dataloader), bit-exact against cv2.  swap_rb = 1 swaps channels 0 and 2. */
int pe_warp_affine(pe_engine* e, const int32_t* frame_idx, const double* trans, int32_t n, int32_t out_h, int32_t out_w,
                   int32_t swap_rb, uint8_t* out_crops);

/* PersonBbox.make (pose_pipeline/pipeline.py:656-687): per frame keep the dicts whose track_id is in
 * keep_tracks; exactly one -> present, bbox = its tlhw; then NaN-mask, bfill(limit 2), ffill(limit 2).
 * Host-only, bit-exact.  counts[f] = #tracks in frame f; track_ids/tlhw are concatenated over frames. */
int pe_person_bbox(const int32_t* counts, int32_t n_frames, const int64_t* track_ids, const double* tlhw,
                   const int64_t* keep_tracks, int32_t n_keep, double* bbox_out /*n_frames*4*/,
                   uint8_t* present_out /*n_frames*/);

/* ---- top-down model (mmpose init_pose_model, wrappers/mmpose.py:57) ---- */
int pe_model_create(pe_engine* e, const pe_model_desc* desc, const pe_op_desc* ops, const pe_tensor_desc* tensors,
                    const int64_t* slot_elems /*n_slots: padded elems per image*/, const float* weights,
                    int64_t n_weight_floats, const float* norm_lut /*3*256: (v/255-mean_c)/std_c*/,
                    const int32_t* flip_perm /*num_joints: channel read for joint k in the flipped map*/,
                    pe_model** out);
int pe_model_destroy(pe_model* m);

/* The hot path == the body of the reference loop wrappers/mmpose.py:60-76 for n (frame, bbox) pairs:
 * bbox->center/scale, affine crop (cv2.warpAffine-exact), normalise, HRNet x2 (flip test), flip-merge,
 * DARK decode, back-projection.  frame_idx[i] indexes the staged frames.  out_kpts: n*K*3 floats
 * [x_px, y_px, score].  Synchronous on return. */
int pe_topdown(pe_model* m, const int32_t* frame_idx, const double* bbox_xywh, int32_t n, float* out_kpts);
/* Same, but results stay on the device until pe_engine_sync(); no host wait (throughput runs). */
int pe_topdown_async(pe_model* m, const int32_t* frame_idx, const double* bbox_xywh, int32_t n, float* out_kpts_pinned);

/* ---- parity hooks (each stage of the path on its own) ---- */
/* mmpose bbox_xywh2cs + get_affine_transform (A.1 steps 2,4): center(2) scale(2) f32, trans 2x3 f64 */
int pe_box_to_affine(const pe_model_desc* desc, const double* bbox_xywh, float* center, float* scale, double* trans);
/* cv2.warpAffine(INTER_LINEAR, BORDER_CONSTANT 0) of staged frames -> uint8 crops n*in_h*in_w*3 */
int pe_warp_crops(pe_model* m, const int32_t* frame_idx, const double* bbox_xywh, int32_t n, uint8_t* out_crops,
                  float* out_center /*n*2*/, float* out_scale /*n*2*/);
/* network only: uint8 crops -> heatmaps (plain pass and raw flipped pass), each n*K*hm_h*hm_w */
int pe_forward_heatmaps(pe_model* m, const uint8_t* crops, int32_t n, float* hm_plain, float* hm_flipped);
/* keypoints_from_heatmaps + flip merge (A.1 step 6-7, A.5) on caller-provided heatmaps */
int pe_decode_heatmaps(pe_model* m, const float* hm_plain, const float* hm_flipped /*NULL = no flip merge*/,
                       const float* center, const float* scale, int32_t n, float* out_kpts);
/* copy one intermediate activation of the last forward (image `img` of the internal batch) as dense CHW fp32 */
int pe_debug_tensor(pe_model* m, int32_t tensor_id, int32_t img, float* out_chw);
/* number of kernels this model has launched so far (bench.py "gpu_launches") */
int pe_model_launch_count(pe_model* m, int64_t* count);
/* device timing of the dominant (conv) kernels between two marks, for bench.py's roofline */
int pe_model_profile(pe_model* m, int32_t enable);
int pe_model_profile_read(pe_model* m, double* conv_ms, double* other_ms, int64_t* conv_launches);
/* per-op accumulated device milliseconds since pe_model_profile(m, 2) (enable=2 times every op, 1 only the convolutions) */
int pe_model_profile_ops(pe_model* m, double* ms_per_op, int32_t n_ops);

/* one convolution layer on its own (k in {1,3}, stride 1, or 3x3 stride 2; BN already folded): dense NCHW in/out; the
 * library packs to its HBM layout, runs the SIMT (use_tc=0) or tcgen05 (use_tc=1) kernel, verifies the zero halo and
 * unpacks.  w_simt: [k*k][Cin][Cout]; w_tc: [k*k][Cin/16][Cout][hi16|lo16] (stride 2: the 2x2 space-to-depth form,
 * [4][4*Cin/16][Cout][32]).  H, W are the input dims; the output is H/stride x W/stride. */
int pe_conv_test(pe_engine* e, const float* in_nchw, int32_t nimg, int32_t Cin, int32_t H, int32_t W, const float* w_simt,
                 const float* w_tc, const float* bias, const float* res_nchw, int32_t Cout, int32_t ks, int32_t stride,
                 int32_t relu, int32_t use_tc, float* out_nchw);

/* host-only: the candidate tilings the tensor-core convolution planner would consider for a layer (no GPU needed; the
 * CPU suite checks their shared-memory / TMEM / alignment invariants).  ks: 1, 3, or 2 (= the 2x2 space-to-depth form of a
 * stride-2 3x3 layer, Cin already x4); gather: 1 = TMA gather mode of that form.  Each candidate fills PE_TC_CAND_FIELDS
 * int32 values: n_split, MT, NC, KC, stages, staging_buffers, stage_bytes, smem_bytes, tmem_cols, rows_per_group, n_drain,
 * window_rows, cta_group (1, or 2 = CTA-pair form: M = 256 MMAs over the two SMs of a TPC),
 * epilogue_sets (1 = 12 epilogue warps on every tile, 2 = two sets of 8 warps on alternate tiles, 3 / 4 = 8 drain warps + 4 / 8
 * finalize warps).  Returns the number of candidates (<= cap) or a negative error. */
#define PE_TC_CAND_FIELDS 14
int pe_tc_plan_candidates(int32_t Cin, int32_t Cout, int32_t ks, int32_t has_residual, int32_t H, int32_t W, int32_t max_img,
                          int32_t gather, int32_t* out, int32_t cap);

/* Host-only view of conv_tc's persistent schedule (nothing in the reference): work item w of a layer with n_split N slices and
 * tiles_m M tiles (tile_rows rows each) -> (*tile, *n_slice).  mode / budget_kb as PE_TC_GROUP / PE_TC_GROUP_KB (1, 49152 are the
 * defaults).  Returns the group size in M tiles (0 = n-major order: every slice sweeps all tiles) or a negative error.  CPU test:
 * every (tile, slice) pair is visited exactly once, whatever the group size. */
int pe_tc_work_item(int32_t Cin, int32_t Cout, int32_t n_split, int32_t tile_rows, int32_t tiles_m, int32_t mode, int32_t budget_kb,
                    int32_t w, int32_t* tile, int32_t* n_slice);

/* ---- VideoPose3D lifter (wrappers/videopose3d.py:46-85; TemporalModelOptimized1f 243 frames) ---- */
/* offsets: 10 layers (expand_conv, layers_conv[0..7], shrink) x 3 float offsets into `weights`: SIMT packing
 * [tap][Cin][Cout], folded bias [Cout], tensor-core packing (engine.pack_tc_weights) or -1. */
int pe_lifter_create(pe_engine* e, const float* weights, int64_t n_floats, const int64_t* offsets, int32_t n_offsets,
                     int32_t channels, pe_lifter** out);
int pe_lifter_destroy(pe_lifter* l);
/* 1 when the temporal convolutions run on the tcgen05 kernel (after the first pe_lift3d), 0 = fp32 SIMT GEMMs */
int pe_lifter_uses_tensor_cores(pe_lifter* l);
int pe_lifter_launch_count(pe_lifter* l, int64_t* count);
/* kp2d_norm: N*17*2 normalised screen coords; out: N*17*3.  Windows are edge-replicated (pad 121). */
int pe_lift3d(pe_lifter* l, const float* kp2d_norm, int32_t n_frames, float* out3d);

/* ---- person detector (YOLOX-X, the detector half of mmtrack.apis.inference_mot, pose_pipeline/wrappers/mmtrack.py:45;
 * architecture 3rdparty/mmtracking/_base_/models/yolox_x_8x8.py:5-26, test pipeline and thresholds
 * mot/bytetrack/bytetrack_yolox_x_crowdhuman_mot17-private-half.py:6,9-20,60-81).  The layer program comes from the host
 * graph builder (posepipeline_b200/yolox_spec.py); operands may be 16-channel-aligned slices of wider tensors. ---- */
enum { PE_GOP_INPUT = 0, PE_GOP_CONV = 1, PE_GOP_MAXPOOL = 2, PE_GOP_UPSAMPLE = 3, PE_GOP_DETHEAD = 4 };
typedef struct pe_gop_desc {
  int32_t kind;                     /* PE_GOP_* */
  int32_t in, out, res;             /* tensor ids (-1 = none); DETHEAD: in = cls tower output, res = reg tower output */
  int32_t in_coff, out_coff, res_coff; /* first channel of each view; DETHEAD: out_coff = index of the level's first prior */
  int32_t cin, cout;                /* channels of the views */
  int32_t ksize, stride;            /* CONV: 1|3, 1|2; MAXPOOL: window; DETHEAD: stride = the level's stride (8/16/32) */
  int32_t act;                      /* 0 none, 1 ReLU, 2 SiLU */
  int64_t w_off, b_off, wtc_off;    /* float offsets into the weight blob (DETHEAD: w = [6][cin] cls,reg x4,obj; b = [6]) */
} pe_gop_desc;
typedef struct pe_det_desc {
  int32_t frame_h, frame_w;         /* staged frame size this detector is built for */
  int32_t resized_h, resized_w;     /* mmcv.imrescale(keep ratio, (800,1440)) size */
  int32_t net_h, net_w;             /* padded to a multiple of 32 (Pad size_divisor) */
  int32_t n_ops, n_tensors, n_slots, max_frames, max_candidates, reserved;
  float score_thr, nms_iou, pad_val, reserved_f;
} pe_det_desc;
int pe_detector_create(pe_engine* e, const pe_det_desc* desc, const pe_gop_desc* ops, const pe_tensor_desc* tensors,
                       const int64_t* slot_elems, const float* weights, int64_t n_weight_floats, pe_detector** out);
int pe_detector_destroy(pe_detector* d);
/* frames = indices of staged frames (pe_stage_frames).  out_dets: n_frames * max_det rows [x1,y1,x2,y2,score] float32 in
 * original-image pixels (rescale=True), score-descending after NMS; out_counts[i] = rows of frame i. */
int pe_detect(pe_detector* d, const int32_t* frame_idx, int32_t n_frames, float* out_dets, int32_t* out_counts, int32_t max_det);
int pe_detector_debug_tensor(pe_detector* d, int32_t tensor_id, int32_t coff, int32_t C, int32_t img, float* out_chw);
int pe_detector_launch_count(pe_detector* d, int64_t* count);

/* ---- ByteTrack association (host only; the per-frame `ByteTracker.track` mmtrack runs inside inference_mot,
 * pose_pipeline/wrappers/mmtrack.py:45; configuration 3rdparty/mmtracking/mot/bytetrack/
 * bytetrack_yolox_x_crowdhuman_mot17-private-half.py:21-28).  cfg = NULL (the reference's values) or 9 floats
 * {obj_score high, low, init_track_thr, match_iou high, low, tentative, weight_iou_with_det_scores, num_tentatives,
 * num_frames_retain}. ---- */
int pe_bytetrack_create(const float* cfg, int32_t n_cfg, pe_bytetrack** out);
int pe_bytetrack_destroy(pe_bytetrack* t);
int pe_bytetrack_reset(pe_bytetrack* t);
/* One frame: dets = n rows [x1,y1,x2,y2,score] in the detector's output order (score-descending NMS output); frame_id 0
 * resets the tracker like ByteTrack.simple_test.  out_rows: up to cap rows [track_id,x1,y1,x2,y2,score] (float64, the dtype
 * of the reference's result["track_bboxes"][0]) in the reference's row order; *n_out = rows written. */
int pe_bytetrack_update(pe_bytetrack* t, int32_t frame_id, const float* dets, int32_t n, double* out_rows, int32_t cap,
                        int32_t* n_out);

#ifdef __cplusplus
}
#endif
#endif /* POSEENGINE_H */
