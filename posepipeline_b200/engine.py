"""Python host over the C ABI (include/poseengine.h): engine / top-down model / lifter objects.

This is the layer the reference-compatible wrappers (posepipeline_b200/wrappers/*.py) call.  It owns no
arithmetic: it builds the layer program (hrnet_spec), folds BatchNorm, packs weights and calls
libposeengine.so.  There is no CPU fallback; constructing an engine without a B200 raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import ModelDesc, OpDesc, TensorDesc, check, ptr
from .hrnet_spec import OP_CONV, OP_FUSE, OP_HEAD, OP_STEM, Program, build_program
from .vit_spec import OP_ATTN, OP_D2S, OP_GEMM, OP_LN, OP_PATCH, build_vitpose_program
from .weights import bn_of, fold_bn

COCO_FLIP_PAIRS = [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [11, 12], [13, 14], [15, 16]]


@dataclass
class TopDownSpec:
    """What the mmpose config of a method contributes to the arithmetic (reference
    3rdparty/mmpose/config/top_down/darkpose/coco/hrnet_w48_coco_384x288_dark.py:40-102,129-144)."""
    variant: str = "w48"
    image_size: Tuple[int, int] = (288, 384)        # (W, H)
    heatmap_size: Tuple[int, int] = (72, 96)        # (W, H)
    num_joints: int = 17
    flip_test: bool = True
    post_process: Optional[str] = "unbiased"
    shift_heatmap: bool = True
    modulate_kernel: int = 17
    padding: float = 1.25
    flip_pairs: List[List[int]] = field(default_factory=lambda: [list(p) for p in COCO_FLIP_PAIRS])
    mean: Tuple[float, float, float] = (0.485, 0.456, 0.406)
    std: Tuple[float, float, float] = (0.229, 0.224, 0.225)
    wrapper_double_swap: bool = True
    use_udp: bool = False                          # TopDownAffine(use_udp=True) + post_dark_udp decode (ViTPose configs)
    checkpoint: Optional[str] = None               # path under MODEL_DATA_DIR (reference wrappers/mmpose.py:33-52)
    config: Optional[str] = None                   # mmcv config path under MODEL_DATA_DIR (same lines)


# Halpe-136 left/right pairs = the `swap` fields of the reference's dataset_info (3rdparty/mmpose/config/_base_/halpe.py, passed
# to the model at halpe/hrnet_w48_halpe_384x288_dark_plus.py:152).  Built-in copy for runs without the config tree; with
# $MODEL_DATA_DIR/mmpose/config present the pairs are read from the file (mmcv_config.flip_pairs_from_dataset_info), and
# tests/test_config.py checks the two agree.
HALPE_FLIP_PAIRS = [[a, b] for a, b in zip(
    [1, 3, 5, 7, 9, 11, 13, 15, 20, 22, 24, 26, 27, 28, 29, 30, 31, 32, 33, 43, 44, 45, 46, 47, 57, 58, 62, 63, 64, 65, 66, 67, 74, 75, 76,
     81, 82, 86, 87, 91] + list(range(94, 115)),
    [2, 4, 6, 8, 10, 12, 14, 16, 21, 23, 25, 42, 41, 40, 39, 38, 37, 36, 35, 52, 51, 50, 49, 48, 61, 60, 71, 70, 69, 68, 73, 72, 80, 79, 78,
     85, 84, 90, 89, 93] + list(range(115, 136)))]


METHODS: Dict[str, TopDownSpec] = {
    # reference wrappers/mmpose.py:33-36
    "HRNet_W48_COCO": TopDownSpec(config="mmpose/config/top_down/darkpose/coco/hrnet_w48_coco_384x288_dark.py",
                                  checkpoint="mmpose/checkpoints/hrnet_w48_coco_384x288_dark-e881a4b6_20210203.pth"),
    # reference wrappers/mmpose.py:41-44; old-style config without dataset_info => mmpose falls back to the COCO-17 body
    # pairs when flipping all 133 channels -- reproduced
    "HRNet_W48_COCOWholeBody": TopDownSpec(num_joints=133, config="mmpose/config/coco-wholebody/hrnet_w48_coco_wholebody_384x288_dark_plus.py",
                                           checkpoint="mmpose/checkpoints/hrnet_w48_coco_wholebody_384x288_dark-f5726563_20200918.pth"),
    # reference wrappers/mmpose.py:49-52 (PosePipe's production default, scripts/process_h36m.py:15)
    "HRNet_W48_HALPE": TopDownSpec(num_joints=136, flip_pairs=[list(p) for p in HALPE_FLIP_PAIRS],
                                   config="mmpose/config/halpe/hrnet_w48_halpe_384x288_dark_plus.py",
                                   checkpoint="mmpose/checkpoints/hrnet_w48_halpe_384x288_dark_plus-d13c2588_20211021.pth"),
    # BASELINE config 1 (upstream hrnet_w32_coco_256x192.py; not configured in the reference)
    # BASELINE configs[2] (upstream ViTPose_base_coco_256x192.py; not configured in the reference)
    "ViTPose_B_COCO": TopDownSpec(variant="vitpose_b", image_size=(192, 256), heatmap_size=(48, 64), post_process="udp",
                                  shift_heatmap=False, modulate_kernel=11, use_udp=True,
                                  config="mmpose/config/top_down/vitpose/coco/ViTPose_base_coco_256x192.py",
                                  checkpoint="mmpose/checkpoints/vitpose-b.pth"),
    "HRNet_W32_COCO": TopDownSpec(variant="w32", image_size=(192, 256), heatmap_size=(48, 64), post_process="default",
                                  modulate_kernel=11, config="mmpose/config/top_down/hrnet/coco/hrnet_w32_coco_256x192.py",
                                  checkpoint="mmpose/checkpoints/hrnet_w32_coco_256x192-c78dce93_20200708.pth"),
}


def spec_for(method: str, model_data_dir: str = "") -> TopDownSpec:
    """The method's settings as the reference would see them: read from its mmcv config file under
    ``model_data_dir`` (the path ``wrappers/mmpose.py:33-52`` passes to ``init_pose_model``) when that file exists -- so an
    edited ``test_cfg`` / ``data_cfg`` is honoured -- else the built-in copy of the shipped values."""
    import dataclasses
    base = METHODS[method]
    path = os.path.join(model_data_dir or "", base.config or "")
    if base.config and os.path.isfile(path):
        from .mmcv_config import load_config, topdown_settings
        return dataclasses.replace(base, **topdown_settings(load_config(path)))
    return base


def program_for(spec: "TopDownSpec") -> Program:
    if spec.variant == "vitpose_b":
        return build_vitpose_program(spec.image_size[1], spec.image_size[0], spec.num_joints)
    return build_program(spec.variant, spec.image_size[1], spec.image_size[0], spec.num_joints)


def deconv_as_conv_weights(w: np.ndarray) -> np.ndarray:
    """ConvTranspose2d(k4, s2, p1) weights (Cin, Cout, 4, 4) -> the equivalent 3x3 stride-1 convolution (4*Cout, Cin, 3, 3) whose
    output channel block (py*2+px) holds output pixels (2a+py, 2b+px): out[2a+py] takes taps ky with (2a+py+1-ky) even,
    i.e. py=0: ky=1 (input row a), ky=3 (row a-1); py=1: ky=0 (row a+1), ky=2 (row a).  csrc/vit.cu d2s_kernel interleaves."""
    cin, cout = w.shape[:2]
    out = np.zeros((4, cout, cin, 3, 3), w.dtype)
    taps = {0: ((1, 1), (3, 0)), 1: ((0, 2), (2, 1))}          # parity -> ((k, 3x3 tap index) ...): tap index = input offset + 1
    for py in range(2):
        for px in range(2):
            for ky, ty in taps[py]:
                for kx, tx in taps[px]:
                    out[py * 2 + px, :, :, ty, tx] = w[:, :, ky, kx].T
    return out.reshape(4 * cout, cin, 3, 3)


def tf32_split(x: np.ndarray):
    """hi = x rounded to TF32 (nearest, ties away: cvt.rna.tf32.f32), lo = x - hi (exact)."""
    x = np.ascontiguousarray(x, np.float32)
    bits = x.view(np.uint32)
    hi = ((bits + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)
    return hi, (x - hi).astype(np.float32)


class WeightBlob:
    def __init__(self):
        self.parts: List[np.ndarray] = []
        self.n = 0

    def add(self, a: np.ndarray) -> int:
        a = np.ascontiguousarray(a, np.float32).ravel()
        off = self.n
        pad = (-a.size) % 64
        self.parts.append(a)
        if pad:
            self.parts.append(np.zeros(pad, np.float32))
        self.n += a.size + pad
        return off

    def array(self) -> np.ndarray:
        return np.concatenate(self.parts) if self.parts else np.zeros(0, np.float32)


def space_to_depth_weights(w: np.ndarray) -> np.ndarray:
    """3x3 stride-2 weights (Cout,Cin,3,3) -> the equivalent 2x2 stride-1 weights (Cout,4*Cin,2,2) over the
    space-to-depth repack of the input (csrc/kernels_simt.cu s2d_kernel): W'[o][(py,px,c)][dy][dx] = w[o][c][2dy+py][2dx+px]."""
    cout, cin = w.shape[:2]
    out = np.zeros((cout, 4, cin, 2, 2), w.dtype)
    for py in range(2):
        for px in range(2):
            for dy in range(2):
                for dx in range(2):
                    ky, kx = 2 * dy + py, 2 * dx + px
                    if ky < 3 and kx < 3:
                        out[:, py * 2 + px, :, dy, dx] = w[:, :, ky, kx]
    return out.reshape(cout, 4 * cin, 2, 2)


FP16_LO_SCALE = np.float32(2048.0)          # csrc/pe_common.cuh PS_LO_SCALE


def fp16_split(x: np.ndarray):
    """h = fp16(x), l = fp16((x - h) * 2^11): 22 significant bits; the scaled low half is a normal fp16 number whenever x
    is (csrc/pe_common.cuh).  Values beyond the fp16 range are an error, not clamped."""
    x = np.ascontiguousarray(x, np.float32)
    if x.size and float(np.abs(x).max()) > 65504.0:
        raise OverflowError("value exceeds the fp16x2 operand range (|x| <= 65504); use the tf32 build (PE_PRECISION=tf32)")
    h = x.astype(np.float16)
    l = ((x - h.astype(np.float32)) * FP16_LO_SCALE).astype(np.float16)
    return h, l


def pack_tc_weights(w: np.ndarray) -> np.ndarray:
    """(Cout,Cin,k,k) folded fp32 -> tensor-core weight blob (float32 words), see csrc/conv_tc.cu:
         [scale 2^-k: Cout floats, padded to 64] [inverse 2^k: same] [operand: [tap][Cin/16][Cout][hi16|lo16]]
    The operand holds TF32 pairs (library built with PE_FP16=0) or FP16 pairs packed two per word (PE_FP16=1).
    Each output channel's weights are multiplied by a power of two so the largest is in [8,16): exact, and it keeps the
    small `lo` halves out of the FP16 subnormal range (a 0.05 weight would otherwise carry only ~20 significant bits);
    the epilogue multiplies the accumulators by `scale` = the inverse power of two (exact again)."""
    if w.ndim == 3:                                # 1-D temporal convolution (Cout, Cin, taps)
        w = w[:, :, :, None]
    cout, cin, k, kw = w.shape
    mx = np.abs(w.reshape(cout, -1)).max(axis=1)
    e = np.where(mx > 0, np.floor(np.log2(16.0 / np.maximum(mx, 1e-30))), 0.0)
    e = np.clip(e, -20, 40)
    while True:                                   # guard the open upper bound against log2 rounding
        over = mx * np.exp2(e) >= 16.0
        if not over.any():
            break
        e = e - over
    up = np.exp2(e).astype(np.float32)
    ws = (w * up[:, None, None, None]).astype(np.float32)                                    # exact
    scale = np.ones(2 * (((cout + 63) // 64) * 64), np.float32)
    scale[:cout] = np.exp2(-e).astype(np.float32)
    scale[len(scale) // 2: len(scale) // 2 + cout] = up
    t = ws.transpose(2, 3, 1, 0).reshape(k * kw, cin // 16, 16, cout).transpose(0, 1, 3, 2)   # tap, chunk, cout, 16
    if _lib.load().pe_precision_mode() == 1:
        h, l = fp16_split(t)
        op = np.ascontiguousarray(np.concatenate([h, l], axis=3)).view(np.float32)
    else:
        hi, lo = tf32_split(t)
        op = np.concatenate([hi, lo], axis=3)
    return np.concatenate([scale, op.ravel()])


class PoseEngine:
    """One per process per GPU.  ``stream``: optional cudaStream_t handle (int) to run on."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self.lib = _lib.load()
        h = C.c_void_p()
        check(self.lib.pe_engine_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h)))
        self.h = h
        self.device = device
        self._frames_keepalive = None
        _lib.track(self)

    def stage_frames(self, frames: np.ndarray):
        """frames (n,H,W,3) uint8 BGR as cv2 returns them.  Asynchronous H2D on the engine stream."""
        if frames.dtype != np.uint8 or frames.ndim != 4 or frames.shape[3] != 3:
            raise ValueError("frames must be (n,H,W,3) uint8")
        if not frames.flags.c_contiguous:
            frames = np.ascontiguousarray(frames)
        self._frames_keepalive = frames
        n, H, W, _ = frames.shape
        check(self.lib.pe_stage_frames(self.h, ptr(frames), n, H, W, 0))
        self.n_frames = n

    def upload_block(self, slot: int, frames: np.ndarray):
        """Asynchronous H2D of a block into device slot 0/1 on the copy stream (callable from a decode thread)."""
        n, H, W, _ = frames.shape
        check(self.lib.pe_frames_upload(self.h, int(slot), ptr(frames), n, H, W, 0))

    def upload_block_to(self, slot: int, dev_ptr: int, frames: np.ndarray):
        """As upload_block, but into caller-owned device memory (the resident frame cache)."""
        n, H, W, _ = frames.shape
        check(self.lib.pe_frames_upload_to(self.h, int(slot), C.c_void_p(dev_ptr), ptr(frames), n, H, W, 0))

    def select_block(self, slot: int, n: int):
        check(self.lib.pe_frames_select(self.h, int(slot)))
        self.n_frames = n

    def slot_ptr(self, slot: int) -> int:
        p = C.c_void_p()
        check(self.lib.pe_frames_slot_ptr(self.h, int(slot), C.byref(p)))
        return p.value or 0

    def warp_affine(self, frame_idx, trans: np.ndarray, out_size, swap_rb: bool = False) -> np.ndarray:
        """cv2.warpAffine(frame, trans_i, out_size=(w, h), INTER_LINEAR) of staged frames -> (n,h,w,3) uint8, bit-exact."""
        fi = np.ascontiguousarray(frame_idx, np.int32)
        t = np.ascontiguousarray(trans, np.float64).reshape(-1, 6)
        w, h = int(out_size[0]), int(out_size[1])
        out = np.empty((len(fi), h, w, 3), np.uint8)
        if len(fi):
            check(self.lib.pe_warp_affine(self.h, ptr(fi), ptr(t), len(fi), h, w, int(swap_rb), ptr(out)))
        return out

    def stage_frames_device(self, dev_ptr: int, n: int, H: int, W: int):
        check(self.lib.pe_stage_frames_device(self.h, C.c_void_p(dev_ptr), n, H, W))
        self.n_frames = n

    def sync(self):
        check(self.lib.pe_engine_sync(self.h))

    def close(self):
        """Destroys the engine and every model / lifter created on it (their Python objects become inert)."""
        if getattr(self, "h", None):
            self.lib.pe_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        if _lib.finalizing():          # interpreter exit: _lib.shutdown() (atexit) has already released everything
            return
        try:
            self.close()
        except Exception:
            pass


def conv_test(engine: PoseEngine, x: np.ndarray, w: np.ndarray, bias: np.ndarray, res: Optional[np.ndarray] = None,
              relu: bool = True, use_tc: bool = True, stride: int = 1) -> np.ndarray:
    """One conv layer through the library (parity hook): x (n,Cin,H,W), w (Cout,Cin,k,k) folded -> (n,Cout,H/s,W/s)."""
    x = np.ascontiguousarray(x, np.float32)
    n, cin, H, W = x.shape
    cout, _, k, _ = w.shape
    ws = np.ascontiguousarray(w.transpose(2, 3, 1, 0).reshape(k * k, cin, cout), np.float32)
    wt = np.ascontiguousarray(pack_tc_weights(space_to_depth_weights(w) if stride == 2 else w), np.float32)
    b = np.ascontiguousarray(bias, np.float32)
    r = np.ascontiguousarray(res, np.float32) if res is not None else None
    out = np.empty((n, cout, H // stride, W // stride), np.float32)
    check(engine.lib.pe_conv_test(engine.h, ptr(x), n, cin, H, W, ptr(ws), ptr(wt), ptr(b), ptr(r) if r is not None else None,
                                  cout, k, stride, int(relu), int(use_tc), ptr(out)))
    return out


def model_desc(spec: TopDownSpec, n_ops=0, n_tensors=0, n_slots=0, max_crops=1, use_tc=False) -> ModelDesc:
    return ModelDesc(in_h=spec.image_size[1], in_w=spec.image_size[0], hm_h=spec.heatmap_size[1], hm_w=spec.heatmap_size[0],
                     num_joints=spec.num_joints, n_ops=n_ops, n_tensors=n_tensors, n_slots=n_slots, max_crops=max_crops,
                     flip_test=int(spec.flip_test), shift_heatmap=int(spec.shift_heatmap),
                     post_process=_lib.PE_POST[spec.post_process], blur_kernel=spec.modulate_kernel,
                     swap_rb=0 if spec.wrapper_double_swap else 1, use_tensor_cores=int(use_tc), reserved=int(spec.use_udp),
                     padding=spec.padding, pixel_std=200.0)


def box_to_affine(spec: TopDownSpec, bbox_xywh):
    """Host-side a5/a6 maths of the library (no GPU needed): -> center f32(2), scale f32(2), trans f64(2,3)."""
    lib = _lib.load()
    d = model_desc(spec)
    bb = np.ascontiguousarray(bbox_xywh, np.float64)
    c, s, t = np.zeros(2, np.float32), np.zeros(2, np.float32), np.zeros(6, np.float64)
    check(lib.pe_box_to_affine(C.byref(d), ptr(bb), ptr(c), ptr(s), ptr(t)))
    return c, s, t.reshape(2, 3)


class TopDownModel:
    """mmpose ``init_pose_model`` + ``inference_top_down_pose_model`` replacement for one method."""

    def __init__(self, engine: PoseEngine, state_dict: Dict[str, np.ndarray], spec: TopDownSpec, max_crops: int = 32,
                 use_tensor_cores: bool = True, unique_slots: bool = False):
        self.engine, self.spec, self.lib = engine, spec, engine.lib
        self.max_crops = max_crops
        prog = program_for(spec)
        self.program = prog
        missing = [k for k in prog.params if k not in state_dict and not k.endswith("num_batches_tracked")]
        if missing:
            raise KeyError(f"checkpoint is missing {len(missing)} tensors, e.g. {missing[:3]}")
        blob = WeightBlob()
        ops = (OpDesc * len(prog.ops))()
        for i, op in enumerate(prog.ops):
            o = ops[i]
            o.kind, o.out = op.kind, op.out
            ins = list(op.ins) + [-1] * (4 - len(op.ins))
            ups = list(op.ups) + [1] * (4 - len(op.ups))
            for j in range(4):
                o.inp[j], o.up[j] = ins[j], ups[j]
            o.n_in, o.ksize, o.stride, o.cin, o.cout = len(op.ins), op.ksize, op.stride, op.cin, op.cout
            o.relu, o.residual, o.reserved = int(op.relu), op.residual, 0
            o.w_off = o.b_off = 0
            o.wtc_off = -1
            if op.kind == OP_GEMM:                                   # Linear (or the patch-embedding conv as a GEMM)
                w = np.asarray(state_dict[f"{op.conv}.weight"], np.float32).reshape(op.cout, op.cin)
                o.b_off = blob.add(np.asarray(state_dict[f"{op.conv}.bias"], np.float32))
                o.wtc_off = blob.add(pack_tc_weights(w[:, :, None]))
                if getattr(op, "const_table", None):                 # x + pos_embed[:, 1:] + pos_embed[:, :1]
                    pos = np.asarray(state_dict[op.const_table], np.float32)
                    o.w_off = blob.add((pos[0, 1:] + pos[0, :1]).astype(np.float32))
                    o.reserved = 1
                continue
            if op.kind == OP_LN:
                o.w_off = blob.add(np.asarray(state_dict[f"{op.conv}.weight"], np.float32))
                o.b_off = blob.add(np.asarray(state_dict[f"{op.conv}.bias"], np.float32))
                continue
            if op.kind in (OP_PATCH, OP_ATTN, OP_D2S):
                continue
            if op.kind in (OP_STEM, OP_CONV, OP_HEAD):
                w = state_dict[f"{op.conv}.weight"]
                if getattr(op, "deconv", False):
                    w = deconv_as_conv_weights(np.asarray(w, np.float32))
                    bn = bn_of(state_dict, op.bn)
                    bn = {k: np.tile(v, 4) for k, v in bn.items()}                   # the same BN for each of the 4 parity blocks
                    wf, bf = fold_bn(w, bn, None)
                    o.b_off = blob.add(bf)
                    o.w_off = 0
                    o.wtc_off = blob.add(pack_tc_weights(wf))
                    continue
                cb = state_dict.get(f"{op.conv}.bias") if op.has_bias else None
                wf, bf = fold_bn(w, bn_of(state_dict, op.bn) if op.bn else None, cb)
                k = op.ksize
                if op.kind == OP_HEAD:
                    o.w_off = blob.add(wf.reshape(op.cout, op.cin).T)                      # [Cin][K]
                else:
                    o.w_off = blob.add(wf.transpose(2, 3, 1, 0).reshape(k * k, op.cin, op.cout))   # [tap][Cin][Cout]
                o.b_off = blob.add(bf)
                if use_tensor_cores and op.kind == OP_CONV and op.cin % 16 == 0 and op.cout % 16 == 0:
                    if op.stride == 1:
                        o.wtc_off = blob.add(pack_tc_weights(wf))
                    elif op.stride == 2 and k == 3:
                        o.wtc_off = blob.add(pack_tc_weights(space_to_depth_weights(wf)))
        if unique_slots:
            slot_of = list(range(len(prog.tensors)))
            slot_elems = [((t.H + 2) * (t.W + 2) if t.W else t.H) * t.C for t in prog.tensors]
        else:
            slot_of, slot_elems = prog.assign_slots()
        tens = (TensorDesc * len(prog.tensors))()
        for t in prog.tensors:
            tens[t.tid].C, tens[t.tid].H, tens[t.tid].W, tens[t.tid].slot = t.C, t.H, t.W, slot_of[t.tid]
        self.slot_elems = np.asarray(slot_elems, np.int64)
        desc = model_desc(spec, len(prog.ops), len(prog.tensors), len(slot_elems), max_crops, use_tensor_cores)
        lut = np.stack([((np.arange(256, dtype=np.float32) / np.float32(255.0)) - np.float32(spec.mean[c])) / np.float32(spec.std[c])
                        for c in range(3)]).astype(np.float32)
        perm = np.arange(spec.num_joints, dtype=np.int32)
        for a, b in spec.flip_pairs:
            perm[a], perm[b] = b, a
        weights = blob.array()
        self.weight_floats = weights.size
        h = C.c_void_p()
        check(self.lib.pe_model_create(engine.h, C.byref(desc), ops, tens, ptr(self.slot_elems), ptr(weights), weights.size,
                                       ptr(lut), ptr(perm), C.byref(h)))
        self.h = h
        self.desc = desc
        _lib.track(self)

    # ---- the hot path
    def topdown(self, frame_idx: Sequence[int], bboxes_xywh: np.ndarray) -> np.ndarray:
        """(n,) staged-frame indices + (n,4) float64 x,y,w,h -> (n,K,3) float32 [x_px, y_px, score]."""
        fi = np.ascontiguousarray(frame_idx, np.int32)
        bb = np.ascontiguousarray(bboxes_xywh, np.float64).reshape(-1, 4)
        out = np.empty((len(fi), self.spec.num_joints, 3), np.float32)
        if len(fi):
            check(self.lib.pe_topdown(self.h, ptr(fi), ptr(bb), len(fi), ptr(out)))
        return out

    def topdown_async(self, frame_idx, bboxes_xywh, out_pinned_ptr: int = 0):
        fi = np.ascontiguousarray(frame_idx, np.int32)
        bb = np.ascontiguousarray(bboxes_xywh, np.float64).reshape(-1, 4)
        check(self.lib.pe_topdown_async(self.h, ptr(fi), ptr(bb), len(fi), C.c_void_p(out_pinned_ptr) if out_pinned_ptr else None))

    # ---- parity hooks
    def warp_crops(self, frame_idx, bboxes_xywh):
        fi = np.ascontiguousarray(frame_idx, np.int32)
        bb = np.ascontiguousarray(bboxes_xywh, np.float64).reshape(-1, 4)
        n = len(fi)
        W, H = self.spec.image_size
        crops = np.empty((n, H, W, 3), np.uint8)
        c, s = np.empty((n, 2), np.float32), np.empty((n, 2), np.float32)
        check(self.lib.pe_warp_crops(self.h, ptr(fi), ptr(bb), n, ptr(crops), ptr(c), ptr(s)))
        return crops, c, s

    def forward_heatmaps(self, crops_u8: np.ndarray):
        crops = np.ascontiguousarray(crops_u8, np.uint8)
        n = crops.shape[0]
        W, H = self.spec.heatmap_size
        hm = np.empty((n, self.spec.num_joints, H, W), np.float32)
        hmf = np.empty_like(hm) if self.spec.flip_test else None
        check(self.lib.pe_forward_heatmaps(self.h, ptr(crops), n, ptr(hm), ptr(hmf) if hmf is not None else None))
        return hm, hmf

    def decode_heatmaps(self, hm, hm_flipped, center, scale):
        hm = np.ascontiguousarray(hm, np.float32)
        n = hm.shape[0]
        hf = np.ascontiguousarray(hm_flipped, np.float32) if hm_flipped is not None else None
        c, s = np.ascontiguousarray(center, np.float32), np.ascontiguousarray(scale, np.float32)
        out = np.empty((n, self.spec.num_joints, 3), np.float32)
        check(self.lib.pe_decode_heatmaps(self.h, ptr(hm), ptr(hf) if hf is not None else None, ptr(c), ptr(s), n, ptr(out)))
        return out

    def debug_tensor(self, tensor_id: int, img: int = 0) -> np.ndarray:
        t = self.program.tensors[tensor_id]
        out = np.empty((t.C, t.H, t.W) if t.W else (t.H, t.C), np.float32)      # W == 0: token matrix [tokens][C]
        check(self.lib.pe_debug_tensor(self.h, tensor_id, img, ptr(out)))
        return out

    def launch_count(self) -> int:
        v = C.c_int64()
        check(self.lib.pe_model_launch_count(self.h, C.byref(v)))
        return v.value

    def profile(self, enable):
        check(self.lib.pe_model_profile(self.h, int(enable)))

    def profile_ops(self) -> np.ndarray:
        ms = np.zeros(len(self.program.ops), np.float64)
        check(self.lib.pe_model_profile_ops(self.h, ptr(ms), len(ms)))
        return ms

    def profile_read(self):
        a, b, n = C.c_double(), C.c_double(), C.c_int64()
        check(self.lib.pe_model_profile_read(self.h, C.byref(a), C.byref(b), C.byref(n)))
        return a.value, b.value, n.value

    def close(self):
        if getattr(self, "h", None):
            self.lib.pe_model_destroy(self.h)      # no-op in C if the engine was destroyed first
            self.h = None

    def __del__(self):
        if _lib.finalizing():
            return
        try:
            self.close()
        except Exception:
            pass


class Lifter:
    """VideoPose3D TemporalModelOptimized1f (243 frames) on the GPU (reference wrappers/videopose3d.py:46-85)."""

    def __init__(self, engine: PoseEngine, state_dict: Dict[str, np.ndarray], channels: int = 1024):
        self.engine, self.lib = engine, engine.lib
        blob = WeightBlob()
        offs: List[int] = []

        def add(wname, bnname, pad_in=None, pad_out=None, bias=None, tc=True):
            w = state_dict[wname]                                        # (Cout, Cin, k)
            wf, bf = fold_bn(w, bn_of(state_dict, bnname) if bnname else None, bias)
            cout, cin, k = wf.shape
            ci, co = pad_in or cin, pad_out or cout
            wp = np.zeros((k, ci, co), np.float32)
            wp[:, :cin, :cout] = wf.transpose(2, 1, 0)
            bp = np.zeros(co, np.float32)
            bp[:cout] = bf
            offs.extend([blob.add(wp), blob.add(bp)])
            if tc:                                                      # tensor-core packing [tap][Cin/16][Cout][h16|l16]
                wt = np.zeros((co, ci, k), np.float32)
                wt[:cout, :cin] = wf
                offs.append(blob.add(pack_tc_weights(wt)))
            else:
                offs.append(-1)

        add("expand_conv.weight", "expand_bn", pad_in=48)
        for i in range(4):
            add(f"layers_conv.{2 * i}.weight", f"layers_bn.{2 * i}")
            add(f"layers_conv.{2 * i + 1}.weight", f"layers_bn.{2 * i + 1}")
        add("shrink.weight", None, pad_out=64, bias=state_dict["shrink.bias"], tc=False)     # 0.4 % of the MACs, fp32 rows out: SIMT
        w = blob.array()
        offs_a = np.asarray(offs, np.int64)
        h = C.c_void_p()
        check(self.lib.pe_lifter_create(engine.h, ptr(w), w.size, ptr(offs_a), len(offs), channels, C.byref(h)))
        self.h = h
        _lib.track(self)

    def uses_tensor_cores(self) -> bool:
        return bool(self.lib.pe_lifter_uses_tensor_cores(self.h))

    def launch_count(self) -> int:
        v = C.c_int64()
        check(self.lib.pe_lifter_launch_count(self.h, C.byref(v)))
        return v.value

    def lift(self, kp2d_norm: np.ndarray) -> np.ndarray:
        """(N,17,2) normalised screen coordinates -> (N,17,3) float32."""
        x = np.ascontiguousarray(kp2d_norm, np.float32).reshape(-1, 34)
        out = np.empty((x.shape[0], 17, 3), np.float32)
        check(self.lib.pe_lift3d(self.h, ptr(x), x.shape[0], ptr(out)))
        return out

    def close(self):
        if getattr(self, "h", None):
            self.lib.pe_lifter_destroy(self.h)
            self.h = None

    def __del__(self):
        if _lib.finalizing():
            return
        try:
            self.close()
        except Exception:
            pass


def person_bbox(tracks, keep_tracks):
    """``PersonBbox.make`` arithmetic (reference pipeline.py:661-685) through the C ABI (host-only, bit-exact)."""
    lib = _lib.load()
    if len(tracks) == 0:
        raise IndexError("list index out of range")                    # reference: LD[0] on an empty video
    counts = np.asarray([len(f) for f in tracks], np.int32)
    ids = np.asarray([int(t["track_id"]) for f in tracks for t in f], np.int64)
    boxes = [np.asarray(t["tlhw"]) for f in tracks for t in f]
    tl = np.asarray(boxes, np.float64).reshape(-1, 4) if boxes else np.zeros((0, 4), np.float64)
    keep = np.asarray(list(keep_tracks), np.int64).reshape(-1)
    n = len(tracks)
    bbox = np.empty((n, 4), np.float64)
    present = np.empty(n, np.uint8)
    check(lib.pe_person_bbox(ptr(counts), n, ptr(ids), ptr(tl), ptr(keep), len(keep), ptr(bbox), ptr(present)))
    present = present.astype(bool)
    # np.array(list-of-bbox) in the reference is float32 only if every frame contributed a float32 tlhw array
    if boxes and all(b.dtype == np.float32 for b in boxes):
        raw_present = np.asarray([sum(int(t["track_id"]) in set(keep.tolist()) for t in f) == 1 for f in tracks])
        if raw_present.all():
            bbox = bbox.astype(np.float32)
    return bbox, present
