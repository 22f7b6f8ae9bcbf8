"""Weights for the engine: checkpoint loading, seeded synthetic weights, BN folding.

The reference loads ``$MODEL_DATA_DIR/mmpose/checkpoints/hrnet_w48_coco_384x288_dark-e881a4b6_20210203.pth``
(``pose_pipeline/wrappers/mmpose.py:35``) and ``$MODEL_DATA_DIR/videopose3d/pretrained_h36m_detectron_coco.bin``
(``wrappers/videopose3d.py:52-57``).  Neither file exists on this box and there is no network
(SURVEY fact 3), so tests and the benchmark use seeded synthetic tensors stored under exactly the
upstream ``state_dict`` key names; ``load_checkpoint`` accepts the real files unchanged.
"""
from __future__ import annotations

import math
import os
import zlib
from typing import Dict, Optional

import numpy as np

from .hrnet_spec import Program, build_program

BN_EPS = 1e-5
_CALIB = os.path.join(os.path.dirname(__file__), "data", "synthetic_head_calibration.npz")


def _rng(name: str, seed: int) -> np.random.Generator:
    return np.random.default_rng([zlib.crc32(name.encode()), seed])


def synthetic_hrnet_state_dict(program: Program, seed: int = 0, calibrated: bool = True) -> Dict[str, np.ndarray]:
    """Seeded weights under mmpose key names.  Scales are chosen so activations stay O(1) through the
    ~100 residual layers.  With ``calibrated`` the last fuse bias and the head come from a committed
    fixture (posepipeline_b200/data/, made by tests/golden/make_synth_calibration.py) that makes the
    heatmaps sparse and peaky like a trained network's -- without that, DARK's Taylor step is
    ill-conditioned and even fp32-vs-fp64 runs of the same network disagree by >1e-3 px."""
    sd: Dict[str, np.ndarray] = {}
    for name, shape in program.params.items():
        parent, leaf = name.rsplit(".", 1)
        r = _rng(name, seed)
        if leaf == "num_batches_tracked":
            sd[name] = np.zeros((), np.int64)
        elif len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            if parent.endswith("final_layer"):
                w = r.random(shape) * (2.0 / fan_in)
            else:
                w = r.standard_normal(shape) * math.sqrt(2.0 / fan_in)
            sd[name] = w.astype(np.float32)
        elif leaf == "weight":
            lo, hi = 0.8, 1.2
            if (parent.endswith("bn2") and ".branches." in parent) or parent.endswith("bn3"):
                lo, hi = 0.2, 0.35          # damp the residual branch
            elif ".fuse_layers." in parent:
                lo, hi = 0.15, 0.3          # damp the multi-branch sums
            sd[name] = (r.random(shape) * (hi - lo) + lo).astype(np.float32)
        elif leaf == "bias":
            if parent.endswith("final_layer"):
                sd[name] = (r.random(shape) * 0.02).astype(np.float32)
            else:
                sd[name] = ((r.random(shape) - 0.5) * 0.2).astype(np.float32)
        elif leaf == "running_mean":
            sd[name] = ((r.random(shape) - 0.5) * 0.2).astype(np.float32)
        elif leaf == "running_var":
            sd[name] = (r.random(shape) * 0.5 + 0.75).astype(np.float32)
        else:
            raise KeyError(name)
    if calibrated:
        apply_head_calibration(sd, program, seed)
    return sd


def apply_head_calibration(sd, program: Program, seed: int):
    key = f"{program.variant}_{program.in_h}x{program.in_w}_k{program.num_joints}_s{seed}"
    if not os.path.exists(_CALIB):
        raise FileNotFoundError(f"{_CALIB} missing: run tests/golden/make_synth_calibration.py")
    z = np.load(_CALIB)
    if f"{key}/fuse_bias" not in z.files:
        raise KeyError(f"no synthetic head calibration for {key}; run tests/golden/make_synth_calibration.py")
    nm = len([k for k in program.params if k.endswith("fuse_layers.0.1.1.bias") and ".stage4." in k])
    sd[f"backbone.stage4.{nm - 1}.fuse_layers.0.1.1.bias"] = z[f"{key}/fuse_bias"].astype(np.float32)
    sd["keypoint_head.final_layer.weight"] = z[f"{key}/head_weight"].astype(np.float32)
    sd["keypoint_head.final_layer.bias"] = z[f"{key}/head_bias"].astype(np.float32)


def load_checkpoint(path: str) -> Dict[str, np.ndarray]:
    """Read an mmpose ``{'state_dict':…, 'meta':…}`` / VideoPose3D ``{'model_pos':…}`` / bare state_dict file."""
    import torch
    ck = torch.load(path, map_location="cpu", weights_only=False)
    for k in ("state_dict", "model_pos"):
        if isinstance(ck, dict) and k in ck:
            ck = ck[k]
            break
    return {k: v.detach().cpu().numpy() for k, v in ck.items()}


def fold_bn(w: np.ndarray, bn: Optional[Dict[str, np.ndarray]], conv_bias: Optional[np.ndarray] = None,
            eps: float = BN_EPS):
    """Eval-mode BN is affine: fold it into the conv.  Done in float64, returned as float32.
    w: (Cout, Cin, k, k) or (Cout, Cin, k).  Returns (w', b')."""
    w64 = w.astype(np.float64)
    cout = w.shape[0]
    b64 = np.zeros(cout) if conv_bias is None else conv_bias.astype(np.float64)
    if bn is not None:
        g = bn["weight"].astype(np.float64) / np.sqrt(bn["running_var"].astype(np.float64) + eps)
        w64 = w64 * g.reshape((cout,) + (1,) * (w.ndim - 1))
        b64 = (b64 - bn["running_mean"].astype(np.float64)) * g + bn["bias"].astype(np.float64)
    return w64.astype(np.float32), b64.astype(np.float32)


def bn_of(sd: Dict[str, np.ndarray], prefix: str) -> Dict[str, np.ndarray]:
    return {k: sd[f"{prefix}.{k}"] for k in ("weight", "bias", "running_mean", "running_var")}


# ------------------------------------------------------------------ VideoPose3D (App. A.6)
def videopose3d_param_shapes(channels: int = 1024, joints_in: int = 17, feat: int = 2, joints_out: int = 17,
                             widths=(3, 3, 3, 3, 3)):
    shapes = {"expand_conv.weight": (channels, joints_in * feat, widths[0])}

    def bn(name):
        for leaf in ("weight", "bias", "running_mean", "running_var"):
            shapes[f"{name}.{leaf}"] = (channels,)
        shapes[f"{name}.num_batches_tracked"] = ()
    bn("expand_bn")
    for i in range(len(widths) - 1):
        shapes[f"layers_conv.{2 * i}.weight"] = (channels, channels, widths[i + 1])
        shapes[f"layers_conv.{2 * i + 1}.weight"] = (channels, channels, 1)
    for i in range(2 * (len(widths) - 1)):
        bn(f"layers_bn.{i}")
    shapes["shrink.weight"] = (joints_out * 3, channels, 1)
    shapes["shrink.bias"] = (joints_out * 3,)
    return shapes


def synthetic_videopose3d_state_dict(seed: int = 0, channels: int = 1024) -> Dict[str, np.ndarray]:
    sd = {}
    for name, shape in videopose3d_param_shapes(channels).items():
        parent, leaf = name.rsplit(".", 1)
        r = _rng(name, seed)
        if leaf == "num_batches_tracked":
            sd[name] = np.zeros((), np.int64)
        elif len(shape) == 3:
            fan_in = shape[1] * shape[2]
            sd[name] = (r.standard_normal(shape) * math.sqrt(2.0 / fan_in)).astype(np.float32)
        elif leaf == "weight":
            lo, hi = (0.3, 0.5) if parent.startswith("layers_bn") and int(parent.split(".")[1]) % 2 == 1 else (0.8, 1.2)
            sd[name] = (r.random(shape) * (hi - lo) + lo).astype(np.float32)
        elif leaf == "bias":
            sd[name] = ((r.random(shape) - 0.5) * 0.2).astype(np.float32)
        elif leaf == "running_mean":
            sd[name] = ((r.random(shape) - 0.5) * 0.2).astype(np.float32)
        elif leaf == "running_var":
            sd[name] = (r.random(shape) * 0.5 + 0.75).astype(np.float32)
    return sd


# ------------------------------------------------------------------ ViTPose-B (App. A.4)
def synthetic_vitpose_state_dict(program: Program, seed: int = 0, calibrated: bool = True) -> Dict[str, np.ndarray]:
    """Seeded ViTPose-B tensors under the upstream key names (there is no checkpoint offline).  Linear layers get
    a ViT-style N(0, s^2) initialisation scaled so the residual stream stays O(1) over the 12 blocks; LayerNorm affine
    parameters are perturbed around (1, 0) so they are exercised.  With ``calibrated`` the 1x1 final layer comes from the
    committed fixture (tests/golden/make_synth_calibration.py) that turns the head features into sparse, peaky heatmaps."""
    sd: Dict[str, np.ndarray] = {}
    for name, shape in program.params.items():
        parent, leaf = name.rsplit(".", 1)
        r = _rng(name, seed)
        if leaf == "num_batches_tracked":
            sd[name] = np.zeros((), np.int64)
        elif name.endswith("pos_embed"):
            sd[name] = (r.standard_normal(shape) * 0.2).astype(np.float32)
        elif "deconv_layers" in parent and len(shape) == 4:
            fan_in = shape[0] * 4                                    # ConvTranspose k4 s2: 4 taps of every input channel per output pixel
            sd[name] = (r.standard_normal(shape) * math.sqrt(2.0 / fan_in)).astype(np.float32)
        elif len(shape) in (2, 4):
            fan_in = int(np.prod(shape[1:]))
            gain = 0.5 if (parent.endswith("attn.proj") or parent.endswith("mlp.fc2")) else 1.0    # damp the residual branches
            sd[name] = (r.standard_normal(shape) * gain / math.sqrt(fan_in)).astype(np.float32)
        elif leaf == "weight":
            sd[name] = (r.random(shape) * 0.4 + 0.8).astype(np.float32)
        elif leaf == "bias":
            sd[name] = ((r.random(shape) - 0.5) * 0.2).astype(np.float32)
        elif leaf == "running_mean":
            sd[name] = ((r.random(shape) - 0.5) * 0.2).astype(np.float32)
        elif leaf == "running_var":
            sd[name] = (r.random(shape) * 0.5 + 0.75).astype(np.float32)
        else:
            raise KeyError(name)
    if calibrated:
        key = f"vitpose_b_{program.in_h}x{program.in_w}_k{program.num_joints}_s{seed}"
        z = np.load(_CALIB)
        if f"{key}/head_weight" not in z.files:
            raise KeyError(f"no synthetic head calibration for {key}; run tests/golden/make_synth_calibration.py")
        sd["keypoint_head.final_layer.weight"] = z[f"{key}/head_weight"].astype(np.float32)
        sd["keypoint_head.final_layer.bias"] = z[f"{key}/head_bias"].astype(np.float32)
    return sd
