"""Drop-in for ``pose_pipeline/wrappers/videopose3d.py`` (reference :19-91): same signature and return dict.

Kept: reads ``(TopDownPerson & key).fetch1("keypoints")`` and ``(VideoInfo & key).fetch1("height", "width")``; normalises
x/y exactly like ``normalize_screen_coordinates`` (:26-33, float64 numpy); drops confidences; feeds COCO-ordered joints and
zero rows as they are (quirk Q9); returns ``{"keypoints_3d": float64 (N,17,3), "keypoints_valid": [True]*N}`` (:87-91).
Changed: the 243-frame windows are not materialised -- the engine evaluates the dilated form of the same network over the
edge-padded sequence on the GPU (csrc/lifter.cu); ``batch_size`` is accepted and ignored (it only chunked the CPU loop).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional

import numpy as np

from .. import engine as E
from . import mmpose as _mm


@dataclass
class VideoPoseArgs:                       # reference :10-16
    causal: bool = False
    architecture: str = "3,3,3,3,3"
    dropout: float = 0.25
    channels: int = 1024
    dense: bool = False


_lifter: Optional["E.Lifter"] = None


def get_lifter():
    global _lifter
    if _lifter is None:
        ckpt = os.path.join(_mm._model_data_dir(), "videopose3d/pretrained_h36m_detectron_coco.bin")   # reference :52-54
        if os.path.exists(ckpt):
            from ..weights import load_checkpoint
            sd = load_checkpoint(ckpt)
        elif os.environ.get("PE_SYNTHETIC_WEIGHTS") == "1":
            from ..weights import synthetic_videopose3d_state_dict
            sd = synthetic_videopose3d_state_dict(0)
        else:
            raise FileNotFoundError(f"{ckpt} not found (set PE_SYNTHETIC_WEIGHTS=1 to run with seeded synthetic weights)")
        _lifter = E.Lifter(_mm.get_engine(), sd, VideoPoseArgs().channels)
    return _lifter


def normalize_screen_coordinates(X, w, h):
    assert X.shape[-1] == 2
    # Normalize so that [0, w] is mapped to [-1, 1], while preserving the aspect ratio
    if w > h:
        return X / w * 2 - [1, h / w]
    else:
        return X / h * 2 - [w / h, 1]


def process_videopose3d(key, batch_size=32, transform_coco=False):
    from pose_pipeline import TopDownPerson, VideoInfo

    keypoints = (TopDownPerson & key).fetch1("keypoints")
    height, width = (VideoInfo & key).fetch1("height", "width")
    N = keypoints.shape[0]
    keypoints = normalize_screen_coordinates(keypoints[:, :, :2], width, height)
    keypoints = keypoints[:, :, :2]
    valid_frames = np.arange(keypoints.shape[0])

    results = get_lifter().lift(keypoints.astype("float32"))

    keypoints_3d = np.zeros((N, 17, 3))
    keypoints_3d[valid_frames] = results
    keypoints_valid = [i in valid_frames.tolist() for i in np.arange(keypoints.shape[0])]
    return {"keypoints_3d": keypoints_3d, "keypoints_valid": keypoints_valid}
