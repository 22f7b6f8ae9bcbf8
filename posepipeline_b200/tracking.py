"""Host-side tracker objects over the C ABI (include/poseengine.h): ByteTrack association.

The reference reaches this arithmetic through ``mmtrack.apis.inference_mot`` (pose_pipeline/wrappers/mmtrack.py:45), whose
ByteTrack model runs ``ByteTracker.track`` once per frame on the detector's boxes.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import check, ptr

# 3rdparty/mmtracking/mot/bytetrack/bytetrack_yolox_x_crowdhuman_mot17-private-half.py:21-28 (+ ByteTracker's default num_tentatives)
BYTETRACK_CFG = dict(obj_score_high=0.6, obj_score_low=0.1, init_track_thr=0.7, match_iou_high=0.1, match_iou_low=0.5,
                     match_iou_tentative=0.3, weight_iou_with_det_scores=True, num_tentatives=3, num_frames_retain=30)


class ByteTracker:
    def __init__(self, **overrides):
        cfg = dict(BYTETRACK_CFG, **overrides)
        self.lib = _lib.load()
        vals = np.asarray([cfg["obj_score_high"], cfg["obj_score_low"], cfg["init_track_thr"], cfg["match_iou_high"],
                           cfg["match_iou_low"], cfg["match_iou_tentative"], float(cfg["weight_iou_with_det_scores"]),
                           cfg["num_tentatives"], cfg["num_frames_retain"]], np.float32)
        h = C.c_void_p()
        check(self.lib.pe_bytetrack_create(ptr(vals), len(vals), C.byref(h)))
        self.h = h

    def reset(self):
        check(self.lib.pe_bytetrack_reset(self.h))

    def update(self, frame_id: int, dets: np.ndarray) -> np.ndarray:
        """dets (n,5) float32 [x1,y1,x2,y2,score] -> (k,6) float64 rows [track_id,x1,y1,x2,y2,score]."""
        d = np.ascontiguousarray(dets, np.float32).reshape(-1, 5)
        out = np.empty((max(len(d), 1), 6), np.float64)
        n = C.c_int32()
        check(self.lib.pe_bytetrack_update(self.h, int(frame_id), ptr(d) if len(d) else None, len(d), ptr(out), len(out), C.byref(n)))
        return out[: n.value].copy()

    def close(self):
        if getattr(self, "h", None):
            self.lib.pe_bytetrack_destroy(self.h)
            self.h = None

    def __del__(self):
        if _lib.finalizing():
            return
        try:
            self.close()
        except Exception:
            pass
