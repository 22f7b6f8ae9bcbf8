"""Drop-in for ``pose_pipeline/wrappers/mmtrack.py`` (reference :8-62): same function name, arguments, return value and
error behaviour.

``mmtrack_bounding_boxes(file_path, "bytetrack")`` runs the B200 detector (YOLOX-X 800x1440 on the tensor-core convolution
kernel, posepipeline_b200.detector -> C ABI) on blocks of frames and the ByteTrack association (posepipeline_b200.tracking,
C++ on the host) frame by frame.  Kept exactly:
  * the accepted method names and the ``Exception(f"Unknown config file for MMTrack method {method}")`` for anything else (:28-29);
  * one ``cap.read()`` per frame of ``CAP_PROP_FRAME_COUNT``; a failed read stops early and returns the shorter list (:37-41);
  * frame ids start at 0 for every video, which resets the tracker (``ByteTrack.simple_test``);
  * per frame a list of ``{"track_id": int, "tlbr": x[1:5], "tlhw": [x1, y1, x2-x1, y2-y1], "confidence": x[5]}`` built from
    float64 rows ``[id, x1, y1, x2, y2, score]`` in the reference's row order (:50-60; "tlhw" holds x,y,w,h -- quirk Q2);
  * the wrapper-level BGR->RGB swap (:43) with ``to_rgb=False`` in the test pipeline: the detector sees R,G,B planes of
    [0, 255] floats, which is how the engine reads the BGR frame.
The Faster R-CNN based methods (tracktor / deepsort / qdtrack) are not built: with a working reference install they go
to the reference's own function (posepipeline_b200.install keeps it), otherwise they raise NotImplementedError.
"""
from __future__ import annotations

import os
from typing import Optional

import cv2
import numpy as np

METHODS = ("tracktor", "deepsort", "bytetrack", "qdtrack")
FRAME_BLOCK = int(os.environ.get("PE_DET_FRAME_BLOCK", "8"))
_reference_impl = None                 # set by posepipeline_b200.install when the reference's mmtrack wrapper is importable
_detector = None


def tracks_from_rows(track_results):
    """reference :50-60: rows [id, x1, y1, x2, y2, score] of one frame -> list of track dicts."""
    return [{"track_id": int(x[0]), "tlbr": x[1:5], "tlhw": np.array([x[1], x[2], x[3] - x[1], x[4] - x[2]]), "confidence": x[5]}
            for x in track_results]


def get_detector():
    """Process-level cache (the reference rebuilds the model for every video, quirk Q8)."""
    global _detector
    if _detector is None:
        from .. import detector as D
        from . import mmpose as _mm
        _detector = D.load_bytetrack_detector(_mm.get_engine(), _mm._model_data_dir())
    return _detector


def mmtrack_bounding_boxes(file_path, method="tracktor"):
    if method not in METHODS:
        raise Exception(f"Unknown config file for MMTrack method {method}")
    if method != "bytetrack":
        if _reference_impl is not None:
            return _reference_impl(file_path, method)
        raise NotImplementedError(f"mmtrack_bounding_boxes({method!r}): only 'bytetrack' is built in this engine "
                                  "(the Faster R-CNN trackers need the reference's mmtrack install)")
    from ..tracking import ByteTracker
    detector = get_detector()
    tracker = ByteTracker()

    cap = cv2.VideoCapture(file_path)
    video_length = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))

    tracks = []
    try:
        frame_id = 0
        done = False
        while frame_id < video_length and not done:
            block = []
            for _ in range(min(FRAME_BLOCK, video_length - frame_id)):
                ret, frame = cap.read()
                if ret != True or frame is None:
                    done = True
                    break
                block.append(frame)
            if not block:
                break
            for dets in detector.detect(np.stack(block)):
                track_results = tracker.update(frame_id, dets)
                tracks.append(tracks_from_rows(track_results))
                frame_id += 1
    finally:
        cap.release()
        tracker.close()

    return tracks
