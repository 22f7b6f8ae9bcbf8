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

from .. import frames as F
from ..sharding import dist_info, gather_rows, max_over_ranks, shard_range

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
    fh, fw = int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT)), int(cap.get(cv2.CAP_PROP_FRAME_WIDTH))
    cap.release()

    engine = getattr(detector, "engine", None)
    rank, world = dist_info()
    start, stop = shard_range(video_length, rank, world)
    # decoded frames go straight into the HBM-resident cache: the pose pass of this video then needs no second decode
    fp = F.fingerprint(file_path) if engine is not None and hasattr(engine, "upload_block_to") else None
    writer = F.CACHE.begin(fp, stop - start, fh, fw, getattr(engine, "device", 0), first=start) if fp else None
    reader = F.BlockReader(file_path, engine, FRAME_BLOCK, start, stop, cache_writer=writer)
    per_frame = []                                 # detections of this rank's frames, in order
    try:
        for blk in reader:
            if blk.n == 0:                         # read failure: the reference breaks out of its loop (:40-41)
                break
            if hasattr(detector, "detect_block"):
                per_frame.extend(detector.detect_block(reader, blk))
            else:
                per_frame.extend(detector.detect(blk.frames))
            if not blk.complete:
                break
    finally:
        reader.close()
    if fp and world == 1 and len(per_frame) == video_length:
        F.mark_valid(fp)                           # every frame decoded: a later get_robust_reader need not test-decode again

    if world > 1:
        # the detector shards by frame; the association is sequential in time and costs microseconds per frame, so every
        # rank gathers all detections and runs the same tracker: identical tracks everywhere, no broadcast
        cap_det = max(1, max_over_ranks(max([len(d) for d in per_frame], default=0)))
        local = np.zeros((stop - start, cap_det * 5 + 1), np.float32)
        local[:, -1] = -1.0                        # -1 = this frame could not be read
        for i, d in enumerate(per_frame):
            local[i, : len(d) * 5] = np.asarray(d, np.float32).ravel()
            local[i, -1] = len(d)
        full = gather_rows(local, video_length)
        per_frame = []
        for row in full:
            if row[-1] < 0:
                break
            per_frame.append(row[: int(row[-1]) * 5].reshape(-1, 5))

    tracks = []
    try:
        for frame_id, dets in enumerate(per_frame):
            track_results = tracker.update(frame_id, dets)
            tracks.append(tracks_from_rows(track_results))
    finally:
        tracker.close()

    return tracks
