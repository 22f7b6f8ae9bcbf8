"""Interface mirror of ``pose_pipeline/wrappers/mmtrack.py`` (reference :8-62).

The signature, the accepted method names, the unknown-method error and the per-frame output schema
(``{"track_id": int, "tlbr": (4,), "tlhw": [x, y, w, h], "confidence": float}``, quirk Q2) are the reference's.
The detector + tracker itself (YOLOX-X 800x1440 + ByteTrack for "bytetrack"; Faster R-CNN trackers for the others) is the
"next" row f1 of the scope table and is NOT built in this round: the call raises NotImplementedError after
validating its arguments.  PersonBbox / TopDownPerson / LiftingPerson consume stored ``tracks`` and do not need this call.
"""
from __future__ import annotations

METHODS = ("tracktor", "deepsort", "bytetrack", "qdtrack")


def tracks_from_rows(track_results):
    """reference :50-60: rows [id, x1, y1, x2, y2, score] of one frame -> list of track dicts."""
    import numpy as np
    return [{"track_id": int(x[0]), "tlbr": x[1:5], "tlhw": np.array([x[1], x[2], x[3] - x[1], x[4] - x[2]]), "confidence": x[5]}
            for x in track_results]


def mmtrack_bounding_boxes(file_path, method="tracktor"):
    if method not in METHODS:
        raise Exception(f"Unknown config file for MMTrack method {method}")
    raise NotImplementedError(
        f"mmtrack_bounding_boxes({method!r}): the detector+tracker front end is the next scope row (SURVEY §8(f) f1) and is "
        "not part of this build; run the reference tracker (or any tracker) to fill TrackingBbox.tracks")
