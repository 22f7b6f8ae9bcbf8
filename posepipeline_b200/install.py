"""Make an existing PosePipeline checkout use this engine without editing it.

    import posepipeline_b200.install as pe; pe.install()

registers this package's wrappers under the module names the reference's ``make()`` methods import lazily
(``pose_pipeline/pipeline.py:526,1021,1271``: ``from .wrappers.mmpose import mmpose_top_down_person`` ...), and swaps the
arithmetic of ``PersonBbox.make`` (``pipeline.py:656-687``) for the bit-exact C-ABI restatement (which also runs on
pandas >= 2.1, where the reference's ``fillna(method=...)`` raises).  DataJoint tables, ``populate()`` and
``standard_pipelines.py`` stay the reference's own.
"""
from __future__ import annotations

import sys


def person_bbox_make(self, key):
    """Body of ``PersonBbox.make`` with the per-frame selection / NaN mask / bfill(2) / ffill(2) done by pe_person_bbox."""
    from pose_pipeline.pipeline import PersonBboxValid, TrackingBbox
    from .engine import person_bbox
    tracks = (TrackingBbox & key).fetch1("tracks")
    keep_tracks = (PersonBboxValid & key).fetch1("keep_tracks")
    bbox, present = person_bbox(tracks, keep_tracks)
    key["present"] = present
    key["bbox"] = bbox
    self.insert1(key)


def install(patch_person_bbox: bool = True):
    from .wrappers import mmpose, mmtrack, videopose3d
    sys.modules["pose_pipeline.wrappers.mmpose"] = mmpose
    sys.modules["pose_pipeline.wrappers.videopose3d"] = videopose3d
    # mmtrack is only an interface mirror in this round: leave the reference's tracker in place
    try:
        import pose_pipeline.wrappers as W
        W.mmpose, W.videopose3d = mmpose, videopose3d
    except Exception:
        pass
    if patch_person_bbox:
        import pose_pipeline.pipeline as P
        P.PersonBbox.make = person_bbox_make
    return mmpose, videopose3d
