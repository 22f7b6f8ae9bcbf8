"""Make an existing PosePipeline checkout use this engine without editing it.

    import posepipeline_b200.install as pe; pe.install()

The reference's ``make()`` methods import their wrapper functions lazily (``pose_pipeline/pipeline.py:526,1021,1271``:
``from .wrappers.mmpose import mmpose_top_down_person`` ...).  ``install()`` therefore imports the reference's own wrapper
modules and overwrites ONLY the functions this engine provides, so everything else in them keeps working
(``mmpose_bottom_up``, used by ``BottomUpPeople.make`` at ``pipeline.py:210``, stays the reference's).  A reference wrapper
module that cannot be imported because its third-party dependency is absent (``pose_pipeline/wrappers/mmtrack.py:5`` does
``import mmtrack.apis`` at module level) is replaced by this package's module of the same name.

It also swaps the arithmetic of ``PersonBbox.make`` (``pipeline.py:656-687``) for the bit-exact C-ABI restatement, which
runs on pandas >= 2.1 where the reference's ``fillna(method=...)`` raises.  DataJoint tables, ``populate()`` and
``standard_pipelines.py`` stay the reference's own.
"""
from __future__ import annotations

import importlib
import sys

# reference module -> names this engine replaces in it
PROVIDED = {
    "mmpose": ("mmpose_top_down_person", "mmpose_joint_dictionary"),
    "videopose3d": ("process_videopose3d", "VideoPoseArgs", "normalize_screen_coordinates"),
    "mmtrack": ("mmtrack_bounding_boxes",),
}


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


def _patch_module(name: str):
    """-> (module now registered as pose_pipeline.wrappers.<name>, 'patched' | 'replaced')"""
    ours = importlib.import_module(f"posepipeline_b200.wrappers.{name}")
    full = f"pose_pipeline.wrappers.{name}"
    try:
        ref = importlib.import_module(full)
    except Exception:                      # ImportError of an absent third-party package, or no such reference module
        ref = None
    if ref is None or ref is ours:
        sys.modules[full] = ours
        try:
            import pose_pipeline.wrappers as W
            setattr(W, name, ours)
        except Exception:
            pass
        return ours, "replaced"
    for attr in PROVIDED[name]:
        if hasattr(ours, attr):
            if (name, attr) in (("mmtrack", "mmtrack_bounding_boxes"), ("mmpose", "mmpose_top_down_person")) and not hasattr(ours, "_reference_impl_set"):
                # methods this engine does not build (tracktor / deepsort / qdtrack; HRFormer_COCO / HRNet_TCFormer_COCOWholeBody)
                # keep going to the reference's own function: install() never breaks a path that worked before it
                ours._reference_impl = getattr(ref, attr, None)
                ours._reference_impl_set = True
            setattr(ref, attr, getattr(ours, attr))
    return ref, "patched"


def install(patch_person_bbox: bool = True, patch_robust_reader: bool = True):
    """Returns {module name: 'patched' | 'replaced'}."""
    status = {}
    for name in PROVIDED:
        _, status[name] = _patch_module(name)
    import pose_pipeline.pipeline as P
    if patch_person_bbox:
        P.PersonBbox.make = person_bbox_make
    if patch_robust_reader:
        # same contract as pipeline.py:47-87; the whole-file validation decode is skipped for content this process has
        # already decoded completely (frames.robust_reader) -- one decode per video instead of four
        from . import frames
        P.Video.get_robust_reader = staticmethod(lambda key, return_cap=True: frames.robust_reader(P.Video, key, return_cap))
    return status
