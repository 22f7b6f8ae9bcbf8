"""Drop-in for ``pose_pipeline/wrappers/mmpose.py`` (reference :26-81): same function name, arguments, return value and
error behaviour, with the per-frame mmpose call replaced by the B200 engine (posepipeline_b200.engine -> C ABI).

What is kept exactly:
  * reads ``(PersonBbox & key).fetch1("bbox")`` and ``Video.get_robust_reader(key, return_cap=False)`` itself, deletes the
    temporary video afterwards (reference :53-55, :79);
  * one ``cap.read()`` per bbox row, ``assert ret and frame is not None`` on a short video (:63-64);
  * a NaN bbox yields ``np.zeros((K, 3))`` (float64) for that frame (:67-69), so the returned array is float64 iff any
    frame is absent, else float32 (quirk Q7);
  * the wrapper-level BGR->RGB swap (:73) followed by mmpose's own swap (quirk Q1) -- the engine consumes the BGR frame
    as decoded and normalises channel 0 with the "R" mean/std, which is the same arithmetic.
What changes: frames are processed in blocks (pinned staging -> one H2D copy per block -> batched crops), the model is
built once per process and cached (the reference rebuilds it per video, quirk Q8), and with torch.distributed
initialised each rank handles a contiguous frame range (posepipeline_b200.sharding).
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import cv2
import numpy as np

from .. import engine as E
from .. import frames as F
from ..sharding import dist_info, gather_rows, shard_range

# reference wrappers/mmpose.py:8-24 (label strings consumed by TopDownPerson.joint_names, pipeline.py:1139-1141)
mmpose_joint_dictionary = {
    'MMPoseWholebody': ["Nose", "Left Eye", "Right Eye", "Left Ear", "Right Ear", "Left Shoulder", "Right Shoulder",
                        "Left Elbow", "Right Elbow", "Left Wrist", "Right Wrist", "Left Hip", "Right Hip", "Left Knee",
                        "Right Knee", "Left Ankle", "Right Ankle", "Left Big Toe", "Left Little Toe", "Left Heel",
                        "Right Big Toe", "Right Little Toe", "Right Heel"],
    'MMPoseHalpe': ["Nose", "Left Eye", "Right Eye", "Left Ear", "Right Ear", "Left Shoulder", "Right Shoulder",
                    "Left Elbow", "Right Elbow", "Left Wrist", "Right Wrist", "Left Hip", "Right Hip", "Left Knee",
                    "Right Knee", "Left Ankle", "Right Ankle", "Head", "Neck", "Pelvis", "Left Big Toe",
                    "Right Big Toe", "Left Little Toe", "Right Little Toe", "Left Heel", "Right Heel"],
    'MMPose': ["Nose", "Left Eye", "Right Eye", "Left Ear", "Right Ear", "Left Shoulder", "Right Shoulder", "Left Elbow",
               "Right Elbow", "Left Wrist", "Right Wrist", "Left Hip", "Right Hip", "Left Knee", "Right Knee",
               "Left Ankle", "Right Ankle"],
}

# tags the reference accepts (:33-52) that this build does not implement yet (HRFormer / TCFormer are
# different backbones and out of the hot-path scope)
_REFERENCE_ONLY = {"HRFormer_COCO": 17, "HRNet_TCFormer_COCOWholeBody": 133}
_reference_impl = None                 # set by posepipeline_b200.install when the reference's mmpose wrapper is importable

FRAME_BLOCK = int(os.environ.get("PE_FRAME_BLOCK", "32"))
_models: Dict[str, "E.TopDownModel"] = {}
_engine: Optional["E.PoseEngine"] = None


def _model_data_dir():
    try:
        from pose_pipeline import MODEL_DATA_DIR
        return MODEL_DATA_DIR
    except Exception:
        return os.environ.get("PIPELINE_3RDPARTY", "")


def get_engine() -> "E.PoseEngine":
    global _engine
    if _engine is None:
        _engine = E.PoseEngine(int(os.environ.get("LOCAL_RANK", "0")))
    return _engine


def get_model(method: str):
    """Process-level cache keyed by method tag (the reference reloads weights for every video, Q8)."""
    if method not in _models:
        spec = E.spec_for(method, _model_data_dir())      # honours an edited $MODEL_DATA_DIR/mmpose/config/**
        ckpt = os.path.join(_model_data_dir(), spec.checkpoint)
        if os.path.exists(ckpt):
            from ..weights import load_checkpoint
            sd = load_checkpoint(ckpt)
        elif os.environ.get("PE_SYNTHETIC_WEIGHTS") == "1":
            from ..weights import synthetic_hrnet_state_dict, synthetic_vitpose_state_dict
            prog = E.program_for(spec)
            sd = synthetic_vitpose_state_dict(prog, 0) if spec.variant == "vitpose_b" else synthetic_hrnet_state_dict(prog, 0)
        else:
            raise FileNotFoundError(f"{ckpt} not found (set PE_SYNTHETIC_WEIGHTS=1 to run with seeded synthetic weights)")
        _models[method] = E.TopDownModel(get_engine(), sd, spec, max_crops=int(os.environ.get("PE_MAX_CROPS", "32")))
    return _models[method]


def mmpose_top_down_person(key, method='HRNet_W48_COCO'):
    from pose_pipeline import Video, PersonBbox      # the reference's own tables (pipeline.py:24, :648)

    if method in _REFERENCE_ONLY:
        if _reference_impl is not None:          # the reference's own mmpose install serves the backbones this engine does not build
            return _reference_impl(key, method)
        raise NotImplementedError(f"top-down method {method} is not built in this engine yet (HRNet_W48_COCO / _COCOWholeBody / _HALPE, HRNet_W32_COCO and ViTPose_B_COCO are)")
    if method not in E.METHODS:
        # the reference falls through its if/elif chain and dies on an unbound `pose_cfg`
        raise UnboundLocalError(f"cannot access local variable 'pose_cfg': unknown top-down method {method!r}")
    num_keypoints = E.METHODS[method].num_joints

    bboxes = (PersonBbox & key).fetch1("bbox")
    video = Video.get_robust_reader(key, return_cap=False)  # returning video allows deleting it
    reader = None
    try:
        model = get_model(method)
        engine = model.engine
        n = len(bboxes)
        rank, world = dist_info()
        start, stop = shard_range(n, rank, world)
        results = []

        def run_block(first, nb):
            # handle the case where person is not tracked in frame
            present = [first + j for j in range(nb) if not np.any(np.isnan(bboxes[first + j]))]
            out_block = [np.zeros((num_keypoints, 3)) for _ in range(nb)]
            if present:
                kp = model.topdown(np.asarray([p - first for p in present], np.int32), np.asarray([bboxes[p] for p in present], np.float64))
                for r, p in enumerate(present):
                    out_block[p - first] = kp[r]
            results.extend(out_block)

        fp = F.fingerprint(video) if hasattr(engine, "stage_frames_device") else None
        cached = F.CACHE.get(fp, start, stop) if fp and stop > start else None
        if cached is not None:
            # this rank's frames are resident in HBM (the tracker pass decoded them): no decode, no host->device copy
            ptr, fh, fw = cached
            i = start
            while i < stop:
                nb = min(FRAME_BLOCK, stop - i)
                engine.stage_frames_device(ptr + (i - start) * fh * fw * 3, nb, fh, fw)
                run_block(i, nb)
                i += nb
        else:
            # the decode thread reads this rank's frame range (seek, not decode-and-drop) one block ahead of the GPU
            reader = F.BlockReader(video, engine, FRAME_BLOCK, start, stop)
            for blk in reader:
                # should match the length of identified person tracks
                assert blk.complete and blk.n > 0
                reader.select(blk)
                run_block(blk.first, blk.n)
            assert len(results) == stop - start
    finally:
        if reader is not None:
            reader.close()
        os.remove(video)

    if world > 1:
        # rows of absent frames are float64 zeros either way; gather float32 model rows + a presence mask
        local = np.asarray([np.asarray(r, np.float32) for r in results], np.float32).reshape(-1, num_keypoints, 3)
        full = gather_rows(local, n)
        absent = np.asarray([bool(np.any(np.isnan(b))) for b in bboxes])
        return full.astype(np.float64) if absent.any() else full
    return np.asarray(results)
