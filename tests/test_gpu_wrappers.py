"""-m gpu: the reference-facing entry points end to end (fake DataJoint tables, real video file, real engine) vs the oracle."""
import os

import cv2
import numpy as np
import pytest

import fakes
import helpers
from oracle import topdown as OT
from oracle import videopose3d as OV
from posepipeline_b200.synthetic import synthetic_bboxes, synthetic_keypoints_2d
from posepipeline_b200.weights import synthetic_videopose3d_state_dict

pytestmark = pytest.mark.gpu


@pytest.fixture()
def synthetic_env(monkeypatch):
    monkeypatch.setenv("PE_SYNTHETIC_WEIGHTS", "1")
    monkeypatch.setenv("PE_MAX_CROPS", "4")


def test_mmpose_top_down_person_on_video(tmp_path, synthetic_env):
    """TopDownPerson.make's call: video file + PersonBbox rows (with absent frames) -> (N,17,3), vs the oracle run on the
    frames as DECODED from the same file (lossy codec: never compare against the pre-encode arrays)."""
    from posepipeline_b200.wrappers import mmpose as W
    ns = fakes.make_fake_pose_pipeline()
    src = helpers.frames(3)
    frames = [np.ascontiguousarray(src[i % 3][:720, :1280]) for i in range(7)]
    path = str(tmp_path / "clip.mp4")
    fakes.write_video(path, frames)
    decoded = fakes.read_video(path)
    assert len(decoded) == 7
    key = {"video_project": "t", "filename": "clip"}
    bbox = synthetic_bboxes(7, 5) * np.array([0.6, 0.6, 0.8, 0.8])
    bbox[2] = np.nan
    ns["Video"].rows.append({**key, "video": path})
    ns["PersonBbox"].rows.append({**key, "bbox": bbox})
    got = W.mmpose_top_down_person(key, "HRNet_W48_COCO")
    assert got.shape == (7, 17, 3) and got.dtype == np.float64 and np.all(got[2] == 0)
    net = helpers.oracle_net("HRNet_W48_COCO")
    ref = OT.top_down_video(net, decoded, bbox, OT.HRNET_W48_COCO)
    net64 = helpers.oracle_net("HRNet_W48_COCO", 0, "float64")
    ref64 = OT.top_down_video(net64, decoded, bbox, OT.HRNET_W48_COCO)
    cond = np.abs(ref[..., :2] - ref64[..., :2]).max(-1)
    good = cond <= 1e-4
    d = np.abs(got[..., :2] - ref[..., :2]).max(-1)
    print("wrapper keypoint |dx| px: UNCONDITIONAL max", d.max(), "well-conditioned max", d[good].max(), "fraction", good.mean())
    # 720p crops run off the frame (zero-padded plateaus): fewer well-conditioned maps than in the 1080p parity tests
    assert good.mean() >= 0.8 and d[good].max() <= 1e-3, (d[good].max(), good.mean())
    assert np.all(d[~good] <= 10 * cond[~good] + 1e-3), (d[~good], cond[~good])
    assert np.abs(got[..., 2] - ref[..., 2]).max() <= 1e-4 * max(1.0, np.abs(ref[..., 2]).max())


def test_process_videopose3d_wrapper(synthetic_env):
    from posepipeline_b200.wrappers import videopose3d as V
    ns = fakes.make_fake_pose_pipeline()
    key = {"video_project": "t", "filename": "clip"}
    kp = synthetic_keypoints_2d(150, seed=3)
    kp[10] = 0                                        # an absent frame is fed as zeros (Q9)
    ns["TopDownPerson"].rows.append({**key, "keypoints": kp})
    ns["VideoInfo"].rows.append({**key, "height": 1080, "width": 1920})
    out = V.process_videopose3d(key)
    ref = OV.process_videopose3d(kp, 1080, 1920, OV.load_lifter(synthetic_videopose3d_state_dict(0)))
    assert out["keypoints_3d"].shape == (150, 17, 3) and out["keypoints_3d"].dtype == np.float64
    assert out["keypoints_valid"] == [True] * 150
    assert np.abs(out["keypoints_3d"] - ref["keypoints_3d"]).max() <= 1e-3


# ------------------------------------------------------------------ handle lifetime / process exit
_EXIT_SCRIPT = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
os.environ["PE_SYNTHETIC_WEIGHTS"] = "1"; os.environ["PE_MAX_CROPS"] = "2"
import numpy as np, torch
torch.zeros(1, device="cuda")                        # a populate() worker normally has torch's CUDA context alive too
import fakes
fakes.make_fake_pose_pipeline()
from posepipeline_b200.wrappers import mmpose as W, videopose3d as V
from posepipeline_b200 import engine as E
m = W.get_model("HRNet_W48_COCO"); l = V.get_lifter()
from posepipeline_b200.synthetic import synthetic_frames, synthetic_bboxes
W.get_engine().stage_frames(synthetic_frames(1, 3))
kp = m.topdown([0], synthetic_bboxes(1, 5))
assert kp.shape == (1, 17, 3)
mode = {mode!r}
if mode == "engine_first":                           # release order a garbage collector may pick
    W.get_engine().close(); m.close(); l.close()
    try:
        m.topdown([0], synthetic_bboxes(1, 5)); raise SystemExit(3)
    except E._lib.PoseEngineError:
        pass
elif mode == "second_engine":                        # handles of two engines, nothing closed
    e2 = E.PoseEngine(0); l2 = E.Lifter(e2, __import__("posepipeline_b200.weights", fromlist=["x"]).synthetic_videopose3d_state_dict(0))
print("ok", flush=True)
"""                                                  # no close(): module globals keep models / lifter / engine alive


@pytest.mark.parametrize("mode", ["leak", "engine_first", "second_engine"])
def test_worker_process_exits_cleanly(mode):
    """A process that used the drop-in wrappers and never closed anything must exit 0 (round 1: rc 139 at interpreter
    exit), as a populate() worker does (reference utils/standard_pipelines.py:100)."""
    import subprocess
    import sys
    from conftest import ROOT
    r = subprocess.run([sys.executable, "-X", "faulthandler", "-c", _EXIT_SCRIPT.format(root=ROOT, mode=mode)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, (r.returncode, r.stdout[-500:], r.stderr[-2000:])
