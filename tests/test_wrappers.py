"""Host logic of the drop-in wrappers, on CPU: a stub model stands in for the GPU engine (the arithmetic is covered by the
-m gpu parity suite); what is tested here is the reference wrapper behaviour and the frame sharding."""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import fakes
from posepipeline_b200 import sharding


class StubEngine:
    def stage_frames(self, frames):
        self.frames = np.array(frames)


class StubModel:
    """keypoints = f(frame mean, bbox): deterministic, so sharded and unsharded runs must agree exactly."""

    def __init__(self):
        self.engine = StubEngine()
        self.calls = 0

    def topdown(self, frame_idx, bboxes):
        self.calls += 1
        out = np.zeros((len(frame_idx), 17, 3), np.float32)
        for r, (fi, bb) in enumerate(zip(frame_idx, bboxes)):
            out[r] = np.float32(self.engine.frames[fi].mean()) + np.float32(bb.sum()) + np.arange(51, dtype=np.float32).reshape(17, 3)
        return out


def _setup(tmp, n_frames=11, absent=(3, 4)):
    ns = fakes.make_fake_pose_pipeline()
    rng = np.random.default_rng(0)
    frames = [np.full((48, 64, 3), 10 * i, np.uint8) for i in range(n_frames)]
    path = os.path.join(tmp, "v.mp4")
    fakes.write_video(path, frames)
    key = {"video_project": "t", "filename": "v"}
    ns["Video"].rows.append({**key, "video": path})
    bbox = rng.uniform(1, 30, (n_frames, 4))
    for a in absent:
        bbox[a] = np.nan
    ns["PersonBbox"].rows.append({**key, "bbox": bbox, "present": ~np.isnan(bbox).any(1)})
    return ns, key, bbox, path


def test_top_down_wrapper_reference_behaviour(tmp_path, monkeypatch):
    from posepipeline_b200.wrappers import mmpose as W
    ns, key, bbox, path = _setup(str(tmp_path))
    stub = StubModel()
    monkeypatch.setattr(W, "get_model", lambda method: stub)
    before = set(os.listdir(tempfile.gettempdir()))
    out = W.mmpose_top_down_person(key, "HRNet_W48_COCO")
    assert out.shape == (11, 17, 3)
    assert out.dtype == np.float64                         # Q7: absent frames make the array float64
    assert np.all(out[3] == 0) and np.all(out[4] == 0)     # NaN bbox -> zeros row (wrappers/mmpose.py:67-69)
    assert np.all(out[5] != 0)
    assert set(os.listdir(tempfile.gettempdir())) == before  # temporary video removed (:79)
    assert os.path.exists(path)
    # no absent frame -> float32, like np.asarray of float32 model rows
    ns, key, bbox, path = _setup(str(tmp_path), absent=())
    out32 = W.mmpose_top_down_person(key)
    assert out32.dtype == np.float32
    # bbox rows beyond the video length -> the reference's assert fires (:64)
    ns["PersonBbox"].rows[0]["bbox"] = np.concatenate([bbox, bbox])
    with pytest.raises(AssertionError):
        W.mmpose_top_down_person(key)
    assert set(os.listdir(tempfile.gettempdir())) == before
    with pytest.raises(NotImplementedError):
        W.mmpose_top_down_person(key, "HRFormer_COCO")
    assert W.mmpose_joint_dictionary["MMPose"][0] == "Nose" and len(W.mmpose_joint_dictionary["MMPoseHalpe"]) == 26


def test_mmtrack_interface():
    from posepipeline_b200.wrappers import mmtrack as T
    with pytest.raises(Exception, match="Unknown config file for MMTrack method nope"):
        T.mmtrack_bounding_boxes("x.mp4", "nope")
    with pytest.raises(NotImplementedError):          # Faster R-CNN trackers: the reference's own install or nothing
        T.mmtrack_bounding_boxes("x.mp4", "tracktor")
    rows = np.array([[3, 10, 20, 50, 80, 0.9]], np.float32)
    d = T.tracks_from_rows(rows)[0]
    assert d["track_id"] == 3 and np.allclose(d["tlhw"], [10, 20, 40, 60]) and np.allclose(d["tlbr"], [10, 20, 50, 80])   # Q2


def test_person_bbox_make_through_install(tmp_path):
    ns = fakes.make_fake_pose_pipeline()
    import posepipeline_b200.install as inst
    key = {"k": 1}
    tracks = [[{"track_id": 1, "tlhw": [1.0, 2.0, 3.0, 4.0]}], [], [{"track_id": 1, "tlhw": [2.0, 2.0, 3.0, 4.0]}]]
    ns["TrackingBbox"].rows.append({**key, "tracks": tracks})
    ns["PersonBboxValid"].rows.append({**key, "keep_tracks": [1]})
    inst.person_bbox_make(ns["PersonBbox"](), dict(key))
    row = ns["PersonBbox"].rows[0]
    assert row["present"].tolist() == [True, True, True] and row["bbox"][1].tolist() == [2.0, 2.0, 3.0, 4.0]


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 8, 9, 1000):
        for w in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def _worker(rank, world, port, tmp, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from posepipeline_b200.wrappers import mmpose as W
    ns, key, bbox, path = _setup(tmp)
    stub = StubModel()
    W.get_model = lambda method: stub
    out = W.mmpose_top_down_person(key)
    q.put((rank, out, stub.calls))
    dist.barrier()
    dist.destroy_process_group()


def test_frame_sharding_world2_gloo_matches_single_rank(tmp_path, monkeypatch):
    from posepipeline_b200.wrappers import mmpose as W
    ns, key, bbox, path = _setup(str(tmp_path))
    monkeypatch.setattr(W, "get_model", lambda method: StubModel())
    single = W.mmpose_top_down_person(key)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    procs = [ctx.Process(target=_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=400) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, out, calls in got:
        assert out.dtype == single.dtype and np.array_equal(out, single), rank     # every rank holds the full result
        assert calls >= 1


# ------------------------------------------------------------------ mmtrack wrapper: detector sharded by frame, replicated tracker
class StubDetector:
    """detections = f(frame content): one box whose position follows the bright square drawn into the frame."""

    def detect(self, frames):
        out = []
        for f in frames:
            ys, xs = np.nonzero(f[..., 1] > 200)
            out.append(np.zeros((0, 5), np.float32) if len(xs) == 0 else
                       np.array([[xs.min(), ys.min(), xs.max() + 1, ys.max() + 1, 0.9], [1, 2, 9, 12, 0.3]], np.float32))
        return out


def _track_video(tmp, n=13):
    frames = []
    for i in range(n):
        f = np.full((96, 128, 3), 30, np.uint8)
        if i != 6:
            f[10 + i:50 + i, 20 + 3 * i:45 + 3 * i] = 255
        frames.append(f)
    path = os.path.join(tmp, "t.mp4")
    fakes.write_video(path, frames)
    return path


def _track_worker(rank, world, port, tmp, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from posepipeline_b200.wrappers import mmtrack as T
    T.get_detector = lambda: StubDetector()
    tracks = T.mmtrack_bounding_boxes(os.path.join(tmp, "t.mp4"), "bytetrack")
    q.put((rank, [[(t["track_id"], t["tlbr"].tolist(), float(t["confidence"])) for t in fr] for fr in tracks]))
    dist.barrier()
    dist.destroy_process_group()


def test_mmtrack_sharded_detector_world2_matches_single_rank(tmp_path, monkeypatch):
    from posepipeline_b200.wrappers import mmtrack as T
    path = _track_video(str(tmp_path))
    monkeypatch.setattr(T, "get_detector", lambda: StubDetector())
    single = T.mmtrack_bounding_boxes(path, "bytetrack")
    assert len(single) == 13 and len(single[6]) <= 1 and single[0][0]["track_id"] == 0
    ref = [[(t["track_id"], t["tlbr"].tolist(), float(t["confidence"])) for t in fr] for fr in single]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    procs = [ctx.Process(target=_track_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=400) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, tr in got:
        assert tr == ref, rank                      # every rank holds the full, identical track list
