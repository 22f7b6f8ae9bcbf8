"""-m gpu: the detector + tracker front end (rows a2 / f1) through the C ABI against the oracle restatement of the
mmdet / mmtrack arithmetic (oracle/yolox.py, oracle/bytetrack.py) -- PARITY UNPINNED upstream, see those files."""
import os

import cv2
import numpy as np
import pytest
import torch

import fakes
import helpers
from oracle import bytetrack as OB
from oracle import yolox as OY
from posepipeline_b200 import detector as D
from posepipeline_b200 import engine as E
from posepipeline_b200.synthetic import synthetic_frame

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = E.PoseEngine(0)
    yield e
    e.close()


@pytest.fixture(scope="module")
def sd():
    return D.synthetic_yolox_state_dict()


@pytest.fixture(scope="module")
def oracle_net(sd):
    return OY.load_detector(sd)


@pytest.fixture(scope="module")
def det1080(eng, sd):
    d = D.Detector(eng, sd, 1080, 1920, max_frames=2)
    yield d
    d.close()


def _match(ours, ref):
    """greedy one-to-one matching by IoU -> list of (i_ours, i_ref, iou)"""
    pairs, used = [], set()
    for j, r in enumerate(ref):
        best, bi = 0.0, -1
        for i, o in enumerate(ours):
            if i in used:
                continue
            iw = min(o[2], r[2]) - max(o[0], r[0])
            ih = min(o[3], r[3]) - max(o[1], r[1])
            if iw <= 0 or ih <= 0:
                continue
            iou = iw * ih / ((o[2] - o[0]) * (o[3] - o[1]) + (r[2] - r[0]) * (r[3] - r[1]) - iw * ih)
            if iou > best:
                best, bi = iou, i
        if bi >= 0:
            used.add(bi)
            pairs.append((bi, j, best))
    return pairs


def test_detector_input_and_every_layer_match_oracle(eng, sd, oracle_net):
    """Fixed-point resize + pad + Focus bit-exact; every ConvModule output of backbone / neck / head towers vs torch fp32."""
    frames = np.stack([synthetic_frame(0), synthetic_frame(1)])
    det1080 = D.Detector(eng, sd, 1080, 1920, max_frames=2, unique_slots=True)
    det1080.detect(frames)
    for img in (0, 1):
        x, sf = OY.preprocess(cv2.cvtColor(frames[img], cv2.COLOR_BGR2RGB))
        xt = torch.from_numpy(x)[None]
        focus = torch.cat((xt[..., ::2, ::2], xt[..., 1::2, ::2], xt[..., ::2, 1::2], xt[..., 1::2, 1::2]), dim=1)[0].numpy()
        got = det1080.debug_tensor("__input__", img)
        assert np.array_equal(got[:12], focus), (got[:12] != focus).mean()          # bit-exact (cv2.resize restated in fixed point)
        assert np.all(got[12:] == 0)
        acts = {}
        hooks = [m.register_forward_hook(lambda m, i, o, n=n: acts.__setitem__(n, o)) for n, m in oracle_net.named_modules()
                 if isinstance(m, (OY.ConvModule, OY.DarknetBottleneck))]
        oracle_net(xt)
        for h in hooks:
            h.remove()
        errs = []
        for name in acts:
            if name not in det1080.program.probes:
                continue
            # the engine fuses the identity add of a Darknet block into its conv2: compare with the block's output
            r = acts[name[: -len(".conv2")] if ".blocks." in name and name.endswith(".conv2") else name][0].numpy()
            g = det1080.debug_tensor(name, img)
            errs.append((float(np.abs(g - r).max() / (np.abs(r).max() + 1e-20)), name))
        # the in-place last Darknet block of every CSP layer overwrites main_conv's slice: those probes hold later values
        errs = [e for e in errs if not e[1].endswith("main_conv")]
        order = {n: i for i, n in enumerate(det1080.program.probes)}
        inorder = sorted(errs, key=lambda e: order[e[1]])
        print("detector layers in program order:", [(f"{e:.1e}", n.replace("backbone.", "b.")) for e, n in inorder[:40]])
        errs.sort(reverse=True)
        print(f"detector worst layers img {img} (max-abs-err / max-abs):", errs[:4], "median", errs[len(errs) // 2][0])
        assert len(errs) > 120
        assert errs[0][0] < 5e-5, errs[:5]
    det1080.close()


def test_detections_match_oracle(eng, det1080, oracle_net, sd):
    frames = np.stack([synthetic_frame(2), synthetic_frame(3)])
    got = det1080.detect(frames)
    net64 = OY.load_detector(sd, torch.float64)
    for img in (0, 1):
        ref = OY.detect(oracle_net, frames[img])
        ref64 = OY.detect(net64, frames[img])
        g = got[img]
        pairs = _match(g, ref)
        p64 = _match(ref, ref64)
        self_err = max(np.abs(ref[i, :4] - ref64[j, :4]).max() for i, j, _ in p64)
        d = np.array([np.abs(g[i, :4] - ref[j, :4]).max() for i, j, _ in pairs])
        ds = np.array([abs(g[i, 4] - ref[j, 4]) for i, j, _ in pairs])
        print(f"detections img {img}: ours {len(g)} oracle {len(ref)} matched {len(pairs)}; box |d| max {d.max():.2e} px (oracle fp32-vs-fp64 {self_err:.2e}), score |d| max {ds.max():.2e}")
        # candidates within 1e-4 of score_thr may fall on either side
        margin = np.sum(np.abs(ref[:, 4] - OY.SCORE_THR) < 1e-4) + np.sum(np.abs(g[:, 4] - OY.SCORE_THR) < 1e-4)
        assert abs(len(g) - len(ref)) <= margin and len(pairs) >= len(ref) - margin
        assert np.all(np.diff(g[:, 4]) <= 0)
        assert d.max() <= max(2e-2, 10 * self_err) and ds.max() <= 1e-4


def test_all_priors_decode_matches_oracle(eng, sd, oracle_net):
    """score_thr = 0: every one of the 23 625 priors becomes a candidate -> the head kernel's decode of all of them."""
    det = D.Detector(eng, sd, 1080, 1920, max_frames=1, score_thr=0.0, nms_iou=1.0, max_candidates=24000, max_det=24000)
    frame = synthetic_frame(4)
    g = det.detect(frame[None])[0]
    det.close()
    img, sf = OY.preprocess(cv2.cvtColor(frame, cv2.COLOR_BGR2RGB))
    cls, reg, obj = oracle_net(torch.from_numpy(img)[None])
    boxes, scores = OY.decode(cls, reg, obj, sf)
    order = np.argsort(-scores, kind="stable")
    assert len(g) == len(scores) == det.program.num_priors          # IoU threshold 1.0 suppresses nothing
    # same order up to score ties / rounding: compare as score-sorted sets
    gs, rs = g[:, 4], scores[order]
    assert np.abs(gs - rs).max() <= 1e-4
    top = 2000                                                       # the well-separated head of the list: rows align
    db = np.abs(g[:top, :4] - boxes[order][:top])
    same_row = db.max(1) < 1.0
    print(f"all-priors decode: score |d| max {np.abs(gs - rs).max():.2e}; top-{top} rows aligned {same_row.mean():.3f}, box |d| max {db[same_row].max():.2e} px")
    assert same_row.mean() > 0.95 and db[same_row].max() <= 2e-2          # near-equal scores may swap rows


def test_mmtrack_bounding_boxes_on_video(tmp_path, monkeypatch, eng, sd, oracle_net):
    """The reference-facing call on a real video file: detector + ByteTrack vs the oracle pipeline on the frames as decoded."""
    monkeypatch.setenv("PE_SYNTHETIC_WEIGHTS", "1")
    from posepipeline_b200.wrappers import mmtrack as W
    fakes.make_fake_pose_pipeline()
    base = synthetic_frame(7, 720, 1280)
    frames = [np.ascontiguousarray(np.roll(base, (3 * i, 5 * i), axis=(0, 1))) for i in range(10)]
    path = str(tmp_path / "clip.mp4")
    fakes.write_video(path, frames)
    decoded = fakes.read_video(path)
    monkeypatch.setattr(W, "_detector", D.DetectorPool(eng, sd, max_frames=4))
    tracks = W.mmtrack_bounding_boxes(path, "bytetrack")
    assert len(tracks) == len(decoded) == 10
    with pytest.raises(Exception, match="Unknown config file"):
        W.mmtrack_bounding_boxes(path, "sort")
    ref_tracker = OB.ByteTracker()
    n_ids, mismatched = set(), 0
    for f, frame in enumerate(decoded):
        rows = ref_tracker.update(f, OY.detect(oracle_net, frame))
        ours = tracks[f]
        for t in ours:
            assert set(t) == {"track_id", "tlbr", "tlhw", "confidence"} and isinstance(t["track_id"], int)
            assert np.allclose(t["tlhw"], [t["tlbr"][0], t["tlbr"][1], t["tlbr"][2] - t["tlbr"][0], t["tlbr"][3] - t["tlbr"][1]])
        ids_o, ids_r = [t["track_id"] for t in ours], rows[:, 0].astype(int).tolist()
        if ids_o != ids_r:
            mismatched += 1
            continue
        n_ids.update(ids_o)
        if len(ours):
            d = np.abs(np.stack([t["tlbr"] for t in ours]) - rows[:, 1:5]).max()
            assert d <= 5e-2, (f, d)
    print(f"bytetrack on video: {len(n_ids)} track ids, frames with differing id lists: {mismatched}/10")
    assert mismatched == 0 and len(n_ids) >= 1
    W._detector.close()


def test_block_reader_detections_equal_host_frame_detections(tmp_path, eng, sd):
    """Many blocks through the frame source (upload slots / the resident frame cache: the staged-frames pointer changes from
    block to block) against the same detector on host frames, bit for bit, from the very first call on.  Regression: the
    detector's input kernel used to be inside the captured CUDA graph with the staged-frames pointer baked in, so from the third
    forward of a batch size on every block was computed on the frames staged at capture time (found by the 2-GPU
    sharded-vs-single check of tools/bench_pipeline.py)."""
    from posepipeline_b200 import frames as F
    base = synthetic_frame(9, 360, 640)
    frames = [np.ascontiguousarray(np.roll(base, (2 * i, 7 * i), axis=(0, 1))) for i in range(44)]
    path = str(tmp_path / "blocks.mp4")
    fakes.write_video(path, frames)
    decoded = np.stack(fakes.read_video(path))
    assert len(decoded) == 44
    for use_cache in (False, True):
        pool = D.DetectorPool(eng, sd, max_frames=4)
        try:
            writer = F.CACHE.begin("test-blocks-%d" % use_cache, len(decoded), 360, 640, eng.device, first=0) if use_cache else None
            reader = F.BlockReader(path, eng, 8, 0, len(decoded), cache_writer=writer)
            got = []
            try:
                for blk in reader:
                    got.extend(pool.detect_block(reader, blk))
            finally:
                reader.close()
            ref = [d for i in range(0, len(decoded), 4) for d in pool.detect(decoded[i:i + 4])]
            again = []
            reader = F.BlockReader(path, eng, 8, 0, len(decoded))
            try:
                for blk in reader:
                    again.extend(pool.detect_block(reader, blk))
            finally:
                reader.close()
        finally:
            pool.close()
        assert len(got) == len(ref) == len(again) == 44
        assert sum(len(d) for d in ref) > 0
        for f in range(44):
            assert got[f].shape == ref[f].shape and np.array_equal(got[f], ref[f]), (use_cache, f)
            assert np.array_equal(again[f], ref[f]), (use_cache, f)
    F.CACHE.clear()
