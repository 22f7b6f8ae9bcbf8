"""ByteTrack association (row a2): the C++ tracker behind the C ABI against the numpy/scipy oracle restatement of mmtrack's
ByteTracker, on synthetic detection sequences -- host-only, so this parity test runs without a GPU."""
import numpy as np
import pytest

from oracle import bytetrack as OB
from posepipeline_b200.tracking import ByteTracker


def synthetic_detections(seed, n_frames=240, n_people=5, width=1920, height=1080):
    """Per frame an (n,5) float32 array [x1,y1,x2,y2,score] sorted by score (descending, like NMS output): people walking
    with smooth motion, entering / leaving, occlusion gaps, score dips into the low band, jitter, and false positives."""
    rng = np.random.default_rng(seed)
    people = []
    for p in range(n_people):
        people.append(dict(x=rng.uniform(100, width - 300), y=rng.uniform(50, height - 500), w=rng.uniform(80, 220),
                           vx=rng.uniform(-9, 9), vy=rng.uniform(-3, 3), born=int(rng.integers(0, n_frames // 3)) if p else 0,
                           dies=int(rng.integers(2 * n_frames // 3, n_frames + 40)), base=rng.uniform(0.55, 0.95)))
    frames = []
    for f in range(n_frames):
        dets = []
        for p in people:
            if not (p["born"] <= f < p["dies"]):
                continue
            p["x"] += p["vx"] + rng.normal(0, 0.8)
            p["y"] += p["vy"] + rng.normal(0, 0.5)
            if rng.random() < 0.04:                              # missed detection
                continue
            h = p["w"] * 2.4
            score = float(np.clip(p["base"] + rng.normal(0, 0.12) - (0.45 if rng.random() < 0.1 else 0.0), 0.02, 0.99))
            jit = rng.normal(0, 1.5, 4)
            dets.append([p["x"] + jit[0], p["y"] + jit[1], p["x"] + p["w"] + jit[2], p["y"] + h + jit[3], score])
        for _ in range(rng.poisson(0.6)):                          # false positives, mostly low score
            x, y, w = rng.uniform(0, width - 100), rng.uniform(0, height - 200), rng.uniform(30, 200)
            dets.append([x, y, x + w, y + 2 * w, float(rng.beta(1.2, 4.0))])
        d = np.asarray(dets, np.float32).reshape(-1, 5)
        frames.append(d[np.argsort(-d[:, 4], kind="stable")])
    return frames


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4, 5])
def test_tracker_matches_oracle(seed):
    frames = synthetic_detections(seed, n_people=3 + seed)
    ours, ref = ByteTracker(), OB.ByteTracker()
    n_ids = set()
    for f, d in enumerate(frames):
        a = ours.update(f, d)
        b = ref.update(f, d)
        assert a.shape == b.shape, (f, a, b)
        assert np.array_equal(a[:, 0], b[:, 0]), (f, a[:, 0], b[:, 0])                 # track ids: bit-exact
        assert np.array_equal(a[:, 1:], b[:, 1:].astype(np.float64)), f                  # boxes / scores are the detections themselves
        assert a.dtype == np.float64
        n_ids.update(a[:, 0].astype(int).tolist())
    assert len(n_ids) >= 3
    ours.close()


def test_tracker_edge_cases():
    t, r = ByteTracker(), OB.ByteTracker()
    empty = np.zeros((0, 5), np.float32)
    for f, d in enumerate([empty, empty, np.array([[10, 10, 60, 160, 0.9]], np.float32), empty,
                           np.array([[12, 11, 62, 161, 0.5], [300, 300, 340, 400, 0.95]], np.float32),
                           np.array([[14, 12, 64, 162, 0.65]], np.float32)]):
        a, b = t.update(f, d), r.update(f, d)
        assert a.shape == b.shape and np.array_equal(a, b.astype(np.float64)), (f, a, b)
    # frame 0 resets ids (ByteTrack.simple_test) and frame-0 tracks are confirmed at once
    a = t.update(0, np.array([[10, 10, 60, 160, 0.9], [100, 10, 160, 160, 0.6]], np.float32))
    assert a[:, 0].tolist() == [0.0] and a.shape == (1, 6)                                # only score > init_track_thr starts a track
    a = t.update(1, np.array([[11, 10, 61, 160, 0.3]], np.float32))                       # low-score detection keeps the confirmed track
    assert a[:, 0].tolist() == [0.0]
    t.close()


def test_extended_assignment_equals_lap_semantics():
    """cost_limit: a pair is only matched when that is cheaper than leaving both unmatched (cost_limit in total)."""
    cost = np.array([[0.2, 0.95], [0.95, 0.97]])
    row, col = OB.lapjv_extended(cost, 0.9)
    assert row.tolist() == [0, -1] and col.tolist() == [0, -1]
