"""Frame source (row f2) host logic on the CPU: block reader with a decode thread, seek == sequential decode, short videos,
the robust reader's contract and its validation skip."""
import os
import shutil
import tempfile

import numpy as np
import pytest

import fakes
from posepipeline_b200 import frames as F


@pytest.fixture(scope="module")
def clip(tmp_path_factory):
    d = tmp_path_factory.mktemp("clip")
    rng = np.random.default_rng(0)
    base = rng.integers(0, 256, (96, 128, 3), dtype=np.uint8)
    frames = [np.ascontiguousarray(np.roll(base, (2 * i, 3 * i), axis=(0, 1))) for i in range(45)]
    path = str(d / "v.mp4")
    fakes.write_video(path, frames)
    return path, fakes.read_video(path)


def test_block_reader_equals_sequential_decode(clip):
    path, decoded = clip
    r = F.BlockReader(path, None, block=8, start=0, stop=len(decoded))
    got, firsts = [], []
    for blk in r:
        assert blk.complete and not blk.on_device
        firsts.append(blk.first)
        got.extend(np.array(blk.frames))                       # copy: the pinned buffer is recycled
    r.close()
    assert firsts == [0, 8, 16, 24, 32, 40] and len(got) == 45
    assert all(np.array_equal(a, b) for a, b in zip(got, decoded))


@pytest.mark.parametrize("start,stop", [(1, 9), (7, 30), (23, 45), (44, 45), (0, 45)])
def test_seek_equals_sequential_decode(clip, start, stop, monkeypatch):
    """A sharded rank seeks to its first frame: the frames must be the ones a decode from frame 0 yields."""
    path, decoded = clip
    for seek in ("1", "0"):                                      # CAP_PROP_POS_FRAMES, and the grab() fallback
        monkeypatch.setenv("PE_FRAME_SEEK", seek)
        r = F.BlockReader(path, None, block=5, start=start, stop=stop)
        got = [np.array(f) for blk in r for f in blk.frames]
        r.close()
        assert len(got) == stop - start
        assert all(np.array_equal(a, b) for a, b in zip(got, decoded[start:stop])), (seek, start, stop)


def test_short_video_reports_incomplete_block(clip):
    path, decoded = clip
    r = F.BlockReader(path, None, block=16, start=32, stop=60)
    blocks = list(r)
    r.close()
    assert [b.n for b in blocks] == [13] and blocks[-1].complete is False
    r = F.BlockReader(path, None, block=15, start=30, stop=60)     # ends exactly on a block boundary: an empty, incomplete block
    blocks = list(r)
    r.close()
    assert [b.n for b in blocks] == [15, 0] and blocks[-1].complete is False


def test_reader_propagates_decoder_errors(tmp_path):
    r = F.BlockReader(str(tmp_path / "missing.mp4"), None, block=4, start=0, stop=3)
    blocks = list(r)
    r.close()
    assert len(blocks) == 1 and blocks[0].n == 0 and not blocks[0].complete


def test_fingerprint_is_content_based(clip, tmp_path):
    path, _ = clip
    cp = str(tmp_path / "other_name.mp4")
    shutil.copy(path, cp)
    assert F.fingerprint(path) == F.fingerprint(cp)
    with open(cp, "ab") as f:
        f.write(b"x")
    assert F.fingerprint(path) != F.fingerprint(cp)


def test_robust_reader_contract_and_validation_skip(clip, tmp_path, monkeypatch):
    """Same contract as Video.get_robust_reader (pipeline.py:47-87): a fresh temp copy the caller deletes; content already
    decoded completely in this process is not test-decoded again."""
    path, decoded = clip
    ns = fakes.make_fake_pose_pipeline()
    key = {"video_project": "t", "filename": "v"}

    class Video(fakes.Table):
        rows = []

    calls = {"reads": 0}
    import cv2
    real = cv2.VideoCapture

    class Counting:
        def __init__(self, *a):
            self.c = real(*a)

        def read(self):
            calls["reads"] += 1
            return self.c.read()

        def __getattr__(self, k):
            return getattr(self.c, k)
    monkeypatch.setattr(F.cv2, "VideoCapture", Counting)
    F._validated.clear()

    def fetch_copy():
        cp = str(tmp_path / f"fetched_{len(os.listdir(tmp_path))}.mp4")      # DataJoint downloads a fresh copy per fetch
        shutil.copy(path, cp)
        Video.rows = [{**key, "video": cp}]
    fetch_copy()
    out = F.robust_reader(Video, key, return_cap=False)
    assert out.startswith(tempfile.gettempdir()) and out.endswith(".mp4") and F.fingerprint(out) == F.fingerprint(path)
    assert calls["reads"] == len(decoded)                                     # first time: every frame test-decoded
    os.remove(out)
    fetch_copy()
    out = F.robust_reader(Video, key, return_cap=False)
    assert calls["reads"] == len(decoded)                                     # known-good content: no second validation decode
    os.remove(out)
    fetch_copy()
    cap = F.robust_reader(Video, key, return_cap=True)
    ok, frame = cap.read()
    assert ok and np.array_equal(frame, decoded[0])
    cap.release()
