"""-m gpu: f4 drop-ins -- person crops for the SMPL wrappers' dataloader (utils/bounding_box.py) through the engine's warp kernel,
bit-exact against the reference's cv2 calls (restated here; the reference module itself is checked against the same
transform on the CPU in tests/test_reference_pipeline.py)."""
import cv2
import numpy as np
import pytest

import fakes
from posepipeline_b200 import engine as E
from posepipeline_b200.utils import bounding_box as BB

pytestmark = pytest.mark.gpu


def _ref_crop(image, bbox, target_size, dilate):
    """pose_pipeline/utils/bounding_box.py:32-53 verbatim arithmetic."""
    bbox = BB.fix_bb_aspect_ratio(bbox, ratio=target_size[0] / target_size[1], dilate=dilate)
    src = np.asarray([[bbox[0], bbox[1]], [bbox[0] + bbox[2], bbox[1] + bbox[3]], [bbox[0], bbox[1] + bbox[3]]])
    dst = np.array([[0, 0], [target_size[0], target_size[1]], [0, target_size[1]]])
    trans = cv2.getAffineTransform(np.float32(src), np.float32(dst))
    return cv2.warpAffine(image, trans, target_size, flags=cv2.INTER_LINEAR), bbox


def test_crop_image_bbox_bit_exact():
    eng = E.PoseEngine(0)
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (720, 1280, 3), dtype=np.uint8)
    for bbox, ts, dil in [(np.array([300.5, 100.25, 180.0, 420.0]), (224, 224), 1.0), (np.array([-40.0, -30.0, 200.0, 300.0]), (224, 224), 1.2),
                          (np.array([1100.0, 500.0, 300.0, 400.0]), (288, 384), 1.2), (np.array([600.0, 300.0, 20.0, 30.0]), (224, 224), 1.0)]:
        got, b = BB.crop_image_bbox(img, bbox, ts, dil, engine=eng)
        ref, rb = _ref_crop(img, bbox, ts, dil)
        assert np.array_equal(b, rb) and got.shape == ref.shape
        assert np.array_equal(got, ref), (got != ref).mean()
    eng.close()


def test_person_crops_of_a_video(tmp_path):
    """get_person_dataloader's loop (reference :123-146): RGB crops of every present frame, absent frames skipped."""
    eng = E.PoseEngine(0)
    rng = np.random.default_rng(2)
    base = rng.integers(0, 256, (360, 640, 3), dtype=np.uint8)
    frames = [np.ascontiguousarray(np.roll(base, (i, 2 * i), axis=(0, 1))) for i in range(40)]
    path = str(tmp_path / "v.mp4")
    fakes.write_video(path, frames)
    decoded = fakes.read_video(path)
    bboxes = np.stack([np.array([100.0 + 3 * i, 50.0 + i, 120.0, 260.0]) for i in range(40)])
    present = np.ones(40, bool)
    present[[3, 17, 18]] = False
    ids, crops, boxes = BB.crop_video_person(path, bboxes, present, (224, 224), 1.0, engine=eng, block=16)
    assert ids == [i for i in range(40) if present[i]] and crops.shape == (37, 224, 224, 3) and boxes.shape == (37, 4)
    for k, i in enumerate(ids):
        ref, rb = _ref_crop(cv2.cvtColor(decoded[i], cv2.COLOR_BGR2RGB), bboxes[i], (224, 224), 1.0)
        assert np.array_equal(crops[k], ref), i
        assert np.array_equal(boxes[k], rb)
    eng.close()
