"""mmcv config reading (VERDICT r1 missing 4): the restricted loader on the reference's own config files, and that the built-in
method table equals what those files say.  Needs /root/reference (build container only)."""
import dataclasses
import os
import shutil

import pytest

from posepipeline_b200 import engine as E
from posepipeline_b200 import mmcv_config as MC
from posepipeline_b200.tracking import BYTETRACK_CFG

REF3 = "/root/reference/3rdparty"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF3), reason="reference checkout not present")


@pytest.mark.parametrize("method", ["HRNet_W48_COCO", "HRNet_W48_COCOWholeBody", "HRNet_W48_HALPE"])
def test_builtin_method_table_equals_reference_configs(method):
    spec = E.METHODS[method]
    cfg = MC.load_config(os.path.join(REF3, spec.config))
    got = MC.topdown_settings(cfg)
    for k, v in got.items():
        ref = getattr(spec, k)
        assert (sorted(map(tuple, v)) == sorted(map(tuple, ref))) if k == "flip_pairs" else (tuple(v) == tuple(ref) if isinstance(v, tuple) else v == ref), (k, v, ref)
    assert E.spec_for(method, REF3) == dataclasses.replace(spec, **got)
    if method == "HRNet_W48_HALPE":
        assert len(got["flip_pairs"]) == 61 and "dataset_info" in cfg          # {{_base_.dataset_info}} resolved through _base_/halpe.py
    if method == "HRNet_W48_COCOWholeBody":
        assert "dataset_info" not in cfg and got["flip_pairs"] == MC.COCO_FALLBACK_PAIRS    # quirk Q3


def test_edited_config_is_honoured(tmp_path):
    rel = E.METHODS["HRNet_W48_COCO"].config
    dst = tmp_path / rel
    os.makedirs(dst.parent)
    text = open(os.path.join(REF3, rel)).read()
    text = text.replace("flip_test=True", "flip_test=False").replace("modulate_kernel=17", "modulate_kernel=11")
    dst.write_text(text)
    spec = E.spec_for("HRNet_W48_COCO", str(tmp_path))
    assert spec.flip_test is False and spec.modulate_kernel == 11 and spec.post_process == "unbiased"
    assert E.spec_for("HRNet_W48_COCO", str(tmp_path / "nowhere")) == E.METHODS["HRNet_W48_COCO"]


def test_loader_is_restricted(tmp_path):
    for body in ("import os\nx = 1\n", "x = __import__('os')\n", "x = (lambda: 1)()\n", "def f():\n    return 1\n"):
        p = tmp_path / "bad.py"
        p.write_text(body)
        with pytest.raises((ValueError, NameError)):
            MC.load_config(str(p))
    p = tmp_path / "ok.py"
    p.write_text("a = dict(b=[1, 2], c=dict(d=3))\ne = a['c']['d'] * 2\n")
    assert MC.load_config(str(p)) == {"a": {"b": [1, 2], "c": {"d": 3}}, "e": 6}


def test_bytetrack_config_matches_builtin_thresholds():
    cfg = MC.load_config(os.path.join(REF3, "mmtracking/mot/bytetrack/bytetrack_yolox_x_crowdhuman_mot17-private-half.py"))
    s = MC.bytetrack_settings(cfg)
    from posepipeline_b200 import detector as D
    assert s["img_scale"] == (800, 1440) and s["num_classes"] == 1
    assert s["score_thr"] == D.SCORE_THR and s["nms_iou"] == D.NMS_IOU
    assert s["tracker"] == BYTETRACK_CFG
    # _base_ merge: the detector architecture comes from _base_/models/yolox_x_8x8.py, the override only changes the head
    det = cfg["model"]["detector"]
    assert det["backbone"] == dict(type="CSPDarknet", deepen_factor=1.33, widen_factor=1.25) and det["neck"]["num_csp_blocks"] == 4
    assert det["bbox_head"]["in_channels"] == 320 and det["bbox_head"]["num_classes"] == 1
