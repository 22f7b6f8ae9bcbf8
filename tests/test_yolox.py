"""Detector (row a2 / f1) on the CPU: the oracle's preprocessing against cv2, the program against the oracle network
(parameter names / shapes, MAC count), NMS semantics, and the host-side wrapper plumbing."""
import numpy as np
import pytest
import torch

from oracle import yolox as OY
from posepipeline_b200 import yolox_spec as YS


def test_resize_restatement_is_cv2():
    """cv2.resize(INTER_LINEAR) on uint8 is fixed-point; the CUDA input kernel follows this restatement bit for bit."""
    import cv2
    rng = np.random.default_rng(0)
    for (h, w) in [(1080, 1920), (720, 1280), (480, 640), (1440, 2560), (360, 202)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        nh, nw = OY.rescale_size(h, w)
        assert (nh, nw) == YS.rescale_size(h, w)
        assert np.array_equal(cv2.resize(img, (nw, nh), interpolation=cv2.INTER_LINEAR), OY.resize_linear_u8(img, nw, nh)), (h, w)
    assert OY.rescale_size(1080, 1920) == (800, 1422) and YS.net_size(1080, 1920) == (800, 1422, 800, 1440)


def test_program_matches_oracle_network():
    prog = YS.YoloxProgram(800, 1440)
    net = OY.YOLOX()
    ref = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    assert set(prog.params) == set(ref)
    assert all(tuple(prog.params[k]) == ref[k] for k in ref)
    n_params = sum(int(np.prod(s)) for k, s in prog.params.items() if not k.endswith("num_batches_tracked"))
    assert 98e6 < n_params < 100e6                                  # YOLOX-X with a 1-class head: ~99 M parameters
    # conv MACs: YOLOX-X is 281.9 GFLOPs (140.95 GMAC) at 640x640 with 80 classes; scale by area, minus the class branch
    macs = prog.conv_macs()
    assert abs(macs / (140.95e9 * 800 * 1440 / 640 / 640) - 1) < 0.02, macs
    assert prog.num_priors == 100 * 180 + 50 * 90 + 25 * 45
    # every oracle ConvModule output has a probe (tensor slice) in the program
    mods = [n for n, m in net.named_modules() if isinstance(m, OY.ConvModule)]
    assert set(mods) == set(prog.probes) - {"__input__"}
    # slot sharing never aliases two live tensors
    slot_of, _ = prog.assign_slots()
    for a in prog.tensors:
        for b in prog.tensors:
            if a.tid < b.tid and slot_of[a.tid] == slot_of[b.tid]:
                assert a.last_use < b.first_def or b.last_use < a.first_def, (a, b)


def test_nms_semantics():
    b = np.array([[0, 0, 10, 10], [1, 1, 11, 11], [20, 20, 30, 30], [0, 0, 10, 10.5]], np.float32)
    s = np.array([0.9, 0.95, 0.5, 0.3], np.float32)
    keep = OY.nms(b, s, 0.7)
    # score order 1, 0, 2, 3: IoU(1,0) = 81/119 = 0.68 <= 0.7 keeps box 0; IoU(0,3) = 100/105 suppresses box 3
    assert keep.tolist() == [1, 0, 2]
    assert OY.nms(b, s, 0.6).tolist() == [1, 2]


def test_synthetic_weights_give_sparse_detections_on_cpu_oracle():
    """Small frame (fast on CPU): the synthetic detector weights must yield a handful of candidates, not all priors."""
    from posepipeline_b200.detector import synthetic_yolox_state_dict
    from posepipeline_b200.synthetic import synthetic_frame
    sd = synthetic_yolox_state_dict()
    net = OY.load_detector(sd)
    frame = synthetic_frame(3, 270, 480)
    dets = OY.detect(net, frame)
    assert dets.shape[1] == 5 and 1 <= len(dets) <= 400, len(dets)
    assert np.all(np.diff(dets[:, 4]) <= 0)                                   # score-descending
    assert np.all(dets[:, 2] > dets[:, 0]) and np.all(dets[:, 3] > dets[:, 1])
