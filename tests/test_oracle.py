"""CPU suite: the oracle against the reference-generated golden vectors, host logic, and the C-ABI surface."""
import ctypes as C
import json
import os
import re

import cv2
import numpy as np
import pytest
import torch

from oracle import hrnet as OH
from oracle import person_bbox as OPB
from oracle import topdown as OT
from oracle import videopose3d as OV
from oracle.warp_fixedpoint import warp_affine_fixedpoint
from posepipeline_b200 import _lib
from posepipeline_b200 import engine as E
from posepipeline_b200.hrnet_spec import build_program, conv_macs
from posepipeline_b200.synthetic import synthetic_bboxes, synthetic_keypoints_2d
from posepipeline_b200.weights import synthetic_videopose3d_state_dict

from conftest import HAS_GPU, ROOT
import helpers

GOLD = os.path.join(ROOT, "tests", "golden")


# ---------------------------------------------------------------- layout / boundary rules
def test_product_never_imports_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b", re.M)
    for d, _, files in os.walk(os.path.join(ROOT, "posepipeline_b200")):
        for f in files:
            if f.endswith(".py"):
                assert not pat.search(open(os.path.join(d, f)).read()), f"{f} imports the oracle"


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _lib.declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert lib.pe_abi_version() == 1


@pytest.mark.skipif(HAS_GPU, reason="only meaningful without a GPU")
def test_engine_fails_loudly_without_gpu():
    with pytest.raises(_lib.PoseEngineError) as ei:
        E.PoseEngine(0)
    assert ei.value.code == _lib.PE_ERR_NOGPU


# ---------------------------------------------------------------- a3 PersonBbox (pinned by reference code)
def _gold_bbox(c):
    return np.array([[float.fromhex(v) for v in r] for r in c["bbox_hex"]])


@pytest.mark.parametrize("impl", ["oracle", "cabi"])
def test_person_bbox_golden(impl):
    cases = json.load(open(os.path.join(GOLD, "person_bbox.json")))
    fn = OPB.person_bbox if impl == "oracle" else E.person_bbox
    for c in cases:
        bb, pr = fn(c["tracks"], c["keep_tracks"])
        gold = _gold_bbox(c)
        assert np.array_equal(pr, np.array(c["present"]))
        assert np.array_equal(np.isnan(bb), np.isnan(gold))
        assert np.array_equal(np.nan_to_num(bb), np.nan_to_num(gold))      # bit-exact


def test_person_bbox_edge_cases():
    tr = [[{"track_id": 1, "tlhw": [1.0, 2.0, 3.0, 4.0]}], [], [], [], [], [],
          [{"track_id": 1, "tlhw": [5.0, 6.0, 7.0, 8.0]}, {"track_id": 2, "tlhw": [0.0, 0.0, 1.0, 1.0]}]]
    for keep in ([1], [1, 2], [3]):
        a, pa = OPB.person_bbox(tr, keep)
        b, pb = E.person_bbox(tr, keep)
        assert np.array_equal(pa, pb) and np.array_equal(np.nan_to_num(a, nan=-7), np.nan_to_num(b, nan=-7))
    b, pb = E.person_bbox(tr, [1])
    assert pb.tolist() == [True, True, True, False, True, True, True]        # bfill 2 then ffill 2
    with pytest.raises(IndexError):
        E.person_bbox([], [1])


# ---------------------------------------------------------------- a9 DARK maths (pinned by utils/inference.py)
@pytest.mark.parametrize("tag", ["a", "b"])
def test_dark_decode_golden(tag):
    z = np.load(os.path.join(GOLD, "dark_decode.npz"))
    hm, k = z[f"{tag}_heatmaps"], int(z[f"{tag}_kernel"])
    preds, maxvals = OT.get_max_preds(hm)
    ref_arg = z[f"{tag}_argmax"]
    pos = z[f"{tag}_maxvals"][..., 0] > 0
    assert np.array_equal(preds[pos], ref_arg[pos].astype(np.float32))
    assert np.all(preds[~pos] == -1)                       # mmpose marks max<=0 as -1; DarkPose's copy uses 0
    assert np.array_equal(maxvals, z[f"{tag}_maxvals"])
    blurred = OT.gaussian_blur(hm.copy(), k)
    assert np.abs(blurred - z[f"{tag}_blurred"]).max() <= 2e-6 * np.abs(blurred).max()
    logged = np.log(np.maximum(blurred, 1e-10))
    ref = z[f"{tag}_refined"]
    for n in range(hm.shape[0]):
        for j in range(hm.shape[1]):
            if not pos[n, j]:
                continue
            c = OT.taylor(logged[n, j], preds[n, j].copy())
            assert np.abs(c - ref[n, j]).max() < 2e-3, (n, j, c, ref[n, j])


# ---------------------------------------------------------------- a5/a6 warp
def test_box_to_affine_bit_exact_vs_oracle():
    spec = E.METHODS["HRNet_W48_COCO"]
    for bb in synthetic_bboxes(300, 5):
        c, s, t = E.box_to_affine(spec, bb)
        oc, os_ = OT.box_to_center_scale(bb, OT.HRNET_W48_COCO)
        ot = OT.get_affine_transform(oc, os_, (288, 384))
        assert np.array_equal(c, oc) and np.array_equal(s, os_)
        assert np.abs(t - ot).max() <= 1e-12 * np.abs(ot).max()


def test_fixedpoint_warp_restatement_is_cv2():
    img = np.random.default_rng(0).integers(0, 256, (1080, 1920, 3), dtype=np.uint8)
    bbs = list(synthetic_bboxes(4, 3)) + [np.array([-200., -100., 500., 900.]), np.array([1700., 800., 400., 500.])]
    for bb in bbs:
        c, s = OT.box_to_center_scale(bb, OT.HRNET_W48_COCO)
        t = OT.get_affine_transform(c, s, (288, 384))
        ref = cv2.warpAffine(img, t, (288, 384), flags=cv2.INTER_LINEAR)
        assert np.array_equal(ref, warp_affine_fixedpoint(img, t, 288, 384))


# ---------------------------------------------------------------- a8 network description
def test_program_matches_oracle_state_dict_and_macs():
    for variant, h, w, macs in (("w48", 384, 288, 35306606592), ("w32", 256, 192, 7645003776)):
        prog = build_program(variant, h, w, 17)
        net = OH.TopDownNet(variant)
        assert list(net.state_dict().keys()) == list(prog.params.keys())
        for k, v in net.state_dict().items():
            assert tuple(v.shape) == tuple(prog.params[k]), k
        assert conv_macs(prog) == macs
    assert sum(1 for o in build_program("w48").ops if o.kind in (0, 1, 3)) == 293


def test_slot_assignment_never_aliases_live_tensors():
    prog = build_program("w48")
    slot_of, sizes = prog.assign_slots()
    for a in prog.tensors:
        for b in prog.tensors:
            if a.tid < b.tid and slot_of[a.tid] == slot_of[b.tid]:
                assert a.last_use < b.first_def or b.last_use < a.first_def, (a, b)
        assert sizes[slot_of[a.tid]] >= (a.H + 2) * (a.W + 2) * a.C


def test_flip_merge_restatement():
    rng = np.random.default_rng(1)
    hm, hf = rng.random((1, 17, 8, 6), dtype=np.float32), rng.random((1, 17, 8, 6), dtype=np.float32)
    m = OT.flip_test_heatmaps(hm, hf, OT.HRNET_W48_COCO)
    assert m[0, 1, 3, 0] == np.float32((hm[0, 1, 3, 0] + hf[0, 2, 3, 5]) * np.float32(0.5))     # col 0 keeps fb[0]
    assert m[0, 1, 3, 4] == np.float32((hm[0, 1, 3, 4] + hf[0, 2, 3, 6 - 1 - 3]) * np.float32(0.5))
    assert m[0, 0, 2, 2] == np.float32((hm[0, 0, 2, 2] + hf[0, 0, 2, 6 - 1 - 1]) * np.float32(0.5))


# ---------------------------------------------------------------- a10 lifter
def test_videopose3d_strided_equals_dilated_whole_sequence():
    """The restructuring the CUDA lifter relies on: per-window strided model == dilated model over the padded video."""
    sd = synthetic_videopose3d_state_dict()
    net = OV.load_lifter(sd, torch.float64)
    kp = synthetic_keypoints_2d(40)
    ref = OV.process_videopose3d(kp, 1080, 1920, net)["keypoints_3d"]
    x = OV.normalize_screen_coordinates(kp[:, :, :2], 1920, 1080)
    out = OV.dilated_whole_sequence(net, x)
    assert out.shape == ref.shape and np.abs(out - ref).max() < 2e-6      # wrapper rounds its windows to float32


def test_videopose3d_oracle_shapes_and_normalisation():
    kp = synthetic_keypoints_2d(5)
    x = OV.normalize_screen_coordinates(kp[:, :, :2], 1920, 1080)
    assert np.allclose(x[..., 0], kp[..., 0] / 1920 * 2 - 1) and np.allclose(x[..., 1], kp[..., 1] / 1920 * 2 - 1080 / 1920)
    w = OV.windows(x, 121)
    assert w.shape == (5, 243, 17, 2) and np.array_equal(w[0, 0], x[0]) and np.array_equal(w[4, -1], x[4])
    assert np.array_equal(w[2, 121], x[2])


# ------------------------------------------------------------------ host logic of the tensor-core planner (no GPU)
def test_conv_planner_candidates_respect_hardware_limits():
    """Every tiling the planner may hand to the auto-tuner must fit the SM: <= 227 KB dynamic shared memory, <= 512 TMEM
    columns, stage bases on the swizzle period, >= 2 pipeline stages, and the HRNet-W48 / W32 layer shapes must all have a plan."""
    import ctypes as C
    from posepipeline_b200 import _lib
    from posepipeline_b200.hrnet_spec import OP_CONV, build_program
    lib = _lib.load()
    fp16 = lib.pe_precision_mode() == 1
    chunk_bytes = 64 if fp16 else 128
    buf = (C.c_int32 * (64 * 14))()
    seen = set()
    for variant, (h, w) in (("w48", (384, 288)), ("w32", (256, 192))):
        prog = build_program(variant, h, w, 17)
        for op in prog.ops:
            if op.kind != OP_CONV or op.cin % 16 or op.cout % 16:
                continue
            to = prog.tensors[op.out]
            if op.stride == 2 and op.ksize != 3:
                continue
            key = (op.cin, op.cout, op.ksize, op.stride, to.H, to.W, op.residual >= 0)
            if key in seen:
                continue
            seen.add(key)
            cin, ks = (4 * op.cin, 2) if op.stride == 2 else (op.cin, op.ksize)
            for gather in ((1, 0) if op.stride == 2 else (0,)):
                n = lib.pe_tc_plan_candidates(cin, op.cout, ks, int(op.residual >= 0), to.H, to.W, 512, gather, buf, 64)
                if gather and (not fp16 or n == 0):
                    continue                                            # gather mode is optional (s2d copy is the fallback)
                assert n > 0 or not fp16, key                          # (tf32x3 build: 128-byte chunks, a few shapes fall back to the SIMT conv)
                for i in range(n):
                    ns, mt, nc, kc, S, nstg, stage, smem, tmem, rpg, ndrain, rows, cg, sets = buf[14 * i:14 * i + 14]
                    assert cg in (1, 2) and (cg == 1 or (nc % 16 == 0 and (nc // 2) % 8 == 0))
                    assert ns * nc == op.cout and nc % 16 == 0 and nc <= 128 and mt in (1, 2)
                    assert 2 <= S <= 4 and 1 <= nstg <= (8 if sets == 3 else 4)
                    assert sets in (1, 2, 3, 4) and smem <= 227 * 1024 and S * stage + {1: 12, 2: 16, 3: 4, 4: 8}[sets] * nstg * 32 * chunk_bytes <= smem
                    assert stage % (8 * chunk_bytes) == 0 and rows % 8 == 0
                    n_main = 3 if mt * nc // 16 <= 6 else 2
                    assert (n_main + 2) * mt * nc <= tmem <= 512
                    assert (cin // 16) % kc == 0 and rpg >= 1 and ndrain >= 1
                    steps = rpg * (ks if ks != 2 else 2) * (1 if fp16 else 2)       # hi*hi MMA steps per accumulation group
                    assert steps <= (8 if ks == 2 and gather else 6) or rpg == 1
    assert len(seen) >= 25
    assert lib.pe_tc_plan_candidates(40, 48, 3, 0, 8, 8, 1, 0, buf, 64) < 0      # channel counts must be multiples of 16
