"""-m gpu: ViTPose-B 256x192 (BASELINE configs[2]) through the C ABI against the oracle (oracle/vitpose.py + the UDP glue of
oracle/topdown.py) -- upstream ViTPose is not in the reference tree: PARITY UNPINNED, see those files."""
import cv2
import numpy as np
import pytest
import torch

import helpers
from oracle import topdown as OT
from oracle import vitpose as OV
from posepipeline_b200 import engine as E
from posepipeline_b200.synthetic import synthetic_bboxes
from posepipeline_b200.vit_spec import OP_ATTN, OP_D2S, OP_GEMM, OP_LN, build_vitpose_program
from posepipeline_b200.weights import synthetic_vitpose_state_dict

pytestmark = pytest.mark.gpu
CFG = OT.VITPOSE_B_COCO
SPEC = E.METHODS["ViTPose_B_COCO"]


@pytest.fixture(scope="module")
def eng():
    e = E.PoseEngine(0)
    yield e
    e.close()


@pytest.fixture(scope="module")
def sd():
    return synthetic_vitpose_state_dict(build_vitpose_program())


@pytest.fixture(scope="module")
def model(eng, sd):
    m = E.TopDownModel(eng, sd, SPEC, max_crops=4)
    yield m
    m.close()


def test_udp_crop_bit_exact(eng, model):
    frames = helpers.frames(3)
    eng.stage_frames(frames)
    bbs = np.concatenate([synthetic_bboxes(6, 11), np.array([[-200., -100., 500., 900.], [1700., 800., 400., 500.], [5., 5., 30., 40.]])])
    fidx = np.arange(len(bbs)) % 3
    crops, c, s = model.warp_crops(fidx, bbs)
    for i, bb in enumerate(bbs):
        x, oc, os_, ref = OT.preprocess(cv2.cvtColor(frames[fidx[i]], cv2.COLOR_BGR2RGB), bb, CFG)
        assert np.array_equal(c[i], oc) and np.array_equal(s[i], os_)
        assert np.array_equal(crops[i], ref), f"crop {i}: {(crops[i] != ref).sum()} pixels differ"


def test_udp_decode_matches_oracle(eng, model):
    rng = np.random.default_rng(5)
    n, K, H, W = 5, 17, 64, 48
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    hm = np.zeros((n, K, H, W), np.float32)
    for i in range(n):
        for k in range(K):
            cx, cy, sg = rng.uniform(-1, W + 1), rng.uniform(-1, H + 1), rng.uniform(1.5, 3.0)
            hm[i, k] = rng.uniform(0.3, 1.0) * np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * sg * sg)) + np.abs(rng.normal(0, 0.003, (H, W)))
    perm = np.arange(K)
    for a, b in OT.COCO_FLIP_PAIRS:
        perm[a], perm[b] = b, a
    hf = hm[:, perm][..., ::-1].copy() + rng.normal(0, 0.002, hm.shape).astype(np.float32)
    bbs = synthetic_bboxes(n, 3)
    cs = [OT.box_to_center_scale(b, CFG) for b in bbs]
    c = np.stack([x[0] for x in cs]); s = np.stack([x[1] for x in cs])
    got = model.decode_heatmaps(hm, hf, c, s)
    ref = OT.decode(OT.flip_test_heatmaps(hm, hf, CFG), c, s, CFG)
    d = np.abs(got[..., :2] - ref[..., :2]).max()
    print(f"UDP decode: max |dx| {d:.2e} px, score diff {np.abs(got[..., 2] - ref[..., 2]).max():.2e}")
    assert np.array_equal(got[..., 2], ref[..., 2])
    assert d <= 1e-3


def _oracle_refs(net, prog, x):
    """reference value of every comparable op output of the program"""
    vals, hooks = {}, []
    mods = dict(net.named_modules())

    def out_hook(name):
        return mods[name].register_forward_hook(lambda m, i, o, name=name: vals.__setitem__(("out", name), o))

    def in_hook(name):
        return mods[name].register_forward_pre_hook(lambda m, i, name=name: vals.__setitem__(("in", name), i[0]))
    for name in mods:
        if name.endswith(("norm1", "norm2", "last_norm", "attn.qkv")) or (name.startswith("backbone.blocks.") and name.count(".") == 2):
            hooks.append(out_hook(name))
        if name.endswith(("attn.proj", "norm2", "mlp.fc2")) or name == "backbone.blocks.0":
            hooks.append(in_hook(name))
    for i in (2, 5):
        hooks.append(out_hook(f"keypoint_head.deconv_layers.{i}"))
    with torch.no_grad():
        hm = net(x)
    for h in hooks:
        h.remove()
    refs = {}
    d2s_i = 0
    for k, op in enumerate(prog.ops):
        if op.kind == OP_LN:
            v = vals[("out", op.conv)]
            refs[k] = v if v.dim() == 3 else None
            if op.conv.endswith("last_norm"):
                refs[k] = ("grid", v)
        elif op.kind == OP_GEMM:
            if op.conv.endswith("patch_embed.proj"):
                refs[k] = vals[("in", "backbone.blocks.0")]
            elif op.conv.endswith("attn.qkv"):
                refs[k] = vals[("out", op.conv)]
            elif op.conv.endswith("attn.proj"):
                refs[k] = vals[("in", op.conv.replace("attn.proj", "norm2"))]
            elif op.conv.endswith("mlp.fc1"):
                refs[k] = vals[("in", op.conv.replace("fc1", "fc2"))]
            elif op.conv.endswith("mlp.fc2"):
                refs[k] = vals[("out", op.conv.rsplit(".mlp", 1)[0])]
        elif op.kind == OP_ATTN:
            blk = prog.ops[k + 1].conv
            refs[k] = vals[("in", blk)]
        elif op.kind == OP_D2S:
            refs[k] = ("chw", vals[("out", f"keypoint_head.deconv_layers.{2 + 3 * d2s_i}")])
            d2s_i += 1
    return refs, hm


def test_every_layer_matches_oracle(eng, sd):
    m = E.TopDownModel(eng, sd, SPEC, max_crops=1, unique_slots=True)
    frames = helpers.frames(3)
    bb = synthetic_bboxes(1, 21)[0]
    x, c, s, crop = OT.preprocess(cv2.cvtColor(frames[1], cv2.COLOR_BGR2RGB), bb, CFG)
    hm, hmf = m.forward_heatmaps(crop[None])
    net = OV.load_net(sd)
    xt = torch.from_numpy(x)[None]
    refs, rh = _oracle_refs(net, m.program, torch.cat([xt, xt.flip(3)]))
    errs = []
    for k, r in refs.items():
        if r is None:
            continue
        op = m.program.ops[k]
        for img in (0, 1):
            got = m.debug_tensor(op.out, img)
            if isinstance(r, tuple) and r[0] == "grid":
                ref = r[1][img].numpy().T.reshape(got.shape)                 # (tokens, C) -> (C, 16, 12)
            elif isinstance(r, tuple):
                ref = r[1][img].numpy()
            else:
                ref = r[img].numpy()
            errs.append((float(np.abs(got - ref).max() / (np.abs(ref).max() + 1e-20)), f"{k}:{op.conv or op.kind}", img))
    rh = rh.numpy()
    e_hm = max(np.abs(hm[0] - rh[0]).max() / np.abs(rh[0]).max(), np.abs(hmf[0] - rh[1]).max() / np.abs(rh[1]).max())
    inorder = [(f"{e:.1e}", n) for e, n, i in errs if i == 0][:16]
    errs.sort(reverse=True)
    print("vitpose layers in program order:", inorder)
    print("vitpose worst layers (max-abs-err / max-abs):", errs[:5], "median", errs[len(errs) // 2][0], "heatmap:", e_hm)
    m.close()
    assert len(errs) >= 2 * (1 + 12 * 7 + 1 + 2)
    assert errs[0][0] < 5e-5, errs[:5]
    assert e_hm <= 1e-4


def test_topdown_end_to_end_keypoints(eng, model, sd):
    frames = helpers.frames(3)
    eng.stage_frames(frames)
    n = 10
    bbs = synthetic_bboxes(n, 77)
    fidx = np.arange(n) % 3
    got = model.topdown(fidx, bbs)
    assert got.shape == (n, 17, 3)
    ref = {}
    for dt in ("float32", "float64"):
        net = OV.load_net(sd, getattr(torch, dt))
        ref[dt] = np.asarray([OT.inference_top_down(net, cv2.cvtColor(frames[fi], cv2.COLOR_BGR2RGB), bb, CFG) for fi, bb in zip(fidx, bbs)])
    cond = np.abs(ref["float32"][..., :2] - ref["float64"][..., :2]).max(-1)
    good = cond <= 1e-4
    d = np.abs(got[..., :2] - ref["float32"][..., :2]).max(-1)
    print(f"vitpose keypoint |dx| px: UNCONDITIONAL max {d.max():.2e} | max over well-conditioned {d[good].max():.2e} well-conditioned {good.mean():.3f} "
          f"oracle fp32-vs-fp64 max {cond.max():.2e}")
    assert (ref["float32"][..., 2] > 0.05).all()                    # real peaks: no keypoint sits at the degenerate (-1, -1)
    assert good.mean() >= 0.9 and d[good].max() <= 1e-3
    assert np.all(d[~good] <= 10 * cond[~good] + 1e-3)
    assert np.abs(got[..., 2] - ref["float32"][..., 2]).max() <= 1e-4 * max(1.0, np.abs(ref["float32"][..., 2]).max())
    assert np.array_equal(got, model.topdown(fidx, bbs))
