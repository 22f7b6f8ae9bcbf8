"""ViTPose-B (BASELINE configs[2]) on the CPU: the program against the oracle network (names / shapes / MACs), the
deconvolution-as-convolution rewrite, and the UDP glue of the oracle (warp matrix, post_dark_udp) on known cases."""
import cv2
import numpy as np
import torch

from oracle import topdown as OT
from oracle import vitpose as OV
from posepipeline_b200 import engine as E
from posepipeline_b200.vit_spec import build_vitpose_program, vit_macs
from posepipeline_b200.weights import synthetic_vitpose_state_dict


def test_program_matches_oracle_network():
    prog = build_vitpose_program()
    ref = {k: tuple(v.shape) for k, v in OV.ViTPose().state_dict().items()}
    assert set(prog.params) == set(ref) and all(tuple(prog.params[k]) == ref[k] for k in ref)
    n = sum(int(np.prod(s)) for k, s in prog.params.items() if not k.endswith("num_batches_tracked"))
    assert abs(n / 89.99e6 - 1) < 0.01                            # ~90 M parameters
    assert prog.tokens == 192 and prog.grid == (16, 12)
    assert abs(vit_macs(prog) / 18.52e9 - 1) < 0.03, vit_macs(prog)      # ~18.52 GMAC per pass
    sd = synthetic_vitpose_state_dict(prog)
    OV.load_net(sd)                                               # strict load under the upstream key names


def test_deconv_as_conv_equals_conv_transpose():
    rng = np.random.default_rng(0)
    w = rng.standard_normal((8, 5, 4, 4)).astype(np.float32)
    x = torch.from_numpy(rng.standard_normal((2, 8, 6, 7)).astype(np.float32))
    ref = torch.nn.functional.conv_transpose2d(x.double(), torch.from_numpy(w).double(), stride=2, padding=1)
    w3 = E.deconv_as_conv_weights(w)
    par = torch.nn.functional.conv2d(x.double(), torch.from_numpy(w3).double(), padding=1)             # (2, 4*5, 6, 7)
    out = torch.zeros_like(ref)
    for py in range(2):
        for px in range(2):
            out[:, :, py::2, px::2] = par[:, (py * 2 + px) * 5:(py * 2 + px + 1) * 5]
    assert out.shape == ref.shape == (2, 5, 12, 14) and torch.allclose(out, ref, atol=1e-12)


def test_udp_glue_of_the_oracle():
    cfg = OT.VITPOSE_B_COCO
    c, s = OT.box_to_center_scale([700.3, 200.7, 310.2, 640.9], cfg)
    m = OT.udp_affine(c, s, cfg.image_size)
    assert m.dtype == np.float32 and m.shape == (2, 3) and m[0, 1] == 0 and m[1, 0] == 0
    # UDP maps the box corners onto pixel CENTRES 0 and size-1 (unit-length convention)
    tl = m @ np.array([c[0] - s[0] * 100, c[1] - s[1] * 100, 1.0])
    br = m @ np.array([c[0] + s[0] * 100, c[1] + s[1] * 100, 1.0])
    assert np.allclose(tl, [0, 0], atol=1e-3) and np.allclose(br, [191, 255], atol=1e-3)
    # post_dark_udp on a sampled Gaussian recovers its sub-pixel centre; transform_preds_udp inverts the warp
    H, W = 64, 48
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    cx, cy = 20.3, 30.6
    hm = np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * 2.0 ** 2)).astype(np.float32)[None, None]
    preds, maxvals = OT.keypoints_from_heatmaps(hm, c[None], s[None], cfg.post_process, cfg.modulate_kernel, use_udp=True)
    back = (np.array([[cx, cy]], np.float32) * (s * 200 / np.array([W - 1, H - 1], np.float32)) + c - s * 100)
    assert np.abs(preds[0] - back).max() < 0.05 * s[0] * 200 / (W - 1)          # within 1/20 heatmap pixel
