"""Shared builders for the tests: synthetic weights, the oracle network, small inputs."""
import functools

import cv2
import numpy as np
import torch

from oracle import hrnet as OH
from oracle import topdown as OT
from posepipeline_b200.engine import METHODS
from posepipeline_b200.hrnet_spec import build_program
from posepipeline_b200.synthetic import synthetic_bboxes, synthetic_frames
from posepipeline_b200.weights import synthetic_hrnet_state_dict

ORACLE_CFG = {"HRNet_W48_COCO": OT.HRNET_W48_COCO, "HRNet_W32_COCO": OT.HRNET_W32_COCO}


@functools.lru_cache(maxsize=None)
def state_dict(method="HRNet_W48_COCO", seed=0):
    spec = METHODS[method]
    prog = build_program(spec.variant, spec.image_size[1], spec.image_size[0], spec.num_joints)
    return synthetic_hrnet_state_dict(prog, seed)


@functools.lru_cache(maxsize=None)
def oracle_net(method="HRNet_W48_COCO", seed=0, dtype="float32"):
    return OH.load_net(state_dict(method, seed), METHODS[method].variant, getattr(torch, dtype))


@functools.lru_cache(maxsize=None)
def frames(n=3, seed0=0):
    return synthetic_frames(n, seed0)


def oracle_keypoints(method, frames_bgr, frame_idx, bboxes, dtype="float32"):
    net = oracle_net(method, 0, dtype)
    cfg = ORACLE_CFG[method]
    out = []
    for fi, bb in zip(frame_idx, bboxes):
        f = cv2.cvtColor(frames_bgr[fi], cv2.COLOR_BGR2RGB)       # wrappers/mmpose.py:73
        out.append(OT.inference_top_down(net, f, bb, cfg))
    return np.asarray(out)
