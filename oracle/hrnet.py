"""Oracle: HRNet backbone + TopdownHeatmapSimpleHead in plain torch (CPU, fp32/fp64).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  PARITY UNPINNED (mmpose 0.x is
an un-vendored dependency); this restates upstream ``mmpose/models/backbones/
hrnet.py`` + ``heads/topdown_heatmap_simple_head.py`` as selected by the
reference config ``3rdparty/mmpose/config/top_down/darkpose/coco/
hrnet_w48_coco_384x288_dark.py:40-85`` and called at
``pose_pipeline/wrappers/mmpose.py:57,75``.  Module / parameter names follow
the upstream ``state_dict`` keys so a real mmpose checkpoint loads unchanged.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

BN_EPS = 1e-5

HRNET_CONFIGS = {
    # cfg :29-37 (channel_cfg), :44-72 (backbone.extra), :73-79 (head)
    "w48": dict(channels=(48, 96, 192, 384), num_modules=(1, 4, 3), num_joints=17),
    "w32": dict(channels=(32, 64, 128, 256), num_modules=(1, 4, 3), num_joints=17),
}


def _conv3x3(cin, cout, stride=1):
    return nn.Conv2d(cin, cout, 3, stride, 1, bias=False)


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, cin, planes):
        super().__init__()
        self.conv1 = _conv3x3(cin, planes)
        self.bn1 = nn.BatchNorm2d(planes, eps=BN_EPS)
        self.conv2 = _conv3x3(planes, planes)
        self.bn2 = nn.BatchNorm2d(planes, eps=BN_EPS)

    def forward(self, x):
        out = F.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        return F.relu(out + x)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, cin, planes, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes, eps=BN_EPS)
        self.conv2 = _conv3x3(planes, planes)
        self.bn2 = nn.BatchNorm2d(planes, eps=BN_EPS)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4, eps=BN_EPS)
        self.downsample = downsample

    def forward(self, x):
        identity = x if self.downsample is None else self.downsample(x)
        out = F.relu(self.bn1(self.conv1(x)))
        out = F.relu(self.bn2(self.conv2(out)))
        out = self.bn3(self.conv3(out))
        return F.relu(out + identity)


class HRModule(nn.Module):
    def __init__(self, channels: Sequence[int], num_blocks=4, multiscale_output=True):
        super().__init__()
        nb = len(channels)
        self.num_branches = nb
        self.branches = nn.ModuleList(
            [nn.Sequential(*[BasicBlock(c, c) for _ in range(num_blocks)]) for c in channels])
        fuse = []
        for i in range(nb if multiscale_output else 1):
            row = []
            for j in range(nb):
                if j > i:
                    row.append(nn.Sequential(
                        nn.Conv2d(channels[j], channels[i], 1, bias=False),
                        nn.BatchNorm2d(channels[i], eps=BN_EPS),
                        nn.Upsample(scale_factor=2 ** (j - i), mode="nearest")))
                elif j == i:
                    row.append(None)
                else:
                    chain = []
                    for k in range(i - j):
                        last = k == i - j - 1
                        cout = channels[i] if last else channels[j]
                        layers = [_conv3x3(channels[j], cout, 2), nn.BatchNorm2d(cout, eps=BN_EPS)]
                        if not last:
                            layers.append(nn.ReLU(inplace=False))
                        chain.append(nn.Sequential(*layers))
                    row.append(nn.Sequential(*chain))
            fuse.append(nn.ModuleList(row))
        self.fuse_layers = nn.ModuleList(fuse)

    def forward(self, xs: List[torch.Tensor]) -> List[torch.Tensor]:
        xs = [b(x) for b, x in zip(self.branches, xs)]
        if self.num_branches == 1:
            return xs
        out = []
        for i, row in enumerate(self.fuse_layers):
            y = 0
            for j in range(self.num_branches):
                y = y + (xs[j] if i == j else row[j](xs[j]))
            out.append(F.relu(y))
        return out


class HRNet(nn.Module):
    def __init__(self, channels=(48, 96, 192, 384), num_modules=(1, 4, 3)):
        super().__init__()
        self.conv1 = _conv3x3(3, 64, 2)
        self.bn1 = nn.BatchNorm2d(64, eps=BN_EPS)
        self.conv2 = _conv3x3(64, 64, 2)
        self.bn2 = nn.BatchNorm2d(64, eps=BN_EPS)
        ds = nn.Sequential(nn.Conv2d(64, 256, 1, bias=False), nn.BatchNorm2d(256, eps=BN_EPS))
        self.layer1 = nn.Sequential(Bottleneck(64, 64, ds), *[Bottleneck(256, 64) for _ in range(3)])
        c = channels
        self.transition1 = nn.ModuleList([
            nn.Sequential(_conv3x3(256, c[0]), nn.BatchNorm2d(c[0], eps=BN_EPS), nn.ReLU()),
            nn.Sequential(nn.Sequential(_conv3x3(256, c[1], 2), nn.BatchNorm2d(c[1], eps=BN_EPS), nn.ReLU())),
        ])
        self.stage2 = nn.Sequential(*[HRModule(c[:2]) for _ in range(num_modules[0])])
        self.transition2 = nn.ModuleList([
            None, None,
            nn.Sequential(nn.Sequential(_conv3x3(c[1], c[2], 2), nn.BatchNorm2d(c[2], eps=BN_EPS), nn.ReLU())),
        ])
        self.stage3 = nn.Sequential(*[HRModule(c[:3]) for _ in range(num_modules[1])])
        self.transition3 = nn.ModuleList([
            None, None, None,
            nn.Sequential(nn.Sequential(_conv3x3(c[2], c[3], 2), nn.BatchNorm2d(c[3], eps=BN_EPS), nn.ReLU())),
        ])
        self.stage4 = nn.Sequential(*[
            HRModule(c[:4], multiscale_output=(m != num_modules[2] - 1)) for m in range(num_modules[2])])

    def forward(self, x):
        x = F.relu(self.bn1(self.conv1(x)))
        x = F.relu(self.bn2(self.conv2(x)))
        x = self.layer1(x)
        xs = [t(x) for t in self.transition1]
        ys = self.stage2(xs)
        xs = [ys[i] if t is None else t(ys[-1]) for i, t in enumerate(self.transition2)]
        ys = self.stage3(xs)
        xs = [ys[i] if t is None else t(ys[-1]) for i, t in enumerate(self.transition3)]
        ys = self.stage4(xs)
        return ys[0]


class SimpleHead(nn.Module):
    """TopdownHeatmapSimpleHead(num_deconv_layers=0, final_conv_kernel=1): cfg :73-79."""

    def __init__(self, cin, num_joints):
        super().__init__()
        self.final_layer = nn.Conv2d(cin, num_joints, 1, bias=True)

    def forward(self, x):
        return self.final_layer(x)


class TopDownNet(nn.Module):
    """TopDown(backbone=HRNet, keypoint_head=SimpleHead) -- forward returns heatmaps (B,K,H/4,W/4)."""

    def __init__(self, variant="w48", num_joints=None):
        super().__init__()
        cfg = HRNET_CONFIGS[variant]
        self.variant = variant
        self.num_joints = num_joints or cfg["num_joints"]
        self.backbone = HRNet(cfg["channels"], cfg["num_modules"])
        self.keypoint_head = SimpleHead(cfg["channels"][0], self.num_joints)
        self.eval()

    @torch.no_grad()
    def forward(self, img):
        return self.keypoint_head(self.backbone(img))


def to_torch_state_dict(sd_numpy) -> Dict[str, torch.Tensor]:
    """numpy state_dict (as produced by the product's weight loader / synthetic generator) -> torch."""
    return {k: torch.from_numpy(np.asarray(v).copy()) for k, v in sd_numpy.items()}


def load_net(state_dict, variant="w48", dtype=torch.float32) -> TopDownNet:
    if not isinstance(next(iter(state_dict.values())), torch.Tensor):
        state_dict = to_torch_state_dict(state_dict)
    k = state_dict["keypoint_head.final_layer.weight"].shape[0]
    net = TopDownNet(variant, k)
    net.load_state_dict(state_dict, strict=True)
    return net.to(dtype).eval()


def count_macs(variant="w48", h=384, w=288):
    """Conv MACs / params / #convs of one forward pass (SURVEY App. B.4 cross-check)."""
    net = TopDownNet(variant)
    macs = [0]
    nconv = [0]

    def hook(m, inp, out):
        macs[0] += out.numel() // out.shape[0] * (m.in_channels // m.groups) * m.kernel_size[0] * m.kernel_size[1]
        nconv[0] += 1

    hs = [m.register_forward_hook(hook) for m in net.modules() if isinstance(m, nn.Conv2d)]
    net(torch.zeros(1, 3, h, w))
    for x in hs:
        x.remove()
    params = sum(p.numel() for p in net.parameters())
    return macs[0], params, nconv[0]
