"""Oracle: ViTPose-B (ViT-B/16 backbone + 2-deconv heatmap head) as upstream ViTPose defines it.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  PARITY UNPINNED: ViTPose-B is BASELINE configs[2] / north_star's "HRNet/ViTPose
backbone" but it is NOT configured anywhere in the reference tree and upstream ViTPose is not
installable here.  Restated from the published ViTPose sources (``mmpose/models/backbones/vit.py``,
``configs/body/2d_kpt_sview_rgb_img/topdown_heatmap/coco/ViTPose_base_coco_256x192.py``) and mmpose's
``TopdownHeatmapSimpleHead``; parameter names follow that ``state_dict`` (``backbone.patch_embed.proj.weight``,
``backbone.blocks.N.attn.qkv.weight`` ..., ``keypoint_head.deconv_layers.N``, ``keypoint_head.final_layer``).
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


class Attention(nn.Module):
    def __init__(self, dim=768, heads=12):
        super().__init__()
        self.num_heads, self.scale = heads, (dim // heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        attn = ((q * self.scale) @ k.transpose(-2, -1)).softmax(dim=-1)
        return self.proj((attn @ v).transpose(1, 2).reshape(B, N, C))


class Mlp(nn.Module):
    def __init__(self, dim=768, hidden=3072):
        super().__init__()
        self.fc1, self.fc2 = nn.Linear(dim, hidden), nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(F.gelu(self.fc1(x)))


class Block(nn.Module):
    def __init__(self, dim=768, heads=12):
        super().__init__()
        self.norm1, self.attn = nn.LayerNorm(dim, eps=1e-6), Attention(dim, heads)
        self.norm2, self.mlp = nn.LayerNorm(dim, eps=1e-6), Mlp(dim, dim * 4)

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))


class PatchEmbed(nn.Module):
    def __init__(self, patch=16, dim=768, ratio=1):
        super().__init__()
        self.proj = nn.Conv2d(3, dim, kernel_size=patch, stride=patch // ratio, padding=4 + 2 * (ratio // 2 - 1))

    def forward(self, x):
        x = self.proj(x)
        Hp, Wp = x.shape[2], x.shape[3]
        return x.flatten(2).transpose(1, 2), (Hp, Wp)


class ViT(nn.Module):
    def __init__(self, img_size=(256, 192), dim=768, depth=12, heads=12):
        super().__init__()
        self.patch_embed = PatchEmbed(16, dim, 1)
        n = (img_size[0] // 16) * (img_size[1] // 16)
        self.pos_embed = nn.Parameter(torch.zeros(1, n + 1, dim))
        self.blocks = nn.ModuleList([Block(dim, heads) for _ in range(depth)])
        self.last_norm = nn.LayerNorm(dim, eps=1e-6)

    def forward(self, x):
        B = x.shape[0]
        x, (Hp, Wp) = self.patch_embed(x)
        x = x + self.pos_embed[:, 1:] + self.pos_embed[:, :1]
        for blk in self.blocks:
            x = blk(x)
        x = self.last_norm(x)
        return x.permute(0, 2, 1).reshape(B, -1, Hp, Wp).contiguous()


class SimpleHead(nn.Module):
    """TopdownHeatmapSimpleHead(in 768, num_deconv_layers 2, filters (256,256), kernels (4,4), final_conv_kernel 1)."""

    def __init__(self, cin=768, num_joints=17):
        super().__init__()
        layers, c = [], cin
        for _ in range(2):
            layers += [nn.ConvTranspose2d(c, 256, 4, stride=2, padding=1, output_padding=0, bias=False), nn.BatchNorm2d(256), nn.ReLU(inplace=True)]
            c = 256
        self.deconv_layers = nn.Sequential(*layers)
        self.final_layer = nn.Conv2d(256, num_joints, 1)

    def forward(self, x):
        return self.final_layer(self.deconv_layers(x))


class ViTPose(nn.Module):
    def __init__(self, num_joints=17):
        super().__init__()
        self.backbone, self.keypoint_head = ViT(), SimpleHead(768, num_joints)
        self.eval()

    @torch.no_grad()
    def forward(self, x):
        return self.keypoint_head(self.backbone(x))


def load_net(state_dict, dtype=torch.float32, num_joints=17):
    net = ViTPose(num_joints)
    sd = {k: (v if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v).copy())) for k, v in state_dict.items()}
    net.load_state_dict(sd, strict=True)
    return net.to(dtype).eval()
