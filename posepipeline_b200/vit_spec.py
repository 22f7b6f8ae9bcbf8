"""ViTPose-B top-down network description (BASELINE configs[2]; north_star "HRNet/ViTPose backbone").

ViTPose is NOT configured in the reference tree; the description follows upstream ViTPose
(``ViTPose_base_coco_256x192.py``: ViT-B/16, PatchEmbed padding 2, 12 blocks of LN -> MHA(12 x 64) -> LN -> MLP(GELU),
``last_norm``, ``TopdownHeatmapSimpleHead`` with two 4x4 stride-2 deconvolutions + BN + ReLU and a 1x1 final conv; test_cfg
flip_test, use_udp, shift_heatmap False, modulate_kernel 11) and uses its ``state_dict`` key names.

Program form = hrnet_spec.Program with token tensors written as W = 0, H = tokens (flat [tokens][C] matrices):
  PATCH (crop -> 192 x 768 patch rows) -> GEMM patch_embed (+ position table through the residual path)
  12 x [ LN -> GEMM qkv -> ATTN -> GEMM proj (+x) -> LN -> GEMM fc1 (GELU) -> GEMM fc2 (+x) ]
  LN last_norm written as the padded 16 x 12 grid -> CONV 3x3 768 -> 4*256 (deconv 1 as four parities) + ReLU -> D2S
  -> CONV 3x3 256 -> 4*256 + ReLU -> D2S -> HEAD 1x1 256 -> K.
Every Linear and both deconvolutions run on the tcgen05 kernel (conv_tc.cu); LN / attention / depth-to-space are vit.cu.
"""
from __future__ import annotations

from .hrnet_spec import OP_CONV, OP_HEAD, Op, Program

OP_PATCH, OP_GEMM, OP_LN, OP_ATTN, OP_D2S = 4, 5, 6, 7, 8
ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 3


def build_vitpose_program(in_h: int = 256, in_w: int = 192, num_joints: int = 17, dim: int = 768, depth: int = 12, heads: int = 12,
                          patch: int = 16, pad: int = 2) -> Program:
    p = Program("vitpose_b", in_h, in_w, num_joints)
    th, tw = (in_h + 2 * pad - patch) // patch + 1, (in_w + 2 * pad - patch) // patch + 1
    T = th * tw
    p.tokens, p.grid, p.heads, p.dim = T, (th, tw), heads, dim
    bb = "backbone"

    def tok(C):
        return p._t(C, T, 0)

    def gemm(x, name, cin, cout, act=ACT_NONE, residual=-1, const=None):
        p.params[f"{name}.weight"] = (cout, cin) if not name.endswith("proj_conv") else (cout, 3, patch, patch)
        p.params[f"{name}.bias"] = (cout,)
        out = tok(cout)
        op = Op(OP_GEMM, out, [x], conv=name, cin=cin, cout=cout, relu=act, residual=residual, has_bias=True)
        op.const_table = const
        p.ops.append(op)
        return out

    def ln(x, name, out=None, grid2d=False):
        p.params[f"{name}.weight"] = (dim,)
        p.params[f"{name}.bias"] = (dim,)
        o = out if out is not None else tok(dim)
        op = Op(OP_LN, o, [x], ups=[int(grid2d), 6], conv=name, cout=dim)
        p.ops.append(op)
        return o

    # patch embedding: Conv2d(3, dim, k16, s16, p2) as a GEMM over gathered patch rows; x + pos[:, 1:] + pos[:, :1]
    p.params[f"{bb}.pos_embed"] = (1, T + 1, dim)
    K = 3 * patch * patch
    rows = tok(K)
    p.ops.append(Op(OP_PATCH, rows, [-1], ksize=patch, stride=pad, cin=3, cout=K))
    p.params[f"{bb}.patch_embed.proj.weight"] = (dim, 3, patch, patch)
    p.params[f"{bb}.patch_embed.proj.bias"] = (dim,)
    x = tok(dim)
    op = Op(OP_GEMM, x, [rows], conv=f"{bb}.patch_embed.proj", cin=K, cout=dim, relu=ACT_NONE, has_bias=True)
    op.const_table = f"{bb}.pos_embed"
    p.ops.append(op)
    for b in range(depth):
        q = f"{bb}.blocks.{b}"
        h = ln(x, f"{q}.norm1")
        qkv = gemm(h, f"{q}.attn.qkv", dim, 3 * dim)
        a = tok(dim)
        p.ops.append(Op(OP_ATTN, a, [qkv], cin=heads, cout=dim))
        x = gemm(a, f"{q}.attn.proj", dim, dim, residual=x)
        h = ln(x, f"{q}.norm2")
        f = gemm(h, f"{q}.mlp.fc1", dim, 4 * dim, act=ACT_GELU)
        x = gemm(f, f"{q}.mlp.fc2", 4 * dim, dim, residual=x)
    grid = p._t(dim, th, tw)
    ln(x, f"{bb}.last_norm", out=grid, grid2d=True)
    # head: 2 x [ConvTranspose2d(k4, s2, p1, no bias) + BN + ReLU] as 3x3 convolutions producing the 4 output parities
    hd = "keypoint_head"
    cur, cin, H, W = grid, dim, th, tw
    for i in range(2):
        p.params[f"{hd}.deconv_layers.{3 * i}.weight"] = (cin, 256, 4, 4)
        p._bn_params(f"{hd}.deconv_layers.{3 * i + 1}", 256)
        par = p._t(4 * 256, H, W)
        op = Op(OP_CONV, par, [cur], conv=f"{hd}.deconv_layers.{3 * i}", bn=f"{hd}.deconv_layers.{3 * i + 1}", ksize=3, stride=1, cin=cin,
                cout=4 * 256, relu=True)
        op.deconv = True
        p.ops.append(op)
        up = p._t(256, 2 * H, 2 * W)
        p.ops.append(Op(OP_D2S, up, [par], cout=256))
        cur, cin, H, W = up, 256, 2 * H, 2 * W
    p.params[f"{hd}.final_layer.weight"] = (num_joints, 256, 1, 1)
    p.params[f"{hd}.final_layer.bias"] = (num_joints,)
    out = p._t(num_joints, H, W)
    p.ops.append(Op(OP_HEAD, out, [cur], conv=f"{hd}.final_layer", ksize=1, cin=256, cout=num_joints, has_bias=True))
    p.out_tensor = out
    p.finalize()
    return p


def vit_macs(p: Program) -> int:
    """Multiply-accumulates per forward pass: Linear layers, attention (QK^T and PV), deconvolutions (their real 4x4 taps), head."""
    total = 0
    T = p.tokens
    for op in p.ops:
        if op.kind == OP_GEMM:
            total += T * op.cin * op.cout
        elif op.kind == OP_ATTN:
            total += 2 * T * T * op.cout
        elif op.kind == OP_CONV:
            t = p.tensors[op.out]
            total += t.H * t.W * op.cin * (op.cout // 4) * 16      # ConvTranspose k4 s2: 4 taps per output pixel, 4 pixels per input position
        elif op.kind == OP_HEAD:
            t = p.tensors[op.out]
            total += t.H * t.W * op.cin * op.cout
    return total
