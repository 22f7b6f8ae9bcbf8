"""HRNet top-down network description: parameter names/shapes and the layer program.

This is host-side graph building for the engine behind ``include/poseengine.h``: it
enumerates, in execution order, every convolution of the network the reference selects with
``3rdparty/mmpose/config/top_down/darkpose/coco/hrnet_w48_coco_384x288_dark.py:44-79``
(HRNet backbone, stage1 BOTTLENECK x4, stages 2-4 BASIC x4 with 1/4/3 modules, 2/3/4
branches, head = one 1x1 conv), using the upstream mmpose ``state_dict`` key names
 so a real checkpoint maps onto the program unchanged.

The program is a flat list of ``Op`` records over symbolic activation tensors; the Python
engine wrapper folds BatchNorm into each conv, packs the weights and hands both to the
C-ABI (``pe_model_create``).  No torch module is instantiated here.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

OP_STEM = 0       # 3x3 s2 conv reading the uint8 crop through the normalisation LUT
OP_CONV = 1       # 3x3 / 1x1 conv, stride 1 / 2, folded BN, optional residual add, optional ReLU
OP_FUSE = 2       # out = ReLU(sum_j nearest_upsample(in_j, up_j))   (HRModule fuse, App. A.2)
OP_HEAD = 3       # 1x1 conv with bias -> planar (K,H,W) fp32 heatmaps

VARIANTS = {
    "w48": dict(channels=(48, 96, 192, 384), num_modules=(1, 4, 3)),
    "w32": dict(channels=(32, 64, 128, 256), num_modules=(1, 4, 3)),
}


@dataclass
class Tensor:
    tid: int
    C: int
    H: int
    W: int
    first_def: int = -1
    last_use: int = -1


@dataclass
class Op:
    kind: int
    out: int
    ins: List[int]
    ups: List[int] = field(default_factory=list)   # OP_FUSE: upsample factor per input
    conv: Optional[str] = None                     # state_dict prefix of the conv ("....conv1")
    bn: Optional[str] = None                       # state_dict prefix of the BN folded into it
    ksize: int = 1
    stride: int = 1
    cin: int = 0
    cout: int = 0
    relu: bool = False
    residual: int = -1                             # tensor id added before the ReLU
    has_bias: bool = False                         # conv has its own bias (head only)


class Program:
    def __init__(self, variant: str, in_h: int, in_w: int, num_joints: int):
        self.variant = variant
        self.in_h, self.in_w, self.num_joints = in_h, in_w, num_joints
        self.tensors: List[Tensor] = []
        self.ops: List[Op] = []
        self.params: Dict[str, Tuple[int, ...]] = {}   # name -> shape, in state_dict order
        self.out_tensor = -1

    # ---- helpers
    def _t(self, C, H, W) -> int:
        self.tensors.append(Tensor(len(self.tensors), C, H, W))
        return len(self.tensors) - 1

    def _bn_params(self, name, c):
        for leaf in ("weight", "bias", "running_mean", "running_var"):
            self.params[f"{name}.{leaf}"] = (c,)
        self.params[f"{name}.num_batches_tracked"] = ()

    def conv(self, x, conv, bn, cin, cout, k, stride=1, relu=True, residual=-1, kind=OP_CONV) -> int:
        tin = self.tensors[x] if x >= 0 else None
        H = (tin.H if tin else self.in_h)
        W = (tin.W if tin else self.in_w)
        pad = k // 2
        Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        self.params[f"{conv}.weight"] = (cout, cin, k, k)
        if bn is not None:
            self._bn_params(bn, cout)
        out = self._t(cout, Ho, Wo)
        self.ops.append(Op(kind, out, [x], conv=conv, bn=bn, ksize=k, stride=stride, cin=cin, cout=cout,
                           relu=relu, residual=residual))
        return out

    def fuse(self, terms: List[Tuple[int, int]]) -> int:
        t0 = self.tensors[terms[0][0]]
        up0 = terms[0][1]
        out = self._t(t0.C, t0.H * up0, t0.W * up0)
        self.ops.append(Op(OP_FUSE, out, [t for t, _ in terms], ups=[u for _, u in terms], cout=t0.C, relu=True))
        return out

    def finalize(self):
        for i, op in enumerate(self.ops):
            self.tensors[op.out].first_def = i
            for t in op.ins + ([op.residual] if op.residual >= 0 else []):
                if t >= 0:
                    self.tensors[t].last_use = i
        self.tensors[self.out_tensor].last_use = len(self.ops)

    # ---- greedy slot assignment: tensors whose live ranges do not overlap share a buffer
    def assign_slots(self):
        """Returns (slot_of_tensor, slot_elems_per_image) where elems counts padded pixels * C."""
        def elems(t):
            return ((t.H + 2) * (t.W + 2) if t.W else t.H) * t.C      # W == 0: flat token matrix of H rows
        slots: List[Tuple[int, int]] = []       # (free_after_op, size)
        slot_of = [-1] * len(self.tensors)
        order = sorted(self.tensors, key=lambda t: t.first_def)
        for t in order:
            best = -1
            for s, (free_after, size) in enumerate(slots):
                if free_after < t.first_def and (best < 0 or abs(size - elems(t)) < abs(slots[best][1] - elems(t))):
                    best = s
            if best < 0:
                slots.append((t.last_use, elems(t)))
                best = len(slots) - 1
            else:
                slots[best] = (t.last_use, max(slots[best][1], elems(t)))
            slot_of[t.tid] = best
        return slot_of, [s[1] for s in slots]


def build_program(variant: str = "w48", in_h: int = 384, in_w: int = 288, num_joints: int = 17) -> Program:
    cfg = VARIANTS[variant]
    C = cfg["channels"]
    p = Program(variant, in_h, in_w, num_joints)
    bb = "backbone"

    # stem: conv1 reads the uint8 crop (tensor id -1)
    x = p.conv(-1, f"{bb}.conv1", f"{bb}.bn1", 3, 64, 3, 2, kind=OP_STEM)
    x = p.conv(x, f"{bb}.conv2", f"{bb}.bn2", 64, 64, 3, 2)

    # layer1: 4 bottlenecks (first has the 1x1 downsample projection)
    for b in range(4):
        pre = f"{bb}.layer1.{b}"
        cin = 64 if b == 0 else 256
        o = p.conv(x, f"{pre}.conv1", f"{pre}.bn1", cin, 64, 1)
        o = p.conv(o, f"{pre}.conv2", f"{pre}.bn2", 64, 64, 3)
        if b == 0:
            # state_dict order: conv3, bn3 come before downsample
            p.params[f"{pre}.conv3.weight"] = (256, 64, 1, 1)
            p._bn_params(f"{pre}.bn3", 256)
            idn = p.conv(x, f"{pre}.downsample.0", f"{pre}.downsample.1", 64, 256, 1, relu=False)
        else:
            idn = x
        x = p.conv(o, f"{pre}.conv3", f"{pre}.bn3", 64, 256, 1, relu=True, residual=idn)

    def basic_blocks(x, pre, c):
        for k in range(4):
            q = f"{pre}.{k}"
            o = p.conv(x, f"{q}.conv1", f"{q}.bn1", c, c, 3)
            x = p.conv(o, f"{q}.conv2", f"{q}.bn2", c, c, 3, relu=True, residual=x)
        return x

    def hr_module(xs, pre, nb, multiscale):
        ys = [None] * nb
        # parameters are registered branches first, then fuse layers (state_dict order)
        for b in range(nb):
            ys[b] = basic_blocks(xs[b], f"{pre}.branches.{b}", C[b])
        outs = []
        for i in range(nb if multiscale else 1):
            terms = []
            for j in range(nb):
                if j == i:
                    terms.append((ys[j], 1))
                elif j > i:
                    q = f"{pre}.fuse_layers.{i}.{j}"
                    t = p.conv(ys[j], f"{q}.0", f"{q}.1", C[j], C[i], 1, relu=False)
                    terms.append((t, 2 ** (j - i)))
                else:
                    t = ys[j]
                    for k in range(i - j):
                        last = k == i - j - 1
                        q = f"{pre}.fuse_layers.{i}.{j}.{k}"
                        t = p.conv(t, f"{q}.0", f"{q}.1", C[j], C[i] if last else C[j], 3, 2, relu=not last)
                    terms.append((t, 1))
            outs.append(p.fuse(terms))
        return outs

    # transition1 + stage2
    xs = [p.conv(x, f"{bb}.transition1.0.0", f"{bb}.transition1.0.1", 256, C[0], 3),
          p.conv(x, f"{bb}.transition1.1.0.0", f"{bb}.transition1.1.0.1", 256, C[1], 3, 2)]
    for m in range(cfg["num_modules"][0]):
        xs = hr_module(xs, f"{bb}.stage2.{m}", 2, True)
    xs = xs + [p.conv(xs[-1], f"{bb}.transition2.2.0.0", f"{bb}.transition2.2.0.1", C[1], C[2], 3, 2)]
    for m in range(cfg["num_modules"][1]):
        xs = hr_module(xs, f"{bb}.stage3.{m}", 3, True)
    xs = xs + [p.conv(xs[-1], f"{bb}.transition3.3.0.0", f"{bb}.transition3.3.0.1", C[2], C[3], 3, 2)]
    nm = cfg["num_modules"][2]
    for m in range(nm):
        xs = hr_module(xs, f"{bb}.stage4.{m}", 4, m != nm - 1)

    # head: TopdownHeatmapSimpleHead, 0 deconvs, 1x1 final conv with bias (cfg :73-79)
    p.params["keypoint_head.final_layer.weight"] = (num_joints, C[0], 1, 1)
    p.params["keypoint_head.final_layer.bias"] = (num_joints,)
    t0 = p.tensors[xs[0]]
    out = p._t(num_joints, t0.H, t0.W)
    p.ops.append(Op(OP_HEAD, out, [xs[0]], conv="keypoint_head.final_layer", ksize=1, cin=C[0], cout=num_joints,
                    has_bias=True))
    p.out_tensor = out
    p.finalize()
    return p


def conv_macs(p: Program) -> int:
    total = 0
    for op in p.ops:
        if op.kind in (OP_STEM, OP_CONV, OP_HEAD):
            t = p.tensors[op.out]
            total += t.H * t.W * op.cout * op.cin * op.ksize * op.ksize
    return total
