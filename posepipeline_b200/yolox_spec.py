"""YOLOX-X detector description: parameter names / shapes and the layer program for the engine's detector executor
(csrc/detector.cu).

Host-side graph building for the detector the reference selects with ``mmtrack_bounding_boxes(video, "bytetrack")``
(``pose_pipeline/wrappers/mmtrack.py:20-23``): ``3rdparty/mmtracking/_base_/models/yolox_x_8x8.py:5-26`` (CSPDarknet deepen
1.33 / widen 1.25, YOLOXPAFPN [320,640,1280] -> 320 with 4 CSP blocks, YOLOXHead 320/320) as overridden by
``mot/bytetrack/bytetrack_yolox_x_crowdhuman_mot17-private-half.py:9-20`` (input (800,1440), one class, score_thr 0.01,
NMS IoU 0.7).  Parameter names are mmdet's ``state_dict`` keys, so the checkpoint the config points at
loads unchanged (with or without mmtrack's ``detector.`` prefix).

Program form: a flat list of ops over activation tensors whose operands may be 16-channel-aligned SLICES of wider tensors.
Every ``torch.cat`` of the network (CSP layers, the SPP bottleneck, both PAFPN paths) is such a wide tensor whose producers
write their slice directly -- no concatenation is ever materialised.  Further fusions decided here:
  * a CSP layer's ``main_conv`` and ``short_conv`` read the same input: ONE 1x1 convolution with the stacked weights;
  * the first convolutions of the classification and the regression tower of a head level likewise;
  * the last Darknet block of a CSP layer writes in place into the ``main`` half of that stacked tensor.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

GOP_INPUT, GOP_CONV, GOP_MAXPOOL, GOP_UPSAMPLE, GOP_DETHEAD = 0, 1, 2, 3, 4
ACT_NONE, ACT_RELU, ACT_SILU = 0, 1, 2
IMG_SCALE = (800, 1440)
SIZE_DIVISOR = 32
STRIDES = (8, 16, 32)
BN_EPS = 1e-3


@dataclass
class GTensor:
    tid: int
    C: int
    H: int
    W: int
    first_def: int = 10 ** 9
    last_use: int = -1


@dataclass
class GOp:
    kind: int
    inp: int = -1
    out: int = -1
    res: int = -1
    in_coff: int = 0
    out_coff: int = 0
    res_coff: int = 0
    cin: int = 0
    cout: int = 0
    ksize: int = 1
    stride: int = 1
    act: int = ACT_SILU
    convs: List[Tuple[str, Optional[str]]] = field(default_factory=list)   # (conv prefix, bn prefix) stacked along Cout
    cin_real: int = 0                                                      # input channels that carry data (stem: 12 of 16)
    head: Optional[Tuple[str, str, str]] = None                            # DETHEAD: conv_cls, conv_reg, conv_obj prefixes


def rescale_size(h: int, w: int, scale=IMG_SCALE) -> Tuple[int, int]:
    """mmcv.rescale_size with keep_ratio: (new_h, new_w)."""
    f = min(max(scale) / max(h, w), min(scale) / min(h, w))
    return int(h * float(f) + 0.5), int(w * float(f) + 0.5)


def net_size(frame_h: int, frame_w: int) -> Tuple[int, int, int, int]:
    """-> (resized_h, resized_w, net_h, net_w): Resize(keep_ratio) then Pad(size_divisor=32)."""
    rh, rw = rescale_size(frame_h, frame_w)
    nh = (rh + SIZE_DIVISOR - 1) // SIZE_DIVISOR * SIZE_DIVISOR
    nw = (rw + SIZE_DIVISOR - 1) // SIZE_DIVISOR * SIZE_DIVISOR
    return rh, rw, nh, nw


class YoloxProgram:
    def __init__(self, net_h: int, net_w: int, num_classes: int = 1, deepen: float = 1.33, widen: float = 1.25):
        assert net_h % 32 == 0 and net_w % 32 == 0
        self.net_h, self.net_w, self.num_classes = net_h, net_w, num_classes
        self.tensors: List[GTensor] = []
        self.ops: List[GOp] = []
        self.params: Dict[str, Tuple[int, ...]] = {}
        self.probes: Dict[str, Tuple[int, int, int]] = {}      # oracle module name -> (tensor, coff, C) holding its output
        self.levels: List[Tuple[int, int, int]] = []           # (stride, H, W) of the head levels
        self._build(deepen, widen)
        self._finalize()

    # ---- helpers
    def _t(self, C, H, W) -> int:
        self.tensors.append(GTensor(len(self.tensors), C, H, W))
        return len(self.tensors) - 1

    def _conv_params(self, conv, bn, cout, cin, k):
        self.params[f"{conv}.weight"] = (cout, cin, k, k)
        if bn:
            for leaf in ("weight", "bias", "running_mean", "running_var"):
                self.params[f"{bn}.{leaf}"] = (cout,)
            self.params[f"{bn}.num_batches_tracked"] = ()

    def conv(self, src, names, cin, couts, k, stride=1, dst=None, res=None, cin_real=None):
        """src/dst/res = (tensor, coff).  names: module prefixes (ConvModule: <p>.conv + <p>.bn) stacked along Cout."""
        t = self.tensors[src[0]]
        Ho, Wo = t.H // stride, t.W // stride
        cout = sum(couts)
        for n, c in zip(names, couts):
            self._conv_params(f"{n}.conv", f"{n}.bn", c, cin_real or cin, k)
        if dst is None:
            dst = (self._t(cout, Ho, Wo), 0)
        assert self.tensors[dst[0]].H == Ho and self.tensors[dst[0]].W == Wo
        self.ops.append(GOp(GOP_CONV, src[0], dst[0], res[0] if res else -1, src[1], dst[1], res[1] if res else 0, cin, cout, k, stride,
                            ACT_SILU, [(f"{n}.conv", f"{n}.bn") for n in names], cin_real or cin))
        off = dst[1]
        for n, c in zip(names, couts):
            self.probes[n] = (dst[0], off, c)
            off += c
        return dst

    def csp(self, src, cin, prefix, cout, nb, add_id, dst=None):
        t = self.tensors[src[0]]
        mid = cout // 2
        ms = self.conv(src, [f"{prefix}.main_conv", f"{prefix}.short_conv"], cin, [mid, mid], 1)
        cur = (ms[0], 0)
        for b in range(nb):
            h = self.conv(cur, [f"{prefix}.blocks.{b}.conv1"], mid, [mid], 1)
            out = (ms[0], 0) if b == nb - 1 else (self._t(mid, t.H, t.W), 0)
            self.conv(h, [f"{prefix}.blocks.{b}.conv2"], mid, [mid], 3, dst=out, res=cur if add_id else None)
            cur = out
        return self.conv((ms[0], 0), [f"{prefix}.final_conv"], 2 * mid, [cout], 1, dst=dst)

    def _build(self, deepen, widen):
        H2, W2 = self.net_h // 2, self.net_w // 2
        base = [64, 128, 256, 512, 1024]
        ch = [int(c * widen) for c in base]                       # 80 160 320 640 1280
        nbs = [max(round(n * deepen), 1) for n in (3, 9, 9, 3)]   # 4 12 12 4
        t_in = self._t(16, H2, W2)
        self.ops.append(GOp(GOP_INPUT, out=t_in, cout=16, act=ACT_NONE))
        self.probes["__input__"] = (t_in, 0, 16)                  # Focus(space-to-depth) of the resized + padded frame
        x = self.conv((t_in, 0), ["backbone.stem.conv"], 16, [ch[0]], 3, cin_real=12)
        # PAFPN concat tensors (producers write their slices)
        H8, W8, H16, W16, H32, W32 = self.net_h // 8, self.net_w // 8, self.net_h // 16, self.net_w // 16, self.net_h // 32, self.net_w // 32
        cat1 = self._t(2 * ch[2], H8, W8)         # [up(red1) | stage2 out]      -> top_down_blocks.1
        cat0 = self._t(2 * ch[3], H16, W16)       # [up(red0) | stage3 out]      -> top_down_blocks.0
        bu0 = self._t(2 * ch[2], H16, W16)        # [down0 | red1]               -> bottom_up_blocks.0
        bu1 = self._t(2 * ch[3], H32, W32)        # [down1 | red0]               -> bottom_up_blocks.1
        # backbone
        y = self.conv(x, ["backbone.stage1.0"], ch[0], [ch[1]], 3, 2)
        x = self.csp(y, ch[1], "backbone.stage1.1", ch[1], nbs[0], True)
        y = self.conv(x, ["backbone.stage2.0"], ch[1], [ch[2]], 3, 2)
        s2 = self.csp(y, ch[2], "backbone.stage2.1", ch[2], nbs[1], True, dst=(cat1, ch[2]))
        y = self.conv(s2, ["backbone.stage3.0"], ch[2], [ch[3]], 3, 2)
        s3 = self.csp(y, ch[3], "backbone.stage3.1", ch[3], nbs[2], True, dst=(cat0, ch[3]))
        y = self.conv(s3, ["backbone.stage4.0"], ch[3], [ch[4]], 3, 2)
        mid = ch[4] // 2
        spp = self._t(4 * mid, H32, W32)
        self.conv(y, ["backbone.stage4.1.conv1"], ch[4], [mid], 1, dst=(spp, 0))
        for i, k in enumerate((5, 9, 13)):
            self.ops.append(GOp(GOP_MAXPOOL, spp, spp, in_coff=0, out_coff=(i + 1) * mid, cin=mid, cout=mid, ksize=k, act=ACT_NONE))
        y = self.conv((spp, 0), ["backbone.stage4.1.conv2"], 4 * mid, [ch[4]], 1)
        s4 = self.csp(y, ch[4], "backbone.stage4.2", ch[4], nbs[3], False)
        # neck: top-down
        red0 = self.conv(s4, ["neck.reduce_layers.0"], ch[4], [ch[3]], 1, dst=(bu1, ch[3]))
        self.ops.append(GOp(GOP_UPSAMPLE, red0[0], cat0, in_coff=red0[1], out_coff=0, cin=ch[3], cout=ch[3], act=ACT_NONE))
        td0 = self.csp((cat0, 0), 2 * ch[3], "neck.top_down_blocks.0", ch[3], nbs[3], False)
        red1 = self.conv(td0, ["neck.reduce_layers.1"], ch[3], [ch[2]], 1, dst=(bu0, ch[2]))
        self.ops.append(GOp(GOP_UPSAMPLE, red1[0], cat1, in_coff=red1[1], out_coff=0, cin=ch[2], cout=ch[2], act=ACT_NONE))
        td1 = self.csp((cat1, 0), 2 * ch[2], "neck.top_down_blocks.1", ch[2], nbs[3], False)
        # neck: bottom-up
        self.conv(td1, ["neck.downsamples.0"], ch[2], [ch[2]], 3, 2, dst=(bu0, 0))
        o1 = self.csp((bu0, 0), 2 * ch[2], "neck.bottom_up_blocks.0", ch[3], nbs[3], False)
        self.conv(o1, ["neck.downsamples.1"], ch[3], [ch[3]], 3, 2, dst=(bu1, 0))
        o2 = self.csp((bu1, 0), 2 * ch[3], "neck.bottom_up_blocks.1", ch[4], nbs[3], False)
        feats = [self.conv(o, [f"neck.out_convs.{i}"], c, [ch[2]], 1) for i, (o, c) in enumerate(zip((td1, o1, o2), (ch[2], ch[3], ch[4])))]
        # head
        F = ch[2]
        prior = 0
        for i, (f, stride) in enumerate(zip(feats, STRIDES)):
            t = self.tensors[f[0]]
            both = self.conv(f, [f"bbox_head.multi_level_cls_convs.{i}.0", f"bbox_head.multi_level_reg_convs.{i}.0"], F, [F, F], 3)
            cls = self.conv((both[0], 0), [f"bbox_head.multi_level_cls_convs.{i}.1"], F, [F], 3)
            reg = self.conv((both[0], F), [f"bbox_head.multi_level_reg_convs.{i}.1"], F, [F], 3)
            names = (f"bbox_head.multi_level_conv_cls.{i}", f"bbox_head.multi_level_conv_reg.{i}", f"bbox_head.multi_level_conv_obj.{i}")
            for n, c in zip(names, (self.num_classes, 4, 1)):
                self.params[f"{n}.weight"] = (c, F, 1, 1)
                self.params[f"{n}.bias"] = (c,)
            self.ops.append(GOp(GOP_DETHEAD, cls[0], -1, reg[0], 0, prior, 0, F, 6, 1, stride, ACT_NONE, head=names))
            self.levels.append((stride, t.H, t.W))
            prior += t.H * t.W
        self.num_priors = prior

    def _finalize(self):
        for i, op in enumerate(self.ops):
            for t in (op.inp, op.res):
                if t >= 0:
                    self.tensors[t].last_use = max(self.tensors[t].last_use, i)
            if op.out >= 0:
                self.tensors[op.out].first_def = min(self.tensors[op.out].first_def, i)
                self.tensors[op.out].last_use = max(self.tensors[op.out].last_use, i)

    def assign_slots(self):
        """Greedy buffer sharing between tensors with disjoint live ranges -> (slot_of_tensor, padded elems per image per slot)."""
        def elems(t):
            return (t.H + 2) * (t.W + 2) * t.C
        slots: List[List[int]] = []                                # [free_after_op, size]
        slot_of = [-1] * len(self.tensors)
        for t in sorted(self.tensors, key=lambda t: t.first_def):
            best = -1
            for s, (free_after, size) in enumerate(slots):
                if free_after < t.first_def and (best < 0 or abs(size - elems(t)) < abs(slots[best][1] - elems(t))):
                    best = s
            if best < 0:
                slots.append([t.last_use, elems(t)])
                best = len(slots) - 1
            else:
                slots[best] = [t.last_use, max(slots[best][1], elems(t))]
            slot_of[t.tid] = best
        return slot_of, [s[1] for s in slots]

    def conv_macs(self) -> int:
        """Multiply-accumulates of one forward pass (convolutions + the head's 1x1 outputs), real input channels only."""
        total = 0
        for op in self.ops:
            if op.kind == GOP_CONV:
                t = self.tensors[op.out]
                total += t.H * t.W * op.cout * op.cin_real * op.ksize * op.ksize
            elif op.kind == GOP_DETHEAD:
                t = self.tensors[op.inp]
                total += t.H * t.W * op.cin * (self.num_classes + 5)
        return total


def build_yolox_program(frame_h: int, frame_w: int) -> YoloxProgram:
    _, _, nh, nw = net_size(frame_h, frame_w)
    return YoloxProgram(nh, nw)
