"""Oracle: the detector half of ``mmtrack.apis.inference_mot`` for ByteTrack -- YOLOX-X at 800x1440 as mmdet 2.x defines it.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  PARITY UNPINNED: mmtrack 0.x / mmdet 2.x / mmcv 1.x are un-vendored, unpinned
dependencies (reference ``requirements.txt:9-12``) reached through ``pose_pipeline/wrappers/mmtrack.py:30,45``; none is
installable here and the reference ships no golden detections.  Restated from the published mmdet 2.x sources
(``mmdet/models/backbones/csp_darknet.py``, ``necks/yolox_pafpn.py``, ``dense_heads/yolox_head.py``, ``utils/csp_layer.py``,
``mmcv.ops.nms``) as configured by the reference's own config files:
  3rdparty/mmtracking/_base_/models/yolox_x_8x8.py:5-26      CSPDarknet deepen 1.33 / widen 1.25, PAFPN [320,640,1280]->320 x4,
                                                             YOLOXHead in/feat 320
  3rdparty/mmtracking/mot/bytetrack/bytetrack_yolox_x_crowdhuman_mot17-private-half.py:6,9-20   input (800,1440), 1 class,
                                                             score_thr 0.01, nms iou 0.7
  ...:60-81  test pipeline: Resize keep_ratio to (800,1440), Normalize mean 0 / std 1 / to_rgb False, Pad to /32 with 114.
The wrapper converts BGR->RGB itself (wrappers/mmtrack.py:43) and the pipeline does not swap again, so the network sees
R,G,B planes of [0,255] floats.  Parameter names follow the mmdet ``state_dict`` (``backbone.stem.conv.conv.weight`` ...;
inside the mmtrack checkpoint they carry a ``detector.`` prefix), so a real checkpoint loads unchanged.
"""
import math

import cv2
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

IMG_SCALE = (800, 1440)          # (short edge, long edge) as mmcv.imrescale reads the config tuple
SIZE_DIVISOR, PAD_VAL = 32, 114.0
SCORE_THR, NMS_IOU = 0.01, 0.7
STRIDES = (8, 16, 32)


class ConvModule(nn.Module):
    """mmcv ConvModule(conv no-bias, BN eps 1e-3 momentum 0.03, Swish)."""

    def __init__(self, cin, cout, k, stride=1):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, stride, (k - 1) // 2, bias=False)
        self.bn = nn.BatchNorm2d(cout, eps=0.001, momentum=0.03)

    def forward(self, x):
        y = self.bn(self.conv(x))
        return y * torch.sigmoid(y)


class Focus(nn.Module):
    def __init__(self, cin, cout, k=3):
        super().__init__()
        self.conv = ConvModule(cin * 4, cout, k)

    def forward(self, x):
        tl, tr = x[..., ::2, ::2], x[..., ::2, 1::2]
        bl, br = x[..., 1::2, ::2], x[..., 1::2, 1::2]
        return self.conv(torch.cat((tl, bl, tr, br), dim=1))


class DarknetBottleneck(nn.Module):
    def __init__(self, cin, cout, expansion=0.5, add_identity=True):
        super().__init__()
        hidden = int(cout * expansion)
        self.conv1 = ConvModule(cin, hidden, 1)
        self.conv2 = ConvModule(hidden, cout, 3)
        self.add_identity = add_identity and cin == cout

    def forward(self, x):
        out = self.conv2(self.conv1(x))
        return out + x if self.add_identity else out


class CSPLayer(nn.Module):
    def __init__(self, cin, cout, expand_ratio=0.5, num_blocks=1, add_identity=True):
        super().__init__()
        mid = int(cout * expand_ratio)
        self.main_conv = ConvModule(cin, mid, 1)
        self.short_conv = ConvModule(cin, mid, 1)
        self.final_conv = ConvModule(2 * mid, cout, 1)
        self.blocks = nn.Sequential(*[DarknetBottleneck(mid, mid, 1.0, add_identity) for _ in range(num_blocks)])

    def forward(self, x):
        return self.final_conv(torch.cat((self.blocks(self.main_conv(x)), self.short_conv(x)), dim=1))


class SPPBottleneck(nn.Module):
    def __init__(self, cin, cout, kernel_sizes=(5, 9, 13)):
        super().__init__()
        mid = cin // 2
        self.conv1 = ConvModule(cin, mid, 1)
        self.poolings = nn.ModuleList([nn.MaxPool2d(k, stride=1, padding=k // 2) for k in kernel_sizes])
        self.conv2 = ConvModule(mid * (len(kernel_sizes) + 1), cout, 1)

    def forward(self, x):
        x = self.conv1(x)
        return self.conv2(torch.cat([x] + [p(x) for p in self.poolings], dim=1))


class CSPDarknet(nn.Module):
    ARCH = [[64, 128, 3, True, False], [128, 256, 9, True, False], [256, 512, 9, True, False], [512, 1024, 3, False, True]]

    def __init__(self, deepen=1.33, widen=1.25):
        super().__init__()
        self.stem = Focus(3, int(64 * widen), 3)
        for i, (cin, cout, nb, add_id, spp) in enumerate(self.ARCH):
            cin, cout, nb = int(cin * widen), int(cout * widen), max(round(nb * deepen), 1)
            stage = [ConvModule(cin, cout, 3, 2)]
            if spp:
                stage.append(SPPBottleneck(cout, cout))
            stage.append(CSPLayer(cout, cout, num_blocks=nb, add_identity=add_id))
            setattr(self, f"stage{i + 1}", nn.Sequential(*stage))

    def forward(self, x):
        x = self.stem(x)
        outs = []
        for i in range(1, 5):
            x = getattr(self, f"stage{i}")(x)
            if i >= 2:
                outs.append(x)
        return outs


class YOLOXPAFPN(nn.Module):
    def __init__(self, in_channels=(320, 640, 1280), out_channels=320, num_csp_blocks=4):
        super().__init__()
        c = list(in_channels)
        self.reduce_layers, self.top_down_blocks = nn.ModuleList(), nn.ModuleList()
        for idx in range(len(c) - 1, 0, -1):
            self.reduce_layers.append(ConvModule(c[idx], c[idx - 1], 1))
            self.top_down_blocks.append(CSPLayer(c[idx - 1] * 2, c[idx - 1], num_blocks=num_csp_blocks, add_identity=False))
        self.downsamples, self.bottom_up_blocks = nn.ModuleList(), nn.ModuleList()
        for idx in range(len(c) - 1):
            self.downsamples.append(ConvModule(c[idx], c[idx], 3, 2))
            self.bottom_up_blocks.append(CSPLayer(c[idx] * 2, c[idx + 1], num_blocks=num_csp_blocks, add_identity=False))
        self.out_convs = nn.ModuleList([ConvModule(ci, out_channels, 1) for ci in c])

    def forward(self, inputs):
        n = len(inputs)
        inner = [inputs[-1]]
        for idx in range(n - 1, 0, -1):
            high = self.reduce_layers[n - 1 - idx](inner[0])
            inner[0] = high
            up = F.interpolate(high, scale_factor=2, mode="nearest")
            inner.insert(0, self.top_down_blocks[n - 1 - idx](torch.cat([up, inputs[idx - 1]], 1)))
        outs = [inner[0]]
        for idx in range(n - 1):
            down = self.downsamples[idx](outs[-1])
            outs.append(self.bottom_up_blocks[idx](torch.cat([down, inner[idx + 1]], 1)))
        return [conv(o) for conv, o in zip(self.out_convs, outs)]


class YOLOXHead(nn.Module):
    def __init__(self, num_classes=1, in_channels=320, feat_channels=320, stacked_convs=2, n_levels=3):
        super().__init__()
        def tower():
            return nn.Sequential(*[ConvModule(in_channels if i == 0 else feat_channels, feat_channels, 3) for i in range(stacked_convs)])
        self.multi_level_cls_convs = nn.ModuleList([tower() for _ in range(n_levels)])
        self.multi_level_reg_convs = nn.ModuleList([tower() for _ in range(n_levels)])
        self.multi_level_conv_cls = nn.ModuleList([nn.Conv2d(feat_channels, num_classes, 1) for _ in range(n_levels)])
        self.multi_level_conv_reg = nn.ModuleList([nn.Conv2d(feat_channels, 4, 1) for _ in range(n_levels)])
        self.multi_level_conv_obj = nn.ModuleList([nn.Conv2d(feat_channels, 1, 1) for _ in range(n_levels)])

    def forward(self, feats):
        cls, reg, obj = [], [], []
        for i, x in enumerate(feats):
            cf, rf = self.multi_level_cls_convs[i](x), self.multi_level_reg_convs[i](x)
            cls.append(self.multi_level_conv_cls[i](cf))
            reg.append(self.multi_level_conv_reg[i](rf))
            obj.append(self.multi_level_conv_obj[i](rf))
        return cls, reg, obj


class YOLOX(nn.Module):
    def __init__(self):
        super().__init__()
        self.backbone, self.neck, self.bbox_head = CSPDarknet(), YOLOXPAFPN(), YOLOXHead()
        self.eval()

    @torch.no_grad()
    def forward(self, img):
        return self.bbox_head(self.neck(self.backbone(img)))


def load_detector(state_dict, dtype=torch.float32):
    net = YOLOX()
    sd = {}
    for k, v in state_dict.items():
        k = k[len("detector."):] if k.startswith("detector.") else k
        sd[k] = v if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v).copy())
    net.load_state_dict(sd, strict=True)
    return net.to(dtype).eval()


# ------------------------------------------------------------------------------------------ test pipeline
def rescale_size(h, w, scale=IMG_SCALE):
    """mmcv.rescale_size(keep ratio): long edge <= max(scale), short edge <= min(scale); int(x * f + 0.5)."""
    f = min(max(scale) / max(h, w), min(scale) / min(h, w))
    return int(h * float(f) + 0.5), int(w * float(f) + 0.5)


def preprocess(frame_rgb, dtype=np.float32):
    """Resize(keep_ratio) -> Normalize(mean 0, std 1, to_rgb False) -> Pad(size_divisor 32, pad_val 114).
    Returns (CHW float image, scale_factor float32[4])."""
    h, w = frame_rgb.shape[:2]
    nh, nw = rescale_size(h, w)
    img = cv2.resize(frame_rgb, (nw, nh), interpolation=cv2.INTER_LINEAR)
    scale_factor = np.array([nw / w, nh / h, nw / w, nh / h], dtype=np.float32)
    img = img.astype(np.float32)                                           # (x - 0) / 1
    ph, pw = int(math.ceil(nh / SIZE_DIVISOR)) * SIZE_DIVISOR, int(math.ceil(nw / SIZE_DIVISOR)) * SIZE_DIVISOR
    out = np.full((ph, pw, 3), PAD_VAL, np.float32)
    out[:nh, :nw] = img
    return np.ascontiguousarray(out.transpose(2, 0, 1)).astype(dtype), scale_factor


def resize_linear_u8(src, dw, dh):
    """cv2.resize(INTER_LINEAR) on uint8 restated: 11-bit fixed-point coefficients, horizontal pass in int32, vertical pass
    ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2 >> 2; x weights are clamped at the borders, y only its row indices.
    Bit-exact against cv2 4.13 here (tests/test_yolox.py); the CUDA preprocessing kernel follows this arithmetic."""
    sh, sw = src.shape[:2]
    scale_x, scale_y = 1.0 / (dw / sw), 1.0 / (dh / sh)

    def coeffs(dn, sn, scale, clamp_weights):
        d = np.arange(dn, dtype=np.float64)
        f = ((d + 0.5) * scale - 0.5).astype(np.float32)
        s = np.floor(f).astype(np.int64)
        f = (f - s.astype(np.float32)).astype(np.float32)
        if clamp_weights:
            lo, hi = s < 0, s >= sn - 1
            f[lo | hi] = 0
            s[lo] = 0
            s[hi] = sn - 1
        a0 = np.rint((np.float32(1.0) - f) * np.float32(2048)).astype(np.int64)
        a1 = np.rint(f * np.float32(2048)).astype(np.int64)
        return np.clip(s, 0, sn - 1), np.clip(s + 1, 0, sn - 1), a0, a1
    sx, sx1, ax0, ax1 = coeffs(dw, sw, scale_x, True)
    sy, sy1, by0, by1 = coeffs(dh, sh, scale_y, False)
    S = src.astype(np.int64)
    H = S[:, sx] * ax0[None, :, None] + S[:, sx1] * ax1[None, :, None]
    out = (((by0[:, None, None] * (H[sy] >> 4)) >> 16) + ((by1[:, None, None] * (H[sy1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


# ------------------------------------------------------------------------------------------ YOLOXHead.get_bboxes
def decode(cls, reg, obj, scale_factor):
    """Flatten the three levels, decode against MlvlPointGenerator(offset 0) priors, rescale to the original image
    (before NMS, as YOLOXHead.get_bboxes does).  -> boxes (P,4) float32, scores (P,) float32 = sigmoid(cls) * sigmoid(obj)."""
    boxes, scores = [], []
    for c, r, o, stride in zip(cls, reg, obj, STRIDES):
        _, _, H, W = r.shape
        ys, xs = torch.meshgrid(torch.arange(H, dtype=r.dtype) * stride, torch.arange(W, dtype=r.dtype) * stride, indexing="ij")
        pri = torch.stack([xs.reshape(-1), ys.reshape(-1)], -1)
        r = r[0].permute(1, 2, 0).reshape(-1, 4)
        xy = r[:, :2] * stride + pri
        wh = r[:, 2:].exp() * stride
        boxes.append(torch.cat([xy - wh / 2, xy + wh / 2], -1))
        scores.append(c[0].permute(1, 2, 0).reshape(-1).sigmoid() * o[0].permute(1, 2, 0).reshape(-1).sigmoid())
    boxes, scores = torch.cat(boxes), torch.cat(scores)
    boxes = boxes / torch.as_tensor(scale_factor, dtype=boxes.dtype)[None]
    return boxes.float().numpy(), scores.float().numpy()


def nms(boxes, scores, iou_thr=NMS_IOU):
    """mmcv.ops.nms (offset 0): descending score order, suppress IoU > iou_thr.  -> kept indices in score order."""
    order = np.argsort(-scores, kind="stable")
    b = boxes[order].astype(np.float32)
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    keep, dead = [], np.zeros(len(b), bool)
    for i in range(len(b)):
        if dead[i]:
            continue
        keep.append(order[i])
        xx1, yy1 = np.maximum(b[i, 0], b[i + 1:, 0]), np.maximum(b[i, 1], b[i + 1:, 1])
        xx2, yy2 = np.minimum(b[i, 2], b[i + 1:, 2]), np.minimum(b[i, 3], b[i + 1:, 3])
        inter = np.maximum(np.float32(0), xx2 - xx1) * np.maximum(np.float32(0), yy2 - yy1)
        iou = inter / (area[i] + area[i + 1:] - inter)
        dead[i + 1:] |= iou > np.float32(iou_thr)
    return np.asarray(keep, np.int64)


def detect(net, frame_bgr, score_thr=SCORE_THR):
    """One frame exactly as the wrapper + inference_mot's detector half treat it -> (n,5) float32 [x1,y1,x2,y2,score]
    sorted by score (descending)."""
    rgb = cv2.cvtColor(frame_bgr, cv2.COLOR_BGR2RGB)                       # wrappers/mmtrack.py:43
    dt = next(net.parameters()).dtype
    img, sf = preprocess(rgb)
    cls, reg, obj = net(torch.from_numpy(img)[None].to(dt))
    boxes, scores = decode(cls, reg, obj, sf)
    valid = scores >= np.float32(score_thr)
    boxes, scores = boxes[valid], scores[valid]
    if len(scores) == 0:
        return np.zeros((0, 5), np.float32)
    keep = nms(boxes, scores)
    return np.concatenate([boxes[keep], scores[keep, None]], axis=1).astype(np.float32)
