"""Oracle: VideoPose3D ``TemporalModelOptimized1f`` + the wrapper's normalisation and windowing.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  PARITY UNPINNED: facebookresearch/VideoPose3D is an
un-vendored dependency (reference ``requirements.txt:32``, imported at
``pose_pipeline/wrappers/videopose3d.py:43-44`` from ``$VIDEOPOSE3D_PATH``).  Restated from the
published model (``common/model.py``: TemporalModelOptimized1f; ``common/generators.py``:
ChunkedGenerator) as configured at ``wrappers/videopose3d.py:10-16,46-75``.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


class TemporalModelOptimized1f(nn.Module):
    """filter_widths=[3,3,3,3,3], causal=False, channels=1024; eval mode (dropout = identity)."""

    def __init__(self, num_joints_in=17, in_features=2, num_joints_out=17, filter_widths=(3, 3, 3, 3, 3), channels=1024):
        super().__init__()
        self.num_joints_in, self.in_features, self.num_joints_out = num_joints_in, in_features, num_joints_out
        self.filter_widths = list(filter_widths)
        self.pad = [filter_widths[0] // 2]
        self.expand_conv = nn.Conv1d(num_joints_in * in_features, channels, filter_widths[0], stride=filter_widths[0], bias=False)
        self.expand_bn = nn.BatchNorm1d(channels, momentum=0.1)
        convs, bns = [], []
        self.causal_shift = [0]
        next_dilation = filter_widths[0]
        for i in range(1, len(filter_widths)):
            self.pad.append((filter_widths[i] - 1) * next_dilation // 2)
            self.causal_shift.append(0)
            convs.append(nn.Conv1d(channels, channels, filter_widths[i], stride=filter_widths[i], bias=False))
            bns.append(nn.BatchNorm1d(channels, momentum=0.1))
            convs.append(nn.Conv1d(channels, channels, 1, dilation=1, bias=False))
            bns.append(nn.BatchNorm1d(channels, momentum=0.1))
            next_dilation *= filter_widths[i]
        self.layers_conv = nn.ModuleList(convs)
        self.layers_bn = nn.ModuleList(bns)
        self.shrink = nn.Conv1d(channels, num_joints_out * 3, 1)
        self.eval()

    def receptive_field(self):
        return 1 + 2 * sum(self.pad)

    @torch.no_grad()
    def forward(self, x):                      # x (B, T, J, F)
        B, T = x.shape[:2]
        x = x.view(B, T, -1).permute(0, 2, 1)
        x = F.relu(self.expand_bn(self.expand_conv(x)))
        for i in range(len(self.pad) - 1):
            res = x[:, :, self.causal_shift[i + 1] + self.filter_widths[i + 1] // 2::self.filter_widths[i + 1]]
            x = F.relu(self.layers_bn[2 * i](self.layers_conv[2 * i](x)))
            x = res + F.relu(self.layers_bn[2 * i + 1](self.layers_conv[2 * i + 1](x)))
        x = self.shrink(x)
        return x.permute(0, 2, 1).view(B, -1, self.num_joints_out, 3)


def normalize_screen_coordinates(X, w, h):      # wrappers/videopose3d.py:26-33
    assert X.shape[-1] == 2
    if w > h:
        return X / w * 2 - [1, h / w]
    return X / h * 2 - [w / h, 1]


def windows(kp_norm: np.ndarray, pad: int = 121) -> np.ndarray:
    """ChunkedGenerator(chunk_length=1, pad=121, causal_shift=0, shuffle=False): one (2*pad+1)-frame
    window per output frame i = [i-pad, i+pad], clipped to the video and edge-replicated."""
    N = kp_norm.shape[0]
    padded = np.pad(kp_norm, ((pad, pad), (0, 0), (0, 0)), "edge")
    idx = np.arange(N)[:, None] + np.arange(2 * pad + 1)[None, :]
    return padded[idx]


def process_videopose3d(keypoints: np.ndarray, height: int, width: int, net: TemporalModelOptimized1f, batch_size: int = 32):
    """== reference ``process_videopose3d`` (wrappers/videopose3d.py:19-91) given fetched inputs."""
    N = keypoints.shape[0]
    kp = normalize_screen_coordinates(keypoints[:, :, :2], width, height)[:, :, :2]
    pad = (net.receptive_field() - 1) // 2
    win = windows(kp, pad)
    results = []
    dt = next(net.parameters()).dtype
    for s in range(0, N, batch_size):
        sample = torch.from_numpy(win[s:s + batch_size].astype("float32")).contiguous().to(dt)
        results.append(net(sample).float().numpy()[:, 0, ...])
    results = np.concatenate(results, axis=0)
    keypoints_3d = np.zeros((N, 17, 3))
    keypoints_3d[np.arange(N)] = results
    return {"keypoints_3d": keypoints_3d, "keypoints_valid": [True] * N}


@torch.no_grad()
def dilated_whole_sequence(net: TemporalModelOptimized1f, kp_norm: np.ndarray) -> np.ndarray:
    """The same network evaluated ONCE over the edge-padded sequence with dilations 1,3,9,27,81 instead of once per 243-frame
    window with strides (identical sums of identical products; tests/test_oracle.py checks it against the windowed form).
    Used as the checker for long sequences, where the windowed form (176 MMAC per frame) is too slow on a CPU."""
    pad = (net.receptive_field() - 1) // 2
    dt = next(net.parameters()).dtype
    xp = np.pad(kp_norm, ((pad, pad), (0, 0), (0, 0)), "edge").reshape(1, -1, 34).transpose(0, 2, 1)
    y = torch.from_numpy(np.ascontiguousarray(xp)).to(dt)
    y = F.relu(net.expand_bn(F.conv1d(y, net.expand_conv.weight)))
    dil = 3
    for i in range(4):
        res = y[:, :, dil:-dil]
        y = F.relu(net.layers_bn[2 * i](F.conv1d(y, net.layers_conv[2 * i].weight, dilation=dil)))
        y = res + F.relu(net.layers_bn[2 * i + 1](F.conv1d(y, net.layers_conv[2 * i + 1].weight)))
        dil *= 3
    return net.shrink(y)[0].T.reshape(-1, 17, 3).double().numpy()


def load_lifter(state_dict, dtype=torch.float32):
    net = TemporalModelOptimized1f()
    sd = {k: (v if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v).copy())) for k, v in state_dict.items()}
    net.load_state_dict(sd, strict=True)
    return net.to(dtype).eval()
