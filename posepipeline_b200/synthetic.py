"""Seeded synthetic inputs for tests and bench.py (there are no videos, no
checkpoints and no network on the box, so every measured workload is synthetic and says so).

Frames are 1080p BGR uint8 with low-frequency structure plus blobs plus mild pixel noise, so
that the network output is not white noise; bounding boxes follow the distribution the
survey fixed for config 2.
"""
from __future__ import annotations

import numpy as np


def synthetic_frame(seed: int, height: int = 1080, width: int = 1920) -> np.ndarray:
    """One BGR uint8 frame (H,W,3)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float32)
    img = np.zeros((height, width, 3), np.float32)
    for c in range(3):
        acc = np.zeros((height, width), np.float32)
        for _ in range(4):
            fx, fy = rng.uniform(0.002, 0.02, 2)
            ph = rng.uniform(0, 2 * np.pi)
            acc += rng.uniform(0.3, 1.0) * np.sin(2 * np.pi * (fx * xx + fy * yy) + ph).astype(np.float32)
        for _ in range(12):
            cx, cy = rng.uniform(0, width), rng.uniform(0, height)
            r = rng.uniform(15, 120)
            acc += rng.uniform(-2, 2) * np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * r * r)).astype(np.float32)
        img[..., c] = acc
    img = (img - img.min()) / (img.max() - img.min() + 1e-6) * 255.0
    img += rng.normal(0, 6.0, img.shape).astype(np.float32)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def synthetic_frames(n: int, seed0: int = 0, height: int = 1080, width: int = 1920) -> np.ndarray:
    return np.stack([synthetic_frame(seed0 + i, height, width) for i in range(n)])


def cheap_frames(n: int, seed: int = 0, height: int = 1080, width: int = 1920) -> np.ndarray:
    """Fast variant for large benchmark batches: a handful of rendered frames, cyclically shifted."""
    base = synthetic_frames(min(n, 4), seed, height, width)
    out = np.empty((n, height, width, 3), np.uint8)
    for i in range(n):
        out[i] = np.roll(base[i % len(base)], shift=(37 * (i // len(base)), 91 * (i // len(base))), axis=(0, 1))
    return out


def synthetic_bboxes(n: int, seed: int = 1234) -> np.ndarray:
    """(n,4) float64 x,y,w,h: x~U(0,1500) y~U(0,400) w~U(120,420) h~U(300,680)  (SURVEY §8(d) config 2)."""
    rng = np.random.default_rng(seed)
    return np.stack([rng.uniform(0, 1500, n), rng.uniform(0, 400, n),
                     rng.uniform(120, 420, n), rng.uniform(300, 680, n)], axis=1)


def synthetic_keypoints_2d(n: int, seed: int = 7, width: int = 1920, height: int = 1080) -> np.ndarray:
    """(n,17,3) [x_px, y_px, conf] for the lifter, smooth in time."""
    rng = np.random.default_rng(seed)
    base = rng.uniform(0.2, 0.8, (1, 17, 2)) * np.array([width, height])
    t = np.arange(n)[:, None, None]
    wob = 40.0 * np.sin(2 * np.pi * t / rng.uniform(30, 200, (1, 17, 2)) + rng.uniform(0, 6.28, (1, 17, 2)))
    kp = base + wob + rng.normal(0, 1.0, (n, 17, 2))
    conf = rng.uniform(0.3, 1.0, (n, 17, 1))
    return np.concatenate([kp, conf], axis=2)
