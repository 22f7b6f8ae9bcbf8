"""Oracle: numpy restatement of OpenCV's fixed-point ``warpAffine(INTER_LINEAR, BORDER_CONSTANT 0)``.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  PINNED against cv2.warpAffine itself (the library
mmpose's TopDownAffine calls; reference call site pose_pipeline/wrappers/mmpose.py:75) -- bit-exact,
tests/test_oracle.py.  It exists so the CUDA crop kernel (csrc/kernels_simt.cu warp_crop_kernel) can be
checked against the *algorithm* independent of which SIMD path a given cv2 build dispatches to.
"""
import numpy as np


def invert_affine(M):
    M = np.array(M, np.float64).reshape(6).copy()
    D = M[0] * M[4] - M[1] * M[3]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = M[4] * D, M[0] * D
    M[0] = A11; M[1] *= -D; M[3] *= -D; M[4] = A22
    b1 = -M[0] * M[2] - M[1] * M[5]
    b2 = -M[3] * M[2] - M[4] * M[5]
    M[2], M[5] = b1, b2
    return M


def warp_affine_fixedpoint(img: np.ndarray, trans: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    H, W = img.shape[:2]
    M = invert_affine(trans)
    x = np.arange(out_w, dtype=np.float64)
    y = np.arange(out_h, dtype=np.float64)
    adelta = np.rint(M[0] * x * 1024).astype(np.int64)
    bdelta = np.rint(M[3] * x * 1024).astype(np.int64)
    X0 = np.rint((M[1] * y + M[2]) * 1024).astype(np.int64) + 16
    Y0 = np.rint((M[4] * y + M[5]) * 1024).astype(np.int64) + 16
    X = (X0[:, None] + adelta[None, :]) >> 5
    Y = (Y0[:, None] + bdelta[None, :]) >> 5
    sx = np.clip(X >> 5, -32768, 32767); sy = np.clip(Y >> 5, -32768, 32767)
    ax = X & 31; ay = Y & 31

    def tap(yy, xx):
        ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
        v = img[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)].astype(np.int64)
        return v * ok[..., None]

    w00 = ((32 - ax) * (32 - ay))[..., None]; w01 = (ax * (32 - ay))[..., None]
    w10 = ((32 - ax) * ay)[..., None]; w11 = (ax * ay)[..., None]
    acc = w00 * tap(sy, sx) + w01 * tap(sy, sx + 1) + w10 * tap(sy + 1, sx) + w11 * tap(sy + 1, sx + 1)
    return ((acc + 512) >> 10).astype(np.uint8)
