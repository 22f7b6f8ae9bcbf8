"""Bring-up driver for the tcgen05 conv kernel: single layers vs a float64 torch reference.
    python tests/tc_bringup.py [case indices...]      (env PE_TC_BO_MODE / PE_TC_MT / PE_TC_NS / PE_TC_VERBOSE)"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from posepipeline_b200 import engine as E

CASES = [  # (Cin, Cout, ks, H, W, nimg, res, relu)
    (16, 16, 1, 6, 6, 1, False, False),
    (64, 48, 1, 10, 7, 2, False, True),
    (16, 16, 3, 6, 6, 1, False, False),
    (48, 48, 3, 96, 72, 2, True, True),
    (96, 96, 3, 48, 36, 3, True, True),
    (192, 192, 3, 24, 18, 5, True, True),
    (384, 384, 3, 12, 9, 7, True, True),
    (256, 48, 3, 96, 72, 1, False, True),
    (64, 256, 1, 96, 72, 2, True, True),
    (256, 64, 1, 96, 72, 2, False, True),
    (384, 48, 1, 12, 9, 4, False, False),
    (64, 64, 3, 96, 72, 2, False, True),
    (16, 16, 3, 8, 8, 1, False, False, 2),        # stride 2 (space-to-depth + 2x2 taps)
    (48, 96, 3, 96, 72, 2, False, True, 2),
    (64, 64, 3, 192, 144, 1, False, True, 2),
    (192, 384, 3, 24, 18, 3, False, False, 2),
    (256, 96, 3, 96, 72, 2, False, True, 2),
    (96, 96, 3, 48, 36, 40, True, True),          # several tiles per CTA (persistent schedule, alternating epilogue sets)
    (48, 48, 3, 96, 72, 24, True, True),
]


def main():
    idx = [int(a) for a in sys.argv[1:]] or range(len(CASES))
    eng = E.PoseEngine(0)
    for i in idx:
        cin, cout, ks, H, W, n, res, relu = CASES[i][:8]
        stride = CASES[i][8] if len(CASES[i]) > 8 else 1
        rng = np.random.default_rng(i)
        x = rng.standard_normal((n, cin, H, W)).astype(np.float32)
        w = (rng.standard_normal((cout, cin, ks, ks)) / np.sqrt(cin * ks * ks)).astype(np.float32)
        b = rng.standard_normal(cout).astype(np.float32)
        r = rng.standard_normal((n, cout, H // stride, W // stride)).astype(np.float32) if res else None
        ref = torch.nn.functional.conv2d(torch.from_numpy(x).double(), torch.from_numpy(w).double(), torch.from_numpy(b).double(), padding=ks // 2, stride=stride)
        if res:
            ref = ref + torch.from_numpy(r).double()
        if relu:
            ref = torch.relu(ref)
        ref = ref.numpy()
        out = {}
        for tc in (0, 1):
            t0 = time.time()
            got = E.conv_test(eng, x, w, b, r, relu, bool(tc), stride)
            err = np.abs(got - ref).max() / np.abs(ref).max()
            out[tc] = err
            print(f"case {i} {CASES[i]} {'TC  ' if tc else 'SIMT'} max rel err {err:.3e}  ({time.time()-t0:.2f}s)", flush=True)
        if not out[1] < 5e-6:
            bad = np.argwhere(np.abs(got - ref) > 1e-3 * np.abs(ref).max())
            print("   FAIL: first bad (n,c,y,x):", bad[:5].tolist(), "count", len(bad), "of", got.size, flush=True)
    print("done")


if __name__ == "__main__":
    main()
