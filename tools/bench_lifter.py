"""VideoPose3D lifter throughput (BASELINE configs[4]): N synthetic 17-keypoint frames -> 3-D joints, host to host.
    python tools/bench_lifter.py [N=16384] [reps=5]
Reports frames/s through the C ABI (pe_lift3d: H2D + 10 layer launches + D2H), the algorithmic FLOP rate in the
reference's per-window accounting (176.3 MMAC per output frame) and the executed rate of the dilated
whole-sequence form (16.9 MMAC per frame)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from posepipeline_b200 import engine as E
from posepipeline_b200.synthetic import synthetic_keypoints_2d
from posepipeline_b200.weights import synthetic_videopose3d_state_dict

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
eng = E.PoseEngine(0)
lf = E.Lifter(eng, synthetic_videopose3d_state_dict())
kp = synthetic_keypoints_2d(n, seed=7)[:, :, :2].astype(np.float32)
x = kp / 1920 * 2 - np.array([1, 1080 / 1920], np.float32)
lf.lift(x)
t0 = time.perf_counter()
for _ in range(reps):
    out = lf.lift(x)
dt = (time.perf_counter() - t0) / reps
print(f"lifter: {n} frames in {dt * 1e3:.2f} ms -> {n / dt:.0f} frames/s; reference-accounting {n * 2 * 176.3e6 / dt / 1e12:.1f} TFLOP/s, "
      f"executed (dilated form) {n * 2 * 16.9e6 / dt / 1e12:.2f} TFLOP/s ({'tcgen05' if lf.uses_tensor_cores() else 'fp32 SIMT'})")
