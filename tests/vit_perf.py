"""Per-op-kind device time of the ViTPose-B forward (CUDA events around every op), grouped by (kind, shape).
    python tests/vit_perf.py [max_crops] [reps]"""
import os, sys, collections
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from posepipeline_b200 import engine as E
from posepipeline_b200.vit_spec import build_vitpose_program
from posepipeline_b200.weights import synthetic_vitpose_state_dict

mc = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
spec = E.METHODS["ViTPose_B_COCO"]
prog = build_vitpose_program()
eng = E.PoseEngine(0)
m = E.TopDownModel(eng, synthetic_vitpose_state_dict(prog, 0), spec, max_crops=mc)
crops = np.random.default_rng(0).integers(0, 256, (mc, spec.image_size[1], spec.image_size[0], 3), dtype=np.uint8)
m.forward_heatmaps(crops)
m.profile(2)
for _ in range(reps):
    m.forward_heatmaps(crops)
ms = m.profile_ops() / reps
agg = collections.OrderedDict()
for op, t in zip(prog.ops, ms):
    to = prog.tensors[op.out]
    key = (op.kind, op.cin, op.cout, op.ksize, to.H, to.W, int(op.residual >= 0))
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1; a[1] += t
tot = ms.sum()
print(f"forward of {2 * mc} images: {tot:.2f} ms -> {mc / tot * 1e3:.1f} crops/s (network only)")
print("kind cin cout k HxW res n ms share")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(k, n, f"{t:.3f} {100 * t / tot:.1f}%")
