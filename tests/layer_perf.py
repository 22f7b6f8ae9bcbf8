"""Per-layer-shape device time of the HRNet-W48 forward (CUDA events around every op), grouped by shape.
    python tests/layer_perf.py [max_crops] [reps]"""
import os, sys, collections
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from posepipeline_b200 import engine as E
from posepipeline_b200.hrnet_spec import build_program
from posepipeline_b200.weights import synthetic_hrnet_state_dict

mc = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
spec = E.METHODS["HRNet_W48_COCO"]
prog = build_program("w48")
sd = synthetic_hrnet_state_dict(prog, 0)
eng = E.PoseEngine(0)
m = E.TopDownModel(eng, sd, spec, max_crops=mc, use_tensor_cores=os.environ.get("PE_NO_TC") != "1")
crops = np.random.default_rng(0).integers(0, 256, (mc, 384, 288, 3), dtype=np.uint8)
m.forward_heatmaps(crops)
m.profile(2)
for _ in range(reps):
    m.forward_heatmaps(crops)
ms = m.profile_ops() / reps
nimg = 2 * mc
agg = collections.OrderedDict()
for op, t in zip(prog.ops, ms):
    to = prog.tensors[op.out]
    key = (["stem", "conv", "fuse", "head"][op.kind], op.cin, op.cout, op.ksize, op.stride, to.H, to.W, int(op.residual >= 0))
    a = agg.setdefault(key, [0, 0.0, 0.0])
    a[0] += 1; a[1] += t
    if op.kind != 2:
        a[2] += 2.0 * to.H * to.W * op.cout * op.cin * op.ksize ** 2 * nimg
tot = ms.sum()
print(f"forward of {nimg} images: {tot:.2f} ms  -> {mc / tot * 1e3:.1f} crops/s (network only)")
print(f"{'kind':5} {'cin':>4} {'cout':>4} k s {'HxW':>7} res {'n':>3} {'ms':>8} {'share':>6} {'TFLOP/s':>8}")
for k, (n, t, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[0]:5} {k[1]:4d} {k[2]:4d} {k[3]} {k[4]} {k[5]:3d}x{k[6]:<3d} {k[7]:3d} {n:3d} {t:8.3f} {100*t/tot:5.1f}% {fl/t/1e9 if t else 0:8.1f}")
