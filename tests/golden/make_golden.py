"""Generates the golden fixtures under tests/golden/ by EXECUTING REFERENCE CODE in this container.

    python tests/golden/make_golden.py          (needs /root/reference; the fixtures are committed)

1. dark_decode.npz  -- ``/root/reference/pose_pipeline/utils/inference.py`` (the reference's in-tree
   DarkPose copy: get_max_preds :27-54, taylor :57-75, gaussian_blur :78-92) is imported as a module
   and run on seeded synthetic heatmaps.  Pins the argmax / 17x17 blur / log / Taylor maths.
2. person_bbox.json -- the body of ``PersonBbox.make`` (``pose_pipeline/pipeline.py:656-687``) is
   extracted from the reference file with ``ast`` and executed against fake tables (datajoint is not
   installable here).  ``fillna(method=...)`` is routed to bfill/ffill because this container's
   pandas 3 rejects the keyword.  Pins bbox/present bit-exactly.

No reference source is copied into the repo: only inputs and outputs are stored.
"""
import ast
import importlib.util
import json
import os
import sys
import textwrap

import numpy as np
import pandas as pd

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def make_heatmaps(rng, N, K, H, W, neg_floor=False):
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    hm = np.zeros((N, K, H, W), np.float64)
    for n in range(N):
        for k in range(K):
            cx, cy = rng.uniform(-2, W + 2), rng.uniform(-2, H + 2)      # some peaks at / over the border
            s = rng.uniform(1.5, 4.0)
            hm[n, k] = rng.uniform(0.2, 1.0) * np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * s * s))
            hm[n, k] += rng.normal(0, 0.004, (H, W))
            if not neg_floor:
                hm[n, k] = np.abs(hm[n, k])
    hm[0, 0] = -np.abs(hm[0, 0])          # max <= 0 case
    return hm.astype(np.float32)


def golden_dark():
    spec = importlib.util.spec_from_file_location("ref_inference", f"{REF}/pose_pipeline/utils/inference.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rng = np.random.default_rng(2024)
    out = {}
    for tag, (N, K, H, W, kernel, neg) in {"a": (2, 6, 96, 72, 17, False), "b": (1, 5, 64, 48, 11, True)}.items():
        hm = make_heatmaps(rng, N, K, H, W, neg)
        coords, maxvals = ref.get_max_preds(hm.copy())
        blurred = ref.gaussian_blur(hm.copy(), kernel)
        logged = np.log(np.maximum(blurred, 1e-10))
        refined = coords.copy()
        for n in range(N):
            for k in range(K):
                refined[n, k] = ref.taylor(logged[n][k], refined[n][k])
        out[f"{tag}_heatmaps"] = hm
        out[f"{tag}_kernel"] = np.int64(kernel)
        out[f"{tag}_argmax"] = coords
        out[f"{tag}_maxvals"] = maxvals
        out[f"{tag}_blurred"] = blurred.astype(np.float32)
        out[f"{tag}_refined"] = refined
    np.savez_compressed(os.path.join(HERE, "dark_decode.npz"), **out)
    print("dark_decode.npz", {k: v.shape for k, v in out.items()})


class _Frame(pd.DataFrame):
    @property
    def _constructor(self):
        return _Frame

    def fillna(self, value=None, *, method=None, axis=None, limit=None, **kw):
        if method == "bfill":
            return self.bfill(axis=axis, limit=limit)
        if method == "ffill":
            return self.ffill(axis=axis, limit=limit)
        return super().fillna(value, axis=axis, limit=limit, **kw)


class _PdShim:
    DataFrame = _Frame


def _reference_person_bbox_make():
    src = open(f"{REF}/pose_pipeline/pipeline.py").read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == "PersonBbox":
            for f in node.body:
                if isinstance(f, ast.FunctionDef) and f.name == "make":
                    return textwrap.dedent(ast.get_source_segment(src, f))
    raise RuntimeError("PersonBbox.make not found")


class _FakeTable:
    def __init__(self, row):
        self.row = row

    def __and__(self, key):
        return self

    def fetch1(self, attr):
        return self.row[attr]


def synth_tracks(rng, n_frames, ids, p_drop, p_dup):
    tracks = []
    for f in range(n_frames):
        frame = []
        for tid in ids:
            if rng.random() < p_drop:
                continue
            x, y, w, h = rng.uniform(0, 1500), rng.uniform(0, 400), rng.uniform(100, 400), rng.uniform(300, 680)
            x, y, w, h = [float(np.float32(v)) for v in (x, y, w, h)]
            frame.append({"track_id": int(tid), "tlbr": [x, y, x + w, y + h], "tlhw": [x, y, w, h], "confidence": float(rng.random())})
            if rng.random() < p_dup:
                frame.append(dict(frame[-1], track_id=int(ids[(ids.index(tid) + 1) % len(ids)])))
        tracks.append(frame)
    return tracks


def golden_person_bbox():
    code = _reference_person_bbox_make()
    ns = {"np": np, "pd": _PdShim}
    exec(code, ns)
    make = ns["make"]
    rng = np.random.default_rng(99)
    cases = []
    specs = [(40, [1], [1], 0.3, 0.0), (60, [1, 2, 5], [2], 0.25, 0.1), (50, [3, 4], [3, 4], 0.4, 0.0),
             (12, [7], [7], 0.0, 0.0), (30, [1, 2], [9], 0.1, 0.0), (25, [1], [1], 0.8, 0.0), (1, [1], [1], 0.0, 0.0)]
    for n, ids, keep, p_drop, p_dup in specs:
        tracks = synth_tracks(rng, n, ids, p_drop, p_dup)

        class Self:
            def insert1(self, key):
                self.key = key
        ns["TrackingBbox"] = _FakeTable({"tracks": tracks})
        ns["PersonBboxValid"] = _FakeTable({"keep_tracks": keep})
        s = Self()
        make(s, {})
        bbox = np.asarray(s.key["bbox"], np.float64)
        present = np.asarray(s.key["present"], bool)
        cases.append({"tracks": tracks, "keep_tracks": keep,
                      "bbox_hex": [[float(v).hex() for v in row] for row in bbox],
                      "present": present.tolist()})
    json.dump(cases, open(os.path.join(HERE, "person_bbox.json"), "w"))
    print("person_bbox.json", len(cases), "cases")


if __name__ == "__main__":
    golden_dark()
    golden_person_bbox()
