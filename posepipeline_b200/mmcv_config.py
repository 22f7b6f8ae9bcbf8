"""Restricted reader for the mmcv-style Python config files the reference hands to ``init_pose_model``
(``pose_pipeline/wrappers/mmpose.py:33-52`` -> ``$MODEL_DATA_DIR/mmpose/config/**``), so that a user-edited config is honoured
exactly where the reference would honour it (SURVEY §2 row 9, §5 "config"): ``model.test_cfg`` (flip_test, post_process,
shift_heatmap, modulate_kernel), ``data_cfg`` (image / heatmap size, joints), the test pipeline (normalisation, bbox padding)
and the top-level ``dataset_info`` (flip pairs from the ``swap`` fields; absent -> mmpose falls back to the COCO-17 pairs,
quirk Q3).

mmcv's ``Config.fromfile`` executes the file as Python, merges the ``_base_`` files (child keys win, ``_delete_=True`` replaces
a dict) and substitutes ``{{_base_.name}}`` references.  This loader does the same with a restricted namespace: no imports, no
attribute access to anything but the base-config view, only literals / dict() / list() / arithmetic -- enough for every
file under ``3rdparty/mmpose/config`` and ``3rdparty/mmtracking``.
"""
from __future__ import annotations

import ast
import os
import re
from typing import Any, Dict, List, Optional, Tuple

_SAFE_BUILTINS = {"dict": dict, "list": list, "tuple": tuple, "set": set, "range": range, "len": len, "int": int, "float": float,
                  "str": str, "bool": bool, "min": min, "max": max, "sum": sum, "abs": abs, "round": round, "True": True,
                  "False": False, "None": None}
_BASE_REF = re.compile(r"\{\{\s*_base_\.([\w.]+)\s*\}\}")


class _View(dict):
    """attribute access into the merged base config: {{_base_.dataset_info}} -> _base_cfg_.dataset_info"""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError:
            raise AttributeError(k)
        return _View(v) if isinstance(v, dict) else v


def _check_ast(tree: ast.AST, path: str):
    for node in ast.walk(tree):
        if isinstance(node, (ast.Import, ast.ImportFrom, ast.Lambda, ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef, ast.With,
                             ast.Try, ast.While, ast.Global, ast.Nonlocal, ast.Await, ast.Yield, ast.YieldFrom)):
            raise ValueError(f"{path}: config files may only contain assignments of literals / dict() / list() expressions "
                             f"({type(node).__name__} at line {getattr(node, 'lineno', '?')})")
        if isinstance(node, ast.Attribute) and not (isinstance(node.value, (ast.Name, ast.Attribute))):
            raise ValueError(f"{path}: attribute access on an expression is not allowed (line {node.lineno})")
        if isinstance(node, ast.Name) and node.id.startswith("__"):
            raise ValueError(f"{path}: dunder names are not allowed (line {node.lineno})")


def _merge(base: Dict[str, Any], child: Dict[str, Any]) -> Dict[str, Any]:
    out = dict(base)
    for k, v in child.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict) and not v.get("_delete_", False):
            out[k] = _merge(out[k], v)
        else:
            out[k] = {kk: vv for kk, vv in v.items() if kk != "_delete_"} if isinstance(v, dict) else v
    return out


def load_config(path: str, _depth: int = 0) -> Dict[str, Any]:
    """-> dict of the config's top-level variables with ``_base_`` files merged in."""
    if _depth > 8:
        raise ValueError(f"{path}: _base_ nesting too deep")
    text = open(path).read()
    # bases first (their values are what {{_base_.x}} refers to)
    tree0 = ast.parse(text.replace("{{", "(").replace("}}", ")"), path)
    bases: List[str] = []
    for node in tree0.body:
        if isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name) and node.targets[0].id == "_base_":
            v = ast.literal_eval(node.value)
            bases = [v] if isinstance(v, str) else list(v)
    base_cfg: Dict[str, Any] = {}
    for b in bases:
        base_cfg = _merge(base_cfg, load_config(os.path.join(os.path.dirname(path), b), _depth + 1))
    src = _BASE_REF.sub(lambda m: f"_base_cfg_.{m.group(1)}", text)
    tree = ast.parse(src, path)
    _check_ast(tree, path)
    ns: Dict[str, Any] = {"__builtins__": _SAFE_BUILTINS, "_base_cfg_": _View(base_cfg)}
    exec(compile(tree, path, "exec"), ns)                          # noqa: S102 -- restricted namespace, AST vetted above
    own = {k: (dict(v) if isinstance(v, _View) else v) for k, v in ns.items() if not k.startswith("_")}
    return _merge(base_cfg, own)


# ------------------------------------------------------------------------------------------ what the engine needs
def flip_pairs_from_dataset_info(dataset_info: Dict[str, Any]) -> List[List[int]]:
    """mmpose ``DatasetInfo``: pairs (id, id of the keypoint named by ``swap``), each once, ordered by the smaller id."""
    kinfo = dataset_info["keypoint_info"]
    name2id = {v["name"]: int(v.get("id", k)) for k, v in kinfo.items()}
    pairs = set()
    for k, v in kinfo.items():
        sw = v.get("swap", "")
        if sw:
            a, b = int(v.get("id", k)), name2id[sw]
            pairs.add((min(a, b), max(a, b)))
    return [list(p) for p in sorted(pairs)]


COCO_FALLBACK_PAIRS = [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [11, 12], [13, 14], [15, 16]]


def topdown_settings(cfg: Dict[str, Any]) -> Dict[str, Any]:
    """The fields of a top-down pose config that change the arithmetic of the path (-> keyword arguments of engine.TopDownSpec)."""
    model, data_cfg = cfg["model"], cfg["data_cfg"]
    test_cfg = model.get("test_cfg", {})
    backbone = model["backbone"]
    if backbone.get("type") != "HRNet":
        raise NotImplementedError(f"backbone {backbone.get('type')!r} is not built in this engine (HRNet is)")
    c0 = tuple(backbone["extra"]["stage2"]["num_channels"])[0]
    variant = {48: "w48", 32: "w32"}.get(c0)
    if variant is None:
        raise NotImplementedError(f"HRNet width {c0} is not built (w32 / w48 are)")
    head = model["keypoint_head"]
    if head.get("num_deconv_layers", 3) != 0 or head.get("extra", {}).get("final_conv_kernel", 1) != 1:
        raise NotImplementedError("only the HRNet head (no deconv layers, 1x1 final conv) is built")
    pipeline = cfg.get("test_pipeline") or cfg.get("val_pipeline") or []
    norm = next((s for s in pipeline if s.get("type") == "NormalizeTensor"), {})
    cs = next((s for s in pipeline if s.get("type") == "TopDownGetBboxCenterScale"), {})
    if any(s.get("type") == "TopDownAffine" and s.get("use_udp") for s in pipeline) or test_cfg.get("use_udp"):
        raise NotImplementedError("UDP affine / decode is not built for the HRNet methods")
    info = cfg.get("dataset_info")
    return dict(variant=variant, image_size=tuple(data_cfg["image_size"]), heatmap_size=tuple(data_cfg["heatmap_size"]),
                num_joints=int(head["out_channels"]), flip_test=bool(test_cfg.get("flip_test", True)),
                post_process=test_cfg.get("post_process", "default"), shift_heatmap=bool(test_cfg.get("shift_heatmap", True)),
                modulate_kernel=int(test_cfg.get("modulate_kernel", 11)), padding=float(cs.get("padding", 1.25)),
                flip_pairs=flip_pairs_from_dataset_info(info) if info else [list(p) for p in COCO_FALLBACK_PAIRS],
                mean=tuple(norm.get("mean", (0.485, 0.456, 0.406))), std=tuple(norm.get("std", (0.229, 0.224, 0.225))))


def bytetrack_settings(cfg: Dict[str, Any]) -> Dict[str, Any]:
    """Detector / tracker thresholds of an mmtracking ByteTrack config (3rdparty/mmtracking/mot/bytetrack/*.py)."""
    m = cfg["model"]
    det, trk = m["detector"], m["tracker"]
    test_cfg = det.get("test_cfg", {})
    return dict(img_scale=tuple(det.get("input_size", (800, 1440))), num_classes=int(det["bbox_head"]["num_classes"]),
                score_thr=float(test_cfg.get("score_thr", 0.01)), nms_iou=float(test_cfg.get("nms", {}).get("iou_threshold", 0.65)),
                tracker=dict(obj_score_high=float(trk["obj_score_thrs"]["high"]), obj_score_low=float(trk["obj_score_thrs"]["low"]),
                             init_track_thr=float(trk["init_track_thr"]), match_iou_high=float(trk["match_iou_thrs"]["high"]),
                             match_iou_low=float(trk["match_iou_thrs"]["low"]), match_iou_tentative=float(trk["match_iou_thrs"]["tentative"]),
                             weight_iou_with_det_scores=bool(trk["weight_iou_with_det_scores"]), num_tentatives=int(trk.get("num_tentatives", 3)),
                             num_frames_retain=int(trk["num_frames_retain"])))
