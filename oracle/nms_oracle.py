"""ORACLE (test infrastructure only) - class-aware NMS exactly as torchvision.ops.boxes.batched_nms (0.26.0).

The reference calls `boxes.batched_nms(det[:, :4], det[:, 4] * det[:, 5], det[:, 6], nms_thres)`
(yolox-drone/models/core/utils_bbox.py:414-419).  torchvision is a third-party dependency that is not vendored
under /root/reference and has no version pin in yolox-drone; the installed 0.26.0 is the pinned oracle.
Restated here from its published algorithm:
  * batched_nms dispatch (torchvision/ops/boxes.py): coordinate trick when boxes.numel() <= 4000 on CPU
    (<= 100000 on CUDA), otherwise one nms() per class;
  * coordinate trick: offsets = idxs * (boxes.max() + 1) in float32, nms(boxes + offsets[:, None]);
  * per-class ("vanilla"): nms() inside every class, kept indices re-sorted by score, descending;
  * nms(): oracle/nms_oracle.c.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_build" / "libnms_oracle.so"
_lib = None


def build(force: bool = False) -> Path:
    src = _HERE / "nms_oracle.c"
    if force or not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
        _SO.parent.mkdir(exist_ok=True)
        subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", str(src), "-o",
                        str(_SO)], check=True)
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(os.fspath(_SO))
        _lib.oracle_nms.restype = ctypes.c_int
        _lib.oracle_nms.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_void_p]
    return _lib


def nms(boxes: np.ndarray, scores: np.ndarray, iou_threshold: float) -> np.ndarray:
    boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 4)
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    n = boxes.shape[0]
    keep = np.empty(n, dtype=np.int32)
    k = _load().oracle_nms(boxes.ctypes.data, scores.ctypes.data, n, ctypes.c_float(iou_threshold), keep.ctypes.data)
    if k < 0:
        raise MemoryError
    return keep[:k].astype(np.int64)


def nms_python(boxes, scores, iou_threshold):
    """Same algorithm in pure Python/numpy scalars (small cases only; cross-checks the C build flags)."""
    boxes = np.asarray(boxes, dtype=np.float32)
    scores = np.asarray(scores, dtype=np.float32)
    order = np.argsort(-scores, kind="stable")
    areas = ((boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])).astype(np.float32)
    sup = np.zeros(len(boxes), dtype=bool)
    keep = []
    thr = np.float32(iou_threshold)
    for a, i in enumerate(order):
        if sup[i]:
            continue
        keep.append(i)
        for j in order[a + 1:]:
            if sup[j]:
                continue
            w = np.float32(min(boxes[i, 2], boxes[j, 2]) - max(boxes[i, 0], boxes[j, 0]))
            h = np.float32(min(boxes[i, 3], boxes[j, 3]) - max(boxes[i, 1], boxes[j, 1]))
            w = w if w > 0 else np.float32(0)
            h = h if h > 0 else np.float32(0)
            inter = np.float32(w * h)
            with np.errstate(divide="ignore", invalid="ignore"):
                ovr = inter / np.float32(np.float32(areas[i] + areas[j]) - inter)
            if ovr > thr:
                sup[j] = True
    return np.asarray(keep, dtype=np.int64)


def resolve_strategy(strategy: str, k: int) -> str:
    """Which torchvision branch applies to K boxes: 'trick' or 'per_class'."""
    if strategy in ("trick", "per_class"):
        return strategy
    if strategy == "auto_cpu":
        return "trick" if 4 * k <= 4000 else "per_class"
    if strategy == "auto_cuda":
        return "trick" if 4 * k <= 100_000 else "per_class"
    raise ValueError(strategy)


def batched_nms(boxes, scores, idxs, iou_threshold: float, strategy: str = "auto_cpu") -> np.ndarray:
    boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 4)
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    idxs = np.ascontiguousarray(idxs, dtype=np.float32)
    k = boxes.shape[0]
    if k == 0:
        return np.empty((0,), dtype=np.int64)
    if resolve_strategy(strategy, k) == "trick":
        max_coordinate = boxes.max()
        offsets = (idxs * np.float32(max_coordinate + np.float32(1))).astype(np.float32)
        return nms((boxes + offsets[:, None]).astype(np.float32), scores, iou_threshold)
    keep_mask = np.zeros(k, dtype=bool)
    for c in np.unique(idxs):
        cur = np.nonzero(idxs == c)[0]
        keep_mask[cur[nms(boxes[cur], scores[cur], iou_threshold)]] = True
    keep = np.nonzero(keep_mask)[0]
    return keep[np.argsort(-scores[keep], kind="stable")]
