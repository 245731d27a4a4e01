"""CPU restatement of the late-fusion box NMS.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows pcdet/models/model_utils/model_nms_utils.py:6-27 (class_agnostic_nms: score mask, top-k, NMS, post max size),
pcdet/ops/iou3d_nms/iou3d_nms_utils.py:84-99 (nms_gpu: descending score order) and the greedy scan of
pcdet/ops/iou3d_nms/src/iou3d_nms.cpp:116-131.  The BEV IoU (iou3d_nms_kernel.cu:104-234: overlap of two rotated
rectangles / union) is evaluated in float64 by polygon clipping - an independent evaluation of the same quantity.
Pinning: the reference's arithmetic for this op is a CUDA kernel, so it cannot run in the GPU-less build container;
``oracle/Makefile`` compiles that kernel unmodified into ``oracle/_ref/libiou3d_ref.so`` and
``tests/test_gpu_nms.py`` checks this oracle AND the product kernels against it on the GPU box (IoU matrix, keep lists).
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np


def _corners(box) -> np.ndarray:
    x, y, dx, dy, a = float(box[0]), float(box[1]), float(box[3]), float(box[4]), float(box[6])
    c, s = np.cos(a), np.sin(a)
    local = np.array([[-dx / 2, -dy / 2], [dx / 2, -dy / 2], [dx / 2, dy / 2], [-dx / 2, dy / 2]])
    rot = np.array([[c, -s], [s, c]])
    return local @ rot.T + np.array([x, y])


def _clip(poly: np.ndarray, p0: np.ndarray, p1: np.ndarray) -> np.ndarray:
    """Keep the part of `poly` on the left of the directed line p0 -> p1."""
    out = []
    n = len(poly)
    d = p1 - p0
    side = lambda p: d[0] * (p[1] - p0[1]) - d[1] * (p[0] - p0[0])
    for i in range(n):
        a, b = poly[i], poly[(i + 1) % n]
        sa, sb = side(a), side(b)
        if sa >= 0:
            out.append(a)
        if (sa > 0 and sb < 0) or (sa < 0 and sb > 0):
            t = sa / (sa - sb)
            out.append(a + t * (b - a))
    return np.array(out) if out else np.zeros((0, 2))


def bev_overlap(box_a, box_b) -> float:
    poly = _corners(box_a)
    cb = _corners(box_b)                      # counter-clockwise
    for k in range(4):
        if len(poly) < 3:
            return 0.0
        poly = _clip(poly, cb[k], cb[(k + 1) % 4])
    if len(poly) < 3:
        return 0.0
    x, y = poly[:, 0], poly[:, 1]
    return 0.5 * abs(float(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1))))


def boxes_iou_bev(boxes_a: np.ndarray, boxes_b: np.ndarray) -> np.ndarray:
    """iou3d_nms_kernel.cu:227-234 for every pair -> (A, B) float64."""
    out = np.zeros((len(boxes_a), len(boxes_b)))
    for i, a in enumerate(boxes_a):
        sa = float(a[3]) * float(a[4])
        for j, b in enumerate(boxes_b):
            so = bev_overlap(a, b)
            out[i, j] = so / max(sa + float(b[3]) * float(b[4]) - so, 1e-8)
    return out


def nms(boxes: np.ndarray, scores: np.ndarray, thresh: float, iou: Optional[np.ndarray] = None) -> np.ndarray:
    """nms_gpu: indices of the kept boxes, highest score first (ties: lower index first); iou3d_nms.cpp:116-131."""
    order = np.lexsort((np.arange(len(scores)), -scores.astype(np.float64)))
    if iou is None:
        iou = boxes_iou_bev(boxes, boxes)
    removed = np.zeros(len(order), dtype=bool)
    keep = []
    for a, i in enumerate(order):
        if removed[a]:
            continue
        keep.append(i)
        for b in range(a + 1, len(order)):
            if iou[i, order[b]] > thresh:
                removed[b] = True
    return np.asarray(keep, dtype=np.int64)


def class_agnostic_nms(box_scores: np.ndarray, box_preds: np.ndarray, nms_thresh: float, pre_max: int, post_max: int,
                       score_thresh: Optional[float] = None, iou: Optional[np.ndarray] = None) -> np.ndarray:
    """model_nms_utils.py:6-27 -> selected indices into the inputs."""
    idx = np.arange(len(box_scores))
    if score_thresh is not None:
        idx = idx[box_scores >= score_thresh]                               # :9
    if len(idx) == 0:
        return np.zeros(0, dtype=np.int64)
    s = box_scores[idx]
    top = np.lexsort((idx, -s.astype(np.float64)))[:min(pre_max, len(idx))]  # torch.topk (:15), ties by index
    cand = idx[top]
    sub_iou = None if iou is None else iou[np.ix_(cand, cand)]
    keep = nms(box_preds[cand, :7], box_scores[cand], nms_thresh, sub_iou)
    return cand[keep[:post_max]]
