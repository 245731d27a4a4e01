"""CPU restatement of the late-fusion box NMS.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows pcdet/models/model_utils/model_nms_utils.py:6-27 (class_agnostic_nms: score mask, top-k, NMS, post max size),
pcdet/ops/iou3d_nms/iou3d_nms_utils.py:84-99 (nms_gpu: descending score order) and the greedy scan of
pcdet/ops/iou3d_nms/src/iou3d_nms.cpp:116-131.  The BEV IoU (iou3d_nms_kernel.cu:104-234: overlap of two rotated
rectangles / union) is evaluated twice: ``bev_overlap`` in float64 by polygon clipping - an independent, exact evaluation of
the same quantity - and ``ref_overlap_f32``, a float32 restatement of the reference kernel's own procedure (edge-pair
intersections, corner test with MARGIN 1e-2, atan2 ordering, fan area), which reproduces the kernel's approximation error
(the kernel itself is matched to rounding, ~1e-6; only FMA contraction differs).
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


_F = np.float32


def _cross3(p1, p2, p0):
    return _F((p1[0] - p0[0]) * (p2[1] - p0[1]) - (p2[0] - p0[0]) * (p1[1] - p0[1]))                  # :39-41


def _in_box2d(box, p) -> bool:
    """iou3d_nms_kernel.cu:51-62."""
    margin = _F(1e-2)
    c, s = np.cos(_F(-box[6])), np.sin(_F(-box[6]))
    rx = _F((p[0] - box[0]) * c + (p[1] - box[1]) * (-s))
    ry = _F((p[0] - box[0]) * s + (p[1] - box[1]) * c)
    return bool(abs(rx) < box[3] / _F(2) + margin and abs(ry) < box[4] / _F(2) + margin)


def _intersection(p1, p0, q1, q0):
    """iou3d_nms_kernel.cu:64-95 -> point or None."""
    if not (min(p0[0], p1[0]) <= max(q0[0], q1[0]) and min(q0[0], q1[0]) <= max(p0[0], p1[0]) and
            min(p0[1], p1[1]) <= max(q0[1], q1[1]) and min(q0[1], q1[1]) <= max(p0[1], p1[1])):
        return None
    s1, s2, s3, s4 = _cross3(q0, p1, p0), _cross3(p1, q1, p0), _cross3(p0, q1, q0), _cross3(q1, p1, q0)
    if not (s1 * s2 > 0 and s3 * s4 > 0):
        return None
    s5 = _cross3(q1, p1, p0)
    if abs(_F(s5 - s1)) > _F(1e-8):
        return np.array([(s5 * q0[0] - s1 * q1[0]) / (s5 - s1), (s5 * q0[1] - s1 * q1[1]) / (s5 - s1)], dtype=_F)
    a0, b0, c0 = p0[1] - p1[1], p1[0] - p0[0], p0[0] * p1[1] - p1[0] * p0[1]
    a1, b1, c1 = q0[1] - q1[1], q1[0] - q0[0], q0[0] * q1[1] - q1[0] * q0[1]
    d = a0 * b1 - a1 * b0
    with np.errstate(all="ignore"):
        return np.array([(b0 * c1 - b1 * c0) / d, (a1 * c0 - a0 * c1) / d], dtype=_F)


def _rot_corners_f32(box) -> np.ndarray:
    hx, hy = box[3] / _F(2), box[4] / _F(2)
    pts = np.array([[box[0] - hx, box[1] - hy], [box[0] + hx, box[1] - hy], [box[0] + hx, box[1] + hy],
                    [box[0] - hx, box[1] + hy]], dtype=_F)
    c, s = np.cos(_F(box[6])), np.sin(_F(box[6]))
    out = np.empty((5, 2), dtype=_F)
    for k in range(4):                                                                                  # :97-101
        dx, dy = pts[k, 0] - box[0], pts[k, 1] - box[1]
        out[k, 0] = dx * c + dy * (-s) + box[0]
        out[k, 1] = dx * s + dy * c + box[1]
    out[4] = out[0]
    return out


def ref_overlap_f32(box_a, box_b) -> float:
    """box_overlap of iou3d_nms_kernel.cu:107-225 in float32 (numpy rounds every operation; the kernel fuses some into FMAs)."""
    a, b = np.asarray(box_a, dtype=_F), np.asarray(box_b, dtype=_F)
    ca, cb = _rot_corners_f32(a), _rot_corners_f32(b)
    pts = []
    for i in range(4):
        for j in range(4):
            p = _intersection(ca[i + 1], ca[i], cb[j + 1], cb[j])
            if p is not None:
                pts.append(p)
    for k in range(4):
        if _in_box2d(a, cb[k]):
            pts.append(cb[k].copy())
        if _in_box2d(b, ca[k]):
            pts.append(ca[k].copy())
    n = len(pts)
    if n == 0:
        return 0.0
    centre = np.zeros(2, dtype=_F)
    for p in pts:
        centre = (centre + p).astype(_F)
    centre = (centre / _F(n)).astype(_F)
    ang = [np.arctan2(_F(p[1] - centre[1]), _F(p[0] - centre[0])) for p in pts]
    for j in range(n - 1):                                                                              # :198-207 bubble sort
        for i in range(n - j - 1):
            if ang[i] > ang[i + 1]:
                pts[i], pts[i + 1] = pts[i + 1], pts[i]
                ang[i], ang[i + 1] = ang[i + 1], ang[i]
    area = _F(0)
    for k in range(n - 1):
        u, v = pts[k] - pts[0], pts[k + 1] - pts[0]
        area = _F(area + _F(u[0] * v[1] - u[1] * v[0]))
    return float(abs(area)) / 2.0


def ref_iou_f32(boxes_a: np.ndarray, boxes_b: np.ndarray) -> np.ndarray:
    """iou_bev (:227-234) with the reference's overlap procedure -> (A, B) float32."""
    out = np.zeros((len(boxes_a), len(boxes_b)), dtype=_F)
    for i, a in enumerate(boxes_a):
        for j, b in enumerate(boxes_b):
            so = _F(ref_overlap_f32(a, b))
            out[i, j] = so / max(_F(_F(a[3]) * _F(a[4]) + _F(b[3]) * _F(b[4]) - so), _F(1e-8))
    return out


def iou_normal_f32(boxes_a: np.ndarray, boxes_b: np.ndarray) -> np.ndarray:
    """iou_normal (iou3d_nms_kernel.cu:316-327): axis-aligned footprints, float32."""
    a, b = np.asarray(boxes_a, dtype=_F), np.asarray(boxes_b, dtype=_F)
    left = np.maximum((a[:, 0] - a[:, 3] / _F(2))[:, None], (b[:, 0] - b[:, 3] / _F(2))[None])
    right = np.minimum((a[:, 0] + a[:, 3] / _F(2))[:, None], (b[:, 0] + b[:, 3] / _F(2))[None])
    top = np.maximum((a[:, 1] - a[:, 4] / _F(2))[:, None], (b[:, 1] - b[:, 4] / _F(2))[None])
    bottom = np.minimum((a[:, 1] + a[:, 4] / _F(2))[:, None], (b[:, 1] + b[:, 4] / _F(2))[None])
    inter = np.maximum(right - left, _F(0)) * np.maximum(bottom - top, _F(0))
    sa, sb = (a[:, 3] * a[:, 4])[:, None], (b[:, 3] * b[:, 4])[None]
    return (inter / np.maximum(sa + sb - inter, _F(1e-8))).astype(_F)


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
