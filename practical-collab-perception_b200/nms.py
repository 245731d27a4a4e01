"""Late-fusion box NMS, drop-in for the reference's call chain
``v2x_late_fusion.py:21-35`` -> ``model_nms_utils.class_agnostic_nms`` (pcdet/models/model_utils/model_nms_utils.py:6-27) ->
``iou3d_nms_utils.nms_gpu`` / ``boxes_iou_bev`` (pcdet/ops/iou3d_nms/iou3d_nms_utils.py:84-99, 27-43).  Same names, argument
meaning and return values; score mask, ordering, suppression mask and the greedy scan all run on the GPU (csrc/nms.cu) - the
reference copies the mask to the host and scans it there."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from .config import cfg_get
from .frontend import _ptr, _require_cuda, _stream, device_guard


@device_guard
def _run(boxes: torch.Tensor, scores: torch.Tensor, iou_thresh: float, score_thresh: Optional[float], pre_max: int,
         post_max: int, normal: bool = False) -> torch.Tensor:
    lib = _lib.load()
    _require_cuda(boxes, "boxes")
    _require_cuda(scores, "scores")
    b = boxes.detach()
    if b.dtype != torch.float32 or b.stride(-1) != 1:
        b = b.float().contiguous()
    s = scores.detach().float().contiguous()
    n = b.shape[0]
    if b.dim() != 2 or b.shape[1] < 7 or s.shape[0] != n:
        raise ValueError(f"boxes must be (N, >= 7) and scores (N,), got {tuple(boxes.shape)} / {tuple(scores.shape)}")
    dev = b.device
    keep = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    count = torch.zeros(1, dtype=torch.int32, device=dev)
    scratch = torch.empty(int(lib.pcp_nms_scratch_bytes(n)) + 256, dtype=torch.uint8, device=dev)
    fn = lib.pcp_nms_normal if normal else lib.pcp_nms_bev
    rc = fn(_ptr(b), b.stride(0) if n else 7, _ptr(s), n, int(score_thresh is not None),
            C.c_float(float(score_thresh) if score_thresh is not None else 0.0), C.c_float(float(iou_thresh)),
            int(pre_max), int(post_max), _ptr(scratch), scratch.numel(), _ptr(keep), _ptr(count), _stream())
    _lib.check(rc, "pcp_nms_normal" if normal else "pcp_nms_bev")
    k = int(count.item())                                       # the one read-back (the reference copies the whole mask)
    if k < 0:
        raise RuntimeError("pcp_nms_bev: more than 4096 candidate boxes and no pre-NMS top-k in [1, 4096] (unsupported)")
    return keep[:k]


def nms_gpu(boxes: torch.Tensor, scores: torch.Tensor, thresh: float, pre_maxsize: Optional[int] = None, **kwargs):
    """iou3d_nms_utils.nms_gpu (iou3d_nms_utils.py:84-99): (N, 7) boxes, (N,) scores -> (indices of the kept boxes in
    descending score order, None)."""
    assert boxes.shape[1] == 7
    return _run(boxes, scores, thresh, None, pre_maxsize or 0, 0), None


def nms_normal_gpu(boxes: torch.Tensor, scores: torch.Tensor, thresh: float, **kwargs):
    """iou3d_nms_utils.nms_normal_gpu (iou3d_nms_utils.py:102-116): NMS on the axis-aligned BEV footprints."""
    assert boxes.shape[1] == 7
    return _run(boxes, scores, thresh, None, 0, 0, normal=True), None


_NMS_TYPES = {"nms_gpu": False, "nms_normal_gpu": True}


def _nms_by_config(box_scores, box_preds, nms_config, score_thresh):
    nms_type = cfg_get(nms_config, "NMS_TYPE", "nms_gpu")
    if nms_type not in _NMS_TYPES:
        raise NotImplementedError(f"NMS_TYPE={nms_type}: 'nms_gpu' and 'nms_normal_gpu' are implemented")
    return _run(box_preds[:, :7], box_scores, nms_config.NMS_THRESH, score_thresh, int(nms_config.NMS_PRE_MAXSIZE),
                int(nms_config.NMS_POST_MAXSIZE), normal=_NMS_TYPES[nms_type])


def class_agnostic_nms(box_scores: torch.Tensor, box_preds: torch.Tensor, nms_config, score_thresh: Optional[float] = None):
    """model_nms_utils.class_agnostic_nms (model_nms_utils.py:6-27) -> (selected indices into the inputs, their scores)."""
    selected = _nms_by_config(box_scores, box_preds, nms_config, score_thresh)
    return selected, box_scores[selected]


def multi_classes_nms(cls_scores: torch.Tensor, box_preds: torch.Tensor, nms_config, score_thresh: Optional[float] = None):
    """model_nms_utils.multi_classes_nms (model_nms_utils.py:28-66): per class k, class_agnostic_nms on cls_scores[:, k];
    -> (scores, labels, boxes) concatenated over the classes."""
    pred_scores, pred_labels, pred_boxes = [], [], []
    for k in range(cls_scores.shape[1]):
        col = cls_scores[:, k].contiguous()
        selected = _nms_by_config(col, box_preds, nms_config, score_thresh)
        pred_scores.append(col[selected])
        pred_labels.append(torch.full((selected.shape[0],), k, dtype=torch.long, device=col.device))
        pred_boxes.append(box_preds[selected])
    return torch.cat(pred_scores, dim=0), torch.cat(pred_labels, dim=0), torch.cat(pred_boxes, dim=0)


@device_guard
def boxes_iou_bev(boxes_a: torch.Tensor, boxes_b: torch.Tensor) -> torch.Tensor:
    """iou3d_nms_utils.boxes_iou_bev (iou3d_nms_utils.py:27-43): (A, 7), (B, 7) -> (A, B)."""
    lib = _lib.load()
    _require_cuda(boxes_a, "boxes_a")
    _require_cuda(boxes_b, "boxes_b")
    assert boxes_a.shape[1] == 7 and boxes_b.shape[1] == 7
    a, b = boxes_a.detach().float().contiguous(), boxes_b.detach().float().contiguous()
    out = torch.zeros((a.shape[0], b.shape[0]), dtype=torch.float32, device=a.device)
    _lib.check(lib.pcp_boxes_iou_bev(_ptr(a), a.shape[0], _ptr(b), b.shape[0], _ptr(out), _stream()), "pcp_boxes_iou_bev")
    return out
