"""Point <-> BEV image exchange of the HunterJr correction head, drop-in for the two hot functions of
pcdet/models/bev_layers/hunter_toolbox.py (``bev_scatter`` :65-96, ``interpolate_points_feat_from_bev_img`` :99-131;
called every forward by pcdet/models/bev_layers/hunter_jr.py:268-279,300).  Same names, argument meaning and return
values; the arithmetic runs in libpcp_b200.so (csrc/bev_points.cu).  Inference only (no autograd), CUDA tensors only.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib
from .frontend import _ptr, _require_cuda, _stream, device_guard


def _no_autograd(name: str, *tensors) -> None:
    """These kernels have no backward: fail loudly instead of silently cutting the graph of a training caller
    (hunter_jr.py trains through both functions)."""
    if torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors):
        raise RuntimeError(f"pcp_b200.{name} is inference-only (no autograd): call it under torch.no_grad() "
                           "or detach its inputs")


def _f32(v) -> float:
    """A Python float holding the fp32 value the reference's CUDA tensor element would hold."""
    if isinstance(v, torch.Tensor):
        return float(v.detach().to(dtype=torch.float32, device="cpu").item())
    return float(np.float32(v))


@device_guard
def bev_scatter(points_bev_coord: torch.Tensor, points_batch_idx: torch.Tensor, points_feat: torch.Tensor,
                bev_img_size: Sequence[int], batch_size: Optional[int] = None) -> torch.Tensor:
    """hunter_toolbox.py:65-96.  (N, 2) bev_x/bev_y, (N,) batch index, (N, C) features, (height, width) ->
    (B, C, H, W) per-pixel mean of the features of the points that fall strictly inside the image, zeros elsewhere.
    ``batch_size`` skips the reference's ``torch.max(points_batch_idx).item() + 1`` synchronisation (:74)."""
    lib = _lib.load()
    _no_autograd("bev_scatter", points_bev_coord, points_feat)
    for t, name in ((points_bev_coord, "points_bev_coord"), (points_batch_idx, "points_batch_idx"), (points_feat, "points_feat")):
        _require_cuda(t, name)
    dev = points_feat.device
    coord = points_bev_coord.detach()
    if coord.dtype != torch.float32 or coord.stride(-1) != 1:
        coord = coord.float().contiguous()
    feat = points_feat.detach()
    if feat.dtype != torch.float32 or feat.stride(-1) != 1:
        feat = feat.float().contiguous()
    bidx = points_batch_idx.detach().to(torch.int64).contiguous()
    n, c = feat.shape
    height, width = int(bev_img_size[0]), int(bev_img_size[1])
    if batch_size is None:
        if n == 0:
            raise RuntimeError("bev_scatter: max() of an empty points_batch_idx (the reference fails here too, :74)")
        nf = torch.empty(1, dtype=torch.int32, device=dev)
        _lib.check(lib.pcp_max_index_i64(_ptr(bidx), n, _ptr(nf), _stream()), "pcp_max_index_i64")
        batch_size = int(nf.item())
    batch_size = int(batch_size)
    nbytes = int(lib.pcp_workspace_bytes(n, batch_size, height, width))
    ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
    cap = max(1, min(n, batch_size * height * width))
    cell_mean = torch.empty((cap, c), dtype=torch.float32, device=dev)
    out = torch.empty((batch_size, c, height, width), dtype=torch.float32, device=dev)
    counts = torch.empty(_lib.PCP_COUNTS_LEN, dtype=torch.int32, device=dev)
    rc = lib.pcp_bev_scatter_mean(_ptr(coord), coord.stride(0), _ptr(bidx), _ptr(feat), feat.stride(0), c, n, batch_size,
                                  height, width, _ptr(ws), ws.numel(), _ptr(cell_mean), _ptr(out), _ptr(counts), _stream())
    _lib.check(rc, "pcp_bev_scatter_mean")
    return out


@device_guard
def interpolate_points_feat_from_bev_img(bev_img: torch.Tensor, points: torch.Tensor,
                                         point_cloud_range: Union[torch.Tensor, Sequence[float]],
                                         bev_pixel_size: Union[torch.Tensor, Sequence[float]],
                                         return_bev_coord: bool = False):
    """hunter_toolbox.py:99-131.  (B, C, H, W) image, (N, 1 + ...) points [batch_idx, x, y, ...] -> (N, C) bilinear
    features (and the (N, 2) BEV coordinates when ``return_bev_coord``)."""
    lib = _lib.load()
    _no_autograd("interpolate_points_feat_from_bev_img", bev_img, points)
    _require_cuda(bev_img, "bev_img")
    _require_cuda(points, "points")
    img = bev_img.detach()
    if img.dtype != torch.float32 or not img.is_contiguous():
        img = img.float().contiguous()
    pts = points.detach()
    if pts.dtype != torch.float32 or pts.stride(-1) != 1:
        pts = pts.float().contiguous()
    b, c, h, w = img.shape
    n = pts.shape[0]
    dev = img.device
    feat = torch.empty((n, c), dtype=torch.float32, device=dev)
    coord = torch.empty((n, 2), dtype=torch.float32, device=dev)
    scratch = torch.empty(img.numel(), dtype=torch.float32, device=dev)
    rc = lib.pcp_bev_interpolate(_ptr(img), 0, b, c, h, w, _ptr(pts), pts.stride(0), n,
                                 C.c_float(_f32(point_cloud_range[0])), C.c_float(_f32(point_cloud_range[1])),
                                 C.c_float(_f32(bev_pixel_size[0])), C.c_float(_f32(bev_pixel_size[1])),
                                 _ptr(scratch), _ptr(feat), _ptr(coord), _stream())
    _lib.check(rc, "pcp_bev_interpolate")
    if return_bev_coord:
        return feat, coord
    return feat
