"""CPU restatement of the reference's DynamicPillarVFE + PointPillarScatter.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  torch-on-CPU / numpy; every
function cites the reference lines it follows (paths relative to the reference
repository root).  It deliberately calls the SAME ATen operators the reference
calls on CPU (``torch.unique(dim=0)``, ``nn.functional.linear``,
``batch_norm`` in eval mode), so that timing it is a faithful stand-in for
the reference's CPU path, and restates torch_scatter (third party, unpinned by
the reference: README.md:68-72 only hints ``pip install torch-scatter``) from its
published semantics:

* ``scatter_mean``  = ``index_add`` sum / count, count clamped to >= 1, true division
* ``scatter_max``   = segment max (rows nobody points at stay 0)
* ``dim_size``      = ``index.max() + 1``

Pinned against outputs of the reference itself: ``tests/golden/*.npz`` are produced
by ``oracle/gen_golden.py`` from the reference modules loaded by
``oracle/ref_loader.py``; ``tests/test_oracle_golden.py`` checks this file against
them, ``tests/test_oracle_vs_reference.py`` re-checks live when ``/root/reference``
exists.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------
# torch_scatter restatement (third-party dependency of the reference, absent here)
# ----------------------------------------------------------------------------------------
def scatter_mean(src: torch.Tensor, index: torch.Tensor, dim_size: Optional[int] = None) -> torch.Tensor:
    """torch_scatter.scatter_mean(src, index, dim=0); call site dynamic_pillar_vfe.py:110."""
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    out = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    out.index_add_(0, index, src)                       # sequential, in row order, on CPU (atomics on a CUDA tensor)
    cnt = torch.zeros(dim_size, dtype=src.dtype, device=src.device)
    cnt.index_add_(0, index, torch.ones(index.shape[0], dtype=src.dtype, device=src.device))
    cnt.clamp_(min=1)
    return out / cnt.view(-1, *([1] * (src.dim() - 1)))


def scatter_max(src: torch.Tensor, index: torch.Tensor, dim_size: Optional[int] = None) -> torch.Tensor:
    """torch_scatter.scatter_max(src, index, dim=0)[0]; call site dynamic_pillar_vfe.py:40."""
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    out = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    if index.numel() == 0:
        return out
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    out.scatter_reduce_(0, idx, src, reduce="amax", include_self=False)
    return out


# ----------------------------------------------------------------------------------------
# parameters
# ----------------------------------------------------------------------------------------
@dataclass
class PFNLayerParams:
    """One PFNLayerV2 (dynamic_pillar_vfe.py:14-33): Linear(+BN eval) + ReLU."""
    weight: torch.Tensor                       # (out, in)
    bias: Optional[torch.Tensor] = None        # only when use_norm is False
    bn_weight: Optional[torch.Tensor] = None
    bn_bias: Optional[torch.Tensor] = None
    bn_mean: Optional[torch.Tensor] = None
    bn_var: Optional[torch.Tensor] = None
    eps: float = 1e-3                          # dynamic_pillar_vfe.py:29


@dataclass
class VFEConfig:
    """Constants DynamicPillarVFE.__init__ derives (dynamic_pillar_vfe.py:50-89)."""
    num_raw_point_features: int
    voxel_size: Sequence[float]
    point_cloud_range: Sequence[float]
    grid_size: Sequence[int]
    use_absolute_xyz: bool = True
    with_distance: bool = False
    use_norm: bool = True
    x_offset: float = field(init=False)
    y_offset: float = field(init=False)
    z_offset: float = field(init=False)

    def __post_init__(self):
        vs, rng = self.voxel_size, self.point_cloud_range
        # dynamic_pillar_vfe.py:80-82, evaluated with the caller's own scalar types
        # (the reference dataset passes python floats for voxel_size and np.float32 for the range).
        self.x_offset = vs[0] / 2 + rng[0]
        self.y_offset = vs[1] / 2 + rng[1]
        self.z_offset = vs[2] / 2 + rng[2]

    @property
    def c_in(self) -> int:
        c = self.num_raw_point_features + (6 if self.use_absolute_xyz else 3)
        return c + (1 if self.with_distance else 0)


def pfn_layer_forward(p: PFNLayerParams, inputs: torch.Tensor, unq_inv: torch.Tensor, last: bool,
                      dim_size: Optional[int] = None) -> torch.Tensor:
    """PFNLayerV2.forward, dynamic_pillar_vfe.py:35-46."""
    x = F.linear(inputs, p.weight, p.bias)
    if p.bn_weight is not None:
        x = F.batch_norm(x, p.bn_mean, p.bn_var, p.bn_weight, p.bn_bias, training=False, eps=p.eps)
    x = torch.relu(x)
    x_max = scatter_max(x, unq_inv, dim_size)
    if last:
        return x_max
    return torch.cat([x, x_max[unq_inv, :]], dim=1)


def dynamic_pillar_vfe(points_in: torch.Tensor, cfg: VFEConfig, layers: List[PFNLayerParams],
                       unique_dim0: bool = True) -> Dict[str, torch.Tensor]:
    """DynamicPillarVFE.forward, dynamic_pillar_vfe.py:94-147.

    ``unique_dim0=True`` issues ``torch.unique(..., dim=0)`` exactly as line 108 does (on CPU this
    selects ATen's slow per-row path and is what the CPU baseline times); ``False`` issues the flat
    ``torch.unique`` - identical results on a 1-D tensor, used by tests for speed.
    Returns the module outputs plus the intermediates the parity tests check.
    """
    points_in = points_in.float()
    scale_xy = int(cfg.grid_size[0]) * int(cfg.grid_size[1])                       # :84
    scale_y = int(cfg.grid_size[1])                                                # :85
    grid = torch.tensor([int(g) for g in cfg.grid_size])                           # :87
    voxel = torch.tensor([float(v) for v in cfg.voxel_size], dtype=torch.float32)  # :88
    rng = torch.tensor([float(v) for v in cfg.point_cloud_range], dtype=torch.float32)  # :89

    points = points_in[:, :1 + cfg.num_raw_point_features]                         # :96
    points_coords = torch.floor((points[:, [1, 2]] - rng[[0, 1]]) / voxel[[0, 1]]).int()      # :98
    mask = ((points_coords >= 0) & (points_coords < grid[[0, 1]])).all(dim=1)      # :99
    points = points[mask]                                                          # :100
    points_coords = points_coords[mask]                                            # :101
    points_xyz = points[:, [1, 2, 3]].contiguous()                                 # :102

    merge_coords = points[:, 0].int() * scale_xy + points_coords[:, 0] * scale_y + points_coords[:, 1]  # :104-106
    if unique_dim0:
        unq_coords, unq_inv, unq_cnt = torch.unique(merge_coords, return_inverse=True,
                                                    return_counts=True, dim=0)     # :108
    else:
        unq_coords, unq_inv, unq_cnt = torch.unique(merge_coords, return_inverse=True, return_counts=True)
    num_pillars = unq_coords.shape[0]

    points_mean = scatter_mean(points_xyz, unq_inv, num_pillars)                   # :110
    f_cluster = points_xyz - points_mean[unq_inv, :]                               # :111

    f_center = torch.zeros_like(points_xyz)                                        # :113
    f_center[:, 0] = points_xyz[:, 0] - (points_coords[:, 0].to(points_xyz.dtype) * cfg.voxel_size[0] + cfg.x_offset)
    f_center[:, 1] = points_xyz[:, 1] - (points_coords[:, 1].to(points_xyz.dtype) * cfg.voxel_size[1] + cfg.y_offset)
    f_center[:, 2] = points_xyz[:, 2] - cfg.z_offset                               # :114-116

    if cfg.use_absolute_xyz:                                                       # :118-121
        features = [points[:, 1:], f_cluster, f_center]
    else:
        features = [points[:, 4:], f_cluster, f_center]
    if cfg.with_distance:                                                          # :123-125
        features.append(torch.norm(points[:, 1:4], 2, dim=1, keepdim=True))
    features = torch.cat(features, dim=-1)                                         # :126
    point_features = features

    for i, lp in enumerate(layers):                                                # :128-129
        features = pfn_layer_forward(lp, features, unq_inv, last=(i == len(layers) - 1), dim_size=num_pillars)

    unq_coords = unq_coords.int()                                                  # :137
    voxel_coords = torch.stack((unq_coords // scale_xy,
                                (unq_coords % scale_xy) // scale_y,
                                unq_coords % scale_y,
                                torch.zeros(unq_coords.shape[0]).int()), dim=1)    # :138-142
    voxel_coords = voxel_coords[:, [0, 3, 2, 1]]                                   # :143
    return {
        "pillar_features": features,          # (P, C_out) fp32          :145
        "voxel_coords": voxel_coords,         # (P, 4) int32 (b, z, y, x) :146
        "unq_coords": unq_coords,             # (P,) int32 ascending
        "unq_inv": unq_inv,                   # (N',) int64 point -> pillar
        "unq_cnt": unq_cnt,                   # (P,) int64
        "keep_mask": mask,                    # (N,) bool
        "points_mean": points_mean,           # (P, 3)
        "point_features": point_features,     # (N', C_in)
    }


def pointpillar_scatter(pillar_features: torch.Tensor, coords: torch.Tensor, grid_size: Sequence[int],
                        num_bev_features: int) -> torch.Tensor:
    """PointPillarScatter.forward, pointpillar_scatter.py:14-37."""
    nx, ny, nz = (int(g) for g in grid_size)
    assert nz == 1                                                                 # :12
    batch_size = int(coords[:, 0].max().int().item()) + 1                          # :17
    out = []
    for b in range(batch_size):                                                    # :18
        spatial_feature = torch.zeros(num_bev_features, nz * nx * ny, dtype=pillar_features.dtype)   # :19-23
        batch_mask = coords[:, 0] == b                                             # :25
        this_coords = coords[batch_mask, :]                                        # :26
        indices = (this_coords[:, 1] + this_coords[:, 2] * nx + this_coords[:, 3]).long()   # :27-28
        pillars = pillar_features[batch_mask, :].t()                               # :29-30
        spatial_feature[:, indices] = pillars                                      # :31
        out.append(spatial_feature)
    out = torch.stack(out, 0)                                                      # :34
    return out.view(batch_size, num_bev_features * nz, ny, nx)                     # :35


def front_end(points: torch.Tensor, cfg: VFEConfig, layers: List[PFNLayerParams],
              unique_dim0: bool = True) -> Dict[str, torch.Tensor]:
    """VFE followed by the BEV scatter (the chain BASELINE.json's metric is quoted on)."""
    out = dynamic_pillar_vfe(points, cfg, layers, unique_dim0=unique_dim0)
    out["spatial_features"] = pointpillar_scatter(out["pillar_features"], out["voxel_coords"],
                                                  cfg.grid_size, layers[-1].weight.shape[0])
    return out


# ----------------------------------------------------------------------------------------
# numpy cross-check of the integer part (independent of torch): used by tests at full sizes
# ----------------------------------------------------------------------------------------
def quantise_keys_numpy(points: np.ndarray, cfg: VFEConfig):
    """Integer half of the path (dynamic_pillar_vfe.py:98-108) in numpy fp32 arithmetic.

    Returns (keep_mask, keys_of_kept, unq_keys, unq_inv, unq_cnt).
    """
    p = np.ascontiguousarray(points, dtype=np.float32)
    rng = np.asarray(cfg.point_cloud_range, dtype=np.float32)
    vs = np.asarray(cfg.voxel_size, dtype=np.float32)
    with np.errstate(invalid="ignore"):     # NaN / inf rows: the int cast is undefined, the rows are culled below
        qf = np.floor((p[:, 1:3] - rng[0:2]) / vs[0:2])
        q = np.where(np.isfinite(qf), qf, -1).astype(np.int64).clip(-2 ** 31, 2 ** 31 - 1).astype(np.int32)
    g = np.asarray(cfg.grid_size[:2], dtype=np.int32)
    keep = ((q >= 0) & (q < g)).all(axis=1)
    qk = q[keep]
    b = p[keep, 0].astype(np.int32)
    keys = b * (int(g[0]) * int(g[1])) + qk[:, 0] * int(g[1]) + qk[:, 1]
    unq, inv, cnt = np.unique(keys, return_inverse=True, return_counts=True)
    return keep, keys, unq.astype(np.int32), inv.astype(np.int64), cnt.astype(np.int64)
