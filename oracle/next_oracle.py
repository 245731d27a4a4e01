"""CPU restatement of the SURVEY.md section 8(f) rows.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Every function cites the reference lines it follows (paths relative to the reference root).  Pinned against outputs of
the reference itself: ``oracle/gen_golden.py`` runs the reference's own ``bev_scatter`` /
``interpolate_points_feat_from_bev_img`` (pcdet/models/bev_layers/hunter_toolbox.py), ``DynamicMeanVFE``
(dynamic_mean_vfe.py), ``DynamicPillarVFESimple2D`` (dynamic_pillar_vfe.py:150-245), ``apply_se3_`` and
``mask_points_by_range`` on seeded inputs and commits the results as ``tests/golden/next_*.npz``;
``tests/test_oracle_golden.py`` checks this file against them.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from . import pillar_oracle as po


# ----------------------------------------------------------------------------------------
# hunter_toolbox.py
# ----------------------------------------------------------------------------------------
def bilinear_interpolate(im: torch.Tensor, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """hunter_toolbox.py:8-41; im (H, W, C)."""
    x0 = torch.floor(x).long()
    x1 = x0 + 1
    y0 = torch.floor(y).long()
    y1 = y0 + 1
    x0 = torch.clamp(x0, 0, im.shape[1] - 1)
    x1 = torch.clamp(x1, 0, im.shape[1] - 1)
    y0 = torch.clamp(y0, 0, im.shape[0] - 1)
    y1 = torch.clamp(y1, 0, im.shape[0] - 1)
    ia, ib, ic, id_ = im[y0, x0], im[y1, x0], im[y0, x1], im[y1, x1]
    wa = (x1.type_as(x) - x) * (y1.type_as(y) - y)
    wb = (x1.type_as(x) - x) * (y - y0.type_as(y))
    wc = (x - x0.type_as(x)) * (y1.type_as(y) - y)
    wd = (x - x0.type_as(x)) * (y - y0.type_as(y))
    return ia * wa[:, None] + ib * wb[:, None] + ic * wc[:, None] + id_ * wd[:, None]      # :40, left to right


def interpolate_points_feat_from_bev_img(bev_img, points, point_cloud_range, bev_pixel_size):
    """hunter_toolbox.py:99-131 -> (points_feat (N, C), points_bev_coord (N, 2))."""
    rng = torch.as_tensor(point_cloud_range, dtype=torch.float32)
    pix = torch.as_tensor(bev_pixel_size, dtype=torch.float32)
    feat = bev_img.new_zeros(points.shape[0], bev_img.shape[1])
    coord = (points[:, 1:3] - rng[:2]) / pix                                             # :114
    bidx = points[:, 0].long()
    for b in range(bev_img.shape[0]):
        img = bev_img[b].permute(1, 2, 0)                                                # 'C H W -> H W C'
        m = bidx == b
        feat[m] = bilinear_interpolate(img, coord[m, 0], coord[m, 1])
    return feat, coord


def bev_scatter(points_bev_coord, points_batch_idx, points_feat, bev_img_size):
    """hunter_toolbox.py:65-96."""
    batch_size = int(points_batch_idx.max()) + 1
    c = points_feat.shape[1]
    height, width = bev_img_size
    area = height * width
    x, y = points_bev_coord[:, 0], points_bev_coord[:, 1]
    mask = (x > 0) & (x < width) & (y > 0) & (y < height)                                 # :78-79
    coord = points_bev_coord[mask].long()
    merge = points_batch_idx[mask] * area + coord[:, 1] * width + coord[:, 0]            # :82
    unq, inv = torch.unique(merge, return_inverse=True)
    img = points_feat.new_zeros(batch_size * area, c)
    if unq.numel():
        img[unq] = po.scatter_mean(points_feat[mask], inv)
    return img.view(batch_size, height, width, c).permute(0, 3, 1, 2).contiguous()      # '(B H W) C -> B C H W'


# ----------------------------------------------------------------------------------------
# dynamic_mean_vfe.py
# ----------------------------------------------------------------------------------------
def dynamic_mean_vfe(points: torch.Tensor, num_point_features: int, voxel_size, point_cloud_range, grid_size):
    """dynamic_mean_vfe.py:53-79 -> dict(voxel_features (V, C), voxel_coords (V, 4) int32, unq_inv)."""
    vs = torch.as_tensor(np.asarray(voxel_size), dtype=torch.float32)
    rng = torch.as_tensor(np.asarray(point_cloud_range), dtype=torch.float32)
    gs = torch.as_tensor(np.asarray(grid_size))
    scale_xyz = int(grid_size[0]) * int(grid_size[1]) * int(grid_size[2])
    scale_yz = int(grid_size[1]) * int(grid_size[2])
    scale_z = int(grid_size[2])
    pts = points[:, :1 + num_point_features]
    pc = torch.floor((pts[:, 1:4] - rng[0:3]) / vs).int()                                # :56
    mask = ((pc >= 0) & (pc < gs)).all(dim=1)
    pts, pc = pts[mask], pc[mask]
    merge = pts[:, 0].int() * scale_xyz + pc[:, 0] * scale_yz + pc[:, 1] * scale_z + pc[:, 2]
    data = pts[:, 1:].contiguous()
    unq, inv, _ = torch.unique(merge, return_inverse=True, return_counts=True)
    mean = po.scatter_mean(data, inv)
    unq = unq.int()
    vc = torch.stack((unq // scale_xyz, (unq % scale_xyz) // scale_yz, (unq % scale_yz) // scale_z, unq % scale_z), dim=1)
    vc = vc[:, [0, 3, 2, 1]]
    return {"voxel_features": mean.contiguous(), "voxel_coords": vc.contiguous(), "unq_inv": inv}


# ----------------------------------------------------------------------------------------
# dynamic_pillar_vfe.py:150-245 (DynamicPillarVFESimple2D)
# ----------------------------------------------------------------------------------------
def simple2d_vfe(points: torch.Tensor, cfg: "po.VFEConfig", layers: List["po.PFNLayerParams"], use_absolute_xyz=True,
                 with_distance=False):
    """-> dict(pillar_features (P, C_out), pillar_coords (P, 3) int32 rows (b, y, x), unq_inv)."""
    vs = torch.as_tensor(np.asarray(cfg.voxel_size), dtype=torch.float32)
    rng = torch.as_tensor(np.asarray(cfg.point_cloud_range), dtype=torch.float32)
    gs = torch.as_tensor(np.asarray(cfg.grid_size[:2]))
    vx, vy, vz = cfg.voxel_size[0], cfg.voxel_size[1], cfg.voxel_size[2]
    x_off = vx / 2 + cfg.point_cloud_range[0]
    y_off = vy / 2 + cfg.point_cloud_range[1]
    z_off = vz / 2 + cfg.point_cloud_range[2]
    scale_xy = int(cfg.grid_size[0]) * int(cfg.grid_size[1])
    scale_y = int(cfg.grid_size[1])
    pc = torch.floor((points[:, [1, 2]] - rng[[0, 1]]) / vs[[0, 1]]).int()               # :201-202
    mask = ((pc >= 0) & (pc < gs)).all(dim=1)
    points, pc = points[mask], pc[mask]
    xyz = points[:, [1, 2, 3]].contiguous()
    merge = points[:, 0].int() * scale_xy + pc[:, 0] * scale_y + pc[:, 1]
    unq, inv, _ = torch.unique(merge, return_inverse=True, return_counts=True, dim=0)
    f_center = torch.zeros_like(xyz)
    f_center[:, 0] = xyz[:, 0] - (pc[:, 0].to(xyz.dtype) * vx + x_off)                  # :213-215
    f_center[:, 1] = xyz[:, 1] - (pc[:, 1].to(xyz.dtype) * vy + y_off)
    f_center[:, 2] = xyz[:, 2] - z_off
    feats = [f_center, points[:, 1:] if use_absolute_xyz else points[:, 4:]]
    if with_distance:
        feats.append(torch.norm(points[:, 1:4], 2, dim=1, keepdim=True))
    x = torch.cat(feats, dim=-1)
    for i, lp in enumerate(layers):
        x = po.pfn_layer_forward(lp, x, inv, last=(i == len(layers) - 1))
    unq = unq.int()
    coords = torch.stack((unq // scale_xy, (unq % scale_xy) // scale_y, unq % scale_y), dim=1)[:, [0, 2, 1]]
    return {"pillar_features": x, "pillar_coords": coords.contiguous(), "unq_inv": inv}


# ----------------------------------------------------------------------------------------
# early-fusion assembly
# ----------------------------------------------------------------------------------------
def fuse_agent_points(ego_points: np.ndarray, agent_points: Sequence[np.ndarray], target_se3_agents: Sequence[np.ndarray],
                      point_cloud_range: Optional[Sequence[float]] = None) -> np.ndarray:
    """v2x_sim_dataset_ego_early.py:85-92 (apply_se3_, nuscenes_temporal_utils.py:62-63, on fp32 arrays: the fp64 result is
    stored back into the fp32 cloud) + np.concatenate + mask_points_by_range (common_utils.py:64-68, fp32 range,
    dataset.py:25).  No shuffle.  Returns fp32 (N', C)."""
    clouds = [np.asarray(ego_points, dtype=np.float32)]
    for pts, tf in zip(agent_points, target_se3_agents):
        p = np.array(pts, dtype=np.float32, copy=True)
        tf = np.asarray(tf, dtype=np.float64)
        p[:, :3] = p[:, :3] @ tf[:3, :3].T + tf[:3, -1]
        clouds.append(p)
    pts = np.concatenate(clouds, axis=0)
    if point_cloud_range is not None:
        r = np.asarray(point_cloud_range, dtype=np.float32)
        m = (pts[:, 0] >= r[0]) & (pts[:, 0] < r[3]) & (pts[:, 1] >= r[1]) & (pts[:, 1] < r[4]) \
            & (pts[:, 2] >= r[2]) & (pts[:, 2] < r[5])
        pts = pts[m]
    return pts


def select_foreground(points, points_cls_logit, points_flow3d, batch_size, threshold=0.3):
    """pcdet/models/bev_layers/hunter_jr.py:377-397, the test-time exchange block, line by line on CPU tensors:
    sigmoid (:379), mask_send = P(background) < 0.3 (:380), cat of points[mask, 1:] | prob[mask] | flow[mask] (:382-385),
    per-sample boolean split (:387-391).  Returns one (F_b, C + 6) tensor per sample (empty when the sample sends nothing;
    the reference skips those at :392)."""
    import torch
    points_cls_prob = torch.sigmoid(points_cls_logit)
    mask_send = points_cls_prob[:, 0] < threshold
    points_to_send = torch.cat([points[mask_send, 1:], points_cls_prob[mask_send], points_flow3d[mask_send]], dim=1)
    points_to_send_batch_idx = points[mask_send, 0].long()
    return [points_to_send[points_to_send_batch_idx == b_idx] for b_idx in range(batch_size)]
