"""CPU restatement of the reference's MoDAR exchange arithmetic.  TEST INFRASTRUCTURE ONLY.

Follows (paths relative to the reference root):

* box membership      pcdet/ops/roiaware_pool3d/src/roiaware_pool3d_kernel.cu:16-36 (test), 313-336 (first
                      box wins, default -1) - the GPU kernel (margin 1e-5), which is what the exchange path
                      calls through roiaware_pool3d_utils.py:28-41; NOT the CPU twin (margin 1e-2).
* propagation         pcdet/datasets/v2x_sim/v2x_sim_dataset_ego.py:203-215 (== workspace/visualize_collab.py:118-142)
* SE(3) + yaw wrap    pcdet/datasets/nuscenes/nuscenes_temporal_utils.py:28-29,66-70
* pack + concat       pcdet/datasets/v2x_sim/v2x_sim_dataset_ego.py:221-232 (ego rows :162-165)

Pinned against the reference's own ``apply_se3_`` (imported by oracle/ref_loader.py) through the fixtures in
tests/golden/modar_*.npz, and against the reference's own CUDA kernel compiled from its source
(oracle/Makefile -> oracle/_ref/libroiaware_ref.so) in the GPU tests.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np

MARGIN_F32 = np.float32(1e-5)          # roiaware_pool3d_kernel.cu:27 "const float MARGIN = 1e-5"


def points_in_boxes(points_xyz: np.ndarray, boxes7: np.ndarray) -> np.ndarray:
    """(F,3),(M,7) fp32 -> (F,) int32 index of the FIRST containing box, -1 if none.

    Arithmetic as the kernel does it: fp32 products/sums for the rotation, comparisons in double
    (``dz / 2.0`` and ``dx / 2.0 + MARGIN`` promote to double in the C source)."""
    pts = np.ascontiguousarray(points_xyz, dtype=np.float32)
    bxs = np.ascontiguousarray(boxes7, dtype=np.float32)
    out = np.full(pts.shape[0], -1, dtype=np.int32)
    undecided = np.ones(pts.shape[0], dtype=bool)
    x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
    for k in range(bxs.shape[0]):                                   # :329 loop, first hit breaks
        cx, cy, cz, dx, dy, dz, rz = bxs[k]
        in_z = ~(np.abs(z - cz).astype(np.float64) > np.float64(dz) / 2.0)          # :33
        cosa, sina = np.cos(np.float32(-rz)), np.sin(np.float32(-rz))              # :17
        sx, sy = (x - cx).astype(np.float32), (y - cy).astype(np.float32)
        lx = (sx * cosa + sy * (-sina)).astype(np.float32)                         # :18
        ly = (sx * sina + sy * cosa).astype(np.float32)                            # :19
        in_xy = (np.abs(lx).astype(np.float64) < np.float64(dx) / 2.0 + np.float64(MARGIN_F32)) & \
                (np.abs(ly).astype(np.float64) < np.float64(dy) / 2.0 + np.float64(MARGIN_F32))   # :35
        hit = undecided & in_z & in_xy
        out[hit] = k
        undecided &= ~hit
    return out


def propagate_modar(modar: np.ndarray, foreground: Optional[np.ndarray], scale: float = 2.0) -> np.ndarray:
    """v2x_sim_dataset_ego.py:203-215.  modar (M,9) fp32, foreground (F,13) fp32 -> modar with xyz shifted by
    ``scale * mean(flow of the foreground points inside the box)``; boxes with no point are untouched.
    ``scale`` is 2.0 in the reference (0.2 s latency, :213)."""
    modar = np.array(modar, dtype=np.float32, copy=True)
    if foreground is None or foreground.shape[0] == 0 or modar.shape[0] == 0:
        return modar
    fg = np.ascontiguousarray(foreground, dtype=np.float32)
    box_idx = points_in_boxes(fg[:, :3], modar[:, :7]).astype(np.int64)            # :203-205
    valid = box_idx > -1                                                           # :206
    fg, box_idx = fg[valid], box_idx[valid]                                        # :207-208
    unq, inv = np.unique(box_idx, return_inverse=True)                             # :210
    sums = np.zeros((unq.shape[0], 3), dtype=np.float32)
    cnt = np.zeros(unq.shape[0], dtype=np.float32)
    flow = fg[:, -3:]
    for i in range(fg.shape[0]):                    # scatter(reduce='mean') on CPU: sequential fp32 adds
        sums[inv[i]] += flow[i]
        cnt[inv[i]] += np.float32(1)
    mean = sums / np.maximum(cnt, np.float32(1))[:, None]
    offset = (mean * np.float32(scale)).astype(np.float32)                         # :213
    modar[unq, :3] += offset                                                       # :215
    return modar


def apply_se3_boxes(se3: np.ndarray, boxes7: np.ndarray) -> np.ndarray:
    """nuscenes_temporal_utils.apply_se3_ boxes branch (:66-70) on an fp32 (M,7) array and an fp64 (4,4)
    matrix: xyz goes through fp64 and is rounded to fp32 on store; yaw += atan2(R10, R00) (fp64 scalar
    added to an fp32 array - numpy >= 2 computes that in fp64 and rounds on store), then wrapped with
    fp32 atan2(sin, cos)."""
    se3 = np.asarray(se3, dtype=np.float64)
    b = np.array(boxes7, dtype=np.float32, copy=True)
    b[:, :3] = b[:, :3] @ se3[:3, :3].T + se3[:3, -1]                               # :67
    b[:, 6] += np.arctan2(se3[1, 0], se3[0, 0])                                     # :68 (+ :28-29)
    b[:, 6] = np.arctan2(np.sin(b[:, 6]), np.cos(b[:, 6]))                          # :69
    return b


def pack_modar_rows(modar: np.ndarray, n_cols: int, max_sweep_idx: float) -> np.ndarray:
    """v2x_sim_dataset_ego.py:221-226: rows [x,y,z, 0, 0, dx,dy,dz,heading,score,label, max_sweep_idx, -1]."""
    rows = np.zeros((modar.shape[0], n_cols))
    rows[:, :3] = modar[:, :3]
    rows[:, 4] = 0.0
    rows[:, 5:11] = modar[:, 3:]
    rows[:, -2] = max_sweep_idx
    rows[:, -1] = -1
    return rows


def modar_exchange(ego_points13: np.ndarray, agents: Sequence[Dict[str, np.ndarray]], max_sweep_idx: float,
                   scale: float = 2.0) -> np.ndarray:
    """The whole exchange block v2x_sim_dataset_ego.py:181-232 for one frame: for every agent
    propagate -> SE(3) -> pack -> concatenate after the ego rows.  ``scale == 0`` is EXCHANGE_NOW (:234-267).
    Returns fp32 (N + sum M, 13) (the dataset keeps fp64 until load_data_to_gpu's ``.float()``,
    pcdet/models/__init__.py:33; every value is already an fp32 value)."""
    pts = np.asarray(ego_points13, dtype=np.float64)
    for ag in agents:
        modar = np.asarray(ag["modar"], dtype=np.float32)
        if scale != 0.0:
            modar = propagate_modar(modar, ag.get("foreground"), scale)
        modar = modar.copy()
        modar[:, :7] = apply_se3_boxes(ag["target_se3_agent"], modar[:, :7])         # :219
        pts = np.concatenate((pts, pack_modar_rows(modar, pts.shape[1], max_sweep_idx)))   # :232
    return pts.astype(np.float32)
