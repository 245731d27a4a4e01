"""Early-fusion input assembly on the GPU: every agent's sweep stack mapped into the ego frame, concatenated after the
ego points, range-masked and given its frame-index column - the numpy steps directly upstream of the pillar path
(pcdet/datasets/v2x_sim/v2x_sim_dataset_ego_early.py:85-92, pcdet/datasets/nuscenes/nuscenes_temporal_utils.py:62-63,
pcdet/datasets/processor/data_processor.py:78-84, pcdet/utils/common_utils.py:64-68, pcdet/datasets/dataset.py:224-229).
The reference then shuffles the rows (data_processor.py:95-104); the pillar path is order-independent, so the rows stay
in input order (ego first, then the agents as given).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from .frontend import _ptr, _require_cuda, _stream, device_guard


@device_guard
def fuse_agent_points(ego_points: torch.Tensor, agent_points: Sequence[torch.Tensor],
                      target_se3_agents: Sequence[np.ndarray], point_cloud_range: Optional[Sequence[float]] = None,
                      batch_idx: Optional[int] = 0) -> torch.Tensor:
    """ego_points (N0, C) and agent_points[i] (Ni, C) fp32 CUDA rows [x, y, z, ...]; target_se3_agents[i] (4, 4) float64.
    Returns (N', 1 + C) rows [batch_idx, x', y', z', ...] (or (N', C) when ``batch_idx`` is None) of the points inside
    ``point_cloud_range`` (all points when it is None).  One 4-byte read-back sizes the result."""
    lib = _lib.load()
    _require_cuda(ego_points, "ego_points")
    clouds = [ego_points] + list(agent_points)
    if len(target_se3_agents) != len(agent_points):
        raise ValueError("one target_se3_agent per agent cloud")
    dev = ego_points.device
    ncols = ego_points.shape[1]
    for t in clouds:
        _require_cuda(t, "agent_points")
        if t.shape[1] != ncols:
            raise ValueError("all clouds must have the same number of columns")
    # the clouds stay where they are: the kernels take one device pointer per agent (no concatenated copy of the inputs)
    clouds = [t.detach() if (t.dtype == torch.float32 and t.is_contiguous()) else t.detach().float().contiguous() for t in clouds]
    offs = np.zeros(len(clouds) + 1, dtype=np.int32)
    offs[1:] = np.cumsum([t.shape[0] for t in clouds])
    n = int(offs[-1])
    se3 = np.zeros((len(clouds), 12), dtype=np.float64)
    se3[0] = np.eye(4)[:3].reshape(-1)
    for i, tf in enumerate(target_se3_agents):
        tf = np.asarray(tf, dtype=np.float64)
        assert tf.shape == (4, 4)
        se3[i + 1] = tf[:3].reshape(-1)
    ptrs = np.asarray([t.data_ptr() if t.shape[0] else 0 for t in clouds], dtype=np.uint64)
    from .modar import _upload
    d_se3, d_offs, d_ptrs = _upload(dev, [se3, offs, ptrs])       # one small asynchronous H2D from pinned staging
    with_b = batch_idx is not None
    out = torch.empty((max(n, 1), ncols + (1 if with_b else 0)), dtype=torch.float32, device=dev)
    scratch = torch.empty(int(lib.pcp_fuse_scratch_bytes(n)) // 4 + 1, dtype=torch.int32, device=dev)
    count = torch.zeros(1, dtype=torch.int32, device=dev)
    rng = None
    if point_cloud_range is not None:
        rng = (C.c_float * 6)(*[float(np.float32(v)) for v in point_cloud_range])   # np.float32 range: dataset.py:25
    rc = lib.pcp_fuse_agent_clouds(_ptr(d_ptrs), ncols, ncols, n, _ptr(d_offs), _ptr(d_se3), len(clouds),
                                   rng, int(with_b), C.c_float(float(batch_idx or 0)), _ptr(scratch), _ptr(out),
                                   out.stride(0), _ptr(count), _stream())
    _lib.check(rc, "pcp_fuse_agent_clouds")
    return out[:int(count.item())]
