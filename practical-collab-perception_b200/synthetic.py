"""Seeded synthetic inputs of the shapes BASELINE.json's configs name.

There is no dataset in the image, so tests, smoke() and bench.py all draw from here.  Everything is
generated on the CPU with an explicit ``torch.Generator`` so the same seed gives the same bytes on the
build container and on the GPU box.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

V2X_RANGE = [-51.2, -51.2, -8.0, 51.2, 51.2, 0.0]      # tools/cfgs/dataset_configs/v2x_sim_dataset_car.yaml:10
V2X_VOXEL = [0.2, 0.2, 8.0]                            # v2x_sim_dataset_car.yaml:53
STRESS_VOXEL = [0.1, 0.1, 8.0]                         # BASELINE.json configs[4]


def grid_size_of(point_cloud_range, voxel_size) -> np.ndarray:
    """pcdet/datasets/processor/data_processor.py:109-110."""
    rng = np.asarray(point_cloud_range, dtype=np.float32)
    g = (rng[3:6] - rng[0:3]) / np.array(voxel_size)
    return np.round(g).astype(np.int64)


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def lidar_frame(n_points: int, seed: int, batch_idx: int = 0, ego_columns: bool = False,
                sigma_r: float = 22.0, uniform_xy: bool = False) -> torch.Tensor:
    """One frame: (n, 8) rows [b,x,y,z,intensity,timelag,sweep_idx,inst_idx] (car / early-fusion layout,
    v2x_sim_dataset_car.yaml:22-26) or (n, 14) rows [b, pt5, 0 x6, sweep_idx, inst_idx] (ego layout,
    v2x_sim_dataset_ego.py:162-165).  Radial lidar-like density; ~1 % of points fall outside +-51.2 m
    and exercise the cull.
    """
    g = _gen(seed)
    if uniform_xy:
        x = (torch.rand(n_points, generator=g) * 2 - 1) * 52.0
        y = (torch.rand(n_points, generator=g) * 2 - 1) * 52.0
    else:
        r = torch.randn(n_points, generator=g).abs() * sigma_r
        th = torch.rand(n_points, generator=g) * (2 * math.pi)
        x, y = r * torch.cos(th), r * torch.sin(th)
    z = torch.rand(n_points, generator=g) * -8.0
    inten = torch.rand(n_points, generator=g)
    lag_i = torch.randint(0, 11, (n_points,), generator=g)
    lag = lag_i.float() * 0.1
    sweep = (10 - lag_i).float()
    inst = torch.full((n_points,), -1.0)
    b = torch.full((n_points,), float(batch_idx))
    if ego_columns:
        zeros = torch.zeros(n_points, 6)
        cols = [b[:, None], x[:, None], y[:, None], z[:, None], inten[:, None], lag[:, None], zeros,
                sweep[:, None], inst[:, None]]
    else:
        cols = [b[:, None], x[:, None], y[:, None], z[:, None], inten[:, None], lag[:, None],
                sweep[:, None], inst[:, None]]
    return torch.cat(cols, dim=1).float().contiguous()


def batch_of_frames(n_frames: int, n_points: int, config_id: int, ego_columns: bool = False,
                    uniform_xy: bool = False, first_frame: int = 0) -> torch.Tensor:
    """collate_batch layout (pcdet/datasets/dataset.py:224-229): frames concatenated, col 0 = frame index.
    seed = 1000*config + global frame number, local batch index restarts at 0."""
    frames = [lidar_frame(n_points, 1000 * config_id + first_frame + f, batch_idx=f, ego_columns=ego_columns,
                          uniform_xy=uniform_xy) for f in range(n_frames)]
    return torch.cat(frames, dim=0).contiguous()


def pfn_state_dict(c_in: int, num_filters=(64, 64), use_norm: bool = True, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Random-init weights with the reference's parameter names (pfn_layers.{i}.linear.weight,
    pfn_layers.{i}.norm.*; dynamic_pillar_vfe.py:24-31,64-75).  nn.Linear default init under
    manual_seed(seed); BN buffers randomised so that BN is not the identity."""
    g = _gen(seed)
    sd: Dict[str, torch.Tensor] = {}
    widths = [c_in] + list(num_filters)
    for i in range(len(widths) - 1):
        last = i >= len(widths) - 2
        cin, cout = widths[i], widths[i + 1]
        if not last:
            cout //= 2
        bound = 1.0 / math.sqrt(cin)
        sd[f"pfn_layers.{i}.linear.weight"] = (torch.rand(cout, cin, generator=g) * 2 - 1) * bound
        if use_norm:
            sd[f"pfn_layers.{i}.norm.weight"] = torch.rand(cout, generator=g) + 0.5
            sd[f"pfn_layers.{i}.norm.bias"] = torch.randn(cout, generator=g) * 0.1
            sd[f"pfn_layers.{i}.norm.running_mean"] = torch.randn(cout, generator=g) * 0.1
            sd[f"pfn_layers.{i}.norm.running_var"] = torch.rand(cout, generator=g) + 0.5
            sd[f"pfn_layers.{i}.norm.num_batches_tracked"] = torch.tensor(1, dtype=torch.long)
        else:
            sd[f"pfn_layers.{i}.linear.bias"] = (torch.rand(cout, generator=g) * 2 - 1) * bound
        # layer i+1 consumes cat[x, x_max] = 2*cout channels
        widths[i + 1] = cout * 2 if not last else cout
    return sd


def modar_agent(seed: int, n_boxes: Optional[int] = None, fg_per_box: Tuple[int, int] = (30, 200),
                stray_fraction: float = 0.1) -> Dict[str, torch.Tensor]:
    """One agent's exchange payload:
    modar (M, 9) = box7 | score | label  (center_head.py:409-427),
    foreground (F, 13) = pt5 | sweep | inst | cls3 | flow3 (hunter_jr.py:377-397),
    target_se3_agent (4, 4) float64.  Foreground points are drawn well inside their box (|local| <=
    0.45*dim) so the box-membership test has no knife-edge cases; a fraction of stray points lies in no box."""
    g = _gen(seed)
    m = int(torch.randint(20, 84, (1,), generator=g)) if n_boxes is None else int(n_boxes)
    ctr = (torch.rand(m, 2, generator=g) * 2 - 1) * 40.0
    cz = -(torch.rand(m, 1, generator=g) * 2 + 1)
    dims = torch.tensor([4.5, 1.9, 1.6]) + torch.randn(m, 3, generator=g) * torch.tensor([0.5, 0.2, 0.2])
    dims = dims.clamp(min=0.5)
    yaw = (torch.rand(m, 1, generator=g) * 2 - 1) * math.pi
    score = torch.rand(m, 1, generator=g) * 0.9 + 0.1
    label = torch.ones(m, 1)
    modar = torch.cat([ctr, cz, dims, yaw, score, label], dim=1).float()
    vel = torch.randn(m, 3, generator=g) * torch.tensor([5.0, 5.0, 0.0])

    rows = []
    for k in range(m):
        f = int(torch.randint(fg_per_box[0], fg_per_box[1] + 1, (1,), generator=g))
        local = (torch.rand(f, 3, generator=g) * 2 - 1) * 0.45 * dims[k]
        c, s = math.cos(float(yaw[k])), math.sin(float(yaw[k]))
        wx = local[:, 0] * c - local[:, 1] * s + modar[k, 0]
        wy = local[:, 0] * s + local[:, 1] * c + modar[k, 1]
        wz = local[:, 2] + modar[k, 2]
        flow = vel[k] * 0.1 + torch.randn(f, 3, generator=g) * 0.05
        row = torch.zeros(f, 13)
        row[:, 0], row[:, 1], row[:, 2] = wx, wy, wz
        row[:, 3] = torch.rand(f, generator=g)
        row[:, 4] = torch.randint(0, 11, (f,), generator=g).float() * 0.1
        row[:, 5] = 10 - row[:, 4] * 10
        row[:, 6] = float(k)
        row[:, 7:10] = torch.tensor([0.05, 0.15, 0.8])
        row[:, 10:13] = flow
        rows.append(row)
    n_stray = int(stray_fraction * sum(r.shape[0] for r in rows))
    if n_stray:
        stray = torch.zeros(n_stray, 13)
        stray[:, 0:2] = (torch.rand(n_stray, 2, generator=g) * 2 - 1) * 50.0
        stray[:, 2] = 20.0 + torch.rand(n_stray, generator=g)       # far above every box
        stray[:, 10:13] = torch.randn(n_stray, 3, generator=g)
        rows.append(stray)
    fg = torch.cat(rows, dim=0)
    fg = fg[torch.randperm(fg.shape[0], generator=g)].contiguous().float()

    t = (torch.rand(2, generator=g, dtype=torch.float64) * 2 - 1) * 30.0
    a = float((torch.rand(1, generator=g, dtype=torch.float64) * 2 - 1) * math.pi)
    se3 = np.eye(4)
    se3[:2, :2] = [[math.cos(a), -math.sin(a)], [math.sin(a), math.cos(a)]]
    se3[:2, 3] = t.numpy()
    se3[2, 3] = float(torch.randn(1, generator=g, dtype=torch.float64) * 0.2)
    return {"modar": modar.contiguous(), "foreground": fg, "target_se3_agent": se3}


def modar_scene(config_id: int = 2, frame: int = 0, n_agents: int = 5, n_ego_points: int = 32768):
    """Config 2: ego sweep stack (14 columns, batch column included) + n_agents payloads."""
    ego = lidar_frame(n_ego_points, 1000 * config_id + frame, batch_idx=0, ego_columns=True)
    agents: List[Dict[str, torch.Tensor]] = [modar_agent(100000 * config_id + 100 * frame + a) for a in range(n_agents)]
    return ego, agents


def model_cfgs(c_raw, num_filters=(64, 64), use_norm=True, with_distance=False, use_abs=True):
    """(VFE cfg, scatter cfg) with the keys of tools/cfgs/v2x_sim_models/v2x_pointpillar_basic_*.yaml (MODEL.VFE,
    MODEL.MAP_TO_BEV)."""
    from .config import CfgDict
    vfe = CfgDict(NAME="DynPillarVFE", NUM_RAW_POINT_FEATURES=int(c_raw), USE_NORM=bool(use_norm),
                  WITH_DISTANCE=bool(with_distance), USE_ABSLOTE_XYZ=bool(use_abs), NUM_FILTERS=list(num_filters))
    scat = CfgDict(NAME="PointPillarScatter", NUM_BEV_FEATURES=int(num_filters[-1]))
    return vfe, scat
