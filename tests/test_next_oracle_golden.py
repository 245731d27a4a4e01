"""CPU: the SURVEY 8(f) oracle (oracle/next_oracle.py) reproduces the committed outputs of the REFERENCE's own
hunter_toolbox functions, DynamicMeanVFE, DynamicPillarVFESimple2D and apply_se3_ (tests/golden/next_*.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import next_oracle as no
from oracle import pillar_oracle as po
from tests.helpers import GOLDEN_DIR, layers_from_state_dict


def _load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: z[k] for k in z.files}


def test_interpolate_and_bev_scatter_bit_exact():
    g = _load("next_hunter_small")
    feat, coord = no.interpolate_points_feat_from_bev_img(torch.from_numpy(g["bev_img"]), torch.from_numpy(g["points"]),
                                                          g["range6"], g["pixel"])
    assert np.array_equal(coord.numpy(), g["ref_bev_coord"])
    assert np.array_equal(feat.numpy(), g["ref_points_feat"])
    pts = torch.from_numpy(g["points"])
    h, w = g["bev_img"].shape[2:]
    bev = no.bev_scatter(coord, pts[:, 0].long(), torch.from_numpy(g["points_feat"]), (h, w))
    assert np.array_equal(bev.numpy(), g["ref_bev_scatter"])


def test_dynamic_mean_vfe_bit_exact():
    g = _load("next_meanvfe_small")
    out = no.dynamic_mean_vfe(torch.from_numpy(g["points"]), int(g["num_point_features"]), [float(v) for v in g["voxel_size"]],
                              g["point_cloud_range"], g["grid_size"])
    assert np.array_equal(out["voxel_coords"].numpy(), g["ref_voxel_coords"])
    assert np.array_equal(out["voxel_features"].numpy(), g["ref_voxel_features"])


@pytest.mark.parametrize("tag", ["abs", "rel_dist"])
def test_simple2d_bit_exact(tag):
    g = _load(f"next_simple2d_{tag}")
    sd = {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd/")}
    cfg = po.VFEConfig(7, [float(v) for v in g["voxel_size"]], g["point_cloud_range"], g["grid_size"])
    out = no.simple2d_vfe(torch.from_numpy(g["points"]), cfg, layers_from_state_dict(sd), bool(g["use_abs"]), bool(g["with_distance"]))
    assert np.array_equal(out["pillar_coords"].numpy(), g["ref_pillar_coords"])
    assert np.array_equal(out["pillar_features"].numpy(), g["ref_pillar_features"])


def test_early_fusion_bit_exact():
    g = _load("next_early_fusion_small")
    clouds = [g[f"cloud{a}"] for a in range(4)]
    tfs = [g[f"se3_{a}"] for a in range(1, 4)]
    assert np.array_equal(no.fuse_agent_points(clouds[0], clouds[1:], tfs, g["range6"]), g["ref_fused"])
    assert np.array_equal(no.fuse_agent_points(clouds[0], clouds[1:], tfs, None), g["ref_fused_nomask"])
