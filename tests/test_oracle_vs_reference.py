"""The oracle against the reference's OWN modules, live (no committed fixture in between): DynamicPillarVFE.forward +
PointPillarScatter.forward (dynamic_pillar_vfe.py:94-147, pointpillar_scatter.py:14-37) and apply_se3_
(nuscenes_temporal_utils.py:66-70), loaded by file path by oracle/ref_loader.py from /root/reference or from the
git-ignored copies build() leaves under oracle/_ref/py.  Skipped when neither is present."""
import numpy as np
import pytest
import torch

from oracle import modar_oracle as mo
from oracle import pillar_oracle as po
from oracle import ref_loader as rl
from tests.helpers import layers_from_state_dict

pytestmark = pytest.mark.skipif(not rl.reference_available(), reason="reference sources not available")


def run_reference(points, c_raw, voxel, rng, grid, sd, **kw):
    vfe, scat = rl.build_reference_front_end(c_raw, voxel, rng, grid, **kw)
    missing = vfe.load_state_dict(sd)
    assert not missing.missing_keys and not missing.unexpected_keys
    with torch.no_grad():
        return scat(vfe({"points": points.clone()}))


@pytest.mark.parametrize("n_frames,n_points,config_id,ego", [(1, 32768, 1, False), (1, 20000, 2, True), (2, 60000, 3, False)])
def test_front_end_matches_the_reference_modules(n_frames, n_points, config_id, ego):
    from pcp_b200 import synthetic as syn
    c_raw = 11 if ego else 5
    rng = np.asarray(syn.V2X_RANGE, dtype=np.float32)
    vox = syn.V2X_VOXEL
    grid = syn.grid_size_of(rng, vox)
    sd = syn.pfn_state_dict(c_raw + 6, (64, 64), True, config_id)
    pts = syn.batch_of_frames(n_frames, n_points, config_id, ego_columns=ego)
    bd = run_reference(pts, c_raw, vox, rng, grid, sd)
    cfg = po.VFEConfig(c_raw, vox, rng, grid)
    for unique_dim0 in (True, False):
        if unique_dim0 and n_points > 40000:
            continue                                   # unique(dim=0) on CPU is the slow loop the reference pays
        want = po.front_end(pts, cfg, layers_from_state_dict(sd), unique_dim0=unique_dim0)
        assert torch.equal(bd["voxel_coords"], want["voxel_coords"])
        assert torch.equal(bd["pillar_features"], want["pillar_features"]), "oracle is not bit-identical to the reference"
        assert torch.equal(bd["spatial_features"], want["spatial_features"])


@pytest.mark.parametrize("opts", [dict(num_filters=(64,)), dict(use_norm=False), dict(with_distance=True),
                                  dict(use_absolute_xyz=False)])
def test_front_end_options_match_the_reference_modules(opts):
    from pcp_b200 import synthetic as syn
    rng = np.asarray(syn.V2X_RANGE, dtype=np.float32)
    vox = syn.V2X_VOXEL
    grid = syn.grid_size_of(rng, vox)
    nf = opts.get("num_filters", (64, 64))
    use_norm, dist, use_abs = opts.get("use_norm", True), opts.get("with_distance", False), opts.get("use_absolute_xyz", True)
    c_in = 5 + (6 if use_abs else 3) + (1 if dist else 0)
    sd = syn.pfn_state_dict(c_in, nf, use_norm, 4)
    pts = syn.batch_of_frames(2, 9000, 21)
    bd = run_reference(pts, 5, vox, rng, grid, sd, **opts)
    cfg = po.VFEConfig(5, vox, rng, grid, use_absolute_xyz=use_abs, with_distance=dist, use_norm=use_norm)
    want = po.front_end(pts, cfg, layers_from_state_dict(sd, use_norm), unique_dim0=False)
    assert torch.equal(bd["voxel_coords"], want["voxel_coords"])
    assert torch.equal(bd["pillar_features"], want["pillar_features"])
    assert torch.equal(bd["spatial_features"], want["spatial_features"])


def test_apply_se3_matches_the_reference():
    ns = rl.load_reference_modules()
    g = np.random.default_rng(5)
    for _ in range(5):
        yaw = g.uniform(-np.pi, np.pi)
        se3 = np.eye(4)
        se3[:2, :2] = [[np.cos(yaw), -np.sin(yaw)], [np.sin(yaw), np.cos(yaw)]]
        se3[:3, 3] = g.uniform(-30, 30, 3)
        boxes = g.uniform(-40, 40, (50, 7)).astype(np.float32)
        boxes[:, 6] = g.uniform(-np.pi, np.pi, 50).astype(np.float32)
        ref = ns.apply_se3_(se3, boxes_=boxes.copy(), return_transformed=True)
        got = mo.apply_se3_boxes(se3, boxes)
        assert np.array_equal(np.asarray(ref, dtype=np.float32), got)
