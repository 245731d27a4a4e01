"""GPU property tests (hypothesis): random small clouds - ragged frames, duplicates, points on cell edges and outside the
range, empty frames - through the sm_100a path vs the CPU oracle.  SURVEY.md section 8c: bit-exact indices / occupancy,
means bit-exact, PFN outputs within 1e-5; plus the size-independent properties (row permutation changes nothing, occupancy
is idempotent under re-voxelising the pillar centres)."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import next_oracle as no
from oracle import pillar_oracle as po
from tests.helpers import assert_features_close, layers_from_state_dict, model_cfgs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SETTINGS = dict(max_examples=25, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)

RNG = np.asarray([-3.2, -1.6, -4.0, 3.2, 1.6, 0.0], dtype=np.float32)      # 32 x 16 pillars of 0.2 m
VOX = [0.2, 0.2, 4.0]


def cloud(seed, n, frames, snap):
    g = torch.Generator().manual_seed(seed)
    pts = torch.zeros(n, 8)
    pts[:, 0] = torch.randint(0, frames, (n,), generator=g).float()
    pts[:, 1] = (torch.rand(n, generator=g) * 2 - 1) * 3.5
    pts[:, 2] = (torch.rand(n, generator=g) * 2 - 1) * 1.8
    pts[:, 3] = torch.rand(n, generator=g) * -4.4 + 0.2
    pts[:, 4:6] = torch.rand(n, 2, generator=g)
    if snap:                                                      # many points exactly on cell edges / duplicates
        k = n // 2
        pts[:k, 1] = torch.round(pts[:k, 1] / 0.2) * 0.2
        pts[:k, 2] = torch.round(pts[:k, 2] / 0.2) * 0.2
    return pts


_modules = {}


def modules():
    if not _modules:
        import pcp_b200
        from pcp_b200 import synthetic as syn
        grid = syn.grid_size_of(RNG, VOX)
        vfe_cfg, scat_cfg = model_cfgs(5)
        sd = syn.pfn_state_dict(11)
        vfe = pcp_b200.DynamicPillarVFE(model_cfg=vfe_cfg, num_point_features=5, voxel_size=VOX, grid_size=grid, point_cloud_range=RNG)
        vfe.load_state_dict(sd)
        _modules.update(vfe=vfe.to(DEV).eval(), scat=pcp_b200.PointPillarScatter(model_cfg=scat_cfg, grid_size=grid).to(DEV).eval(),
                        sd=sd, grid=grid)
    return _modules


@settings(**SETTINGS)
@given(seed=st.integers(0, 10 ** 6), n=st.integers(1, 3000), frames=st.integers(1, 5), snap=st.booleans())
def test_pillar_chain_matches_oracle(seed, n, frames, snap):
    from pcp_b200.modules import CTX_KEY
    m = modules()
    pts = cloud(seed, n, frames, snap)
    cfg = po.VFEConfig(5, VOX, RNG, m["grid"])
    if po.dynamic_pillar_vfe(pts, cfg, layers_from_state_dict(m["sd"]), unique_dim0=False)["voxel_coords"].shape[0] == 0:
        return                                                    # the reference's scatter fails on an empty batch (:17)
    want = po.front_end(pts, cfg, layers_from_state_dict(m["sd"]), unique_dim0=False)
    with torch.no_grad():
        bd = m["scat"](m["vfe"]({"points": pts.to(DEV), "batch_size": frames}))
    assert torch.equal(bd["voxel_coords"].cpu(), want["voxel_coords"])
    pp = bd[CTX_KEY]["point_pillar"].cpu().long()
    assert torch.equal(pp[pp >= 0], want["unq_inv"].long())
    assert_features_close(bd["pillar_features"].cpu().numpy(), want["pillar_features"].numpy(), "pillar_features")
    sf = bd["spatial_features"].cpu()
    assert tuple(sf.shape) == tuple(want["spatial_features"].shape)
    assert torch.equal((sf != 0).any(1), (want["spatial_features"] != 0).any(1))
    # permutation of the rows: identical outputs (means are summed in ascending row order of the PERMUTED rows, so
    # the features are compared with the tolerance, the indices exactly)
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(seed + 1))
    with torch.no_grad():
        bd2 = m["scat"](m["vfe"]({"points": pts[perm].to(DEV), "batch_size": frames}))
    assert torch.equal(bd2["voxel_coords"], bd["voxel_coords"])
    assert_features_close(bd2["pillar_features"].cpu().numpy(), bd["pillar_features"].cpu().numpy(), "permuted rows")
    assert torch.equal((bd2["spatial_features"] != 0).any(1), (bd["spatial_features"] != 0).any(1))


@settings(**SETTINGS)
@given(seed=st.integers(0, 10 ** 6), n=st.integers(1, 3000), frames=st.integers(1, 4), c=st.integers(1, 40), snap=st.booleans())
def test_bev_scatter_and_interpolate_match_oracle(seed, n, frames, c, snap):
    import pcp_b200
    g = torch.Generator().manual_seed(seed)
    h, w = 16, 32
    pts = cloud(seed, n, frames, snap)
    pts[0, 0] = frames - 1
    img = torch.randn(frames, c, h, w, generator=g)
    feat, coord = pcp_b200.interpolate_points_feat_from_bev_img(img.to(DEV), pts.to(DEV), RNG, np.asarray([0.2, 0.2], dtype=np.float32), True)
    wf, wc = no.interpolate_points_feat_from_bev_img(img, pts, RNG[:2], np.asarray([0.2, 0.2], dtype=np.float32))
    assert torch.equal(coord.cpu(), wc) and torch.equal(feat.cpu(), wf)
    bev = pcp_b200.bev_scatter(coord, pts[:, 0].long().to(DEV), feat, (h, w))
    assert torch.equal(bev.cpu(), no.bev_scatter(wc, pts[:, 0].long(), wf, (h, w)))


@settings(**SETTINGS)
@given(seed=st.integers(0, 10 ** 6), n=st.integers(1, 3000), frames=st.integers(1, 4), c=st.integers(3, 7), snap=st.booleans())
def test_dynamic_mean_vfe_matches_oracle(seed, n, frames, c, snap):
    import pcp_b200
    vox = [0.2, 0.2, 0.25]
    grid = np.asarray([32, 16, 16])
    pts = cloud(seed, n, frames, snap)
    m = pcp_b200.DynamicMeanVFE(model_cfg=pcp_b200.CfgDict(), num_point_features=c, voxel_size=vox, grid_size=grid, point_cloud_range=RNG)
    bd = m({"points": pts.to(DEV), "batch_size": frames})
    want = no.dynamic_mean_vfe(pts, c, vox, RNG, grid)
    assert torch.equal(bd["voxel_coords"].cpu(), want["voxel_coords"])
    assert torch.equal(bd["voxel_features"].cpu(), want["voxel_features"])
