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


class _RadixWhereItApplies(str):
    """Marker value of the fixture below: FrontEnd falls back to "auto" for shapes the radix method does not cover."""


@pytest.fixture(autouse=True, params=["auto", "radix", "binned"])
def voxelize_method(request):
    """Every test of this file runs on both compaction algorithms of pcp_voxelize_method: the dense-histogram path ("auto")
    and the stable radix sort (wherever it applies).  Results must be identical."""
    from pcp_b200 import frontend
    old = frontend.DEFAULT_VOXELIZE_METHOD
    frontend.DEFAULT_VOXELIZE_METHOD = {"radix": "radix_or_auto", "binned": "binned_or_auto"}.get(request.param, "auto")
    yield request.param
    frontend.DEFAULT_VOXELIZE_METHOD = old


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
    from pcp_b200 import frontend
    return _modules_for(frontend.DEFAULT_VOXELIZE_METHOD)


def _modules_for(method, _cache={}):
    _modules = _cache.setdefault(method, {})        # the module's FrontEnd binds the method when it is first used
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


# ---- late-fusion NMS: random clustered boxes vs the oracle ------------------------------------
def _boxes(seed, n_obj, per_obj, degenerate):
    g = torch.Generator().manual_seed(seed)
    rows = []
    for _ in range(n_obj):
        c = (torch.rand(2, generator=g) * 2 - 1) * 20
        dims = torch.tensor([4.0, 1.8, 1.5]) + torch.rand(3, generator=g)
        yaw = float((torch.rand(1, generator=g) * 2 - 1) * 3.2)
        for _ in range(per_obj):
            j = torch.randn(7, generator=g) * torch.tensor([0.4, 0.4, 0.1, 0.2, 0.1, 0.05, 0.15])
            rows.append(torch.cat([c + j[:2], torch.tensor([-1.0]) + j[2:3], dims + j[3:6], torch.tensor([yaw]) + j[6:7],
                                   torch.rand(1, generator=g)]))
    b = torch.stack(rows).float()
    if degenerate:                        # exact duplicates, axis-aligned boxes, equal scores
        b[1] = b[0]
        b[2, 6] = 0.0
        b[3, 6] = float(np.pi / 2)
        b[4, 7] = b[5, 7]
    return b[torch.randperm(b.shape[0], generator=g)].contiguous()


@settings(**SETTINGS)
@given(seed=st.integers(0, 10_000), n_obj=st.integers(1, 8), per_obj=st.integers(1, 5),
       thresh=st.sampled_from([0.01, 0.1, 0.2, 0.5, 0.7]), degenerate=st.booleans(),
       score_thresh=st.sampled_from([None, 0.3]), pre=st.sampled_from([1000, 7]), post=st.sampled_from([100, 3]))
def test_nms_random_scenes_against_the_oracle(seed, n_obj, per_obj, thresh, degenerate, score_thresh, pre, post):
    import pcp_b200
    from oracle import nms_oracle as nmo
    b = _boxes(seed, max(n_obj, 6 if degenerate else 1), per_obj, degenerate)
    bn = b.numpy().astype(np.float64)
    # the product follows the reference kernel's procedure (corner test with a 1e-2 m margin): the oracle's float32
    # restatement of that procedure matches it to rounding, the exact float64 area to the kernel's own error
    iou = nmo.ref_iou_f32(b.numpy()[:, :7], b.numpy()[:, :7])
    got_iou = pcp_b200.boxes_iou_bev(b[:, :7].contiguous().to(DEV), b[:, :7].contiguous().to(DEV)).cpu().numpy()
    assert np.abs(got_iou - iou).max() < 1e-5
    assert np.abs(got_iou - nmo.boxes_iou_bev(bn[:, :7], bn[:, :7])).max() < 5e-3
    off = iou[~np.eye(len(iou), dtype=bool)]
    if off.size and np.any(np.abs(off - thresh) < 1e-4):
        return                                                    # a pair sits on the threshold: the last bits decide
    cfg = pcp_b200.CfgDict(NMS_TYPE="nms_gpu", NMS_THRESH=thresh, NMS_PRE_MAXSIZE=pre, NMS_POST_MAXSIZE=post)
    bd = b.to(DEV)
    sel, sc = pcp_b200.class_agnostic_nms(bd[:, 7], bd[:, :7], cfg, score_thresh=score_thresh)
    want = nmo.class_agnostic_nms(b[:, 7].numpy(), b[:, :7].numpy(), thresh, pre, post, score_thresh, iou)
    assert sel.cpu().tolist() == want.tolist()
    assert torch.equal(sc, bd[sel, 7])
