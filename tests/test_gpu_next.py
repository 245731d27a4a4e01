"""GPU parity of the SURVEY 8(f) rows (bev_scatter / interpolate, DynamicMeanVFE, DynamicPillarVFESimple2D, early-fusion
assembly): the sm_100a kernels, called through the drop-in Python surface / C ABI, against the committed outputs of the
reference's own code (tests/golden/next_*.npz) and against the CPU oracle on larger seeded inputs and edge cases.

Bars: coordinates / indices / occupancy bit-exact; bilinear interpolation, per-pixel and per-voxel means bit-exact (every
product and sum is rounded separately in the reference's order; sums run in ascending row order like the CPU scatter_mean);
PFN outputs within 1e-5 relative; SE(3) transform: fp64 evaluation rounded once to fp32 (bit-exact up to the last-bit
ambiguity of a BLAS FMA, allowed on < 1e-5 of the values, 1 ulp).
"""
import os

import numpy as np
import pytest
import torch

from oracle import next_oracle as no
from oracle import pillar_oracle as po
from tests.helpers import GOLDEN_DIR, assert_features_close, layers_from_state_dict

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: z[k] for k in z.files}


# ------------------------------------------------------------------------------------------------ hunter_toolbox
def test_interpolate_golden():
    import pcp_b200
    g = _load("next_hunter_small")
    feat, coord = pcp_b200.interpolate_points_feat_from_bev_img(torch.from_numpy(g["bev_img"]).to(DEV), torch.from_numpy(g["points"]).to(DEV),
                                                                torch.from_numpy(g["range6"]).to(DEV), torch.from_numpy(g["pixel"]).to(DEV),
                                                                return_bev_coord=True)
    assert np.array_equal(coord.cpu().numpy(), g["ref_bev_coord"])
    assert np.array_equal(feat.cpu().numpy(), g["ref_points_feat"])
    only = pcp_b200.interpolate_points_feat_from_bev_img(torch.from_numpy(g["bev_img"]).to(DEV), torch.from_numpy(g["points"]).to(DEV),
                                                         g["range6"], g["pixel"])
    assert torch.equal(only, feat)


def test_bev_scatter_golden():
    import pcp_b200
    g = _load("next_hunter_small")
    pts = torch.from_numpy(g["points"])
    h, w = g["bev_img"].shape[2:]
    bev = pcp_b200.bev_scatter(torch.from_numpy(g["ref_bev_coord"]).to(DEV), pts[:, 0].long().to(DEV),
                               torch.from_numpy(g["points_feat"]).to(DEV), (h, w))
    assert np.array_equal(bev.cpu().numpy(), g["ref_bev_scatter"])


@pytest.mark.parametrize("n,b,c,h,w", [(50000, 4, 64, 128, 128), (3000, 1, 7, 33, 47), (1, 1, 3, 4, 4), (20000, 2, 32, 256, 256)])
def test_hunter_round_trip_vs_oracle(n, b, c, h, w):
    """image -> points -> image on seeded inputs; odd channel counts and image sizes; points on and outside the border."""
    import pcp_b200
    gen = torch.Generator().manual_seed(n + c)
    img = torch.randn(b, c, h, w, generator=gen)
    pts = torch.zeros(n, 6)
    pts[:, 0] = torch.randint(0, b, (n,), generator=gen).float()
    pts[:, 0][0] = b - 1                                           # the last frame is present
    pix = np.asarray([0.4, 0.8], dtype=np.float32)
    rng = np.asarray([-0.4 * w / 2, -0.8 * h / 2, -3.0], dtype=np.float32)
    pts[:, 1] = (torch.rand(n, generator=gen) * 1.1 - 0.55) * (0.4 * w)
    pts[:, 2] = (torch.rand(n, generator=gen) * 1.1 - 0.55) * (0.8 * h)
    feat, coord = pcp_b200.interpolate_points_feat_from_bev_img(img.to(DEV), pts.to(DEV), rng, pix, return_bev_coord=True)
    wf, wc = no.interpolate_points_feat_from_bev_img(img, pts, rng[:2], pix)
    assert torch.equal(coord.cpu(), wc)
    assert torch.equal(feat.cpu(), wf)
    bev = pcp_b200.bev_scatter(coord, pts[:, 0].long().to(DEV), feat, (h, w))
    want = no.bev_scatter(wc, pts[:, 0].long(), wf, (h, w))
    assert tuple(bev.shape) == tuple(want.shape)
    assert torch.equal(bev.cpu(), want)


def test_bev_scatter_edges():
    import pcp_b200
    # every point outside the image: an all-zero image of the right shape; frame index of a culled point still counts (:74)
    coord = torch.tensor([[-1.0, 2.0], [0.0, 0.5], [4.0, 1.0], [1.0, 4.0], [float("nan"), 1.0]])
    bidx = torch.tensor([0, 2, 1, 0, 1])
    feat = torch.ones(5, 3)
    bev = pcp_b200.bev_scatter(coord.to(DEV), bidx.to(DEV), feat.to(DEV), (4, 4))
    assert tuple(bev.shape) == (3, 3, 4, 4) and float(bev.abs().sum()) == 0.0
    # many points in one pixel: sequential sum in row order
    gen = torch.Generator().manual_seed(3)
    for n in (4000, 9000):
        coord = torch.rand(n, 2, generator=gen) * 0.9 + 1.05      # all in pixel (1, 1)
        feat = torch.randn(n, 5, generator=gen) * 100
        bidx = torch.zeros(n, dtype=torch.long)
        bev = pcp_b200.bev_scatter(coord.to(DEV), bidx.to(DEV), feat.to(DEV), (3, 3), batch_size=1)
        want = no.bev_scatter(coord, bidx, feat, (3, 3))
        assert torch.equal(bev.cpu(), want)                       # rows of a cell are visited in ascending order, any count


# ------------------------------------------------------------------------------------------------ DynamicMeanVFE
def _mean_vfe(points, c, vox, rng, grid, batch_size=None):
    import pcp_b200
    m = pcp_b200.DynamicMeanVFE(model_cfg=pcp_b200.CfgDict(), num_point_features=c, voxel_size=vox, grid_size=grid,
                                point_cloud_range=rng)
    bd = {"points": points.to(DEV)}
    if batch_size is not None:
        bd["batch_size"] = batch_size
    return m(bd)


def test_dynamic_mean_vfe_golden():
    g = _load("next_meanvfe_small")
    vox = [float(v) for v in g["voxel_size"]]
    bd = _mean_vfe(torch.from_numpy(g["points"]), int(g["num_point_features"]), vox, g["point_cloud_range"], g["grid_size"])
    assert bd["voxel_coords"].dtype == torch.int32
    assert np.array_equal(bd["voxel_coords"].cpu().numpy(), g["ref_voxel_coords"])
    assert np.array_equal(bd["voxel_features"].cpu().numpy(), g["ref_voxel_features"])


@pytest.mark.parametrize("n,frames,c,ego", [(60000, 2, 5, False), (40000, 3, 11, True), (200, 1, 4, False)])
def test_dynamic_mean_vfe_vs_oracle(n, frames, c, ego):
    """SECOND-style grid (v2x_second_car.yaml: 0.1 m x 0.1 m x 0.2 m -> 1024 x 1024 x 40)."""
    from pcp_b200 import synthetic as syn
    vox = [0.1, 0.1, 0.2]
    rng = np.asarray(syn.V2X_RANGE, dtype=np.float32)
    grid = syn.grid_size_of(rng, vox)
    pts = syn.batch_of_frames(frames, n, 11, ego_columns=ego)
    pts[::7, 3] += 1.0                                             # some z above the range: culled here
    bd = _mean_vfe(pts, c, vox, rng, grid, batch_size=frames)
    want = no.dynamic_mean_vfe(pts, c, vox, rng, grid)
    assert torch.equal(bd["voxel_coords"].cpu(), want["voxel_coords"])
    assert torch.equal(bd["voxel_features"].cpu(), want["voxel_features"])


def test_dynamic_mean_vfe_dense_pillar_and_empty():
    """hundreds of points in one pillar spread over every z cell; then a batch where everything is culled."""
    vox, rng = [0.2, 0.2, 0.5], np.asarray([-0.8, -0.8, -8.0, 0.8, 0.8, 0.0], dtype=np.float32)
    grid = np.asarray([8, 8, 16])
    gen = torch.Generator().manual_seed(1)
    n = 3000
    pts = torch.zeros(n, 6)
    pts[:, 1:3] = torch.rand(n, 2, generator=gen) * 0.19 + 0.2
    pts[:, 3] = torch.rand(n, generator=gen) * -8.0
    pts[:, 4:] = torch.randn(n, 2, generator=gen)
    bd = _mean_vfe(pts, 5, vox, rng, grid)
    want = no.dynamic_mean_vfe(pts, 5, vox, rng, grid)
    assert torch.equal(bd["voxel_coords"].cpu(), want["voxel_coords"])
    assert torch.equal(bd["voxel_features"].cpu(), want["voxel_features"])
    far = pts.clone()
    far[:, 1] += 10.0
    bd = _mean_vfe(far, 5, vox, rng, grid, batch_size=1)
    assert bd["voxel_coords"].shape == (0, 4) and bd["voxel_features"].shape == (0, 5)


# ------------------------------------------------------------------------------------------------ Simple2D
@pytest.mark.parametrize("tag", ["abs", "rel_dist"])
def test_simple2d_golden(tag):
    import pcp_b200
    g = _load(f"next_simple2d_{tag}")
    sd = {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd/")}
    cfg = pcp_b200.CfgDict(USE_NORM=True, WITH_DISTANCE=bool(g["with_distance"]), USE_ABSLOTE_XYZ=bool(g["use_abs"]), NUM_FILTERS=[64, 64])
    m = pcp_b200.DynamicPillarVFESimple2D(model_cfg=cfg, num_point_features=7, voxel_size=[float(v) for v in g["voxel_size"]],
                                          grid_size=g["grid_size"], point_cloud_range=g["point_cloud_range"])
    missing = m.load_state_dict(sd)
    assert not missing.missing_keys and not missing.unexpected_keys
    m = m.to(DEV).eval()
    with torch.no_grad():
        bd = m({"points": torch.from_numpy(g["points"]).to(DEV)})
    assert "voxel_coords" not in bd and "voxel_features" not in bd
    assert np.array_equal(bd["pillar_coords"].cpu().numpy(), g["ref_pillar_coords"])
    assert_features_close(bd["pillar_features"].cpu().numpy(), g["ref_pillar_features"], "pillar_features")


# ------------------------------------------------------------------------------------------------ early fusion
def _close_fp64_rounded(got, want):
    assert got.shape == want.shape
    same = got == want
    frac = 1.0 - float(same.mean())
    assert frac < 1e-5, f"{frac:.2e} of the values differ"
    if frac:
        ulp = np.abs(got[~same].view(np.int32).astype(np.int64) - want[~same].view(np.int32).astype(np.int64))
        assert ulp.max() <= 1


def test_early_fusion_golden():
    import pcp_b200
    g = _load("next_early_fusion_small")
    clouds = [torch.from_numpy(g[f"cloud{a}"]).to(DEV) for a in range(4)]
    tfs = [g[f"se3_{a}"] for a in range(1, 4)]
    fused = pcp_b200.fuse_agent_points(clouds[0], clouds[1:], tfs, g["range6"], batch_idx=3).cpu().numpy()
    assert np.all(fused[:, 0] == 3.0)
    assert fused.shape[0] == g["ref_fused"].shape[0]
    _close_fp64_rounded(fused[:, 1:], g["ref_fused"])
    nomask = pcp_b200.fuse_agent_points(clouds[0], clouds[1:], tfs, None, batch_idx=None).cpu().numpy()
    _close_fp64_rounded(nomask, g["ref_fused_nomask"])


def test_early_fusion_ragged_agents():
    """An agent that sent nothing, an empty ego cloud, non-contiguous / float64 inputs: the per-agent pointer path
    (pcp_fuse_agent_clouds) against the oracle."""
    import pcp_b200
    from pcp_b200 import synthetic as syn
    rng = np.asarray(syn.V2X_RANGE, dtype=np.float32)
    clouds = [syn.lidar_frame(n, 8300 + a)[:, 1:].contiguous() for a, n in enumerate((500, 0, 1234, 1))]
    tfs = [syn.modar_agent(8400 + a, n_boxes=1)["target_se3_agent"] for a in range(3)]
    want = no.fuse_agent_points(clouds[0].numpy(), [c.numpy() for c in clouds[1:]], tfs, rng)
    got = pcp_b200.fuse_agent_points(clouds[0].to(DEV), [clouds[1].to(DEV), clouds[2].double().to(DEV), clouds[3].to(DEV)], tfs, rng,
                                     batch_idx=0)
    assert got.shape == (want.shape[0], 8)
    _close_fp64_rounded(got.cpu().numpy()[:, 1:], want)
    # empty ego, strided agent rows (a column slice of a wider tensor)
    wide = torch.cat([clouds[2], torch.zeros(clouds[2].shape[0], 3)], dim=1).to(DEV)
    want2 = no.fuse_agent_points(clouds[1].numpy(), [clouds[2].numpy()], tfs[:1], rng)
    got2 = pcp_b200.fuse_agent_points(clouds[1].to(DEV), [wide[:, :7]], tfs[:1], rng, batch_idx=None)
    assert got2.shape == want2.shape
    _close_fp64_rounded(got2.cpu().numpy(), want2)


def test_early_fusion_feeds_the_pillar_path():
    """BASELINE configs[2] shape in small: 6 clouds fused on the GPU, then DynamicPillarVFE + PointPillarScatter; checked
    against the oracle chain on the CPU-fused cloud."""
    import pcp_b200
    from pcp_b200 import synthetic as syn
    from tests.helpers import model_cfgs
    rng = np.asarray(syn.V2X_RANGE, dtype=np.float32)
    vox = syn.V2X_VOXEL
    grid = syn.grid_size_of(rng, vox)
    clouds = [syn.lidar_frame(20000, 8100 + a)[:, 1:].contiguous() for a in range(6)]
    tfs = [syn.modar_agent(8200 + a, n_boxes=1)["target_se3_agent"] for a in range(5)]
    fused = pcp_b200.fuse_agent_points(clouds[0].to(DEV), [c.to(DEV) for c in clouds[1:]], tfs, rng, batch_idx=0)
    want_pts = no.fuse_agent_points(clouds[0].numpy(), [c.numpy() for c in clouds[1:]], tfs, rng)
    assert fused.shape == (want_pts.shape[0], 8)
    got = fused.cpu().numpy()
    _close_fp64_rounded(got[:, 1:], want_pts)
    vfe_cfg, scat_cfg = model_cfgs(5)
    sd = syn.pfn_state_dict(11)
    vfe = pcp_b200.DynamicPillarVFE(model_cfg=vfe_cfg, num_point_features=5, voxel_size=vox, grid_size=grid, point_cloud_range=rng)
    vfe.load_state_dict(sd)
    vfe = vfe.to(DEV).eval()
    scat = pcp_b200.PointPillarScatter(model_cfg=scat_cfg, grid_size=grid).to(DEV).eval()
    with torch.no_grad():
        bd = scat(vfe({"points": fused, "batch_size": 1}))
    want = po.front_end(fused.cpu(), po.VFEConfig(5, vox, rng, grid), layers_from_state_dict(sd), unique_dim0=False)
    assert torch.equal(bd["voxel_coords"].cpu(), want["voxel_coords"])
    assert_features_close(bd["pillar_features"].cpu().numpy(), want["pillar_features"].numpy(), "pillar_features")
    assert torch.equal((bd["spatial_features"].cpu() != 0).any(1), (want["spatial_features"] != 0).any(1))


def test_load_points_to_gpu_rebuilds_collate_batch_rows():
    """SURVEY 8f rank 3, host->device half: collate_points + load_points_to_gpu against the reference's collate_batch
    (dataset.py:224-229) + load_data_to_gpu (models/__init__.py:23-34) restated with numpy / torch."""
    import pcp_b200
    g = np.random.default_rng(3)
    frames = [g.normal(size=(n, 7)) * 20 for n in (4000, 1, 0, 2500)]
    want = torch.from_numpy(np.concatenate([np.pad(f, ((0, 0), (1, 0)), mode="constant", constant_values=i)
                                            for i, f in enumerate(frames)])).float()
    full = pcp_b200.load_points_to_gpu(pcp_b200.collate_points(frames), DEV)
    torch.cuda.synchronize()
    assert torch.equal(full.cpu(), want)
    packed = pcp_b200.collate_points(frames, columns=(0, 1, 2, 3, 4))
    staging = {}
    out = torch.full((7000, 8), 7.0, device=DEV)
    got = pcp_b200.load_points_to_gpu(packed, DEV, out=out, staging=staging)
    torch.cuda.synchronize()
    assert got.shape == (6501, 8) and got.data_ptr() == out.data_ptr()
    assert torch.equal(got[:, :6].cpu(), want[:, :6]) and float(got[:, 6:].abs().max()) == 0.0
    assert float(out[6501:].min()) == 7.0                                  # rows past N untouched
    # the array collate_batch built (float64) through the plain path
    plain = pcp_b200.load_points_to_gpu(want.double().numpy(), DEV)
    assert torch.equal(plain.cpu(), want)
    # the pillar path gives the same result on the shipped-columns rows as on the full rows
    from pcp_b200 import synthetic as syn
    rng = np.asarray(syn.V2X_RANGE, dtype=np.float32)
    grid = syn.grid_size_of(rng, syn.V2X_VOXEL)
    vfe_cfg, _ = syn.model_cfgs(5)
    vfe = pcp_b200.DynamicPillarVFE(model_cfg=vfe_cfg, num_point_features=5, voxel_size=syn.V2X_VOXEL, grid_size=grid,
                                    point_cloud_range=rng)
    vfe.load_state_dict(syn.pfn_state_dict(11))
    vfe = vfe.to(DEV).eval()
    a = vfe({"points": full, "batch_size": 4})
    b = vfe({"points": got, "batch_size": 4})
    assert torch.equal(a["voxel_coords"], b["voxel_coords"]) and torch.equal(a["pillar_features"], b["pillar_features"])


def test_points_prefetcher_streams_batches_in_order():
    import pcp_b200
    g = np.random.default_rng(9)
    batches = []
    for j in range(7):
        frames = [g.normal(size=(int(g.integers(0, 3000)), 7)) * 30 for _ in range(3)]
        batches.append((pcp_b200.collate_points(frames, columns=(0, 1, 2, 4)), frames))
    pre = pcp_b200.PointsPrefetcher(DEV, depth=2)
    got = []
    pre.submit(batches[0][0])
    for j in range(7):
        if j + 1 < 7:
            pre.submit(batches[j + 1][0])
        got.append(pre.get().clone())
    torch.cuda.synchronize()
    for (packed, frames), rows in zip(batches, got):
        want = torch.from_numpy(np.concatenate([np.pad(f, ((0, 0), (1, 0)), mode="constant", constant_values=i)
                                                for i, f in enumerate(frames)])).float()
        assert rows.shape == want.shape
        for c in (0, 1, 2, 3, 5):
            assert torch.equal(rows[:, c].cpu(), want[:, c])
        assert float(rows[:, [4, 6, 7]].abs().max()) == 0.0 if rows.shape[0] else True
    with pytest.raises(RuntimeError):
        pre.get()


@pytest.mark.parametrize("n,frames,sorted_by_frame", [(50000, 4, True), (3001, 3, False), (0, 2, True), (7, 5, True)])
def test_select_foreground_matches_the_reference_block(n, frames, sorted_by_frame):
    """pcp_b200.select_foreground vs the restated hunter_jr.py:377-397 block: the same rows, in the same order, per sample;
    probabilities to 1e-6 (expf of the GPU vs the CPU's vectorised exp), everything else bit for bit.  Rows within 2e-6 of
    the 0.3 threshold are steered away from it (a last-bit difference of sigmoid may flip them)."""
    import pcp_b200
    from oracle import next_oracle as no
    g = torch.Generator().manual_seed(n + frames)
    pts = torch.randn(n, 8, generator=g)
    b = torch.randint(0, frames, (n,), generator=g).float()
    pts[:, 0] = torch.sort(b).values if sorted_by_frame else b
    if n > 5:
        pts[3, 0] = float(frames)                      # a row of a sample the batch does not have: never sent
    logit = torch.randn(n, 3, generator=g) * 2
    p0 = torch.sigmoid(logit[:, 0])
    logit[(p0 - 0.3).abs() < 2e-6, 0] += 0.01
    flow = torch.randn(n, 3, generator=g)
    want = no.select_foreground(pts, logit, flow, frames)
    got = pcp_b200.select_foreground(pts.to(DEV), logit.to(DEV), flow.to(DEV), frames)
    assert len(got) == frames
    for w, gt in zip(want, got):
        gt = gt.cpu()
        assert gt.shape == w.shape
        assert torch.equal(gt[:, :7], w[:, :7]) and torch.equal(gt[:, 10:], w[:, 10:])
        assert torch.allclose(gt[:, 7:10], w[:, 7:10], rtol=0, atol=1e-6)


def test_exchange_payloads_round_trip_through_the_wire_format():
    """select_foreground + exchange_payloads -> unpack_exchange -> modar_exchange consumes the records unchanged."""
    import pcp_b200
    from pcp_b200 import exchange
    g = torch.Generator().manual_seed(3)
    n = 4000
    pts = torch.randn(n, 8, generator=g)
    pts[:, 0] = torch.sort(torch.randint(0, 2, (n,), generator=g).float()).values
    logit, flow = torch.randn(n, 3, generator=g) * 2, torch.randn(n, 3, generator=g)
    fg = pcp_b200.select_foreground(pts.to(DEV), logit.to(DEV), flow.to(DEV), 2)
    preds = [{"pred_boxes": torch.randn(5, 7, generator=g).to(DEV), "pred_scores": torch.rand(5, generator=g).to(DEV),
              "pred_labels": torch.ones(5).to(DEV)},
             {"pred_boxes": torch.zeros(0, 7).to(DEV), "pred_scores": torch.zeros(0).to(DEV), "pred_labels": torch.zeros(0).to(DEV)}]
    msgs = exchange.exchange_payloads(preds, fg, agent_ids=[1, 2], timestamps=[0.2, 0.2])
    assert msgs[1] is None                              # no boxes: nothing is sent (center_head.py:413)
    m = exchange.unpack_exchange(msgs[0])
    assert m.agent_id == 1 and m.boxes.shape == (5, 9) and torch.equal(m.foreground, fg[0])
    assert torch.equal(m.boxes[:, :7], preds[0]["pred_boxes"])
