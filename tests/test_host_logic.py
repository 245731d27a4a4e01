"""CPU: host-side logic of the drop-in - constructor arithmetic, state_dict names, registries, sharding
(world_size-2 gloo), latency -> flow scale."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from tests.helpers import golden_cfg, load_golden, model_cfgs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_vfe(c_raw=5, **opts):
    import pcp_b200
    from pcp_b200 import synthetic as syn
    rng = np.asarray(syn.V2X_RANGE, dtype=np.float32)
    grid = syn.grid_size_of(rng, syn.V2X_VOXEL)
    vfe_cfg, scat_cfg = model_cfgs(c_raw, **opts)
    vfe = pcp_b200.DynamicPillarVFE(model_cfg=vfe_cfg, num_point_features=99, voxel_size=syn.V2X_VOXEL, grid_size=grid,
                                    point_cloud_range=rng, depth_downsample_factor=None)
    scat = pcp_b200.PointPillarScatter(model_cfg=scat_cfg, grid_size=grid)
    return vfe, scat


def test_constructor_matches_reference_attributes():
    vfe, scat = make_vfe()
    assert vfe.num_raw_point_features == 5                      # NUM_RAW_POINT_FEATURES overrides the argument (:53-54)
    assert vfe.get_output_feature_dim() == 64 and scat.num_bev_features == 64
    assert (scat.nx, scat.ny, scat.nz) == (512, 512, 1)
    assert vfe.scale_xy == 512 * 512 and vfe.scale_y == 512
    # offsets evaluated like the reference: python float / 2 + np.float32 -> np.float32 under numpy >= 2
    assert vfe.x_offset == 0.2 / 2 + np.float32(-51.2) and vfe.z_offset == 8.0 / 2 + np.float32(-8.0)
    g = load_golden("vfe_car_small")
    cfg, *_ = golden_cfg(g)
    assert np.float32(vfe.x_offset) == np.float32(cfg.x_offset)


def test_state_dict_names_are_the_reference_ones():
    vfe, _ = make_vfe()
    keys = set(vfe.state_dict().keys())
    want = {"pfn_layers.0.linear.weight", "pfn_layers.1.linear.weight"}
    for i in (0, 1):
        want |= {f"pfn_layers.{i}.norm.{k}" for k in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked")}
    assert keys == want
    assert vfe.pfn_layers[0].linear.weight.shape == (32, 11) and vfe.pfn_layers[1].linear.weight.shape == (64, 64)
    assert vfe.pfn_layers[0].norm.eps == 1e-3 and vfe.pfn_layers[0].norm.momentum == 0.01
    g = load_golden("vfe_car_small")
    res = vfe.load_state_dict(g["sd"])
    assert not res.missing_keys and not res.unexpected_keys
    nonorm, _ = make_vfe(use_norm=False)
    assert set(nonorm.state_dict()) == {"pfn_layers.0.linear.weight", "pfn_layers.0.linear.bias",
                                        "pfn_layers.1.linear.weight", "pfn_layers.1.linear.bias"}
    ego, _ = make_vfe(c_raw=11)
    assert ego.pfn_layers[0].linear.weight.shape == (32, 17)


def test_cpu_inputs_and_training_mode_are_rejected_not_emulated():
    vfe, scat = make_vfe()
    vfe.eval()
    with pytest.raises(RuntimeError, match="no CPU path"):
        vfe({"points": torch.zeros(4, 8), "batch_size": 1})
    with pytest.raises(RuntimeError, match="no CPU path"):
        scat({"pillar_features": torch.zeros(1, 64), "voxel_coords": torch.zeros(1, 4, dtype=torch.int32)})
    vfe.train()
    with pytest.raises(RuntimeError, match="inference-only"):
        vfe({"points": torch.zeros(4, 8), "batch_size": 1})
    with pytest.raises(AssertionError):
        import pcp_b200
        pcp_b200.PointPillarScatter(model_cfg=pcp_b200.CfgDict(NUM_BEV_FEATURES=64), grid_size=(512, 512, 2))


def test_registry_patch():
    from pcp_b200 import registry, modules
    vfe_reg, bev_reg = {"DynPillarVFE": object, "MeanVFE": int}, {"PointPillarScatter": object}
    registry.patch_pcdet(vfe_reg, bev_reg)
    assert vfe_reg["DynPillarVFE"] is modules.DynamicPillarVFE and vfe_reg["MeanVFE"] is int
    assert bev_reg["PointPillarScatter"] is modules.PointPillarScatter


def test_cfgdict_behaves_like_easydict():
    from pcp_b200 import CfgDict
    c = CfgDict(USE_NORM=True)
    assert c.USE_NORM and c.get("MISSING", None) is None
    with pytest.raises(AttributeError):
        c.MISSING


def test_flow_scale():
    from pcp_b200.modar import flow_scale
    assert flow_scale(0.0, 0.2) == 2.0                           # the literal in v2x_sim_dataset_ego.py:213
    assert flow_scale(1234.5, 1234.7) == 2.0                     # float timestamps snap to whole intervals
    assert flow_scale(7.0, 7.0) == 0.0                           # EXCHANGE_NOW
    assert flow_scale(0.0, 0.4) == 4.0 and abs(flow_scale(0.0, 0.1) - 1.0) < 1e-12
    assert abs(flow_scale(0.0, 0.13) - 1.3) < 1e-12
    with pytest.raises(ValueError):
        flow_scale(1.0, 0.5)


def test_frame_ranges_partition_the_batch():
    from pcp_b200.sharding import frame_range
    for frames in (1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [frame_range(frames, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == frames
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [l - f for f, l in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shard_points_renumbers_frames():
    from pcp_b200 import synthetic as syn
    from pcp_b200.sharding import shard_points
    pts = syn.batch_of_frames(5, 50, 4)
    seen = 0
    for r in range(2):
        local, nf = shard_points(pts, 5, r, 2)
        assert nf == (3, 2)[r] and local.shape[0] == nf * 50
        assert set(local[:, 0].tolist()) == set(float(i) for i in range(nf))
        first = (0, 3)[r]
        assert torch.equal(local[:, 1:], pts[(pts[:, 0] >= first) & (pts[:, 0] < first + nf)][:, 1:])
        seen += local.shape[0]
    assert seen == pts.shape[0]


def _gather_worker(rank, world, port, frames, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from pcp_b200.sharding import frame_range, gather_bev
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    f, l = frame_range(frames, rank, world)
    local = torch.stack([torch.full((2, 3, 4), float(b)) for b in range(f, l)]) if l > f else torch.zeros(0, 2, 3, 4)
    full = gather_bev(local, frames)
    ok = full.shape == (frames, 2, 3, 4) and all(float(full[b].mean()) == float(b) for b in range(frames))
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("frames", [4, 5])
def test_gather_bev_world_size_2_gloo(frames):
    """The only collective of the path (validation all-gather of per-rank BEV blocks), on CPU with gloo."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + frames) % 2000
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_synthetic_generators_are_seeded_and_shaped():
    from pcp_b200 import synthetic as syn
    a, b = syn.lidar_frame(1000, 5), syn.lidar_frame(1000, 5)
    assert torch.equal(a, b) and a.shape == (1000, 8)
    e = syn.lidar_frame(100, 5, ego_columns=True)
    assert e.shape == (100, 14) and float(e[:, 6:12].abs().max()) == 0.0
    outside = ((a[:, 1].abs() > 51.2) | (a[:, 2].abs() > 51.2)).float().mean()
    assert 0.0 < float(outside) < 0.05
    ego, agents = syn.modar_scene(2, 0, n_agents=2, n_ego_points=64)
    assert ego.shape == (64, 14) and all(ag["modar"].shape[1] == 9 and ag["foreground"].shape[1] == 13 for ag in agents)
    assert list(syn.grid_size_of(syn.V2X_RANGE, syn.STRESS_VOXEL)) == [1024, 1024, 1]


def test_collate_points_replaces_the_points_branch_of_collate_batch():
    """pcp_b200.collate_points: the frame-index padding + concatenation of collate_batch (dataset.py:224-229) and the
    .float() of load_data_to_gpu, into one (pinned) buffer, optionally a subset of the columns."""
    import numpy as np
    from pcp_b200.loader import collate_points
    g = np.random.default_rng(0)
    frames = [g.normal(size=(n, 7)) for n in (5, 0, 11)]                 # float64, as the dataset returns them
    p = collate_points(frames)
    want = np.concatenate([np.pad(f, ((0, 0), (1, 0)), mode="constant", constant_values=i) for i, f in enumerate(frames)])
    assert p.columns == tuple(range(7)) and p.n_point_cols == 7 and p.batch_size == 3 and p.n_points == 16
    assert p.frame_offsets.tolist() == [0, 5, 5, 16]
    assert np.array_equal(p.data.numpy(), want[:, 1:].astype(np.float32))
    sub = collate_points(frames, columns=(0, 1, 2, 3, 4))
    assert np.array_equal(sub.data.numpy(), want[:, 1:6].astype(np.float32))
    again = collate_points([f[:3] for f in frames if len(f)], columns=(0, 1, 2, 3, 4), out=sub)      # refill in place
    assert again.data.data_ptr() == sub.data.data_ptr() and again.frame_offsets.tolist() == [0, 3, 6]
    import pytest
    with pytest.raises(ValueError):
        collate_points(frames, columns=(0, 0))
    with pytest.raises(ValueError):
        collate_points(frames, columns=(7,))
