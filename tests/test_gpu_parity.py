"""GPU parity: the sm_100a kernels, called through the drop-in modules / C ABI, against the committed
reference outputs (tests/golden) and against the CPU oracle on seeded inputs.

Bars (BASELINE.json north_star): voxel_coords, point->pillar map and BEV occupancy bit-exact; per-pillar
mean bit-exact (sums run in the reference CPU order); canvas = exact copy of pillar_features; PFN outputs
within 1e-5 relative (fp32; absolute floor 1e-5 x tensor scale for values near zero).
"""
import numpy as np
import pytest
import torch

from oracle import pillar_oracle as po
from tests.helpers import (assert_features_close, golden_cases, golden_cfg, layers_from_state_dict, load_golden,
                           model_cfgs)

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


def build_modules(c_raw, vox, rng, grid, sd, **opts):
    import pcp_b200
    vfe_cfg, scat_cfg = model_cfgs(c_raw, **opts)
    vfe = pcp_b200.DynamicPillarVFE(model_cfg=vfe_cfg, num_point_features=c_raw, voxel_size=vox, grid_size=grid,
                                    point_cloud_range=rng)
    missing = vfe.load_state_dict(sd)
    assert not missing.missing_keys and not missing.unexpected_keys
    scat = pcp_b200.PointPillarScatter(model_cfg=scat_cfg, grid_size=grid)
    return vfe.to(DEV).eval(), scat.to(DEV).eval()


def run_modules(vfe, scat, points, batch_size=None):
    bd = {"points": points.to(DEV)}
    if batch_size is not None:
        bd["batch_size"] = batch_size
    with torch.no_grad():
        bd = scat(vfe(bd))
    torch.cuda.synchronize()
    return bd


def check_against(bd, want, n_points):
    """want: oracle-style dict of CPU tensors."""
    from pcp_b200.modules import CTX_KEY
    vc = bd["voxel_coords"].cpu()
    assert vc.dtype == torch.int32
    assert torch.equal(vc, want["voxel_coords"]), "voxel_coords differ"
    pp = bd[CTX_KEY]["point_pillar"].cpu().long()
    assert pp.shape[0] == n_points
    assert torch.equal(pp[pp >= 0], want["unq_inv"].long()), "point->pillar map differs"
    if "keep_mask" in want:
        assert torch.equal(pp >= 0, want["keep_mask"]), "cull mask differs"
    assert bd["voxel_features"] is bd["pillar_features"]
    pf = bd["pillar_features"].cpu()
    assert_features_close(pf.numpy(), want["pillar_features"].numpy(), "pillar_features")
    sf = bd["spatial_features"].cpu()
    assert tuple(sf.shape) == tuple(want["canvas_shape"])
    occ = torch.nonzero((sf != 0).any(dim=1).flatten()).flatten()
    assert torch.equal(occ.int(), want["occupied"].int()), "BEV occupancy differs"
    # the canvas is a pure copy: must equal the reference scatter of OUR pillar features bit for bit
    grid = (sf.shape[3], sf.shape[2], 1)
    assert torch.equal(sf, po.pointpillar_scatter(pf, vc, grid, pf.shape[1])), "canvas is not an exact copy"


@pytest.mark.parametrize("name", golden_cases())
def test_golden_reference_outputs(name):
    g = load_golden(name)
    cfg, rng, vox, grid = golden_cfg(g)
    vfe, scat = build_modules(int(g["c_raw"]), vox, rng, grid, g["sd"], num_filters=tuple(int(v) for v in g["num_filters"]),
                              use_norm=bool(g["use_norm"]), with_distance=bool(g["with_distance"]), use_abs=bool(g["use_abs"]))
    pts = torch.from_numpy(g["points"])
    nb = int(np.nanmax(np.where(np.isfinite(g["points"][:, 0]), g["points"][:, 0], 0))) + 1
    bd = run_modules(vfe, scat, pts, batch_size=nb)
    want = {"voxel_coords": torch.from_numpy(g["voxel_coords"]), "unq_inv": torch.from_numpy(g["unq_inv"]),
            "pillar_features": torch.from_numpy(g["pillar_features"]), "canvas_shape": g["canvas_shape"],
            "occupied": torch.from_numpy(g["occupied"])}
    check_against(bd, want, pts.shape[0])
    if "spatial_features" in g:
        assert_features_close(bd["spatial_features"].cpu().numpy(), g["spatial_features"], "spatial_features")


def oracle_want(points, cfg, layers):
    out = po.front_end(points, cfg, layers, unique_dim0=False)
    sf = out["spatial_features"]
    out["canvas_shape"] = tuple(sf.shape)
    out["occupied"] = torch.nonzero((sf != 0).any(dim=1).flatten()).flatten()
    return out


def v2x_setup(c_raw, voxel=None, seed=0):
    from pcp_b200 import synthetic as syn
    voxel = voxel or syn.V2X_VOXEL
    rng = np.asarray(syn.V2X_RANGE, dtype=np.float32)
    grid = syn.grid_size_of(rng, voxel)
    sd = syn.pfn_state_dict(c_raw + 6, (64, 64), True, seed)
    cfg = po.VFEConfig(c_raw, voxel, rng, grid)
    return syn, rng, voxel, grid, sd, cfg


@pytest.mark.parametrize("n_frames,n_points,config_id,ego", [
    (1, 32768, 1, False),     # BASELINE config 1: one 32k-point ego frame, car layout
    (1, 32768, 2, True),      # config 2 layout: 14 columns, C_raw 11
    (1, 300000, 3, False),    # config 3: early fusion ~300k points
    (4, 20000, 4, False),     # a sharded batch's local block
])
def test_seeded_configs_against_oracle(n_frames, n_points, config_id, ego):
    c_raw = 11 if ego else 5
    syn, rng, vox, grid, sd, cfg = v2x_setup(c_raw, seed=config_id)
    pts = syn.batch_of_frames(n_frames, n_points, config_id, ego_columns=ego)
    if ego:   # populate the MoDAR columns of some rows
        g = torch.Generator().manual_seed(3)
        sel = torch.randperm(pts.shape[0], generator=g)[:400]
        pts[sel, 6:12] = torch.rand(400, 6, generator=g) * 4
    vfe, scat = build_modules(c_raw, vox, rng, grid, sd)
    bd = run_modules(vfe, scat, pts, batch_size=n_frames)
    want = oracle_want(pts, cfg, layers_from_state_dict(sd))
    check_against(bd, want, pts.shape[0])


def test_batch_size_inferred_when_absent():
    syn, rng, vox, grid, sd, cfg = v2x_setup(5)
    pts = syn.batch_of_frames(3, 5000, 9)
    vfe, scat = build_modules(5, vox, rng, grid, sd)
    bd = run_modules(vfe, scat, pts, batch_size=None)
    check_against(bd, oracle_want(pts, cfg, layers_from_state_dict(sd)), pts.shape[0])


def test_pillar_mean_and_segment_reductions_bit_exact():
    """scatter_mean (sum in ascending row order / count) and scatter_max are exact, not approximate."""
    from pcp_b200.frontend import FrontEnd, GridSpec
    syn, rng, vox, grid, sd, cfg = v2x_setup(5)
    pts = syn.batch_of_frames(2, 60000, 11)
    fe = FrontEnd(GridSpec(vox, rng, grid), 5)
    l = layers_from_state_dict(sd)
    fe.pack_params(sd["pfn_layers.0.linear.weight"].to(DEV), [sd[f"pfn_layers.0.norm.{k}"].to(DEV) for k in ("weight", "bias", "running_mean", "running_var")],
                   sd["pfn_layers.1.linear.weight"].to(DEV), [sd[f"pfn_layers.1.norm.{k}"].to(DEV) for k in ("weight", "bias", "running_mean", "running_var")])
    p_dev = pts.to(DEV)
    out = fe.voxelize(p_dev, 2, want_point_pillar=True, want_counts_per_pillar=True)
    fe.pfn(p_dev, out, want_mean=True)
    counts = fe.read_counts(out)
    P = int(counts[0])
    want = po.dynamic_pillar_vfe(pts, cfg, l, unique_dim0=False)
    assert P == want["voxel_coords"].shape[0] and int(counts[1]) == want["unq_inv"].shape[0]
    assert torch.equal(out["pillar_count_buf"][:P].cpu().long(), want["unq_cnt"])
    assert int(counts[4]) == int(want["unq_cnt"].max())
    assert torch.equal(out["pillar_mean_buf"][:P].cpu(), want["points_mean"]), "pillar mean is not bit exact"
    # standalone segmented reductions over arbitrary per-point values (indexed by original row)
    g = torch.Generator().manual_seed(1)
    vals = torch.randn(pts.shape[0], 64, generator=g)
    kept = want["keep_mask"]
    got_max = fe.segment_reduce(vals.to(DEV), "max")[:P].cpu()
    got_mean = fe.segment_reduce(vals.to(DEV), "mean")[:P].cpu()
    idx = want["unq_inv"].view(-1, 1).expand(-1, 64)
    ref_max = torch.full((P, 64), float("-inf")).scatter_reduce_(0, idx, vals[kept], reduce="amax", include_self=True)
    assert torch.equal(got_max, ref_max), "segment max is not bit exact"
    assert torch.equal(got_mean, po.scatter_mean(vals[kept], want["unq_inv"], P)), "segment mean is not bit exact"


@pytest.mark.parametrize("n_in_pillar", [2, 8, 9, 12, 13, 16, 17, 24, 25, 31, 32, 33, 127, 128, 129, 500, 1024, 1025, 4096, 5000, 20000])
def test_long_pillars(n_in_pillar):
    """Pillars longer than a 128-point chunk stream through the multi-chunk path; sorted-segment code
    paths switch at 8 / 32 / 4096 points."""
    syn, rng, vox, grid, sd, cfg = v2x_setup(5, seed=2)
    g = torch.Generator().manual_seed(n_in_pillar)
    base = syn.lidar_frame(3000, 77)
    dense = syn.lidar_frame(n_in_pillar, 78)
    dense[:, 1] = 10.0 + torch.rand(n_in_pillar, generator=g) * 0.19
    dense[:, 2] = -7.0 + torch.rand(n_in_pillar, generator=g) * 0.19
    dense2 = dense.clone()
    dense2[:, 1] += 0.2                               # a second long pillar right behind the first
    pts = torch.cat([base, dense, dense2], 0)
    pts = pts[torch.randperm(pts.shape[0], generator=g)].contiguous()
    vfe, scat = build_modules(5, vox, rng, grid, sd)
    bd = run_modules(vfe, scat, pts, batch_size=1)
    want = oracle_want(pts, cfg, layers_from_state_dict(sd))
    check_against(bd, want, pts.shape[0])


@pytest.mark.parametrize("n_in_pillar", [3, 16, 17, 33, 128, 129, 700, 1024, 1025, 4096, 4097, 9000, 40000])
def test_pillar_means_bit_exact_on_every_ordering_path(n_in_pillar):
    """The per-pillar mean is a sequential fp32 sum in ascending row order on every path of pillar_prep_kernel
    (thread / half warp / warp / CTA with counting rank / CTA with bitonic network / above 4096 rows: counting sort by
    4096-row bucket, then the same network per group of buckets): bit-equal to index_add_ on the CPU."""
    from pcp_b200.frontend import FrontEnd, GridSpec
    syn, rng, vox, grid, sd, cfg = v2x_setup(5, seed=3)
    g = torch.Generator().manual_seed(100 + n_in_pillar)
    base = syn.lidar_frame(2000, 91)
    dense = syn.lidar_frame(n_in_pillar, 92)
    dense[:, 1] = 30.0 + torch.rand(n_in_pillar, generator=g) * 0.19      # far from the origin: sums are order sensitive
    dense[:, 2] = -41.0 + torch.rand(n_in_pillar, generator=g) * 0.19
    pts = torch.cat([base, dense], 0)
    pts = pts[torch.randperm(pts.shape[0], generator=g)].contiguous()
    fe = FrontEnd(GridSpec(vox, rng, grid), 5)
    bn = lambda i: [sd[f"pfn_layers.{i}.norm.{k}"].to(DEV) for k in ("weight", "bias", "running_mean", "running_var")]
    fe.pack_params(sd["pfn_layers.0.linear.weight"].to(DEV), bn(0), sd["pfn_layers.1.linear.weight"].to(DEV), bn(1))
    p_dev = pts.to(DEV)
    out = fe.voxelize(p_dev, 1, want_point_pillar=True)
    fe.pfn(p_dev, out, want_mean=True)
    P = int(fe.read_counts(out)[0])
    want = po.dynamic_pillar_vfe(pts, cfg, layers_from_state_dict(sd), unique_dim0=False)
    assert P == want["voxel_coords"].shape[0]
    assert torch.equal(out["pillar_mean_buf"][:P].cpu(), want["points_mean"]), "pillar mean is not bit exact"


def test_all_points_in_one_pillar_and_all_culled_frame():
    syn, rng, vox, grid, sd, cfg = v2x_setup(5)
    pts = syn.lidar_frame(3000, 5)
    pts[:, 1], pts[:, 2] = 0.05, 0.05
    far = syn.lidar_frame(100, 6, batch_idx=1)
    far[:, 1] = 500.0                                 # frame 1 entirely outside the range
    pts = torch.cat([pts, far], 0)
    vfe, scat = build_modules(5, vox, rng, grid, sd)
    bd = run_modules(vfe, scat, pts, batch_size=2)
    assert bd["voxel_coords"].shape[0] == 1 and bd["spatial_features"].shape[0] == 1   # B shrinks (scatter :17)
    check_against(bd, oracle_want(pts, cfg, layers_from_state_dict(sd)), pts.shape[0])


def test_generic_scatter_path_matches_reference_semantics():
    """PointPillarScatter on coords it did not produce: shuffled rows, a duplicate cell (last row wins as in the
    CPU reference's sequential index_put), channels != 64."""
    import pcp_b200
    from tests.helpers import model_cfgs
    g = torch.Generator().manual_seed(0)
    nx, ny, B, C, P = 96, 40, 3, 32, 700
    cells = torch.randperm(B * nx * ny, generator=g)[:P]
    coords = torch.stack([cells // (nx * ny), torch.zeros(P, dtype=torch.long), (cells % (nx * ny)) // nx, cells % nx], 1).int()
    coords = torch.cat([coords, coords[:5]], 0)       # 5 duplicate cells at higher row numbers
    feats = torch.randn(coords.shape[0], C, generator=g)
    scat = pcp_b200.PointPillarScatter(model_cfg=pcp_b200.CfgDict(NUM_BEV_FEATURES=C), grid_size=(nx, ny, 1))
    bd = scat({"pillar_features": feats.to(DEV), "voxel_coords": coords.to(DEV)})
    want = po.pointpillar_scatter(feats, coords, (nx, ny, 1), C)
    assert torch.equal(bd["spatial_features"].cpu(), want)


def test_results_are_deterministic_and_inputs_untouched():
    syn, rng, vox, grid, sd, cfg = v2x_setup(5)
    pts = syn.batch_of_frames(2, 100000, 12)
    vfe, scat = build_modules(5, vox, rng, grid, sd)
    p_dev = pts.to(DEV)
    a = scat(vfe({"points": p_dev, "batch_size": 2}))
    b = scat(vfe({"points": p_dev, "batch_size": 2}))
    torch.cuda.synchronize()
    assert torch.equal(p_dev.cpu(), pts)
    for k in ("pillar_features", "voxel_coords", "spatial_features"):
        assert torch.equal(a[k], b[k]), f"{k} differs between two runs"
        assert a[k].data_ptr() != b[k].data_ptr()


def test_permuting_points_permutes_nothing_in_the_outputs():
    syn, rng, vox, grid, sd, cfg = v2x_setup(5)
    pts = syn.batch_of_frames(2, 50000, 13)
    perm = torch.randperm(pts.shape[0], generator=torch.Generator().manual_seed(4))
    vfe, scat = build_modules(5, vox, rng, grid, sd)
    a = run_modules(vfe, scat, pts, 2)
    b = run_modules(vfe, scat, pts[perm].contiguous(), 2)
    assert torch.equal(a["voxel_coords"], b["voxel_coords"])
    assert_features_close(a["pillar_features"].cpu().numpy(), b["pillar_features"].cpu().numpy(), "permuted")


def test_error_behaviour():
    import pcp_b200
    syn, rng, vox, grid, sd, cfg = v2x_setup(5)
    vfe, scat = build_modules(5, vox, rng, grid, sd)
    pts = syn.batch_of_frames(2, 1000, 1)
    with pytest.raises(RuntimeError, match="no CPU path"):
        vfe({"points": pts, "batch_size": 2})
    with pytest.raises(RuntimeError, match="frame index outside"):
        vfe({"points": pts.to(DEV), "batch_size": 1})
    vfe.train()
    with pytest.raises(RuntimeError, match="inference-only"):
        vfe({"points": pts.to(DEV), "batch_size": 2})
    with pytest.raises(NotImplementedError):
        vfe.pfn_layers[0](torch.zeros(1, 11), torch.zeros(1, dtype=torch.long))
    vfe_cfg, _ = model_cfgs(5, num_filters=(64, 128, 128))
    bad = pcp_b200.DynamicPillarVFE(model_cfg=vfe_cfg, num_point_features=5, voxel_size=vox, grid_size=grid,
                                    point_cloud_range=rng).to(DEV).eval()
    with pytest.raises(NotImplementedError, match="NUM_FILTERS"):
        bad({"points": pts.to(DEV), "batch_size": 2})


def test_checkpoint_reload_repacks_parameters():
    syn, rng, vox, grid, sd, cfg = v2x_setup(5, seed=0)
    vfe, scat = build_modules(5, vox, rng, grid, sd)
    pts = syn.batch_of_frames(1, 8000, 3)
    a = run_modules(vfe, scat, pts, 1)["pillar_features"].clone()
    sd2 = syn.pfn_state_dict(11, (64, 64), True, seed=99)
    vfe.load_state_dict(sd2)
    b = run_modules(vfe, scat, pts, 1)["pillar_features"]
    want = po.dynamic_pillar_vfe(pts, cfg, layers_from_state_dict(sd2), unique_dim0=False)["pillar_features"]
    assert not torch.equal(a, b)
    assert_features_close(b.cpu().numpy(), want.numpy(), "after reload")


@pytest.mark.parametrize("n_frames,n_points,voxel,uniform", [
    (8, 300000, None, False),                 # bench workload: 8 early-fusion frames on one GPU
    (1, 4000000, [0.1, 0.1, 8.0], False),     # BASELINE config 5: 4 M points, 1024^2 canvas
    (1, 2000000, [0.1, 0.1, 8.0], True),      # uniform xy: maximum occupancy
])
def test_full_size_properties(n_frames, n_points, voxel, uniform):
    """At sizes the torch oracle does not finish quickly: integer half against numpy, occupancy == P,
    canvas column sums == pillar feature column sums (every pillar lands exactly once), determinism."""
    from pcp_b200.modules import CTX_KEY
    syn, rng, vox, grid, sd, cfg = v2x_setup(5, voxel=voxel)
    pts = syn.batch_of_frames(n_frames, n_points, 5, uniform_xy=uniform)
    vfe, scat = build_modules(5, vox, rng, grid, sd)
    bd = run_modules(vfe, scat, pts, n_frames)
    keep, keys, unq, inv, cnt = po.quantise_keys_numpy(pts.numpy(), cfg)
    nxy, ny = int(grid[0]) * int(grid[1]), int(grid[1])
    coords = np.stack([unq // nxy, np.zeros_like(unq), unq % ny, (unq % nxy) // ny], axis=1).astype(np.int32)
    assert np.array_equal(bd["voxel_coords"].cpu().numpy(), coords)
    pp = bd[CTX_KEY]["point_pillar"].cpu().numpy()
    assert np.array_equal(pp >= 0, keep) and np.array_equal(pp[keep], inv.astype(np.int32))
    sf, pf = bd["spatial_features"], bd["pillar_features"]
    assert int((sf != 0).any(dim=1).sum()) == unq.shape[0]
    assert torch.allclose(sf.double().sum(dim=(0, 2, 3)), pf.double().sum(dim=0), rtol=1e-12, atol=0)
    assert bool(torch.isfinite(pf).all()) and float(pf.min()) >= 0.0
    bd2 = run_modules(vfe, scat, pts, n_frames)
    assert torch.equal(bd2["pillar_features"], pf)


@pytest.mark.parametrize("depth,ego", [(1, False), (2, False), (3, False), (2, True)])
def test_pipelined_front_end_is_bit_identical_to_the_serial_one(depth, ego):
    """Steady-state mode (canvas of batch i under the voxelize kernels of batch i + 1, separate streams and buffer sets):
    every batch of a stream of different-sized batches equals the serial chain bit for bit."""
    from pcp_b200.frontend import FrontEnd, GridSpec, PipelinedFrontEnd
    c_raw = 11 if ego else 5                      # ego: the 14-column lately-fusion rows (PFN row layout 2)
    syn, rng, vox, grid, sd, cfg = v2x_setup(c_raw)
    gs = GridSpec(vox, rng, grid)
    bn = lambda i: [sd[f"pfn_layers.{i}.norm.{k}"].to(DEV) for k in ("weight", "bias", "running_mean", "running_var")]
    w = lambda i: sd[f"pfn_layers.{i}.linear.weight"].to(DEV)
    B = 3
    batches = [syn.batch_of_frames(B, n, 40 + j, ego_columns=ego).to(DEV)
               for j, n in enumerate([60000, 1000, 90000, 250, 40000, 70000, 5])]
    serial = FrontEnd(gs, c_raw)
    serial.pack_params(w(0), bn(0), w(1), bn(1))
    want = []
    for pts in batches:
        o = serial.forward_device(pts, B, {}, None)
        torch.cuda.synchronize()
        p = int(serial.read_counts(o)[0])
        want.append((p, o["voxel_coords_buf"][:p].clone(), o["pillar_features_buf"][:p].clone(), o["spatial_features"].clone()))
    pipe = PipelinedFrontEnd(gs, c_raw, B, depth=depth)
    pipe.pack_params(w(0), bn(0), w(1), bn(1))
    for rep in range(2):
        got = []
        for pts in batches:
            o = pipe.submit(pts)
            torch.cuda.current_stream().wait_event(o["done"])
            # the consumer copies the results out on its own stream, then releases the buffer set
            got.append((o["counts"].clone(), o["voxel_coords_buf"].clone(), o["pillar_features_buf"].clone(),
                        o["spatial_features"].clone()))
            pipe.release(o)
        pipe.drain()
        torch.cuda.synchronize()
        for (p, vc, pf, sf), (cnt, gvc, gpf, gsf) in zip(want, got):
            assert int(cnt[0]) == p
            assert torch.equal(gvc[:p], vc) and torch.equal(gpf[:p], pf) and torch.equal(gsf, sf)


def test_graph_captured_steady_state_is_bit_identical_to_the_serial_chain():
    """PipelinedFrontEnd.capture(): one CUDA graph per buffer set (voxelize + PFN of set k beside the canvas of set k - 1);
    a stream of batches copied into the static input buffers gives the serial chain's results bit for bit."""
    from pcp_b200.frontend import FrontEnd, GridSpec, PipelinedFrontEnd
    syn, rng, vox, grid, sd, cfg = v2x_setup(5)
    gs = GridSpec(vox, rng, grid)
    bn = lambda i: [sd[f"pfn_layers.{i}.norm.{k}"].to(DEV) for k in ("weight", "bias", "running_mean", "running_var")]
    w = lambda i: sd[f"pfn_layers.{i}.linear.weight"].to(DEV)
    B, N = 2, 50000
    batches = [syn.batch_of_frames(B, N, 60 + j).to(DEV) for j in range(5)]
    # a shorter batch: the tail of the static buffer is filled with out-of-range rows, which the cull drops
    short = syn.batch_of_frames(B, N // 2, 70).to(DEV)
    pad = short.new_full((N * B - short.shape[0], short.shape[1]), 1.0e6)
    pad[:, 0] = 0
    batches.append(torch.cat([short, pad]))
    serial = FrontEnd(gs, 5)
    serial.pack_params(w(0), bn(0), w(1), bn(1))
    want = []
    for pts in batches[:5] + [short]:
        o = serial.forward_device(pts, B, {}, None)
        torch.cuda.synchronize()
        p = int(serial.read_counts(o)[0])
        want.append((p, o["voxel_coords_buf"][:p].clone(), o["pillar_features_buf"][:p].clone(), o["spatial_features"].clone()))
    pipe = PipelinedFrontEnd(gs, 5, B, depth=2)
    pipe.pack_params(w(0), bn(0), w(1), bn(1))
    static = [torch.empty_like(batches[0]) for _ in range(2)]
    static[0].copy_(batches[0])
    static[1].copy_(batches[1])
    pipe.capture(static)
    got = []
    snap = lambda o: (o["counts"].clone(), o["voxel_coords_buf"].clone(), o["pillar_features_buf"].clone(), o["spatial_features"].clone())
    for i, pts in enumerate(batches):
        k = i % 2
        static[k].copy_(pts)
        pipe.replay(k)
        if i > 0:
            got.append(snap(pipe.sets[1 - k]))                # batch i - 1 is complete once graph i has run
    got.append(snap(pipe.flush((len(batches) - 1) % 2)))
    torch.cuda.synchronize()
    for (p, vc, pf, sf), (cnt, gvc, gpf, gsf) in zip(want, got):
        assert int(cnt[0]) == p
        assert torch.equal(gvc[:p], vc) and torch.equal(gpf[:p], pf) and torch.equal(gsf, sf)


@pytest.mark.parametrize("alt", [0, 1])
def test_bench_workload_against_oracle(alt):
    """The workload bench.py times (8 early-fusion frames of 300 k points, both alternating batches), through the
    modules, against the oracle with the sort-based unique: pillar coordinates, point->pillar map, occupancy bit-exact,
    pillar features within the path's fp32 tolerance, canvas an exact copy."""
    syn, rng, vox, grid, sd, cfg = v2x_setup(5)
    pts = syn.batch_of_frames(8, 300000, 3, first_frame=alt * 1000)
    vfe, scat = build_modules(5, vox, rng, grid, sd)
    bd = run_modules(vfe, scat, pts, 8)
    want = oracle_want(pts, cfg, layers_from_state_dict(sd))
    check_against(bd, want, pts.shape[0])
    # frame by frame: the pillars of frame b are exactly the oracle's pillars of that frame run alone
    vc = bd["voxel_coords"].cpu()
    for b in (0, 7):
        one = pts[pts[:, 0] == b].clone()
        one[:, 0] = 0
        w1 = po.front_end(one, cfg, layers_from_state_dict(sd), unique_dim0=False)
        sel = vc[:, 0] == b
        assert torch.equal(vc[sel][:, 1:], w1["voxel_coords"][:, 1:])
        assert_features_close(bd["pillar_features"].cpu()[sel].numpy(), w1["pillar_features"].numpy(), f"frame {b}")


@pytest.mark.parametrize("n_points,uniform", [(1000000, False), (1000000, True)])
def test_stress_config_against_oracle(n_points, uniform):
    """BASELINE config 5 at 1 M points (0.1 m pillars, 1024 x 1024 canvas): full comparison with the oracle."""
    voxel = [0.1, 0.1, 8.0]
    syn, rng, vox, grid, sd, cfg = v2x_setup(5, voxel=voxel)
    pts = syn.batch_of_frames(1, n_points, 5, uniform_xy=uniform)
    vfe, scat = build_modules(5, vox, rng, grid, sd)
    bd = run_modules(vfe, scat, pts, 1)
    check_against(bd, oracle_want(pts, cfg, layers_from_state_dict(sd)), pts.shape[0])


@pytest.mark.parametrize("n_frames,n_points,voxel,ego", [(8, 300000, None, False), (1, 32768, None, True), (1, 4000000, [0.1, 0.1, 8.0], False),
                                                         (3, 777, None, False), (16, 1000, None, False)])
def test_compaction_methods_agree_bit_for_bit(n_frames, n_points, voxel, ego):
    """pcp_voxelize_method: the stable radix sort and the dense-histogram path give the same pillars, point->pillar map,
    counts, per-pillar means (sequential sums in row order on both), PFN features, canvas and segment reductions."""
    from pcp_b200.frontend import FrontEnd, GridSpec
    c_raw = 11 if ego else 5
    syn, rng, vox, grid, sd, cfg = v2x_setup(c_raw, voxel=voxel)
    pts = syn.batch_of_frames(n_frames, n_points, 17, ego_columns=ego).to(DEV)
    vals = torch.randn(pts.shape[0], 8, generator=torch.Generator().manual_seed(5)).to(DEV)
    bn = lambda i: [sd[f"pfn_layers.{i}.norm.{k}"].to(DEV) for k in ("weight", "bias", "running_mean", "running_var")]
    res = {}
    for method in ("radix", "histogram", "binned"):
        fe = FrontEnd(GridSpec(vox, rng, grid), c_raw, voxelize_method=method)
        fe.pack_params(sd["pfn_layers.0.linear.weight"].to(DEV), bn(0), sd["pfn_layers.1.linear.weight"].to(DEV), bn(1))
        out = fe.voxelize(pts, n_frames, want_point_pillar=True, want_counts_per_pillar=True)
        fe.pfn(pts, out, want_mean=True)
        canvas = fe.scatter_ws(out["pillar_features_buf"], n_frames)
        smax, smean = fe.segment_reduce(vals, "max"), fe.segment_reduce(vals, "mean")
        torch.cuda.synchronize()
        counts = fe.read_counts(out)
        p = int(counts[0])
        res[method] = dict(counts=counts.copy(), vc=out["voxel_coords_buf"][:p].clone(), pp=out["point_pillar"][:pts.shape[0]].clone(),
                           pc=out["pillar_count_buf"][:p].clone(), pf=out["pillar_features_buf"][:p].clone(),
                           mean=out["pillar_mean_buf"][:p].clone(), canvas=canvas, smax=smax[:p].clone(), smean=smean[:p].clone())
    b = res["histogram"]
    for other in ("radix", "binned"):
        a = res[other]
        assert np.array_equal(a["counts"], b["counts"]), (other, a["counts"], b["counts"])
        assert a["counts"][0] > 0
        for k in ("vc", "pp", "pc", "mean", "pf", "canvas", "smax", "smean"):      # pillars of ANY length: row-order sums on every method
            assert torch.equal(a[k], b[k]), (other, k)


def test_radix_method_reports_what_it_does_not_cover():
    from pcp_b200.frontend import FrontEnd, GridSpec
    syn, rng, vox, grid, sd, cfg = v2x_setup(5)
    pts = syn.batch_of_frames(1, 1000, 3).to(DEV)
    fe = FrontEnd(GridSpec(vox, rng, grid), 5, voxelize_method="radix")
    with pytest.raises(RuntimeError, match="radix method covers"):
        fe.voxelize(pts, 40)                           # 40 x 512 x 512 cells > 4 M: histogram path only
    auto = FrontEnd(GridSpec(vox, rng, grid), 5, voxelize_method="auto")
    out = auto.voxelize(pts, 40)
    torch.cuda.synchronize()
    assert int(auto.read_counts(out)[0]) > 0
    with pytest.raises(ValueError):
        FrontEnd(GridSpec(vox, rng, grid), 5, voxelize_method="sort")
