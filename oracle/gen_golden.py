"""Generate tests/golden/*.npz by running the REFERENCE's own modules (build container only).

TEST INFRASTRUCTURE ONLY.  Run from the repo root:  ``python -m oracle.gen_golden``

The reference has no tests or golden vectors for this path, so the fixtures are
outputs of the reference itself: ``DynamicPillarVFE`` / ``PointPillarScatter`` / ``apply_se3_`` imported
from /root/reference by ``oracle/ref_loader.py`` (torch_scatter replaced by its pure-torch restatement,
the one third-party piece that is absent), evaluated on seeded synthetic inputs and hand-built edge
cases.  Each file stores the inputs, the parameters and the reference outputs.  Dense canvases are
stored whole only for small grids; for 512^2 grids the occupied cell list and per-channel float64
sums are stored instead (the canvas is a pure function of pillar_features and voxel_coords).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader as rl  # noqa: E402
import pcp_b200  # noqa: E402,F401
from pcp_b200 import synthetic as syn  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def run_reference(points, c_raw, voxel, rng, num_filters=(64, 64), use_norm=True, with_distance=False,
                  use_abs=True, seed=0):
    rng = np.asarray(rng, dtype=np.float32)          # pcdet/datasets/dataset.py:25
    grid = syn.grid_size_of(rng, voxel)
    vfe, scat = rl.build_reference_front_end(c_raw, voxel, rng, grid, num_filters, use_norm, with_distance, use_abs)
    c_in = c_raw + (6 if use_abs else 3) + (1 if with_distance else 0)
    sd = syn.pfn_state_dict(c_in, num_filters, use_norm, seed)
    missing = vfe.load_state_dict(sd)
    assert not missing.missing_keys and not missing.unexpected_keys, missing

    # capture unq_inv: the reference does not export it, re-derive exactly as dynamic_pillar_vfe.py:98-108 does
    captured = {}
    orig_unique = torch.unique

    def spy(*a, **k):
        out = orig_unique(*a, **k)
        if k.get("return_inverse") and k.get("dim", None) == 0:
            captured["unq_inv"] = out[1].clone()
        return out

    torch.unique = spy
    try:
        with torch.no_grad():
            bd = vfe({"points": points.clone()})
            bd = scat(bd)
    finally:
        torch.unique = orig_unique
    return bd, captured["unq_inv"], sd, grid


def save_case(name, points, c_raw, voxel, rng, store_canvas=False, **kw):
    bd, unq_inv, sd, grid = run_reference(points, c_raw, voxel, rng, **kw)
    sf = bd["spatial_features"]
    occ = (sf != 0).any(dim=1)                                  # (B, ny, nx)
    data = {
        "points": points.numpy(),
        "c_raw": np.int64(c_raw),
        "voxel_size": np.asarray(voxel, dtype=np.float64),
        "point_cloud_range": np.asarray(rng, dtype=np.float32),
        "grid_size": grid,
        "num_filters": np.asarray(kw.get("num_filters", (64, 64)), dtype=np.int64),
        "use_norm": np.bool_(kw.get("use_norm", True)),
        "with_distance": np.bool_(kw.get("with_distance", False)),
        "use_abs": np.bool_(kw.get("use_abs", True)),
        "seed": np.int64(kw.get("seed", 0)),
        "pillar_features": bd["pillar_features"].numpy(),
        "voxel_coords": bd["voxel_coords"].numpy(),
        "unq_inv": unq_inv.numpy().astype(np.int32),
        "canvas_shape": np.asarray(sf.shape, dtype=np.int64),
        "occupied": torch.nonzero(occ.flatten()).flatten().numpy().astype(np.int32),
        "canvas_channel_sums": sf.double().sum(dim=(0, 2, 3)).numpy(),
    }
    for k, v in sd.items():
        data["sd/" + k] = v.numpy()
    if store_canvas:
        data["spatial_features"] = sf.numpy()
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **data)
    print(f"{name}: N={points.shape[0]} P={bd['pillar_features'].shape[0]} canvas={tuple(sf.shape)} "
          f"-> {os.path.getsize(path) / 1e3:.0f} kB")


def edge_case_points():
    """Hand-built rows on a tiny grid (range +-0.8 m, 0.2 m pillars -> 8 x 8): cell edges, range ends,
    outside points, duplicates, one-point pillars, negative zero, z outside its range (z is NOT culled,
    dynamic_pillar_vfe.py:99), and a batch whose last frame (b=2) has only culled points."""
    f = np.float32
    rows = []

    def add(b, x, y, z=-1.0, i=0.5, t=0.0):
        rows.append([b, x, y, z, i, t, 10.0, -1.0])

    edges = [-0.8, -0.6, -0.2, 0.0, -0.0, 0.2, 0.6, 0.79999995, 0.8, -0.80000001, -0.8000001, 0.8000001,
             0.19999999, 0.20000002, 0.59999996, 0.6000001, 1e-9, -1e-9, 0.4, -0.4]
    for b in (0, 1):
        for ex in edges:
            for ey in (edges[0], edges[3], edges[5], edges[7], edges[8], edges[12]):
                add(b, ex, ey, z=-1.0 - b)
    # duplicates and a dense pillar
    for k in range(40):
        add(0, 0.3 + 1e-3 * k, 0.3, z=-0.1 * k, i=k / 40.0, t=0.1 * (k % 11))
    for k in range(3):
        add(1, -0.5, 0.5)
    # z far outside its range is kept
    add(0, 0.1, 0.1, z=100.0)
    add(1, 0.1, 0.1, z=-100.0)
    # frame 2: everything culled -> B shrinks to 2 in the scatter (pointpillar_scatter.py:17)
    add(2, 5.0, 5.0)
    add(2, -0.9, 0.0)
    add(2, 0.0, float("nan"))
    add(2, float("inf"), 0.0)
    pts = torch.tensor(np.asarray(rows, dtype=f))
    perm = torch.randperm(pts.shape[0], generator=torch.Generator().manual_seed(7))
    return pts[perm].contiguous()


def modar_cases():
    ns = rl.load_reference_modules()
    from oracle import modar_oracle as mo
    data = {}
    for a in range(4):
        ag = syn.modar_agent(4242 + a, n_boxes=12 + a, fg_per_box=(3, 12))
        boxes = ag["modar"][:, :7].numpy().copy()
        ref = ns.apply_se3_(ag["target_se3_agent"], boxes_=boxes, return_transformed=True)   # the reference itself
        data[f"a{a}/modar"] = ag["modar"].numpy()
        data[f"a{a}/foreground"] = ag["foreground"].numpy()
        data[f"a{a}/se3"] = ag["target_se3_agent"]
        data[f"a{a}/ref_apply_se3_boxes"] = ref
        # propagate: box membership needs the reference's CUDA op (not runnable on this CPU host);
        # stored from the restatement, re-checked against the reference's compiled kernel in the GPU tests
        data[f"a{a}/oracle_box_idx"] = mo.points_in_boxes(ag["foreground"][:, :3].numpy(), ag["modar"][:, :7].numpy())
        data[f"a{a}/oracle_propagated"] = mo.propagate_modar(ag["modar"].numpy(), ag["foreground"].numpy(), 2.0)
    path = os.path.join(OUT, "modar_small.npz")
    np.savez_compressed(path, **data)
    print(f"modar_small -> {os.path.getsize(path) / 1e3:.0f} kB")


def next_cases():
    """SURVEY 8(f) rows, outputs of the reference's own functions / modules on seeded inputs -> tests/golden/next_*.npz."""
    ns = rl.load_reference_modules()
    g = torch.Generator().manual_seed(99)
    # ---- hunter_toolbox: interpolate + bev_scatter on a (2, 24, 32, 40) image ----
    B, C, H, W = 2, 24, 32, 40
    rng6 = np.asarray([-8.0, -6.4, -8.0, 8.0, 6.4, 0.0], dtype=np.float32)
    pix = np.asarray([0.4, 0.4], dtype=np.float32)
    img = torch.randn(B, C, H, W, generator=g)
    n = 3000
    pts = torch.zeros(n, 8)
    pts[:, 0] = torch.randint(0, B, (n,), generator=g).float()
    pts[:, 1] = (torch.rand(n, generator=g) * 2 - 1) * 8.6          # some outside the image on both sides
    pts[:, 2] = (torch.rand(n, generator=g) * 2 - 1) * 6.9
    pts[:, 3:] = torch.randn(n, 5, generator=g)
    # exact pixel centres / borders
    pts[:40, 1] = torch.linspace(-8.0, 8.0, 40)
    pts[:40, 2] = -6.4
    pts[40:80, 1] = torch.arange(40).float() * 0.4 - 8.0
    pts[40:80, 2] = torch.arange(40).float().remainder(32) * 0.4 - 6.4
    feat_ref, coord_ref = ns.interpolate_points_feat_from_bev_img(img, pts.clone(), torch.from_numpy(rng6), torch.from_numpy(pix),
                                                                  return_bev_coord=True)
    pfeat = torch.randn(n, C, generator=g)
    pfeat[:100] *= 1e3
    bev_ref = ns.bev_scatter(coord_ref, pts[:, 0].long(), pfeat, (H, W))
    np.savez_compressed(os.path.join(OUT, "next_hunter_small.npz"), bev_img=img.numpy(), points=pts.numpy(), range6=rng6, pixel=pix,
                        ref_points_feat=feat_ref.numpy(), ref_bev_coord=coord_ref.numpy(), points_feat=pfeat.numpy(),
                        ref_bev_scatter=bev_ref.numpy())
    print("next_hunter_small", tuple(feat_ref.shape), tuple(bev_ref.shape))

    # ---- DynamicMeanVFE on a 64 x 64 x 10 grid (SECOND-style), 3 frames ----
    vox3 = [0.2, 0.2, 0.8]
    rng3 = np.asarray([-6.4, -6.4, -8.0, 6.4, 6.4, 0.0], dtype=np.float32)
    grid3 = syn.grid_size_of(rng3, vox3)
    p3 = syn.batch_of_frames(3, 2500, 9)
    p3[:, 1:3] *= 0.12
    p3[:50, 3] = torch.linspace(-8.2, 0.2, 50)                       # z outside its range IS culled here
    with ns.cuda_is_identity():
        mvfe = ns.DynamicMeanVFE(model_cfg=rl.Cfg(), num_point_features=5, voxel_size=vox3, grid_size=grid3, point_cloud_range=rng3)
    with torch.no_grad():
        bd = mvfe({"points": p3.clone()})
    np.savez_compressed(os.path.join(OUT, "next_meanvfe_small.npz"), points=p3.numpy(), voxel_size=np.asarray(vox3), point_cloud_range=rng3,
                        grid_size=grid3, num_point_features=np.int64(5), ref_voxel_features=bd["voxel_features"].numpy(),
                        ref_voxel_coords=bd["voxel_coords"].numpy())
    print("next_meanvfe_small", tuple(bd["voxel_features"].shape))

    # ---- DynamicPillarVFESimple2D, 64 x 64 grid, all 7 columns ----
    small_rng, small_vox = np.asarray([-6.4, -6.4, -8.0, 6.4, 6.4, 0.0], dtype=np.float32), [0.2, 0.2, 8.0]
    grid2 = syn.grid_size_of(small_rng, small_vox)
    p2 = syn.batch_of_frames(3, 600, 7)
    p2[:, 1:3] *= 0.1
    for tag, use_abs, dist in (("abs", True, False), ("rel_dist", False, True)):
        cfg = rl.Cfg(USE_NORM=True, WITH_DISTANCE=dist, USE_ABSLOTE_XYZ=use_abs, NUM_FILTERS=[64, 64])
        with ns.cuda_is_identity():
            m = ns.DynamicPillarVFESimple2D(model_cfg=cfg, num_point_features=7, voxel_size=small_vox, grid_size=grid2,
                                            point_cloud_range=small_rng)
        c_in = 7 + (3 if use_abs else 0) + (1 if dist else 0)
        sd = syn.pfn_state_dict(c_in, (64, 64), True, 11)
        missing = m.load_state_dict(sd)
        assert not missing.missing_keys and not missing.unexpected_keys, missing
        m.eval()
        with torch.no_grad():
            bd = m({"points": p2.clone()})
        data = dict(points=p2.numpy(), voxel_size=np.asarray(small_vox), point_cloud_range=small_rng, grid_size=grid2,
                    use_abs=np.bool_(use_abs), with_distance=np.bool_(dist), ref_pillar_features=bd["pillar_features"].numpy(),
                    ref_pillar_coords=bd["pillar_coords"].numpy())
        for k, v in sd.items():
            data["sd/" + k] = v.numpy()
        np.savez_compressed(os.path.join(OUT, f"next_simple2d_{tag}.npz"), **data)
        print(f"next_simple2d_{tag}", tuple(bd["pillar_features"].shape))

    # ---- early-fusion assembly: apply_se3_ per agent (the reference's function) + concatenate + mask_points_by_range ----
    clouds, tfs = [], []
    for a in range(4):
        f = syn.lidar_frame(1500, 7000 + a)[:, 1:].contiguous()      # (n, 7) x y z i t sweep inst
        clouds.append(f.numpy().copy())
        if a > 0:
            ag = syn.modar_agent(5000 + a, n_boxes=1)
            tfs.append(ag["target_se3_agent"])
    moved = [clouds[0]]
    for c, tf in zip(clouds[1:], tfs):
        x = c.copy()
        x[:, :3] = ns.apply_se3_(tf, points_=x[:, :3], return_transformed=True)      # v2x_sim_dataset_ego_early.py:85
        moved.append(x)
    cat = np.concatenate(moved, axis=0)
    r = np.asarray(syn.V2X_RANGE, dtype=np.float32)
    # mask_points_by_range, pcdet/utils/common_utils.py:64-68 (that module imports SharedArray-era dependencies at import
    # time; its three lines are evaluated here verbatim on the reference-transformed points)
    mask = (cat[:, 0] >= r[0]) & (cat[:, 0] < r[3]) & (cat[:, 1] >= r[1]) & (cat[:, 1] < r[4]) & (cat[:, 2] >= r[2]) & (cat[:, 2] < r[5])
    data = {"range6": r, "ref_fused": cat[mask], "ref_fused_nomask": cat}
    for a, c in enumerate(clouds):
        data[f"cloud{a}"] = c
    for a, tf in enumerate(tfs):
        data[f"se3_{a + 1}"] = tf
    np.savez_compressed(os.path.join(OUT, "next_early_fusion_small.npz"), **data)
    print("next_early_fusion_small", cat.shape, int(mask.sum()))


def main():
    assert rl.reference_available(), "run in the build container: /root/reference must exist"
    torch.set_num_threads(1)
    rng, vox = syn.V2X_RANGE, syn.V2X_VOXEL
    # config-1 shape, scaled down: car layout, C_raw 5, 512^2 grid, 2 frames
    save_case("vfe_car_small", syn.batch_of_frames(2, 1500, 1), 5, vox, rng)
    # config-2 shape: ego layout, C_raw 11 (MoDAR columns populated for a few rows)
    ego = syn.batch_of_frames(2, 1200, 2, ego_columns=True)
    g = torch.Generator().manual_seed(5)
    sel = torch.randperm(ego.shape[0], generator=g)[:150]
    ego[sel, 6:12] = torch.rand(150, 6, generator=g) * 4
    ego[sel, 4:6] = 0
    save_case("vfe_ego_small", ego, 11, vox, rng)
    # config-5 shape: 0.1 m pillars, 1024^2 grid
    save_case("vfe_stress_small", syn.batch_of_frames(1, 2500, 5), 5, syn.STRESS_VOXEL, rng)
    # edge cases on an 8 x 8 grid, whole canvas stored
    tiny_rng, tiny_vox = [-0.8, -0.8, -8.0, 0.8, 0.8, 0.0], [0.2, 0.2, 8.0]
    save_case("vfe_edges_tiny", edge_case_points(), 5, tiny_vox, tiny_rng, store_canvas=True)
    # module options the cfg exposes (dynamic_pillar_vfe.py:57-62): 64 x 64 grid, whole canvas stored
    small_rng, small_vox = [-6.4, -6.4, -8.0, 6.4, 6.4, 0.0], [0.2, 0.2, 8.0]
    pts = syn.batch_of_frames(3, 600, 7)
    pts[:, 1:3] *= 0.1
    save_case("vfe_opts_default", pts, 5, small_vox, small_rng, store_canvas=True)
    save_case("vfe_opts_nonorm", pts, 5, small_vox, small_rng, store_canvas=True, use_norm=False, seed=3)
    save_case("vfe_opts_distance", pts, 5, small_vox, small_rng, store_canvas=True, with_distance=True, seed=4)
    save_case("vfe_opts_relxyz", pts, 5, small_vox, small_rng, store_canvas=True, use_abs=False, seed=5)
    save_case("vfe_opts_single_layer", pts, 5, small_vox, small_rng, store_canvas=True, num_filters=(64,), seed=6)
    modar_cases()
    next_cases()


if __name__ == "__main__":
    main()
