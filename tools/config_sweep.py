"""Device-timed sweep over every BASELINE.json config and the section 8(f) rows.
    python tools/config_sweep.py [--iters 20] > profiles/rNN_config_sweep.json

One JSON object per line.  Inputs are resident in HBM; every number is the median of `iters` CUDA-event timings after 5
warm-up runs, on the stream the kernels are launched on.  `roofline_frac` = algorithmic bytes (every point row
read once + pillar_features + voxel_coords + canvas written once) / time / the measured copy bandwidth.
Not the bench contract (bench.py is): these are the parity-test shapes, timed for DESIGN.md.
"""
import argparse
import json
import os
import statistics
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcp_b200  # noqa: E402
from pcp_b200 import synthetic as syn  # noqa: E402
from pcp_b200.frontend import FrontEnd, GridSpec  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=20)
a = ap.parse_args()
dev = torch.device("cuda", 0)
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6550.0


def timed(fn, iters=a.iters, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)


def emit(d):
    print(json.dumps(d), flush=True)


def chain(name, frames, n_points, voxel, ego=False, uniform=False, config_id=1, pts=None):
    rng = np.asarray(syn.V2X_RANGE, dtype=np.float32)
    grid = syn.grid_size_of(rng, voxel)
    gs = GridSpec(voxel, rng, grid)
    c_raw = 11 if ego else 5
    sd = syn.pfn_state_dict(c_raw + 6)
    fe = FrontEnd(gs, c_raw)
    bn = lambda i: [sd[f"pfn_layers.{i}.norm.{k}"].to(dev) for k in ("weight", "bias", "running_mean", "running_var")]
    fe.pack_params(sd["pfn_layers.0.linear.weight"].to(dev), bn(0), sd["pfn_layers.1.linear.weight"].to(dev), bn(1))
    if pts is None:
        pts = syn.batch_of_frames(frames, n_points, config_id, ego_columns=ego, uniform_xy=uniform)
    pts = pts.to(dev)
    out, canvas = {}, torch.empty((frames, 64, gs.ny, gs.nx), dtype=torch.float32, device=dev)
    t_vox = timed(lambda: fe.voxelize(pts, frames, out, want_point_pillar=False))
    t_pfn = timed(lambda: fe.pfn(pts, out))
    t_can = timed(lambda: fe.scatter_ws(out["pillar_features_buf"], frames, canvas))
    t_all = timed(lambda: fe.forward_device(pts, frames, out, canvas))
    # config 5 asks for the segment reductions on their own: scatter-mean of xyz and scatter-max of 64 channels
    vals = torch.randn(pts.shape[0], 64, device=dev)
    t_smean = timed(lambda: fe.segment_reduce(pts[:, 1:4], "mean"))
    t_smax = timed(lambda: fe.segment_reduce(vals, "max"))
    c = fe.read_counts(out)
    p, kept = int(c[0]), int(c[1])
    n = pts.shape[0]
    alg = n * 4 * pts.shape[1] + p * 64 * 4 + p * 16 + frames * 64 * gs.ny * gs.nx * 4
    emit({"config": name, "frames": frames, "points": n, "kept": kept, "pillars": p, "grid": [gs.nx, gs.ny], "c_raw": c_raw,
          "us": {"voxelize": t_vox, "pfn": t_pfn, "canvas": t_can, "chain": t_all, "segment_mean_xyz": t_smean,
                 "segment_max_64ch": t_smax},
          "frames_per_s": frames / (t_all * 1e-6), "mpts_per_s": n / t_all, "algorithmic_bytes": alg,
          "roofline_frac": alg / (t_all * 1e-6) / 1e9 / PEAK})
    del canvas, out, fe
    torch.cuda.empty_cache()


# ---- BASELINE configs ----
chain("1: v2x_pointpillar_basic_car, one 32k-pt frame", 1, 32768, syn.V2X_VOXEL, config_id=1)
# config 2: lately fusion = MoDAR exchange (5 agents, 0.2 s latency) + the ego-layout chain
ego14, agents = syn.modar_scene(2, 0, n_agents=5, n_ego_points=32768)
ego_d = ego14.to(dev)
ag_d = [{k: (v.to(dev) if hasattr(v, "to") else v) for k, v in ag.items()} for ag in agents]
exch = lambda: pcp_b200.modar_exchange([g["modar"] for g in ag_d], [g["foreground"] for g in ag_d],
                                       [g["target_se3_agent"] for g in ag_d], 0.0, 0.2, ego_d, max_sweep_idx=10.0)
t_modar = timed(exch)
fused = exch()
emit({"config": "2: MoDAR exchange, 5 agents", "ego_points": int(ego_d.shape[0]), "modar_rows": int(fused.shape[0] - ego_d.shape[0]),
      "foreground_points": int(sum(g["foreground"].shape[0] for g in ag_d)), "us": {"modar_exchange (incl. small H2D of offsets/poses)": t_modar}})
chain("2: v2x_pointpillar_basic_ego, ego sweeps + MoDAR rows", 1, 0, syn.V2X_VOXEL, ego=True, pts=fused.cpu())
chain("3: v2x_pointpillar_basic_ego_early, one 300k-pt frame", 1, 300000, syn.V2X_VOXEL, config_id=3)
chain("4: 8 early-fusion frames per GPU (the bench workload)", 8, 300000, syn.V2X_VOXEL, config_id=3)
for n in (1_000_000, 2_000_000, 4_000_000):
    chain(f"5: stress {n // 1_000_000}M pts, 0.1 m pillars, 1024x1024, radial", 1, n, syn.STRESS_VOXEL, config_id=5)
chain("5: stress 4M pts, 0.1 m pillars, 1024x1024, uniform xy", 1, 4_000_000, syn.STRESS_VOXEL, uniform=True, config_id=5)

# ---- SURVEY 8(f) rows ----
g = torch.Generator().manual_seed(0)
B, C, H, W = 4, 64, 256, 256
img = torch.randn(B, C, H, W, generator=g).to(dev)
pts = syn.batch_of_frames(B, 75000, 6).to(dev)                       # 300k points over 4 frames
rng = np.asarray(syn.V2X_RANGE, dtype=np.float32)
pix = np.asarray([0.4, 0.4], dtype=np.float32)
t_int = timed(lambda: pcp_b200.interpolate_points_feat_from_bev_img(img, pts, rng, pix, True))
feat, coord = pcp_b200.interpolate_points_feat_from_bev_img(img, pts, rng, pix, True)
bidx = pts[:, 0].long()
t_sc = timed(lambda: pcp_b200.bev_scatter(coord, bidx, feat, (H, W), batch_size=B))
n = pts.shape[0]
emit({"config": "8f-1: hunter_toolbox on a (4, 64, 256, 256) image, 300k points", "us": {"interpolate_points_feat_from_bev_img": t_int, "bev_scatter": t_sc},
      "interpolate_GBps": (img.numel() * 8 + n * C * 4 * 5) / (t_int * 1e-6) / 1e9,
      "bev_scatter_GBps": (n * C * 4 + B * C * H * W * 4) / (t_sc * 1e-6) / 1e9,
      "note": "includes the workspace / output allocations of the drop-in functions (torch caching allocator)"})

vox3 = [0.1, 0.1, 0.2]
grid3 = syn.grid_size_of(rng, vox3)
mvfe = pcp_b200.DynamicMeanVFE(model_cfg=pcp_b200.CfgDict(), num_point_features=5, voxel_size=vox3, grid_size=grid3, point_cloud_range=rng)
p3 = syn.batch_of_frames(2, 150000, 8).to(dev)
t_mv = timed(lambda: mvfe({"points": p3, "batch_size": 2}))
bd = mvfe({"points": p3, "batch_size": 2})
emit({"config": "8f-2: DynamicMeanVFE, SECOND grid 1024 x 1024 x 40, 2 x 150k points", "voxels": int(bd["voxel_coords"].shape[0]),
      "us": {"forward (incl. the 32-byte counts read-back)": t_mv}, "mpts_per_s": p3.shape[0] / t_mv})

clouds = [syn.lidar_frame(50000, 9000 + k)[:, 1:].contiguous().to(dev) for k in range(6)]
tfs = [syn.modar_agent(9100 + k, n_boxes=1)["target_se3_agent"] for k in range(5)]
t_fu = timed(lambda: pcp_b200.fuse_agent_points(clouds[0], clouds[1:], tfs, rng, batch_idx=0))
emit({"config": "8f-3: early-fusion assembly, 6 clouds x 50k points", "us": {"fuse_agent_points (per-agent pointers, one small H2D, count read-back)": t_fu},
      "mpts_per_s": 300000 / t_fu})
