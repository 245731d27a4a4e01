"""Runs a few device-only steps of the chain (voxelize -> PFN -> canvas) for ncu / timing experiments.
    python tools/profile_step.py [--steps 3] [--frames 8] [--points 300000] [--voxel 0.2]
Prints per-stage CUDA-event times (not valid under ncu)."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcp_b200  # noqa: E402
from pcp_b200 import _lib  # noqa: E402
if os.environ.get("PCP_LIB"):                     # tuning aid: time an alternative build of the library
    _lib.LIB_PATH = os.path.abspath(os.environ["PCP_LIB"])
from pcp_b200 import synthetic as syn  # noqa: E402
from pcp_b200.frontend import FrontEnd, GridSpec  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--frames", type=int, default=8)
ap.add_argument("--points", type=int, default=300000)
ap.add_argument("--voxel", type=float, default=0.2)
ap.add_argument("--ego", action="store_true")
ap.add_argument("--uniform", action="store_true")
ap.add_argument("--presorted", action="store_true", help="rows pre-sorted by pillar: upper bound of what sequential row reads buy")
a = ap.parse_args()

dev = torch.device("cuda", 0)
rng = np.asarray(syn.V2X_RANGE, dtype=np.float32)
vox = [a.voxel, a.voxel, 8.0]
grid = syn.grid_size_of(rng, vox)
c_raw = 11 if a.ego else 5
gs = GridSpec(vox, rng, grid)
sd = syn.pfn_state_dict(c_raw + 6)
fe = FrontEnd(gs, c_raw)
bn = lambda i: [sd[f"pfn_layers.{i}.norm.{k}"].to(dev) for k in ("weight", "bias", "running_mean", "running_var")]
fe.pack_params(sd["pfn_layers.0.linear.weight"].to(dev), bn(0), sd["pfn_layers.1.linear.weight"].to(dev), bn(1))
pts = syn.batch_of_frames(a.frames, a.points, 3, ego_columns=a.ego, uniform_xy=a.uniform).to(dev)
if a.presorted:
    cx = torch.floor((pts[:, 1] - float(rng[0])) / a.voxel).long().clamp(-1, gs.nx)
    cy = torch.floor((pts[:, 2] - float(rng[1])) / a.voxel).long().clamp(-1, gs.ny)
    key = (pts[:, 0].long() * (gs.nx + 2) + (cx + 1)) * (gs.ny + 2) + (cy + 1)
    pts = pts[torch.sort(key, stable=True)[1]].contiguous()
out, canvas = {}, torch.empty((a.frames, 64, gs.ny, gs.nx), dtype=torch.float32, device=dev)
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(a.steps)]
for s in range(a.steps):
    ev[s][0].record()
    fe.voxelize(pts, a.frames, out, want_point_pillar=False)
    ev[s][1].record()
    fe.pfn(pts, out)
    ev[s][2].record()
    fe.scatter_ws(out["pillar_features_buf"], a.frames, canvas)
    ev[s][3].record()
torch.cuda.synchronize()
c = fe.read_counts(out)
for s in range(a.steps):
    t = [ev[s][j].elapsed_time(ev[s][j + 1]) * 1e3 for j in range(3)]
    print(f"step {s}: voxelize {t[0]:.1f} us  pfn {t[1]:.1f} us  canvas {t[2]:.1f} us  total {sum(t):.1f} us")
print(f"N={pts.shape[0]} kept={c[1]} P={c[0]} max/pillar={c[4]}")
