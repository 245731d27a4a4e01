"""Device time of the stand-alone segment reductions (pcp_segment_reduce: scatter_mean / scatter_max over the pillars) on the
bench batch and the stress clouds; PCP_LIB=... times an alternative build.
    python tools/segment_time.py"""
import os, sys, json, statistics
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from pcp_b200 import _lib
if os.environ.get("PCP_LIB"): _lib.LIB_PATH = os.path.abspath(os.environ["PCP_LIB"])
from pcp_b200 import synthetic as syn
from pcp_b200.frontend import FrontEnd, GridSpec
dev = "cuda:0"
rng = np.asarray(syn.V2X_RANGE, dtype=np.float32)
res = {}
for name, frames, npts, vox in (("bench", 8, 300000, syn.V2X_VOXEL), ("stress4M", 1, 4000000, [0.1, 0.1, 8.0]), ("stress1M", 1, 1000000, [0.1, 0.1, 8.0])):
    gs = GridSpec(vox, rng, syn.grid_size_of(rng, vox))
    fe = FrontEnd(gs, 5)
    pts = syn.batch_of_frames(frames, npts, 1).to(dev)
    out = fe.voxelize(pts, frames, {}, want_point_pillar=False)
    vals = torch.randn(pts.shape[0], 64, device=dev)
    for mode, v in (("mean3", pts[:, 1:4]), ("max64", vals), ("mean64", vals)):
        m = "mean" if mode.startswith("mean") else "max"
        for _ in range(3): fe.segment_reduce(v, m)
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fe.segment_reduce(v, m); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        res[f"{name}_{mode}"] = round(statistics.median(ts), 1)
print(os.path.basename(_lib.LIB_PATH), json.dumps(res))
