"""Tuning aid: steady-state step time of PipelinedFrontEnd (two CUDA graphs, as bench.py times it) for one build of the
library (PCP_LIB=...) and one pair of stream priorities.
    PCP_LIB=practical-collab-perception_b200/libpcp_b200_x.so python tools/pipe_matrix.py --prio -5,0"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcp_b200 import _lib  # noqa: E402
if os.environ.get("PCP_LIB"):
    _lib.LIB_PATH = os.path.abspath(os.environ["PCP_LIB"])
from pcp_b200 import synthetic as syn  # noqa: E402
from pcp_b200.frontend import GridSpec, PipelinedFrontEnd  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--prio", default="-5,0")
ap.add_argument("--steps", type=int, default=60)
a = ap.parse_args()
prio = tuple(int(v) for v in a.prio.split(","))
dev = torch.device("cuda", 0)
rng = np.asarray(syn.V2X_RANGE, dtype=np.float32)
gs = GridSpec(syn.V2X_VOXEL, rng, syn.grid_size_of(rng, syn.V2X_VOXEL))
sd = syn.pfn_state_dict(11)
B = 8
batches = [syn.batch_of_frames(B, 300000, 3, first_frame=alt * 1000).to(dev) for alt in range(2)]
bn = lambda i: [sd[f"pfn_layers.{i}.norm.{k}"].to(dev) for k in ("weight", "bias", "running_mean", "running_var")]
pipe = PipelinedFrontEnd(gs, 5, B, depth=2, priorities=prio)
pipe.pack_params(sd["pfn_layers.0.linear.weight"].to(dev), bn(0), sd["pfn_layers.1.linear.weight"].to(dev), bn(1))
ev = lambda: torch.cuda.Event(enable_timing=True)
st = [[ev() for _ in range(6)] for _ in range(12)]
for i in range(4):
    pipe.submit(batches[i & 1])
pipe.drain()
for i in range(12):
    pipe.submit(batches[i & 1], st[i])
pipe.drain()
torch.cuda.synchronize()
stage = [float(np.mean([st[i][2 * j].elapsed_time(st[i][2 * j + 1]) for i in range(2, 12)])) * 1e3 for j in range(3)]
pipe.capture(batches)
for i in range(6):
    pipe.replay(i & 1)
torch.cuda.synchronize()
e0, e1 = ev(), ev()
e0.record()
for i in range(a.steps):
    pipe.replay(i & 1)
e1.record()
torch.cuda.synchronize()
print(json.dumps({"lib": os.path.basename(_lib.LIB_PATH), "prio": prio, "us_per_step": e0.elapsed_time(e1) / a.steps * 1e3,
                  "eager_stage_us": {"voxelize": stage[0], "pfn": stage[1], "canvas": stage[2]}}), flush=True)
