"""Debug: wall-clock intervals (globaltimer) of the kernels of ONE pipelined step (graph replay: voxelize + PFN of batch i + 1
beside the canvas of batch i) and of one serial step.  Needs `make -C .../csrc dbg` (libpcp_b200_st.so).
    python tools/stage_trace.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcp_b200 import _lib  # noqa: E402
_lib.LIB_PATH = _lib.LIB_PATH.replace("libpcp_b200.so", "libpcp_b200_st.so")
from pcp_b200 import synthetic as syn  # noqa: E402
from pcp_b200.frontend import FrontEnd, GridSpec, PipelinedFrontEnd  # noqa: E402

dev = torch.device("cuda", 0)
rng = np.asarray(syn.V2X_RANGE, dtype=np.float32)
gs = GridSpec(syn.V2X_VOXEL, rng, syn.grid_size_of(rng, syn.V2X_VOXEL))
sd = syn.pfn_state_dict(11)
B = 8
batches = [syn.batch_of_frames(B, 300000, 3, first_frame=alt * 1000).to(dev) for alt in range(2)]
bn = lambda i: [sd[f"pfn_layers.{i}.norm.{k}"].to(dev) for k in ("weight", "bias", "running_mean", "running_var")]
lib = _lib.load()
readers = [("voxelize", lib.pcp_debug_stage_voxelize, ["quantise_count", "cell scan", "place", "pillar_prep"]),
           ("pfn", lib.pcp_debug_stage_pfn, ["pfn_slot"]), ("canvas", lib.pcp_debug_stage_canvas, ["canvas"])]
for _, fn, _n in readers:
    fn.argtypes = [C.c_void_p, C.c_int]


def reset():
    torch.cuda.synchronize()
    for _, fn, _n in readers:
        assert fn(None, 1) == 0


def read(title):
    rows = []
    for _, fn, names in readers:
        buf = np.zeros((8, 2), dtype=np.uint64)
        assert fn(buf.ctypes.data, 0) == 0
        for i, nm in enumerate(names):
            if buf[i, 1] > 0:
                rows.append((nm, int(buf[i, 0]), int(buf[i, 1])))
    t0 = min(r[1] for r in rows)
    print(f"== {title}")
    for nm, a, b in sorted(rows, key=lambda r: r[1]):
        print(f"   {nm:16s} start {(a - t0) / 1e3:8.1f} us   end {(b - t0) / 1e3:8.1f} us   span {(b - a) / 1e3:7.1f} us")
    print(f"   step span {(max(r[2] for r in rows) - t0) / 1e3:.1f} us")


pipe = PipelinedFrontEnd(gs, 5, B, depth=2)
pipe.pack_params(sd["pfn_layers.0.linear.weight"].to(dev), bn(0), sd["pfn_layers.1.linear.weight"].to(dev), bn(1))
pipe.capture(batches)
for i in range(8):
    pipe.replay(i & 1)
for trial in range(3):
    reset()
    pipe.replay(trial & 1)
    read(f"pipelined step (graph replay {trial})")
    pipe.replay((trial + 1) & 1)           # keep the alternation
fe = FrontEnd(gs, 5)
fe.packed = pipe.stages[0].packed
out, canvas = {}, torch.empty((B, 64, gs.ny, gs.nx), dtype=torch.float32, device=dev)
for i in range(3):
    fe.forward_device(batches[i & 1], B, out, canvas)
reset()
fe.forward_device(batches[1], B, out, canvas)
read("serial step (one batch, one stream, issued eagerly)")
