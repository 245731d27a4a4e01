"""Debug: per-role event trace (clock64) of CTA 0 of the PFN kernel.  Needs `make -C .../csrc dbg` (libpcp_b200_dbg.so).
    python tools/pfn_timing.py [first_event] [n_events]
Event ids - MMA warp: 20/21 before/after wait A0 (layer-0 operand), 22/23 before/after wait A1, 24 group done;
producer: 13/10 slot start (13 = first slot of a group), 14 rows landed, 11/12 before/after wait for the A0 buffer;
epilogue: 30/31 before/after wait D0, 32 layer-0 epilogue done, 33/34 before/after wait D1, 37 hoist added, 38 rows written."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcp_b200 import _lib  # noqa: E402
_lib.LIB_PATH = _lib.LIB_PATH.replace("libpcp_b200.so", "libpcp_b200_dbg.so")
from pcp_b200 import synthetic as syn  # noqa: E402
from pcp_b200.frontend import FrontEnd, GridSpec  # noqa: E402

dev = torch.device("cuda", 0)
rng = np.asarray(syn.V2X_RANGE, dtype=np.float32)
vox = syn.V2X_VOXEL
grid = syn.grid_size_of(rng, vox)
gs = GridSpec(vox, rng, grid)
sd = syn.pfn_state_dict(11)
fe = FrontEnd(gs, 5)
bn = lambda i: [sd[f"pfn_layers.{i}.norm.{k}"].to(dev) for k in ("weight", "bias", "running_mean", "running_var")]
fe.pack_params(sd["pfn_layers.0.linear.weight"].to(dev), bn(0), sd["pfn_layers.1.linear.weight"].to(dev), bn(1))
pts = syn.batch_of_frames(8, 300000, 3).to(dev)
out = {}
for _ in range(3):
    fe.voxelize(pts, 8, out, want_point_pillar=False)
    fe.pfn(pts, out)
torch.cuda.synchronize()
CAP = 8192
buf = np.zeros((5, CAP, 2), dtype=np.int64)
cnt = np.zeros(5, dtype=np.int32)
lib = _lib.load()
lib.pcp_debug_read_timing.argtypes = [C.c_void_p, C.c_void_p]
print("rc", lib.pcp_debug_read_timing(buf.ctypes.data, cnt.ctypes.data), "events", cnt)
first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n_ev = int(sys.argv[2]) if len(sys.argv) > 2 else 120
t0 = min(int(buf[r, 0, 1]) for r in range(5) if cnt[r] > 0)
names = ["mma", "producer set 0", "E0", "E1", "producer set 1"]
for r in range(5):
    ev = buf[r, :min(int(cnt[r]), CAP)]
    if len(ev) == 0:
        continue
    print(f"== {names[r]}: {len(ev)} events, span {int(ev[-1, 1] - ev[0, 1])} cycles")
    # time spent between consecutive events, summed by (from id -> to id)
    agg = {}
    for i in range(1, len(ev)):
        k = (int(ev[i - 1, 0]), int(ev[i, 0]))
        a = agg.setdefault(k, [0, 0])
        a[0] += int(ev[i, 1] - ev[i - 1, 1]); a[1] += 1
    for k, (tot, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"   {k[0]:3d} -> {k[1]:3d}: total {tot:9d} cycles over {n:5d} = {tot / n:8.1f} each")
    line = []
    for i in range(first, min(first + n_ev, len(ev))):
        line.append(f"[{int(ev[i, 0])}]@{int(ev[i, 1]) - t0}")
    print(" ".join(line))
