"""Debug: per-phase clock64 trace of one CTA of the PFN kernel (needs `make -C csrc dbg`: libpcp_b200_dbg.so)."""
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
buf = (C.c_longlong * 8192)()
lib = _lib.load()
lib.pcp_debug_read_timing.argtypes = [C.c_void_p]
print("rc", lib.pcp_debug_read_timing(buf))
a = np.frombuffer(buf, dtype=np.int64).reshape(2, 1024, 4)
first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for t in range(2):
    ev = a[t]
    ev = ev[ev[:, 1] > 0]
    print(f"thread {'0' if t == 0 else '200'}: {len(ev)} events, total {int(ev[-1,1]-ev[0,1])} cycles")
    prev = None
    line = []
    for i in range(len(ev)):
        sid, clk = int(ev[i, 0]), int(ev[i, 1])
        d = 0 if prev is None else clk - prev
        prev = clk
        if sid >= 100 and sid not in (101, 102, 103, 104, 105, 106, 107, 108, 110) or sid == 1:
            if line and i > first and i < first + 400:
                print(" ".join(line))
            line = []
        line.append(f"[{sid}]+{d}")
    print(" ".join(line))
