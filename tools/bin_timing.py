"""Debug: per-tile phase times of bin_finish_kernel2 (the binned voxelize path).  Needs `make -C .../csrc dbg`.
    python tools/bin_timing.py
Stamps (globaltimer, ns): 0 CTA start, 1 ticket + bin range known, 2 rows per cell counted, 3 record published,
4 records placed, 5 prefix of the earlier tiles known, 6 rank-ordered outputs written, 7 short pillars finished."""
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
gs = GridSpec(vox, rng, syn.grid_size_of(rng, vox))
fe = FrontEnd(gs, 5, voxelize_method="binned")
pts = syn.batch_of_frames(8, 300000, 3).to(dev)
out = {}
for _ in range(3):
    fe.voxelize(pts, 8, out, want_point_pillar=False)
torch.cuda.synchronize()
tiles = 8 * 512 * 512 // 2048
buf = np.zeros((tiles, 12), dtype=np.uint64)
lib = _lib.load()
lib.pcp_debug_read_bin_timing.argtypes = [C.c_void_p, C.c_int]
print("rc", lib.pcp_debug_read_bin_timing(buf.ctypes.data, tiles))
t = buf[:, :8].astype(np.int64)
t0 = t[:, 0].min()
rows = buf[:, 10].astype(np.int64)
print(f"kernel span {(t[:, 7].max() - t0) / 1e3:.1f} us; tiles {tiles}; rows/tile mean {rows.mean():.0f} max {rows.max()}")
names = ["ticket+range", "count", "scan+publish", "place", "prefix wait", "rank outputs", "short pillars"]
d = np.diff(t, axis=1) / 1e3
print("phase             mean us   p50    p90    max   | heavy tiles (rows > 8000) mean")
heavy = rows > 8000
for i, nm in enumerate(names):
    print(f"{nm:16s} {d[:, i].mean():7.2f} {np.median(d[:, i]):6.2f} {np.percentile(d[:, i], 90):6.2f} {d[:, i].max():6.2f}   | {d[heavy, i].mean():6.2f}")
life = (t[:, 7] - t[:, 0]) / 1e3
print(f"tile lifetime mean {life.mean():.2f} p50 {np.median(life):.2f} max {life.max():.2f} us; heavy mean {life[heavy].mean():.2f}")
order = np.argsort(t[:, 0])
print("start times (us) of every 64th tile in start order:", [(int(i), round((t[i, 0] - t0) / 1e3, 1), round((t[i, 7] - t0) / 1e3, 1), int(rows[i])) for i in order[::64]])
late = np.argsort(-t[:, 7])[:10]
print("last finishers:", [(int(i), round((t[i, 0] - t0) / 1e3, 1), round((t[i, 7] - t0) / 1e3, 1), int(rows[i])) for i in late])
