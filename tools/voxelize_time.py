"""Device time of pcp_voxelize_method alone, radix sort vs dense histogram (CUDA events, medians): the bench batch
(8 early-fusion frames), one early-fusion frame, one 32 k frame, the 1 M / 4 M stress clouds.
    python tools/voxelize_time.py [--json out.json]"""
import argparse
import json
import os
import statistics
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcp_b200 import synthetic as syn                      # noqa: E402
from pcp_b200.frontend import FrontEnd, GridSpec           # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--case", type=int, default=-1, help="run only this case (for an ncu launch list)")
    args = ap.parse_args()
    dev = "cuda:0"
    rng = np.asarray(syn.V2X_RANGE, dtype=np.float32)
    cases = [("bench 8 x 300k", 8, 300000, syn.V2X_VOXEL), ("1 x 300k", 1, 300000, syn.V2X_VOXEL), ("1 x 32k", 1, 32768, syn.V2X_VOXEL),
             ("stress 1M 0.1m", 1, 1000000, [0.1, 0.1, 8.0]), ("stress 4M 0.1m", 1, 4000000, [0.1, 0.1, 8.0])]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    res = {}
    if args.case >= 0:
        cases = cases[args.case:args.case + 1]
    for name, frames, npts, vox in cases:
        gs = GridSpec(vox, rng, syn.grid_size_of(rng, vox))
        batches = [syn.batch_of_frames(frames, npts, 3, first_frame=1000 * a).to(dev) for a in range(2)]
        row = {}
        for method in ("radix", "histogram"):
            fe = FrontEnd(gs, 5, voxelize_method=method)
            out = {}
            for cold in (False, True):
                ts = []
                for i in range(args.iters + 3):
                    if cold:
                        flush.zero_()                     # L2 flushed: every kernel starts from HBM
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    fe.voxelize(batches[i & 1], frames, out, want_point_pillar=False)
                    e1.record()
                    torch.cuda.synchronize()
                    if i >= 3:
                        ts.append(e0.elapsed_time(e1) * 1e3)
                row[method + ("_cold" if cold else "")] = statistics.median(ts)
            row["pillars"] = int(fe.read_counts(out)[0])
        res[name] = row
        print(f"{name:16s} pillars {row['pillars']:8d}  radix {row['radix']:7.1f} us (L2 flushed {row['radix_cold']:7.1f})   "
              f"histogram {row['histogram']:7.1f} us (L2 flushed {row['histogram_cold']:7.1f})", flush=True)
    if args.json:
        json.dump(res, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
