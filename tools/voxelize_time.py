"""Device time of pcp_voxelize_method alone, per compaction method (CUDA events, medians; `graph`: the same call captured
in a CUDA graph and replayed, i.e. without the host's launch gaps): the bench batch
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
    ap.add_argument("--methods", default="binned,histogram,radix")
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
        for method in args.methods.split(","):
            fe = FrontEnd(gs, 5, voxelize_method=method)
            out = {}
            if not getattr(fe, method + "_applies", lambda *a: True)(npts * frames, frames):
                continue
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
            # the same call as one CUDA graph launch
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fe.voxelize(batches[0], frames, out, want_point_pillar=False)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    fe.voxelize(batches[0], frames, out, want_point_pillar=False)
            torch.cuda.synchronize()
            ts = []
            for i in range(args.iters + 3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                g.replay()
                e1.record()
                torch.cuda.synchronize()
                if i >= 3:
                    ts.append(e0.elapsed_time(e1) * 1e3)
            row[method + "_graph"] = statistics.median(ts)
            row["pillars_" + method] = int(fe.read_counts(out)[0])
            del g
        res[name] = row
        print(f"{name:16s} " + "  ".join(f"{k} {v:.1f}" if isinstance(v, float) else f"{k} {v}" for k, v in row.items()), flush=True)
    if args.json:
        json.dump(res, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
