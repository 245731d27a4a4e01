"""BASELINE.json configs[3]: a batch of synthetic V2X-Sim frames sharded across the GPUs of one box, the per-GPU BEV
maps NCCL-gathered for validation.  No collective on the hot path; the all-gather runs after it.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/validate_multi_gpu.py [--frames 64] [--points 32768]

Every rank voxelises / encodes / scatters its own contiguous block of frames (frame indices renumbered from 0), then
    1. `sharding.gather_bev` all-gathers the (B_local, 64, ny, nx) blocks over NCCL;
    2. rank 0 runs the WHOLE batch on its own GPU, chunk by chunk, and checks that the gathered maps, the pillar
       coordinates and the pillar features are bit-identical to the unsharded run (same kernels, same order of
       operations per frame: sharding must not change a single bit);
    3. rank 0 prints one JSON line with the outcome and the gather's bus bandwidth.
Exit code 0 only if every comparison is exact.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcp_b200  # noqa: E402
from pcp_b200 import synthetic as syn  # noqa: E402
from pcp_b200.frontend import FrontEnd, GridSpec  # noqa: E402
from pcp_b200.sharding import frame_range, gather_bev, shard_points  # noqa: E402


def run_block(fe, pts, n_frames, dev):
    """voxelize -> PFN -> canvas of one block of frames; returns (coords, features, canvas) trimmed to P."""
    out = fe.forward_device(pts, n_frames)
    c = fe.read_counts(out)
    p = int(c[0])
    return out["voxel_coords_buf"][:p].clone(), out["pillar_features_buf"][:p].clone(), out["spatial_features"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--points", type=int, default=32768)
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    rng = np.asarray(syn.V2X_RANGE, dtype=np.float32)
    vox = syn.V2X_VOXEL
    grid = syn.grid_size_of(rng, vox)
    gs = GridSpec(vox, rng, grid)
    sd = syn.pfn_state_dict(11)
    fe = FrontEnd(gs, 5)
    bn = lambda i: [sd[f"pfn_layers.{i}.norm.{k}"].to(dev) for k in ("weight", "bias", "running_mean", "running_var")]
    fe.pack_params(sd["pfn_layers.0.linear.weight"].to(dev), bn(0), sd["pfn_layers.1.linear.weight"].to(dev), bn(1))

    # the same seeded batch on every rank (CPU generator), sharded by frame
    batch = syn.batch_of_frames(a.frames, a.points, 4)
    local, n_local = shard_points(batch, a.frames, rank, world)
    coords, feats, bev = run_block(fe, local.to(dev), n_local, dev)
    torch.cuda.synchronize()

    gather_gbs = None
    if world > 1:
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        full = gather_bev(bev, a.frames)                       # warm-up (communicator setup)
        e0.record()
        full = gather_bev(bev, a.frames)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        gather_gbs = full.numel() * 4 * (world - 1) / world / (ms * 1e-3) / 1e9
    else:
        full = bev

    ok = True
    detail = {}
    if rank == 0:
        # unsharded reference run on this GPU, in blocks of the shard size so that the canvas fits comfortably
        first0, last0 = frame_range(a.frames, 0, world)
        mism_bev = mism_coords = mism_feats = 0
        for r in range(world):
            f, l = frame_range(a.frames, r, world)
            blk, nb = shard_points(batch, a.frames, r, world)
            c2, f2, b2 = run_block(fe, blk.to(dev), nb, dev)
            mism_bev += int((full[f:l] != b2).sum().item())
            if r == 0:
                mism_coords += int((coords != c2).sum().item())
                mism_feats += int((feats != f2).sum().item())
        occ = int((full != 0).any(1).sum().item())
        ok = (mism_bev == 0 and mism_coords == 0 and mism_feats == 0 and occ > 0)
        detail = {"bev_mismatches": mism_bev, "coord_mismatches": mism_coords, "feature_mismatches": mism_feats,
                  "occupied_cells": occ}
        print(json.dumps({"check": "multi_gpu_bev_gather", "ok": ok, "n_gpus": world, "frames": a.frames,
                          "points_per_frame": a.points, "bev_shape": list(full.shape),
                          "all_gather_bus_GBps": gather_gbs, **detail}), flush=True)
    if world > 1:
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.broadcast(flag, 0)
        ok = bool(flag.item())
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
