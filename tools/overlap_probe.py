#!/usr/bin/env python
"""Steady-state pipelining probe: do the stages of CONSECUTIVE batches overlap on one GPU?

voxelize is bound by L2 atomics / random 4-byte transactions, the PFN by instruction issue and the canvas by HBM
writes, so the stages of different batches compete for different resources.  This tool runs the bench workload
(8 early-fusion frames per step) with the stages on separate streams (priorities configurable) and N_BUF
independent buffer sets, and prints the step time of every arrangement next to the serial one.  It also checks
that every pipelined canvas is bit-identical to the serial result.

    python tools/overlap_probe.py [--steps 40] [--out gpurun_out/overlap.json]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--points", type=int, default=300_000)
    ap.add_argument("--out", default=None)
    ap.add_argument("--l2-persist", type=float, default=0.0,
                    help="fraction of the input rows pinned in L2 (access policy window on the voxelize / PFN streams); 0 = off")
    ap.add_argument("--l2-persist-ws", type=float, default=0.0,
                    help="MB at the start of each voxelize workspace (histogram, rank map, keys, slots, sorted rows) marked "
                         "persisting in L2 on the voxelize / PFN stream, so that the canvas stream cannot evict them; 0 = off")
    args = ap.parse_args()
    from pcp_b200 import synthetic as syn
    from pcp_b200.frontend import FrontEnd, GridSpec

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    rng = np.asarray(syn.V2X_RANGE, dtype=np.float32)
    gs = GridSpec(syn.V2X_VOXEL, rng, syn.grid_size_of(rng, syn.V2X_VOXEL))
    sd = syn.pfn_state_dict(5 + 6)
    B = args.frames
    batches = [syn.batch_of_frames(B, args.points, 3, first_frame=alt * 1000).to(dev) for alt in range(2)]
    bn = lambda i: [sd[f"pfn_layers.{i}.norm.{k}"].to(dev) for k in ("weight", "bias", "running_mean", "running_var")]

    n_buf = 3
    fes, outs, canvases = [], [], []
    for _ in range(n_buf):
        fe = FrontEnd(gs, 5)
        fe.pack_params(sd["pfn_layers.0.linear.weight"].to(dev), bn(0), sd["pfn_layers.1.linear.weight"].to(dev), bn(1))
        fes.append(fe)
        outs.append({})
        canvases.append(torch.empty((B, 64, gs.ny, gs.nx), dtype=torch.float32, device=dev))
    torch.cuda.synchronize()

    # serial truth for both alternating batches
    truth = []
    for alt in range(2):
        fes[0].forward_device(batches[alt], B, outs[0], canvases[0])
        torch.cuda.synchronize()
        truth.append((canvases[0].double().sum().item(), canvases[0].view(torch.int32).to(torch.int64).sum().item()))

    # ---- optional: L2 persistence for the point rows (cudaStreamAttributeAccessPolicyWindow) ----
    import ctypes

    class _Win(ctypes.Structure):
        _fields_ = [("base_ptr", ctypes.c_void_p), ("num_bytes", ctypes.c_size_t), ("hitRatio", ctypes.c_float),
                    ("hitProp", ctypes.c_int), ("missProp", ctypes.c_int)]

    class _Attr(ctypes.Union):
        _fields_ = [("win", _Win), ("pad", ctypes.c_char * 64)]

    rt = None
    if args.l2_persist > 0 or args.l2_persist_ws > 0:
        rt = ctypes.CDLL("libcudart.so.12")
        mx, mw = ctypes.c_int(), ctypes.c_int()
        rt.cudaDeviceGetAttribute(ctypes.byref(mx), 108, 0)      # cudaDevAttrMaxPersistingL2CacheSize
        rt.cudaDeviceGetAttribute(ctypes.byref(mw), 109, 0)      # cudaDevAttrMaxAccessPolicyWindowSize
        want = mx.value if args.l2_persist > 0 else min(mx.value, int(args.l2_persist_ws * 2**20))
        rc = rt.cudaDeviceSetLimit(6, ctypes.c_size_t(want))      # cudaLimitPersistingL2CacheSize
        mx = ctypes.c_int(want)
        print(json.dumps({"max_persisting_l2": mx.value, "max_window": mw.value, "set_limit_rc": rc}), flush=True)
        max_persist, max_window = mx.value, mw.value

    def set_window(stream, t, k=0):
        if rt is None:
            return
        if args.l2_persist_ws > 0:
            ws = fes[k].ws.buf
            if ws is None:                      # first use of this buffer set: the workspace is allocated by the call itself
                return
            a = _Attr()
            nbytes = min(int(args.l2_persist_ws * 2**20), ws.numel(), max_window)
            a.win = _Win(ws.data_ptr(), nbytes, min(1.0, max_persist / nbytes), 2, 1)      # persisting hits / streaming misses
            rc = rt.cudaStreamSetAttribute(ctypes.c_void_p(stream.cuda_stream), 1, ctypes.byref(a))
            assert rc == 0, rc
            return
        a = _Attr()
        nbytes = min(t.numel() * 4, max_window)
        a.win = _Win(t.data_ptr(), nbytes, min(1.0, args.l2_persist * max_persist / nbytes), 2, 1)   # persisting / streaming
        rc = rt.cudaStreamSetAttribute(ctypes.c_void_p(stream.cuda_stream), 1, ctypes.byref(a))
        assert rc == 0, rc

    # torch.cuda.Stream: lower number = higher priority; CUDA clamps values outside the device's range
    greatest, least = -5, 0

    def run(mode, pv, pp, pc, nb, steps):
        """mode 'serial' | 'vp_c' (voxelize+PFN on one stream, canvas on another) | 'v_p_c' (three streams)"""
        ev = lambda: torch.cuda.Event(enable_timing=True)
        if mode == "serial":
            s = torch.cuda.Stream(device=dev)
            t0, t1 = ev(), ev()
            with torch.cuda.stream(s):
                for i in range(5):
                    fes[0].forward_device(batches[i & 1], B, outs[0], canvases[0])
                t0.record(s)
                for i in range(steps):
                    set_window(s, batches[i & 1])
                    fes[0].forward_device(batches[i & 1], B, outs[0], canvases[0])
                t1.record(s)
            torch.cuda.synchronize()
            return t0.elapsed_time(t1) / steps * 1e3, True
        sv = torch.cuda.Stream(device=dev, priority=pv)
        sp = sv if mode == "vp_c" else torch.cuda.Stream(device=dev, priority=pp)
        sc = torch.cuda.Stream(device=dev, priority=pc)
        done_c = [None] * nb
        t0, t1 = ev(), ev()
        ok = True

        def loop(n, check):
            nonlocal ok
            for i in range(n):
                k = i % nb
                pts = batches[i & 1]
                if done_c[k] is not None:
                    sv.wait_event(done_c[k])            # buffer set k is free again
                set_window(sv, pts, k)
                if sp is not sv:
                    set_window(sp, pts, k)
                with torch.cuda.stream(sv):
                    fes[k].voxelize(pts, B, outs[k], want_point_pillar=False)
                    e_v = torch.cuda.Event()
                    e_v.record(sv)
                if sp is not sv:
                    sp.wait_event(e_v)
                with torch.cuda.stream(sp):
                    fes[k].pfn(pts, outs[k])
                    e_p = torch.cuda.Event()
                    e_p.record(sp)
                sc.wait_event(e_p)
                with torch.cuda.stream(sc):
                    fes[k].scatter_ws(outs[k]["pillar_features_buf"], B, canvases[k])
                    e_c = torch.cuda.Event()
                    e_c.record(sc)
                done_c[k] = e_c
                if check and i >= n - nb:
                    e_c.synchronize()
                    got = (canvases[k].double().sum().item(), canvases[k].view(torch.int32).to(torch.int64).sum().item())
                    ok = ok and got == truth[i & 1]

        loop(6, False)
        torch.cuda.synchronize()
        t0.record(sv)
        loop(steps, False)
        for e in done_c:
            if e is not None:
                sv.wait_event(e)
        t1.record(sv)
        torch.cuda.synchronize()
        us = t0.elapsed_time(t1) / steps * 1e3
        for k in range(nb):
            done_c[k] = None
        loop(2 * nb, True)
        torch.cuda.synchronize()
        return us, ok

    results = []
    plans = [("serial", 0, 0, 0, 1)]
    for nb in (2, 3):
        plans += [("vp_c", least, least, least, nb), ("vp_c", greatest, greatest, least, nb), ("vp_c", least, least, greatest, nb)]
    plans += [("v_p_c", least, least, least, 3), ("v_p_c", greatest, least, least, 3), ("v_p_c", greatest, greatest, least, 3),
              ("v_p_c", least, greatest, least, 3), ("v_p_c", greatest, least, greatest, 3)]
    if args.l2_persist > 0 or args.l2_persist_ws > 0:
        plans = [("serial", 0, 0, 0, 1), ("vp_c", greatest, greatest, least, 2), ("vp_c", greatest, greatest, least, 3)]
    for mode, pv, pp, pc, nb in plans:
        us, ok = run(mode, pv, pp, pc, nb, args.steps)
        r = {"mode": mode, "prio_voxelize": pv, "prio_pfn": pp, "prio_canvas": pc, "buffers": nb, "us_per_step": us,
             "frames_per_s": B / (us * 1e-6), "bit_identical": ok}
        results.append(r)
        print(json.dumps(r), flush=True)
    if args.out:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        json.dump({"priority_range": [greatest, least], "env": {k: v for k, v in os.environ.items() if k.startswith("PCP_")},
                   "results": results}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
