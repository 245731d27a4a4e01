"""MoDAR exchange: received detections + a timestamp -> points appended to the ego sweep stack.

Replaces the exchange block of the reference's lately-fusion dataset
(pcdet/datasets/v2x_sim/v2x_sim_dataset_ego.py:181-232) and its online twin
(workspace/visualize_collab.py:118-142,248-265): box-pool the agent's foreground flow, shift each box by
``scale * mean(flow)`` (scale = 2 for the reference's 0.2 s latency), map centre + heading into the ego
frame with the fp64 SE(3), and pack ``[x,y,z,0,0,dx,dy,dz,heading,score,label,max_sweep_idx,-1]`` rows
after the ego rows.  All agents of a frame are processed by ONE kernel launch (csrc/modar.cu).
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import Dict, Optional, Sequence, Union

import numpy as np
import torch

from . import _lib
from .exchange import ExchangeMessage
from .frontend import _ptr, _stream

SAMPLE_INTERVAL_S = 0.2   # V2X-Sim key frames are 5 Hz (README.md:45-46); flow = displacement to the newest sweep


def flow_scale(t_detect: float, t_query: float, sample_interval: float = SAMPLE_INTERVAL_S) -> float:
    """Multiplier applied to the mean foreground flow.  The mean flow over uniformly populated sweeps is half
    the displacement over one sample interval, so one interval of latency gives exactly 2.0 - the literal
    in v2x_sim_dataset_ego.py:213; zero latency gives 0.0 = EXCHANGE_NOW (:234-267)."""
    latency = float(t_query) - float(t_detect)
    if latency < 0:
        raise ValueError(f"t_query {t_query} is earlier than t_detect {t_detect}")
    ratio = latency / sample_interval
    if abs(ratio - round(ratio)) < 1e-6:      # timestamps are float seconds: snap 0.2000000001 to one interval
        ratio = float(round(ratio))
    return 2.0 * ratio


def _as_boxes9(det: Union[Dict[str, torch.Tensor], torch.Tensor], device) -> torch.Tensor:
    """detections dict {'pred_boxes' (M,7), 'pred_scores' (M,), 'pred_labels' (M,)} (the layout CenterHead
    emits, center_head.py:409-427), an (M, 9) tensor box7|score|label, or an unpacked exchange.ExchangeMessage."""
    if isinstance(det, ExchangeMessage):
        det = det.boxes
    if isinstance(det, dict):
        b = det["pred_boxes"].to(device=device, dtype=torch.float32)
        s = det["pred_scores"].to(device=device, dtype=torch.float32).reshape(-1, 1)
        l = det["pred_labels"].to(device=device, dtype=torch.float32).reshape(-1, 1)
        if b.shape[1] != 7:
            raise ValueError(f"pred_boxes must be (M, 7), got {tuple(b.shape)}")
        return torch.cat([b, s, l], dim=1).contiguous()
    t = det.to(device=device, dtype=torch.float32).contiguous()
    if t.dim() != 2 or t.shape[1] != 9:
        raise ValueError(f"modar tensor must be (M, 9) box7|score|label, got {tuple(t.shape)}")
    return t


_staging = {}   # (device, stream, thread) -> [pinned uint8 buffer, event of the last H2D that read it]
_staging_lock = threading.Lock()


class _DevPtr:
    """A device address inside a staging upload (what ``frontend._ptr`` returns for a tensor); ``keep`` pins the buffer."""
    __slots__ = ("ptr", "keep")

    def __init__(self, ptr, keep):
        self.ptr, self.keep = ptr, keep

    def data_ptr(self):
        return self.ptr


_RING = 4     # staging buffers per (device, stream, thread): a call only waits for the copy issued four calls earlier


def _upload(dev, arrays):
    """One asynchronous H2D copy for all the small host-side arrays of a call (poses, offsets, pointers): they are packed
    into a pinned staging buffer (8-byte aligned pieces) and come back as device addresses inside one device buffer.  A small
    ring of staging buffers per (device, stream, host thread): concurrent callers never share one, and a caller does not wait
    for its own previous copy."""
    sizes = [(a.nbytes + 7) // 8 * 8 for a in arrays]
    total = max(sum(sizes), 8)
    key = (dev, torch.cuda.current_stream(dev).cuda_stream, threading.get_ident())
    with _staging_lock:
        ring = _staging.get(key)
        if ring is None:
            ring = {"slots": [None] * _RING, "next": 0}
            _staging[key] = ring
        i = ring["next"]
        ring["next"] = (i + 1) % _RING
        st = ring["slots"][i]
        if st is None or st[0].numel() < total:
            st = [torch.empty(max(total, 4096), dtype=torch.uint8).pin_memory(), None]
            ring["slots"][i] = st
    if st[1] is not None:
        st[1].synchronize()                      # the copy issued _RING calls ago has left this staging buffer
    host = st[0].numpy()
    offs, o = [], 0
    for a, sz in zip(arrays, sizes):
        host[o:o + a.nbytes].view(a.dtype)[:] = a.reshape(-1)
        offs.append(o)
        o += sz
    d = torch.empty(total, dtype=torch.uint8, device=dev)
    d.copy_(st[0][:total], non_blocking=True)
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(dev))
    st[1] = ev
    base = d.data_ptr()
    return [_DevPtr(base + off, d) for off in offs]


def modar_exchange(detections, foreground, target_se3_agent, t_detect: float, t_query: float,
                   ego_points: torch.Tensor, max_sweep_idx: Optional[float] = None, batch_idx: float = 0.0,
                   sample_interval: float = SAMPLE_INTERVAL_S, return_box_idx: bool = False):
    """One frame of lately fusion.

    detections / foreground / target_se3_agent: one agent, or equal-length lists for several agents
        (detections: dict, (M,9) tensor or an unpacked exchange.ExchangeMessage - whose foreground records are used when
        `foreground` is None; foreground: (F,13) tensor or None; target_se3_agent: (4,4) float64).
    t_detect, t_query: when the agents detected and when the ego asks (seconds).
    ego_points: (N,13) ego sweep stack ``[pt5, 0 x6, sweep_idx, inst_idx]`` (v2x_sim_dataset_ego.py:162-165) or
        (N,14) with the collate_batch frame-index column in front; must live on the GPU.
    max_sweep_idx: written to every MoDAR row (:225); default = ego_points[:, sweep column].max().
    Returns (N + sum M, same columns) fp32 on the GPU.
    """
    if not ego_points.is_cuda:
        raise RuntimeError("ego_points must be on the GPU: pcp_b200 has no CPU path")
    # received records may live anywhere (host, another GPU): they are moved to the ego cloud's GPU, which is made current
    with torch.cuda.device(ego_points.device):
        return _modar_exchange(detections, foreground, target_se3_agent, t_detect, t_query, ego_points, max_sweep_idx,
                               batch_idx, sample_interval, return_box_idx)


def _modar_exchange(detections, foreground, target_se3_agent, t_detect, t_query, ego_points, max_sweep_idx, batch_idx,
                    sample_interval, return_box_idx):
    lib = _lib.load()
    dev = ego_points.device
    # ExchangeMessage is a NamedTuple: test for it before the list / tuple test, or one message reads as four agents
    if isinstance(detections, ExchangeMessage) or not isinstance(detections, (list, tuple)):
        detections, foreground, target_se3_agent = [detections], [foreground], [target_se3_agent]
        if foreground[0] is None:
            foreground = None
    if foreground is None:          # messages carry their own foreground records
        foreground = [d.foreground if isinstance(d, ExchangeMessage) and d.foreground.shape[0] else None for d in detections]
    if not (len(detections) == len(foreground) == len(target_se3_agent)):
        raise ValueError("detections, foreground and target_se3_agent must have the same length")
    ncol = ego_points.shape[1]
    if ncol not in (13, 14):
        raise ValueError(f"ego_points must have 13 or 14 columns, got {ncol}")
    with_b = ncol == 14
    ego = ego_points if (ego_points.dtype == torch.float32 and ego_points.is_contiguous()) else ego_points.float().contiguous()
    scale = flow_scale(t_detect, t_query, sample_interval)

    boxes = [_as_boxes9(d, dev) for d in detections]
    fgs = []
    for f in foreground:
        if f is None or scale == 0.0:
            fgs.append(torch.empty((0, 13), dtype=torch.float32, device=dev))
        else:
            f = f.to(device=dev, dtype=torch.float32).contiguous()
            if f.dim() != 2 or f.shape[1] != 13:
                raise ValueError(f"foreground must be (F, 13), got {tuple(f.shape)}")
            fgs.append(f)
    n_agents = len(boxes)
    m_tot = sum(b.shape[0] for b in boxes)
    n_ego = ego.shape[0]
    out = torch.empty((n_ego + m_tot, ncol), dtype=torch.float32, device=dev)
    out[:n_ego].copy_(ego)                                   # np.concatenate((points_, modar_)) (:232): D2D copy
    if m_tot == 0:
        return (out, None) if return_box_idx else out

    box_off = np.zeros(n_agents + 1, dtype=np.int32)
    fg_off = np.zeros(n_agents + 1, dtype=np.int32)
    box_off[1:] = np.cumsum([b.shape[0] for b in boxes])
    fg_off[1:] = np.cumsum([f.shape[0] for f in fgs])
    for t in target_se3_agent:
        if np.asarray(t).shape != (4, 4):
            raise ValueError("target_se3_agent must be (4, 4)")
    se3 = np.stack([np.asarray(t, dtype=np.float64)[:3, :4].reshape(12) for t in target_se3_agent])
    # Nothing is gathered and nothing is read back: the kernel takes one device pointer per agent (records stay where the
    # exchange format left them) and the sweep index from device memory (reduced there when the caller did not give it).
    box_ptrs = np.asarray([b.data_ptr() if b.shape[0] else 0 for b in boxes], dtype=np.uint64)
    fg_ptrs = np.asarray([f.data_ptr() if f.shape[0] else 0 for f in fgs], dtype=np.uint64)
    msi_host = np.asarray([0.0 if max_sweep_idx is None else float(max_sweep_idx)], dtype=np.float32)
    se3_d, meta, bp_d, fp_d, msi_d = _upload(dev, [se3, np.concatenate([box_off, fg_off]), box_ptrs, fg_ptrs, msi_host])
    if max_sweep_idx is None:                                                       # points_[:, -2].max() (:174)
        rc = lib.pcp_column_max(_ptr(ego), ego.stride(0), n_ego, ncol - 2, _ptr(msi_d), _stream())
        _lib.check(rc, "pcp_column_max")
    n_fg = int(fg_off[-1])
    have_fg = n_fg > 0
    box_idx = torch.empty((max(n_fg, 1),), dtype=torch.int32, device=dev)
    rows = out[n_ego:]
    rc = lib.pcp_modar_agents(_ptr(bp_d), C.c_void_p(meta.data_ptr()), _ptr(fp_d) if have_fg else None,
                              C.c_void_p(meta.data_ptr() + 4 * (n_agents + 1)), _ptr(se3_d), n_agents,
                              int(max(b.shape[0] for b in boxes)), int(max(f.shape[0] for f in fgs)),
                              C.c_float(scale), _ptr(msi_d), int(with_b), C.c_float(batch_idx),
                              _ptr(rows), ncol, _ptr(box_idx), _stream())
    _lib.check(rc, "pcp_modar_agents")
    if return_box_idx:
        return out, (box_idx[:n_fg] if have_fg else None)
    return out
