"""Exchange wire format: what one agent broadcasts per key frame, as ONE flat buffer of fixed fp32 records.

The reference hands detections from agent to ego as torch-pickled ``.pth`` files - ``{token}_id{lidar}_modar.pth`` =
(M, 9) ``box7 | score | label`` written by CenterHead (pcdet/models/dense_heads/center_head.py:409-427) and
``{token}_id{lidar}_foreground.pth`` = (F, 13) ``point5 | sweep_idx | inst_idx | cls_prob3 | flow3`` - and reads them back
with ``torch.load`` in the dataset (pcdet/datasets/v2x_sim/v2x_sim_dataset_ego.py:192-200,246-250); the late-fusion
detector receives the same (M, 9) records through ``metadata['exchange_boxes']`` (v2x_late_fusion.py:21-26).

Here a message is a contiguous byte buffer (little endian) that can be sent over a socket, written to a file, or handed
from GPU to GPU without any (un)pickling; the records are the reference's own row layouts, so unpacking is a view:

    offset  0  uint32  magic 'PCPX'          offset 16  uint32  n_boxes   M
            4  uint32  version (1)                   20  uint32  n_foregr  F
            8  int32   agent (lidar) id              24  uint32  floats per box record  (9)
           12  uint32  header bytes (64)             28  uint32  floats per foreground record (13)
           32  float64 timestamp of the detection [s]
           40  24 bytes reserved (zero)
           64  M x 9 fp32 box records, then F x 13 fp32 foreground records

``pack_exchange`` / ``unpack_exchange`` work on whatever device the tensors live on (device-to-device copies, no host
round trip); ``modar_exchange`` and ``class_agnostic_nms`` consume the unpacked views directly.
"""
from __future__ import annotations

import struct
from typing import Dict, NamedTuple, Optional, Union

import numpy as np
import torch

MAGIC = 0x58504350            # b'PCPX' little endian
VERSION = 1
HEADER_BYTES = 64
BOX_FLOATS = 9                # box7 | score | label (center_head.py:413-417)
FOREGROUND_FLOATS = 13        # point5 | sweep_idx | inst_idx | cls_prob3 | flow3 (v2x_sim_dataset_ego.py:200)
_HEADER = struct.Struct("<IIiIIIIId24x")
assert _HEADER.size == HEADER_BYTES


class ExchangeMessage(NamedTuple):
    agent_id: int
    timestamp: float
    boxes: torch.Tensor                  # (M, 9) fp32 view of the buffer
    foreground: torch.Tensor             # (F, 13) fp32 view of the buffer (F may be 0)

    @property
    def detections(self) -> Dict[str, torch.Tensor]:
        """The dict layout CenterHead emits / modar_exchange takes."""
        return {"pred_boxes": self.boxes[:, :7], "pred_scores": self.boxes[:, 7], "pred_labels": self.boxes[:, 8].long()}


def message_bytes(n_boxes: int, n_foreground: int = 0) -> int:
    return HEADER_BYTES + 4 * (BOX_FLOATS * int(n_boxes) + FOREGROUND_FLOATS * int(n_foreground))


def _boxes9(det) -> torch.Tensor:
    if isinstance(det, dict):
        b = det["pred_boxes"]
        if b.dim() != 2 or b.shape[1] != 7:
            raise ValueError(f"pred_boxes must be (M, 7), got {tuple(b.shape)}")
        return torch.cat([b.float(), det["pred_scores"].float().reshape(-1, 1), det["pred_labels"].float().reshape(-1, 1)], dim=1)
    if det.dim() != 2 or det.shape[1] != BOX_FLOATS:
        raise ValueError(f"detections must be a dict or an (M, 9) tensor, got {tuple(det.shape)}")
    return det.float()


def pack_exchange(detections: Union[Dict[str, torch.Tensor], torch.Tensor], foreground: Optional[torch.Tensor] = None,
                  agent_id: int = 0, timestamp: float = 0.0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """-> uint8 buffer of ``message_bytes(M, F)`` bytes on the device of the detections."""
    boxes = _boxes9(detections)
    dev = boxes.device
    m = boxes.shape[0]
    if foreground is None:
        f = 0
    else:
        if foreground.dim() != 2 or foreground.shape[1] != FOREGROUND_FLOATS:
            raise ValueError(f"foreground must be (F, 13), got {tuple(foreground.shape)}")
        f = foreground.shape[0]
    n = message_bytes(m, f)
    if out is None:
        out = torch.empty(n, dtype=torch.uint8, device=dev)
    elif out.dtype != torch.uint8 or out.numel() < n or not out.is_contiguous():
        raise ValueError(f"out must be a contiguous uint8 buffer of at least {n} bytes")
    buf = out[:n]
    head = torch.from_numpy(np.frombuffer(_HEADER.pack(MAGIC, VERSION, int(agent_id), HEADER_BYTES, m, f, BOX_FLOATS,
                                                        FOREGROUND_FLOATS, float(timestamp)), dtype=np.uint8).copy())
    buf[:HEADER_BYTES].copy_(head, non_blocking=True)
    body = buf[HEADER_BYTES:].view(torch.float32)
    body[:BOX_FLOATS * m].view(m, BOX_FLOATS).copy_(boxes)
    if f:
        body[BOX_FLOATS * m:].view(f, FOREGROUND_FLOATS).copy_(foreground.to(device=dev, dtype=torch.float32))
    return buf


def select_foreground(points: torch.Tensor, points_cls_logit: torch.Tensor, points_flow3d: torch.Tensor, batch_size: int,
                      threshold: float = 0.3):
    """The foreground records every sample of a batch sends away - drop-in for the test-time block of HunterJr.forward
    (pcdet/models/bev_layers/hunter_jr.py:377-397): ``sigmoid`` of the 3-class logits, ``P(background) < threshold``, the
    ``torch.cat`` of ``[points[mask, 1:], prob[mask], flow[mask]]`` and the per-sample split, as ONE stable partition on the
    GPU (csrc/fusion.cu: pcp_select_foreground) followed by one (batch_size + 1)-int read-back.

    points (N, 1 + C) with the sample index in column 0 (C = 7: point5 | sweep_idx | inst_idx), points_cls_logit (N, 3),
    points_flow3d (N, 3), all on the same GPU.  Returns a list of ``batch_size`` (F_b, C + 6) fp32 tensors (views of one
    buffer; a sample that sends nothing gets an empty one - the reference skips those)."""
    import ctypes as C
    from . import _lib
    from .frontend import _ptr, _stream
    if not points.is_cuda:
        raise RuntimeError("select_foreground: tensors must be on the GPU (pcp_b200 has no CPU path)")
    dev = points.device
    if points_cls_logit.device != dev or points_flow3d.device != dev:
        raise RuntimeError("select_foreground: points, points_cls_logit and points_flow3d must be on the same device")
    n = int(points.shape[0])
    if points_cls_logit.shape != (n, 3) or points_flow3d.shape != (n, 3) or points.dim() != 2 or points.shape[1] < 2:
        raise ValueError("select_foreground: expected points (N, 1 + C), points_cls_logit (N, 3), points_flow3d (N, 3)")
    batch_size = int(batch_size)
    f32 = lambda t: t.detach() if (t.dtype == torch.float32 and t.stride(1) == 1) else t.detach().float().contiguous()
    pts, lg, fl = f32(points), f32(points_cls_logit), f32(points_flow3d)
    c = int(pts.shape[1]) - 1
    with torch.cuda.device(dev):
        lib = _lib.load()
        rows = torch.empty((max(n, 1), c + 6), dtype=torch.float32, device=dev)
        offs = torch.empty((batch_size + 1,), dtype=torch.int32, device=dev)
        scratch = torch.empty(int(lib.pcp_select_scratch_bytes(n, batch_size)) // 4 + 1, dtype=torch.int32, device=dev)
        rc = lib.pcp_select_foreground(_ptr(pts), pts.stride(0), c, _ptr(lg), lg.stride(0), _ptr(fl), fl.stride(0), n, batch_size,
                                       C.c_float(threshold), _ptr(scratch), _ptr(rows), rows.stride(0), _ptr(offs), _stream())
        _lib.check(rc, "pcp_select_foreground")
        o = offs.cpu().tolist()                              # the one host read (the reference syncs on torch.any(mask_send))
    return [rows[o[b]:o[b + 1]] for b in range(batch_size)]


def exchange_payloads(pred_dicts, foreground_per_sample, agent_ids, timestamps):
    """One ExchangeMessage buffer per sample of a batch: the (M, 9) ``mo_pts`` records CenterHead assembles
    (center_head.py:409-427: cat of pred_boxes | pred_scores | pred_labels) and the sample's foreground records
    (``select_foreground``), packed on the device - what the reference torch.save()s as ``*_modar.pth`` /
    ``*_foreground.pth``.  Samples without boxes send nothing (None), as in the reference."""
    out = []
    for pred, fg, aid, ts in zip(pred_dicts, foreground_per_sample, agent_ids, timestamps):
        if pred["pred_boxes"].shape[0] == 0:
            out.append(None)
            continue
        out.append(pack_exchange(pred, fg if (fg is not None and fg.shape[0]) else None, agent_id=int(aid), timestamp=float(ts)))
    return out


def parse_header(header: bytes):
    """-> (agent_id, timestamp, M, F); raises ValueError on a foreign or newer message."""
    if len(header) < HEADER_BYTES:
        raise ValueError(f"exchange message shorter than its {HEADER_BYTES}-byte header")
    magic, version, agent, hbytes, m, f, bf, ff, ts = _HEADER.unpack(bytes(header[:HEADER_BYTES]))
    if magic != MAGIC:
        raise ValueError(f"not an exchange message (magic {magic:#x})")
    if version != VERSION or hbytes != HEADER_BYTES or bf != BOX_FLOATS or ff != FOREGROUND_FLOATS:
        raise ValueError(f"unsupported exchange message: version {version}, header {hbytes}, records {bf}/{ff}")
    return agent, ts, m, f


def unpack_exchange(buf: torch.Tensor) -> ExchangeMessage:
    """Zero-copy views of the records; the 64 header bytes are the only thing read on the host."""
    if buf.dtype != torch.uint8 or buf.dim() != 1 or not buf.is_contiguous():
        raise ValueError("exchange message must be a contiguous 1-D uint8 tensor")
    if buf.numel() < HEADER_BYTES:
        raise ValueError(f"exchange message shorter than its {HEADER_BYTES}-byte header")
    agent, ts, m, f = parse_header(buf[:HEADER_BYTES].cpu().numpy().tobytes())
    need = message_bytes(m, f)
    if buf.numel() < need:
        raise ValueError(f"truncated exchange message: {buf.numel()} bytes, header announces {need}")
    body = buf[HEADER_BYTES:need].view(torch.float32)
    return ExchangeMessage(agent, ts, body[:BOX_FLOATS * m].view(m, BOX_FLOATS), body[BOX_FLOATS * m:].view(f, FOREGROUND_FLOATS))


def write_exchange(path, buf: torch.Tensor) -> None:
    """Raw bytes to a file - the replacement of ``torch.save(mo_pts, save_path)`` (center_head.py:424-425)."""
    buf.detach().cpu().numpy().tofile(str(path))


def read_exchange(path, device=None) -> ExchangeMessage:
    """The replacement of ``torch.load(path_modar)`` / ``torch.load(path_foregr)`` (v2x_sim_dataset_ego.py:196,200): one
    read into pinned memory (when a GPU is the destination) and one asynchronous H2D copy."""
    raw = np.fromfile(str(path), dtype=np.uint8)
    parse_header(raw[:HEADER_BYTES].tobytes())
    t = torch.from_numpy(raw)
    if device is not None and torch.device(device).type == "cuda":
        t = t.pin_memory().to(device, non_blocking=True)
    return unpack_exchange(t)
