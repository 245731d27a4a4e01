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
