"""Points loader: the host->device half of the early-fusion row of SURVEY.md section 8(f), drop-in for what the reference
does to ``batch_dict['points']`` between the dataset and the first module:

* ``collate_batch`` (pcdet/datasets/dataset.py:224-229): every sample's (Ni, C) array gets its frame index padded in
  front (``np.pad``) and the samples are concatenated;
* ``load_data_to_gpu`` (pcdet/models/__init__.py:23-34): ``torch.from_numpy(val).float().cuda().contiguous()`` - a float64
  -> float32 cast on the host and a pageable H2D copy of all 1 + C columns.

Here ``collate_points`` writes the samples straight into ONE pinned fp32 buffer, keeping only the per-point columns the
consumer reads (the pillar encoder reads the first NUM_RAW_POINT_FEATURES: x, y, z, intensity, time - 20 of an
early-fusion row's 32 bytes), and ``load_points_to_gpu`` ships that buffer with one asynchronous copy and rebuilds the
(N, 1 + C) rows on the device (csrc/fusion.cu: pcp_unpack_points; columns that were not shipped read as zero).
``columns=None`` ships every column: the rows are then exactly collate_batch's.
"""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib
from .frontend import _ptr, _stream


class PackedPoints(NamedTuple):
    data: torch.Tensor            # (N, k) fp32 host tensor (pinned when CUDA is available): shipped columns, frames in order
    frame_offsets: torch.Tensor   # (B + 1,) int32 host tensor (pinned): first row of every frame
    columns: Tuple[int, ...]      # per-point column (0-based, without the frame index) behind each shipped column
    n_point_cols: int             # C: per-point columns of a full row

    @property
    def batch_size(self) -> int:
        return int(self.frame_offsets.shape[0]) - 1

    @property
    def n_points(self) -> int:
        return int(self.data.shape[0])


def _pinned(shape, dtype):
    t = torch.empty(shape, dtype=dtype)
    return t.pin_memory() if torch.cuda.is_available() else t


def collate_points(frames: Sequence[np.ndarray], columns: Optional[Sequence[int]] = None,
                   out: Optional[PackedPoints] = None) -> PackedPoints:
    """frames[b]: (Nb, C) array (any float dtype) of sample b, as the dataset's ``__getitem__`` returns it.
    columns: per-point columns to ship (default: all).  ``out``: a PackedPoints of sufficient capacity to refill (its
    pinned buffers are reused; the returned views are trimmed to this batch)."""
    if len(frames) == 0:
        raise ValueError("collate_points: empty batch")
    c = int(frames[0].shape[1])
    cols = tuple(range(c)) if columns is None else tuple(int(v) for v in columns)
    if not cols or len(cols) > 16 or min(cols) < 0 or max(cols) >= c or len(set(cols)) != len(cols):
        raise ValueError(f"columns must be 1..16 distinct indices below {c}, got {cols}")
    counts = [int(f.shape[0]) for f in frames]
    n = sum(counts)
    if out is not None and (out.data.shape[0] < n or out.data.shape[1] != len(cols) or out.frame_offsets.shape[0] < len(frames) + 1):
        out = None
    data = out.data[:n] if out is not None else _pinned((n, len(cols)), torch.float32)
    offs = out.frame_offsets[:len(frames) + 1] if out is not None else _pinned((len(frames) + 1,), torch.int32)
    dst = data.numpy()
    o = 0
    offs_np = offs.numpy()
    for b, f in enumerate(frames):
        if f.ndim != 2 or f.shape[1] != c:
            raise ValueError("every frame must be (Nb, C) with the same C")
        offs_np[b] = o
        # one pass: column selection + the float64 -> float32 cast load_data_to_gpu does (.float())
        np.copyto(dst[o:o + counts[b]], f if columns is None else f[:, cols], casting="same_kind")
        o += counts[b]
    offs_np[len(frames)] = o
    return PackedPoints(data, offs, cols, c)


def load_points_to_gpu(points: Union[PackedPoints, np.ndarray, torch.Tensor], device, out: Optional[torch.Tensor] = None,
                       staging: Optional[dict] = None) -> torch.Tensor:
    """-> (N, 1 + C) fp32 rows on ``device``, enqueued on the current stream (no host synchronisation).

    points: a PackedPoints (see collate_points), or the (N, 1 + C) array collate_batch produced (then this is
    load_data_to_gpu's ``torch.from_numpy(val).float().cuda()`` through pinned memory).
    out: optional (>= N, 1 + C) device buffer to fill (rows past N are left alone).
    staging: optional dict that keeps the small device buffers between calls (no allocation per step)."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("load_points_to_gpu: the destination must be a CUDA device (pcp_b200 has no CPU path)")
    with torch.cuda.device(dev):
        if not isinstance(points, PackedPoints):
            t = torch.from_numpy(points) if isinstance(points, np.ndarray) else points
            t = t.float()                                              # models/__init__.py:33
            if out is not None:
                out[:t.shape[0]].copy_(t, non_blocking=True)
                return out[:t.shape[0]]
            return t.to(dev, non_blocking=True).contiguous()
        lib = _lib.load()
        n, k, c = points.n_points, len(points.columns), points.n_point_cols
        st = staging if staging is not None else {}
        d_packed = st.get("packed")
        if d_packed is None or d_packed.shape[0] < n or d_packed.shape[1] != k or d_packed.device != dev:
            d_packed = torch.empty((max(n, 1), k), dtype=torch.float32, device=dev)
            st["packed"] = d_packed
        d_off = st.get("offsets")
        if d_off is None or d_off.shape[0] < points.frame_offsets.shape[0] or d_off.device != dev:
            d_off = torch.empty((points.frame_offsets.shape[0],), dtype=torch.int32, device=dev)
            st["offsets"] = d_off
        d_packed[:n].copy_(points.data, non_blocking=True)             # the one large H2D copy: 4 k bytes per point
        d_off[:points.frame_offsets.shape[0]].copy_(points.frame_offsets, non_blocking=True)
        if out is None:
            out = torch.empty((max(n, 1), 1 + c), dtype=torch.float32, device=dev)
        elif out.shape[0] < n or out.shape[1] != 1 + c or out.device != dev or out.dtype != torch.float32 or out.stride(1) != 1:
            raise ValueError(f"out must be a (>= {n}, {1 + c}) fp32 tensor on {dev}")
        cols = (C.c_int32 * k)(*points.columns)
        rc = lib.pcp_unpack_points(_ptr(d_packed), k, cols, n, _ptr(d_off), points.batch_size, c, _ptr(out), out.stride(0),
                                   _stream())
        _lib.check(rc, "pcp_unpack_points")
        return out[:n]


class PointsPrefetcher:
    """Double-buffered form of ``load_points_to_gpu`` for a stream of batches: ``submit`` starts the H2D copy of a batch on
    a private copy stream and returns at once; ``get`` makes the caller's stream wait for the oldest submitted copy and
    rebuilds its rows there.  Copies of consecutive batches run back to back under the kernels of the batch before - the
    pin_memory + non_blocking prefetch of a PyTorch data loader, with the row rebuild kept off the copy stream.

    ``depth`` row buffers are cycled: a buffer returned by ``get`` is overwritten ``depth`` gets later, so the consumer must
    have enqueued its reads of it by then (it has, if it consumes batches in order on one stream)."""

    def __init__(self, device, depth: int = 3):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("PointsPrefetcher: the destination must be a CUDA device")
        self.depth = int(depth)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._slots = [dict(staging={}, rows=None, copied=None, consumed=None, meta=None) for _ in range(self.depth)]
        self._head = 0          # next slot to submit into
        self._tail = 0          # next slot to get from
        self._pending = 0

    def submit(self, packed: PackedPoints) -> None:
        if self._pending >= self.depth:
            raise RuntimeError("PointsPrefetcher: more than `depth` batches in flight; call get() first")
        s = self._slots[self._head]
        n, k = packed.n_points, len(packed.columns)
        st = s["staging"]
        with torch.cuda.device(self.device), torch.cuda.stream(self.copy_stream):
            if s["consumed"] is not None:
                self.copy_stream.wait_event(s["consumed"])            # the unpack that last read this slot's staging is done
            if st.get("packed") is None or st["packed"].shape[0] < n or st["packed"].shape[1] != k:
                st["packed"] = torch.empty((max(n, 1), k), dtype=torch.float32, device=self.device)
            if st.get("offsets") is None or st["offsets"].shape[0] < packed.frame_offsets.shape[0]:
                st["offsets"] = torch.empty((packed.frame_offsets.shape[0],), dtype=torch.int32, device=self.device)
            st["packed"][:n].copy_(packed.data, non_blocking=True)
            st["offsets"][:packed.frame_offsets.shape[0]].copy_(packed.frame_offsets, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        s["copied"], s["meta"] = ev, packed
        self._head = (self._head + 1) % self.depth
        self._pending += 1

    def get(self) -> torch.Tensor:
        """(N, 1 + C) rows of the oldest submitted batch, complete on the caller's current stream."""
        if self._pending == 0:
            raise RuntimeError("PointsPrefetcher: nothing submitted")
        lib = _lib.load()
        s = self._slots[self._tail]
        packed = s["meta"]
        n, k, c = packed.n_points, len(packed.columns), packed.n_point_cols
        with torch.cuda.device(self.device):
            cur = torch.cuda.current_stream()
            cur.wait_event(s["copied"])
            rows = s["rows"]
            if rows is None or rows.shape[0] < n or rows.shape[1] != 1 + c:
                rows = torch.empty((max(n, 1), 1 + c), dtype=torch.float32, device=self.device)
                s["rows"] = rows
            cols = (C.c_int32 * k)(*packed.columns)
            rc = lib.pcp_unpack_points(_ptr(s["staging"]["packed"]), k, cols, n, _ptr(s["staging"]["offsets"]), packed.batch_size, c,
                                       _ptr(rows), rows.stride(0), _stream())
            _lib.check(rc, "pcp_unpack_points")
            ev = torch.cuda.Event()
            ev.record(cur)
            s["consumed"] = ev
        self._tail = (self._tail + 1) % self.depth
        self._pending -= 1
        return rows[:n]
