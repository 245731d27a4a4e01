"""Host side of the point->BEV front end: owns device buffers and enqueues the C-ABI calls.

PyTorch is used for device memory and streams only; every arithmetic step runs in libpcp_b200.so.
``FrontEnd`` is the object both the drop-in modules (modules.py) and bench.py drive.
"""
from __future__ import annotations

import ctypes as C
import functools
from dataclasses import dataclass
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import PcpGrid, PcpPfnDesc


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: pcp_b200 has no CPU path")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream() -> C.c_void_p:
    """The current stream of the CURRENT device: every public wrapper runs under ``device_guard``, which makes the
    device of its tensor arguments current, so this is the stream torch itself would launch on for them."""
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _cuda_devices(obj, found, depth=0):
    if isinstance(obj, torch.Tensor):
        if obj.is_cuda:
            found.add(obj.device)
    elif isinstance(obj, dict) and depth < 3:
        for v in obj.values():
            _cuda_devices(v, found, depth + 1)
    elif isinstance(obj, (list, tuple)) and depth < 3:
        for v in obj:
            _cuda_devices(v, found, depth + 1)


def device_guard(fn):
    """Run ``fn`` with the GPU of its tensor arguments as the current device (the library launches on the current
    device's current stream and never calls cudaSetDevice).  Tensors on two different GPUs raise: the reference's
    torch ops would raise there too, these kernels would silently dereference a foreign pointer."""
    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        found = set()
        _cuda_devices(args, found)
        _cuda_devices(kwargs, found)
        if len(found) > 1:
            raise RuntimeError(f"{fn.__qualname__}: tensors on different GPUs {sorted(str(d) for d in found)}")
        if not found:
            return fn(*args, **kwargs)
        with torch.cuda.device(next(iter(found))):
            return fn(*args, **kwargs)
    return wrapped


DEFAULT_VOXELIZE_METHOD = "auto"


@dataclass
class GridSpec:
    """Constants DynamicPillarVFE.__init__ derives (dynamic_pillar_vfe.py:77-89), evaluated with the
    caller's own scalar types exactly as the reference does, then rounded to fp32 for the kernels."""
    voxel_size: Sequence[float]
    point_cloud_range: Sequence[float]
    grid_size: Sequence[int]

    def __post_init__(self):
        vs, rng = self.voxel_size, self.point_cloud_range
        self.voxel_x, self.voxel_y, self.voxel_z = vs[0], vs[1], vs[2]
        self.x_offset = self.voxel_x / 2 + rng[0]          # :80
        self.y_offset = self.voxel_y / 2 + rng[1]          # :81
        self.z_offset = self.voxel_z / 2 + rng[2]          # :82
        self.nx, self.ny, self.nz = (int(g) for g in self.grid_size)
        # torch.tensor(voxel_size).cuda() / torch.tensor(point_cloud_range).cuda() are fp32 tensors (:88-89)
        f32 = lambda v: float(np.float32(v))
        self.c = PcpGrid(f32(rng[0]), f32(rng[1]), f32(vs[0]), f32(vs[1]),
                         f32(self.x_offset), f32(self.y_offset), f32(self.z_offset), self.nx, self.ny)


class Workspace:
    """Caller-owned scratch for one batch; remembers which problem filled it."""

    def __init__(self):
        self.buf: Optional[torch.Tensor] = None
        self.n_points = 0
        self.max_frames = 0
        self.generation = 0

    def ensure(self, nbytes: int, device) -> torch.Tensor:
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, device=device)
        return self.buf


class FrontEnd:
    """voxelize -> PFN -> BEV scatter on one GPU.  No host synchronisation inside any method except
    ``read_counts`` (the single D2H read of P the data-dependent output shape needs)."""

    def __init__(self, grid: GridSpec, c_raw: int, use_absolute_xyz: bool = True, with_distance: bool = False,
                 num_filters: Sequence[int] = (64, 64), voxelize_method: Optional[str] = None):
        """voxelize_method: "auto" (= "histogram", the faster one on the B200), "histogram", "radix" (stable radix sort;
        raises outside the sizes it covers) or "radix_or_auto" (radix where it applies) - pcp_voxelize_method, identical
        results; None = DEFAULT_VOXELIZE_METHOD."""
        self.lib = _lib.load()
        self.grid = grid
        voxelize_method = voxelize_method or DEFAULT_VOXELIZE_METHOD
        if voxelize_method not in _lib.VOXELIZE_METHODS:
            raise ValueError(f"voxelize_method must be one of {sorted(_lib.VOXELIZE_METHODS)}")
        self.voxelize_method = voxelize_method
        self._counts_pinned = None
        nf = list(num_filters)
        if len(nf) not in (1, 2) or nf[-1] != 64 or (len(nf) == 2 and nf[0] != 64):
            raise NotImplementedError(
                f"NUM_FILTERS={nf}: the fused sm_100a PFN kernel implements [64] and [64, 64] "
                "(every DynPillarVFE config the reference ships uses [64, 64])")
        self.desc = PcpPfnDesc(int(c_raw), int(bool(use_absolute_xyz)), int(bool(with_distance)), len(nf),
                               nf[0] // 2 if len(nf) == 2 else 0, nf[-1])
        self.c_out = nf[-1]
        self.c_in = int(c_raw) + (6 if use_absolute_xyz else 3) + (1 if with_distance else 0)
        self.n_param_floats = int(self.lib.pcp_pfn_param_floats(C.byref(self.desc)))
        if self.n_param_floats == 0:
            raise RuntimeError("pcp_pfn_param_floats returned 0")
        self.ws = Workspace()
        self.packed: Optional[torch.Tensor] = None

    # ------------------------------------------------------------------ parameters
    @device_guard
    def pack_params(self, w0, bn0, w1=None, bn1=None, lin_bias0=None, lin_bias1=None, eps: float = 1e-3):
        """bn = (weight, bias, running_mean, running_var) or None.  Runs pcp_pack_pfn_params on the GPU."""
        dev = w0.device
        _require_cuda(w0, "w0")
        self.packed = torch.empty(self.n_param_floats, dtype=torch.float32, device=dev)
        keep = []

        def f(t):
            if t is None:
                return None
            t = t.detach().to(device=dev, dtype=torch.float32).contiguous()
            keep.append(t)
            return _ptr(t)

        b0 = [f(t) for t in bn0] if bn0 is not None else [None] * 4
        b1 = [f(t) for t in bn1] if bn1 is not None else [None] * 4
        rc = self.lib.pcp_pack_pfn_params(C.byref(self.desc), f(w0), f(lin_bias0), *b0, f(w1), f(lin_bias1), *b1,
                                          C.c_float(eps), _ptr(self.packed), _stream())
        _lib.check(rc, "pcp_pack_pfn_params")
        self._keep = keep  # keep sources alive until the stream has consumed them
        return self.packed

    # ------------------------------------------------------------------ stages
    def radix_applies(self, n_points: int, max_frames: int) -> bool:
        """What PCP_VOXELIZE_RADIX covers (include/pcp_b200.h): 1 .. 16.6 M rows, at most 4 M cells."""
        return 1 <= n_points <= 256 * 65024 and int(max_frames) * self.grid.nx * self.grid.ny <= 4096 * 1024

    def binned_applies(self, n_points: int, max_frames: int) -> bool:
        """What PCP_VOXELIZE_BINNED covers (include/pcp_b200.h): at least one row, at most 2048 x 2048 = 4.2 M cells."""
        return n_points >= 1 and int(max_frames) * self.grid.nx * self.grid.ny <= 2048 * 2048

    def capacity(self, n_points: int, max_frames: int) -> int:
        return max(1, min(int(n_points), int(max_frames) * self.grid.nx * self.grid.ny))

    @device_guard
    def voxelize(self, points: torch.Tensor, max_frames: int, out: Optional[Dict[str, torch.Tensor]] = None,
                 want_point_pillar: bool = True, want_counts_per_pillar: bool = False,
                 host_counts: bool = False) -> Dict[str, torch.Tensor]:
        """host_counts: also start the 32-byte copy of the counts block to pinned host memory right behind the voxelize
        kernels and record an event, so that ``read_counts`` waits for voxelize only - not for whatever the caller
        enqueues next (the module enqueues the PFN before it reads P)."""
        _require_cuda(points, "points")
        if points.dtype != torch.float32 or points.dim() != 2 or points.stride(1) != 1:
            raise RuntimeError("points must be a 2-D fp32 tensor with unit column stride")
        n, stride = points.shape[0], points.stride(0)
        dev = points.device
        cap = self.capacity(n, max_frames)
        nbytes = int(self.lib.pcp_workspace_bytes(n, max_frames, self.grid.nx, self.grid.ny))
        ws = self.ws.ensure(nbytes, dev)
        out = {} if out is None else out

        def buf(name, shape, dtype):
            t = out.get(name)
            if t is None or t.shape[0] < shape[0] or t.device != dev:
                t = torch.empty(shape, dtype=dtype, device=dev)
                out[name] = t
            return t

        coords = buf("voxel_coords_buf", (cap, 4), torch.int32)
        counts = buf("counts", (_lib.PCP_COUNTS_LEN,), torch.int32)
        pp = buf("point_pillar", (max(n, 1),), torch.int32) if want_point_pillar else None
        pc = buf("pillar_count_buf", (cap,), torch.int32) if want_counts_per_pillar else None
        method = _lib.VOXELIZE_METHODS[self.voxelize_method]
        if self.voxelize_method == "radix_or_auto" and not self.radix_applies(n, max_frames):
            method = _lib.VOXELIZE_METHODS["auto"]
        if self.voxelize_method == "binned_or_auto" and not self.binned_applies(n, max_frames):
            method = _lib.VOXELIZE_METHODS["auto"]
        rc = self.lib.pcp_voxelize_method(_ptr(points), stride, n, int(max_frames), C.byref(self.grid.c), _ptr(ws), ws.numel(),
                                          _ptr(pp), _ptr(coords), _ptr(pc), coords.shape[0], _ptr(counts), method, _stream())
        _lib.check(rc, "pcp_voxelize_method")
        self.ws.n_points, self.ws.max_frames = n, int(max_frames)
        self.ws.generation += 1
        out["capacity"] = cap
        out.pop("counts_host", None)
        if host_counts:
            # a small ring of pinned blocks: the counts of a call stay readable until eight later calls have been made
            if self._counts_pinned is None:
                self._counts_pinned = torch.empty((8, _lib.PCP_COUNTS_LEN), dtype=torch.int32).pin_memory()
                self._counts_slot = 0
            pinned = self._counts_pinned[self._counts_slot]
            self._counts_slot = (self._counts_slot + 1) % 8
            pinned.copy_(counts, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            out["counts_host"] = (pinned, ev)
        return out

    @device_guard
    def pfn(self, points: torch.Tensor, out: Dict[str, torch.Tensor], want_mean: bool = False,
            stages: int = 3) -> Dict[str, torch.Tensor]:
        """stages: 1 = the tensor-core kernel, 2 = the pillars above 32 points, 3 = both (pcp_pfn_stages)."""
        if self.packed is None:
            raise RuntimeError("pack_params() must be called before pfn()")
        n, stride = points.shape[0], points.stride(0)
        cap = out["voxel_coords_buf"].shape[0]
        pf = out.get("pillar_features_buf")
        if pf is None or pf.shape[0] < cap:
            pf = torch.empty((cap, self.c_out), dtype=torch.float32, device=points.device)
            out["pillar_features_buf"] = pf
        mean = None
        if want_mean:
            mean = out.get("pillar_mean_buf")
            if mean is None or mean.shape[0] < cap:
                mean = torch.empty((cap, 3), dtype=torch.float32, device=points.device)
                out["pillar_mean_buf"] = mean
        ws = self.ws.buf
        rc = self.lib.pcp_pfn_stages(_ptr(points), stride, n, self.ws.max_frames, C.byref(self.grid.c), C.byref(self.desc),
                                     _ptr(self.packed), _ptr(ws), ws.numel(), _ptr(pf), _ptr(mean), cap, int(stages), _stream())
        _lib.check(rc, "pcp_pfn_stages")
        return out

    @device_guard
    def scatter_ws(self, pillar_features: torch.Tensor, num_frames: int,
                   canvas: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Dense canvas from the workspace the last voxelize() left (no search, no memset)."""
        g = self.grid
        c = pillar_features.shape[1]
        if canvas is None:
            canvas = torch.empty((num_frames, c, g.ny, g.nx), dtype=torch.float32, device=pillar_features.device)
        ws = self.ws.buf
        rc = self.lib.pcp_bev_scatter_ws(_ptr(pillar_features), c, int(num_frames), self.ws.n_points, self.ws.max_frames,
                                         C.byref(g.c), _ptr(ws), ws.numel(), _ptr(canvas), _stream())
        _lib.check(rc, "pcp_bev_scatter_ws")
        return canvas

    @device_guard
    def segment_reduce(self, values: torch.Tensor, mode: str, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """scatter_mean / scatter_max of per-point ``values`` (indexed by original row) over the pillars of
        the last voxelize().  Returns the (capacity, C) buffer; the first P rows are valid."""
        _require_cuda(values, "values")
        assert values.dim() == 2 and values.stride(1) == 1 and values.dtype == torch.float32
        assert values.shape[0] == self.ws.n_points
        cap = self.capacity(self.ws.n_points, self.ws.max_frames)
        c = values.shape[1]
        if out is None:
            out = torch.empty((cap, c), dtype=torch.float32, device=values.device)
        ws = self.ws.buf
        rc = self.lib.pcp_segment_reduce(_ptr(values), values.stride(0), c, {"mean": 0, "max": 1}[mode],
                                         self.ws.n_points, self.ws.max_frames, self.grid.nx, self.grid.ny,
                                         _ptr(ws), ws.numel(), _ptr(out), cap, _stream())
        _lib.check(rc, "pcp_segment_reduce")
        return out

    # ------------------------------------------------------------------ whole chain
    @device_guard
    def forward_device(self, points: torch.Tensor, max_frames: int, out: Optional[Dict[str, torch.Tensor]] = None,
                       canvas: Optional[torch.Tensor] = None, want_point_pillar: bool = False):
        """voxelize -> PFN -> canvas for ``max_frames`` frames, all enqueued, nothing read back.
        The canvas covers all ``max_frames`` frames (the reference sizes it from the last non-empty frame,
        pointpillar_scatter.py:17; the modules do that after reading the counts)."""
        out = self.voxelize(points, max_frames, out, want_point_pillar=want_point_pillar)
        self.pfn(points, out)
        out["spatial_features"] = self.scatter_ws(out["pillar_features_buf"], max_frames, canvas)
        return out

    @staticmethod
    def read_counts(out: Dict[str, torch.Tensor]) -> np.ndarray:
        """The one D2H read of the path: 32 bytes (P, N', frames, bad-frame count, max points per pillar)."""
        early = out.get("counts_host")
        if early is not None:
            pinned, ev = early
            ev.synchronize()
            return pinned.numpy().copy()
        return out["counts"].cpu().numpy()


class PipelinedFrontEnd:
    """Steady-state throughput mode of the chain: batch i + 1 is voxelised while the canvas of batch i is written.

    voxelize + PFN run on a high-priority stream, the canvas writer on a low-priority one, over ``depth`` independent
    buffer sets (workspace, pillar buffers, canvas).  The voxelize kernels are bound by scattered L1/L2 transactions and
    the canvas by HBM writes, so the block scheduler can interleave them; results are bit-identical to the serial
    ``FrontEnd.forward_device`` (tests/test_gpu_parity.py).  No host synchronisation: ``submit`` returns the buffer set
    with a ``done`` event the consumer waits on; a set is reused ``depth`` submits later, after its canvas has been
    written AND the consumer has had the chance to read it (``release`` records the consumer's stream).
    """

    def __init__(self, grid: GridSpec, c_raw: int, max_frames: int, depth: int = 2, priorities: Tuple[int, int] = (-5, 0),
                 **kwargs):
        """priorities: CUDA stream priorities of the (voxelize + PFN, canvas) branches; lower = more urgent."""
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.priorities = (int(priorities[0]), int(priorities[1]))
        self.max_frames = int(max_frames)
        self.stages = [FrontEnd(grid, c_raw, **kwargs) for _ in range(depth)]
        self.sets = [{} for _ in range(depth)]
        self.grid = grid
        self._n = 0
        self._streams = None

    def pack_params(self, *args, **kwargs):
        packed = self.stages[0].pack_params(*args, **kwargs)
        for fe in self.stages[1:]:
            fe.packed = packed
        return packed

    def _ensure_streams(self, device):
        if self._streams is None or self._streams[0].device != device:
            # lower number = higher priority; CUDA clamps to the device's range
            self._streams = (torch.cuda.Stream(device=device, priority=self.priorities[0]),
                             torch.cuda.Stream(device=device, priority=self.priorities[1]))
        return self._streams

    @device_guard
    def submit(self, points: torch.Tensor, record=None) -> Dict[str, torch.Tensor]:
        """Enqueue one batch.  ``record``: optional 6 timing events (voxelize / PFN / canvas begin and end)."""
        _require_cuda(points, "points")
        dev = points.device
        s_main, s_canvas = self._ensure_streams(dev)
        k = self._n % len(self.sets)
        self._n += 1
        fe, out = self.stages[k], self.sets[k]
        caller = torch.cuda.current_stream(dev)
        ready = torch.cuda.Event()
        ready.record(caller)                                # the caller's writes to `points` are ordered before us
        s_main.wait_event(ready)
        if out.get("done") is not None:
            s_main.wait_event(out["done"])                  # this set's previous canvas has been written
        if out.get("released") is not None:
            s_main.wait_event(out["released"])              # ... and read by its consumer
        with torch.cuda.stream(s_main):
            if record:
                record[0].record(s_main)
            fe.voxelize(points, self.max_frames, out, want_point_pillar=False)
            if record:
                record[1].record(s_main)
                record[2].record(s_main)
            fe.pfn(points, out, stages=1)
            if record:
                record[3].record(s_main)
            e_p = torch.cuda.Event()
            e_p.record(s_main)
            if out.get("spatial_features") is None:
                g = self.grid
                out["spatial_features"] = torch.empty((self.max_frames, fe.c_out, g.ny, g.nx), dtype=torch.float32, device=dev)
        s_canvas.wait_event(e_p)
        with torch.cuda.stream(s_canvas):
            fe.pfn(points, out, stages=2)                   # the few pillars above 32 points: off the main stream's critical path
            if record:
                record[4].record(s_canvas)
            fe.scatter_ws(out["pillar_features_buf"], self.max_frames, out["spatial_features"])
            if record:
                record[5].record(s_canvas)
            done = torch.cuda.Event()
            done.record(s_canvas)
        out["done"] = done
        out["released"] = None
        out["points"] = points                              # keep the input alive until the set is reused
        return out

    # ------------------------------------------------------------------ CUDA graphs
    @device_guard
    def capture(self, static_points: Sequence[torch.Tensor]) -> None:
        """Steady state as CUDA graphs (one per buffer set): graph k holds voxelize + PFN of set k on the high-priority branch
        beside the canvas of set k - 1 on the low-priority branch, i.e. exactly the overlap ``submit`` produces, in ONE launch
        per batch instead of nine launches and a dozen event operations (the host stops being the bottleneck when eight ranks
        share one CPU).  ``static_points[k]`` is the device buffer set k always reads: the server copies each batch into it
        (rows past the batch's length filled with out-of-range points, which the cull drops) before ``replay(k)``.
        After ``replay(k)`` every output of set k - 1 (pillar features, coordinates, canvas) is complete on the current stream;
        of set k the pillars of up to 32 points are (the few longer ones are finished on the canvas branch of the next graph).
        ``flush(k)`` completes the last batch."""
        depth = len(self.sets)
        if depth < 2:
            raise ValueError("capture() needs depth >= 2 (the canvas of one set runs beside the voxelize kernels of the next)")
        if len(static_points) != depth:
            raise ValueError(f"capture() takes one static input buffer per buffer set ({depth})")
        dev = static_points[0].device
        s_main, s_canvas = self._ensure_streams(dev)
        self._n = 0
        for pts in static_points:                 # eager pass: allocates every buffer, sets the kernel attributes
            self.submit(pts)
        self.drain()
        torch.cuda.synchronize(dev)
        graphs = []
        for k in range(depth):
            prev = (k - 1) % depth
            g = torch.cuda.CUDAGraph()
            # thread_local: other host threads (NCCL watchdog, clock sampler) may keep calling the CUDA runtime during the capture
            with torch.cuda.graph(g, stream=s_main, capture_error_mode="thread_local"):
                fork = torch.cuda.Event()
                fork.record(s_main)
                s_canvas.wait_event(fork)
                with torch.cuda.stream(s_canvas):
                    self.stages[prev].pfn(static_points[prev], self.sets[prev], stages=2)
                    self.stages[prev].scatter_ws(self.sets[prev]["pillar_features_buf"], self.max_frames,
                                                 self.sets[prev]["spatial_features"])
                    join = torch.cuda.Event()
                    join.record(s_canvas)
                self.stages[k].voxelize(static_points[k], self.max_frames, self.sets[k], want_point_pillar=False)
                self.stages[k].pfn(static_points[k], self.sets[k], stages=1)
                s_main.wait_event(join)
            graphs.append(g)
        self._graphs = graphs
        self._static = list(static_points)
        for out in self.sets:                     # graph replays are ordered by the stream they are launched on
            out["done"] = None
            out["released"] = None

    def replay(self, k: int) -> Dict[str, torch.Tensor]:
        """One batch through graph k on the current stream; returns buffer set k (complete after the NEXT replay / flush)."""
        with torch.cuda.device(self._static[k].device):
            self._graphs[k].replay()
        return self.sets[k]

    def flush(self, k: int) -> Dict[str, torch.Tensor]:
        """Long pillars + canvas of the batch last replayed through graph k (end of a stream of batches)."""
        self.stages[k].pfn(self._static[k], self.sets[k], stages=2)
        self.stages[k].scatter_ws(self.sets[k]["pillar_features_buf"], self.max_frames, self.sets[k]["spatial_features"])
        return self.sets[k]

    @staticmethod
    def release(out: Dict[str, torch.Tensor]) -> None:
        """Call on the consumer's stream after enqueuing its reads of ``out``."""
        e = torch.cuda.Event()
        e.record(torch.cuda.current_stream())
        out["released"] = e

    def drain(self) -> None:
        """Make the caller's current stream wait for everything submitted so far."""
        cur = torch.cuda.current_stream()
        for out in self.sets:
            if out.get("done") is not None:
                cur.wait_event(out["done"])


@device_guard
def generic_scatter(pillar_features: torch.Tensor, voxel_coords: torch.Tensor, nx: int, ny: int,
                    num_frames: Optional[int] = None) -> torch.Tensor:
    """PointPillarScatter for arbitrary (pillar_features, voxel_coords) (pointpillar_scatter.py:14-37)."""
    lib = _lib.load()
    _require_cuda(pillar_features, "pillar_features")
    _require_cuda(voxel_coords, "voxel_coords")
    pf = pillar_features.float().contiguous()
    vc = voxel_coords.to(torch.int32).contiguous()
    p = vc.shape[0]
    if num_frames is None:
        nf = torch.zeros(1, dtype=torch.int32, device=vc.device)
        _lib.check(lib.pcp_num_frames(_ptr(vc), p, _ptr(nf), _stream()), "pcp_num_frames")
        num_frames = int(nf.item())             # the reference syncs here too (pointpillar_scatter.py:17)
    c = pf.shape[1]
    cell_map = torch.empty((max(num_frames, 1) * ny * nx,), dtype=torch.int32, device=vc.device)
    canvas = torch.empty((num_frames, c, ny, nx), dtype=torch.float32, device=vc.device)
    rc = lib.pcp_bev_scatter(_ptr(pf), _ptr(vc), p, c, num_frames, nx, ny, _ptr(cell_map), _ptr(canvas), _stream())
    _lib.check(rc, "pcp_bev_scatter")
    return canvas
