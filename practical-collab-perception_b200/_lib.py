"""ctypes binding of libpcp_b200.so (include/pcp_b200.h).  No fallback: a missing library is an error."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpcp_b200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

PCP_COUNTS_LEN = 8
# radix_or_auto / binned_or_auto: the named method where it applies, else "auto" (see FrontEnd.voxelize)
VOXELIZE_METHODS = {"auto": 0, "histogram": 1, "radix": 2, "radix_or_auto": 2, "binned": 3, "binned_or_auto": 3}
COUNT_PILLARS, COUNT_KEPT, COUNT_FRAMES, COUNT_BAD_FRAME, COUNT_MAX_PER_PILLAR, COUNT_VOXELS = 0, 1, 2, 3, 4, 5


class PcpGrid(C.Structure):
    _fields_ = [("range_min_x", C.c_float), ("range_min_y", C.c_float),
                ("voxel_x", C.c_float), ("voxel_y", C.c_float),
                ("x_offset", C.c_float), ("y_offset", C.c_float), ("z_offset", C.c_float),
                ("nx", C.c_int32), ("ny", C.c_int32)]


class PcpPfnDesc(C.Structure):
    _fields_ = [("c_raw", C.c_int32), ("use_absolute_xyz", C.c_int32), ("with_distance", C.c_int32),
                ("num_layers", C.c_int32), ("hidden", C.c_int32), ("c_out", C.c_int32)]


_P = C.c_void_p
_SIGNATURES = {
    "pcp_abi_version": (C.c_int, []),
    "pcp_last_error_string": (C.c_char_p, []),
    "pcp_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32, C.c_int32, C.c_int32]),
    "pcp_pfn_param_floats": (C.c_size_t, [C.POINTER(PcpPfnDesc)]),
    "pcp_pack_pfn_params": (C.c_int, [C.POINTER(PcpPfnDesc)] + [_P] * 12 + [C.c_float, _P, _P]),
    "pcp_voxelize": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int32, C.POINTER(PcpGrid), _P, C.c_size_t,
                               _P, _P, _P, C.c_int64, _P, _P]),
    "pcp_voxelize_method": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int32, C.POINTER(PcpGrid), _P, C.c_size_t,
                                      _P, _P, _P, C.c_int64, _P, C.c_int32, _P]),
    "pcp_pfn": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int32, C.POINTER(PcpGrid), C.POINTER(PcpPfnDesc), _P,
                          _P, C.c_size_t, _P, _P, C.c_int64, _P]),
    "pcp_pfn_stages": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int32, C.POINTER(PcpGrid), C.POINTER(PcpPfnDesc), _P,
                                 _P, C.c_size_t, _P, _P, C.c_int64, C.c_int32, _P]),
    "pcp_segment_reduce": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                     _P, C.c_size_t, _P, C.c_int64, _P]),
    "pcp_bev_scatter_ws": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.POINTER(PcpGrid),
                                     _P, C.c_size_t, _P, _P]),
    "pcp_bev_scatter": (C.c_int, [_P, _P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P]),
    "pcp_num_frames": (C.c_int, [_P, C.c_int64, _P, _P]),
    "pcp_selftest_umma": (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, _P]),
    "pcp_selftest_umma_ts": (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, _P]),
    "pcp_selftest_umma_cycles": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "pcp_max_index_i64": (C.c_int, [_P, C.c_int64, _P, _P]),
    "pcp_bev_scatter_mean": (C.c_int, [_P, C.c_int64, _P, _P, C.c_int64, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                       _P, C.c_size_t, _P, _P, _P, _P]),
    "pcp_bev_interpolate": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, C.c_int64, C.c_int64,
                                      C.c_float, C.c_float, C.c_float, C.c_float, _P, _P, _P, _P]),
    "pcp_voxel3d_scratch_bytes": (C.c_size_t, [C.c_int64, C.c_int32, C.c_int32, C.c_int32]),
    "pcp_voxelize3d_mean": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int32, C.POINTER(PcpGrid), C.c_float, C.c_float, C.c_int32,
                                      C.c_int32, _P, C.c_size_t, _P, C.c_size_t, _P, _P, _P, C.c_int64, _P, _P]),
    "pcp_fuse_scratch_bytes": (C.c_size_t, [C.c_int64]),
    "pcp_fuse_agent_points": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int64, _P, _P, C.c_int32, _P, C.c_int32, C.c_float, _P,
                                        _P, C.c_int64, _P, _P]),
    "pcp_fuse_agent_clouds": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int64, _P, _P, C.c_int32, _P, C.c_int32, C.c_float, _P,
                                        _P, C.c_int64, _P, _P]),
    "pcp_select_scratch_bytes": (C.c_size_t, [C.c_int64, C.c_int32]),
    "pcp_select_foreground": (C.c_int, [_P, C.c_int64, C.c_int32, _P, C.c_int64, _P, C.c_int64, C.c_int64, C.c_int32, C.c_float,
                                        _P, _P, C.c_int64, _P, _P]),
    "pcp_unpack_points": (C.c_int, [_P, C.c_int32, _P, C.c_int64, _P, C.c_int32, C.c_int32, _P, C.c_int64, _P]),
    "pcp_boxes_iou_bev": (C.c_int, [_P, C.c_int64, _P, C.c_int64, _P, _P]),
    "pcp_nms_scratch_bytes": (C.c_size_t, [C.c_int64]),
    "pcp_nms_bev": (C.c_int, [_P, C.c_int64, _P, C.c_int64, C.c_int32, C.c_float, C.c_float, C.c_int32, C.c_int32, _P, C.c_size_t,
                              _P, _P, _P]),
    "pcp_nms_normal": (C.c_int, [_P, C.c_int64, _P, C.c_int64, C.c_int32, C.c_float, C.c_float, C.c_int32, C.c_int32, _P, C.c_size_t,
                                 _P, _P, _P]),
    "pcp_modar": (C.c_int, [_P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_int32, C.c_float,
                            _P, C.c_int64, _P, _P]),
    "pcp_modar_agents": (C.c_int, [_P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_float, _P, C.c_int32, C.c_float,
                                   _P, C.c_int64, _P, _P]),
    "pcp_column_max": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int32, _P, _P]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lock = threading.Lock()
_lib = None


def build(verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into libpcp_b200.so (nvcc cross-compiles without a GPU)."""
    proc = subprocess.run(["make", "-C", CSRC_DIR, "-j8"], capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        print(proc.stdout[-4000:])
        print(proc.stderr[-4000:])
    if proc.returncode != 0:
        raise RuntimeError("building libpcp_b200.so failed (see output above)")
    return LIB_PATH


def load() -> C.CDLL:
    """Returns the loaded library; raises if it has not been built (there is no CPU fallback)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: the CUDA extension is required (no CPU/torch fallback). "
                "Build it with `python -c 'import __graft_entry__ as g; g.build()'` or `make -C "
                f"{CSRC_DIR}`.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so is stale
            fn.restype, fn.argtypes = res, args
        if lib.pcp_abi_version() != 1:
            raise RuntimeError(f"libpcp_b200.so ABI {lib.pcp_abi_version()} != 1; rebuild")
        _lib = lib
        return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().pcp_last_error_string().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (rc={rc}): {msg}")
