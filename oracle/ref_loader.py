"""Import the REFERENCE's own modules for the hot path from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box: nothing that runs
there may call this.  It is used by ``oracle/gen_golden.py`` (fixture generation) and by
``tests/test_oracle_vs_reference.py`` (skipped when the reference tree is absent).

The ``pcdet`` package cannot be imported as a package (its ``__init__`` chain pulls compiled CUDA
ops, easydict, nuscenes ...), so the three files on the path are loaded BY FILE PATH under a synthetic
parent package:

* ``pcdet/models/backbones_3d/vfe/dynamic_pillar_vfe.py`` (+ ``vfe_template.py``)
* ``pcdet/models/backbones_2d/map_to_bev/pointpillar_scatter.py``
* ``pcdet/datasets/nuscenes/nuscenes_temporal_utils.py`` (``apply_se3_``)
* ``pcdet/models/backbones_3d/vfe/dynamic_mean_vfe.py``, ``pcdet/models/bev_layers/hunter_toolbox.py``

``torch_scatter`` is absent and unpinned by the reference; a pure-torch stand-in that follows its
published semantics is placed in ``sys.modules`` (same restatement as oracle/pillar_oracle.py).
The three hard ``.cuda()`` calls in the reference constructor (dynamic_pillar_vfe.py:87-89) are
neutralised while constructing on a CPU-only machine.  No reference source is copied.
"""
from __future__ import annotations

import contextlib
import importlib.util
import os
import sys
import types

import torch

# /root/reference in the build container; on the GPU box the git-ignored copies build() leaves under oracle/_ref/py
# (same relative paths), which travel with the snapshot like the compiled oracle/_ref/*.so
_LOCAL_COPY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "py")
REFERENCE_FILES = (
    "pcdet/models/backbones_3d/vfe/vfe_template.py",
    "pcdet/models/backbones_3d/vfe/dynamic_pillar_vfe.py",
    "pcdet/models/backbones_3d/vfe/dynamic_mean_vfe.py",
    "pcdet/models/backbones_2d/map_to_bev/pointpillar_scatter.py",
    "pcdet/models/bev_layers/hunter_toolbox.py",
    "pcdet/datasets/nuscenes/nuscenes_temporal_utils.py",
)


def _default_root() -> str:
    env = os.environ.get("PCP_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isfile(os.path.join("/root/reference", REFERENCE_FILES[1])):
        return "/root/reference"
    return _LOCAL_COPY


REFERENCE_ROOT = _default_root()


def stage_reference_files(src_root: str = "/root/reference") -> int:
    """Copy the reference's own .py files of the path (unmodified) into oracle/_ref/py so that `bench.py --impl reference`
    and tests/test_oracle_vs_reference.py can run them where /root/reference does not exist.  The directory is
    git-ignored: the sources never enter this repository's history.  Returns the number of files copied."""
    import shutil
    n = 0
    for rel in REFERENCE_FILES:
        src = os.path.join(src_root, rel)
        if not os.path.isfile(src):
            continue
        dst = os.path.join(_LOCAL_COPY, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        n += 1
    return n


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "pcdet/models/backbones_3d/vfe/dynamic_pillar_vfe.py"))


def _install_torch_scatter_shim() -> None:
    if "torch_scatter" in sys.modules:
        return
    from . import pillar_oracle as po

    shim = types.ModuleType("torch_scatter")
    shim.__doc__ = "pure-torch stand-in for rusty1s/pytorch_scatter (absent in this image)"

    def scatter_mean(src, index, dim=0, out=None, dim_size=None):
        assert dim == 0 and out is None
        return po.scatter_mean(src, index, dim_size)

    def scatter_max(src, index, dim=0, out=None, dim_size=None):
        assert dim == 0 and out is None
        val = po.scatter_max(src, index, dim_size)
        return val, None  # the reference only takes [0] (dynamic_pillar_vfe.py:40)

    def scatter(src, index, dim=0, out=None, dim_size=None, reduce="sum"):
        assert dim == 0 and out is None
        if reduce == "mean":
            return po.scatter_mean(src, index, dim_size)
        if reduce == "max":
            return po.scatter_max(src, index, dim_size)
        raise NotImplementedError(reduce)

    shim.scatter_mean, shim.scatter_max, shim.scatter = scatter_mean, scatter_max, scatter
    sys.modules["torch_scatter"] = shim


def _load(name: str, relpath: str, package: str | None = None):
    path = os.path.join(REFERENCE_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    if package is not None:
        mod.__package__ = package
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@contextlib.contextmanager
def _cuda_is_identity(force: bool = False):
    """dynamic_pillar_vfe.py:87-89 call .cuda() unconditionally; make that a no-op on CPU-only hosts, and - force - on a
    GPU box when the reference is to run on the host cores (the CPU baseline of bench.py)."""
    if torch.cuda.is_available() and not force:
        yield
        return
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig


_cache = {}


def load_reference_modules():
    """Returns a namespace with DynamicPillarVFE, PFNLayerV2, PointPillarScatter, apply_se3_."""
    if "ns" in _cache:
        return _cache["ns"]
    if not reference_available():
        raise FileNotFoundError(f"reference tree not found under {REFERENCE_ROOT}")
    _install_torch_scatter_shim()
    pkg = "_pcp_ref_vfe"
    parent = types.ModuleType(pkg)
    parent.__path__ = []  # mark as package so that "from .vfe_template import" resolves
    sys.modules[pkg] = parent
    _load(f"{pkg}.vfe_template", "pcdet/models/backbones_3d/vfe/vfe_template.py", pkg)
    vfe = _load(f"{pkg}.dynamic_pillar_vfe", "pcdet/models/backbones_3d/vfe/dynamic_pillar_vfe.py", pkg)
    mean_vfe = _load(f"{pkg}.dynamic_mean_vfe", "pcdet/models/backbones_3d/vfe/dynamic_mean_vfe.py", pkg)
    scat = _load("_pcp_ref_pointpillar_scatter", "pcdet/models/backbones_2d/map_to_bev/pointpillar_scatter.py")
    # the two hot functions of the HunterJr head (imports torch, torch_scatter (shim), einops only)
    hunter = _load("_pcp_ref_hunter_toolbox", "pcdet/models/bev_layers/hunter_toolbox.py")

    # nuscenes_temporal_utils imports nuscenes + pyquaternion at module scope (only type names are used
    # by apply_se3_): satisfy them with empty stand-ins.
    for missing, attrs in (("nuscenes", ("NuScenes",)), ("pyquaternion", ("Quaternion",))):
        if missing not in sys.modules:
            try:
                __import__(missing)
            except Exception:
                m = types.ModuleType(missing)
                for a in attrs:
                    setattr(m, a, type(a, (), {}))
                sys.modules[missing] = m
    se3 = _load("_pcp_ref_temporal_utils", "pcdet/datasets/nuscenes/nuscenes_temporal_utils.py")

    ns = types.SimpleNamespace(
        DynamicPillarVFE=vfe.DynamicPillarVFE, PFNLayerV2=vfe.PFNLayerV2,
        PointPillarScatter=scat.PointPillarScatter, apply_se3_=se3.apply_se3_,
        DynamicMeanVFE=mean_vfe.DynamicMeanVFE, DynamicPillarVFESimple2D=vfe.DynamicPillarVFESimple2D,
        bev_scatter=hunter.bev_scatter, interpolate_points_feat_from_bev_img=hunter.interpolate_points_feat_from_bev_img,
        cuda_is_identity=_cuda_is_identity)
    _cache["ns"] = ns
    return ns


class Cfg(dict):
    """Minimal EasyDict look-alike (attribute access + .get), what the reference modules read."""
    __getattr__ = dict.__getitem__


def build_reference_front_end(num_raw_point_features, voxel_size, point_cloud_range, grid_size,
                              num_filters=(64, 64), use_norm=True, with_distance=False, use_absolute_xyz=True):
    """Constructs the reference's DynamicPillarVFE + PointPillarScatter in eval mode on the CPU (also on a GPU box)."""
    ns = load_reference_modules()
    vfe_cfg = Cfg(NUM_RAW_POINT_FEATURES=num_raw_point_features, USE_NORM=use_norm, WITH_DISTANCE=with_distance,
                  USE_ABSLOTE_XYZ=use_absolute_xyz, NUM_FILTERS=list(num_filters))
    with ns.cuda_is_identity(force=True):
        vfe = ns.DynamicPillarVFE(model_cfg=vfe_cfg, num_point_features=num_raw_point_features,
                                  voxel_size=voxel_size, grid_size=grid_size,
                                  point_cloud_range=point_cloud_range)
    scatter = ns.PointPillarScatter(model_cfg=Cfg(NUM_BEV_FEATURES=num_filters[-1]), grid_size=grid_size)
    return vfe.eval(), scatter.eval()
