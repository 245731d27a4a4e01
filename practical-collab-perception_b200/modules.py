"""Drop-in pcdet modules: ``DynamicPillarVFE``, ``PFNLayerV2``, ``PointPillarScatter`` and
``DynamicMeanVFE``, ``DynamicPillarVFESimple2D``.

Same constructor signatures, config keys, ``batch_dict`` keys and ``state_dict`` names as the reference
(pcdet/models/backbones_3d/vfe/dynamic_pillar_vfe.py:14-147,
pcdet/models/backbones_2d/map_to_bev/pointpillar_scatter.py:5-37), so
``vfe.__all__['DynPillarVFE'] = pcp_b200.DynamicPillarVFE`` (see registry.py) and reference checkpoints load
unchanged.  The arithmetic runs in libpcp_b200.so; inference (eval mode) only - ``forward`` raises in
training mode rather than silently differing (training needs batch-statistics BatchNorm and autograd
through the segment reductions, which are outside this path).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .config import cfg_get
from .frontend import FrontEnd, GridSpec, _ptr, _stream, generic_scatter

CTX_KEY = "pcp_b200_ctx"   # private hand-off from the VFE to the scatter inside batch_dict


class PFNLayerV2(nn.Module):
    """Parameter container with the reference's names (``linear``, ``norm``); dynamic_pillar_vfe.py:14-33.
    The fused kernel executes the whole stack, so a layer is not callable on its own."""

    def __init__(self, in_channels, out_channels, use_norm=True, last_layer=False):
        super().__init__()
        self.last_vfe = last_layer
        self.use_norm = use_norm
        if not self.last_vfe:
            out_channels = out_channels // 2
        if self.use_norm:
            self.linear = nn.Linear(in_channels, out_channels, bias=False)
            self.norm = nn.BatchNorm1d(out_channels, eps=1e-3, momentum=0.01)
        else:
            self.linear = nn.Linear(in_channels, out_channels, bias=True)
        self.relu = nn.ReLU()

    def forward(self, inputs, unq_inv):
        raise NotImplementedError(
            "pcp_b200.PFNLayerV2 holds parameters only: DynamicPillarVFE.forward runs every PFN layer inside one "
            "fused sm_100a kernel (Linear+BN+ReLU+segment max); there is no per-layer torch path")


class VFETemplate(nn.Module):
    """pcdet/models/backbones_3d/vfe/vfe_template.py:4-22."""

    def __init__(self, model_cfg, **kwargs):
        super().__init__()
        self.model_cfg = model_cfg

    def get_output_feature_dim(self):
        raise NotImplementedError

    def forward(self, **kwargs):
        raise NotImplementedError


class DynamicPillarVFE(VFETemplate):
    def __init__(self, model_cfg, num_point_features, voxel_size, grid_size, point_cloud_range, **kwargs):
        super().__init__(model_cfg=model_cfg)
        if cfg_get(self.model_cfg, "NUM_RAW_POINT_FEATURES", None) is not None:      # :53-54
            num_point_features = self.model_cfg.NUM_RAW_POINT_FEATURES
        self.num_raw_point_features = num_point_features
        self.use_norm = self.model_cfg.USE_NORM
        self.with_distance = self.model_cfg.WITH_DISTANCE
        self.use_absolute_xyz = self.model_cfg.USE_ABSLOTE_XYZ
        num_point_features += 6 if self.use_absolute_xyz else 3
        if self.with_distance:
            num_point_features += 1

        self.num_filters = self.model_cfg.NUM_FILTERS
        assert len(self.num_filters) > 0
        num_filters = [num_point_features] + list(self.num_filters)
        pfn_layers = []
        for i in range(len(num_filters) - 1):                                        # :68-75
            pfn_layers.append(PFNLayerV2(num_filters[i], num_filters[i + 1], self.use_norm,
                                         last_layer=(i >= len(num_filters) - 2)))
        self.pfn_layers = nn.ModuleList(pfn_layers)

        self._grid = GridSpec(voxel_size, point_cloud_range, grid_size)
        self.voxel_x, self.voxel_y, self.voxel_z = self._grid.voxel_x, self._grid.voxel_y, self._grid.voxel_z
        self.x_offset, self.y_offset, self.z_offset = self._grid.x_offset, self._grid.y_offset, self._grid.z_offset
        self.scale_xy = int(grid_size[0]) * int(grid_size[1])                        # :84
        self.scale_y = int(grid_size[1])                                             # :85
        # the reference keeps these as CUDA tensors (:87-89); kept on the host here, kernels take scalars
        self.grid_size = torch.tensor(np.asarray(grid_size))
        self.voxel_size = torch.tensor(np.asarray(voxel_size, dtype=np.float32))
        self.point_cloud_range = torch.tensor(np.asarray(point_cloud_range, dtype=np.float32))

        self._fe: Optional[FrontEnd] = None
        self._param_key = None
        self._bufs = {}

    def get_output_feature_dim(self):
        return self.num_filters[-1]

    # -- kernel-side state -------------------------------------------------------------------------
    voxelize_method = None        # None (frontend.DEFAULT_VOXELIZE_METHOD) | "auto" | "radix" | "histogram": FrontEnd

    def _front_end(self) -> FrontEnd:
        if self._fe is None:
            self._fe = FrontEnd(self._grid, self.num_raw_point_features, self.use_absolute_xyz,
                                self.with_distance, list(self.num_filters), voxelize_method=self.voxelize_method)
        return self._fe

    def _sync_params(self, device) -> None:
        """Re-pack the PFN parameters when any of them changed (checkpoint load, .to(), in-place edit)."""
        tensors = []
        for layer in self.pfn_layers:
            tensors.append(layer.linear.weight)
            if layer.use_norm:
                tensors += [layer.norm.weight, layer.norm.bias, layer.norm.running_mean, layer.norm.running_var]
            else:
                tensors.append(layer.linear.bias)
        key = tuple((t.data_ptr(), t._version, str(t.device)) for t in tensors) + (str(device),)
        if key == self._param_key:
            return
        fe = self._front_end()

        def bn(layer):
            return (layer.norm.weight, layer.norm.bias, layer.norm.running_mean, layer.norm.running_var) \
                if layer.use_norm else None

        l0 = self.pfn_layers[0]
        l1 = self.pfn_layers[1] if len(self.pfn_layers) > 1 else None
        to = lambda t: None if t is None else t.detach().to(device)
        fe.pack_params(to(l0.linear.weight), None if bn(l0) is None else [to(t) for t in bn(l0)],
                       None if l1 is None else to(l1.linear.weight),
                       None if (l1 is None or bn(l1) is None) else [to(t) for t in bn(l1)],
                       lin_bias0=None if l0.use_norm else to(l0.linear.bias),
                       lin_bias1=None if (l1 is None or l1.use_norm) else to(l1.linear.bias),
                       eps=l0.norm.eps if l0.use_norm else 1e-3)
        self._param_key = key

    # -- forward ------------------------------------------------------------------------------------
    def forward(self, batch_dict, **kwargs):
        if self.training:
            raise RuntimeError("pcp_b200.DynamicPillarVFE is inference-only: call .eval() "
                               "(training-mode BatchNorm statistics and autograd are not implemented)")
        points = batch_dict["points"]
        if not points.is_cuda:
            raise RuntimeError("batch_dict['points'] must be on the GPU (load_data_to_gpu): no CPU path")
        with torch.cuda.device(points.device):      # the kernels launch on the current device's current stream
            return self._forward(batch_dict, points)

    def _forward(self, batch_dict, points):
        if points.dtype != torch.float32 or not points.is_contiguous():
            points = points.float().contiguous()
        if points.shape[1] < 1 + self.num_raw_point_features:
            raise RuntimeError(f"points has {points.shape[1]} columns, need 1 + {self.num_raw_point_features}")
        fe = self._front_end()
        self._sync_params(points.device)

        # dense key space needs the number of frames; collate_batch always provides it (dataset.py:320)
        max_frames = batch_dict.get("batch_size", None)
        if max_frames is None:
            max_frames = int(points[:, 0].max().item()) + 1 if points.shape[0] else 1
        max_frames = max(int(max_frames), 1)

        out = fe.voxelize(points, max_frames, self._bufs, want_point_pillar=True, host_counts=True)
        fe.pfn(points, out)
        # the one host sync of the module (32 bytes): it waits for the voxelize kernels only, the PFN keeps running while
        # the caller (PointPillarScatter, ...) enqueues what comes next
        counts = fe.read_counts(out)
        if counts[_lib.COUNT_BAD_FRAME] > 0:
            raise RuntimeError(f"{int(counts[_lib.COUNT_BAD_FRAME])} points carry a frame index outside "
                               f"[0, {max_frames}) (batch_dict['batch_size'] too small?)")
        p = int(counts[_lib.COUNT_PILLARS])
        features = out["pillar_features_buf"][:p]
        voxel_coords = out["voxel_coords_buf"][:p]
        batch_dict["voxel_features"] = batch_dict["pillar_features"] = features      # :145
        batch_dict["voxel_coords"] = voxel_coords                                    # :146
        batch_dict[CTX_KEY] = {
            "front_end": fe, "generation": fe.ws.generation, "voxel_coords": voxel_coords,
            "num_frames": int(counts[_lib.COUNT_FRAMES]), "num_pillars": p,
            "point_pillar": out["point_pillar"][:points.shape[0]],
        }
        # buffers are handed to the caller; fresh ones are allocated next call so results are not overwritten
        self._bufs = {}
        return batch_dict


class PointPillarScatter(nn.Module):
    def __init__(self, model_cfg, grid_size, **kwargs):
        super().__init__()
        self.model_cfg = model_cfg
        self.num_bev_features = self.model_cfg.NUM_BEV_FEATURES
        self.nx, self.ny, self.nz = (int(g) for g in grid_size)
        assert self.nz == 1

    def forward(self, batch_dict, **kwargs):
        pillar_features, coords = batch_dict["pillar_features"], batch_dict["voxel_coords"]
        if not pillar_features.is_cuda:
            raise RuntimeError("pillar_features must be on the GPU: pcp_b200 has no CPU path")
        if coords.device != pillar_features.device:
            raise RuntimeError(f"pillar_features on {pillar_features.device}, voxel_coords on {coords.device}")
        with torch.cuda.device(pillar_features.device):
            return self._forward(batch_dict, pillar_features, coords)

    def _forward(self, batch_dict, pillar_features, coords):
        if coords.shape[0] == 0:
            # the reference fails here too: coords[:, 0].max() of an empty tensor (pointpillar_scatter.py:17)
            raise RuntimeError("PointPillarScatter: no pillars (max() of an empty voxel_coords)")
        ctx = batch_dict.get(CTX_KEY)
        if (ctx is not None and ctx["voxel_coords"] is coords
                and ctx["front_end"].ws.generation == ctx["generation"]
                and pillar_features.shape == (ctx["num_pillars"], self.num_bev_features)
                and pillar_features.dtype == torch.float32 and pillar_features.is_contiguous()
                and (ctx["front_end"].grid.nx, ctx["front_end"].grid.ny) == (self.nx, self.ny)):
            # fast path: cell -> pillar map is still in the VFE's workspace; batch size = last frame that
            # owns a pillar + 1, which is what coords[:, 0].max() + 1 evaluates to (:17)
            canvas = ctx["front_end"].scatter_ws(pillar_features, ctx["num_frames"])
        else:
            canvas = generic_scatter(pillar_features, coords, self.nx, self.ny)
        batch_dict["spatial_features"] = canvas          # (B, C * nz, ny, nx)  :35-36
        return batch_dict


class DynamicPillarVFESimple2D(DynamicPillarVFE):
    """pcdet/models/backbones_3d/vfe/dynamic_pillar_vfe.py:150-245: the pillar encoder without cluster offsets.
    Per-point features are [f_center (3), points[:, 1:] (every column, or points[:, 4:] without absolute xyz),
    (distance)] (:210-227); outputs ``pillar_features`` and ``pillar_coords`` (b, y, x) (:233-244).

    Runs on the same fused kernel as DynamicPillarVFE: its feature vector [raw | f_cluster | f_center] is a superset,
    so the first layer's weight columns are permuted into that order and the f_cluster columns get zero weights."""

    def __init__(self, model_cfg, num_point_features, voxel_size, grid_size, point_cloud_range, **kwargs):
        VFETemplate.__init__(self, model_cfg=model_cfg)
        self.use_norm = self.model_cfg.USE_NORM
        self.with_distance = self.model_cfg.WITH_DISTANCE
        self.use_absolute_xyz = self.model_cfg.USE_ABSLOTE_XYZ
        self.num_raw_point_features = num_point_features          # every column after the frame index (:199,216)
        if self.use_absolute_xyz:
            num_point_features += 3                                  # :159-160
        if self.with_distance:
            num_point_features += 1
        self.num_filters = self.model_cfg.NUM_FILTERS
        assert len(self.num_filters) > 0
        num_filters = [num_point_features] + list(self.num_filters)
        self.pfn_layers = nn.ModuleList([
            PFNLayerV2(num_filters[i], num_filters[i + 1], self.use_norm, last_layer=(i >= len(num_filters) - 2))
            for i in range(len(num_filters) - 1)])
        self._grid = GridSpec(voxel_size, point_cloud_range, grid_size)
        self.voxel_x, self.voxel_y, self.voxel_z = self._grid.voxel_x, self._grid.voxel_y, self._grid.voxel_z
        self.x_offset, self.y_offset, self.z_offset = self._grid.x_offset, self._grid.y_offset, self._grid.z_offset
        self.scale_xy = int(grid_size[0]) * int(grid_size[1])
        self.scale_y = int(grid_size[1])
        self.grid_size = torch.tensor(np.asarray(grid_size[:2]))
        self.voxel_size = torch.tensor(np.asarray(voxel_size, dtype=np.float32))
        self.point_cloud_range = torch.tensor(np.asarray(point_cloud_range, dtype=np.float32))
        self._fe = None
        self._param_key = None
        self._bufs = {}

    def _sync_params(self, device) -> None:
        tensors = []
        for layer in self.pfn_layers:
            tensors.append(layer.linear.weight)
            if layer.use_norm:
                tensors += [layer.norm.weight, layer.norm.bias, layer.norm.running_mean, layer.norm.running_var]
            else:
                tensors.append(layer.linear.bias)
        key = tuple((t.data_ptr(), t._version, str(t.device)) for t in tensors) + (str(device),)
        if key == self._param_key:
            return
        fe = self._front_end()
        l0 = self.pfn_layers[0]
        l1 = self.pfn_layers[1] if len(self.pfn_layers) > 1 else None
        to = lambda t: None if t is None else t.detach().to(device)
        bn = lambda layer: [to(layer.norm.weight), to(layer.norm.bias), to(layer.norm.running_mean), to(layer.norm.running_var)] \
            if layer.use_norm else None
        # reference column order [f_center(3) | raw(n) | dist] -> kernel order [raw(n) | f_cluster(3) = 0 | f_center(3) | dist]
        w = to(l0.linear.weight)
        n_raw = self.num_raw_point_features if self.use_absolute_xyz else self.num_raw_point_features - 3
        cols = [w[:, 3:3 + n_raw], torch.zeros((w.shape[0], 3), dtype=w.dtype, device=w.device), w[:, 0:3]]
        if self.with_distance:
            cols.append(w[:, 3 + n_raw:4 + n_raw])
        w_perm = torch.cat(cols, dim=1).contiguous()
        fe.pack_params(w_perm, bn(l0), None if l1 is None else to(l1.linear.weight), None if l1 is None else bn(l1),
                       lin_bias0=None if l0.use_norm else to(l0.linear.bias),
                       lin_bias1=None if (l1 is None or l1.use_norm) else to(l1.linear.bias),
                       eps=l0.norm.eps if l0.use_norm else 1e-3)
        self._param_key = key

    def forward(self, batch_dict, **kwargs):
        points = batch_dict["points"]
        if points.shape[1] != 1 + self.num_raw_point_features:
            raise RuntimeError(f"points has {points.shape[1]} columns, DynamicPillarVFESimple2D was built for "
                               f"1 + {self.num_raw_point_features} (it consumes every column, :216)")
        had = {k: batch_dict.get(k) for k in ("voxel_features", "voxel_coords")}
        batch_dict = super().forward(batch_dict, **kwargs)
        coords = batch_dict["voxel_coords"]
        batch_dict["pillar_coords"] = coords[:, [0, 2, 3]].contiguous()            # (b, y, x)  :238-241
        for k, v in had.items():                                                     # the reference writes only pillar_* (:243-244)
            if v is None:
                batch_dict.pop(k, None)
            else:
                batch_dict[k] = v
        return batch_dict


class DynamicMeanVFE(VFETemplate):
    """pcdet/models/backbones_3d/vfe/dynamic_mean_vfe.py:14-79: dynamic 3-D voxelisation, per-voxel mean of the point
    features.  ``voxel_features`` (V, C), ``voxel_coords`` (V, 4) int32 rows (b, z, y, x) in torch.unique order."""

    def __init__(self, model_cfg, num_point_features, voxel_size, grid_size, point_cloud_range, **kwargs):
        super().__init__(model_cfg=model_cfg)
        if cfg_get(model_cfg, "NUM_POINT_FEATURES", None) is not None:             # :18-19
            num_point_features = model_cfg.NUM_POINT_FEATURES
        self.num_point_features = num_point_features
        self.num_raw_point_features = num_point_features
        self._grid = GridSpec(voxel_size, point_cloud_range, grid_size)
        self.voxel_x, self.voxel_y, self.voxel_z = self._grid.voxel_x, self._grid.voxel_y, self._grid.voxel_z
        self.x_offset, self.y_offset, self.z_offset = self._grid.x_offset, self._grid.y_offset, self._grid.z_offset
        self.scale_xyz = int(grid_size[0]) * int(grid_size[1]) * int(grid_size[2])
        self.scale_yz = int(grid_size[1]) * int(grid_size[2])
        self.scale_z = int(grid_size[2])
        self.grid_size = torch.tensor(np.asarray(grid_size))
        self.voxel_size = torch.tensor(np.asarray(voxel_size, dtype=np.float32))
        self.point_cloud_range = torch.tensor(np.asarray(point_cloud_range, dtype=np.float32))
        self._min_z = float(np.float32(point_cloud_range[2]))
        self._voxel_z = float(np.float32(voxel_size[2]))
        self._ws = None

    def get_output_feature_dim(self):
        return self.num_point_features

    @torch.no_grad()
    def forward(self, batch_dict, **kwargs):
        points = batch_dict["points"]
        if not points.is_cuda:
            raise RuntimeError("batch_dict['points'] must be on the GPU (load_data_to_gpu): no CPU path")
        with torch.cuda.device(points.device):
            return self._forward(batch_dict, points)

    def _forward(self, batch_dict, points):
        lib = _lib.load()
        if points.dtype != torch.float32 or not points.is_contiguous():
            points = points.float().contiguous()
        c = self.num_raw_point_features
        if points.shape[1] < 1 + c:
            raise RuntimeError(f"points has {points.shape[1]} columns, need 1 + {c}")
        dev = points.device
        n = points.shape[0]
        max_frames = batch_dict.get("batch_size", None)
        if max_frames is None:
            max_frames = int(points[:, 0].max().item()) + 1 if n else 1
        max_frames = max(int(max_frames), 1)
        g = self._grid
        nbytes = int(lib.pcp_workspace_bytes(n, max_frames, g.nx, g.ny))
        sbytes = int(lib.pcp_voxel3d_scratch_bytes(n, max_frames, g.nx, g.ny))
        if self._ws is None or self._ws[0].numel() < nbytes or self._ws[1].numel() < sbytes or self._ws[0].device != dev:
            self._ws = (torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, device=dev),
                        torch.empty(int(sbytes * 1.25) + 4096, dtype=torch.uint8, device=dev))
        ws, scratch = self._ws
        cap = max(1, min(n, max_frames * g.nx * g.ny * g.nz))
        vf = torch.empty((cap, c), dtype=torch.float32, device=dev)
        vc = torch.empty((cap, 4), dtype=torch.int32, device=dev)
        counts = torch.empty(_lib.PCP_COUNTS_LEN, dtype=torch.int32, device=dev)
        rc = lib.pcp_voxelize3d_mean(_ptr(points), points.stride(0), n, max_frames, C.byref(g.c), C.c_float(self._min_z),
                                     C.c_float(self._voxel_z), g.nz, c, _ptr(ws), ws.numel(), _ptr(scratch), scratch.numel(),
                                     _ptr(vf), _ptr(vc), None, cap, _ptr(counts), _stream())
        _lib.check(rc, "pcp_voxelize3d_mean")
        cnt = counts.cpu().numpy()                                                 # the one host sync (32 bytes)
        if cnt[_lib.COUNT_BAD_FRAME] > 0:
            raise RuntimeError(f"{int(cnt[_lib.COUNT_BAD_FRAME])} points carry a frame index outside [0, {max_frames})")
        v = int(cnt[_lib.COUNT_VOXELS])
        batch_dict["voxel_features"] = vf[:v]                                      # :77
        batch_dict["voxel_coords"] = vc[:v]                                        # :78
        return batch_dict
