"""Shared test helpers: golden loading and oracle parameter plumbing (tests may import oracle/)."""
import glob
import os

import numpy as np
import torch

from oracle import pillar_oracle as po

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases(prefix="vfe_"):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    g["sd"] = {k[3:]: torch.from_numpy(g[k]) for k in z.files if k.startswith("sd/")}
    return g


def golden_cfg(g):
    """(VFEConfig, range fp32 array, voxel list of python floats, grid) with the reference's scalar types."""
    rng = np.asarray(g["point_cloud_range"], dtype=np.float32)
    vox = [float(v) for v in g["voxel_size"]]
    grid = g["grid_size"]
    cfg = po.VFEConfig(int(g["c_raw"]), vox, rng, grid, use_absolute_xyz=bool(g["use_abs"]),
                       with_distance=bool(g["with_distance"]), use_norm=bool(g["use_norm"]))
    return cfg, rng, vox, grid


def layers_from_state_dict(sd, use_norm=True):
    layers, i = [], 0
    while f"pfn_layers.{i}.linear.weight" in sd:
        p = f"pfn_layers.{i}."
        if use_norm:
            layers.append(po.PFNLayerParams(sd[p + "linear.weight"].float(), None, sd[p + "norm.weight"].float(),
                                            sd[p + "norm.bias"].float(), sd[p + "norm.running_mean"].float(),
                                            sd[p + "norm.running_var"].float()))
        else:
            layers.append(po.PFNLayerParams(sd[p + "linear.weight"].float(), sd[p + "linear.bias"].float()))
        i += 1
    return layers


def model_cfgs(*args, **kwargs):
    from pcp_b200.synthetic import model_cfgs as f
    return f(*args, **kwargs)


def assert_features_close(got, want, what, rtol=1e-5, atol_scale=1e-5):
    """fp32 tolerance of the path (BASELINE.json north_star): 1e-5 relative, with an absolute floor of
    1e-5 x the tensor's scale for values that come out of cancellation / ReLU near zero."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    scale = max(float(np.abs(want).max()) if want.size else 0.0, 1e-30)
    err = np.abs(got - want)
    tol = rtol * np.abs(want) + atol_scale * scale
    bad = err > tol
    assert not bad.any(), (f"{what}: {int(bad.sum())} / {bad.size} outside rtol={rtol} atol={atol_scale}*{scale:.3g}; "
                           f"max err {err.max():.3e} at {np.unravel_index(err.argmax(), err.shape)}")
