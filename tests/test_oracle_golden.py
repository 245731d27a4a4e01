"""CPU: the oracle restatement reproduces the committed outputs of the REFERENCE's own modules bit for bit."""
import numpy as np
import pytest
import torch

from oracle import pillar_oracle as po
from oracle import modar_oracle as mo
from tests.helpers import golden_cases, golden_cfg, layers_from_state_dict, load_golden


@pytest.mark.parametrize("name", golden_cases())
def test_vfe_and_scatter_bit_exact(name):
    g = load_golden(name)
    cfg, rng, vox, grid = golden_cfg(g)
    layers = layers_from_state_dict(g["sd"], bool(g["use_norm"]))
    out = po.front_end(torch.from_numpy(g["points"]), cfg, layers, unique_dim0=False)
    assert np.array_equal(out["voxel_coords"].numpy(), g["voxel_coords"])
    assert np.array_equal(out["unq_inv"].numpy().astype(np.int32), g["unq_inv"])
    assert np.array_equal(out["pillar_features"].numpy(), g["pillar_features"])        # same ATen ops: bit exact
    sf = out["spatial_features"]
    assert tuple(sf.shape) == tuple(g["canvas_shape"])
    occ = torch.nonzero((sf != 0).any(dim=1).flatten()).flatten().numpy().astype(np.int32)
    assert np.array_equal(occ, g["occupied"])
    assert np.array_equal(sf.double().sum(dim=(0, 2, 3)).numpy(), g["canvas_channel_sums"])
    if "spatial_features" in g:
        assert np.array_equal(sf.numpy(), g["spatial_features"])


@pytest.mark.parametrize("name", ["vfe_car_small", "vfe_edges_tiny"])
def test_unique_dim0_is_the_same_result(name):
    """torch.unique(dim=0), the call the reference makes (:108), equals the flat unique the tests use."""
    g = load_golden(name)
    cfg, *_ = golden_cfg(g)
    layers = layers_from_state_dict(g["sd"], bool(g["use_norm"]))
    a = po.dynamic_pillar_vfe(torch.from_numpy(g["points"]), cfg, layers, unique_dim0=True)
    assert np.array_equal(a["voxel_coords"].numpy(), g["voxel_coords"])
    assert np.array_equal(a["pillar_features"].numpy(), g["pillar_features"])


@pytest.mark.parametrize("name", golden_cases())
def test_integer_half_in_numpy(name):
    """Independent numpy restatement of quantise / cull / linearise / unique agrees with the reference output."""
    g = load_golden(name)
    cfg, *_ = golden_cfg(g)
    keep, keys, unq, inv, cnt = po.quantise_keys_numpy(g["points"], cfg)
    nxy, ny = int(g["grid_size"][0]) * int(g["grid_size"][1]), int(g["grid_size"][1])
    coords = np.stack([unq // nxy, np.zeros_like(unq), unq % ny, (unq % nxy) // ny], axis=1).astype(np.int32)
    assert np.array_equal(coords, g["voxel_coords"])
    assert np.array_equal(inv.astype(np.int32), g["unq_inv"])
    assert cnt.sum() == keep.sum() == g["unq_inv"].shape[0]


def test_edge_case_fixture_covers_what_it_claims():
    g = load_golden("vfe_edges_tiny")
    pts = g["points"]
    assert np.isnan(pts).any() and np.isinf(pts).any()
    assert (pts[:, 0] == 2).sum() >= 4 and g["canvas_shape"][0] == 2          # trailing frame is empty
    assert g["unq_inv"].shape[0] < pts.shape[0]                               # some rows culled
    cnt = np.bincount(g["unq_inv"])
    assert cnt.max() >= 40 and (cnt == 1).any()                               # dense pillar + one-point pillars


def test_modar_se3_matches_reference_apply_se3():
    z = np.load("tests/golden/modar_small.npz") if False else None
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "modar_small.npz"))
    for a in range(4):
        got = mo.apply_se3_boxes(z[f"a{a}/se3"], z[f"a{a}/modar"][:, :7])
        assert np.array_equal(got, z[f"a{a}/ref_apply_se3_boxes"])             # reference's own function, bit exact
        assert np.array_equal(mo.points_in_boxes(z[f"a{a}/foreground"][:, :3], z[f"a{a}/modar"][:, :7]),
                              z[f"a{a}/oracle_box_idx"])
        assert np.array_equal(mo.propagate_modar(z[f"a{a}/modar"], z[f"a{a}/foreground"], 2.0),
                              z[f"a{a}/oracle_propagated"])


def test_modar_membership_known_answers():
    """Hand-checkable cases of the box test (roiaware_pool3d_kernel.cu:23-36): z uses |dz| <= h/2 inclusive,
    xy uses a strict < with a 1e-5 margin, rotation by -heading, the first containing box wins."""
    boxes = np.array([[0, 0, 0, 4, 2, 2, 0.0], [0, 0, 0, 10, 10, 10, 0.0],
                      [20, 0, 0, 4, 2, 2, np.pi / 2]], dtype=np.float32)
    pts = np.array([[1.9, 0.9, 0.0],      # inside box 0 (and 1): first wins
                    [2.0, 0.0, 0.0],      # on the x face of box 0: |lx| = 2 < 2 + 1e-5 -> inside
                    [2.1, 0.0, 0.0],      # outside box 0, inside box 1
                    [0.0, 0.0, 1.0],      # z face of box 0: |dz| = 1 is not > 1 -> inside
                    [0.0, 0.0, 1.01],     # above box 0, inside box 1
                    [20.0, 1.9, 0.0],     # box 2 is rotated 90 deg: its long side lies along y
                    [21.9, 0.0, 0.0],     # outside box 2 (short side along x now)
                    [100, 100, 0]], dtype=np.float32)
    assert mo.points_in_boxes(pts, boxes).tolist() == [0, 0, 1, 0, 1, 2, -1, -1]


def test_modar_exchange_layout():
    rng = np.random.default_rng(0)
    ego = np.zeros((5, 13), dtype=np.float32)
    ego[:, :5] = rng.random((5, 5)); ego[:, -2] = [0, 3, 10, 2, 1]; ego[:, -1] = -1
    modar = np.array([[1, 2, -1, 4, 2, 1.5, 0.3, 0.9, 1]], dtype=np.float32)
    out = mo.modar_exchange(ego, [{"modar": modar, "foreground": None, "target_se3_agent": np.eye(4)}], 10.0, 2.0)
    assert out.shape == (6, 13) and np.array_equal(out[:5], ego)
    np.testing.assert_allclose(out[5], [1, 2, -1, 0, 0, 4, 2, 1.5, 0.3, 0.9, 1, 10, -1], rtol=1e-6)
