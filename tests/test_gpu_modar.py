"""GPU parity of the MoDAR exchange kernel: against the CPU oracle, against the reference's own
apply_se3_ outputs (tests/golden/modar_small.npz) and - for box membership - against the REFERENCE'S OWN
CUDA KERNEL compiled from its source for sm_100a (oracle/_ref/libroiaware_ref.so, built by oracle/Makefile)."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import modar_oracle as mo

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libroiaware_ref.so")


def run_exchange(ego, agents, t_detect=0.0, t_query=0.2, **kw):
    import pcp_b200
    out = pcp_b200.modar_exchange([a["modar"] for a in agents], [a.get("foreground") for a in agents],
                                  [a["target_se3_agent"] for a in agents], t_detect, t_query, ego.to(DEV), **kw)
    torch.cuda.synchronize()
    return out


def oracle_exchange(ego13, agents, max_sweep, scale):
    ags = [{"modar": a["modar"].numpy(), "foreground": None if a.get("foreground") is None else a["foreground"].numpy(),
            "target_se3_agent": a["target_se3_agent"]} for a in agents]
    return mo.modar_exchange(ego13.numpy(), ags, max_sweep, scale)


def compare_rows(got, want, n_ego):
    got, want = got.cpu().numpy(), np.asarray(want)
    assert got.shape == want.shape
    assert np.array_equal(got[:n_ego], want[:n_ego]), "ego rows must be copied verbatim"
    g, w = got[n_ego:], want[n_ego:]
    # exact columns: zeros, dims, score, label, sweep idx, instance idx
    for c in (3, 4, 5, 6, 7, 9, 10, 11, 12):
        assert np.array_equal(g[:, c], w[:, c]), f"column {c}"
    # xyz: fp32 sums + fp64 SE(3); heading: fp32 sin/cos/atan2 (libdevice vs numpy SIMD kernels, few ulp)
    np.testing.assert_allclose(g[:, :3], w[:, :3], rtol=1e-5, atol=1e-5)
    d = np.abs(g[:, 8] - w[:, 8])
    d = np.minimum(d, 2 * np.pi - d)
    assert d.max() < 2e-6, f"heading differs by {d.max()}"


@pytest.mark.parametrize("n_agents", [1, 5])
def test_config2_scene_against_oracle(n_agents):
    from pcp_b200 import synthetic as syn
    ego14, agents = syn.modar_scene(2, 0, n_agents=n_agents, n_ego_points=4096)
    ego13 = ego14[:, 1:].contiguous()
    max_sweep = float(ego13[:, -2].max())
    got, box_idx = run_exchange(ego13, agents, return_box_idx=True)
    want = oracle_exchange(ego13, agents, max_sweep, 2.0)
    compare_rows(got, want, ego13.shape[0])
    want_idx = np.concatenate([mo.points_in_boxes(a["foreground"][:, :3].numpy(), a["modar"][:, :7].numpy()) for a in agents])
    assert np.array_equal(box_idx.cpu().numpy(), want_idx)
    # 14-column layout (collate_batch frame column in front)
    got14 = run_exchange(ego14, agents, batch_idx=0.0)
    assert torch.equal(got14[:, 1:], got) and float(got14[:, 0].abs().max()) == 0.0


def test_exchange_now_and_missing_foreground_skip_propagation():
    from pcp_b200 import synthetic as syn
    ego14, agents = syn.modar_scene(2, 1, n_agents=3, n_ego_points=512)
    ego13 = ego14[:, 1:].contiguous()
    max_sweep = float(ego13[:, -2].max())
    now = run_exchange(ego13, agents, t_detect=3.0, t_query=3.0)
    compare_rows(now, oracle_exchange(ego13, agents, max_sweep, 0.0), ego13.shape[0])
    nofg = [dict(a, foreground=None) for a in agents]
    got = run_exchange(ego13, nofg)
    compare_rows(got, oracle_exchange(ego13, nofg, max_sweep, 2.0), ego13.shape[0])
    assert torch.equal(got, now)


def test_detections_dict_and_latency_scaling():
    from pcp_b200 import synthetic as syn
    import pcp_b200
    ego14, agents = syn.modar_scene(2, 2, n_agents=2, n_ego_points=256)
    ego13 = ego14[:, 1:].contiguous()
    dets = [{"pred_boxes": a["modar"][:, :7], "pred_scores": a["modar"][:, 7], "pred_labels": a["modar"][:, 8].long()} for a in agents]
    a = pcp_b200.modar_exchange(dets, [x["foreground"] for x in agents], [x["target_se3_agent"] for x in agents],
                                10.0, 10.2, ego13.to(DEV))
    b = run_exchange(ego13, agents)
    assert torch.equal(a, b)
    # 0.4 s of latency = two sample intervals -> scale 4
    c = run_exchange(ego13, agents, t_detect=1.0, t_query=1.4)
    compare_rows(c, oracle_exchange(ego13, agents, float(ego13[:, -2].max()), 4.0), ego13.shape[0])


def test_golden_apply_se3_from_the_reference():
    z = np.load(os.path.join(ROOT, "tests", "golden", "modar_small.npz"))
    for a in range(4):
        modar = torch.from_numpy(z[f"a{a}/modar"])
        ego = torch.zeros(1, 13)
        got = run_exchange(ego, [{"modar": modar, "foreground": None, "target_se3_agent": z[f"a{a}/se3"]}],
                           max_sweep_idx=10.0).cpu().numpy()[1:]
        ref = z[f"a{a}/ref_apply_se3_boxes"]                      # the reference's own apply_se3_
        np.testing.assert_allclose(got[:, :3], ref[:, :3], rtol=1e-6, atol=1e-6)
        d = np.abs(got[:, 8] - ref[:, 6]); d = np.minimum(d, 2 * np.pi - d)
        assert d.max() < 2e-6
        assert np.array_equal(got[:, 5:8], ref[:, 3:6])


@pytest.mark.skipif(not os.path.isfile(REF_SO), reason="oracle/_ref/libroiaware_ref.so not built (make -C oracle)")
def test_box_membership_bit_exact_against_reference_cuda_kernel():
    """points_in_boxes_kernel of the reference (roiaware_pool3d_kernel.cu:313-359), compiled unmodified, run on
    the same GPU: identical box index for every point, including knife-edge points on box faces."""
    from pcp_b200 import synthetic as syn
    ref = ctypes.CDLL(REF_SO)
    launcher = getattr(ref, "_Z24points_in_boxes_launcheriiiPKfS0_Pi")
    launcher.argtypes = [ctypes.c_int] * 3 + [ctypes.c_void_p] * 3
    launcher.restype = None
    g = torch.Generator().manual_seed(0)
    for seed in range(3):
        ag = syn.modar_agent(900 + seed, n_boxes=83, fg_per_box=(50, 120), stray_fraction=0.3)
        boxes, fg = ag["modar"].clone(), ag["foreground"].clone()
        # add points exactly on / next to faces and corners of random boxes
        k = torch.randint(0, boxes.shape[0], (4000,), generator=g)
        sgn = torch.randint(0, 2, (4000, 3), generator=g).float() * 2 - 1
        eps = (torch.randint(-2, 3, (4000, 3), generator=g).float()) * 1e-5
        local = sgn * boxes[k, 3:6] / 2 + eps
        c, s = torch.cos(boxes[k, 6]), torch.sin(boxes[k, 6])
        edge = torch.zeros(4000, 13)
        edge[:, 0] = local[:, 0] * c - local[:, 1] * s + boxes[k, 0]
        edge[:, 1] = local[:, 0] * s + local[:, 1] * c + boxes[k, 1]
        edge[:, 2] = local[:, 2] + boxes[k, 2]
        fg = torch.cat([fg, edge], 0).contiguous()
        b_dev, f_dev = boxes[:, :7].contiguous().to(DEV), fg[:, :3].contiguous().to(DEV)
        want = torch.full((fg.shape[0],), -1, dtype=torch.int32, device=DEV)
        torch.cuda.synchronize()
        launcher(1, boxes.shape[0], fg.shape[0], b_dev.data_ptr(), f_dev.data_ptr(), want.data_ptr())
        torch.cuda.synchronize()
        _, got = run_exchange(torch.zeros(1, 13), [{"modar": boxes, "foreground": fg, "target_se3_agent": np.eye(4)}],
                              return_box_idx=True, max_sweep_idx=0.0)
        assert torch.equal(got, want), f"{int((got != want).sum())} of {fg.shape[0]} box indices differ"
        assert int((want >= 0).sum()) > 1000


def test_exchange_messages_packed_on_the_gpu_feed_modar_and_nms(tmp_path):
    """the fixed-record wire format replaces the torch-pickled .pth hand-off.  Messages are packed on the
    GPU, unpacked as views, and give bit-identical MoDAR rows / NMS survivors to passing the tensors directly."""
    from pcp_b200 import synthetic as syn
    import pcp_b200
    ego14, agents = syn.modar_scene(2, 5, n_agents=3, n_ego_points=512)
    ego13 = ego14[:, 1:].contiguous()
    want = run_exchange(ego13, agents)
    msgs = []
    for i, a in enumerate(agents):
        buf = pcp_b200.pack_exchange(a["modar"].to(DEV), a["foreground"].to(DEV), agent_id=i, timestamp=7.0)
        assert buf.is_cuda
        if i == 0:                                           # one of them through a file, as the offline database does
            pcp_b200.write_exchange(tmp_path / "m.bin", buf)
            msgs.append(pcp_b200.read_exchange(tmp_path / "m.bin", device=DEV))
        else:
            msgs.append(pcp_b200.unpack_exchange(buf))
    assert [m.agent_id for m in msgs] == [0, 1, 2] and all(m.boxes.is_cuda for m in msgs)
    got = pcp_b200.modar_exchange(msgs, None, [a["target_se3_agent"] for a in agents], msgs[0].timestamp, 7.2, ego13.to(DEV))
    assert torch.equal(got, want)
    # late fusion: NMS over every agent's records (v2x_late_fusion.py:21-35)
    allb = torch.cat([m.boxes for m in msgs])
    direct = torch.cat([a["modar"] for a in agents]).to(DEV)
    cfg = pcp_b200.CfgDict(NMS_TYPE="nms_gpu", NMS_THRESH=0.2, NMS_PRE_MAXSIZE=1000, NMS_POST_MAXSIZE=100)
    s1, sc1 = pcp_b200.class_agnostic_nms(allb[:, 7], allb[:, :7], cfg, score_thresh=0.3)
    s2, sc2 = pcp_b200.class_agnostic_nms(direct[:, 7], direct[:, :7], cfg, score_thresh=0.3)
    assert torch.equal(s1, s2) and torch.equal(sc1, sc2) and s1.numel() > 0


def test_agents_without_boxes_or_foreground():
    """Ragged exchange: one agent reports no boxes, one reports boxes but no foreground points - the per-agent pointer entry
    (pcp_modar_agents) against the oracle."""
    import pcp_b200
    from pcp_b200 import synthetic as syn
    ego14, agents = syn.modar_scene(2, 7, n_agents=3, n_ego_points=2048)
    agents[1]["modar"] = agents[1]["modar"][:0]
    agents[1]["foreground"] = agents[1]["foreground"][:0]
    agents[2]["foreground"] = agents[2]["foreground"][:0]
    pts = pcp_b200.modar_exchange([a["modar"] for a in agents], [a["foreground"] for a in agents],
                                  [a["target_se3_agent"] for a in agents], 0.0, 0.2, ego14.to(DEV))
    want = mo.modar_exchange(ego14[:, 1:].numpy(), [{k: (v.numpy() if hasattr(v, "numpy") else v) for k, v in a.items()} for a in agents],
                             float(ego14[:, -2].max()), 2.0)
    got = pts.cpu().numpy()
    assert got.shape[0] == want.shape[0] == 2048 + agents[0]["modar"].shape[0] + agents[2]["modar"].shape[0]
    np.testing.assert_allclose(got[:, 1:], want, rtol=1e-5, atol=2e-5)


def test_single_message_call_forms():
    """The documented one-agent forms: a bare ExchangeMessage (a NamedTuple - it must not be iterated as four agents),
    a bare detections dict and a bare (M, 9) tensor, each with a bare (4, 4) pose."""
    from pcp_b200 import synthetic as syn
    import pcp_b200
    ego14, agents = syn.modar_scene(2, 7, n_agents=1, n_ego_points=300)
    ego13 = ego14[:, 1:].contiguous()
    a = agents[0]
    want = run_exchange(ego13, agents)
    msg = pcp_b200.unpack_exchange(pcp_b200.pack_exchange(a["modar"].to(DEV), a["foreground"].to(DEV), agent_id=3, timestamp=1.0))
    got = pcp_b200.modar_exchange(msg, None, a["target_se3_agent"], msg.timestamp, 1.2, ego13.to(DEV))
    assert torch.equal(got, want)
    got = pcp_b200.modar_exchange(msg, a["foreground"], a["target_se3_agent"], 1.0, 1.2, ego13.to(DEV))
    assert torch.equal(got, want)
    got = pcp_b200.modar_exchange(a["modar"], a["foreground"], a["target_se3_agent"], 1.0, 1.2, ego13.to(DEV))
    assert torch.equal(got, want)
    got = pcp_b200.modar_exchange(msg.detections, a["foreground"], a["target_se3_agent"], 1.0, 1.2, ego13.to(DEV))
    assert torch.equal(got, want)


def test_hunter_toolbox_refuses_to_cut_an_autograd_graph():
    import pcp_b200
    from pcp_b200 import hunter_toolbox as ht
    coord = torch.rand(64, 2, device=DEV) * 16
    feat = torch.rand(64, 4, device=DEV, requires_grad=True)
    bidx = torch.zeros(64, dtype=torch.long, device=DEV)
    with pytest.raises(RuntimeError, match="inference-only"):
        ht.bev_scatter(coord, bidx, feat, (16, 16), batch_size=1)
    with torch.no_grad():
        out = ht.bev_scatter(coord, bidx, feat, (16, 16), batch_size=1)
    assert out.shape == (1, 4, 16, 16)
    out = ht.bev_scatter(coord, bidx, feat.detach(), (16, 16), batch_size=1)
    assert not out.requires_grad


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_tensors_on_a_non_current_gpu_run_on_their_own_device():
    """Advisor finding: the library launches on the current device's stream.  With cuda:0 current, a model and a cloud
    on cuda:1 must give the same result as on cuda:0; tensors on two GPUs in one call must raise."""
    import pcp_b200
    from pcp_b200 import synthetic as syn
    from pcp_b200.frontend import FrontEnd, GridSpec
    rng = np.asarray(syn.V2X_RANGE, dtype=np.float32)
    gs = GridSpec(syn.V2X_VOXEL, rng, syn.grid_size_of(rng, syn.V2X_VOXEL))
    pts = syn.batch_of_frames(1, 20000, 1)
    outs = []
    torch.cuda.set_device(0)
    for dev in ("cuda:0", "cuda:1"):
        fe = FrontEnd(gs, 5)
        o = fe.voxelize(pts.to(dev), 1)
        torch.cuda.synchronize(dev)
        p = int(fe.read_counts(o)[0])
        outs.append(o["voxel_coords_buf"][:p].cpu())
    assert torch.equal(outs[0], outs[1]) and torch.cuda.current_device() == 0
    with pytest.raises(RuntimeError, match="different GPUs"):
        pcp_b200.boxes_iou_bev(torch.zeros(1, 7, device="cuda:0"), torch.zeros(1, 7, device="cuda:1"))


@pytest.mark.skipif(not os.path.isfile(REF_SO), reason="oracle/_ref/libroiaware_ref.so not built (make -C oracle)")
def test_propagation_against_the_reference_composition():
    """v2x_sim_dataset_ego.py:203-215 composed literally: the REFERENCE'S OWN points_in_boxes kernel (compiled unmodified)
    -> torch.unique(return_inverse) -> torch_scatter.scatter(reduce='mean') * 2 (pure-torch restatement of the absent
    third-party op) -> modar[unq, :3] += offset.  Both the oracle and the product must reproduce it: the product sums each
    box's flow in ascending point order, which is the CPU scatter's order, so the comparison is bit for bit."""
    from pcp_b200 import synthetic as syn
    from oracle import pillar_oracle as po
    ref = ctypes.CDLL(REF_SO)
    launcher = getattr(ref, "_Z24points_in_boxes_launcheriiiPKfS0_Pi")
    launcher.argtypes = [ctypes.c_int] * 3 + [ctypes.c_void_p] * 3
    launcher.restype = None
    for seed in range(4):
        ag = syn.modar_agent(1300 + seed, n_boxes=40 + 10 * seed, fg_per_box=(30, 200), stray_fraction=0.25)
        modar, foregr = ag["modar"].clone(), ag["foreground"].clone()
        b_dev, f_dev = modar[:, :7].contiguous().to(DEV), foregr[:, :3].contiguous().to(DEV)
        idx = torch.full((foregr.shape[0],), -1, dtype=torch.int32, device=DEV)
        torch.cuda.synchronize()
        launcher(1, modar.shape[0], foregr.shape[0], b_dev.data_ptr(), f_dev.data_ptr(), idx.data_ptr())
        torch.cuda.synchronize()
        # ---- the literal lines :203-215 ----
        box_idx_of_foregr = idx.cpu().long()
        mask_valid_foregr = box_idx_of_foregr > -1
        fg = foregr[mask_valid_foregr]
        box_idx_of_foregr = box_idx_of_foregr[mask_valid_foregr]
        unq_box_idx, inv_unq_box_idx = torch.unique(box_idx_of_foregr, return_inverse=True)
        boxes_offset = po.scatter_mean(fg[:, -3:], inv_unq_box_idx, unq_box_idx.shape[0]) * 2.
        want = modar.clone()
        want[unq_box_idx, :3] += boxes_offset
        # ---- oracle ----
        got_oracle = mo.propagate_modar(modar.numpy(), foregr.numpy(), 2.0)
        assert np.array_equal(got_oracle, want.numpy()), "oracle differs from the reference composition"
        # ---- product (identity pose: the fp64 SE(3) is exact, heading wrap leaves |yaw| < pi unchanged up to 1 ulp) ----
        rows = run_exchange(torch.zeros(1, 13), [{"modar": modar, "foreground": foregr, "target_se3_agent": np.eye(4)}],
                            max_sweep_idx=0.0).cpu()[1:]
        assert torch.equal(rows[:, :3], want[:, :3]), "product differs from the reference composition"
        assert int((unq_box_idx.shape[0])) > 10
