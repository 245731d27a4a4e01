"""Exchange wire format: host logic of pcp_b200.exchange on CPU tensors - the records are the
reference's own row layouts (center_head.py:413-417, v2x_sim_dataset_ego.py:196-200), so a round trip is bit-exact."""
import struct

import numpy as np
import pytest
import torch

from pcp_b200 import exchange as ex


def _msg(m=7, f=33, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(m, 9, generator=g), torch.randn(f, 13, generator=g)


@pytest.mark.parametrize("m,f", [(7, 33), (1, 0), (0, 0), (83, 5000)])
def test_round_trip_is_bit_exact(m, f):
    boxes, fg = _msg(m, f)
    buf = ex.pack_exchange(boxes, fg if f else None, agent_id=3, timestamp=12.4)
    assert buf.dtype == torch.uint8 and buf.numel() == ex.message_bytes(m, f) == 64 + 36 * m + 52 * f
    msg = ex.unpack_exchange(buf)
    assert msg.agent_id == 3 and msg.timestamp == 12.4
    assert torch.equal(msg.boxes, boxes) and msg.foreground.shape == (f, 13)
    if f:
        assert torch.equal(msg.foreground, fg)
    # views, not copies
    assert msg.boxes.untyped_storage().data_ptr() == buf.untyped_storage().data_ptr()
    d = msg.detections
    assert torch.equal(d["pred_boxes"], boxes[:, :7]) and torch.equal(d["pred_scores"], boxes[:, 7])
    assert d["pred_labels"].dtype == torch.int64


def test_detections_dict_input_and_preallocated_buffer():
    boxes, fg = _msg()
    det = {"pred_boxes": boxes[:, :7], "pred_scores": boxes[:, 7], "pred_labels": boxes[:, 8].round().long()}
    out = torch.zeros(4096, dtype=torch.uint8)
    buf = ex.pack_exchange(det, fg, out=out)
    assert buf.untyped_storage().data_ptr() == out.untyped_storage().data_ptr()
    msg = ex.unpack_exchange(buf)
    assert torch.equal(msg.boxes[:, :8], boxes[:, :8]) and torch.equal(msg.boxes[:, 8], boxes[:, 8].round())
    with pytest.raises(ValueError):
        ex.pack_exchange(det, fg, out=torch.zeros(100, dtype=torch.uint8))


def test_header_layout_is_the_documented_one():
    boxes, fg = _msg(5, 2)
    raw = ex.pack_exchange(boxes, fg, agent_id=-2, timestamp=0.2).numpy().tobytes()
    magic, version, agent, hbytes, m, f, bf, ff = struct.unpack_from("<IIiIIIII", raw, 0)
    assert raw[:4] == b"PCPX" and (magic, version, agent, hbytes, m, f, bf, ff) == (ex.MAGIC, 1, -2, 64, 5, 2, 9, 13)
    assert struct.unpack_from("<d", raw, 32)[0] == 0.2 and raw[40:64] == bytes(24)
    assert np.array_equal(np.frombuffer(raw, dtype="<f4", count=45, offset=64).reshape(5, 9), boxes.numpy())


def test_foreign_truncated_and_malformed_messages_are_rejected():
    boxes, fg = _msg()
    buf = ex.pack_exchange(boxes, fg)
    with pytest.raises(ValueError, match="truncated"):
        ex.unpack_exchange(buf[:-4].clone())
    bad = buf.clone()
    bad[0] = 0
    with pytest.raises(ValueError, match="magic"):
        ex.unpack_exchange(bad)
    newer = buf.clone()
    newer[4] = 9
    with pytest.raises(ValueError, match="unsupported"):
        ex.unpack_exchange(newer)
    with pytest.raises(ValueError):
        ex.unpack_exchange(buf[:10].clone())
    with pytest.raises(ValueError):
        ex.pack_exchange(torch.zeros(3, 8))
    with pytest.raises(ValueError):
        ex.pack_exchange(boxes, torch.zeros(3, 12))


def test_file_round_trip(tmp_path):
    boxes, fg = _msg(11, 70, seed=4)
    p = tmp_path / "tok_id2_exchange.bin"
    ex.write_exchange(p, ex.pack_exchange(boxes, fg, agent_id=2, timestamp=3.2))
    assert p.stat().st_size == ex.message_bytes(11, 70)
    msg = ex.read_exchange(p)
    assert msg.agent_id == 2 and msg.timestamp == 3.2 and torch.equal(msg.boxes, boxes) and torch.equal(msg.foreground, fg)


def test_round_trip_property():
    """hypothesis: any record counts, agent ids and timestamps survive pack -> bytes -> unpack bit for bit."""
    from hypothesis import HealthCheck, given, settings, strategies as st

    @settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
    @given(m=st.integers(0, 90), f=st.integers(0, 300), agent=st.integers(-2**31, 2**31 - 1),
           ts=st.floats(allow_nan=False, allow_infinity=False, width=64), seed=st.integers(0, 1000))
    def check(m, f, agent, ts, seed):
        boxes, fg = _msg(m, f, seed)
        raw = ex.pack_exchange(boxes, fg if f else None, agent_id=agent, timestamp=ts).numpy().tobytes()
        msg = ex.unpack_exchange(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
        assert (msg.agent_id, msg.timestamp) == (agent, ts)
        assert torch.equal(msg.boxes, boxes) and (f == 0 or torch.equal(msg.foreground, fg))

    check()
