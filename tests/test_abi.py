"""CPU: libpcp_b200.so loads, exports every symbol include/pcp_b200.h declares, validates arguments without
touching a GPU, and the product package never reaches into oracle/."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pcp_b200.h")


@pytest.fixture(scope="module")
def lib():
    from pcp_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        _lib.build()
    return _lib.load()


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pcp_[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_agree(lib):
    from pcp_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 12
    assert sorted(_lib.EXPORTED_SYMBOLS) == names
    for n in names:
        assert getattr(lib, n) is not None, n


def test_every_entry_point_cites_the_reference():
    src = open(HEADER).read()
    for needle in ("dynamic_pillar_vfe.py:98-108", "dynamic_pillar_vfe.py:110-129", "pointpillar_scatter.py:14-37",
                   "v2x_sim_dataset_ego.py:203-232", "roiaware_pool3d_kernel.cu", "nuscenes_temporal_utils.py:66-70"):
        assert needle in src, needle


def test_struct_layouts_match_the_header():
    from pcp_b200._lib import PcpGrid, PcpPfnDesc
    assert C.sizeof(PcpGrid) == 7 * 4 + 2 * 4 and C.sizeof(PcpPfnDesc) == 6 * 4


def test_workspace_bytes_is_monotone_and_aligned(lib):
    a = lib.pcp_workspace_bytes(32768, 1, 512, 512)
    b = lib.pcp_workspace_bytes(300000, 1, 512, 512)
    c = lib.pcp_workspace_bytes(300000, 8, 512, 512)
    assert 0 < a < b < c and a % 256 == 0 and c % 256 == 0
    assert lib.pcp_workspace_bytes(-1, 1, 512, 512) == 0 and lib.pcp_workspace_bytes(10, 0, 512, 512) == 0
    # cells + 4 int32 per point dominate
    assert c >= 8 * 512 * 512 * 4 + 300000 * 12


def test_argument_validation_returns_codes_not_crashes(lib):
    from pcp_b200._lib import PcpGrid, PcpPfnDesc
    g = PcpGrid(-51.2, -51.2, 0.2, 0.2, -51.1, -51.1, -4.0, 512, 512)
    rc = lib.pcp_voxelize(None, 8, 10, 1, C.byref(g), None, 0, None, None, None, 10, None, None)
    assert rc == -1 and b"null" in lib.pcp_last_error_string()
    d = PcpPfnDesc(5, 1, 0, 3, 32, 64)
    # hi/lo operand panels (K padded to 16) of W0, W1[:, :32], W1[:, 32:], their sum (one-point pillars) + folded BN
    # + fp32 W1[:, 32:] (long pillars)
    assert lib.pcp_pfn_param_floats(C.byref(PcpPfnDesc(5, 1, 0, 2, 32, 64))) == 2 * 16 * 32 + 6 * 32 * 64 + 64 + 128 + 64 * 32
    assert lib.pcp_pfn_param_floats(C.byref(PcpPfnDesc(5, 1, 0, 1, 0, 64))) == 2 * 16 * 64 + 128
    rc = lib.pcp_pack_pfn_params(C.byref(d), *([None] * 12), C.c_float(1e-3), None, None)
    assert rc == -3 and b"num_layers" in lib.pcp_last_error_string()
    assert lib.pcp_modar(None, None, None, None, None, 0, 0, 0, 2.0, 10.0, 0, 0.0, None, 13, None, None) == 0
    assert lib.pcp_modar(None, None, None, None, None, 2, 1, 1, 2.0, 10.0, 0, 0.0, None, 13, None, None) == -1
    assert lib.pcp_segment_reduce(None, 4, 4, 7, 10, 1, 8, 8, None, 0, None, 10, None) == -1


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "practical-collab-perception_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "/root/reference" not in txt, f


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from pcp_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU/torch fallback"):
        _lib.load()


def test_pfn_issue_warp_stays_inside_its_register_budget():
    """pfn_slot_kernel lowers the MMA-issue warp's register budget with setmaxnreg.dec 56; ptxas does not enforce
    the budget on the code that follows, so check the SASS: no register above R55 between the USETMAXREG and the
    first instruction of another role (cp.async / tensor-memory loads and stores)."""
    import shutil
    import subprocess
    from pcp_b200 import _lib
    obj = os.path.join(os.path.dirname(_lib.LIB_PATH), "csrc", "build", "pfn_tc.o")
    if not (os.path.isfile(obj) and shutil.which("cuobjdump")):
        pytest.skip("no object file / cuobjdump")
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    seen = 0
    for fn in sass.split("Function : ")[1:]:
        if "pfn_slot_kernel" not in fn.split("\n")[0]:
            continue
        lines = fn.split("\n")
        start = [i for i, l in enumerate(lines) if "USETMAXREG" in l]
        assert len(start) == 1, "expected exactly one setmaxnreg in pfn_slot_kernel"
        assert "0x38" in lines[start[0]], lines[start[0]]          # 56 registers
        end = start[0]
        while end < len(lines) and not re.search(r"LDGSTS|STTM|LDTM", lines[end]):
            end += 1
        regs = [int(x) for l in lines[start[0]:end] for x in re.findall(r"\bR(\d+)\b", l)]
        assert max(regs) < 56, f"issue warp uses R{max(regs)} after lowering its budget to 56"
        seen += 1
    assert seen >= 3
