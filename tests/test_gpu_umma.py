"""GPU: the hand-written tcgen05 path (operand panels -> tcgen05.mma kind::tf32 x3 -> TMEM -> tcgen05.ld)
against an fp64 matmul.  Pins the descriptor / layout plumbing the PFN kernel relies on."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k,n", [(8, 32), (16, 32), (24, 32), (32, 64), (64, 64), (64, 32)])
def test_3xtf32_gemm_matches_fp64(k, n):
    from pcp_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(k * 100 + n)
    a = (torch.randn(128, k, generator=g) * 3).cuda()
    b = torch.randn(n, k, generator=g).cuda()
    c = torch.full((128, n), float("nan"), device="cuda")
    rc = lib.pcp_selftest_umma(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), k, n, C.c_void_p(c.data_ptr()),
                               C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "pcp_selftest_umma")
    torch.cuda.synchronize()
    want = a.double() @ b.double().t()
    scale = (a.double().abs() @ b.double().abs().t())           # sum |a_k b_k|: the natural error scale
    err = ((c.double() - want).abs() / scale).max().item()
    assert torch.isfinite(c).all()
    assert err < 4e-6, f"3xTF32 relative error {err:.3e} (single-pass TF32 would be ~5e-4)"


def test_structured_operands_hit_the_right_rows_and_columns():
    """A = one-hot rows, B = distinct integers: any row / column / K-panel mix-up changes the answer exactly."""
    from pcp_b200 import _lib
    lib = _lib.load()
    k, n = 64, 64
    a = torch.zeros(128, k)
    a[torch.arange(128), torch.arange(128) % k] = 1.0
    a[:, 0] += torch.arange(128).float() * 0.5
    b = (torch.arange(n * k).reshape(n, k) % 97).float()
    c = torch.empty(128, n, device="cuda")
    a_d, b_d = a.cuda(), b.cuda()
    rc = lib.pcp_selftest_umma(C.c_void_p(a_d.data_ptr()), C.c_void_p(b_d.data_ptr()), k, n,
                               C.c_void_p(c.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "pcp_selftest_umma")
    torch.cuda.synchronize()
    assert torch.equal(c.cpu(), a @ b.t())          # small integers and halves: exact in every format involved


@pytest.mark.parametrize("k,n", [(8, 32), (16, 64), (32, 64), (32, 32)])
def test_3xtf32_gemm_with_a_in_tensor_memory(k, n):
    """Layer 1 of the PFN reads its A operand (the layer-0 activations) from TMEM: tcgen05.st + TS-form MMA."""
    from pcp_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(k * 1000 + n)
    a = (torch.randn(128, k, generator=g) * 3).cuda()
    b = torch.randn(n, k, generator=g).cuda()
    c = torch.full((128, n), float("nan"), device="cuda")
    rc = lib.pcp_selftest_umma_ts(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), k, n, C.c_void_p(c.data_ptr()),
                                  C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "pcp_selftest_umma_ts")
    torch.cuda.synchronize()
    want = a.double() @ b.double().t()
    scale = (a.double().abs() @ b.double().abs().t())
    err = ((c.double() - want).abs() / scale).max().item()
    assert torch.isfinite(c).all()
    assert err < 4e-6, f"3xTF32 (A in TMEM) relative error {err:.3e}"


def test_tmem_a_operand_rows_and_columns_exact():
    from pcp_b200 import _lib
    lib = _lib.load()
    k, n = 32, 64
    a = torch.zeros(128, k)
    a[torch.arange(128), torch.arange(128) % k] = 1.0
    a[:, 0] += torch.arange(128).float() * 0.5
    b = (torch.arange(n * k).reshape(n, k) % 97).float()
    c = torch.empty(128, n, device="cuda")
    a_d, b_d = a.cuda(), b.cuda()
    rc = lib.pcp_selftest_umma_ts(C.c_void_p(a_d.data_ptr()), C.c_void_p(b_d.data_ptr()), k, n,
                                  C.c_void_p(c.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "pcp_selftest_umma_ts")
    torch.cuda.synchronize()
    assert torch.equal(c.cpu(), a @ b.t())
