"""GPU: the tcgen05 building blocks (smem descriptors, instruction descriptor, TMEM load pattern, commit -> mbarrier,
CTA-pair mode) against a plain fp32 GEMM of the same fp16 inputs."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def run_probe(cg, K, N, variant=0, seed=0):
    import ceeus_b200

    lib = ceeus_b200._lib.load_probe()
    g = torch.Generator(device="cpu").manual_seed(seed)
    M = 128 * cg
    A = (torch.randn(M, K, generator=g) * 0.5).half().cuda()
    Bt = (torch.randn(N, K, generator=g) * 0.5).half().cuda()
    D = torch.full((M, N), float("nan"), device="cuda")
    st = C.c_int(0)
    rc = lib.ceeus_selftest_umma(cg, C.c_void_p(A.data_ptr()), C.c_void_p(Bt.data_ptr()), C.c_void_p(D.data_ptr()), K, N,
                                 variant, C.byref(st))
    want = A.float() @ Bt.float().t()
    return rc, st.value, D, want


@pytest.mark.parametrize("cg,K,N", [(1, 16, 32), (1, 48, 128), (1, 128, 128), (1, 160, 128), (1, 128, 16), (1, 128, 64),
                                     (2, 16, 32), (2, 48, 128), (2, 128, 128), (2, 160, 128), (2, 128, 16)])
def test_umma_tile_matches_gemm(cg, K, N):
    rc, st, D, want = run_probe(cg, K, N)
    assert rc == 0 and st == 0, (rc, st)
    err = (D - want).abs().max().item()
    assert err < 2e-3 * max(1.0, want.abs().max().item()), err


@pytest.mark.parametrize("cg,K,N", [(1, 128, 128), (1, 48, 64), (2, 128, 128), (2, 160, 128), (2, 48, 128), (2, 128, 16)])
def test_umma_tile_a_from_tmem(cg, K, N):
    """TS mode: the A operand is written to TMEM by the epilogue threads (tcgen05.st) instead of shared memory."""
    rc, st, D, want = run_probe(cg, K, N, variant=2)
    assert rc == 0 and st == 0, (rc, st)
    err = (D - want).abs().max().item()
    assert err < 2e-3 * max(1.0, want.abs().max().item()), err
