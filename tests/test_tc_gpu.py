"""GPU: the tcgen05 / TMEM building block of the Cogen plant model (3xTF32 GEMM tile) vs float64."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,K", [(128, 8), (128, 32), (256, 32), (256, 64)])
def test_tc_gemm_tile(native_lib, N, K):
    from sustaingym_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda"); g.manual_seed(N + K)
    A = (torch.rand(128, K, generator=g, device="cuda") * 2 - 1) * 3
    Bm = (torch.rand(N, K, generator=g, device="cuda") * 2 - 1)
    D = torch.full((128, N), float("nan"), device="cuda")
    _lib.check(lib.sg_tc_gemm_selftest(A.data_ptr(), Bm.data_ptr(), D.data_ptr(), N, K,
                                       torch.cuda.current_stream().cuda_stream), "sg_tc_gemm_selftest")
    torch.cuda.synchronize()
    ref = A.double() @ Bm.double().T
    err = (D.double() - ref).abs().max().item()
    scale = (A.double().abs() @ Bm.double().abs().T).max().item()
    assert err <= 2e-6 * scale, (err, scale)       # 3xTF32: ~2^-21 relative to the magnitude sum
