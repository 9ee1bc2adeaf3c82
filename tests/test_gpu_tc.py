"""tcgen05 / TMEM building blocks (csrc/tc05.cuh) pinned by a plain bf16 GEMM through the C ABI."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K", [(128, 64, 16), (128, 16, 16), (256, 64, 64), (300, 128, 208), (128, 256, 32), (1000, 64, 448)])
def test_tc_gemm_selftest(M, N, K):
    from carla_imitation_learning_b200 import _lib
    assert torch.cuda.is_available()
    dev = torch.device("cuda", 0)
    gen = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A = torch.randn((M, K), generator=gen).to(torch.bfloat16)
    B = torch.randn((N, K), generator=gen).to(torch.bfloat16)
    Ad, Bd = A.to(dev), B.to(dev)
    D = torch.full((M, N), float("nan"), dtype=torch.float32, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(_lib.lib().bc_tc_gemm_selftest(Ad.data_ptr(), Bd.data_ptr(), D.data_ptr(), M, N, K, err.data_ptr(),
                                             torch.cuda.current_stream().cuda_stream), "bc_tc_gemm_selftest")
    torch.cuda.synchronize()
    assert int(err[0]) == 0, "mbarrier wait timed out inside the kernel"
    ref = A.double() @ B.double().t()
    got = D.cpu().double()
    assert torch.isfinite(got).all()
    # bf16 products are exact in f32; only the f32 accumulation order differs
    assert float((got - ref).abs().max()) <= 1e-5 * float(ref.abs().max()) * max(1, K // 64)
