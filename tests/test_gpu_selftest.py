"""Tensor-core building blocks on the GPU: split-precision tcgen05 GEMM through the same producer /
loader / MMA / drain code the fused kernel uses, against an fp64 matmul."""
import pytest
import torch

from anerf_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fmt", [1, 0])
@pytest.mark.parametrize("N,K", [(256, 128), (256, 256), (256, 512), (128, 1024), (64, 128), (64, 384)])
def test_split_gemm_matches_fp64(fmt, N, K):
    g = torch.Generator(device="cpu").manual_seed(N * 1000 + K)
    A = torch.randn(256, K, generator=g).cuda()
    B = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    D = _lib.selftest_gemm(A, B, fmt)
    torch.cuda.synchronize()
    ref = (A.double() @ B.double().t())
    scale = ref.abs().max().item()
    for rep in range(2):   # two passes: TMEM regions 0 and 1, ring wrap-around
        err = (D[rep].double() - ref).abs().max().item() / scale
        # hi*hi + lo*hi + hi*lo keeps ~2^-16 (bf16) / ~2^-21 (fp16) of each product
        assert err < (3e-5 if fmt == 1 else 8e-6), (rep, err)   # (the self test does not pre-scale B, so fp16 lo parts of small weights go subnormal)
