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


def _ref(A, B):
    return (A.double() @ B.double().t()).float()


def test_tc_gemm_forward_dgrad_wgrad_forms():
    """The training path's tensor-core GEMM in its three uses, with ragged sizes (rows not a multiple of 256, K not
    a multiple of 128, N split into tiles, strided views), against fp64 matmul.  bf16 hi/lo split: 1e-4 of max."""
    torch.manual_seed(0)
    dev = torch.device("cuda")
    rows, P, W = 700, 432, 256
    XS = torch.randn(rows, P + W, device=dev)                    # [encoding | h] with leading dimension 688
    Wl = torch.randn(W, P, device=dev) / 20
    b = torch.randn(W, device=dev)
    # forward: H = relu(X W^T + b), X = XS[:, :P]
    H = torch.empty(rows, W, device=dev)
    _lib.selftest_tc_gemm(XS, (P + W, 1), rows, P, Wl, (P, 1), W, H, (W, 1), bias=b, relu=True)
    ref = torch.relu(_ref(XS[:, :P], Wl) + b)
    assert float((H - ref).abs().max() / ref.abs().max()) < 1e-4
    # dgrad into the skip layer's input: dX[:, P:] = (G W5[:, P:]) . (h > 0), accumulated onto existing values
    W5 = torch.randn(W, P + W, device=dev) / 20
    G = torch.randn(rows, W, device=dev) * 1e-4                  # gradient-sized magnitudes
    dXS = torch.randn(rows, P + W, device=dev) * 1e-4
    before = dXS.clone()
    _lib.selftest_tc_gemm(G, (W, 1), rows, W, W5[:, P:], (1, P + W), W, dXS[:, P:], (P + W, 1), mask=XS[:, P:], mask_ms=P + W, mode=1)
    ref = (before[:, P:] + _ref(G, W5[:, P:].t().contiguous())) * (XS[:, P:] > 0)
    assert float((dXS[:, P:] - ref).abs().max() / ref.abs().max()) < 1e-4
    assert torch.equal(dXS[:, :P], before[:, :P])
    # dgrad, wide output split into N tiles: dX[:, :P] = G W5[:, :P]
    _lib.selftest_tc_gemm(G, (W, 1), rows, W, W5, (1, P + W), P, dXS, (P + W, 1))
    ref = _ref(G, W5[:, :P].t().contiguous())
    assert float((dXS[:, :P] - ref).abs().max() / ref.abs().max()) < 1e-4
    # wgrad: dW5[n, k'] += sum_rows G[row, n] XS[row, k'] (A = XS^T, B = G^T, transposed atomic output, split over rows)
    rows2 = 1500
    XS2 = torch.randn(rows2, P + W, device=dev)
    G2 = torch.randn(rows2, W, device=dev) * 1e-4
    dW = torch.zeros(W, P + W, device=dev)
    _lib.selftest_tc_gemm(XS2, (1, P + W), P + W, rows2, G2, (1, W), W, dW, (1, P + W), mode=2, slice_chunks=8)
    ref = _ref(G2.t().contiguous(), XS2.t().contiguous())
    assert float((dW - ref).abs().max() / ref.abs().max()) < 1e-4
    # wgrad of the narrow views layer (N = 128) with a 920-wide input
    VIN = torch.randn(rows2, 920, device=dev)
    GHV = torch.randn(rows2, 128, device=dev) * 1e-4
    dWv = torch.zeros(128, 920, device=dev)
    _lib.selftest_tc_gemm(VIN, (1, 920), 920, rows2, GHV, (1, 128), 128, dWv, (1, 920), mode=2, slice_chunks=32)
    ref = _ref(GHV.t().contiguous(), VIN.t().contiguous())
    assert float((dWv - ref).abs().max() / ref.abs().max()) < 1e-4
