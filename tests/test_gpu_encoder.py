"""Encoder-side kernels on a real B200: the tcgen05 (TF32) dense layer against a torch fp32 reference."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _gemm(a, bt, bias, relu):
    import torch
    from hashgan_b200 import _native

    lib = _native.lib()
    M, K = a.shape
    N = bt.shape[0]
    c = torch.full((M, N), float("nan"), dtype=torch.float32, device=a.device)
    stream = torch.cuda.current_stream().cuda_stream
    _native.check(lib.hg_gemm_tf32(a.data_ptr(), a.stride(0), bt.data_ptr(), bt.stride(0), bias.data_ptr() if bias is not None else None,
                                   c.data_ptr(), c.stride(0), M, N, K, 1 if relu else 0, stream))
    torch.cuda.synchronize()
    return c


@pytest.mark.parametrize("M,N,K,relu", [(128, 128, 32, False), (128, 128, 256, False), (256, 384, 512, True), (1280, 4096, 1024, True),
                                        (1280, 64, 4096, False), (100, 48, 4096, False), (333, 200, 96, True), (1280, 4096, 9216, True)])
def test_gemm_tf32_matches_fp32_reference(M, N, K, relu):
    import torch

    torch.manual_seed(M * 7 + N * 3 + K)
    dev = torch.device("cuda:0")
    a = torch.randn(M, K, device=dev)
    bt = torch.randn(N, K, device=dev) / K ** 0.5
    bias = torch.randn(N, device=dev)
    got = _gemm(a, bt, bias, relu)
    torch.backends.cuda.matmul.allow_tf32 = False
    want = a.double() @ bt.double().t() + bias.double()
    if relu:
        want = want.clamp_min(0)
    assert not torch.isnan(got).any()
    err = (got.double() - want).abs().max().item()
    scale = want.abs().max().item()
    # TF32 keeps 10 mantissa bits: |err| ~ 2^-11 * sqrt(K) * |a||b| per term; allow 2e-3 of the output scale
    assert err <= 2e-3 * max(scale, 1.0), (err, scale)
    # exact integers survive TF32: a stricter structural check (small integer operands are exact in tf32)
    ai = torch.randint(-3, 4, (M, K), device=dev).float()
    bi = torch.randint(-3, 4, (N, K), device=dev).float()
    goti = _gemm(ai, bi, None, False)
    wanti = (ai.double() @ bi.double().t()).float()
    assert torch.equal(goti, wanti)
