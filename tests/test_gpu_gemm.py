"""tcgen05 3xTF32 GEMM (csrc/hn_gemm.cu) against float64 matmul.  ``-m gpu``."""
import pytest
import torch

from hermnet_b200 import ops

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,K,N", [(128, 32, 64), (128, 128, 128), (1000, 128, 384), (4097, 256, 128), (300, 128, 512),
                                   (65, 384, 128), (20000, 128, 384), (777, 256, 768)])
def test_gemm_tf32x3_matches_fp64(M, K, N):
    g = torch.Generator().manual_seed(M + K + N)
    a = (torch.randn(M, K, generator=g) * torch.exp(torch.randn(M, 1, generator=g))).cuda()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    b = torch.randn(N, generator=g).cuda()
    hi, lo = ops.split_tf32(w)
    assert torch.equal(hi + lo, w)
    out = ops.gemm_tf32x3(a, hi, lo, b)
    ref = a.double() @ w.double().t() + b.double()
    scale = (a.double().abs() @ w.double().abs().t()) + b.double().abs()      # condition-aware bound
    err = ((out.double() - ref).abs() / scale).max().item()
    assert err < 2e-6, err
    fp32 = torch.nn.functional.linear(a, w, b)
    err32 = ((fp32.double() - ref).abs() / scale).max().item()
    assert err < 4 * err32 + 1e-6          # no worse than cuBLAS fp32 by more than a small factor
    out2 = ops.gemm_tf32x3(a, hi, lo, None)
    assert torch.allclose(out2 + b, out, rtol=1e-6, atol=1e-6)


def test_gemm_strided_input_and_determinism():
    g = torch.Generator().manual_seed(1)
    big = torch.randn(5000, 512, generator=g).cuda()
    w = torch.randn(384, 128, generator=g).cuda() / 11.3
    hi, lo = ops.split_tf32(w)
    a = big[:, 128:256]                                 # column slice: row pitch 512
    out = ops.gemm_tf32x3(a, hi, lo, None)
    ref = a.double() @ w.double().t()
    assert ((out.double() - ref).abs().max() / ref.abs().max()).item() < 1e-5
    assert torch.equal(out, ops.gemm_tf32x3(a, hi, lo, None))
