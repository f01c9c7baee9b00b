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


def _ssilu(z):
    return z * torch.sigmoid(z) / 0.6


@pytest.mark.parametrize("M,K,N", [(1000, 128, 512), (4097, 256, 128), (333, 128, 192), (20000, 384, 128), (129, 64, 64)])
def test_gemm_activation_epilogues(M, K, N):
    """mode 1 (ScaledSiLU + stored pre-activation) and mode 2 (times ScaledSiLU'(aux)) of hn_gemm_tf32x3_ex, written into
    column blocks of wider buffers (row-pitched outputs / aux), both tile widths (N % 128 == 0 and N % 64 == 0)."""
    g = torch.Generator().manual_seed(M + 3 * K + N)
    a = torch.randn(M, K, generator=g).cuda()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    b = torch.randn(N, generator=g).cuda()
    hi, lo = ops.split_tf32(w)
    pre_ref = a.double() @ w.double().t() + b.double()
    wide = torch.full((M, N + 64), 7.0, device="cuda")
    wide2 = torch.full((M, 2 * N), -3.0, device="cuda")
    out = ops.gemm_tf32x3_ex(a, hi, lo, b, out=wide[:, 64:], mode=1, out2=wide2[:, :N])
    assert out.data_ptr() == wide[:, 64:].data_ptr()
    assert ((wide2[:, :N].double() - pre_ref).abs().max() / pre_ref.abs().max()).item() < 1e-5
    assert ((wide[:, 64:].double() - _ssilu(pre_ref)).abs().max() / pre_ref.abs().max()).item() < 1e-5
    assert torch.all(wide[:, :64] == 7.0) and torch.all(wide2[:, N:] == -3.0)          # neighbours untouched
    only = ops.gemm_tf32x3_ex(a, hi, lo, b, mode=1)                                    # without the pre-activation output
    assert torch.equal(only, wide[:, 64:])
    # mode 2: (a . w^T) * ScaledSiLU'(aux)
    aux_wide = torch.randn(M, N + 128, generator=g).cuda()
    aux = aux_wide[:, 128:]
    z = aux.double().requires_grad_(True)
    (dz,) = torch.autograd.grad(_ssilu(z).sum(), z)
    ref2 = (a.double() @ w.double().t()) * dz
    out2 = ops.gemm_tf32x3_ex(a, hi, lo, None, mode=2, aux=aux)
    assert ((out2.double() - ref2).abs().max() / ref2.abs().max()).item() < 1e-5
    assert torch.equal(out2, ops.gemm_tf32x3_ex(a, hi, lo, None, mode=2, aux=aux))     # deterministic
