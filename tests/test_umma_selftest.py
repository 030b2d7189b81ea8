"""tcgen05 building blocks (naruto_b200/csrc/umma.cuh) against an fp64 matmul: descriptors, operand layouts, TMEM read-back."""
import pytest
import torch

from naruto_b200 import _lib as L

pytestmark = pytest.mark.gpu


def _run(mode, a, b, k, n, passes):
    lib = L.load()
    d = torch.full((128, n), float('nan'), device='cuda')
    L.check(lib.nrt_selftest_umma(mode, L.ptr(a), L.ptr(b), k, n, passes, L.ptr(d), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return d


@pytest.mark.parametrize('k,n', [(8, 16), (32, 16), (80, 32), (64, 32), (32, 32)])
@pytest.mark.parametrize('passes', [1, 3])
def test_kmajor_gemm(k, n, passes):
    g = torch.Generator().manual_seed(k * 100 + n)
    a = (torch.rand(128, k, generator=g) * 2 - 1).cuda()
    b = (torch.rand(n, k, generator=g) * 2 - 1).cuda()
    d = _run(0, a, b, k, n, passes)
    ref = (a.double() @ b.double().T)
    err = (d.double() - ref).abs().max().item()
    tol = 2e-6 * k if passes == 3 else 2e-3 * k ** 0.5
    assert err < tol, (err, tol)


@pytest.mark.parametrize('k,n', [(8, 16), (32, 16), (80, 32), (64, 32), (16, 32)])
@pytest.mark.parametrize('passes', [1, 3])
def test_tmem_a_gemm(k, n, passes):
    """A operand staged in tensor memory (tcgen05.st), B K-major in shared memory."""
    g = torch.Generator().manual_seed(k * 100 + n + 7)
    a = (torch.rand(128, k, generator=g) * 2 - 1).cuda()
    b = (torch.rand(n, k, generator=g) * 2 - 1).cuda()
    d = _run(2, a, b, k, n, passes)
    ref = (a.double() @ b.double().T)
    err = (d.double() - ref).abs().max().item()
    tol = 2e-6 * k if passes == 3 else 2e-3 * k ** 0.5
    assert err < tol, (err, tol)


@pytest.mark.parametrize('ma,n', [(128, 160), (88, 160), (32, 80), (16, 32), (8, 32)])
def test_weight_gradient_gemm(ma, n):
    """D[ma, n] = Y[128 pts, ma]^T X[128 pts, n] through transposed K-major operands (single-pass TF32, RN operands)."""
    g = torch.Generator().manual_seed(n + ma)
    y = (torch.rand(128, ma, generator=g) * 2 - 1).cuda()
    x = (torch.rand(128, n, generator=g) * 2 - 1).cuda()
    d = _run(1, y, x, ma, n, 1)[:ma]
    ref = y.double().T @ x.double()
    err = (d.double() - ref).abs().max().item()
    assert err < 2e-2, err
