"""GPU parity: loss-algebra kernels vs the torch restatement of the reference expressions (oracle/residuals_oracle.py),
values and autograd gradients, tolerance 1e-5 relative (north_star)."""
import pytest
import torch

from oracle import residuals_oracle as ro

pytestmark = pytest.mark.gpu


def _pair(x, dev):
    return x.clone().double().requires_grad_(True), x.clone().to(dev).requires_grad_(True)


@pytest.mark.parametrize("T,F", [(3, 1), (5, 78), (30, 69), (12, 3 * 10475), (300, 78)])
def test_second_diff(fpv, cuda_dev, T, F):
    x = torch.randn(T, F, generator=torch.Generator().manual_seed(T * F))
    x[1] = x[0]                                                        # exact zeros: sign(0) = 0 path
    xo, xg = _pair(x, cuda_dev)
    lo, lg = ro.second_diff_l1(xo), fpv.second_diff_l1(xg)
    assert lg.item() == pytest.approx(lo.item(), rel=1e-5)
    (lo * 3.0).backward()
    (lg * 3.0).backward()
    torch.testing.assert_close(xg.grad.cpu().double(), xo.grad, rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("weighted", [False, True])
def test_first_diff(fpv, cuda_dev, weighted):
    T, K = 40, 500
    g = torch.Generator().manual_seed(5)
    x = torch.randn(T, K, 3, generator=g)
    w = (torch.rand(T, generator=g) > 0.4).float() * torch.rand(T, generator=g) if weighted else None
    xo, xg = _pair(x, cuda_dev)
    if weighted:
        lo = ro.weighted_first_diff_l1(xo, w.double())
        lg = fpv.first_diff_l1(xg, w.to(cuda_dev))
    else:
        lo, lg = ro.first_diff_l1(xo), fpv.first_diff_l1(xg)
    assert lg.item() == pytest.approx(lo.item(), rel=1e-5)
    lo.backward()
    lg.backward()
    torch.testing.assert_close(xg.grad.cpu().double(), xo.grad, rtol=1e-5, atol=1e-10)


def test_robust_contact(fpv, cuda_dev):
    d = torch.rand(30, 777, generator=torch.Generator().manual_seed(1)) ** 3
    d[0, :5] = 0.0
    do, dg = _pair(d, cuda_dev)
    lo, lg = ro.contact_robust_loss(do, 0.1), fpv.contact_robust_loss(dg, 0.1)
    assert lg.item() == pytest.approx(lo.item(), rel=1e-5)
    lo.backward()
    lg.backward()
    torch.testing.assert_close(dg.grad.cpu().double(), do.grad, rtol=2e-5, atol=1e-12)


def test_verts_transform(fpv, cuda_dev):
    g = torch.Generator().manual_seed(2)
    T, P = 7, 10475
    v = torch.randn(T, P, 3, generator=g)
    M = torch.randn(T, 4, 4, generator=g)
    gout = torch.randn(T, P, 3, generator=g)
    vo, vg = _pair(v, cuda_dev)
    Mo, Mg = _pair(M, cuda_dev)
    yo, yg = ro.verts_transform(vo, Mo), fpv.verts_transform(vg, Mg)
    torch.testing.assert_close(yg.cpu().double(), yo, rtol=1e-5, atol=1e-5)
    (yo * gout.double()).sum().backward()
    (yg * gout.to(cuda_dev)).sum().backward()
    torch.testing.assert_close(vg.grad.cpu().double(), vo.grad, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(Mg.grad.cpu().double(), Mo.grad, rtol=1e-5, atol=1e-5 * float(Mo.grad.abs().max()))
    with pytest.raises(RuntimeError):
        fpv.verts_transform(vg, Mg[:3])


def test_errors(fpv, cuda_dev):
    with pytest.raises(RuntimeError):
        fpv.second_diff_l1(torch.zeros(2, 5, device=cuda_dev))
    with pytest.raises(RuntimeError):
        fpv.first_diff_l1(torch.zeros(1, 5, device=cuda_dev))
    with pytest.raises(RuntimeError):
        fpv.contact_robust_loss(torch.zeros(0, device=cuda_dev))
