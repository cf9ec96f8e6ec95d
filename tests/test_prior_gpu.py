"""GPU parity of the parameter front-end and the DCT prior (csrc/prior.cu) through the C ABI: against goldens made by
the real reference code where it exists (6D decoder wiring, cal_dctloss) and the float64 oracle elsewhere.
Tolerance: 1e-5 relative (north_star), written per assertion."""
import importlib
import os

import numpy as np
import pytest
import torch

from oracle import prior_oracle as po

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def test_convert_rows_match_reference_goldens(fpv, cuda_dev):
    d = np.load(os.path.join(G, "prior_codec.npz"))
    rows78 = torch.tensor(d["rows78"], dtype=torch.float32, device=cuda_dev)
    rows75 = fpv.convert_to_3D_rot(rows78)
    want = po.convert_to_3D_rot(torch.tensor(d["rows78"]).float().double())
    np.testing.assert_allclose(rows75.cpu().numpy(), want.numpy(), rtol=1e-5, atol=2e-6)
    back = fpv.convert_to_6D_rot(rows75)
    want_back = po.convert_to_6D_rot(want)
    np.testing.assert_allclose(back.cpu().numpy(), want_back.numpy(), rtol=1e-5, atol=2e-6)
    # untouched columns pass through bit-exactly
    assert torch.equal(rows75[:, :3], rows78[:, :3]) and torch.equal(rows75[:, 6:], rows78[:, 9:])


def test_rot6d_to_aa_all_branches_and_gradient(fpv, cuda_dev):
    g = torch.Generator().manual_seed(11)
    x = torch.randn(6000, 6, generator=g)
    x[0] = torch.tensor([1., 0., 0., 1., 0., 0.])
    x[1] = torch.tensor([-1., 0., 0., -1., 0., 0.2])
    xo = x.double().requires_grad_(True)
    ao = po.rot6d_to_aa(xo)
    w = torch.randn(6000, 3, generator=g)
    (ao * w.double()).sum().backward()
    xg = x.to(cuda_dev).requires_grad_(True)
    ag = fpv.rot6d_to_aa(xg)
    (ag * w.to(cuda_dev)).sum().backward()
    # away from the branch boundaries of the quaternion selection / the pi wrap the map is smooth: compare there
    m = po.rot6d_decode(x.double()).transpose(1, 2)
    margin = torch.minimum(torch.minimum((m[:, 2, 2] - 1e-6).abs(), (m[:, 0, 0] - m[:, 1, 1]).abs()), (m[:, 0, 0] + m[:, 1, 1]).abs())
    ok = (margin > 1e-3) & (ao.detach().norm(dim=1) < 3.1)
    assert ok.sum() > 5000
    np.testing.assert_allclose(ag.detach().cpu().numpy()[ok], ao.detach().numpy()[ok], rtol=1e-5, atol=3e-6)
    go, gg = xo.grad.numpy()[ok], xg.grad.cpu().numpy()[ok]
    np.testing.assert_allclose(gg, go, rtol=2e-4, atol=2e-4 * np.abs(go).max())
    # on the whole set (branch flips allowed) the ROTATION is the same
    Rg = po.aa2matrot(ag.detach().cpu().double())
    Ro = po.rot6d_decode(x.double())
    assert float((Rg - Ro).abs().max()) < 2e-5


def test_vposer_decode_forward_backward(fpv, cuda_dev):
    w = fpv.make_vposer_weights(seed=7)
    dec = fpv.VPoserDecoderB200(w).to(cuda_dev)
    g = torch.Generator().manual_seed(3)
    T = 37
    z = torch.randn(T, 32, generator=g)
    wd = {k: v.double() for k, v in w.items()}
    zo = z.double().requires_grad_(True)
    ao = po.vposer_decode_aa(wd, zo)
    cot = torch.randn(T, 1, 21, 3, generator=g)
    (ao * cot.double()).sum().backward()
    zg = z.to(cuda_dev).requires_grad_(True)
    ag = dec.decode(zg, output_type="aa")
    assert ag.shape == (T, 1, 21, 3)
    (ag * cot.to(cuda_dev)).sum().backward()
    np.testing.assert_allclose(ag.detach().cpu().numpy(), ao.detach().numpy(), rtol=1e-5, atol=1e-5)
    go = zo.grad.numpy()
    np.testing.assert_allclose(zg.grad.cpu().numpy(), go, rtol=1e-4, atol=1e-5 * np.abs(go).max())
    assert ag.view(T, -1).shape == (T, 63)                           # what the reference does with it (:271)
    with pytest.raises(RuntimeError):
        dec.decode(z, output_type="aa")                              # CPU tensor: no fallback


def test_dct_loss_matches_reference_golden(fpv, cuda_dev):
    d = np.load(os.path.join(G, "prior_dct.npz"))
    joints = torch.tensor(d["joints"], dtype=torch.float32, device=cuda_dev, requires_grad=True)
    c = torch.tensor(d["c_dct"], dtype=torch.float32, device=cuda_dev, requires_grad=True)
    basis = torch.tensor(d["basis"], dtype=torch.float32, device=cuda_dev)
    loss = fpv.cal_dctloss(joints, basis, c)
    loss.backward()
    assert abs(float(loss.detach()) - float(d["loss"])) <= 1e-5 * abs(float(d["loss"]))
    np.testing.assert_allclose(joints.grad.cpu().numpy(), d["g_joints"], rtol=1e-4, atol=1e-5 * np.abs(d["g_joints"]).max())
    np.testing.assert_allclose(c.grad.cpu().numpy(), d["g_c"], rtol=1e-4, atol=1e-5 * np.abs(d["g_c"]).max())
    # repeatable bit for bit (fixed-order reductions)
    loss2 = fpv.cal_dctloss(joints.detach(), basis, c.detach())
    assert torch.equal(loss2, loss.detach())
    # the synthetic basis helper reproduces the golden's basis
    np.testing.assert_allclose(fpv.dct_basis(60, 5).numpy(), d["basis"], atol=1e-7)
    with pytest.raises(RuntimeError):
        fpv.cal_dctloss(joints[:100], basis, c)
