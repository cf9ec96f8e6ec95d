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


def test_front_end_split_equals_the_composed_reference_calls(fpv, cuda_dev):
    """front_end_split == body_params_encapsulate_batch(convert_to_3D_rot(x)) (global_optimization.py:261-268): same
    values bit for bit, same gradient (the fused node only replaces the slice / cat bookkeeping)."""
    g = torch.Generator(device="cpu").manual_seed(5)
    x = torch.randn(37, 78, generator=g).to(cuda_dev)
    x[:, 3:9] = fpv.convert_to_6D_rot(torch.cat([x[:, :3], 0.4 * x[:, 3:6], x[:, 9:]], 1))[:, 3:9]
    w = {k: torch.randn(37, n, generator=g).to(cuda_dev)
         for k, n in zip(("transl", "global_orient", "betas", "body_pose_vp", "left_hand_pose", "right_hand_pose",
                          "camera_translation"), (3, 3, 10, 32, 12, 12, 3))}
    xa = x.clone().requires_grad_(True)
    xb = x.clone().requires_grad_(True)
    fused = fpv.prior.front_end_split(xa)
    ref = fpv.body_params_encapsulate_batch(fpv.convert_to_3D_rot(xb))
    assert list(fused) == list(ref)
    for k in ref:
        assert torch.equal(fused[k], ref[k]), k
    sum((fused[k] * w[k]).sum() for k in w if k != "betas").backward()     # one block without a gradient
    sum((ref[k] * w[k]).sum() for k in w if k != "betas").backward()
    assert torch.equal(xa.grad, xb.grad)
    with pytest.raises(RuntimeError):
        fpv.prior.front_end_split(x[:, :75])


def test_vposer_decode_does_not_depend_on_frames_per_cta(fpv, cuda_dev):
    """Long clips put 2 or 4 frames on one CTA (one weight fetch serves them all); a frame's result and gradient are
    bit-identical to the one-frame-per-CTA launch that short batches get."""
    dec = fpv.VPoserDecoderB200(fpv.make_vposer_weights(seed=7)).to(cuda_dev)
    g = torch.Generator().manual_seed(4)
    for T in (301, 1801 if torch.cuda.get_device_properties(cuda_dev).multi_processor_count <= 450 else 4001):
        z = torch.randn(T, 32, generator=g).to(cuda_dev)
        cot = torch.randn(T, 1, 21, 3, generator=g).to(cuda_dev)
        zb = z.clone().requires_grad_(True)
        big = dec.decode(zb, output_type="aa")
        (big * cot).sum().backward()
        for s in range(0, T, 100):                      # 100 frames: always one frame per CTA
            zs = z[s:s + 100].clone().requires_grad_(True)
            small = dec.decode(zs, output_type="aa")
            (small * cot[s:s + 100]).sum().backward()
            assert torch.equal(small, big[s:s + 100])
            assert torch.equal(zs.grad, zb.grad[s:s + 100])
