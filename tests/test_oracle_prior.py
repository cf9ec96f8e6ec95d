"""CPU: the prior/front-end oracle against goldens produced by the REAL reference code
(tests/golden/make_golden_prior.py: cvae.ContinousRotReprDecoder.decode, global_optimization.convert_to_*D_rot,
FittingOP.cal_dctloss).  torchgeometry itself is absent: that part of the oracle is a restatement (parity unpinned,
see oracle/prior_oracle.py) and is checked here through invariants only."""
import os

import numpy as np
import torch

from oracle import prior_oracle as po

G = os.path.join(os.path.dirname(__file__), "golden")


def test_rot6d_decoder_matches_reference_class():
    d = np.load(os.path.join(G, "prior_codec.npz"))
    R = po.rot6d_decode(torch.tensor(d["x6"]))
    assert np.array_equal(R.numpy(), d["R"])            # same torch ops in the same order: bit-identical
    eye = torch.eye(3, dtype=torch.float64)
    assert torch.allclose(R.transpose(1, 2) @ R, eye.expand_as(R), atol=1e-12)
    assert torch.allclose(torch.linalg.det(R), torch.ones(R.shape[0], dtype=torch.float64), atol=1e-12)


def test_row_conversions_match_reference_wiring():
    d = np.load(os.path.join(G, "prior_codec.npz"))
    rows75 = po.convert_to_3D_rot(torch.tensor(d["rows78"]))
    assert np.array_equal(rows75.numpy(), d["rows75"])
    assert np.array_equal(po.convert_to_6D_rot(rows75).numpy(), d["rows78_back"])


def test_angle_axis_round_trip_and_rotation_consistency():
    g = torch.Generator().manual_seed(5)
    aa = torch.randn(500, 3, generator=g, dtype=torch.float64)
    aa = aa / aa.norm(dim=1, keepdim=True) * (torch.rand(500, 1, generator=g, dtype=torch.float64) * 3.0 + 0.01)
    R = po.aa2matrot(aa)
    # the package normalises by (theta + 1e-6): R is a rotation to ~1e-6, and the log map inverts it to the same order
    assert torch.allclose(R.transpose(1, 2) @ R, torch.eye(3, dtype=torch.float64).expand_as(R), atol=5e-6)
    back = po.matrot2aa(R)
    assert torch.allclose(back, aa, atol=2e-5)
    # every quaternion branch is reachable and agrees with the matrix exponential of the result
    x6 = torch.randn(4000, 6, generator=g, dtype=torch.float64)
    Rm = po.rot6d_decode(x6)
    a = po.matrot2aa(Rm)
    th = a.norm(dim=1)
    K = torch.zeros(4000, 3, 3, dtype=torch.float64)
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0], K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -a[:, 2], a[:, 1], a[:, 2], -a[:, 0], -a[:, 1], a[:, 0]
    Rexp = torch.matrix_exp(K)
    assert torch.allclose(Rexp, Rm, atol=1e-9)
    m = Rm.transpose(1, 2)
    d2 = m[:, 2, 2] < 1e-6
    branches = [(d2 & (m[:, 0, 0] > m[:, 1, 1])).sum(), (d2 & ~(m[:, 0, 0] > m[:, 1, 1])).sum(),
                (~d2 & (m[:, 0, 0] < -m[:, 1, 1])).sum(), (~d2 & ~(m[:, 0, 0] < -m[:, 1, 1])).sum()]
    assert all(int(b) > 50 for b in branches)
    assert th.max() <= np.pi + 1e-9


def test_dct_loss_matches_reference_method():
    d = np.load(os.path.join(G, "prior_dct.npz"))
    joints = torch.tensor(d["joints"], requires_grad=True)
    c = torch.tensor(d["c_dct"], requires_grad=True)
    loss = po.dct_loss(joints, torch.tensor(d["basis"]), c)
    loss.backward()
    assert abs(float(loss.detach()) - float(d["loss"])) < 1e-12 * max(1.0, abs(float(d["loss"])))
    np.testing.assert_allclose(joints.grad.numpy(), d["g_joints"], rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(c.grad.numpy(), d["g_c"], rtol=1e-10, atol=1e-14)


def test_world_placement_oracle_matches_reference_functions():
    """oracle/residuals_oracle.py against the literal verts_transform and FittingOP.body2world of the reference."""
    from oracle import residuals_oracle as ro
    d = np.load(os.path.join(G, "prior_world.npz"))
    rec, scale, cam = torch.tensor(d["rec"]), torch.tensor(d["scale"]), torch.tensor(d["cam_ext"])
    b2w = ro.body2world(rec[:, -3:], scale, cam)
    assert np.array_equal(b2w.numpy(), d["b2w"])
    vt = ro.verts_transform(torch.tensor(d["verts"]) * scale, b2w)
    assert np.array_equal(vt.numpy(), d["vt"])
