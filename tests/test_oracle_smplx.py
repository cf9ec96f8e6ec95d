"""CPU: self-consistency of the SMPL-X restatement (oracle/smplx_oracle.py).

PARITY UNPINNED against the real [3P] smplx package (not installable, model files licence-gated):
these tests pin the restatement to the invariants the published algorithm guarantees.
"""
import math

import pytest
import torch

from conftest import load_pkg
from oracle import residuals_oracle as ro
from oracle import smplx_oracle as so

fpv = load_pkg()


@pytest.fixture(scope="module")
def consts():
    return fpv.synthetic.make_body_constants(seed=7, num_verts=600)


def _params(T, seed=0, zero_pose=False):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    z = 0.0 if zero_pose else 1.0
    return dict(betas=r(T, 10), global_orient=0.5 * z * r(T, 3), body_pose=0.3 * z * r(T, 63), transl=r(T, 3),
                left_hand_pose=0.3 * z * r(T, 12), right_hand_pose=0.3 * z * r(T, 12))


def test_shapes_and_kinematic_tree(consts):
    p = _params(3)
    v, j = so.smplx_forward(consts, **p)
    assert v.shape == (3, 600, 3) and j.shape == (3, 55 + 21, 3)
    parents = consts["parents"].tolist()
    assert parents[0] == -1 and all(parents[i] < i for i in range(1, 55))
    depth = [0] * 55
    for i in range(1, 55):
        depth[i] = depth[parents[i]] + 1
    assert max(depth) == 10 and 9 * 54 == 486  # SURVEY.md section 8a shape algebra


def test_zero_pose_is_shape_blend_plus_transl(consts):
    p = _params(2, zero_pose=True)
    v, j, inter = so.smplx_forward(consts, **p, return_intermediates=True)
    c64 = {k: (t.double() if t.is_floating_point() else t) for k, t in consts.items()}
    beta20 = torch.cat([p["betas"], torch.zeros(2, 10, dtype=torch.float64)], 1)
    v_shaped = c64["v_template"] + torch.einsum("bl,mkl->bmk", beta20, c64["shapedirs"])
    # fp32 skinning weights sum to 1 only within ~6e-8, so the blend of identity transforms scales by that
    torch.testing.assert_close(v, v_shaped + p["transl"].unsqueeze(1), rtol=0, atol=2e-7)
    J = torch.einsum("bik,ji->bjk", v_shaped, c64["J_regressor"])
    torch.testing.assert_close(j[:, :55], J + p["transl"].unsqueeze(1), rtol=0, atol=1e-9)


def test_rodrigues_is_a_rotation():
    g = torch.Generator().manual_seed(1)
    r = torch.randn(50, 3, generator=g, dtype=torch.float64)
    R = so.batch_rodrigues(r)
    eye = torch.eye(3, dtype=torch.float64).expand(50, 3, 3)
    torch.testing.assert_close(R @ R.transpose(1, 2), eye, rtol=0, atol=1e-7)
    torch.testing.assert_close(torch.linalg.det(R), torch.ones(50, dtype=torch.float64), rtol=0, atol=1e-7)
    ang = torch.acos(((R.diagonal(dim1=1, dim2=2).sum(-1) - 1) / 2).clamp(-1, 1))
    torch.testing.assert_close(ang, torch.remainder(r.norm(dim=1) + math.pi, 2 * math.pi).sub(math.pi).abs(), rtol=0, atol=1e-6)


def test_root_rotation_acts_rigidly_about_the_root_joint(consts):
    """Changing only global_orient rotates every vertex rigidly about the (shape-dependent) root joint."""
    p = _params(1, seed=4)
    p0 = dict(p, global_orient=torch.zeros(1, 3, dtype=torch.float64))
    v0, j0, i0 = so.smplx_forward(consts, **p0, return_intermediates=True)
    v1, j1 = so.smplx_forward(consts, **p)
    R = so.batch_rodrigues(p["global_orient"])[0]
    root = i0["J"][0, 0]
    t = p["transl"][0]
    expect = (v0[0] - t - root) @ R.T + root + t
    torch.testing.assert_close(v1[0], expect, rtol=0, atol=1e-9)


def test_fp32_oracle_tracks_fp64(consts):
    p = _params(4, seed=9)
    v64, j64 = so.smplx_forward(consts, **p, dtype=torch.float64)
    v32, j32 = so.smplx_forward(consts, **p, dtype=torch.float32)
    assert (v32.double() - v64).abs().max() < 2e-5 and (j32.double() - j64).abs().max() < 2e-5


def test_verts_transform_and_body2world_match_reference_expressions():
    g = torch.Generator().manual_seed(2)
    v = torch.randn(3, 17, 3, generator=g, dtype=torch.float64)
    M = torch.randn(3, 4, 4, generator=g, dtype=torch.float64)
    out = ro.verts_transform(v, M)
    expect = torch.einsum("trc,tpc->tpr", M[:, :3, :3], v) + M[:, :3, 3].unsqueeze(1)
    torch.testing.assert_close(out, expect, rtol=0, atol=1e-12)
    ct = torch.randn(3, 3, generator=g, dtype=torch.float64)
    s = torch.tensor(1.7, dtype=torch.float64)
    b2w = ro.body2world(ct, s, M)
    for t in range(3):  # literal per-frame loop of global_optimization.py:194-205
        pose = torch.eye(4, dtype=torch.float64)
        pose[:3, 3] = ct[t] * s
        torch.testing.assert_close(b2w[t], M[t] @ pose, rtol=0, atol=1e-12)


def test_residual_known_answers():
    x = torch.tensor([[0.0], [1.0], [4.0], [9.0], [7.0]])
    assert ro.second_diff_l1(x).item() == pytest.approx((2 + 2 + 7) / 3)
    assert ro.first_diff_l1(x).item() == pytest.approx((1 + 3 + 5 + 2) / 4)
    d = torch.tensor([0.0, 1.0 - 1e-4])
    assert ro.contact_robust_loss(d, 2.0).item() == pytest.approx(2.0 * 0.5 * (0.01 / 1.01 + 0.5), rel=1e-5)
