"""GPU parity: SMPL-X forward/backward kernels vs the float64 torch restatement of smplx.lbs
(oracle/smplx_oracle.py).  Tolerance: 1e-5 relative to the coordinate scale (~1 m) for vertices and
joints; gradients 1e-5 relative to the largest gradient entry of each parameter block."""
import pytest
import torch

from oracle import smplx_oracle as so

pytestmark = pytest.mark.gpu

KEYS = ["betas", "global_orient", "body_pose", "transl", "left_hand_pose", "right_hand_pose",
        "expression", "jaw_pose", "leye_pose", "reye_pose"]
DIMS = dict(betas=10, global_orient=3, body_pose=63, transl=3, left_hand_pose=12, right_hand_pose=12,
            expression=10, jaw_pose=3, leye_pose=3, reye_pose=3)


def _params(T, seed, scale=0.4):
    g = torch.Generator().manual_seed(seed)
    p = {k: scale * torch.randn(T, DIMS[k], generator=g) for k in KEYS}
    p["betas"] = torch.randn(T, 10, generator=g)
    p["transl"] = torch.randn(T, 3, generator=g)
    return p


@pytest.mark.parametrize("V,T", [(10475, 6), (700, 33), (129, 1)])
def test_forward_matches_oracle(fpv, cuda_dev, V, T):
    c = fpv.synthetic.make_body_constants(seed=V, num_verts=V)
    model = fpv.create(constants=c, batch_size=T).to(cuda_dev)
    p = _params(T, seed=V + T)
    v64, j64 = so.smplx_forward(c, **p, dtype=torch.float64)
    out = model(return_verts=True, **{k: v.to(cuda_dev) for k, v in p.items()})
    assert out.vertices.shape == (T, V, 3) and out.joints.shape == (T, 76, 3)
    torch.testing.assert_close(out.vertices.cpu().double(), v64, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(out.joints.cpu().double(), j64, rtol=1e-5, atol=1e-5)


def test_reference_call_signature_and_defaults(fpv, cuda_dev):
    """The exact call of global_optimization.py:280-283: six keyword tensors, the rest default to the
    module's zero parameters."""
    T, V = 5, 900
    c = fpv.synthetic.make_body_constants(seed=3, num_verts=V)
    model = fpv.create("./models", model_type="smplx", gender="neutral", ext="npz", num_pca_comps=12,
                       create_global_orient=True, create_body_pose=True, create_betas=True,
                       create_left_hand_pose=True, create_right_hand_pose=True, create_expression=True,
                       create_jaw_pose=True, create_leye_pose=True, create_reye_pose=True, create_transl=True,
                       batch_size=T, constants=c).to(cuda_dev)
    p = _params(T, 11)
    body_param_ = {k: p[k].to(cuda_dev) for k in ["transl", "global_orient", "betas", "left_hand_pose", "right_hand_pose"]}
    smplx_output = model(return_verts=True, body_pose=p["body_pose"].to(cuda_dev), **body_param_)
    ref = {k: p[k] for k in ["betas", "global_orient", "body_pose", "transl", "left_hand_pose", "right_hand_pose"]}
    v64, j64 = so.smplx_forward(c, **ref, dtype=torch.float64)
    torch.testing.assert_close(smplx_output.vertices.cpu().double(), v64, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(smplx_output.joints[:, 0:23, :].cpu().double(), j64[:, 0:23], rtol=1e-5, atol=1e-5)
    with pytest.raises(RuntimeError):
        model(betas=torch.zeros(T + 1, 10, device=cuda_dev))


@pytest.mark.parametrize("use_joints", [False, True])
def test_backward_matches_float64_autograd(fpv, cuda_dev, use_joints):
    T, V = 4, 1500
    c = fpv.synthetic.make_body_constants(seed=21, num_verts=V)
    model = fpv.create(constants=c, batch_size=T).to(cuda_dev)
    p = _params(T, 5)
    g = torch.Generator().manual_seed(77)
    gv = torch.randn(T, V, 3, generator=g)
    gj = torch.randn(T, 76, 3, generator=g)
    po = {k: v.clone().double().requires_grad_(True) for k, v in p.items()}
    v64, j64 = so.smplx_forward(c, **po, dtype=torch.float64)
    lo = (v64 * gv.double()).sum() + ((j64 * gj.double()).sum() if use_joints else 0.0)
    lo.backward()
    # the same loss through plain torch float32 autograd (the arithmetic the reference itself runs): where a
    # gradient is an ill-conditioned signed sum of thousands of terms, ITS distance from the float64 truth is the
    # floor any fp32 implementation shares, and the bar is 1e-5 or that floor, whichever is larger
    p32 = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    v32, j32 = so.smplx_forward(c, **p32, dtype=torch.float32)
    ((v32 * gv).sum() + ((j32 * gj).sum() if use_joints else 0.0)).backward()
    grads = []
    for _ in range(2):
        pg = {k: v.clone().to(cuda_dev).requires_grad_(True) for k, v in p.items()}
        out = model(**pg)
        lg = (out.vertices * gv.to(cuda_dev)).sum() + ((out.joints * gj.to(cuda_dev)).sum() if use_joints else 0.0)
        lg.backward()
        grads.append({k: pg[k].grad.cpu() for k in KEYS})
    for k in KEYS:
        ref = po[k].grad
        tol = 1e-5 * float(ref.abs().max()) + 1e-7
        floor32 = float((p32[k].grad.double() - ref).abs().max())
        err = float((grads[0][k].double() - ref).abs().max())
        assert err <= max(tol, 1.5 * floor32), (k, err, tol, floor32)
        assert torch.equal(grads[0][k], grads[1][k]), f"{k}: backward is not run-to-run deterministic"


def test_zero_pose_jaw_eyes_have_finite_gradients(fpv, cuda_dev):
    """jaw / eye poses are exactly zero in the reference call; Rodrigues backward must stay finite there."""
    T, V = 2, 400
    c = fpv.synthetic.make_body_constants(seed=2, num_verts=V)
    model = fpv.create(constants=c, batch_size=T).to(cuda_dev)
    jaw = torch.zeros(T, 3, device=cuda_dev, requires_grad=True)
    go = torch.zeros(T, 3, device=cuda_dev, requires_grad=True)
    out = model(jaw_pose=jaw, global_orient=go)
    out.vertices.square().sum().backward()
    assert torch.isfinite(jaw.grad).all() and torch.isfinite(go.grad).all()
    jo = torch.zeros(T, 3, dtype=torch.float64, requires_grad=True)
    goo = torch.zeros(T, 3, dtype=torch.float64, requires_grad=True)
    z = lambda d: torch.zeros(T, d, dtype=torch.float64)
    v64, _ = so.smplx_forward(c, betas=z(10), global_orient=goo, body_pose=z(63), transl=z(3), left_hand_pose=z(12),
                              right_hand_pose=z(12), jaw_pose=jo)
    v64.square().sum().backward()
    torch.testing.assert_close(go.grad.cpu().double(), goo.grad, rtol=1e-4, atol=1e-4 * float(goo.grad.abs().max()))
    torch.testing.assert_close(jaw.grad.cpu().double(), jo.grad, rtol=1e-4, atol=1e-4 * float(jo.grad.abs().max()) + 1e-6)
