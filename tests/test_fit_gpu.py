"""GPU: one whole optimisation step (fit.FitProblem) vs the same step assembled from the CPU oracles."""
import numpy as np
import pytest
import torch

from oracle import chamfer_oracle as co
from oracle import prior_oracle as po
from oracle import residuals_oracle as ro
from oracle import smplx_oracle as so

pytestmark = pytest.mark.gpu


def _oracle_step(fpv, prob, dtype=torch.float64):
    """dtype=float32 runs the SAME assembly in plain torch float32 -- the arithmetic the reference itself runs."""
    from importlib import import_module
    fit = import_module("4dcapture-fpv_b200.fit")
    W = fit.LOSS_WEIGHTS
    p = prob.params.detach().cpu().to(dtype).requires_grad_(True)
    scale = prob.scale.detach().cpu().to(dtype).requires_grad_(True)
    cam = prob.camera_ext.detach().cpu().to(dtype).requires_grad_(True)
    data = prob.data.cpu().to(dtype)
    sl = lambda r: p[:, r[0]:r[1]]
    extra = {}
    c_dct = None
    if prob.front_end:
        r75 = po.convert_to_3D_rot(p)
        z = r75[:, 16:48]
        extra["vposer"] = torch.mean(z ** 2)
        wd = {k: getattr(prob.vposer, k).detach().cpu().to(dtype) for k in ("w1", "b1", "w2", "b2", "w3", "b3")}
        b2w = ro.body2world(r75[:, 72:75], scale, cam)
        v, j = so.smplx_forward(prob.constants, betas=r75[:, 6:16], global_orient=r75[:, 3:6],
                                body_pose=po.vposer_decode_aa(wd, z).view(p.shape[0], -1), transl=r75[:, 0:3],
                                left_hand_pose=r75[:, 48:60], right_hand_pose=r75[:, 60:72], dtype=dtype)
    else:
        b2w = ro.body2world(sl(fit.P_CAM), scale, cam)
        v, j = so.smplx_forward(prob.constants, betas=sl(fit.P_BETAS), global_orient=sl(fit.P_ORIENT),
                                body_pose=sl(fit.P_POSE), transl=sl(fit.P_TRANSL), left_hand_pose=sl(fit.P_LH),
                                right_hand_pose=sl(fit.P_RH), dtype=dtype)
    verts = ro.verts_transform(v * scale, b2w)
    joints = ro.verts_transform(j[:, 0:23], b2w)          # global_optimization.py:296-297: joints are NOT scaled
    scene = prob.host_scene.to(dtype)
    # indices from the canonical fp32 oracle on the fp32 vertices the GPU path sees; distances re-derived in
    # float64 through those indices so autograd flows exactly as torch.min would route it
    v32 = verts.detach().float().numpy()
    _, _, i_b2a, i_a2b = co.dist_chamfer(v32, prob.host_scene.numpy())
    i_b2a, i_a2b = torch.tensor(i_b2a), torch.tensor(i_a2b)
    d_a2b = ((verts - scene[i_a2b]) ** 2).sum(-1)
    d_b2a = ((scene.unsqueeze(0) - torch.gather(verts, 1, i_b2a.unsqueeze(-1).expand(-1, -1, 3))) ** 2).sum(-1)
    cid = prob.contact_ids.cpu()
    losses = dict(rec=torch.mean(torch.abs(data - p)), smoothing=ro.second_diff_l1(p),
                  contact=ro.contact_robust_loss(d_a2b[:, cid]), scene2body=d_b2a.mean(),
                  world_smoothing=ro.first_diff_l1(joints), vert_smoothing=ro.second_diff_l1(verts))
    if prob.front_end and prob.dct_batches:
        c_dct = prob.c_dct.detach().cpu().to(dtype).requires_grad_(True)
        extra["dct"] = po.dct_loss(joints, prob.dct_mtx.cpu().to(dtype), c_dct)
    losses.update(extra)
    total = sum(W[k] * x for k, x in losses.items())
    total.backward()
    if c_dct is not None:
        return total.item(), p.grad, scale.grad, cam.grad, c_dct.grad
    return total.item(), p.grad, scale.grad, cam.grad


def _check_grads(pairs, floors=None):
    """1e-5 of the largest entry, or 1.5x the distance of plain torch float32 from the float64 truth where that is
    larger (gradients that are ill-conditioned signed sums of ~30k vertex terms: the floor any fp32 path shares)."""
    for k, (got, ref, name) in enumerate(pairs):
        err = (got.cpu().double() - ref).abs().max().item()
        tol = 1e-5 * ref.abs().max().item() + 1e-8
        if floors is not None:
            tol = max(tol, 1.5 * (floors[k].double() - ref).abs().max().item())
        assert err <= tol, (name, err, tol)


def test_fit_step_matches_oracle(fpv, cuda_dev):
    prob = fpv.FitProblem(T=6, M=20000, device=cuda_dev, seed=1235)
    loss = prob.step()
    ref_loss, gp, gs, gc = _oracle_step(fpv, prob)
    _, *floors = _oracle_step(fpv, prob, torch.float32)
    assert loss.item() == pytest.approx(ref_loss, rel=1e-5)
    _check_grads([(prob.params.grad, gp, "params"), (prob.scale.grad, gs, "scale"), (prob.camera_ext.grad, gc, "camera_ext")], floors)
    l2 = prob.step()
    assert l2.item() == loss.item()                                   # same inputs -> bitwise same loss
    host = prob.step_e2e()
    assert host[0].item() == pytest.approx(ref_loss, rel=1e-5) and host[1].shape == (6, 106)


def test_fit_step_cuda_graph_replay_matches_eager(fpv, cuda_dev):
    """f1: the whole step captured as one CUDA graph reproduces the eager step bit for bit (the kernels are
    deterministic), and follows in-place parameter updates between replays."""
    prob = fpv.FitProblem(T=5, M=12000, device=cuda_dev, seed=1237, scene_order="morton")
    loss_e = prob.step().clone()
    grads_e = [t.grad.clone() for t in prob.leaves()]
    prob.capture()
    loss_g = prob.step_graph().clone()
    assert torch.equal(loss_g, loss_e)
    for g, t in zip(grads_e, prob.leaves()):
        assert torch.equal(t.grad, g)
    with torch.no_grad():
        prob.params.add_(0.003 * torch.randn_like(prob.params))
    loss_g2 = prob.step_graph().clone()
    grads_g2 = [t.grad.clone() for t in prob.leaves()]
    loss_e2 = prob.step().clone()
    assert torch.equal(loss_g2, loss_e2)
    for g, t in zip(grads_g2, prob.leaves()):
        assert torch.equal(t.grad, g)


def test_fit_step_with_reference_front_end_matches_oracle(fpv, cuda_dev):
    """f2 + f3 inside the step: 78-D 6D row -> convert_to_3D_rot -> VPoser decode -> body model, plus the VPoser
    regulariser (:263) and the DCT prior (:310), against the float64 oracle assembly."""
    prob = fpv.FitProblem(T=8, M=6000, device=cuda_dev, seed=1238, front_end=True, dct_frames=4)
    assert prob.params.shape == (8, 78) and prob.dct_batches == 2 and len(prob.leaves()) == 4
    loss = prob.step()
    ref_loss, gp, gs, gc, gd = _oracle_step(fpv, prob)
    _, *floors = _oracle_step(fpv, prob, torch.float32)
    assert loss.item() == pytest.approx(ref_loss, rel=1e-5)
    _check_grads([(prob.params.grad, gp, "params"), (prob.scale.grad, gs, "scale"),
                  (prob.camera_ext.grad, gc, "camera_ext"), (prob.c_dct.grad, gd, "c_dct")], floors)
    assert prob.step().item() == loss.item()
    prob.capture()
    assert torch.equal(prob.step_graph(), loss)


def test_fit_e2e_graph_matches_eager(fpv, cuda_dev):
    """The end-to-end path through the captured step (host inputs -> static buffers -> replay -> host results) returns
    what the eager step returns."""
    prob = fpv.FitProblem(T=5, M=9000, device=cuda_dev, seed=1239, front_end=True, dct_frames=2)
    eager = prob.step_e2e()
    prob.capture()
    graph = prob.step_e2e_graph()
    assert len(graph) == len(eager)
    for g, e in zip(graph, eager):
        assert torch.equal(g, e)


def _oracle_local_step(fpv, prob):
    """FittingOP.cal_loss2 (global_optimization.py:368-447) assembled from the float64 oracles."""
    p = prob.params.detach().cpu().double().requires_grad_(True)
    scale = prob.scale.detach().cpu().double()
    cam = prob.camera_ext.detach().cpu().double()
    data = prob.data.cpu().double()
    r75 = po.convert_to_3D_rot(p)
    wd = {k: getattr(prob.vposer, k).detach().cpu().double() for k in ("w1", "b1", "w2", "b2", "w3", "b3")}
    b2w = ro.body2world(r75[:, 72:75], scale, cam)
    v, _ = so.smplx_forward(prob.constants, betas=r75[:, 6:16], global_orient=r75[:, 3:6],
                            body_pose=po.vposer_decode_aa(wd, r75[:, 16:48]).view(p.shape[0], -1), transl=r75[:, 0:3],
                            left_hand_pose=r75[:, 48:60], right_hand_pose=r75[:, 60:72], dtype=torch.float64)
    verts = ro.verts_transform(v * scale, b2w)
    loss_rec = torch.mean(torch.abs(data - p))                                   # :376 (weights all one)
    diff_local = p[0:-1] - p[1:]
    loss_local = torch.mean(torch.abs(diff_local[0:-1] - diff_local[1:]))        # :381-382
    diff = verts[0:-1] - verts[1:]
    loss_smooth = torch.mean(torch.abs(diff[0:-1] - diff[1:]))                   # :404-405
    vl, vr = verts[:, prob.left_ids.cpu()], verts[:, prob.right_ids.cpu()]
    dl, dr = vl[0:-1] - vl[1:], vr[0:-1] - vr[1:]                                # :412-413
    w_right = prob.contact_weight.cpu().double().clone()                         # :415-420
    w_left = 1 - w_right
    w_left[w_left < 0.5] = 0.0
    w_right[w_right < 0.5] = 0.0
    wl = w_left[1:].unsqueeze(1).unsqueeze(1).repeat(1, dl.shape[1], dl.shape[2])
    wr = w_right[1:].unsqueeze(1).unsqueeze(1).repeat(1, dr.shape[1], dr.shape[2])
    loss_cs = torch.mean(torch.abs(dl * wl)) + torch.mean(torch.abs(dr * wr))    # :429
    total = loss_smooth + loss_local + loss_rec + loss_cs                        # :549
    total.backward()
    return total.item(), p.grad


def test_local_mode_step_matches_cal_loss2(fpv, cuda_dev):
    """mode='local' (second stage of fitting(mode='local'), :536-556): the vertex-space second difference over all
    10,475 vertices and the contact-weighted leg velocity, forward and backward to the 78-D row."""
    prob = fpv.FitProblem(T=7, M=0, device=cuda_dev, seed=1241, front_end=True, mode="local")
    loss = prob.step()
    ref_loss, gp = _oracle_local_step(fpv, prob)
    assert loss.item() == pytest.approx(ref_loss, rel=1e-5)
    _check_grads([(prob.params.grad, gp, "params")])
    assert prob.scale.grad is None or float(prob.scale.grad.abs().max()) >= 0.0
    before = prob.params.detach().clone()
    l2 = prob.step(update=True)
    assert l2.item() == loss.item() and not torch.equal(prob.params.detach(), before)
    prob.capture(update=True)
    a = prob.step_graph().item()
    b = prob.step_graph().item()
    assert np.isfinite(a) and np.isfinite(b) and b != a                          # the captured update moves the body


def test_batched_clips_equal_independent_clips(fpv, cuda_dev):
    """BASELINE.json configs[4] (independent clips batched into one step): with equal-length clips every loss term is a
    mean over frames, so the batched step must equal the average of the single-clip steps -- loss and per-frame
    gradients -- and no temporal residual may couple frames of different clips."""
    C, Tc, M = 3, 6, 8000
    big = fpv.FitProblem(T=C * Tc, M=M, device=cuda_dev, seed=1242, front_end=True, dct_frames=3, clips=C)
    assert big.dct_batches == C * (Tc // 3)
    loss = big.step()
    per = big.dct_batches // C
    singles, total = [], 0.0
    for c in range(C):
        one = fpv.FitProblem(T=Tc, M=M, device=cuda_dev, seed=1242, front_end=True, dct_frames=3)
        with torch.no_grad():
            sl = slice(c * Tc, (c + 1) * Tc)
            one.params.copy_(big.params[sl]); one.data.copy_(big.data[sl]); one.camera_ext.copy_(big.camera_ext[sl])
            one.scale.copy_(big.scale); one.c_dct.copy_(big.c_dct[c * per:(c + 1) * per])
        assert torch.equal(one.scene, big.scene)
        total += one.step().item() / C
        singles.append(one)
    assert loss.item() == pytest.approx(total, rel=1e-5)
    for c, one in enumerate(singles):
        sl = slice(c * Tc, (c + 1) * Tc)
        ref = one.params.grad / C
        err = (big.params.grad[sl] - ref).abs().max().item()
        assert err <= 1e-5 * ref.abs().max().item() + 1e-9, (c, err)
        refc = one.camera_ext.grad / C
        assert (big.camera_ext.grad[sl] - refc).abs().max().item() <= 1e-5 * refc.abs().max().item() + 1e-9
    big.capture(update=True)
    a = big.step_graph().item()
    assert np.isfinite(a)


def test_reference_literal_call_sequence(fpv, cuda_dev):
    """mode='reference': cal_loss as written -- repeated scene (:176), chamferDist on the contact vertices with dist2
    discarded (:290-294), loss of :570 -- against the float64 assembly; the repeated scene must take the indexed path."""
    prob = fpv.FitProblem(T=5, M=9000, device=cuda_dev, seed=1243, front_end=True, dct_frames=2, mode="reference")
    loss = prob.step()
    assert prob.s_verts_batch.shape == (5, 9000, 3) and prob.s_verts_batch.stride(0) == 27000     # a real copy
    p = prob.params.detach().cpu().double().requires_grad_(True)
    scale = prob.scale.detach().cpu().double().requires_grad_(True)
    cam = prob.camera_ext.detach().cpu().double()
    r75 = po.convert_to_3D_rot(p)
    wd = {k: getattr(prob.vposer, k).detach().cpu().double() for k in ("w1", "b1", "w2", "b2", "w3", "b3")}
    b2w = ro.body2world(r75[:, 72:75], scale, cam)
    v, _ = so.smplx_forward(prob.constants, betas=r75[:, 6:16], global_orient=r75[:, 3:6],
                            body_pose=po.vposer_decode_aa(wd, r75[:, 16:48]).view(5, -1), transl=r75[:, 0:3],
                            left_hand_pose=r75[:, 48:60], right_hand_pose=r75[:, 60:72], dtype=torch.float64)
    verts = ro.verts_transform(v * scale, b2w)
    cv = verts[:, prob.contact_ids.cpu()]
    scene = prob.host_scene.double()
    _, _, _, i_a2b = co.dist_chamfer(cv.detach().float().numpy(), prob.host_scene.numpy())
    d = ((cv - scene[torch.tensor(i_a2b)]) ** 2).sum(-1)
    total = 0.1 * ro.contact_robust_loss(d) + ro.second_diff_l1(p) + torch.mean(torch.abs(prob.data.cpu().double() - p))
    total.backward()
    assert loss.item() == pytest.approx(total.item(), rel=1e-5)
    _check_grads([(prob.params.grad, p.grad, "params"), (prob.scale.grad, scale.grad, "scale")])
