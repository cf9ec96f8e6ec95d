"""GPU parity of the round-2 operators around the chamfer searches:
  * the single-direction forms and the unbatched helpers of chamfer_python.py (pairwise_dist :4-9, NN_loss :12-15);
  * chamferDist's lazy second direction and the detection of the reference's materialised T-fold scene (:176);
  * the fused scene->body sum (no [T,M] output; gradient from per-vertex integer accumulators);
  * the capturable Adam update (global_optimization.py:188, :592);
  * a multi-step DRIFTING fit at the benchmarked scene size with seeds carried from step to step.
Bar: indices and distances bit-exact against the C oracle; sums, losses and gradients within 1e-5 relative."""
import importlib

import numpy as np
import pytest
import torch

from oracle import chamfer_oracle as co
from oracle import chamfer_ref_port as port

pytestmark = pytest.mark.gpu


def _clouds(rng, T, N, M):
    a = (rng.standard_normal((T, N, 3)) * 0.4 + [1.0, -0.5, 1.0]).astype(np.float32)
    b = (rng.random((M, 3)) * [8, 8, 3] - [4, 4, 0]).astype(np.float32)
    return a, b


@pytest.mark.parametrize("T,N,M", [(1, 5, 7), (3, 700, 300), (2, 1500, 9000), (4, 2500, 30000)])
def test_single_direction_ops_match_oracle(fpv, cuda_dev, T, N, M):
    rng = np.random.default_rng(T * 31 + N + M)
    a, b = _clouds(rng, T, N, M)
    want = co.dist_chamfer(a, b)
    ta, tb = torch.tensor(a, device=cuda_dev), torch.tensor(b, device=cuda_dev)
    d, i = fpv.body_to_scene(ta, tb, idx_dtype=torch.int64)
    assert np.array_equal(d.cpu().numpy(), want[1]) and np.array_equal(i.cpu().numpy(), want[3])
    d, i = fpv.scene_to_body(ta, tb, idx_dtype=torch.int64)
    assert np.array_equal(d.cpu().numpy(), want[0]) and np.array_equal(i.cpu().numpy(), want[2])
    # per-batch scene (not shared) takes the brute-force single-direction path
    tb3 = torch.tensor(np.stack([b + 0.01 * k for k in range(T)]).astype(np.float32), device=cuda_dev)
    want3 = co.dist_chamfer(a, tb3.cpu().numpy())
    d, i = fpv.body_to_scene(ta, tb3, idx_dtype=torch.int64)
    assert np.array_equal(d.cpu().numpy(), want3[1]) and np.array_equal(i.cpu().numpy(), want3[3])
    d, i = fpv.scene_to_body(ta, tb3, idx_dtype=torch.int64)
    assert np.array_equal(d.cpu().numpy(), want3[0]) and np.array_equal(i.cpu().numpy(), want3[2])


def test_single_direction_gradients(fpv, cuda_dev):
    rng = np.random.default_rng(5)
    a, b = _clouds(rng, 3, 900, 6000)
    g2 = rng.standard_normal((3, 900)).astype(np.float32)
    g1 = rng.standard_normal((3, 6000)).astype(np.float32)
    _, _, i1, i2 = co.dist_chamfer(a, b)
    ta = torch.tensor(a, device=cuda_dev, requires_grad=True)
    tb = torch.tensor(b, device=cuda_dev, requires_grad=True)
    d, _ = fpv.body_to_scene(ta, tb)
    (d * torch.tensor(g2, device=cuda_dev)).sum().backward()
    ga, gb = co.dist_chamfer_bwd(a, b, None, g2, i1, i2)
    np.testing.assert_allclose(ta.grad.cpu().numpy(), ga, rtol=1e-5, atol=1e-5 * np.abs(ga).max())
    np.testing.assert_allclose(tb.grad.cpu().numpy().reshape(gb.sum(0).shape), gb.sum(0), rtol=1e-5, atol=1e-5 * np.abs(gb).max())
    ta.grad = None
    d, _ = fpv.scene_to_body(ta, tb.detach())
    (d * torch.tensor(g1, device=cuda_dev)).sum().backward()
    ga, _ = co.dist_chamfer_bwd(a, b, g1, None, i1, i2)
    np.testing.assert_allclose(ta.grad.cpu().numpy(), ga, rtol=1e-5, atol=1e-5 * np.abs(ga).max())


def test_pairwise_dist_and_nn_loss_match_reference_arithmetic(fpv, cuda_dev):
    """chamfer_python.py:4-15 restated in oracle/chamfer_ref_port.py (expanded form): equal within its own rounding on
    unit-scale clouds; NN_loss both dims against the canonical oracle."""
    rng = np.random.default_rng(11)
    x = rng.standard_normal((257, 3)).astype(np.float32)
    y = rng.standard_normal((411, 3)).astype(np.float32)
    tx, ty = torch.tensor(x, device=cuda_dev), torch.tensor(y, device=cuda_dev)
    P = fpv.pairwise_dist(tx, ty).cpu()
    Pref = port.pairwise_dist(torch.tensor(x), torch.tensor(y))
    assert P.shape == (257, 411)
    torch.testing.assert_close(P, Pref, rtol=1e-5, atol=2e-5)
    d_yx, i_yx = co.nn(y, x)            # for every y_j the nearest x_i  (min over dim 0)
    d_xy, i_xy = co.nn(x, y)            # for every x_i the nearest y_j  (min over dim 1)
    assert fpv.NN_loss(tx, ty, 0).item() == pytest.approx(float(d_yx.astype(np.float64).mean()), rel=1e-6)
    assert fpv.NN_loss(tx, ty, 1).item() == pytest.approx(float(d_xy.astype(np.float64).mean()), rel=1e-6)
    for dim in (0, 1):
        assert fpv.NN_loss(tx, ty, dim).item() == pytest.approx(port.NN_loss(torch.tensor(x), torch.tensor(y), dim).item(), rel=2e-5)


def test_chamferdist_directions_and_literal_reference_call(fpv, cuda_dev):
    """The reference's literal call sequence (global_optimization.py:173-176, :290-295): a materialised .repeat(T,1,1)
    scene, chamferDist()(contact_verts.contiguous(), scene.contiguous()), dist2 discarded.  It must take the shared
    (indexed) path, compute dist1 only, and equal the oracle."""
    ch = importlib.import_module("4dcapture-fpv_b200.chamfer")
    rng = np.random.default_rng(21)
    T, Nc, M = 6, 1200, 40000
    a, b = _clouds(rng, T, Nc, M)
    want = co.dist_chamfer(a, b)
    s_verts_batch = torch.tensor(b, device=cuda_dev).unsqueeze(0).repeat(T, 1, 1)       # :175-176
    body = torch.tensor(a, device=cuda_dev, requires_grad=True)
    _, b_shared, shared = ch._prep(body, s_verts_batch.contiguous())
    assert shared and b_shared.shape[0] == 1                                           # routed to the shared path
    for _ in range(2):                                                                 # second call: cached verdict
        contact_dist, none = fpv.chamferDist()(body.contiguous(), s_verts_batch.contiguous())
        assert none is None
        assert np.array_equal(contact_dist.detach().cpu().numpy(), want[1])
    loss = torch.mean(torch.sqrt(contact_dist + 1e-4) / (torch.sqrt(contact_dist + 1e-4) + 1.0))   # :295
    loss.backward()
    assert torch.isfinite(body.grad).all()
    d1, d2 = fpv.chamferDist(directions="both")(body.detach(), s_verts_batch)
    assert np.array_equal(d1.cpu().numpy(), want[1]) and np.array_equal(d2.cpu().numpy(), want[0])
    n1, d2b = fpv.chamferDist(directions="dist2")(body.detach(), s_verts_batch)
    assert n1 is None and torch.equal(d2b, d2)
    # a batch that is NOT a repeat keeps the general path and its own per-frame answers
    s2 = s_verts_batch.clone()
    s2[3, 17] += 0.5
    _, _, shared2 = ch._prep(body, s2)
    assert not shared2
    want2 = co.dist_chamfer(a, s2.cpu().numpy())
    d1b, _ = fpv.chamferDist()(body.detach(), s2)
    assert np.array_equal(d1b.cpu().numpy(), want2[1])
    s_verts_batch[2, 5] += 1.0                                                         # in-place edit bumps the version
    _, _, shared3 = ch._prep(body, s_verts_batch)
    assert not shared3


@pytest.mark.parametrize("T,N,M,presorted", [(3, 1000, 9000, False), (5, 2500, 40000, True), (2, 10475, 200000, True)])
def test_scene_to_body_sum_matches_full_outputs(fpv, cuda_dev, T, N, M, presorted):
    fit = importlib.import_module("4dcapture-fpv_b200.fit")
    rng = np.random.default_rng(N + M)
    a, b = _clouds(rng, T, N, M)
    if presorted:
        b = fit._morton_sorted(torch.tensor(b)).numpy()
    d1, _, i1, _ = co.dist_chamfer(a, b)
    want_sum = d1.astype(np.float64).sum(1)
    bt = torch.tensor(b, device=cuda_dev).unsqueeze(0)
    g = rng.standard_normal(T).astype(np.float32)
    ga, _ = co.dist_chamfer_bwd(a, b, np.repeat(g[:, None], M, 1), None, i1, np.zeros((T, N), np.int64))
    state = fpv.SearchState()
    grads, sums = [], []
    for call in range(3):
        ta = torch.tensor(a, device=cuda_dev, requires_grad=True)
        if call == 2:   # garbage seeds: hints only
            sd = state.seeds[("b2a", T, M, cuda_dev.index)]
            sd.copy_(torch.randint(-7, 2 * N, sd.shape, device=cuda_dev, dtype=torch.int32))
        s = fpv.scene_to_body_sum(ta, bt, clip=True, state=state)
        assert s.shape == (T,)
        (s * torch.tensor(g, device=cuda_dev)).sum().backward()
        sums.append(s.detach().cpu().numpy())
        grads.append(ta.grad.cpu().numpy())
    np.testing.assert_allclose(sums[0], want_sum, rtol=1e-6)
    np.testing.assert_allclose(grads[0], ga, rtol=1e-5, atol=1e-5 * np.abs(ga).max())
    for k in (1, 2):    # seeded, garbage-seeded: bitwise the same (integer accumulators, fixed-order sums)
        assert np.array_equal(sums[k], sums[0]) and np.array_equal(grads[k], grads[0])
    # a coarser fixed point (every |value| < 2^41: the accumulate pass then reduces two limbs per coordinate instead of
    # three -- the path a 1 M-point scene takes) gives the same gradient up to its 2^-33 relative resolution
    sc = fpv.spatial.cached_scene(bt)
    assert sc.M == M
    shift = sc.fix_shift()
    sc._fix_shift = shift - 8
    try:
        ta = torch.tensor(a, device=cuda_dev, requires_grad=True)
        s = fpv.scene_to_body_sum(ta, bt, clip=True, state=state)
        (s * torch.tensor(g, device=cuda_dev)).sum().backward()
        assert np.array_equal(s.detach().cpu().numpy(), sums[0])
        np.testing.assert_allclose(ta.grad.cpu().numpy(), grads[0], rtol=1e-6, atol=1e-7 * np.abs(ga).max())
    finally:
        sc._fix_shift = shift
    # the combined form returns the same sum plus the body->scene direction
    s2, d_a2b, i_a2b = fpv.fit_chamfer_terms(torch.tensor(a, device=cuda_dev), bt, state=fpv.SearchState())
    _, d2, _, i2 = co.dist_chamfer(a, b)
    assert np.array_equal(s2.cpu().numpy(), sums[0])
    assert np.array_equal(d_a2b.cpu().numpy(), d2) and np.array_equal(i_a2b.cpu().numpy(), i2)


def test_scene_to_body_sum_special_values(fpv, cuda_dev):
    """A non-finite scene point poisons the sum like the reference's min would (inf / nan) and adds no gradient; ties go
    to the lowest vertex index (counts land on it)."""
    rng = np.random.default_rng(3)
    a = rng.integers(-4, 5, (2, 600, 3)).astype(np.float32)
    b = rng.integers(-4, 5, (5000, 3)).astype(np.float32)
    d1, _, i1, _ = co.dist_chamfer(a, b)
    ta = torch.tensor(a, device=cuda_dev, requires_grad=True)
    s = fpv.scene_to_body_sum(ta, torch.tensor(b, device=cuda_dev), options=fpv.SearchOptions(engine="spatial"))
    np.testing.assert_allclose(s.detach().cpu().numpy(), d1.astype(np.float64).sum(1), rtol=1e-6)
    s.sum().backward()
    ga, _ = co.dist_chamfer_bwd(a, b, np.ones((2, 5000), np.float32), None, i1, np.zeros((2, 600), np.int64))
    np.testing.assert_allclose(ta.grad.cpu().numpy(), ga, rtol=1e-5, atol=1e-5 * np.abs(ga).max())
    b2 = b.copy()
    b2[77] = [np.nan, 0, 0]
    tb = torch.tensor(b2, device=cuda_dev)
    ta2 = torch.tensor(a, device=cuda_dev, requires_grad=True)
    s2 = fpv.scene_to_body_sum(ta2, tb)
    assert not np.isfinite(s2.detach().cpu().numpy()).any()
    s2.sum().backward()
    keep = np.ones(5000, bool)
    keep[77] = False
    ga2, _ = co.dist_chamfer_bwd(a, b[keep], np.ones((2, 4999), np.float32), None, co.dist_chamfer(a, b[keep])[2],
                                 np.zeros((2, 600), np.int64))
    np.testing.assert_allclose(ta2.grad.cpu().numpy(), ga2, rtol=1e-5, atol=1e-5 * np.abs(ga2).max())


def test_adam_update_matches_torch(fpv, cuda_dev):
    """fpv_adam_update against torch.optim.Adam (the reference's optimiser, :188) over 25 steps."""
    fit = importlib.import_module("4dcapture-fpv_b200.fit")
    L = fpv._lib.lib()
    g = torch.Generator().manual_seed(4)
    p0 = torch.randn(300, 78, generator=g)
    ref = p0.clone().to(cuda_dev).requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=fit.ADAM["lr"])
    mine = p0.clone().to(cuda_dev)
    m, v = torch.zeros_like(mine), torch.zeros_like(mine)
    step = torch.zeros(1, device=cuda_dev)
    for k in range(25):
        grad = (torch.randn(300, 78, generator=g) * (0.1 + k)).to(cuda_dev)
        ref.grad = grad.clone()
        opt.step()
        fpv._lib.check(L.fpv_adam_tick(fpv._lib.ptr(step), fpv._lib.stream_ptr()))
        fpv._lib.check(L.fpv_adam_update(fpv._lib.ptr(mine), fpv._lib.ptr(grad), fpv._lib.ptr(m), fpv._lib.ptr(v),
                                         mine.numel(), fit.ADAM["lr"], 0.9, 0.999, 1e-8, fpv._lib.ptr(step),
                                         fpv._lib.stream_ptr()))
    assert step.item() == 25.0
    torch.testing.assert_close(mine, ref.detach(), rtol=1e-5, atol=1e-6)


def test_drifting_fit_at_benchmark_scene_size_stays_exact(fpv, cuda_dev):
    """T = 64 frames against the 1 M-point scene, 6 optimiser steps with the Adam update moving the body every step and
    the seeds / frozen body order carried from step to step: after every step the production engines (seeded) are
    compared with the C oracle on sampled rows of both directions, and the fused sum with the full [T,M] output."""
    T, M = 64, 1_000_000
    prob = fpv.FitProblem(T=T, M=M, device=cuda_dev, seed=1240, front_end=True, idx_dtype=torch.int32)
    rng = np.random.default_rng(9)
    rows = rng.integers(0, M, 150)
    cols = rng.integers(0, 10475, 60)
    s_np = prob.scene[0].cpu().numpy()
    check_state = fpv.SearchState()               # the checker's own seeds also go stale from step to step
    losses = []
    for step in range(6):
        losses.append(prob.step(update=True).item())
        with torch.no_grad():
            verts, _, _ = prob._body()
        d_b2a, d_a2b, i_b2a, i_a2b = fpv.distChamfer(verts, prob.scene, idx_dtype=torch.int32, clip=True, state=check_state)
        v_np = verts.cpu().numpy()
        for t in (0, 17, T - 1):
            d, i = co.nn(s_np[rows], v_np[t])
            assert np.array_equal(d, d_b2a[t].cpu().numpy()[rows]) and np.array_equal(i, i_b2a[t].cpu().numpy()[rows]), step
            d, i = co.nn(v_np[t][cols], s_np)
            assert np.array_equal(d, d_a2b[t].cpu().numpy()[cols]) and np.array_equal(i, i_a2b[t].cpu().numpy()[cols]), step
        # the fused kernel (seeded from the PREVIOUS step's winners inside prob.search_state) against the full output
        s = fpv.scene_to_body_sum(verts, prob.scene, clip=True, state=prob.search_state)
        np.testing.assert_allclose(s.cpu().numpy(), d_b2a.double().sum(1).cpu().numpy(), rtol=1e-6)
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]          # the optimiser is really descending
    moved = (prob.params.detach().cpu() - prob._host_init).abs().max().item()
    assert moved > 0.02                                                 # 6 Adam steps of lr 0.005: the body drifted


def test_fused_terms_tolerate_retain_graph(fpv, cuda_dev):
    """The reference loop calls loss.backward(retain_graph=True) (global_optimization.py:591): the backward of the fused
    terms must not consume or change what the forward saved -- a second backward adds exactly the same gradient."""
    rng = np.random.default_rng(17)
    a, b = _clouds(rng, 3, 1200, 20000)
    ta = torch.tensor(a, device=cuda_dev, requires_grad=True)
    tb = torch.tensor(b, device=cuda_dev).unsqueeze(0)
    s, d, _ = fpv.fit_chamfer_terms(ta, tb, state=fpv.SearchState())
    loss = s.sum() / b.shape[0] + d.mean()
    loss.backward(retain_graph=True)
    g1 = ta.grad.clone()
    loss.backward()
    assert torch.equal(ta.grad, g1 + g1)
    assert float(g1.abs().max()) > 0
