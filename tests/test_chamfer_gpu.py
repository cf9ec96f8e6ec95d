"""GPU parity: the CUDA chamfer path (through the C ABI) vs the CPU oracle and the golden vectors.
Bar: NN indices bit-exact (lowest index on ties), distances bit-exact (same canonical fp32 arithmetic),
gradients within 1e-5 relative (north_star tolerance)."""
import numpy as np
import pytest
import torch

from conftest import golden_cases, load_golden
from oracle import chamfer_oracle as co

pytestmark = pytest.mark.gpu


def _run(fpv, a, b, dev, idx_dtype=torch.int64):
    ta = torch.tensor(a, device=dev)
    tb = torch.tensor(b, device=dev)
    out = fpv.distChamfer(ta, tb, idx_dtype=idx_dtype)
    return [o.cpu().numpy() for o in out]


def _assert_exact(got, want):
    d1, d2, i1, i2 = got
    w1, w2, j1, j2 = want
    assert np.array_equal(i1, j1), f"i_b2a mismatches: {(i1 != j1).sum()}"
    assert np.array_equal(i2, j2), f"i_a2b mismatches: {(i2 != j2).sum()}"
    assert np.array_equal(d1, w1) and np.array_equal(d2, w2)


@pytest.mark.parametrize("case", golden_cases())
def test_golden_vectors_from_the_real_reference(fpv, cuda_dev, case):
    g = load_golden(case)
    d1, d2, i1, i2 = _run(fpv, g["a"], g["b"], cuda_dev)
    assert i1.dtype == np.int64 and d1.dtype == np.float32           # reference dtypes (chamfer_python.py:28)
    assert np.array_equal(i1, g["i_b2a"]) and np.array_equal(i2, g["i_a2b"])
    if case.startswith("random"):
        np.testing.assert_allclose(d1, g["d_b2a"], rtol=1e-5, atol=3e-5)
        np.testing.assert_allclose(d2, g["d_a2b"], rtol=1e-5, atol=3e-5)
    else:
        assert np.array_equal(d1, g["d_b2a"]) and np.array_equal(d2, g["d_a2b"])


@pytest.mark.parametrize("bs,N,M", [(1, 1, 1), (2, 5, 3), (1, 7, 1025), (3, 1000, 999), (2, 2049, 17),
                                    (1, 33, 8193), (4, 257, 4100), (1, 10475, 20000)])
def test_oracle_parity_ragged_sizes(fpv, cuda_dev, bs, N, M):
    rng = np.random.default_rng(bs * 1000003 + N * 101 + M)
    a = rng.standard_normal((bs, N, 3)).astype(np.float32)
    b = rng.standard_normal((bs, M, 3)).astype(np.float32)
    _assert_exact(_run(fpv, a, b, cuda_dev), co.dist_chamfer(a, b))


def test_lattice_ties_lowest_index(fpv, cuda_dev):
    rng = np.random.default_rng(1)
    a = rng.integers(-8, 9, (2, 3000, 3)).astype(np.float32)
    b = rng.integers(-8, 9, (2, 5000, 3)).astype(np.float32)
    _assert_exact(_run(fpv, a, b, cuda_dev), co.dist_chamfer(a, b))
    same = np.zeros((1, 600, 3), np.float32)                        # every candidate ties: index 0 must win
    d1, d2, i1, i2 = _run(fpv, same, same, cuda_dev)
    assert (i1 == 0).all() and (i2 == 0).all() and (d1 == 0).all()


def test_shared_scene_forms_agree(fpv, cuda_dev):
    """[M,3], [1,M,3], a stride-0 expand and a materialised repeat (global_optimization.py:176) all agree."""
    rng = np.random.default_rng(2)
    a = torch.tensor(rng.standard_normal((5, 700, 3)).astype(np.float32), device=cuda_dev)
    s = torch.tensor(rng.standard_normal((3001, 3)).astype(np.float32), device=cuda_dev)
    ref = co.dist_chamfer(a.cpu().numpy(), s.cpu().numpy())
    for form in (s, s.unsqueeze(0), s.unsqueeze(0).expand(5, -1, -1), s.unsqueeze(0).repeat(5, 1, 1)):
        out = [o.cpu().numpy() for o in fpv.distChamfer(a, form)]
        _assert_exact(out, ref)


def test_split_candidate_path_and_int32(fpv, cuda_dev):
    """Few queries x many candidates takes the split + 64-bit atomicMin merge path."""
    rng = np.random.default_rng(3)
    a = rng.standard_normal((1, 300, 3)).astype(np.float32)
    b = rng.standard_normal((1, 200_003, 3)).astype(np.float32)
    b[0, 150_000:150_300] = b[0, 5:305]                              # ties across split boundaries
    got = _run(fpv, a, b, cuda_dev, idx_dtype=torch.int32)
    assert got[2].dtype == np.int32
    _assert_exact(got, co.dist_chamfer(a, b))
    L = fpv._lib.lib()
    try:
        for qpt, ns, packed in [(4, 1, 0), (8, 1, 8), (4, 7, 1), (8, 13, 2), (4, 3, 4), (8, 2, 0), (4, 0, 3), (8, 0, 6)]:
            L.fpv_nn_set_tuning(qpt, ns, packed)
            _assert_exact(_run(fpv, a, b, cuda_dev), co.dist_chamfer(a, b))
    finally:
        L.fpv_nn_set_tuning(0, 0, -1)


def test_far_from_origin_and_special_values(fpv, cuda_dev):
    rng = np.random.default_rng(4)
    a = (100.0 + 0.01 * rng.standard_normal((1, 500, 3))).astype(np.float32)
    d1, d2, i1, i2 = _run(fpv, a, a.copy(), cuda_dev)
    assert (d1 == 0).all() and np.array_equal(i1[0], np.arange(500)) and np.array_equal(i2[0], np.arange(500))
    x = np.array([[[0, 0, 0], [np.nan, 0, 0], [3e38, 3e38, 3e38]]], np.float32)
    y = np.array([[[np.nan, 0, 0], [1, 0, 0], [1, 0, 0], [-3e38, -3e38, -3e38]]], np.float32)
    got = _run(fpv, x, y, cuda_dev)
    want = co.dist_chamfer(x, y)
    assert np.array_equal(got[2], want[2]) and np.array_equal(got[3], want[3])
    assert np.array_equal(got[0], want[0], equal_nan=True) and np.array_equal(got[1], want[1], equal_nan=True)


def test_errors(fpv, cuda_dev):
    a = torch.zeros(2, 4, 3, device=cuda_dev)
    with pytest.raises(RuntimeError):
        fpv.distChamfer(a, torch.zeros(2, 0, 3, device=cuda_dev))
    with pytest.raises(RuntimeError):
        fpv.distChamfer(a, torch.zeros(3, 4, 3, device=cuda_dev))
    with pytest.raises(RuntimeError):
        fpv.distChamfer(a.double(), a.double())
    with pytest.raises(RuntimeError):
        fpv.distChamfer(a, torch.zeros(2, 4, 3))


@pytest.mark.parametrize("shared", [False, True])
def test_backward_matches_oracle_and_is_deterministic(fpv, cuda_dev, shared):
    rng = np.random.default_rng(6)
    bs, N, M = 3, 900, 2500
    a = rng.standard_normal((bs, N, 3)).astype(np.float32)
    b = rng.standard_normal((M, 3) if shared else (bs, M, 3)).astype(np.float32)
    g1 = rng.standard_normal((bs, M)).astype(np.float32)
    g2 = rng.standard_normal((bs, N)).astype(np.float32)
    _, _, i1, i2 = co.dist_chamfer(a, b)
    ga, gb = co.dist_chamfer_bwd(a, b, g1, g2, i1, i2)
    if shared:
        gb = gb.sum(0, keepdims=True)
    grads = []
    for _ in range(2):
        ta = torch.tensor(a, device=cuda_dev, requires_grad=True)
        tb = torch.tensor(b, device=cuda_dev, requires_grad=True)
        d1, d2, _, _ = fpv.distChamfer(ta, tb)
        loss = (d1 * torch.tensor(g1, device=cuda_dev)).sum() + (d2 * torch.tensor(g2, device=cuda_dev)).sum()
        loss.backward(retain_graph=True)                             # as the reference loop does (:591)
        grads.append((ta.grad.cpu().numpy(), tb.grad.cpu().numpy().reshape(gb.shape)))
    scale_a, scale_b = np.abs(ga).max(), np.abs(gb).max()
    np.testing.assert_allclose(grads[0][0], ga, rtol=1e-5, atol=1e-5 * scale_a)
    np.testing.assert_allclose(grads[0][1], gb, rtol=1e-5, atol=1e-5 * scale_b)
    assert np.array_equal(grads[0][0], grads[1][0]) and np.array_equal(grads[0][1], grads[1][1])  # bitwise


def test_backward_single_direction_only(fpv, cuda_dev):
    """The reference loop consumes dist1 only (contact_dist, _ = ..., :293): the unused direction gets no grad."""
    rng = np.random.default_rng(8)
    a = rng.standard_normal((2, 300, 3)).astype(np.float32)
    b = rng.standard_normal((2, 800, 3)).astype(np.float32)
    ta = torch.tensor(a, device=cuda_dev, requires_grad=True)
    dist1, _ = fpv.chamferDist()(ta, torch.tensor(b, device=cuda_dev))
    dist1.sum().backward()
    _, _, i1, i2 = co.dist_chamfer(a, b)
    ga, _ = co.dist_chamfer_bwd(a, b, None, np.ones((2, 300), np.float32), i1, i2)
    np.testing.assert_allclose(ta.grad.cpu().numpy(), ga, rtol=1e-5, atol=1e-5)


def test_golden_gradients(fpv, cuda_dev):
    for case in golden_cases():
        g = load_golden(case)
        ta = torch.tensor(g["a"], device=cuda_dev, requires_grad=True)
        tb = torch.tensor(g["b"], device=cuda_dev, requires_grad=True)
        d1, d2, _, _ = fpv.distChamfer(ta, tb)
        ((d1 * torch.tensor(g["g_b2a"], device=cuda_dev)).sum() + (d2 * torch.tensor(g["g_a2b"], device=cuda_dev)).sum()).backward()
        np.testing.assert_allclose(ta.grad.cpu().numpy(), g["grad_a"], rtol=1e-5, atol=3e-5)
        np.testing.assert_allclose(tb.grad.cpu().numpy(), g["grad_b"], rtol=1e-5, atol=3e-5)


def test_full_size_properties_config2_shapes(fpv, cuda_dev):
    """BASELINE.json config-2 scale in the candidate dimension (1M scene points; 8 frames keep the test
    short): size-independent properties -- self-search is the identity with d == 0; a sampled subset
    agrees with the oracle bit for bit; every reported distance is attained by its reported index."""
    T, V, M = 8, 10475, 1_000_000
    gen = torch.Generator().manual_seed(12)
    scene = torch.rand(M, 3, generator=gen) * torch.tensor([8.0, 8.0, 3.0]) - torch.tensor([4.0, 4.0, 0.0])
    verts = torch.rand(T, V, 3, generator=gen) * torch.tensor([1.0, 1.0, 1.8]) + torch.tensor([-0.5, -0.5, 0.0])
    sd, vd = scene.to(cuda_dev), verts.to(cuda_dev)
    d_b2a, d_a2b, i_b2a, i_a2b = fpv.distChamfer(vd, sd.unsqueeze(0), idx_dtype=torch.int32)
    assert d_b2a.shape == (T, M) and d_a2b.shape == (T, V)
    won = torch.gather(sd.unsqueeze(0).expand(T, -1, -1), 1, i_a2b.long().unsqueeze(-1).expand(-1, -1, 3))
    dx = vd - won
    re = torch.addcmul(torch.addcmul(dx[..., 0] * dx[..., 0], dx[..., 1], dx[..., 1]), dx[..., 2], dx[..., 2])
    assert torch.allclose(re, d_a2b, rtol=1e-6, atol=0)
    # oracle on a sample: 200 body vertices of frame 3 against the full scene, 20k scene points against frame 5
    sel = torch.randperm(V, generator=gen)[:200]
    od, oi = co.nn(verts[3, sel].numpy(), scene.numpy())
    assert np.array_equal(i_a2b[3, sel].cpu().numpy(), oi) and np.array_equal(d_a2b[3, sel].cpu().numpy(), od)
    sel2 = torch.randperm(M, generator=gen)[:20000]
    od2, oi2 = co.nn(scene[sel2].numpy(), verts[5].numpy())
    assert np.array_equal(i_b2a[5, sel2].cpu().numpy(), oi2) and np.array_equal(d_b2a[5, sel2].cpu().numpy(), od2)
    # idempotence: the scene against itself
    planes = fpv.pack_planes(sd)
    d, i = fpv.nn_search(sd[:200_000].unsqueeze(0), planes, M)
    assert (d == 0).all() and torch.equal(i[0].long(), torch.arange(200_000, device=cuda_dev))


def test_keys_and_unpack(fpv, cuda_dev):
    rng = np.random.default_rng(9)
    x = rng.standard_normal((2, 500, 3)).astype(np.float32)
    y = rng.standard_normal((1, 3000, 3)).astype(np.float32)
    planes = fpv.pack_planes(torch.tensor(y, device=cuda_dev))
    keys = fpv.nn_search(torch.tensor(x, device=cuda_dev), planes, 3000, idx_base=1000, want_keys=True)
    d, i = fpv.unpack_keys(keys, torch.int64)
    for s in range(2):
        od, oi = co.nn(x[s], y[0])
        assert np.array_equal(d[s].cpu().numpy(), od) and np.array_equal(i[s].cpu().numpy(), oi + 1000)
    assert int(keys[0, 0]) == co.pack_key(float(d[0, 0]), int(i[0, 0]))


def test_backward_broadcast_weights_equal_materialised(fpv, cuda_dev):
    """sum()/mean() hand the backward a stride-0 expanded scalar: it is passed as one float (fpv_chamfer_bwd_bcast) and
    must give exactly the gradient of the materialised weights."""
    rng = np.random.default_rng(77)
    a = torch.tensor(rng.standard_normal((3, 900, 3)).astype(np.float32), device=cuda_dev)
    b = torch.tensor(rng.standard_normal((1, 5000, 3)).astype(np.float32), device=cuda_dev)
    grads = []
    for materialise in (False, True):
        ar, br = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
        d1, d2, _, _ = fpv.distChamfer(ar, br)
        if materialise:
            loss = (d1 * torch.full_like(d1, 0.3)).sum() + (d2 * torch.full_like(d2, 1.0 / d2.numel())).sum()
        else:
            loss = d1.sum() * 0.3 + d2.mean()
        loss.backward()
        grads.append((ar.grad.clone(), br.grad.clone()))
    assert torch.equal(grads[0][0], grads[1][0]) and torch.equal(grads[0][1], grads[1][1])
