"""CPU: pin the chamfer oracle against the golden vectors generated from the REAL reference
(tests/golden/make_golden.py imports /root/reference/chamfer_python.py) and check its own invariants."""
import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from conftest import golden_cases, load_golden
from oracle import chamfer_oracle as co
from oracle import chamfer_ref_port as port

EXACT = [c for c in golden_cases() if not c.startswith("random")]
RANDOM = [c for c in golden_cases() if c.startswith("random")]


@pytest.mark.parametrize("case", EXACT)
def test_oracle_bit_exact_on_lattice_golden(case):
    """Lattice / dyadic inputs: every fp32 form of the distance is exact, so the canonical oracle must
    reproduce the literal reference bit for bit -- distances, indices (ties included) and gradients."""
    g = load_golden(case)
    d1, d2, i1, i2 = co.dist_chamfer(g["a"], g["b"])
    assert np.array_equal(i1, g["i_b2a"]) and np.array_equal(i2, g["i_a2b"])
    assert np.array_equal(d1, g["d_b2a"]) and np.array_equal(d2, g["d_a2b"])
    ga, gb = co.dist_chamfer_bwd(g["a"], g["b"], g["g_b2a"], g["g_a2b"], i1, i2)
    assert np.array_equal(ga.astype(np.float32), g["grad_a"])
    assert np.array_equal(gb.astype(np.float32), g["grad_b"])


@pytest.mark.parametrize("case", RANDOM)
def test_oracle_matches_reference_on_random_golden(case):
    """Unit-cube floats: indices identical (NN margins exceed the expanded form's error here),
    distances and gradients within 1e-5 relative of the cloud scale (tolerance stated by north_star)."""
    g = load_golden(case)
    d1, d2, i1, i2 = co.dist_chamfer(g["a"], g["b"])
    assert np.array_equal(i1, g["i_b2a"]) and np.array_equal(i2, g["i_a2b"])
    np.testing.assert_allclose(d1, g["d_b2a"], rtol=1e-5, atol=1e-5 * 3.0)
    np.testing.assert_allclose(d2, g["d_a2b"], rtol=1e-5, atol=1e-5 * 3.0)
    ga, gb = co.dist_chamfer_bwd(g["a"], g["b"], g["g_b2a"], g["g_a2b"], i1, i2)
    np.testing.assert_allclose(ga, g["grad_a"], rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(gb, g["grad_b"], rtol=1e-5, atol=2e-5)


def test_probe4_known_answer():
    """Hand-computed case of SURVEY.md section 8c."""
    g = load_golden("probe4")
    d1, d2, i1, i2 = co.dist_chamfer(g["a"], g["b"])
    assert i2.tolist() == [[0, 0, 0, 0]] and d2.tolist() == [[1.0, 66.0, 0.0, 226.0]]
    assert i1.tolist() == [[2, 2, 0, 2]] and d1.tolist() == [[0.0, 0.0, 1.0, 0.0]]


@pytest.mark.parametrize("case", EXACT)
def test_port_matches_reference_on_lattice(case):
    """The tiled torch restatement (the CPU-baseline arithmetic) equals the literal reference on exact inputs."""
    g = load_golden(case)
    d1, d2, i1, i2 = port.distChamfer(torch.tensor(g["a"]), torch.tensor(g["b"]), tile=100)
    assert np.array_equal(i1.numpy(), g["i_b2a"]) and np.array_equal(i2.numpy(), g["i_a2b"])
    assert np.array_equal(d1.numpy(), g["d_b2a"]) and np.array_equal(d2.numpy(), g["d_a2b"])


@pytest.mark.parametrize("N,M", [(1, 1), (3, 1000), (517, 33), (64, 513), (1000, 1537)])
def test_blocked_equals_naive(N, M):
    rng = np.random.default_rng(N * 7919 + M)
    x = rng.standard_normal((N, 3)).astype(np.float32)
    y = rng.standard_normal((M, 3)).astype(np.float32)
    y[M // 2:] = y[: M - M // 2]  # duplicates -> ties
    d, i = co.nn(x, y)
    dn, i_n = co.nn(x, y, naive=True)
    assert np.array_equal(i, i_n) and np.array_equal(d, dn)


def test_n_neq_m_and_shared_scene():
    rng = np.random.default_rng(5)
    a = rng.standard_normal((3, 40, 3)).astype(np.float32)
    b = rng.standard_normal((70, 3)).astype(np.float32)
    d1, d2, i1, i2 = co.dist_chamfer(a, b)
    rep = np.repeat(b[None], 3, 0)
    e1, e2, j1, j2 = co.dist_chamfer(a, rep)
    assert d1.shape == (3, 70) and d2.shape == (3, 40)
    assert np.array_equal(d1, e1) and np.array_equal(i2, j2) and np.array_equal(i1, j1) and np.array_equal(d2, e2)


def test_far_from_origin_stays_non_negative_and_exact():
    """Where the reference's expanded form goes negative, the canonical form does not."""
    rng = np.random.default_rng(11)
    a = (100.0 + 0.01 * rng.standard_normal((1, 200, 3))).astype(np.float32)
    d1, d2, i1, i2 = co.dist_chamfer(a, a.copy())
    assert (d1 >= 0).all() and (d2 >= 0).all()
    assert np.array_equal(i1[0], np.arange(200)) and np.array_equal(i2[0], np.arange(200))
    assert (d1 == 0).all()


def test_nan_and_empty():
    x = np.array([[0, 0, 0], [np.nan, 0, 0]], np.float32)
    y = np.array([[np.nan, 0, 0], [1, 0, 0], [1, 0, 0]], np.float32)
    d, i = co.nn(x, y)
    assert i.tolist() == [1, 0] and d[0] == 1.0 and np.isinf(d[1])
    with pytest.raises(ValueError):
        co.nn(x, np.zeros((0, 3), np.float32))


def test_gradient_against_float64_autograd():
    rng = np.random.default_rng(3)
    a = rng.standard_normal((2, 50, 3)).astype(np.float32)
    b = rng.standard_normal((2, 80, 3)).astype(np.float32)
    g1 = rng.standard_normal((2, 80)).astype(np.float32)
    g2 = rng.standard_normal((2, 50)).astype(np.float32)
    d1, d2, i1, i2 = co.dist_chamfer(a, b)
    ga, gb = co.dist_chamfer_bwd(a, b, g1, g2, i1, i2)
    ta = torch.tensor(a, dtype=torch.float64, requires_grad=True)
    tb = torch.tensor(b, dtype=torch.float64, requires_grad=True)
    P = ((ta.unsqueeze(2) - tb.unsqueeze(1)) ** 2).sum(-1)
    loss = (P.min(1)[0] * torch.tensor(g1, dtype=torch.float64)).sum() + (P.min(2)[0] * torch.tensor(g2, dtype=torch.float64)).sum()
    loss.backward()
    np.testing.assert_allclose(ga, ta.grad.numpy(), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(gb, tb.grad.numpy(), rtol=1e-6, atol=1e-6)


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 60), st.integers(1, 90), st.integers(1, 5), st.integers(0, 2 ** 31 - 1))
def test_sharded_min_equals_unsharded(N, M, shards, seed):
    """Property: for ANY contiguous partition of the scene, the integer minimum of
    the packed (distance, global index) keys equals the unsharded lexicographic minimum."""
    rng = np.random.default_rng(seed)
    x = rng.integers(-3, 4, (N, 3)).astype(np.float32)
    y = rng.integers(-3, 4, (M, 3)).astype(np.float32)  # lattice: many exact ties
    d, i = co.nn(x, y)
    cuts = sorted(set([0, M] + rng.integers(0, M + 1, shards).tolist()))
    best = np.full(N, np.iinfo(np.uint64).max, np.uint64)
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        ds, is_ = co.nn(x, y[lo:hi])
        keys = (ds.view(np.uint32).astype(np.uint64) << np.uint64(32)) | (is_.astype(np.uint64) + np.uint64(lo))
        best = np.minimum(best, keys)
    assert np.array_equal((best >> np.uint64(32)).astype(np.uint32).view(np.float32), d)
    assert np.array_equal((best & np.uint64(0xFFFFFFFF)).astype(np.int32), i)
    assert co.pack_key(float(d[0]), int(i[0])) == int(best[0])


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 40), st.integers(2, 60), st.integers(0, 2 ** 31 - 1))
def test_permutation_consistency(N, M, seed):
    """Permuting the scene permutes the winners: distances unchanged, and the winner under the
    permutation is the lowest permuted index among the points at the minimum distance."""
    rng = np.random.default_rng(seed)
    x = rng.integers(-2, 3, (N, 3)).astype(np.float32)
    y = rng.integers(-2, 3, (M, 3)).astype(np.float32)
    perm = rng.permutation(M)
    d, i = co.nn(x, y)
    dp, ip = co.nn(x, y[perm])
    assert np.array_equal(d, dp)
    full = ((x[:, None, :] - y[perm][None]) ** 2).sum(-1)
    assert np.array_equal(ip, full.argmin(1))
