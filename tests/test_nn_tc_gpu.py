"""GPU parity of the tensor-core filtered search (nn_tc_kernel, engine 2): it must return bit-identical
distances and indices to the oracle -- the tensor cores only pre-select candidates for the exact fp32
re-check -- on ragged sizes, ties, far-from-origin data, special values, both chamfer directions and the
candidate-split path."""
import numpy as np
import pytest
import torch

from oracle import chamfer_oracle as co

pytestmark = pytest.mark.gpu


@pytest.fixture()
def tc_engine(fpv):
    L = fpv._lib.lib()
    L.fpv_nn_set_engine(2, 0)
    yield L
    L.fpv_nn_set_engine(0, 0)
    L.fpv_nn_set_tuning(0, 0, -1)


def _run(fpv, a, b, dev, idx_dtype=torch.int64):
    out = fpv.distChamfer(torch.tensor(a, device=dev), torch.tensor(b, device=dev), idx_dtype=idx_dtype)
    return [o.cpu().numpy() for o in out]


def _assert_exact(got, want):
    assert np.array_equal(got[2], want[2]), f"i_b2a mismatches: {(got[2] != want[2]).sum()}"
    assert np.array_equal(got[3], want[3]), f"i_a2b mismatches: {(got[3] != want[3]).sum()}"
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])


@pytest.mark.parametrize("bs,N,M", [(1, 1, 1), (2, 5, 3), (1, 7, 1025), (3, 1000, 999), (2, 2049, 17), (1, 33, 8193),
                                    (4, 257, 4100), (1, 10475, 20000), (2, 513, 256), (1, 128, 255)])
def test_tc_parity_ragged(fpv, cuda_dev, tc_engine, bs, N, M):
    rng = np.random.default_rng(bs * 7 + N * 13 + M)
    a = (rng.standard_normal((bs, N, 3)) * 2).astype(np.float32)
    b = (rng.standard_normal((bs, M, 3)) * 2).astype(np.float32)
    _assert_exact(_run(fpv, a, b, cuda_dev), co.dist_chamfer(a, b))


def test_tc_ties_and_duplicates(fpv, cuda_dev, tc_engine):
    rng = np.random.default_rng(1)
    a = rng.integers(-8, 9, (2, 3000, 3)).astype(np.float32)
    b = rng.integers(-8, 9, (2, 5000, 3)).astype(np.float32)
    _assert_exact(_run(fpv, a, b, cuda_dev), co.dist_chamfer(a, b))
    same = np.zeros((1, 700, 3), np.float32)
    d1, d2, i1, i2 = _run(fpv, same, same, cuda_dev)
    assert (i1 == 0).all() and (i2 == 0).all() and (d1 == 0).all()
    b2 = rng.standard_normal((1, 6000, 3)).astype(np.float32)
    b2[0, 3000:] = b2[0, :3000]                       # every point duplicated 3000 indices later
    a2 = b2[:, ::7].copy()
    _assert_exact(_run(fpv, a2, b2, cuda_dev), co.dist_chamfer(a2, b2))


def test_tc_far_from_origin_and_special_values(fpv, cuda_dev, tc_engine):
    rng = np.random.default_rng(4)
    a = (100.0 + 0.01 * rng.standard_normal((1, 900, 3))).astype(np.float32)   # filter nearly useless, still exact
    b = (100.0 + 0.01 * rng.standard_normal((1, 1300, 3))).astype(np.float32)
    _assert_exact(_run(fpv, a, b, cuda_dev), co.dist_chamfer(a, b))
    x = np.zeros((1, 300, 3), np.float32)
    x[0, :, 0] = np.arange(300)
    x[0, 1] = [np.nan, 0, 0]
    x[0, 2] = [3e38, 3e38, 3e38]
    x[0, 3] = [1e19, 1e19, 0]
    x[0, 4] = [np.inf, 0, 0]
    y = np.zeros((1, 400, 3), np.float32)
    y[0, :, 1] = np.arange(400) * 0.5
    y[0, 0] = [np.nan, 0, 0]
    y[0, 7] = [-3e38, -3e38, -3e38]
    y[0, 9] = [1e19, 1e19, 1.0]
    y[0, 11] = [np.inf, np.inf, 0]
    got = _run(fpv, x, y, cuda_dev)
    want = co.dist_chamfer(x, y)
    assert np.array_equal(got[2], want[2]) and np.array_equal(got[3], want[3])
    assert np.array_equal(got[0], want[0], equal_nan=True) and np.array_equal(got[1], want[1], equal_nan=True)


def test_tc_split_keys_and_shared_scene(fpv, cuda_dev, tc_engine):
    rng = np.random.default_rng(3)
    a = rng.standard_normal((3, 700, 3)).astype(np.float32)
    s = rng.standard_normal((100_003, 3)).astype(np.float32)
    s[60_000:60_300] = s[5:305]
    want = co.dist_chamfer(a, s)
    for ns in (0, 1, 5, 13):
        tc_engine.fpv_nn_set_tuning(0, ns, -1)
        _assert_exact(_run(fpv, a, s[None], cuda_dev, torch.int32), want)
    tc_engine.fpv_nn_set_tuning(0, 0, -1)
    planes = fpv.pack_planes(torch.tensor(s, device=cuda_dev))
    keys = fpv.nn_search(torch.tensor(a, device=cuda_dev), planes, s.shape[0], idx_base=1000, want_keys=True)
    d, i = fpv.unpack_keys(keys, torch.int64)
    assert np.array_equal(d.cpu().numpy(), want[1]) and np.array_equal(i.cpu().numpy(), want[3] + 1000)


def test_tc_equals_simt_on_room_scale_data(fpv, cuda_dev, tc_engine):
    """Config-2-like geometry (room box, body-sized query cloud), both directions, larger sizes; also sweeps the
    error-bound exponent: the default (15) must be exact; tighter bounds are reported to document the margin."""
    gen = torch.Generator().manual_seed(12)
    T, V, M = 3, 10475, 300_000
    scene = torch.rand(M, 3, generator=gen) * torch.tensor([8.0, 8.0, 3.0]) - torch.tensor([4.0, 4.0, 0.0])
    verts = torch.rand(T, V, 3, generator=gen) * torch.tensor([0.6, 0.6, 1.8]) + torch.tensor([0.5, -1.0, 0.0])
    sd, vd = scene.to(cuda_dev), verts.to(cuda_dev)
    tc_engine.fpv_nn_set_engine(1, 0)
    ref = [o.clone() for o in fpv.distChamfer(vd, sd.unsqueeze(0), idx_dtype=torch.int32)]
    margins = {}
    for e in (15, 17, 19, 21, 23):
        tc_engine.fpv_nn_set_engine(2, e)
        out = fpv.distChamfer(vd, sd.unsqueeze(0), idx_dtype=torch.int32)
        bad = sum(int((o != r).sum()) for o, r in zip(out, ref))
        margins[e] = bad
    print("tc filter mismatches by error-bound exponent:", margins)
    assert margins[15] == 0 and margins[17] == 0
