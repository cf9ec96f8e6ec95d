"""CPU: the C-ABI library loads, exports every symbol include/fpv_b200.h declares, answers its size
queries without a GPU, and the product refuses CPU tensors instead of falling back."""
import numpy as np
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT, load_pkg

fpv = load_pkg()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fpv_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fpv_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    L = fpv._lib.lib()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/fpv_b200.h but not exported"
        assert n in fpv._lib.SIGNATURES, f"{n} has no ctypes signature in _lib.py"
    assert L.fpv_abi_version() == 1


def test_struct_layout_matches_header():
    # 4 int32 + 16 pointers
    assert ctypes.sizeof(fpv._lib.SmplxModelStruct) == 16 + 16 * 8


def test_size_queries_run_without_gpu():
    L = fpv._lib.lib()
    assert L.fpv_nn_planes_bytes(1, 1000) == 3 * 1024 * 4
    assert L.fpv_nn_planes_bytes(0, 5) == 0
    # config 2: planes for 300 x 10,475 body vertices + 1M scene points
    ws = L.fpv_chamfer_fwd_workspace_bytes(300, 10475, 1_000_000, 1)
    assert 49_000_000 < ws < 120_000_000
    assert L.fpv_chamfer_bwd_workspace_bytes(300, 10475, 1_000_000, 1, 0) >= 300 * 10475 * 3 * 8
    assert L.fpv_transform_bwd_workspace_bytes(300, 10475) > 0
    assert L.fpv_reduce_workspace_bytes(10) > 0


def test_argument_errors_are_reported_not_crashed():
    L = fpv._lib.lib()
    rc = L.fpv_chamfer_fwd(None, None, 1, 1, 1, 0, None, None, None, None, 8, None, 0, None)
    assert rc != 0 and b"null" in L.fpv_last_error()
    rc = L.fpv_tdiff_l1_fwd(ctypes.c_void_p(256), 2, 4, 2, None, ctypes.c_void_p(256), None, 0, None)
    assert rc != 0 and b"T > order" in L.fpv_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback():
    a = torch.zeros(1, 4, 3)
    with pytest.raises(RuntimeError, match="no CPU"):
        fpv.distChamfer(a, a)
    with pytest.raises(RuntimeError, match="no CPU"):
        fpv.second_diff_l1(torch.zeros(5, 3))
    m = fpv.create(constants=fpv.synthetic.make_body_constants(1, 200), batch_size=2)
    with pytest.raises(RuntimeError, match="no CPU"):
        m(betas=torch.zeros(2, 10))
    L = fpv._lib.lib()
    assert L.fpv_device_query(0, None, None, None) != 0


def test_body_model_constant_preparation():
    c = fpv.synthetic.make_body_constants(3, 500)
    m = fpv.create(constants=c, batch_size=2)
    V = 500
    assert m.basis_kn.shape == (512, 3 * V) and m.ell_width == 4
    torch.testing.assert_close(m.basis_kn[:486], c["posedirs"])
    torch.testing.assert_close(m.basis_kn[506].view(V, 3), c["v_template"])
    torch.testing.assert_close(m.basis_kn[486:506].view(20, V, 3).permute(1, 2, 0), c["shapedirs"])
    assert float(m.basis_kn[507:].abs().max()) == 0.0
    # ELL and the per-joint lists are two views of the same sparse matrix
    W = torch.zeros(V, 55)
    for w in range(m.ell_width):
        j = m.ell_joint[w].long()
        ok = j >= 0
        W[torch.nonzero(ok).squeeze(1), j[ok]] = m.ell_weight[w][ok]
    torch.testing.assert_close(W, c["lbs_weights"])
    W2 = torch.zeros(V, 55)
    for j in range(55):
        s, e = int(m.csr_ptr[j]), int(m.csr_ptr[j + 1])
        vs = m.csr_vert[s:e].long()
        assert (vs[1:] > vs[:-1]).all()
        W2[vs, j] = m.csr_weight[s:e]
    torch.testing.assert_close(W2, c["lbs_weights"])
    torch.testing.assert_close(m.j_template, (c["J_regressor"].double() @ c["v_template"].double()).float())


def test_shard_ranges_partition_the_scene():
    for M, G in [(10, 3), (1_000_000, 8), (7, 8), (20_000_000, 8)]:
        r = [fpv.shard_range(M, G, k) for k in range(G)]
        assert r[0][0] == 0 and r[-1][1] == M and all(r[k][1] == r[k + 1][0] for k in range(G - 1))
        sizes = [e - b for b, e in r]
        assert max(sizes) - min(sizes) <= 1


def test_morton_keys_and_scene_presort_host_logic():
    """The host-side ordering helpers (pure torch, CPU): bit interleave against a plain Python reference, and the
    pre-sorted scene is a permutation of the input that is sorted by its own keys."""
    import importlib
    sp = importlib.import_module("4dcapture-fpv_b200.spatial")
    fit = importlib.import_module("4dcapture-fpv_b200.fit")
    g = torch.Generator().manual_seed(3)
    pts = torch.rand(500, 3, generator=g) * torch.tensor([8.0, 8.0, 3.0]) - torch.tensor([4.0, 4.0, 0.0])
    pts[7] = torch.tensor([float("nan"), 0.0, 0.0])
    pts[9] = torch.tensor([float("inf"), -float("inf"), 1.0])
    lo, inv = sp.grid_of(pts)
    keys = sp.morton_keys(pts, lo, inv)

    def ref_key(p):
        q = []
        for a in range(3):
            v = (float(p[a]) - float(lo[a])) * float(inv[a])
            v = np.float32(p[a].item() - lo[a].item()) * np.float32(inv[a].item()) if np.isfinite(p[a].item()) else v
            if v != v or v == float("inf"):
                v = 1023.0
            if v == -float("inf"):
                v = 0.0
            q.append(int(min(max(float(v), 0.0), 1023.0)))
        k = 0
        for bit in range(10):
            for a in range(3):
                k |= ((q[a] >> bit) & 1) << (3 * bit + a)
        return k

    assert [int(k) for k in keys] == [ref_key(p) for p in pts]
    clean = pts[torch.isfinite(pts).all(dim=1)]
    srt = fit._morton_sorted(clean)
    assert sorted(map(tuple, srt.tolist())) == sorted(map(tuple, clean.tolist()))
    lo2, inv2 = sp.grid_of(srt)
    k2 = sp.morton_keys(srt, lo2, inv2)
    assert bool((k2[1:] >= k2[:-1]).all())
