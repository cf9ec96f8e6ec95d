"""CPU: the on-disk formats of SURVEY.md section 8f row f4 (io_formats.py) against the reference's own conventions."""
import importlib
import os
import pickle
import sys

import numpy as np
import pytest
import torch

io = importlib.import_module("4dcapture-fpv_b200.io_formats")


def test_qvec2rotmat_convention_and_camerapose_round_trip(tmp_path):
    # hand-checked: 90 degrees about z, COLMAP (w,x,y,z)
    R = io.qvec2rotmat([np.sqrt(0.5), 0, 0, np.sqrt(0.5)])
    np.testing.assert_allclose(R, [[0, -1, 0], [1, 0, 0], [0, 0, 1]], atol=1e-12)
    g = torch.Generator().manual_seed(0)
    q = torch.randn(9, 4, generator=g, dtype=torch.float64)
    q = q / q.norm(dim=1, keepdim=True)
    t = torch.randn(9, 3, generator=g, dtype=torch.float64)
    p = tmp_path / "camerapose.txt"
    with open(p, "w") as f:
        for i in range(9):
            f.write(" ".join([f"img{i}.jpg"] + [repr(float(v)) for v in q[i]] + [repr(float(v)) for v in t[i]]) + "\n")
    ext = io.read_camerapose(str(p))
    assert ext.shape == (9, 4, 4) and ext.dtype == torch.float32
    for i in range(9):                       # extract_ext: inv([R|t])
        m = np.eye(4)
        m[:3, :3] = io.qvec2rotmat(q[i].numpy())
        m[:3, 3] = t[i].numpy()
        np.testing.assert_allclose(ext[i].numpy(), np.linalg.inv(m), atol=1e-6)
    p2 = tmp_path / "again.txt"
    io.write_camerapose(str(p2), ext)
    np.testing.assert_allclose(io.read_camerapose(str(p2)).numpy(), ext.numpy(), atol=2e-6)
    with pytest.raises(ValueError):
        (tmp_path / "bad.txt").write_text("x 1 0 0\n")
        io.read_camerapose(str(tmp_path / "bad.txt"))


@pytest.mark.parametrize("binary", [True, False])
def test_ply_vertices_round_trip(tmp_path, binary):
    g = torch.Generator().manual_seed(1)
    pts = torch.randn(1000, 3, generator=g)
    p = tmp_path / "scene.ply"
    io.write_ply_vertices(str(p), pts, binary=binary)
    back = io.read_ply_vertices(str(p))
    assert back.dtype == torch.float32 and back.shape == (1000, 3)
    assert torch.equal(back, pts) if binary else torch.allclose(back, pts, rtol=1e-7, atol=0)


def test_ply_with_extra_properties_and_faces(tmp_path):
    # a mesh as scanners write it: normals + colours on the vertices, then a face list
    n = 5
    dt = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("nx", "<f4"), ("ny", "<f4"), ("nz", "<f4"),
                   ("red", "u1"), ("green", "u1"), ("blue", "u1")])
    rec = np.zeros(n, dt)
    rec["x"], rec["y"], rec["z"] = np.arange(n), np.arange(n) * 2, np.arange(n) * 3
    hdr = ("ply\nformat binary_little_endian 1.0\ncomment made by a test\nelement vertex 5\nproperty float x\nproperty float y\n"
           "property float z\nproperty float nx\nproperty float ny\nproperty float nz\nproperty uchar red\nproperty uchar green\n"
           "property uchar blue\nelement face 1\nproperty list uchar int vertex_indices\nend_header\n")
    p = tmp_path / "mesh.ply"
    with open(p, "wb") as f:
        f.write(hdr.encode())
        f.write(rec.tobytes())
        f.write(bytes([3]) + np.array([0, 1, 2], "<i4").tobytes())
    v = io.read_ply_vertices(str(p))
    np.testing.assert_array_equal(v.numpy(), np.stack([np.arange(n), np.arange(n) * 2, np.arange(n) * 3], 1).astype(np.float32))
    with pytest.raises(ValueError):
        (tmp_path / "no.ply").write_text("obj\n")
        io.read_ply_vertices(str(tmp_path / "no.ply"))


def test_result_pickles_round_trip_and_reader_keys(tmp_path):
    g = torch.Generator().manual_seed(2)
    T = 7
    rec = torch.randn(T, 75, generator=g)
    ext = torch.eye(4).repeat(T, 1, 1) + 0.01 * torch.randn(T, 4, 4, generator=g)
    names = io.save_result(rec, torch.tensor([1.25]), ext, str(tmp_path / "fit"))
    assert [os.path.basename(n) for n in names[:2]] == ["body_gen_000000.pkl", "body_gen_000001.pkl"]
    with open(names[3], "rb") as f:
        d = pickle.load(f)
    # the keys global_vis.py:116-124 reads, with the per-frame [1,D] shapes of cvae.py:196-202
    assert d["betas"].shape == (1, 10) and d["body_pose"].shape == (1, 32) and d["camera_translation"].shape == (1, 3)
    assert d["scale"] == pytest.approx(1.25) and d["camera_ext"].shape == (4, 4)
    np.testing.assert_array_equal(d["transl"], rec[3:4, 0:3].numpy())
    r2, s2, e2 = io.load_result(str(tmp_path / "fit"))
    assert torch.equal(r2, rec) and s2 == pytest.approx(1.25) and torch.equal(e2, ext)
    # the SMPLify-X input side: same dict keys -> body_params_parse order (global_optimization.py:64-76)
    rows = io.load_smplifyx_results(str(tmp_path / "fit" / "*.pkl"))
    assert torch.equal(rows, rec)
    with pytest.raises(FileNotFoundError):
        io.load_smplifyx_results(str(tmp_path / "nothing" / "*.pkl"))


def test_formats_match_reference_functions():
    """qvec2rotmat and body_params_parse against outputs of the reference's own functions
    (tests/golden/prior_formats.npz, made by tests/golden/make_golden_prior.py)."""
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "prior_formats.npz"))
    for q, R in zip(d["q"], d["Rq"]):
        assert np.array_equal(io.qvec2rotmat(q), R)
    frame = {k[len("frame_"):]: d[k] for k in d.files if k.startswith("frame_")}
    assert np.array_equal(io.body_params_parse(frame), d["row"])


def test_load_smplx_npz_v10_and_v11_layouts(tmp_path):
    """body_model.load_smplx_npz reads both shapedirs layouts of the real model files: 400 columns (v1.1: expression
    directions from column 300) and 20 columns (v1.0: 10 shape + 10 expression, what the smplx package falls back to)."""
    import numpy as np
    import torch
    from conftest import load_pkg
    fpv = load_pkg()
    V = 64
    c = fpv.synthetic.make_body_constants(5, V)
    rng = np.random.default_rng(0)
    base = dict(v_template=c["v_template"].numpy(), posedirs=c["posedirs"].numpy().T.reshape(V, 3, 486),
                J_regressor=c["J_regressor"].numpy(), weights=c["lbs_weights"].numpy(),
                kintree_table=np.stack([np.array([2 ** 32 - 1] + c["parents"].tolist()[1:]), np.arange(55)]).astype(np.int64),
                hands_componentsl=rng.standard_normal((45, 45)), hands_componentsr=rng.standard_normal((45, 45)),
                hands_meanl=rng.standard_normal(45), hands_meanr=rng.standard_normal(45))
    sd400 = rng.standard_normal((V, 3, 400))
    sd20 = np.concatenate([sd400[:, :, :10], sd400[:, :, 300:310]], -1)
    out = {}
    for tag, sd in (("v11", sd400), ("v10", sd20)):
        path = tmp_path / f"SMPLX_{tag}.npz"
        np.savez(path, shapedirs=sd, **base)
        k = fpv.load_smplx_npz(str(path))
        assert k["shapedirs"].shape == (V, 3, 20) and k["posedirs"].shape == (486, 3 * V)
        assert k["lh_components"].shape == (12, 45) and int(k["parents"][0]) == -1
        out[tag] = k
        fpv.create(str(path), batch_size=2)          # the module accepts the constants (CPU construction only)
    assert torch.equal(out["v10"]["shapedirs"], out["v11"]["shapedirs"])
    np.testing.assert_allclose(out["v11"]["shapedirs"][:, :, 10:].numpy(), sd400[:, :, 300:310].astype(np.float32))
