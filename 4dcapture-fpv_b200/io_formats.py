"""On-disk formats either side of the fit loop, host-side Python like the reference's.

    load_smplifyx_results(pattern)     the SMPLify-X per-frame pickles -> [T,75] rows
                                       (global_optimization.py:64-76 body_params_parse, :688-705)
    read_camerapose / write_camerapose COLMAP `camerapose.txt` lines  `name qw qx qy qz tx ty tz`  ->  camera_ext [T,4,4]
                                       = inv([R(q) | t])  (:51-61 qvec2rotmat, :208-230 extract_ext)
    read_ply_vertices                  scene mesh vertices (replaces open3d.io.read_triangle_mesh(...).vertices, :173-175);
                                       ascii and binary little/big-endian PLY
    save_result / load_result          per-frame `body_gen_%06d.pkl` (:637-653) with the keys the readers expect
                                       (global_vis.py:116-124): the seven body_params_encapsulate keys (cvae.py:189-208)
                                       + `scale` + `camera_ext`
No device code here: these run once per clip, outside the step.
"""
from __future__ import annotations

import glob
import os
import pickle
from typing import Dict, List, Sequence

import numpy as np
import torch

PARAM_KEYS = ["transl", "global_orient", "betas", "body_pose", "left_hand_pose", "right_hand_pose", "camera_translation"]
PARAM_SLICES = {"transl": (0, 3), "global_orient": (3, 6), "betas": (6, 16), "body_pose": (16, 48),
                "left_hand_pose": (48, 60), "right_hand_pose": (60, 72), "camera_translation": (72, 75)}


def qvec2rotmat(qvec: Sequence[float]) -> np.ndarray:
    """COLMAP quaternion (w, x, y, z) -> rotation matrix (global_optimization.py:51-61)."""
    w, x, y, z = (float(v) for v in qvec)
    return np.array([[1 - 2 * y * y - 2 * z * z, 2 * x * y - 2 * w * z, 2 * z * x + 2 * w * y],
                     [2 * x * y + 2 * w * z, 1 - 2 * x * x - 2 * z * z, 2 * y * z - 2 * w * x],
                     [2 * z * x - 2 * w * y, 2 * y * z + 2 * w * x, 1 - 2 * x * x - 2 * y * y]])


def rotmat2qvec(R: np.ndarray) -> np.ndarray:
    """Inverse of qvec2rotmat (w >= 0), used by write_camerapose."""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = [0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s]
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
        q = [0.0, 0.0, 0.0, 0.0]
        q[0] = (R[k, j] - R[j, k]) / s
        q[1 + i] = 0.25 * s
        q[1 + j] = (R[j, i] + R[i, j]) / s
        q[1 + k] = (R[k, i] + R[i, k]) / s
    q = np.asarray(q, dtype=np.float64)
    return q if q[0] >= 0 else -q


def read_camerapose(path: str) -> torch.Tensor:
    """camera_ext [T,4,4] float32: per line, the inverse of the world-to-camera [R(q) | t] (extract_ext, :208-230)."""
    mats = []
    with open(path) as f:
        for ln, line in enumerate(f):
            line = line.rstrip("\n")
            if not line:
                continue
            items = line.split(" ")
            if len(items) < 8:
                raise ValueError(f"{path}:{ln + 1}: expected `name qw qx qy qz tx ty tz`, got {len(items)} fields")
            m = np.eye(4)
            m[:3, :3] = qvec2rotmat([float(v) for v in items[1:5]])
            m[:3, 3] = [float(v) for v in items[5:8]]
            mats.append(np.linalg.inv(m))
    if not mats:
        raise ValueError(f"{path}: no camera poses")
    return torch.tensor(np.stack(mats), dtype=torch.float32)


def write_camerapose(path: str, camera_ext: torch.Tensor, names: Sequence[str] = None) -> None:
    """Inverse of read_camerapose: camera_ext [T,4,4] (camera-to-world) -> COLMAP lines."""
    ext = camera_ext.detach().cpu().double().numpy()
    with open(path, "w") as f:
        for i, m in enumerate(ext):
            w2c = np.linalg.inv(m)
            q, t = rotmat2qvec(w2c[:3, :3]), w2c[:3, 3]
            name = names[i] if names is not None else f"frame_{i:06d}.jpg"
            f.write(" ".join([name] + [repr(float(v)) for v in list(q) + list(t)]) + "\n")


_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2",
              "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4",
              "double": "f8", "float64": "f8"}


def read_ply_vertices(path: str) -> torch.Tensor:
    """[M,3] float32 vertex positions of a PLY file (ascii, binary_little_endian or binary_big_endian)."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, nvert, props, in_vertex = None, None, [], False
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: truncated PLY header")
            tok = line.decode("ascii", "replace").strip().split()
            if not tok:
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                in_vertex = tok[1] == "vertex"
                if in_vertex:
                    nvert = int(tok[2])
                elif nvert is None:
                    raise ValueError(f"{path}: an element precedes `vertex`; unsupported layout")
            elif tok[0] == "property" and in_vertex:
                if tok[1] == "list":
                    raise ValueError(f"{path}: list property on the vertex element")
                props.append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if fmt is None or nvert is None:
            raise ValueError(f"{path}: PLY header lacks format / vertex element")
        names = [p[0] for p in props]
        if not all(k in names for k in "xyz"):
            raise ValueError(f"{path}: vertex element has no x/y/z")
        if fmt == "ascii":
            rows = np.loadtxt(f, max_rows=nvert, ndmin=2, dtype=np.float64) if nvert else np.zeros((0, len(props)))
            xyz = rows[:, [names.index(k) for k in "xyz"]]
        else:
            end = "<" if fmt == "binary_little_endian" else ">"
            dt = np.dtype([(n, end + t) for n, t in props])
            rec = np.frombuffer(f.read(dt.itemsize * nvert), dtype=dt, count=nvert)
            xyz = np.stack([rec[k] for k in "xyz"], axis=1)
    return torch.tensor(np.ascontiguousarray(xyz), dtype=torch.float32)


def write_ply_vertices(path: str, points: torch.Tensor, binary: bool = True) -> None:
    p = points.detach().cpu().to(torch.float32).numpy().reshape(-1, 3)
    hdr = "ply\nformat {} 1.0\nelement vertex {}\nproperty float x\nproperty float y\nproperty float z\nend_header\n".format(
        "binary_little_endian" if binary else "ascii", p.shape[0])
    with open(path, "wb") as f:
        f.write(hdr.encode("ascii"))
        if binary:
            f.write(p.astype("<f4").tobytes())
        else:
            for r in p:
                f.write(("%.9g %.9g %.9g\n" % tuple(r)).encode("ascii"))


def body_params_parse(body_params: Dict[str, np.ndarray]) -> np.ndarray:
    """One SMPLify-X pickle -> [1,75] row (global_optimization.py:64-76)."""
    return np.concatenate([np.asarray(body_params[k]).reshape(1, -1) for k in PARAM_KEYS], axis=-1)


def load_smplifyx_results(pattern: str) -> torch.Tensor:
    """sorted(glob(pattern)) pickles -> [T,75] float32 (the reference's main block, :688-705)."""
    files = sorted(glob.glob(pattern))
    if not files:
        raise FileNotFoundError(f"no result pickles match {pattern}")
    rows = []
    for fn in files:
        with open(fn, "rb") as f:
            rows.append(body_params_parse(pickle.load(f, encoding="latin1")))
    data = np.vstack(rows)
    if data.shape[1] != 75:
        raise ValueError(f"expected 75 parameters per frame, got {data.shape[1]}")
    return torch.tensor(data, dtype=torch.float32)


def body_params_encapsulate(body_rec: torch.Tensor, scale, camera_ext: torch.Tensor) -> List[Dict[str, np.ndarray]]:
    """[T,75] rows -> per-frame dicts: the cvae.py:189-208 keys as [1,D] arrays + `scale` + `camera_ext` [4,4]
    (what save_result passes, :644, and what global_vis.py:116-124 reads back)."""
    rec = body_rec.detach().cpu().numpy()
    ext = camera_ext.detach().cpu().numpy()
    s = float(np.asarray(scale.detach().cpu() if torch.is_tensor(scale) else scale).squeeze())
    out = []
    for b in range(rec.shape[0]):
        d = {k: rec[b:b + 1, lo:hi] for k, (lo, hi) in PARAM_SLICES.items()}
        d["scale"] = s
        d["camera_ext"] = ext[b] if ext.ndim == 3 else ext
        out.append(d)
    return out


def save_result(body_rec: torch.Tensor, scale, camera_ext: torch.Tensor, fit_path: str) -> List[str]:
    """FittingOP.save_result (:637-653): fit_path/body_gen_%06d.pkl, one per frame."""
    os.makedirs(fit_path, exist_ok=True)
    names = []
    for i, d in enumerate(body_params_encapsulate(body_rec, scale, camera_ext)):
        fn = os.path.join(fit_path, "body_gen_" + str(i).zfill(6) + ".pkl")
        with open(fn, "wb") as f:
            pickle.dump(d, f)
        names.append(fn)
    return names


def load_result(fit_path: str):
    """Inverse of save_result: ([T,75] rows, scale, camera_ext [T,4,4])."""
    files = sorted(glob.glob(os.path.join(fit_path, "body_gen_*.pkl")))
    if not files:
        raise FileNotFoundError(f"no body_gen_*.pkl under {fit_path}")
    rows, exts, scale = [], [], None
    for fn in files:
        with open(fn, "rb") as f:
            d = pickle.load(f)
        rows.append(body_params_parse(d))
        exts.append(np.asarray(d["camera_ext"]))
        scale = d["scale"]
    return torch.tensor(np.vstack(rows), dtype=torch.float32), float(scale), torch.tensor(np.stack(exts), dtype=torch.float32)
