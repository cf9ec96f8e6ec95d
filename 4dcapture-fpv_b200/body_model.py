"""SMPL-X body model with the [3P] `smplx` call signature, backed by the sm_100a kernels.

Mirrors what the reference constructs and calls:
  smplx.create(model_path, model_type='smplx', gender='neutral', ext='npz', num_pca_comps=12,
               create_*=True, batch_size=T)                       global_optimization.py:154-168
  body_mesh_model(return_verts=True, body_pose=..., transl=..., global_orient=..., betas=...,
                  left_hand_pose=..., right_hand_pose=...)        global_optimization.py:280-283
  -> output.vertices [T,10475,3], output.joints [T,>=23,3]        :283, :298

Omitted arguments fall back to the module's own zero-initialised parameters, as in smplx.
The constants come either from a real SMPL-X .npz (same keys the smplx package reads) or from
synthetic.make_body_constants() when the licence-gated file is absent.
"""
from __future__ import annotations

import os
from collections import namedtuple
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .synthetic import NUM_PCA, make_body_constants

SMPLXOutput = namedtuple("SMPLXOutput", ["vertices", "joints", "full_pose", "betas", "global_orient", "body_pose",
                                         "expression", "left_hand_pose", "right_hand_pose", "jaw_pose", "transl"])

NJ, NPF, NSH, KP, NTH = 55, 486, 20, 512, 122


def load_smplx_npz(path: str, num_betas: int = 10, num_expression_coeffs: int = 10,
                   num_pca_comps: int = NUM_PCA, flat_hand_mean: bool = False) -> Dict[str, torch.Tensor]:
    """Read a real SMPL-X model file into the canonical constant layout ([3P] smplx.body_models.SMPLX.__init__)."""
    d = np.load(path, allow_pickle=True)
    shapedirs = np.asarray(d["shapedirs"], np.float64)
    # expression directions: columns 300.. of the 400-column SMPL-X v1.1 files; SMPL-X v1.0 files carry only 20 columns
    # (10 shape + 10 expression) and the smplx package then reads the expression part from column 10 on
    expr_start = 300 if shapedirs.shape[-1] >= 300 + num_expression_coeffs else 10
    n_expr = min(num_expression_coeffs, shapedirs.shape[-1] - expr_start)
    if n_expr <= 0:
        raise RuntimeError(f"load_smplx_npz: shapedirs has {shapedirs.shape[-1]} columns, no expression directions found")
    sd = np.concatenate([shapedirs[:, :, :num_betas], shapedirs[:, :, expr_start:expr_start + n_expr]], -1)
    if sd.shape[-1] < num_betas + num_expression_coeffs:     # pad missing expression directions with zeros
        sd = np.concatenate([sd, np.zeros(sd.shape[:2] + (num_betas + num_expression_coeffs - sd.shape[-1],))], -1)
    posedirs = np.asarray(d["posedirs"], np.float64)
    posedirs = posedirs.reshape(-1, posedirs.shape[-1]).T                      # [486, 3V]
    parents = np.asarray(d["kintree_table"])[0].astype(np.int64)
    parents[0] = -1
    lh = np.asarray(d["hands_componentsl"], np.float64)[:num_pca_comps]
    rh = np.asarray(d["hands_componentsr"], np.float64)[:num_pca_comps]
    pose_mean = np.zeros(165, np.float64)
    if not flat_hand_mean:
        pose_mean[75:120] = np.asarray(d["hands_meanl"], np.float64)
        pose_mean[120:165] = np.asarray(d["hands_meanr"], np.float64)
    f32 = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float32)
    return dict(v_template=f32(d["v_template"]), shapedirs=f32(sd), posedirs=f32(posedirs),
                J_regressor=f32(np.asarray(d["J_regressor"], np.float64)), parents=torch.tensor(parents),
                lbs_weights=f32(d["weights"]), lh_components=f32(lh), rh_components=f32(rh),
                pose_mean=f32(pose_mean), extra_vertex_ids=torch.zeros(0, dtype=torch.int64))


def split_tf32(x: torch.Tensor):
    """x -> (hi, lo), each exactly representable in TF32 (10 explicit mantissa bits), hi + lo = x to 2^-22.
    Same bit recipe as the device code (round-half-up on the magnitude, csrc/tc_gemm.cu::tf32_round)."""
    def rn(t):
        return ((t.contiguous().view(torch.int32) + 0x1000) & -0x2000).view(torch.float32)
    hi = rn(x)
    return hi, rn(x - hi)


class _SMPLXFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, theta, module):
        _lib.require_cuda(theta)
        th = theta.contiguous()
        T = th.shape[0]
        dev = th.device
        L = _lib.lib()
        ms = module._struct(dev)
        V, E = module.num_verts, module.num_extra
        verts = torch.empty(T, V, 3, dtype=torch.float32, device=dev)
        joints = torch.empty(T, NJ + E, 3, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            saved = _lib.workspace(L.fpv_smplx_saved_bytes(ms, T), dev)
            _lib.check(L.fpv_smplx_fwd(ms, T, _lib.ptr(th), _lib.ptr(verts), _lib.ptr(joints), _lib.ptr(saved),
                                       None, 0, _lib.stream_ptr()), "fpv_smplx_fwd")
        ctx.save_for_backward(th, saved)
        ctx.module = module
        ctx.set_materialize_grads(False)
        return verts, joints

    @staticmethod
    def backward(ctx, g_verts, g_joints):
        th, saved = ctx.saved_tensors
        if g_verts is None and g_joints is None:
            return None, None
        module = ctx.module
        T = th.shape[0]
        dev = th.device
        L = _lib.lib()
        ms = module._struct(dev)
        gv = g_verts.contiguous().float() if g_verts is not None else None
        gj = g_joints.contiguous().float() if g_joints is not None else None
        g_theta = torch.empty_like(th)
        with torch.cuda.device(dev):
            ws = _lib.workspace(L.fpv_smplx_workspace_bytes(ms, T), dev)
            _lib.check(L.fpv_smplx_bwd(ms, T, _lib.ptr(th), _lib.ptr(saved), _lib.ptr(gv), _lib.ptr(gj),
                                       _lib.ptr(g_theta), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()),
                       "fpv_smplx_bwd")
        return g_theta, None


class SMPLXB200(nn.Module):
    """nn.Module with SMPLX.forward's keyword signature; all T frames in four kernel launches."""

    NUM_JOINTS = NJ
    NUM_BODY_JOINTS = 21

    def __init__(self, constants: Dict[str, torch.Tensor], batch_size: int = 1, num_pca_comps: int = NUM_PCA,
                 create_global_orient=True, create_body_pose=True, create_betas=True,
                 create_left_hand_pose=True, create_right_hand_pose=True, create_expression=True,
                 create_jaw_pose=True, create_leye_pose=True, create_reye_pose=True, create_transl=True,
                 dtype=torch.float32, **_ignored):
        super().__init__()
        if dtype != torch.float32:
            raise RuntimeError("SMPLXB200 computes in float32 only")
        c = constants
        V = c["v_template"].shape[0]
        if c["posedirs"].shape != (NPF, 3 * V) or c["shapedirs"].shape != (V, 3, NSH):
            raise RuntimeError("SMPL-X constants have unexpected shapes")
        if c["lh_components"].shape[0] != num_pca_comps:
            raise RuntimeError("num_pca_comps does not match the hand components")
        if num_pca_comps != 12:
            raise RuntimeError("this build fixes num_pca_comps=12 (global_optimization.py:156)")
        self.batch_size = batch_size
        self.num_verts = V
        self.num_pca_comps = num_pca_comps
        f64 = lambda t: t.to(torch.float64)
        # --- one-time constant preparation (host, float64 -> float32) ---
        basis = torch.zeros(KP, 3 * V, dtype=torch.float64)
        basis[:NPF] = f64(c["posedirs"])
        basis[NPF:NPF + NSH] = f64(c["shapedirs"]).permute(2, 0, 1).reshape(NSH, 3 * V)
        basis[NPF + NSH] = f64(c["v_template"]).reshape(3 * V)
        Jr = f64(c["J_regressor"])
        j_template = Jr @ f64(c["v_template"])                                   # [55,3]
        j_shapedirs = torch.einsum("jv,vcl->jcl", Jr, f64(c["shapedirs"]))       # [55,3,20]
        W = c["lbs_weights"]
        nnz = (W != 0).sum(1)
        width = int(max(1, nnz.max().item()))
        order = torch.argsort((W != 0).to(torch.int8), dim=1, descending=True, stable=True)[:, :width]
        ell_w = torch.gather(W, 1, order)
        ell_j = torch.where(ell_w != 0, order, torch.full_like(order, -1)).to(torch.int32)
        # per-joint influence lists (transpose of the ELL matrix), vertices ascending
        vv, jj = torch.nonzero(W.t().contiguous() != 0, as_tuple=True)[::-1]
        jj_sorted, perm = torch.sort(jj, stable=True)
        csr_vert = vv[perm].to(torch.int32)
        csr_w = W[vv[perm], jj_sorted].to(torch.float32)
        csr_ptr = torch.zeros(NJ + 1, dtype=torch.int32)
        csr_ptr[1:] = torch.cumsum(torch.bincount(jj_sorted, minlength=NJ), 0).to(torch.int32)
        extra = c.get("extra_vertex_ids", torch.zeros(0, dtype=torch.int64))
        self.num_extra = int(extra.numel())
        f32 = lambda t: t.to(torch.float32).contiguous()
        # TF32 hi/lo splits of the basis in both tensor-core operand orientations (K contiguous):
        #   forward  v_posed = coef x basis_nk^T : basis_nk [3V, 512]
        #   backward gC = g_vposed x basis_kn^T  : basis_kn [512, P], P = 3V rounded up to 4 (TMA row pitch)
        b32 = f32(basis)
        P = (3 * V + 3) // 4 * 4
        kn = torch.zeros(KP, P, dtype=torch.float32)
        kn[:, :3 * V] = b32
        kn_hi, kn_lo = split_tf32(kn)
        nk_hi, nk_lo = split_tf32(b32.t().contiguous())
        self.register_buffer("basis_kn", b32, persistent=False)
        self.register_buffer("basis_kn_hi", kn_hi, persistent=False)
        self.register_buffer("basis_kn_lo", kn_lo, persistent=False)
        self.register_buffer("basis_nk_hi", nk_hi, persistent=False)
        self.register_buffer("basis_nk_lo", nk_lo, persistent=False)
        self.register_buffer("j_template", f32(j_template))
        self.register_buffer("j_shapedirs", f32(j_shapedirs))
        self.register_buffer("parents", c["parents"].to(torch.int32).contiguous())
        self.register_buffer("hand_comps", f32(torch.stack([c["lh_components"], c["rh_components"]])))
        self.register_buffer("pose_mean", f32(c["pose_mean"]))
        self.register_buffer("ell_joint", ell_j.t().contiguous())
        self.register_buffer("ell_weight", f32(ell_w.t()))
        self.register_buffer("csr_ptr", csr_ptr)
        self.register_buffer("csr_vert", csr_vert.contiguous())
        self.register_buffer("csr_weight", csr_w.contiguous())
        self.register_buffer("extra_vertex_ids", extra.to(torch.int32).contiguous())
        self.ell_width = width
        # --- default (zero) parameters, as smplx creates them ---
        def param(flag, name, dim):
            if flag:
                self.register_parameter(name, nn.Parameter(torch.zeros(batch_size, dim, dtype=torch.float32)))
        param(create_global_orient, "global_orient", 3)
        param(create_body_pose, "body_pose", 63)
        param(create_betas, "betas", 10)
        param(create_left_hand_pose, "left_hand_pose", num_pca_comps)
        param(create_right_hand_pose, "right_hand_pose", num_pca_comps)
        param(create_expression, "expression", 10)
        param(create_jaw_pose, "jaw_pose", 3)
        param(create_leye_pose, "leye_pose", 3)
        param(create_reye_pose, "reye_pose", 3)
        param(create_transl, "transl", 3)
        self._struct_cache = {}

    def _apply(self, fn, *a, **k):
        self._struct_cache = {}
        return super()._apply(fn, *a, **k)

    def _struct(self, device):
        key = (device.type, device.index)
        s = self._struct_cache.get(key)
        if s is None:
            if self.basis_kn.device != device:
                raise RuntimeError(f"SMPLXB200 constants live on {self.basis_kn.device}, input on {device}")
            s = _lib.SmplxModelStruct()
            s.num_verts, s.num_extra, s.ell_width, s.reserved = self.num_verts, self.num_extra, self.ell_width, 0
            p = lambda t: t.data_ptr() if t.numel() else None
            s.basis_kn = p(self.basis_kn)
            s.basis_nk_hi, s.basis_nk_lo = p(self.basis_nk_hi), p(self.basis_nk_lo)
            s.basis_kn_hi, s.basis_kn_lo = p(self.basis_kn_hi), p(self.basis_kn_lo)
            s.j_template, s.j_shapedirs, s.parents = p(self.j_template), p(self.j_shapedirs), p(self.parents)
            s.hand_comps, s.pose_mean = p(self.hand_comps), p(self.pose_mean)
            s.ell_joint, s.ell_weight = p(self.ell_joint), p(self.ell_weight)
            s.csr_ptr, s.csr_vert, s.csr_weight = p(self.csr_ptr), p(self.csr_vert), p(self.csr_weight)
            s.extra_vertex_ids = p(self.extra_vertex_ids)
            self._struct_cache[key] = s
        return s

    def _default(self, name, T, dim, device):
        p = getattr(self, name, None)
        if p is None:
            return torch.zeros(T, dim, dtype=torch.float32, device=device)
        if p.shape[0] != T:
            if p.shape[0] == 1:
                return p.expand(T, dim)
            raise RuntimeError(f"SMPLXB200: batch {T} does not match the module's batch_size {p.shape[0]} for `{name}`")
        return p

    def forward(self, betas=None, global_orient=None, body_pose=None, left_hand_pose=None, right_hand_pose=None,
                transl=None, expression=None, jaw_pose=None, leye_pose=None, reye_pose=None,
                return_verts: bool = True, return_full_pose: bool = False, pose2rot: bool = True, **_ignored):
        if not pose2rot:
            raise RuntimeError("SMPLXB200: pose2rot=False (rotation-matrix input) is not supported")
        given = [t for t in (betas, global_orient, body_pose, left_hand_pose, right_hand_pose, transl, expression,
                             jaw_pose, leye_pose, reye_pose) if t is not None]
        T = given[0].shape[0] if given else self.batch_size
        dev = given[0].device if given else self.basis_kn.device

        def arg(t, name, dim):
            if t is None:
                return self._default(name, T, dim, dev)
            t = t.reshape(T, -1)
            if t.shape[1] != dim:
                raise RuntimeError(f"SMPLXB200: `{name}` must have {dim} values per frame, got {t.shape[1]}")
            return t.to(torch.float32)

        parts = [arg(global_orient, "global_orient", 3), arg(body_pose, "body_pose", 63),
                 arg(jaw_pose, "jaw_pose", 3), arg(leye_pose, "leye_pose", 3), arg(reye_pose, "reye_pose", 3),
                 arg(left_hand_pose, "left_hand_pose", self.num_pca_comps),
                 arg(right_hand_pose, "right_hand_pose", self.num_pca_comps),
                 arg(betas, "betas", 10), arg(expression, "expression", 10), arg(transl, "transl", 3)]
        theta = torch.cat(parts, dim=1)                                          # [T,122]
        verts, joints = _SMPLXFn.apply(theta, self)
        return SMPLXOutput(vertices=verts if return_verts else None, joints=joints, full_pose=None,
                           betas=parts[7], global_orient=parts[0], body_pose=parts[1], expression=parts[8],
                           left_hand_pose=parts[5], right_hand_pose=parts[6], jaw_pose=parts[2], transl=parts[9])


def create(model_path: Optional[str] = None, model_type: str = "smplx", gender: str = "neutral", ext: str = "npz",
           num_pca_comps: int = NUM_PCA, batch_size: int = 1, constants: Optional[Dict[str, torch.Tensor]] = None,
           seed: int = 1234, **kwargs) -> SMPLXB200:
    """smplx.create(...) look-alike (global_optimization.py:154-168).

    Looks for <model_path>/smplx/SMPLX_<GENDER>.<ext> (or model_path itself when it is a file); when no
    model file exists -- it is licence-gated -- falls back to the seeded random-init SMPL-X-topology
    constants BASELINE.json prescribes (`constants` overrides both).
    """
    if model_type != "smplx":
        raise RuntimeError("only model_type='smplx' is implemented (the reference uses nothing else)")
    if constants is None:
        cand = None
        if model_path:
            if os.path.isfile(model_path):
                cand = model_path
            else:
                p = os.path.join(model_path, "smplx", f"SMPLX_{gender.upper()}.{ext}")
                cand = p if os.path.isfile(p) else None
        constants = load_smplx_npz(cand, num_pca_comps=num_pca_comps) if cand else make_body_constants(seed)
    return SMPLXB200(constants, batch_size=batch_size, num_pca_comps=num_pca_comps, **kwargs)
