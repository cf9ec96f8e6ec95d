"""Parameter front-end and trajectory prior: the steps immediately before the
SMPL-X forward and the temporal residual of the `dct` mode, with the reference's own names and argument meaning.

    convert_to_3D_rot(x)      global_optimization.py:107-115   [T,78] 6D row -> [T,75] axis-angle row (differentiable)
    convert_to_6D_rot(x)      global_optimization.py:96-104    [T,75] -> [T,78]
    VPoserDecoderB200.decode  [3P] human_body_prior v1 VPoser.decode(z, output_type='aa'), call site :270-271
    body_params_encapsulate_batch   column split of the 75-D row (cvae.py:189-208, call site :268)
    front_end_split(x)        the two above composed as one autograd node (what FitProblem runs)
    cal_dctloss(joints, dct_mtx, c_dct)   FittingOP.cal_dctloss, :232-246
    dct_basis(F, K)           orthonormal DCT-II basis (the reference loads ../Data/DCT_Basis/60.mat, absent)

Everything runs through libfpv_b200.so (csrc/prior.cu); CPU tensors raise.
"""
from __future__ import annotations

import ctypes
import math
from typing import Dict, Optional

import torch

from . import _lib


def _f32c(x: torch.Tensor, what: str) -> torch.Tensor:
    if x.dtype != torch.float32:
        raise RuntimeError(f"{what}: float32 required, got {x.dtype}")
    _lib.require_cuda(x)
    return x.contiguous()


class _Rot6dToAAFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x6):
        x = _f32c(x6, "rot6d_to_aa")
        n = x.numel() // 6
        aa = torch.empty(n, 3, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().fpv_rot6d_to_aa_fwd(_lib.ptr(x), n, _lib.ptr(aa), _lib.stream_ptr()), "fpv_rot6d_to_aa_fwd")
        ctx.save_for_backward(x)
        return aa

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        n = x.numel() // 6
        g = g.contiguous().float()
        gx = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().fpv_rot6d_to_aa_bwd(_lib.ptr(x), n, _lib.ptr(g), _lib.ptr(gx), _lib.stream_ptr()),
                       "fpv_rot6d_to_aa_bwd")
        return gx


def rot6d_to_aa(x6: torch.Tensor) -> torch.Tensor:
    """[n,6] (the (3,2) matrix of cvae.py:72, row-major) -> [n,3] axis-angle; ContinousRotReprDecoder.decode + matrot2aa."""
    if x6.shape[-1] != 6:
        raise RuntimeError(f"rot6d_to_aa: last dimension must be 6, got {tuple(x6.shape)}")
    return _Rot6dToAAFn.apply(x6.reshape(-1, 6))


def aa_to_rot6d(aa: torch.Tensor) -> torch.Tensor:
    """[n,3] -> [n,6]: aa2matrot(aa)[:, :, :-1].reshape(-1, 6) (global_optimization.py:101-102).  Not differentiated
    (the reference applies it to the observed data only)."""
    if aa.shape[-1] != 3:
        raise RuntimeError(f"aa_to_rot6d: last dimension must be 3, got {tuple(aa.shape)}")
    a = _f32c(aa.detach().reshape(-1, 3), "aa_to_rot6d")
    out = torch.empty(a.shape[0], 6, dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(_lib.lib().fpv_aa_to_rot6d(_lib.ptr(a), a.shape[0], _lib.ptr(out), _lib.stream_ptr()), "fpv_aa_to_rot6d")
    return out


def convert_to_6D_rot(x_batch: torch.Tensor) -> torch.Tensor:
    return torch.cat([x_batch[:, :3], aa_to_rot6d(x_batch[:, 3:6]), x_batch[:, 6:]], dim=-1)


def convert_to_3D_rot(x_batch: torch.Tensor) -> torch.Tensor:
    return torch.cat([x_batch[:, :3], rot6d_to_aa(x_batch[:, 3:9]), x_batch[:, 9:]], dim=-1)


def body_params_encapsulate_batch(x_body_rec: torch.Tensor) -> Dict[str, torch.Tensor]:
    """The 75-D row split into the keyword arguments of the body model (+ the VPoser latent).  The batched static method
    the loop calls (global_optimization.py:268) is missing from the reference's cvae.py; this follows its per-frame
    sibling body_params_encapsulate (cvae.py:189-208) and the key the caller reads (`body_pose_vp`, :270)."""
    return {"transl": x_body_rec[:, :3], "global_orient": x_body_rec[:, 3:6], "betas": x_body_rec[:, 6:16],
            "body_pose_vp": x_body_rec[:, 16:48], "left_hand_pose": x_body_rec[:, 48:60],
            "right_hand_pose": x_body_rec[:, 60:72], "camera_translation": x_body_rec[:, 72:75]}


_ROW78 = (3, 6, 10, 32, 12, 12, 3)   # transl | 6D global orientation | betas | VPoser latent | hands | camera translation
_ROW78_KEYS = ("transl", "global_orient", "betas", "body_pose_vp", "left_hand_pose", "right_hand_pose", "camera_translation")


class _FrontEndSplitFn(torch.autograd.Function):
    """body_params_encapsulate_batch(convert_to_3D_rot(x)) as ONE autograd node: the column blocks of the 78-D row as
    contiguous tensors, the orientation block decoded to axis-angle.  The backward is the 6D codec's VJP and one
    concatenation -- the chain of slice / cat nodes it replaces costs a zero-fill, a copy and an add per block."""

    @staticmethod
    def forward(ctx, x):
        xc = _f32c(x, "front_end_split")
        blocks = [b.contiguous() for b in xc.split(_ROW78, dim=1)]
        n = xc.shape[0]
        aa = torch.empty(n, 3, dtype=torch.float32, device=xc.device)
        with torch.cuda.device(xc.device):
            _lib.check(_lib.lib().fpv_rot6d_to_aa_fwd(_lib.ptr(blocks[1]), n, _lib.ptr(aa), _lib.stream_ptr()),
                       "fpv_rot6d_to_aa_fwd")
        ctx.save_for_backward(blocks[1])
        blocks[1] = aa
        return tuple(blocks)

    @staticmethod
    def backward(ctx, *grads):
        (x6,) = ctx.saved_tensors
        n = x6.shape[0]
        cols = []
        for k, (g, w) in enumerate(zip(grads, _ROW78)):
            if k == 1:
                g6 = torch.zeros_like(x6) if g is None else torch.empty_like(x6)
                if g is not None:
                    gc = g.contiguous().float()
                    with torch.cuda.device(x6.device):
                        _lib.check(_lib.lib().fpv_rot6d_to_aa_bwd(_lib.ptr(x6), n, _lib.ptr(gc), _lib.ptr(g6), _lib.stream_ptr()),
                                   "fpv_rot6d_to_aa_bwd")
                cols.append(g6)
            else:
                cols.append(g if g is not None else x6.new_zeros(n, w))
        return torch.cat(cols, dim=1)


def front_end_split(x_batch: torch.Tensor) -> Dict[str, torch.Tensor]:
    """== body_params_encapsulate_batch(convert_to_3D_rot(x_batch)) (global_optimization.py:261-268), fused."""
    if x_batch.dim() != 2 or x_batch.shape[1] != sum(_ROW78):
        raise RuntimeError(f"front_end_split: expected [T,{sum(_ROW78)}], got {tuple(x_batch.shape)}")
    return dict(zip(_ROW78_KEYS, _FrontEndSplitFn.apply(x_batch)))


# ---------------------------------------------------------------------------------------------
# VPoser decoder
# ---------------------------------------------------------------------------------------------
def make_vposer_weights(seed: int = 1234, latent: int = 32, hidden: int = 512, joints: int = 21) -> Dict[str, torch.Tensor]:
    """Random-init decoder weights of the VPoser v1 shapes (the snapshot is licence-gated and absent): nn.Linear init,
    with the output layer damped and biased to the identity rotation so that decoded poses stay in the range of a
    plausible body (|angle| of a few tenths of a radian), like the axis-angle clips of synthetic.make_clip_params."""
    g = torch.Generator().manual_seed(seed)

    def linear(o, i):
        b = 1.0 / math.sqrt(i)
        return (torch.rand(o, i, generator=g) * 2 - 1) * b, (torch.rand(o, generator=g) * 2 - 1) * b

    w1, b1 = linear(hidden, latent)
    w2, b2 = linear(hidden, hidden)
    w3, b3 = linear(6 * joints, hidden)
    w3 = w3 * 0.5
    b3 = b3 * 0.5 + torch.tensor([1.0, 0.0, 0.0, 1.0, 0.0, 0.0]).repeat(joints)
    return dict(w1=w1, b1=b1, w2=w2, b2=b2, w3=w3, b3=b3)


class _VposerDecodeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, dec):
        zc = _f32c(z, "VPoser.decode")
        T = zc.shape[0]
        L = _lib.lib()
        aa = torch.empty(T, dec.joints, 3, dtype=torch.float32, device=zc.device)
        saved = torch.empty(L.fpv_vposer_saved_floats(ctypes.byref(dec.struct), T), dtype=torch.float32, device=zc.device)
        with torch.cuda.device(zc.device):
            _lib.check(L.fpv_vposer_decode_fwd(ctypes.byref(dec.struct), _lib.ptr(zc), T, _lib.ptr(aa), _lib.ptr(saved),
                                               _lib.stream_ptr()), "fpv_vposer_decode_fwd")
        ctx.dec = dec
        ctx.save_for_backward(saved)
        return aa

    @staticmethod
    def backward(ctx, g):
        (saved,) = ctx.saved_tensors
        dec = ctx.dec
        g = g.contiguous().float()
        T = g.shape[0]
        gz = torch.empty(T, dec.latent, dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            _lib.check(_lib.lib().fpv_vposer_decode_bwd(ctypes.byref(dec.struct), _lib.ptr(saved), _lib.ptr(g), T,
                                                        _lib.ptr(gz), _lib.stream_ptr()), "fpv_vposer_decode_bwd")
        return gz, None


class VPoserDecoderB200(torch.nn.Module):
    """The decoder half of VPoser v1 with frozen weights (the fit never updates them, global_optimization.py:153)."""

    def __init__(self, weights: Optional[Dict[str, torch.Tensor]] = None):
        super().__init__()
        w = weights if weights is not None else make_vposer_weights()
        for k in ("w1", "b1", "w2", "b2", "w3", "b3"):
            self.register_buffer(k, w[k].detach().to(torch.float32).contiguous())
        for k in ("w1", "w2", "w3"):
            self.register_buffer(k + "t", w[k].detach().to(torch.float32).t().contiguous())
        self.hidden, self.latent = self.w1.shape
        self.joints = self.w3.shape[0] // 6
        if self.w2.shape != (self.hidden, self.hidden) or self.w3.shape != (6 * self.joints, self.hidden):
            raise RuntimeError("VPoserDecoderB200: inconsistent layer shapes")
        self._struct_key = None
        self._struct = None

    @property
    def struct(self) -> _lib.VposerModelStruct:
        key = tuple(getattr(self, k).data_ptr() for k in ("w1", "w2", "w3", "w1t", "w2t", "w3t", "b1", "b2", "b3"))
        if key != self._struct_key:
            s = _lib.VposerModelStruct()
            for k in ("w1", "b1", "w2", "b2", "w3", "b3", "w1t", "w2t", "w3t"):
                setattr(s, k, getattr(self, k).data_ptr())
            s.latent, s.hidden, s.joints = self.latent, self.hidden, self.joints
            self._struct, self._struct_key = s, key
        return self._struct

    def decode(self, Zin: torch.Tensor, output_type: str = "aa") -> torch.Tensor:
        """[T,latent] -> [T,1,joints,3] (the reference reshapes it with .view(batch_size, -1), :271)."""
        if output_type != "aa":
            raise RuntimeError("VPoserDecoderB200.decode: only output_type='aa' is on the hot path")
        if Zin.dim() != 2 or Zin.shape[1] != self.latent:
            raise RuntimeError(f"VPoser.decode: expected [T,{self.latent}], got {tuple(Zin.shape)}")
        _lib.require_cuda(Zin, self.w1)
        return _VposerDecodeFn.apply(Zin, self).view(Zin.shape[0], 1, self.joints, 3)


# ---------------------------------------------------------------------------------------------
# DCT prior
# ---------------------------------------------------------------------------------------------
def dct_basis(frames: int, num: int, device=None) -> torch.Tensor:
    """[frames,num] orthonormal DCT-II basis: what load_dct_base() returns (rows of D, transposed, :131-137)."""
    n = torch.arange(frames, dtype=torch.float64)
    cols = [torch.cos(math.pi * (n + 0.5) * k / frames) * (math.sqrt(1.0 / frames) if k == 0 else math.sqrt(2.0 / frames))
            for k in range(num)]
    return torch.stack(cols, dim=1).to(torch.float32).to(device)


class _DctPriorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, basis, coef):
        NB, C, K = coef.shape
        F = basis.shape[0]
        L = _lib.lib()
        out = torch.empty(1, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            ws = _lib.workspace(L.fpv_dct_prior_workspace_bytes(NB, F, C), x.device)
            _lib.check(L.fpv_dct_prior_fwd(_lib.ptr(x), _lib.ptr(basis), _lib.ptr(coef), NB, F, C, K, _lib.ptr(out),
                                           _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "fpv_dct_prior_fwd")
        ctx.save_for_backward(x, basis, coef)
        return out.reshape(())

    @staticmethod
    def backward(ctx, g):
        x, basis, coef = ctx.saved_tensors
        NB, C, K = coef.shape
        F = basis.shape[0]
        g = g.reshape(1).contiguous().float()
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gc = torch.empty_like(coef) if ctx.needs_input_grad[2] else None
        if gx is not None or gc is not None:
            with torch.cuda.device(x.device):
                _lib.check(_lib.lib().fpv_dct_prior_bwd(_lib.ptr(x), _lib.ptr(basis), _lib.ptr(coef), NB, F, C, K, _lib.ptr(g),
                                                        _lib.ptr(gx), _lib.ptr(gc), _lib.stream_ptr()), "fpv_dct_prior_bwd")
        return gx, None, gc


def cal_dctloss(body_joints_batch: torch.Tensor, dct_mtx: torch.Tensor, c_dct: torch.Tensor) -> torch.Tensor:
    """FittingOP.cal_dctloss: body_joints_batch [>=NB*F, J, 3] (J = 23 joints), dct_mtx [F,K], c_dct [NB,J,3,K]."""
    if c_dct.dim() != 4 or dct_mtx.dim() != 2 or body_joints_batch.dim() != 3:
        raise RuntimeError("cal_dctloss: expected joints [T,J,3], dct_mtx [F,K], c_dct [NB,J,3,K]")
    NB, J, A, K = c_dct.shape
    F = dct_mtx.shape[0]
    if dct_mtx.shape[1] != K or body_joints_batch.shape[1] != J or body_joints_batch.shape[2] != A:
        raise RuntimeError("cal_dctloss: shape mismatch between joints, dct_mtx and c_dct")
    if body_joints_batch.shape[0] < NB * F:
        raise RuntimeError(f"cal_dctloss: {body_joints_batch.shape[0]} frames < NUM_BATCHES*BATCH_FRAME_NUM = {NB * F}")
    x = _f32c(body_joints_batch[:NB * F].reshape(NB * F, J * A), "cal_dctloss")
    return _DctPriorFn.apply(x, _f32c(dct_mtx, "cal_dctloss"), _f32c(c_dct.reshape(NB, J * A, K), "cal_dctloss"))
