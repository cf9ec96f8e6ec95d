"""Fused loss-algebra operators around the chamfer term (CUDA, autograd-aware).

Each mirrors an expression of /root/reference/global_optimization.py:
  verts_transform(verts, cam_ext)      :119-127
  body2world(cam_transl, scale, ext)   :191-206 (vectorised on the device: no per-frame host loop)
  contact_robust_loss(dist, weight)    :295
  second_diff_l1(x)                    :266-267, :381-382, :404-405
  first_diff_l1(x, frame_weight=None)  :304, :415-429
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib


class _TransformFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, verts, mats):
        _lib.require_cuda(verts, mats)
        v = verts.contiguous()
        m = mats.contiguous()
        T, P, _ = v.shape
        out = torch.empty_like(v)
        L = _lib.lib()
        with torch.cuda.device(v.device):
            _lib.check(L.fpv_transform_fwd(_lib.ptr(v), _lib.ptr(m), T, P, _lib.ptr(out), _lib.stream_ptr()),
                       "fpv_transform_fwd")
        ctx.save_for_backward(v, m)
        return out

    @staticmethod
    def backward(ctx, g):
        v, m = ctx.saved_tensors
        T, P, _ = v.shape
        g = g.contiguous()
        need_v, need_m = ctx.needs_input_grad
        L = _lib.lib()
        g_v = torch.empty_like(v) if need_v else None
        g_m = torch.empty_like(m) if need_m else None
        if not (need_v or need_m):
            return None, None
        with torch.cuda.device(v.device):
            ws = _lib.workspace(L.fpv_transform_bwd_workspace_bytes(T, P), v.device)
            _lib.check(L.fpv_transform_bwd(_lib.ptr(v), _lib.ptr(m), _lib.ptr(g), T, P, _lib.ptr(g_v),
                                           _lib.ptr(g_m), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()),
                       "fpv_transform_bwd")
        return g_v, g_m


def verts_transform(verts_batch: torch.Tensor, cam_ext_batch: torch.Tensor) -> torch.Tensor:
    """global_optimization.py:119-127 -- [T,P,3] x [T,4,4] -> [T,P,3] (homogeneous pad, M^T, drop w)."""
    if verts_batch.dim() != 3 or verts_batch.shape[-1] != 3:
        raise RuntimeError(f"verts_transform: expected [T,P,3], got {tuple(verts_batch.shape)}")
    if cam_ext_batch.shape != (verts_batch.shape[0], 4, 4):
        raise RuntimeError(f"verts_transform: expected cam_ext [T,4,4], got {tuple(cam_ext_batch.shape)}")
    return _TransformFn.apply(verts_batch, cam_ext_batch)


def body2world(cam_transl_batch: torch.Tensor, scale: torch.Tensor, camera_ext: torch.Tensor) -> torch.Tensor:
    """global_optimization.py:191-206: camera_ext_t @ [I | cam_transl_t*scale] for all frames at once
    (the reference builds T 4x4 matrices on the host and copies each to the GPU every step)."""
    T = cam_transl_batch.shape[0]
    pose = torch.eye(4, dtype=cam_transl_batch.dtype, device=cam_transl_batch.device).repeat(T, 1, 1)
    pose[:, :3, 3] = cam_transl_batch * scale
    return torch.matmul(camera_ext, pose)


class _RobustMeanFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, d, eps: float):
        _lib.require_cuda(d)
        dc = d.contiguous()
        out = torch.empty(1, dtype=torch.float32, device=d.device)
        L = _lib.lib()
        with torch.cuda.device(d.device):
            ws = _lib.workspace(L.fpv_reduce_workspace_bytes(dc.numel()), d.device)
            _lib.check(L.fpv_robust_mean_fwd(_lib.ptr(dc), dc.numel(), eps, _lib.ptr(out), _lib.ptr(ws),
                                             ws.numel(), _lib.stream_ptr()), "fpv_robust_mean_fwd")
        ctx.save_for_backward(dc)
        ctx.eps = eps
        return out.reshape(())

    @staticmethod
    def backward(ctx, g):
        (dc,) = ctx.saved_tensors
        grad = torch.empty_like(dc)
        gc = g.reshape(1).contiguous().float()
        L = _lib.lib()
        with torch.cuda.device(dc.device):
            _lib.check(L.fpv_robust_mean_bwd(_lib.ptr(dc), dc.numel(), ctx.eps, _lib.ptr(gc), _lib.ptr(grad),
                                             _lib.stream_ptr()), "fpv_robust_mean_bwd")
        return grad, None


def contact_robust_loss(contact_dist: torch.Tensor, weight: float = 1.0, eps: float = 1e-4) -> torch.Tensor:
    """global_optimization.py:295 -- weight * mean( sqrt(d+1e-4) / (sqrt(d+1e-4) + 1) ), one fused pass."""
    if contact_dist.numel() == 0:
        raise RuntimeError("contact_robust_loss: empty input")
    return weight * _RobustMeanFn.apply(contact_dist, float(eps))


class _TDiffFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, order: int, frame_w):
        _lib.require_cuda(x, frame_w)
        xc = x.contiguous()
        T = xc.shape[0]
        F = xc.numel() // T
        wc = frame_w.contiguous().float() if frame_w is not None else None
        out = torch.empty(1, dtype=torch.float32, device=x.device)
        L = _lib.lib()
        with torch.cuda.device(x.device):
            ws = _lib.workspace(L.fpv_reduce_workspace_bytes(xc.numel()), x.device)
            _lib.check(L.fpv_tdiff_l1_fwd(_lib.ptr(xc), T, F, order, _lib.ptr(wc), _lib.ptr(out), _lib.ptr(ws),
                                          ws.numel(), _lib.stream_ptr()), "fpv_tdiff_l1_fwd")
        ctx.save_for_backward(xc, wc)
        ctx.order = order
        return out.reshape(())

    @staticmethod
    def backward(ctx, g):
        xc, wc = ctx.saved_tensors
        T = xc.shape[0]
        F = xc.numel() // T
        grad = torch.empty_like(xc)
        gc = g.reshape(1).contiguous().float()
        L = _lib.lib()
        with torch.cuda.device(xc.device):
            _lib.check(L.fpv_tdiff_l1_bwd(_lib.ptr(xc), T, F, ctx.order, _lib.ptr(wc), _lib.ptr(gc),
                                          _lib.ptr(grad), _lib.stream_ptr()), "fpv_tdiff_l1_bwd")
        return grad, None, None


def _check_seq(x, order):
    if x.dim() < 2:
        raise RuntimeError("temporal residual: expected [T, ...]")
    if x.shape[0] <= order:
        raise RuntimeError(f"temporal residual of order {order} needs more than {order} frames (T={x.shape[0]})")
    if x.dtype != torch.float32:
        raise RuntimeError("temporal residual: float32 required")


def second_diff_l1(x: torch.Tensor) -> torch.Tensor:
    """mean |(x_t - x_{t+1}) - (x_{t+1} - x_{t+2})| over frames  (:266-267, :381-382, :404-405)."""
    _check_seq(x, 2)
    return _TDiffFn.apply(x, 2, None)


def first_diff_l1(x: torch.Tensor, frame_weight: Optional[torch.Tensor] = None) -> torch.Tensor:
    """mean |x_t - x_{t+1}| (:304); with frame_weight [T]: mean |(x_t - x_{t+1}) * w_{t+1}| (:415-429)."""
    _check_seq(x, 1)
    if frame_weight is not None and frame_weight.shape != (x.shape[0],):
        raise RuntimeError("first_diff_l1: frame_weight must have shape [T]")
    return _TDiffFn.apply(x, 1, frame_weight)
