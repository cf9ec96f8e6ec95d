"""Torch-CPU restatement of the reference's pure-torch chamfer (TEST / BASELINE INFRASTRUCTURE).

Follows /root/reference/chamfer_python.py line by line, with the three changes SURVEY.md
section 8c lists as necessary to run it at all outside a CUDA box with N == M:

  chamfer_python.py:21-23  xx/yy/zz = bmm(...)      -> zz is tiled over M; |x|^2 and |y|^2 are
                                                       computed as the row-wise dot product the
                                                       diagonal of the bmm holds (no N x N matrix)
  chamfer_python.py:24     torch.cuda.LongTensor    -> not needed (no diagonal gather)
  chamfer_python.py:27     P = rx^T + ry - 2 zz     -> same expression, same operand order
  chamfer_python.py:28     min over dim 1 / dim 2   -> running (value, index) minimum per tile with
                                                       strict '<' so the FIRST occurrence wins, which
                                                       is what torch.min returns on CPU

It is the arithmetic ("port") the CPU baseline of bench.py times, and the literal-reference
stand-in for the lattice known-answer tests; it is NOT the bit-exact index oracle (that is
oracle/chamfer_oracle.c with the canonical direct-difference arithmetic).
"""
from __future__ import annotations

import torch


def pairwise_dist(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """chamfer_python.py:4-9 for N != M: P[i,j] = |x_i|^2 + |y_j|^2 - 2 x_i.y_j."""
    rx = (x * x).sum(-1)
    ry = (y * y).sum(-1)
    zz = torch.mm(x, y.t())
    return rx.unsqueeze(1) + ry.unsqueeze(0) - 2 * zz


def NN_loss(x: torch.Tensor, y: torch.Tensor, dim: int = 0) -> torch.Tensor:
    """chamfer_python.py:12-15."""
    values, _ = pairwise_dist(x, y).min(dim=dim)
    return values.mean()


def distChamfer(a: torch.Tensor, b: torch.Tensor, tile: int = 8192):
    """chamfer_python.py:18-28 semantics for any N, M; b may be [bs,M,3] or a shared [1,M,3]/[M,3].

    Returns (d_b2a [bs,M], d_a2b [bs,N], i_b2a [bs,M] int64, i_a2b [bs,N] int64).
    Not differentiable (use distChamfer_autograd for gradients on small inputs).
    """
    x = a
    bs, N, _ = x.shape
    if b.dim() == 2:
        b = b.unsqueeze(0)
    M = b.shape[1]
    d_a2b = torch.empty(bs, N, dtype=x.dtype)
    i_a2b = torch.empty(bs, N, dtype=torch.int64)
    d_b2a = torch.empty(bs, M, dtype=x.dtype)
    i_b2a = torch.empty(bs, M, dtype=torch.int64)
    for s in range(bs):
        xs = x[s]
        ys = b[s if b.shape[0] > 1 else 0]
        rx = (xs * xs).sum(-1)                       # diag(xx), :21,:25
        best_d = torch.full((N,), float("inf"), dtype=x.dtype)
        best_i = torch.zeros(N, dtype=torch.int64)
        for j0 in range(0, M, tile):
            yt = ys[j0:j0 + tile]
            ry = (yt * yt).sum(-1)                   # diag(yy), :22,:26
            zz = torch.mm(xs, yt.t())                # :23
            P = rx.unsqueeze(1) + ry.unsqueeze(0) - 2 * zz   # :27
            v, i = P.min(dim=1)                      # torch.min(P, 2) of the batched form
            upd = v < best_d
            best_d = torch.where(upd, v, best_d)
            best_i = torch.where(upd, i + j0, best_i)
            v0, i0 = P.min(dim=0)                    # torch.min(P, 1) of the batched form
            d_b2a[s, j0:j0 + tile] = v0
            i_b2a[s, j0:j0 + tile] = i0
        d_a2b[s] = best_d
        i_a2b[s] = best_i
    return d_b2a, d_a2b, i_b2a, i_a2b


def distChamfer_autograd(a: torch.Tensor, b: torch.Tensor):
    """Untiled, differentiable form (small inputs only) -- the literal expression of :21-28."""
    rx = (a * a).sum(-1)
    ry = (b * b).sum(-1)
    zz = torch.bmm(a, b.transpose(2, 1))
    P = rx.unsqueeze(2) + ry.unsqueeze(1) - 2 * zz
    m1 = torch.min(P, 1)
    m2 = torch.min(P, 2)
    return m1[0], m2[0], m1[1], m2[1]
