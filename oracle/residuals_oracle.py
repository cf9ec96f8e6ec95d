"""Torch restatement of the loss algebra around the hot path (TEST / BASELINE INFRASTRUCTURE ONLY).

Each function follows the cited lines of /root/reference/global_optimization.py verbatim (same
operand order), generalised only where the reference hard-codes a device.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def verts_transform(verts_batch: torch.Tensor, cam_ext_batch: torch.Tensor) -> torch.Tensor:
    """global_optimization.py:119-127 -- pad to homogeneous, right-multiply by M^T, drop w."""
    verts_batch_homo = F.pad(verts_batch, (0, 1), mode="constant", value=1)
    verts_batch_homo_transformed = torch.matmul(verts_batch_homo, cam_ext_batch.permute(0, 2, 1))
    return verts_batch_homo_transformed[:, :, :-1]


def body2world(cam_transl_batch: torch.Tensor, scale: torch.Tensor, camera_ext: torch.Tensor) -> torch.Tensor:
    """global_optimization.py:191-206 without the per-frame host loop:
    body2world_t = camera_ext_t @ [I | cam_transl_t * scale ; 0 0 0 1]."""
    T = cam_transl_batch.shape[0]
    pose = torch.eye(4, dtype=cam_transl_batch.dtype, device=cam_transl_batch.device).repeat(T, 1, 1)
    pose = pose.clone()
    pose[:, :3, 3] = cam_transl_batch * scale
    return torch.matmul(camera_ext, pose)


def contact_robust_loss(contact_dist: torch.Tensor, weight: float = 1.0) -> torch.Tensor:
    """global_optimization.py:295 -- w * mean( sqrt(d+1e-4) / (sqrt(d+1e-4) + 1) )."""
    return weight * torch.mean(torch.sqrt(contact_dist + 1e-4) / (torch.sqrt(contact_dist + 1e-4) + 1.0))


def second_diff_l1(x: torch.Tensor) -> torch.Tensor:
    """global_optimization.py:266-267 (parameters), :381-382, :404-405 (vertices):
    diff = x[0:-1]-x[1:];  mean |diff[0:-1]-diff[1:]|."""
    diff = x[0:-1] - x[1:]
    return torch.mean(torch.abs(diff[0:-1] - diff[1:]))


def first_diff_l1(x: torch.Tensor) -> torch.Tensor:
    """global_optimization.py:304 -- mean |x[0:-1]-x[1:]| (world joint smoothness)."""
    return torch.mean(torch.abs(x[0:-1] - x[1:]))


def weighted_first_diff_l1(x_sel: torch.Tensor, frame_weight: torch.Tensor) -> torch.Tensor:
    """global_optimization.py:415-429 for one leg: x_sel [T,K,3] gathered leg vertices,
    frame_weight [T] already thresholded (:421-422); uses weight[1:] (:424-425)."""
    diff = x_sel[0:-1] - x_sel[1:]
    w = frame_weight[1:].unsqueeze(1).unsqueeze(1)
    return torch.mean(torch.abs(diff * w))
