"""Synthetic inputs of the named shapes (SURVEY.md section 8d "Synthetic inputs").

The SMPL-X model files are licence-gated and absent (global_optimization.py:669 './models'), so
BASELINE.json prescribes a "random-init SMPL-X-topology mesh": same sizes (V=10,475 vertices,
55 joints, 486 pose features, 20 shape+expression coefficients, 12 hand PCA components), the real
kinematic tree, random constants.  Everything is generated on the CPU with seeded generators so the
oracle, the CUDA path and the CPU baseline see identical bytes on every machine.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

NUM_VERTS = 10475
NUM_JOINTS = 55
NUM_POSE_FEAT = 9 * (NUM_JOINTS - 1)          # 486
NUM_SHAPE = 20                                # 10 betas + 10 expression
NUM_PCA = 12

# [3P] SMPL-X kinematic tree (public model definition)
SMPLX_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 15, 15, 15,
                 20, 25, 26, 20, 28, 29, 20, 31, 32, 20, 34, 35, 20, 37, 38,
                 21, 40, 41, 21, 43, 44, 21, 46, 47, 21, 49, 50, 21, 52, 53]


def _rest_joints() -> torch.Tensor:
    """A 1.7 m stick person in SMPL's y-up frame, pelvis near the origin."""
    J = torch.zeros(NUM_JOINTS, 3, dtype=torch.float64)
    body = {
        0: (0.0, 0.0, 0.0), 1: (0.09, -0.08, 0.0), 2: (-0.09, -0.08, 0.0), 3: (0.0, 0.11, 0.0),
        4: (0.10, -0.48, 0.0), 5: (-0.10, -0.48, 0.0), 6: (0.0, 0.25, 0.0),
        7: (0.10, -0.90, -0.02), 8: (-0.10, -0.90, -0.02), 9: (0.0, 0.32, 0.0),
        10: (0.11, -0.95, 0.10), 11: (-0.11, -0.95, 0.10), 12: (0.0, 0.52, 0.0),
        13: (0.06, 0.44, 0.0), 14: (-0.06, 0.44, 0.0), 15: (0.0, 0.62, 0.02),
        16: (0.18, 0.46, 0.0), 17: (-0.18, 0.46, 0.0), 18: (0.44, 0.46, 0.0), 19: (-0.44, 0.46, 0.0),
        20: (0.69, 0.46, 0.0), 21: (-0.69, 0.46, 0.0), 22: (0.0, 0.60, 0.05),
        23: (0.03, 0.67, 0.08), 24: (-0.03, 0.67, 0.08),
    }
    for k, v in body.items():
        J[k] = torch.tensor(v, dtype=torch.float64)
    for side, wrist, base in ((1.0, 20, 25), (-1.0, 21, 40)):
        for f in range(5):
            for k in range(3):
                J[base + 3 * f + k] = J[wrist] + torch.tensor(
                    [side * (0.08 + 0.03 * k), 0.0, (f - 2) * 0.02], dtype=torch.float64)
    return J


def make_body_constants(seed: int = 1234, num_verts: int = NUM_VERTS) -> Dict[str, torch.Tensor]:
    """Random-init SMPL-X-topology constants in the package's canonical (smplx-module) layout.

    v_template [V,3]; shapedirs [V,3,20] ~ N(0,0.01^2); posedirs [486,3V] ~ N(0,0.001^2);
    J_regressor [55,V] row-stochastic, <=32 non-zeros/row; lbs_weights [V,55] row-stochastic,
    <=4 non-zeros/row; parents [55]; lh/rh_components [12,45] ~ N(0,0.1^2); pose_mean = 0;
    extra_vertex_ids [21] (the VertexJointSelector picks: 5 face, 6 feet, 10 finger tips).
    """
    g = torch.Generator().manual_seed(seed)
    V = num_verts
    J = _rest_joints()
    parents = torch.tensor(SMPLX_PARENTS, dtype=torch.int64)
    # vertices on capsules around the bones (joint -> parent segment), count ~ length * radius
    radius = torch.full((NUM_JOINTS,), 0.05, dtype=torch.float64)
    radius[[0, 3, 6, 9]] = 0.14
    radius[[1, 2, 4, 5]] = 0.07
    radius[[12, 15, 22]] = 0.08
    radius[25:] = 0.008
    radius[[23, 24]] = 0.012
    seg_a = J.clone()
    seg_b = J[torch.clamp(parents, min=0)].clone()
    length = (seg_a - seg_b).norm(dim=1) + 2 * radius
    share = length * radius
    counts = torch.floor(share / share.sum() * V).to(torch.int64)
    counts[0] += V - counts.sum()
    bone = torch.repeat_interleave(torch.arange(NUM_JOINTS), counts)
    u = torch.rand(V, generator=g, dtype=torch.float64)
    p = seg_b[bone] + (seg_a[bone] - seg_b[bone]) * u.unsqueeze(1)
    dirs = torch.randn(V, 3, generator=g, dtype=torch.float64)
    dirs = dirs / dirs.norm(dim=1, keepdim=True)
    v_template = p + dirs * radius[bone].unsqueeze(1)
    v_template = v_template[torch.randperm(V, generator=g)]            # mesh order is not spatial
    # skinning weights: 4 nearest joints, inverse-distance, row-stochastic
    dj = torch.cdist(v_template, J)                                      # [V,55]
    near_d, near_j = torch.topk(dj, 4, dim=1, largest=False)
    w = 1.0 / (near_d + 0.02) ** 2
    w = w / w.sum(dim=1, keepdim=True)
    lbs_weights = torch.zeros(V, NUM_JOINTS, dtype=torch.float64)
    lbs_weights.scatter_(1, near_j, w)
    # joint regressor: 32 nearest vertices per joint, inverse-distance, row-stochastic
    nd, nv = torch.topk(dj.t().contiguous(), 32, dim=1, largest=False)   # [55,32]
    wj = 1.0 / (nd + 0.01)
    wj = wj / wj.sum(dim=1, keepdim=True)
    J_regressor = torch.zeros(NUM_JOINTS, V, dtype=torch.float64)
    J_regressor.scatter_(1, nv, wj)
    shapedirs = torch.randn(V, 3, NUM_SHAPE, generator=g, dtype=torch.float64) * 0.01
    posedirs = torch.randn(NUM_POSE_FEAT, V * 3, generator=g, dtype=torch.float64) * 0.001
    lh = torch.randn(NUM_PCA, 45, generator=g, dtype=torch.float64) * 0.1
    rh = torch.randn(NUM_PCA, 45, generator=g, dtype=torch.float64) * 0.1
    extra = torch.randint(0, V, (21,), generator=g, dtype=torch.int64)
    f32 = lambda t: t.to(torch.float32).contiguous()
    return dict(v_template=f32(v_template), shapedirs=f32(shapedirs), posedirs=f32(posedirs),
                J_regressor=f32(J_regressor), parents=parents, lbs_weights=f32(lbs_weights),
                lh_components=f32(lh), rh_components=f32(rh),
                pose_mean=torch.zeros(165, dtype=torch.float32), extra_vertex_ids=extra)


def make_clip_params(T: int, seed: int = 1234) -> Dict[str, torch.Tensor]:
    """Per-frame SMPL-X parameters: smooth random walk (sigma 0.02 rad/frame, |aa| <= 0.6 rad),
    betas ~ N(0,1) constant over the clip, translation random walk, plus the world placement the
    reference optimises (scale, camera_ext [T,4,4], camera translation; global_optimization.py:
    179-185,191-206).  The body is y-up; camera_ext maps it into the z-up room, feet on the floor."""
    g = torch.Generator().manual_seed(seed)

    def walk(dim, sigma, clamp):
        steps = torch.randn(T, dim, generator=g) * sigma
        steps[0] = torch.randn(dim, generator=g) * 0.2
        return torch.clamp(torch.cumsum(steps, 0), -clamp, clamp)

    betas = torch.randn(1, 10, generator=g).repeat(T, 1)
    global_orient = walk(3, 0.02, 0.6)
    body_pose = walk(63, 0.02, 0.6)
    lh = walk(NUM_PCA, 0.02, 0.6)
    rh = walk(NUM_PCA, 0.02, 0.6)
    transl = walk(3, 0.01, 0.5)
    cam_transl = walk(3, 0.01, 0.5)
    # world placement: rotate y-up -> z-up, stand at a slowly moving floor position
    R = torch.tensor([[1.0, 0.0, 0.0], [0.0, 0.0, -1.0], [0.0, 1.0, 0.0]])
    pos = torch.cumsum(torch.randn(T, 2, generator=g) * 0.01, 0) + (torch.rand(2, generator=g) * 4 - 2)
    pos = torch.clamp(pos, -3.0, 3.0)
    cam_ext = torch.eye(4).repeat(T, 1, 1)
    cam_ext[:, :3, :3] = R
    cam_ext[:, 0, 3] = pos[:, 0]
    cam_ext[:, 1, 3] = pos[:, 1]
    cam_ext[:, 2, 3] = 1.0
    return dict(betas=betas, global_orient=global_orient, body_pose=body_pose,
                left_hand_pose=lh, right_hand_pose=rh, transl=transl,
                cam_transl=cam_transl, camera_ext=cam_ext, scale=torch.tensor(1.0))


def make_scene(M: int, kind: str = "uniform", seed: int = 1234) -> torch.Tensor:
    """Scene cloud [M,3] fp32 in the room box [-4,4] x [-4,4] x [0,3] m.

    kind='uniform' : uniform in the volume (configs 1, 2, 4)
    kind='surface' : jittered-grid blue-noise surrogate on the six box faces plus floor-standing boxes
                     (configs 3, 5, "Poisson-sampled")
    kind='lattice' : integer coordinates in [-8,8]^3 with duplicates (tie tests)
    """
    g = torch.Generator().manual_seed(seed)
    if kind == "uniform":
        p = torch.rand(M, 3, generator=g)
        lo = torch.tensor([-4.0, -4.0, 0.0])
        hi = torch.tensor([4.0, 4.0, 3.0])
        return (lo + p * (hi - lo)).contiguous()
    if kind == "lattice":
        return torch.randint(-8, 9, (M, 3), generator=g).to(torch.float32)
    if kind == "surface":
        # faces: floor, ceiling, 4 walls, and 4 boxes of furniture; area-proportional jittered grids
        rects = []  # (origin, edge_u, edge_v)
        def add(o, u, v):
            rects.append((torch.tensor(o), torch.tensor(u), torch.tensor(v)))
        add([-4.0, -4.0, 0.0], [8.0, 0.0, 0.0], [0.0, 8.0, 0.0])
        add([-4.0, -4.0, 3.0], [8.0, 0.0, 0.0], [0.0, 8.0, 0.0])
        add([-4.0, -4.0, 0.0], [8.0, 0.0, 0.0], [0.0, 0.0, 3.0])
        add([-4.0, 4.0, 0.0], [8.0, 0.0, 0.0], [0.0, 0.0, 3.0])
        add([-4.0, -4.0, 0.0], [0.0, 8.0, 0.0], [0.0, 0.0, 3.0])
        add([4.0, -4.0, 0.0], [0.0, 8.0, 0.0], [0.0, 0.0, 3.0])
        for cx, cy, sx, sy, h in [(-2.5, -2.0, 1.6, 0.8, 0.75), (2.0, 2.5, 0.9, 2.0, 0.45),
                                  (2.8, -2.8, 0.6, 0.6, 1.8), (-3.0, 2.0, 0.5, 1.5, 1.0)]:
            x0, y0 = cx - sx / 2, cy - sy / 2
            add([x0, y0, h], [sx, 0.0, 0.0], [0.0, sy, 0.0])
            add([x0, y0, 0.0], [sx, 0.0, 0.0], [0.0, 0.0, h])
            add([x0, y0 + sy, 0.0], [sx, 0.0, 0.0], [0.0, 0.0, h])
            add([x0, y0, 0.0], [0.0, sy, 0.0], [0.0, 0.0, h])
            add([x0 + sx, y0, 0.0], [0.0, sy, 0.0], [0.0, 0.0, h])
        areas = torch.tensor([float(u.norm() * v.norm()) for _, u, v in rects])
        cnt = torch.floor(areas / areas.sum() * M).to(torch.int64)
        cnt[0] += M - cnt.sum()
        out = []
        for (o, u, v), n in zip(rects, cnt.tolist()):
            if n == 0:
                continue
            lu, lv = float(u.norm()), float(v.norm())
            nu = max(1, int(round(math.sqrt(n * lu / lv))))
            nv = (n + nu - 1) // nu
            ii = torch.arange(nu * nv)[:n]
            cu = (ii % nu).to(torch.float32)
            cv = (ii // nu).to(torch.float32)
            ju = torch.rand(n, generator=g)
            jv = torch.rand(n, generator=g)
            s = ((cu + ju) / nu).unsqueeze(1)
            t = ((cv + jv) / nv).unsqueeze(1)
            out.append(o + s * u + t * v)
        pts = torch.cat(out, 0)
        return pts[torch.randperm(M, generator=g)].contiguous()
    raise ValueError(f"unknown scene kind {kind!r}")
