"""Torch restatement of the SMPL-X body-model forward (TEST / BASELINE INFRASTRUCTURE ONLY).

The reference calls a third-party package that is NOT vendored under /root/reference and is not
installable here (no network, licence-gated model files):

    [3P] `smplx` (pip; version UNPINNED: /root/reference/README.md:5 only says "same dependencies
    as SMPLify-X").  Call sites: /root/reference/global_optimization.py:154-168 (construction:
    model_type='smplx', gender='neutral', num_pca_comps=12, batch_size=T, every create_*=True),
    :280-283 / :333-335 / :396-398 (forward), optimization.py:107-121, global_vis.py:48-63,144.

What follows restates the PUBLISHED algorithm of smplx.lbs (blend_shapes, vertices2joints,
batch_rodrigues, batch_rigid_transform, lbs) and smplx.body_models.SMPLX.forward, the eight steps
of SURVEY.md section 8 row a4.  PARITY UNPINNED against the real package: there is no smplx
install, model file, golden vector or test in the reference to pin it to.  It is pinned instead by
(i) float64 self-consistency (finite-difference gradients, rigid-motion invariants in
tests/test_oracle_smplx.py) and (ii) the kinematic-tree shape algebra SURVEY.md probed.

Everything is dtype-generic so the same code gives the float64 "truth" and the float32 CPU
baseline that bench.py times.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

NUM_JOINTS = 55          # 22 body + jaw + 2 eyes + 2 x 15 hand joints
NUM_BODY_JOINTS = 21
NUM_HAND_JOINTS = 15


def batch_rodrigues(rot_vecs: torch.Tensor, epsilon: float = 1e-8) -> torch.Tensor:
    """[3P] smplx.lbs.batch_rodrigues: axis-angle [N,3] -> rotation matrices [N,3,3]."""
    n = rot_vecs.shape[0]
    dtype, device = rot_vecs.dtype, rot_vecs.device
    angle = torch.norm(rot_vecs + epsilon, dim=1, keepdim=True)
    rot_dir = rot_vecs / angle
    cos = torch.unsqueeze(torch.cos(angle), dim=1)
    sin = torch.unsqueeze(torch.sin(angle), dim=1)
    rx, ry, rz = torch.split(rot_dir, 1, dim=1)
    zeros = torch.zeros((n, 1), dtype=dtype, device=device)
    K = torch.cat([zeros, -rz, ry, rz, zeros, -rx, -ry, rx, zeros], dim=1).view((n, 3, 3))
    ident = torch.eye(3, dtype=dtype, device=device).unsqueeze(dim=0)
    return ident + sin * K + (1 - cos) * torch.bmm(K, K)


def blend_shapes(betas: torch.Tensor, shape_disps: torch.Tensor) -> torch.Tensor:
    """[3P] smplx.lbs.blend_shapes: [B,L] x [V,3,L] -> [B,V,3]."""
    return torch.einsum("bl,mkl->bmk", [betas, shape_disps])


def vertices2joints(J_regressor: torch.Tensor, vertices: torch.Tensor) -> torch.Tensor:
    """[3P] smplx.lbs.vertices2joints: [J,V] x [B,V,3] -> [B,J,3]."""
    return torch.einsum("bik,ji->bjk", [vertices, J_regressor])


def batch_rigid_transform(rot_mats, joints, parents):
    """[3P] smplx.lbs.batch_rigid_transform.

    rot_mats [B,J,3,3], joints [B,J,3], parents [J] -> (posed_joints [B,J,3], A [B,J,4,4]) where A is
    the world transform of every joint with the rest pose removed.
    """
    B, J = joints.shape[:2]
    dtype, device = joints.dtype, joints.device
    joints = torch.unsqueeze(joints, dim=-1)
    rel_joints = joints.clone()
    rel_joints[:, 1:] -= joints[:, parents[1:]]
    bottom = torch.zeros(B, J, 1, 4, dtype=dtype, device=device)
    bottom[..., 3] = 1
    transforms_mat = torch.cat([torch.cat([rot_mats, rel_joints], dim=-1), bottom], dim=-2)
    chain = [transforms_mat[:, 0]]
    for i in range(1, J):
        chain.append(torch.matmul(chain[int(parents[i])], transforms_mat[:, i]))
    transforms = torch.stack(chain, dim=1)
    posed_joints = transforms[:, :, :3, 3]
    joints_homogen = torch.nn.functional.pad(joints, [0, 0, 0, 1])
    rel_transforms = transforms - torch.nn.functional.pad(
        torch.matmul(transforms, joints_homogen), [3, 0, 0, 0, 0, 0, 0, 0])
    return posed_joints, rel_transforms


def lbs(betas, pose, v_template, shapedirs, posedirs, J_regressor, parents, lbs_weights):
    """[3P] smplx.lbs.lbs with pose2rot=True.

    betas [B,L]; pose [B,J*3] axis-angle; v_template [V,3]; shapedirs [V,3,L]; posedirs [(J-1)*9, V*3];
    J_regressor [J,V]; parents [J]; lbs_weights [V,J].
    Returns verts [B,V,3], posed joints [B,J,3], and the intermediates the backward tests use.
    """
    B = betas.shape[0]
    dtype, device = betas.dtype, betas.device
    v_shaped = v_template + blend_shapes(betas, shapedirs)                    # step 2
    J = vertices2joints(J_regressor, v_shaped)                                # step 3
    ident = torch.eye(3, dtype=dtype, device=device)
    rot_mats = batch_rodrigues(pose.view(-1, 3)).view([B, -1, 3, 3])          # step 4
    pose_feature = (rot_mats[:, 1:, :, :] - ident).view([B, -1])              # step 5
    pose_offsets = torch.matmul(pose_feature, posedirs).view(B, -1, 3)
    v_posed = pose_offsets + v_shaped
    J_transformed, A = batch_rigid_transform(rot_mats, J, parents)            # step 6
    W = lbs_weights.unsqueeze(dim=0).expand([B, -1, -1])                      # step 7
    num_joints = J_regressor.shape[0]
    T = torch.matmul(W, A.view(B, num_joints, 16)).view(B, -1, 4, 4)
    homogen_coord = torch.ones([B, v_posed.shape[1], 1], dtype=dtype, device=device)
    v_posed_homo = torch.cat([v_posed, homogen_coord], dim=2)
    v_homo = torch.matmul(T, torch.unsqueeze(v_posed_homo, dim=-1))
    verts = v_homo[:, :, :3, 0]
    return verts, J_transformed, dict(v_shaped=v_shaped, J=J, rot_mats=rot_mats,
                                      pose_feature=pose_feature, v_posed=v_posed, A=A)


def smplx_forward(model: Dict[str, torch.Tensor], *, betas, global_orient, body_pose, transl,
                  left_hand_pose, right_hand_pose, expression: Optional[torch.Tensor] = None,
                  jaw_pose=None, leye_pose=None, reye_pose=None, dtype=torch.float64,
                  return_intermediates: bool = False):
    """[3P] smplx.body_models.SMPLX.forward (use_pca=True, flat_hand_mean=False, no face contour).

    `model` holds the constants in the package's canonical layout (see
    4dcapture-fpv_b200/synthetic.py::make_body_constants): v_template [V,3], shapedirs [V,3,20]
    (10 shape + 10 expression), posedirs [486,3V], J_regressor [55,V], parents [55],
    lbs_weights [V,55], lh_components/rh_components [12,45], pose_mean [165],
    extra_vertex_ids [E] (vertex-picked extra joints, VertexJointSelector).
    Omitted expression / jaw / eye arguments default to zeros, exactly as the module's own
    zero-initialised parameters do in the reference call (global_optimization.py:280-283).
    Returns vertices [T,V,3] and joints [T,55+E,3], both with transl added.
    """
    c = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in model.items()}
    cast = lambda t: t.to(dtype)
    betas, global_orient, body_pose, transl = map(cast, (betas, global_orient, body_pose, transl))
    left_hand_pose, right_hand_pose = cast(left_hand_pose), cast(right_hand_pose)
    T = betas.shape[0]
    z3 = torch.zeros(T, 3, dtype=dtype)
    jaw_pose = cast(jaw_pose) if jaw_pose is not None else z3
    leye_pose = cast(leye_pose) if leye_pose is not None else z3
    reye_pose = cast(reye_pose) if reye_pose is not None else z3
    expression = cast(expression) if expression is not None else torch.zeros(T, 10, dtype=dtype)
    lh = torch.einsum("bi,ij->bj", [left_hand_pose, c["lh_components"]])      # step 1
    rh = torch.einsum("bi,ij->bj", [right_hand_pose, c["rh_components"]])
    full_pose = torch.cat([global_orient.reshape(-1, 1, 3),
                           body_pose.reshape(-1, NUM_BODY_JOINTS, 3),
                           jaw_pose.reshape(-1, 1, 3), leye_pose.reshape(-1, 1, 3),
                           reye_pose.reshape(-1, 1, 3),
                           lh.reshape(-1, NUM_HAND_JOINTS, 3),
                           rh.reshape(-1, NUM_HAND_JOINTS, 3)], dim=1).reshape(-1, 165)
    full_pose = full_pose + c["pose_mean"]
    shape_components = torch.cat([betas, expression], dim=-1)
    verts, joints, inter = lbs(shape_components, full_pose, c["v_template"], c["shapedirs"],
                               c["posedirs"], c["J_regressor"], c["parents"], c["lbs_weights"])
    extra = c.get("extra_vertex_ids")
    if extra is not None and extra.numel() > 0:                               # VertexJointSelector
        joints = torch.cat([joints, torch.index_select(verts, 1, extra)], dim=1)
    joints = joints + transl.unsqueeze(dim=1)                                 # step 8
    verts = verts + transl.unsqueeze(dim=1)
    if return_intermediates:
        inter["full_pose"] = full_pose
        return verts, joints, inter
    return verts, joints
