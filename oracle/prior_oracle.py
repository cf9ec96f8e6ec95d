"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch, any dtype; float64 in the tests) of the parameter front-end
and the DCT prior.  Imported by tests/ only; the product path never touches it.

Sources followed:
  * ContinousRotReprDecoder.decode          /root/reference/cvae.py:71-81   (pinned: tests/golden/prior_*.npz are
                                            outputs of the real class, see tests/golden/make_golden_prior.py)
  * convert_to_6D_rot / convert_to_3D_rot   /root/reference/global_optimization.py:96-115 (wiring pinned the same way)
  * FittingOP.cal_dctloss                   /root/reference/global_optimization.py:232-246 (pinned: literal method run
                                            on a stand-in self)
  * [3P] torchgeometry 0.1.2 (unpinned in the reference, README.md; NOT installed here):
        rotation_matrix_to_quaternion, quaternion_to_angle_axis, angle_axis_to_rotation_matrix
    restated from the published algorithm.  PARITY UNPINNED for these three: no copy of the package exists in this
    image, so the goldens of convert_to_3D_rot were produced with THIS restatement plugged in as `torchgeometry`.
  * [3P] human_body_prior v1 VPoser.decode  (unpinned, absent): Linear-lrelu(0.2)-[dropout, eval]-Linear-lrelu(0.2)-
    Linear -> ContinousRotReprDecoder -> matrot2aa.  PARITY UNPINNED (weights are licence-gated as well).
"""
import torch
import torch.nn.functional as F


# ---- [3P] torchgeometry restatement ------------------------------------------------------------------------
def rotation_matrix_to_quaternion(rotation_matrix: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """[N,3,4] -> [N,4] (w,x,y,z); four-branch selection on the TRANSPOSED matrix, as the package does."""
    m = torch.transpose(rotation_matrix, 1, 2)
    mask_d2 = m[:, 2, 2] < eps
    mask_d0_d1 = m[:, 0, 0] > m[:, 1, 1]
    mask_d0_nd1 = m[:, 0, 0] < -m[:, 1, 1]
    t0 = 1 + m[:, 0, 0] - m[:, 1, 1] - m[:, 2, 2]
    q0 = torch.stack([m[:, 1, 2] - m[:, 2, 1], t0, m[:, 0, 1] + m[:, 1, 0], m[:, 2, 0] + m[:, 0, 2]], -1)
    t1 = 1 - m[:, 0, 0] + m[:, 1, 1] - m[:, 2, 2]
    q1 = torch.stack([m[:, 2, 0] - m[:, 0, 2], m[:, 0, 1] + m[:, 1, 0], t1, m[:, 1, 2] + m[:, 2, 1]], -1)
    t2 = 1 - m[:, 0, 0] - m[:, 1, 1] + m[:, 2, 2]
    q2 = torch.stack([m[:, 0, 1] - m[:, 1, 0], m[:, 2, 0] + m[:, 0, 2], m[:, 1, 2] + m[:, 2, 1], t2], -1)
    t3 = 1 + m[:, 0, 0] + m[:, 1, 1] + m[:, 2, 2]
    q3 = torch.stack([t3, m[:, 1, 2] - m[:, 2, 1], m[:, 2, 0] - m[:, 0, 2], m[:, 0, 1] - m[:, 1, 0]], -1)
    c0 = (mask_d2 & mask_d0_d1).unsqueeze(-1)
    c1 = (mask_d2 & ~mask_d0_d1).unsqueeze(-1)
    c2 = (~mask_d2 & mask_d0_nd1).unsqueeze(-1)
    q = torch.where(c0, q0, torch.where(c1, q1, torch.where(c2, q2, q3)))
    t = torch.where(c0, t0.unsqueeze(-1), torch.where(c1, t1.unsqueeze(-1), torch.where(c2, t2.unsqueeze(-1), t3.unsqueeze(-1))))
    return q / torch.sqrt(t) * 0.5


def quaternion_to_angle_axis(q: torch.Tensor) -> torch.Tensor:
    q1, q2, q3 = q[..., 1], q[..., 2], q[..., 3]
    sin_sq = q1 * q1 + q2 * q2 + q3 * q3
    pos = sin_sq > 0.0
    sin_t = torch.sqrt(torch.where(pos, sin_sq, torch.ones_like(sin_sq)))   # keeps the unused branch finite
    cos_t = q[..., 0]
    two_theta = 2.0 * torch.where(cos_t < 0.0, torch.atan2(-sin_t, -cos_t), torch.atan2(sin_t, cos_t))
    k = torch.where(pos, two_theta / sin_t, 2.0 * torch.ones_like(sin_t))
    return torch.stack([q1 * k, q2 * k, q3 * k], -1)


def rotation_matrix_to_angle_axis(rotation_matrix: torch.Tensor) -> torch.Tensor:
    return quaternion_to_angle_axis(rotation_matrix_to_quaternion(rotation_matrix))


def angle_axis_to_rotation_matrix(angle_axis: torch.Tensor) -> torch.Tensor:
    """[N,3] -> [N,4,4]; Rodrigues with the package's (theta + 1e-6) normalisation and first-order Taylor branch."""
    eps = 1e-6
    theta2 = (angle_axis * angle_axis).sum(-1)
    theta = torch.sqrt(theta2)
    w = angle_axis / (theta + eps).unsqueeze(-1)
    wx, wy, wz = w[:, 0], w[:, 1], w[:, 2]
    c, s = torch.cos(theta), torch.sin(theta)
    k = 1.0 - c
    normal = torch.stack([c + wx * wx * k, wx * wy * k - wz * s, wy * s + wx * wz * k,
                          wz * s + wx * wy * k, c + wy * wy * k, -wx * s + wy * wz * k,
                          -wy * s + wx * wz * k, wx * s + wy * wz * k, c + wz * wz * k], -1).view(-1, 3, 3)
    rx, ry, rz = angle_axis[:, 0], angle_axis[:, 1], angle_axis[:, 2]
    one = torch.ones_like(rx)
    taylor = torch.stack([one, -rz, ry, rz, one, -rx, -ry, rx, one], -1).view(-1, 3, 3)
    rot = torch.where((theta2 > eps).view(-1, 1, 1), normal, taylor)
    out = torch.eye(4, dtype=angle_axis.dtype).repeat(angle_axis.shape[0], 1, 1)
    out[:, :3, :3] = rot
    return out


# ---- reference code paths ----------------------------------------------------------------------------------
def rot6d_decode(module_input: torch.Tensor) -> torch.Tensor:
    """cvae.py:71-81."""
    x = module_input.view(-1, 3, 2)
    b1 = F.normalize(x[:, :, 0], dim=1)
    dot = torch.sum(b1 * x[:, :, 1], dim=1, keepdim=True)
    b2 = F.normalize(x[:, :, 1] - dot * b1, dim=-1)
    b3 = torch.cross(b1, b2, dim=1)
    return torch.stack([b1, b2, b3], dim=-1)


def matrot2aa(pose_matrot: torch.Tensor) -> torch.Tensor:
    """cvae.py:84-93."""
    homogen = F.pad(pose_matrot.reshape(-1, 3, 3), [0, 1])
    return rotation_matrix_to_angle_axis(homogen).view(-1, 3).contiguous()


def aa2matrot(pose: torch.Tensor) -> torch.Tensor:
    """cvae.py:95-101."""
    return angle_axis_to_rotation_matrix(pose.reshape(-1, 3))[:, :3, :3].contiguous()


def rot6d_to_aa(x6: torch.Tensor) -> torch.Tensor:
    return matrot2aa(rot6d_decode(x6))


def aa_to_rot6d(aa: torch.Tensor) -> torch.Tensor:
    return aa2matrot(aa)[:, :, :-1].reshape(-1, 6)


def convert_to_6D_rot(x_batch: torch.Tensor) -> torch.Tensor:
    """global_optimization.py:96-104."""
    return torch.cat([x_batch[:, :3], aa_to_rot6d(x_batch[:, 3:6]), x_batch[:, 6:]], dim=-1)


def convert_to_3D_rot(x_batch: torch.Tensor) -> torch.Tensor:
    """global_optimization.py:107-115."""
    return torch.cat([x_batch[:, :3], rot6d_to_aa(x_batch[:, 3:9]), x_batch[:, 9:]], dim=-1)


def vposer_decode_aa(w, z: torch.Tensor) -> torch.Tensor:
    """[3P] VPoser v1 decode(z, output_type='aa'); w = dict(w1,b1,w2,b2,w3,b3) in nn.Linear layout. -> [T,1,J,3]"""
    h = F.leaky_relu(F.linear(z, w["w1"], w["b1"]), negative_slope=0.2)
    h = F.leaky_relu(F.linear(h, w["w2"], w["b2"]), negative_slope=0.2)
    y = F.linear(h, w["w3"], w["b3"])
    J = w["w3"].shape[0] // 6
    return matrot2aa(rot6d_decode(y)).view(z.shape[0], 1, J, 3)


def dct_loss(joints: torch.Tensor, dct_mtx: torch.Tensor, c_dct: torch.Tensor) -> torch.Tensor:
    """global_optimization.py:232-246 with the triple loop folded: joints [NB*F,J,3], dct_mtx [F,K], c_dct [NB,J,3,K]."""
    NB, J, A, K = c_dct.shape
    Fr = dct_mtx.shape[0]
    traj = joints[:NB * Fr].reshape(NB, Fr, J, A)
    hat = torch.einsum("fk,njak->nfja", dct_mtx, c_dct)
    err = (traj - hat) ** 2
    return (err / (err + 1.0)).sum(dim=1).mean()
