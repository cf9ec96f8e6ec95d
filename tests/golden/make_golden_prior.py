"""Generates tests/golden/prior_*.npz from the REAL reference code (run in the build container, where /root/reference
exists).  The reference modules import packages that are absent here (torchgeometry, smplx, open3d, ...): they are
stubbed with empty modules, except `torchgeometry`, which is given oracle/prior_oracle.py's restatement of the three
functions the reference calls -- so these goldens pin the reference's OWN code (Gram-Schmidt decoder, 6D<->3D wiring,
cal_dctloss) and the parameter layout, not torchgeometry itself (stated in the oracle header).

    python tests/golden/make_golden_prior.py
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import prior_oracle as po  # noqa: E402

tgm = types.ModuleType("torchgeometry")
tgm.rotation_matrix_to_angle_axis = po.rotation_matrix_to_angle_axis
tgm.angle_axis_to_rotation_matrix = po.angle_axis_to_rotation_matrix
sys.modules["torchgeometry"] = tgm


class _Anything(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        sub = _Anything(self.__name__ + "." + name)
        sys.modules[sub.__name__] = sub
        return sub

    def __call__(self, *a, **k):
        return None


for name in ["smplx", "open3d", "human_body_prior", "human_body_prior.tools", "human_body_prior.tools.model_loader",
             "ChamferDistancePytorch", "ChamferDistancePytorch.dist_chamfer", "MotionGeneration", "trimesh", "pyrender",
             "cv2", "tqdm"]:
    if name not in sys.modules:
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = _Anything(name)
torch.cuda.LongTensor = torch.LongTensor
sys.path.insert(0, "/root/reference")
import cvae  # noqa: E402
import global_optimization as go  # noqa: E402

out = os.path.dirname(os.path.abspath(__file__))
g = torch.Generator().manual_seed(2024)

# ---- 6D decoder (pure reference code) and the 6D <-> 3D row conversion ----
x6 = torch.randn(64, 6, generator=g, dtype=torch.float64)
x6[0] = torch.tensor([1., 0., 0., 1., 0., 0.])                     # identity
x6[1] = torch.tensor([-1., 0., 0., -1., 0., 0.2])                  # near a half turn: exercises the other branches
x6[2] = torch.tensor([0., 1., -1., 0., 0., 0.3])
R = cvae.ContinousRotReprDecoder.decode(x6)
rows78 = torch.randn(16, 78, generator=g, dtype=torch.float64)
rows75 = go.convert_to_3D_rot(rows78)
rows78_back = go.convert_to_6D_rot(rows75)
np.savez(os.path.join(out, "prior_codec.npz"), x6=x6.numpy(), R=R.numpy(), rows78=rows78.numpy(), rows75=rows75.numpy(),
         rows78_back=rows78_back.numpy())

# ---- cal_dctloss: the literal method on a stand-in self ----
NB, Fr, K = go.NUM_BATCHES, go.BATCH_FRAME_NUM, go.DCT_NUM
n = torch.arange(Fr, dtype=torch.float64)
basis = torch.stack([torch.cos(np.pi * (n + 0.5) * k / Fr) * (np.sqrt(1.0 / Fr) if k == 0 else np.sqrt(2.0 / Fr))
                     for k in range(K)], dim=1)                      # orthonormal DCT-II, [F,K] (the .mat file is absent)
joints = (torch.randn(NB * Fr, 23, 3, generator=g, dtype=torch.float64) * 0.5).requires_grad_(True)
c_dct = torch.randn(NB, 23, 3, K, generator=g, dtype=torch.float64).requires_grad_(True)
fake = types.SimpleNamespace(dct_mtx=basis, c_dct=c_dct)
loss = go.FittingOP.cal_dctloss(fake, joints)
loss.backward()
np.savez(os.path.join(out, "prior_dct.npz"), basis=basis.numpy(), joints=joints.detach().numpy(), c_dct=c_dct.detach().numpy(),
         loss=loss.detach().numpy(), g_joints=joints.grad.numpy(), g_c=c_dct.grad.numpy())
print("wrote prior_codec.npz, prior_dct.npz; loss_dct =", float(loss.detach()))

# ---- the world placement: the literal verts_transform (:119-127) and FittingOP.body2world (:191-206) ----
torch.Tensor.cuda = lambda self, *a, **k: self                        # CPU shim for the per-frame `.cuda()` of :200
T = 9
verts = torch.randn(T, 57, 3, generator=g, dtype=torch.float32)
cam_ext = torch.eye(4).repeat(T, 1, 1) + 0.05 * torch.randn(T, 4, 4, generator=g)
cam_ext[:, 3, :] = torch.tensor([0.0, 0.0, 0.0, 1.0])
rec = torch.randn(T, 78, generator=g)
scale = torch.tensor([1.3])
fake = types.SimpleNamespace(body_rotation_rec=rec, num_body=T, scale=scale, camera_ext=cam_ext)
b2w = go.FittingOP.body2world(fake)
vt = go.verts_transform(verts * scale, b2w)
np.savez(os.path.join(out, "prior_world.npz"), verts=verts.numpy(), cam_ext=cam_ext.numpy(), rec=rec.numpy(), scale=scale.numpy(),
         b2w=b2w.numpy(), vt=vt.numpy())
print("wrote prior_world.npz")

# ---- host-side formats: the literal qvec2rotmat (:51-61) and body_params_parse (:64-76) ----
q = torch.randn(6, 4, generator=g, dtype=torch.float64)
q = (q / q.norm(dim=1, keepdim=True)).numpy()
Rq = np.stack([go.qvec2rotmat(v) for v in q])
keys = ["transl", "global_orient", "betas", "body_pose", "left_hand_pose", "right_hand_pose", "camera_translation"]
dims = [3, 3, 10, 32, 12, 12, 3]
frame = {k: torch.randn(1, n, generator=g).numpy() for k, n in zip(keys, dims)}
row = go.body_params_parse(frame)
np.savez(os.path.join(out, "prior_formats.npz"), q=q, Rq=Rq, row=row, **{"frame_" + k: v for k, v in frame.items()})
print("wrote prior_formats.npz")
