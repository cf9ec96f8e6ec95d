"""One global-optimisation step on synthetic data of the named shapes -- the unit bench.py times.

Mirrors FittingOP.cal_loss + loss.backward() of /root/reference/global_optimization.py
(:249-312, :560-593) restricted to the hot path SURVEY.md section 8 scopes in:

    body2world (:191-206)  ->  SMPL-X forward (:280-283)  ->  verts*scale, verts_transform (:284-285)
    ->  chamfer, both directions (:292-294 / chamfer_python.distChamfer)
    ->  robust contact mean on the contact vertices (:290,:295)
    ->  parameter 2nd-difference (:266-267), world-joint 1st-difference (:298-304),
        vertex 2nd-difference (cal_loss2 :404-405), reconstruction L1 (:259)
    ->  backward to the per-frame parameters, scale and camera_ext.

front_end=True adds the reference's parameter front-end and the `dct` prior (prior.py):
the optimised variable becomes the 78-D 6D-rotation row with the VPoser latent, decoded every step.  The default
(front_end=False) optimises the axis-angle parameter row directly.
mode="local" assembles FittingOP.cal_loss2 instead (:368-447, the second stage of fitting(mode='local')): reconstruction,
parameter 2nd-difference, VERTEX 2nd-difference over all 10,475 vertices and the contact-weighted leg velocity -- no chamfer.
optimizer_step() is the reference's Adam update (:188, :592) as capturable kernels; step(update=True) / a capture with
update=True ends with it, so the body moves from step to step as it does under the reference's optimiser.
With world_size > 1 the scene is sharded across ranks (sharded.py); the shards are combined through peer memory (p2p.py)
and the whole sharded step can be captured too.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist

from . import _lib, chamfer, p2p, prior, residuals, sharded
from .body_model import SMPLXB200
from .synthetic import make_body_constants, make_clip_params, make_scene

# column layout of the per-frame parameter row (cf. the reference's 75-D layout, cvae.py:196-202,
# with the 32-D VPoser latent replaced by the 63-D body pose it decodes to)
P_TRANSL, P_ORIENT, P_BETAS, P_POSE, P_LH, P_RH, P_CAM = (0, 3), (3, 6), (6, 16), (16, 79), (79, 91), (91, 103), (103, 106)
PARAM_DIM = 106

LOSS_WEIGHTS = dict(rec=1.0, contact=0.1, smoothing=1.0, world_smoothing=1.0, vert_smoothing=0.5, scene2body=0.1,
                    vposer=0.001, dct=0.0001)   # vposer: global_optimization.py:683; dct: the weight of the `dct` mode (:620)

# front_end=True: the reference's own optimisation variable (global_optimization.py:96-104, :454): the 78-D row
#   [transl 0:3 | rot6d 3:9 | betas 9:19 | vposer latent 19:51 | lh 51:63 | rh 63:75 | cam_transl 75:78]
# decoded every step by convert_to_3D_rot + VPoser.decode before the body model.
FRONT_END_DIM = 78
DCT_FRAMES, DCT_NUM = 60, 5   # BATCH_FRAME_NUM, DCT_NUM (global_optimization.py:41-44)


def pack_params(p: Dict[str, torch.Tensor]) -> torch.Tensor:
    return torch.cat([p["transl"], p["global_orient"], p["betas"], p["body_pose"], p["left_hand_pose"],
                      p["right_hand_pose"], p["cam_transl"]], dim=1).contiguous()


def contact_vertex_ids(constants: Dict[str, torch.Tensor]) -> torch.Tensor:
    """Synthetic stand-in for body_segments/{L_Leg,R_Leg}.json (absent; global_optimization.py:675,682):
    the vertices whose dominant skinning joint is a knee, ankle or foot."""
    dom = constants["lbs_weights"].argmax(dim=1)
    leg = torch.zeros(55, dtype=torch.bool)
    leg[[4, 5, 7, 8, 10, 11]] = True
    return torch.nonzero(leg[dom]).squeeze(1)


SHARD_BLOCK = 2048        # Morton-ordered scene: points per dealt block
SHARD_BLOCK_KD = 128      # k-d ordered scene: one query group (= two 64-point cells) per dealt block
ADAM = dict(lr=0.005, beta1=0.9, beta2=0.999, eps=1e-8)     # optim.Adam(..., lr=self.init_lr_h) :188, init_lr_h = 0.005 :671


def _morton_sorted(points: torch.Tensor) -> torch.Tensor:
    """[M,3] CPU cloud re-ordered along the Morton curve of its own bounding grid (same arithmetic as spatial.py)."""
    from . import spatial
    lo, inv_cell = spatial.grid_of(points)
    return points[torch.argsort(spatial.morton_keys(points, lo, inv_cell), stable=True)].contiguous()


def _kd_sorted(points: torch.Tensor) -> torch.Tensor:
    """[M,3] CPU cloud re-ordered as a left-balanced k-d partition: every aligned block of 64 * 2^k points is a cell."""
    from . import spatial
    out = points[spatial.kd_order(points, leaf=64, align="pow2")]
    # inside a 64-point cell: along the Morton curve, so that the 32 consecutive points a warp row holds are neighbours
    # (neighbours share their nearest body vertex: fewer groups per row in the accumulate pass)
    lo, inv_cell = spatial.grid_of(out)
    fine = spatial.morton_keys(out, lo, inv_cell)
    cell = torch.arange(out.shape[0]) // 64
    return out[torch.argsort(cell * (1 << 30) + fine, stable=True)].contiguous()


def _deal_blocks(M: int, world: int, blk: int) -> torch.Tensor:
    """Permutation of arange(M) that deals blocks of `blk` consecutive indices round-robin to `world` ranks and lays the
    ranks' shares out one after the other, each EXACTLY sharded.shard_range(M, world, r) long: whole blocks stay whole
    and aligned inside every share (a share that ends up a few points long hands its tail to the short ones)."""
    blocks = torch.arange(M).split(blk)
    share = [torch.cat(blocks[r::world]) if len(blocks) > r else torch.zeros(0, dtype=torch.int64) for r in range(world)]
    sizes = [e - b for b, e in (sharded.shard_range(M, world, r) for r in range(world))]
    pool = torch.cat([share[r][sizes[r]:] for r in range(world)])
    out = []
    for r in range(world):
        need = sizes[r] - min(sizes[r], share[r].numel())
        out.append(torch.cat([share[r][:sizes[r]], pool[:need]]))
        pool = pool[need:]
    return torch.cat(out)


def leg_vertex_ids(constants: Dict[str, torch.Tensor]):
    """Synthetic stand-ins for body_segments/L_Leg.json and R_Leg.json (:401-409): (left ids, right ids)."""
    dom = constants["lbs_weights"].argmax(dim=1)
    left = torch.zeros(55, dtype=torch.bool)
    right = torch.zeros(55, dtype=torch.bool)
    left[[4, 7, 10]] = True
    right[[5, 8, 11]] = True
    return torch.nonzero(left[dom]).squeeze(1), torch.nonzero(right[dom]).squeeze(1)


class FitProblem:
    """Synthetic clip + scene + body model on one device, and the per-step forward/backward."""

    def __init__(self, T: int, M: int, device, seed: int = 1234, scene_kind: str = "uniform",
                 rank: int = 0, world_size: int = 1, group=None, idx_dtype=torch.int64, presort_scene: bool = True,
                 front_end: bool = False, dct_frames: int = DCT_FRAMES, mode: str = "global", fused: bool = True,
                 comm: str = "p2p", options: Optional[chamfer.SearchOptions] = None, clips: int = 1,
                 scene_order: str = "kd", shard_frames: bool = True, scene_points: Optional[torch.Tensor] = None):
        """clips > 1: T is the TOTAL number of frames of `clips` independent clips of T/clips frames batched into one
        step (BASELINE.json configs[4]); the temporal residuals never couple frames of different clips.
        fused: scene -> body reduced inside the search kernel (chamfer.scene_to_body_sum) instead of materialising
        [T,M] distances and indices.  comm: "p2p" (peer-memory mailbox) or "nccl" for the sharded key / gradient exchange.
        scene_points: a ready host scene [M,3] in its final stored order (what a previous FitProblem built as
        .host_scene for the same M, order and world size) -- skips generation, ordering and dealing.
        shard_frames (sharded runs over the mailbox): the per-frame part of the step -- front-end, body model, world
        placement and their backward -- runs on T / world frames per rank; the vertices are all-gathered through the
        mailbox (backward: reduce-scatter of the vertex gradient), so only the searches' fixed costs stay replicated."""
        if mode not in ("global", "local", "reference"):
            raise RuntimeError("FitProblem: mode must be 'global' (cal_loss, both chamfer directions), 'local' (cal_loss2) "
                               "or 'reference' (the literal call sequence of cal_loss)")
        self.literal = mode == "reference"
        if self.literal:
            if world_size != 1 or clips != 1:
                raise RuntimeError("FitProblem: mode 'reference' is single-rank, single-clip (like the reference)")
            mode = "global"
        if T % clips != 0:
            raise RuntimeError("FitProblem: T must be a multiple of clips")
        self.T, self.M, self.device = T, M, torch.device(device)
        self.clips, self.Tc = clips, T // clips
        self.rank, self.world, self.group = rank, world_size, group
        self.idx_dtype = idx_dtype
        self.mode, self.fused = mode, bool(fused)
        self.options = options or chamfer.DEFAULT_OPTIONS
        self.search_state = chamfer.SearchState()          # seeds, frozen body order: per problem, never global
        self.report_terms = False                          # forward() also returns the individual (scaled) loss terms
        constants = make_body_constants(seed)
        self.constants = constants
        self.shard_frames = bool(shard_frames and world_size > 1 and comm == "p2p" and mode == "global"
                                 and self.device.type == "cuda")
        self.frame_ranges = [sharded.shard_range(T, world_size, r) for r in range(world_size)]
        self.f0, self.f1 = self.frame_ranges[rank] if self.shard_frames else (0, T)
        self.model = SMPLXB200(constants, batch_size=self.f1 - self.f0).to(self.device)
        # the module's own default parameters (jaw / eye poses, expression: zeros the loop never optimises, :154-168) get no
        # gradient: otherwise every backward splits and accumulates four more slices into buffers nobody reads
        self.model.requires_grad_(False)
        parts = [make_clip_params(self.Tc, seed + 1000 * c) for c in range(clips)]
        clip = {k: (torch.cat([p[k] for p in parts]) if parts[0][k].dim() > 0 else parts[0][k]) for k in parts[0]}
        self.front_end = front_end
        self.dct_batches = 0
        if front_end:
            self.vposer = prior.VPoserDecoderB200(prior.make_vposer_weights(seed)).to(self.device)
            g = torch.Generator().manual_seed(seed + 77)
            latent = torch.cat([torch.clamp(torch.cumsum(torch.randn(self.Tc, 32, generator=g) * 0.02, 0), -2.0, 2.0)
                                for _ in range(clips)])
            row75 = torch.cat([clip["transl"], clip["global_orient"], clip["betas"], latent, clip["left_hand_pose"],
                               clip["right_hand_pose"], clip["cam_transl"]], dim=1)
            # the observed data in the 6D form, converted once like the reference's loader does (:96-104)
            self.host_params = prior.convert_to_6D_rot(row75.to(self.device)).cpu().contiguous()      # [T,78]
            self.dct_batches = (self.Tc // dct_frames) * clips if mode == "global" else 0
            self.dct_frames = dct_frames
            if self.dct_batches:
                self.dct_mtx = prior.dct_basis(dct_frames, min(DCT_NUM, dct_frames), self.device)
                self.host_c_dct = torch.randn(self.dct_batches, 23, 3, self.dct_mtx.shape[1], generator=g)   # :186
        else:
            self.host_params = pack_params(clip)                       # [T,106] observed data (CPU)
        if scene_points is not None:
            if tuple(scene_points.shape) != (M, 3):
                raise RuntimeError("FitProblem: scene_points must be [M,3]")
            self.host_scene, presort_scene_now = scene_points, False
        else:
            self.host_scene, presort_scene_now = (make_scene(M, scene_kind, seed) if mode == "global" else torch.zeros(0, 3)), True
        self.host_camera_ext = clip["camera_ext"].clone()
        self.host_scale = clip["scale"].clone().reshape(1)
        self.contact_ids = contact_vertex_ids(constants).to(self.device)
        left, right = leg_vertex_ids(constants)
        self.left_ids, self.right_ids = left.to(self.device), right.to(self.device)
        # cal_loss2's per-frame contact weight (detect_contact, :315-365): a left/right stance pattern in [0,1]
        gw = torch.Generator().manual_seed(seed + 5)
        self.host_contact_weight = torch.clamp(0.5 + 0.6 * torch.sin(torch.arange(T) * 0.21 + torch.rand(1, generator=gw) * 6.28), 0, 1)
        self.begin, self.end = sharded.shard_range(M, world_size, rank)
        if presort_scene and presort_scene_now and mode == "global":
            # One-time host-side data preparation: the losses do not depend on the order of the scene points (every
            # term is a min / mean over them), so the scene is stored in a spatial order (k-d partition or Morton curve).
            # With several ranks the ordered points are dealt round-robin in blocks, so every rank's shard covers the
            # WHOLE room at full local density (a spatially contiguous shard leaves the ranks far from the body with
            # far-field-only work and unbalances the scene->body search: 58 vs 30 ms/step at 2 GPUs in round 1; dealing
            # single points thins every query group out).  k-d order: blocks of 128 points -- one query group of the
            # sphere search, two 64-point cells of the box search -- kept as they are: ~7,800 / world groups per rank
            # balance the heavy-tailed per-group cost statistically (blocks of 2048 left 15 % between two ranks).
            # Morton order: blocks of 2048, each shard then re-sorted on its own grid.
            kd = scene_order == "kd"
            order = _kd_sorted if kd else _morton_sorted
            whole = order(self.host_scene)
            if world_size > 1:
                whole = whole[_deal_blocks(M, world_size, SHARD_BLOCK_KD if kd else SHARD_BLOCK)]
            parts = [whole[slice(*sharded.shard_range(M, world_size, r))] for r in range(world_size)]
            self.host_scene = torch.cat(parts if kd else [order(p) for p in parts]).contiguous()
        self.scene_presorted = bool(presort_scene and mode == "global" and scene_order == "kd")
        self.comm = None
        if world_size > 1 and comm == "p2p" and self.device.type == "cuda":
            nfloats = self.host_params.numel() + 1 + 16 * T + 64 + (self.host_c_dct.numel() if self.dct_batches else 0)
            if self.shard_frames:   # the all-gather / reduce-scatter pieces: vertices + 23 joints of the widest frame range
                nv = constants["v_template"].shape[0]
                nfloats = max(nfloats, max(e - b for b, e in self.frame_ranges) * (nv + 23) * 3)
            self.comm = p2p.Mailbox(self.device, group, key_capacity=T * constants["v_template"].shape[0], float_capacity=nfloats)
        self._adam = None
        self.upload()

    def upload(self, non_blocking: bool = False):
        """Host -> device copies of every per-step input (the eager e2e leg calls this inside the timed region)."""
        dev = self.device
        self.data = self.host_params.to(dev, non_blocking=non_blocking)
        g = torch.Generator().manual_seed(99)
        if not hasattr(self, "_host_init"):
            # the optimised copy starts a small perturbation away from the data, like an outlier-fixed init
            self._host_init = self.host_params + 0.01 * torch.randn(self.host_params.shape, generator=g)
            if dev.type == "cuda":
                self.host_params = self.host_params.pin_memory()
                self._host_init = self._host_init.pin_memory()
                if self.host_scene.numel():
                    self.host_scene = self.host_scene.pin_memory()
                self.host_camera_ext = self.host_camera_ext.pin_memory()
                self.host_scale = self.host_scale.pin_memory()
        self.params = self._host_init.to(dev, non_blocking=non_blocking).requires_grad_(True)
        self.scale = self.host_scale.to(dev, non_blocking=non_blocking).requires_grad_(True)
        self.camera_ext = self.host_camera_ext.to(dev, non_blocking=non_blocking).requires_grad_(True)
        if self.mode == "global":
            self.scene = self.host_scene[self.begin:self.end].to(dev, non_blocking=non_blocking).unsqueeze(0)
            if self.scene_presorted and dev.type == "cuda":
                from . import spatial
                spatial.cached_scene(self.scene, presorted=True)     # index it in the order it arrives in
        self.contact_weight = self.host_contact_weight.to(dev, non_blocking=non_blocking)
        if self.front_end and self.dct_batches:
            if dev.type == "cuda" and not self.host_c_dct.is_pinned():
                self.host_c_dct = self.host_c_dct.pin_memory()
            self.c_dct = self.host_c_dct.to(dev, non_blocking=non_blocking).requires_grad_(True)
        self._adam = None
        return self

    def h2d_bytes(self) -> int:
        extra = self.host_c_dct.numel() if (self.front_end and self.dct_batches) else 0
        scene = (self.end - self.begin) * 3 if self.mode == "global" else 0
        return 4 * (self.host_params.numel() * 2 + self.host_scale.numel() + self.host_camera_ext.numel() + scene + extra)

    def d2h_bytes(self) -> int:
        return 4 * (1 + sum(t.numel() for t in self.leaves()))

    def leaves(self):
        extra = [self.c_dct] if (self.front_end and self.dct_batches) else []
        return [self.params, self.scale, self.camera_ext] + extra

    def _refresh_leaves(self) -> None:
        self.params = self.params.detach().requires_grad_(True)
        self.scale = self.scale.detach().requires_grad_(True)
        self.camera_ext = self.camera_ext.detach().requires_grad_(True)
        if self.front_end and self.dct_batches:
            self.c_dct = self.c_dct.detach().requires_grad_(True)

    # ---- temporal residuals, clip-aware: frames of different clips are never differenced ----
    def _per_clip(self, fn, x, *rest):
        if self.clips == 1:
            return fn(x, *rest)
        Tc = self.Tc
        vals = [fn(x[c * Tc:(c + 1) * Tc], *[r[c * Tc:(c + 1) * Tc] for r in rest]) for c in range(self.clips)]
        return torch.stack(vals).mean()

    def _body(self):
        """Front-end + body model + world placement: (vertices [T,V,3], joints [T,23,3], extra losses).  With
        shard_frames every rank runs its own frame range and the results are all-gathered through the mailbox."""
        f0, f1 = self.f0, self.f1
        Tr = f1 - f0
        p = self.params[f0:f1] if self.shard_frames else self.params
        cam_ext = self.camera_ext[f0:f1] if self.shard_frames else self.camera_ext
        sl = lambda r: p[:, r[0]:r[1]]
        inv_world = 1.0 / self.world
        extra_losses = {}
        if self.front_end:
            # global_optimization.py:261-283: 6D row -> axis-angle row -> VPoser decode -> body model
            bp = prior.front_end_split(p)
            z = bp.pop("body_pose_vp")
            cam_transl = bp.pop("camera_translation")
            if self.mode == "global":
                # mean over ALL frames: a rank that holds only its own rows contributes its share of the sum
                extra_losses["vposer"] = (torch.sum(z ** 2) / float(self.T * z.shape[1])) if self.shard_frames \
                    else torch.mean(z ** 2) * inv_world
            b2w = residuals.body2world(cam_transl, self.scale, cam_ext)
            out = self.model(return_verts=True, body_pose=self.vposer.decode(z, output_type="aa").view(Tr, -1), **bp)
        else:
            b2w = residuals.body2world(sl(P_CAM), self.scale, cam_ext)
            out = self.model(return_verts=True, body_pose=sl(P_POSE), transl=sl(P_TRANSL),
                             global_orient=sl(P_ORIENT), betas=sl(P_BETAS),
                             left_hand_pose=sl(P_LH), right_hand_pose=sl(P_RH))
        verts = residuals.verts_transform(out.vertices * self.scale, b2w)
        # joints are transformed UNSCALED, as in the reference (only the vertices are multiplied by scale, :284-285 vs :296-297)
        joints = residuals.verts_transform(out.joints[:, 0:23, :].contiguous(), b2w)
        if self.shard_frames:
            V = verts.shape[1]
            rows = torch.cat([verts.reshape(Tr, V * 3), joints.reshape(Tr, 69)], dim=1)
            full = p2p.all_gather_rows(rows, self.comm, self.frame_ranges)
            verts = full[:, :V * 3].reshape(self.T, V, 3)
            joints = full[:, V * 3:].reshape(self.T, 23, 3)
        return verts, joints, extra_losses

    def forward_reference(self) -> Dict[str, torch.Tensor]:
        """The reference's LITERAL call sequence (FittingOP.cal_loss, global_optimization.py:249-312, and the first-stage
        loss of fitting(mode='global'), :570): the scene as the materialised .repeat(T,1,1) copy of :176, the contact
        term through ext.chamferDist()(contact_verts.contiguous(), s_verts_batch.contiguous()) with its second value
        discarded (:290-294), every term of cal_loss computed, loss = 0.1 contact + smoothing + rec."""
        p = self.params
        verts, joints, extra = self._body()
        if not hasattr(self, "s_verts_batch"):
            self.s_verts_batch = self.scene.repeat(self.T, 1, 1)                                   # :175-176
        body_verts_contact_batch = verts[:, self.contact_ids, :]                                   # :290
        contact_dist, _ = chamfer.chamferDist()(body_verts_contact_batch.contiguous(), self.s_verts_batch.contiguous())
        losses = {
            "rec": torch.mean(torch.abs(self.data - p)),                                           # :259
            "vposer": extra.get("vposer", p.new_zeros(())),                                        # :262-263
            "smoothing": residuals.second_diff_l1(p),                                              # :266-267
            "contact": residuals.contact_robust_loss(contact_dist),                                # :295
            "world_smoothing": residuals.first_diff_l1(joints),                                    # :304
        }
        if self.front_end and self.dct_batches:
            losses["dct"] = self._dct(joints)                                                      # :310
        losses["total"] = 0.1 * losses["contact"] + losses["smoothing"] + losses["rec"]           # :570
        return losses

    def forward(self) -> Dict[str, torch.Tensor]:
        if self.mode == "local":
            return self.forward_local()
        if self.literal:
            return self.forward_reference()
        p, W = self.params, LOSS_WEIGHTS
        inv_world = 1.0 / self.world
        verts, joints, extra_losses = self._body()
        if self.world > 1:
            s_b2a, d_a2b, _, _ = sharded.distChamferSharded(verts, self.scene, self.begin, self.group, clip=True,
                                                            comm=self.comm, options=self.options, state=self.search_state,
                                                            fused=self.fused)
        elif self.fused:
            s_b2a, d_a2b, _ = chamfer.fit_chamfer_terms(verts, self.scene, clip=True, options=self.options,
                                                        state=self.search_state, idx_dtype=self.idx_dtype)
        else:
            s_b2a, d_a2b, _, _ = chamfer.distChamfer(verts, self.scene, idx_dtype=self.idx_dtype, clip=True,
                                                     options=self.options, state=self.search_state)
        # raw terms; each is scaled so that the sum over ranks of the local totals is the global loss: replicated terms by
        # 1/world, the shard-local scene -> body sum by the GLOBAL point count, the frame-sharded VPoser term already is
        # a share.  The weighted total is ONE stack . weights product (it replaced ~25 scalar multiply / add launches
        # and as many in the backward).
        terms = {
            "rec": (torch.mean(torch.abs(self.data - p)), inv_world),
            "smoothing": (self._per_clip(residuals.second_diff_l1, p), inv_world),
            "contact": (residuals.contact_robust_loss(d_a2b.index_select(1, self.contact_ids)), inv_world),
            "scene2body": (s_b2a.sum(), 1.0 / float(self.T * self.M)),
            "world_smoothing": (self._per_clip(residuals.first_diff_l1, joints), inv_world),
            "vert_smoothing": (self._per_clip(residuals.second_diff_l1, verts), inv_world),
        }
        if self.front_end and self.dct_batches:
            terms["dct"] = (self._dct(joints), inv_world)      # :310
        for k, v in extra_losses.items():
            terms[k] = (v, 1.0)                                  # (already scaled in _body)
        keys = tuple(terms)
        if getattr(self, "_loss_keys", None) != keys:
            self._loss_keys = keys
            self._loss_w = torch.tensor([W[k] * terms[k][1] for k in keys], dtype=torch.float32, device=self.device)
        stacked = torch.stack([terms[k][0].reshape(()) for k in keys])
        losses = {k: stacked[i].detach() * terms[k][1] for i, k in enumerate(keys)} if self.report_terms else {}
        losses["total"] = torch.dot(stacked, self._loss_w)
        return losses

    def _dct(self, joints):
        if self.clips == 1:
            return prior.cal_dctloss(joints, self.dct_mtx, self.c_dct)
        per = self.dct_batches // self.clips
        vals = [prior.cal_dctloss(joints[c * self.Tc:(c + 1) * self.Tc], self.dct_mtx, self.c_dct[c * per:(c + 1) * per])
                for c in range(self.clips)]
        return torch.stack(vals).mean()

    def forward_local(self) -> Dict[str, torch.Tensor]:
        """FittingOP.cal_loss2 (global_optimization.py:368-447): loss = loss_smoothing + loss_local_smoothing + loss_rec
        + loss_contact_smoothing (:549); no chamfer (its block is commented out, :432-445).  Not sharded (replicas)."""
        p = self.params
        verts, _, _ = self._body()
        w_right = self.contact_weight.clone()                       # :411-416
        w_left = 1.0 - w_right
        w_left = torch.where(w_left < 0.5, torch.zeros_like(w_left), w_left)
        w_right = torch.where(w_right < 0.5, torch.zeros_like(w_right), w_right)
        vl = verts.index_select(1, self.left_ids)
        vr = verts.index_select(1, self.right_ids)
        losses = {
            "rec": torch.mean(torch.abs(self.data - p)),                                        # :376
            "local_smoothing": self._per_clip(residuals.second_diff_l1, p),                     # :381-382
            "smoothing": self._per_clip(residuals.second_diff_l1, verts),                       # :404-405
            "contact_smoothing": self._per_clip(residuals.first_diff_l1, vl, w_left)            # :415-429
            + self._per_clip(residuals.first_diff_l1, vr, w_right),
        }
        losses["total"] = losses["rec"] + losses["local_smoothing"] + losses["smoothing"] + losses["contact_smoothing"]
        return losses

    # ---- the optimiser update (global_optimization.py:188, :592) ----
    def trainable(self):
        """The leaves the update touches: the first stage of fitting() trains body_rotation_rec and scale
        (:565-568; camera_ext and c_dct are frozen there); mode 'local' trains body_rotation_rec only (:543-546)."""
        return [self.params] if self.mode == "local" else [self.params, self.scale]

    def optimizer_step(self) -> None:
        """One torch.optim.Adam step (lr 0.005) on the trainable leaves, as kernels with a device-side step counter."""
        L = _lib.lib()
        if self._adam is None:
            self._adam = {"step": torch.zeros(1, dtype=torch.float32, device=self.device),
                          "state": [(torch.zeros_like(t), torch.zeros_like(t)) for t in self.trainable()]}
        with torch.cuda.device(self.device), torch.no_grad():
            _lib.check(L.fpv_adam_tick(_lib.ptr(self._adam["step"]), _lib.stream_ptr()), "fpv_adam_tick")
            for t, (m, v) in zip(self.trainable(), self._adam["state"]):
                _lib.check(L.fpv_adam_update(_lib.ptr(t), _lib.ptr(t.grad), _lib.ptr(m), _lib.ptr(v), t.numel(),
                                             ADAM["lr"], ADAM["beta1"], ADAM["beta2"], ADAM["eps"],
                                             _lib.ptr(self._adam["step"]), _lib.stream_ptr()), "fpv_adam_update")

    def _backward_and_reduce(self, losses):
        """backward, then (sharded) ONE exchange that sums the parameter gradients and the loss over the ranks."""
        losses["total"].backward()
        loss = losses["total"].detach()
        if self.world > 1 and self.mode == "global":
            loss = sharded.allreduce_grads(self.leaves(), self.group, comm=self.comm, extra=loss.reshape(1)).reshape(())
        return loss

    def step(self, update: bool = False) -> torch.Tensor:
        """zero_grad -> forward -> backward (-> gradient + loss exchange) (-> Adam update).  Returns the detached
        total loss (the GLOBAL loss on every rank when sharded)."""
        for t in self.leaves():
            t.grad = None
        loss = self._backward_and_reduce(self.forward())
        if update:
            self.optimizer_step()
        return loss

    def capture(self, warmup: int = 3, update: bool = False):
        """Capture zero_grad -> forward -> backward (-> exchange) (-> Adam update) as ONE CUDA graph (SURVEY.md section
        8f, row f1): no host synchronisation, no allocation at replay.  Returns self; step_graph() replays.

        Everything the step touches is capture-safe by construction: the library entry points only enqueue work
        on the caller's stream into caller-provided workspaces, the scene index, the body ordering and the seed buffers
        are cached (built during warm-up; the kernels update the seeds in place), the per-step tensors come from torch's
        graph-private pool.  After capture the leaves, the observed data and the scene are STATIC buffers: change them
        in place (copy_), never rebind them -- and the scene must stay what it was (its index is not part of the graph;
        spatial.invalidate_scene after refilling it).
        Sharded steps capture too when the shards are combined through the p2p mailbox (kernels only); with NCCL as the
        transport (comm="nccl") the step stays eager: capturing the NCCL collectives deadlocked in round 1."""
        if self.world != 1 and self.comm is None and self.mode == "global":
            raise RuntimeError("FitProblem.capture: the sharded step is capturable with comm='p2p' only")
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            # Fresh leaf tensors on the same storage.  A leaf's AccumulateGrad node keeps the stream it was CREATED on and
            # is reused for as long as anything holds it -- the autograd worker thread still holds the last node of a
            # backward for a moment after backward() returns, so the node of an eager step on the default stream can be
            # handed from step to step.  During capture the engine would then enqueue a wait on the legacy stream:
            # cudaErrorStreamCaptureImplicit (round 1: seen under compute-sanitizer, where that window is wide).  New
            # leaves get their node in the first warm-up step below, on the capture stream.  (Holders of the old tensor
            # objects still see the values -- same storage -- but not the .grad.)
            self._refresh_leaves()
            for _ in range(max(1, warmup)):
                self.step(update=update)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        for t in self.leaves():
            t.grad = None
        self._graph = torch.cuda.CUDAGraph()
        # capture on the stream the warm-up ran on: the leaves' AccumulateGrad nodes then live on the capture stream and
        # autograd inserts no dependency on the legacy default stream (illegal while a stream is capturing)
        with torch.cuda.graph(self._graph, stream=side):
            self._graph_loss = self._backward_and_reduce(self.forward())
            if update:
                self.optimizer_step()
        self._graph_update = update
        return self

    def step_graph(self) -> torch.Tensor:
        """Replay the captured step: gradients land in the leaves' .grad (static buffers), returns the (global) loss
        tensor (static; overwritten by the next replay)."""
        self._graph.replay()
        return self._graph_loss

    def step_e2e_graph(self):
        """End to end through the captured step.  The observed data (the step's input) goes from pinned host memory
        into the static device buffer, the graph is replayed, the loss and the gradients (update=False capture) or the
        updated leaves (update=True: the optimiser state lives on the device, as in the reference) come back to the
        host.  The scene is resident, as in the reference (uploaded once before the loop, :173-176)."""
        with torch.no_grad():
            self.data.copy_(self.host_params, non_blocking=True)
            if not self._graph_update:
                self.params.copy_(self._host_init, non_blocking=True)
                self.scale.copy_(self.host_scale, non_blocking=True)
                self.camera_ext.copy_(self.host_camera_ext, non_blocking=True)
                if self.front_end and self.dct_batches:
                    self.c_dct.copy_(self.host_c_dct, non_blocking=True)
        loss = self.step_graph()
        outs = [t.detach() for t in self.leaves()] if self._graph_update else [t.grad for t in self.leaves()]
        host = [loss.to("cpu", non_blocking=True)] + [t.to("cpu", non_blocking=True) for t in outs]
        torch.cuda.synchronize(self.device)
        return host

    def h2d_bytes_graph(self) -> int:
        if getattr(self, "_graph_update", False):
            return 4 * self.host_params.numel()
        scene = (self.end - self.begin) * 3 if self.mode == "global" else 0
        return self.h2d_bytes() - 4 * scene

    def step_e2e(self):
        """The same step from HOST buffers: pinned inputs -> device, step, loss + gradients -> host."""
        self.upload(non_blocking=True)
        loss = self.step()
        host = [loss.to("cpu", non_blocking=True)] + [t.grad.to("cpu", non_blocking=True) for t in self.leaves()]
        torch.cuda.synchronize(self.device)
        return host

    def close(self):
        if self.comm is not None:
            self.comm.close()
            self.comm = None
