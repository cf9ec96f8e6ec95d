"""Scene-sharded chamfer across the GPUs of one box.

The scene cloud is split into contiguous index ranges, one per rank, resident for the whole fit.
Per step and per rank:
  body -> scene : the local search over the shard writes packed 64-bit keys
                  (float_bits(d) << 32 | GLOBAL index) straight from the kernel epilogue; the shards are combined by an
                  element-wise MIN over the keys -- the integer minimum IS the lexicographic (d, idx) minimum, so ties
                  still resolve to the lowest global index.  Two transports:
                    * comm = p2p.Mailbox : the search epilogue also stores every key into the peers' mailboxes over
                      NVLink (the transfer is fused into the search), a flag barrier, then every rank reduces its own
                      mailbox -- kernels only, capturable in a CUDA graph;
                    * comm = None        : one NCCL all_reduce(MIN) on the key tensor (the survey's baseline design);
  scene -> body : stays shard-local ([T, M/G] per rank, or reduced in the kernel by the fused form), no communication;
  backward      : each rank back-propagates its own shard (the scatter of its scene points, and the
                  body->scene term only for the queries whose winner it owns); the caller sums the
                  small PARAMETER gradients across ranks (allreduce_grads) instead of the 37.7 MB
                  vertex gradient.
For the host-side logic tests the same code runs on gloo with CPU tensors minus the CUDA kernels
(tests/test_sharded_gloo.py injects an oracle search).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib

KEY_IDX_MASK = 0xFFFFFFFF


def shard_range(M: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of scene indices owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(M, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def combine_keys(keys: torch.Tensor, group=None) -> torch.Tensor:
    """All-reduce(MIN) of packed (distance, index) keys, in place.  int64 view of the uint64 keys:
    canonical distances are >= 0, so the sign bit is clear and signed order == unsigned order."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(keys, op=dist.ReduceOp.MIN, group=group)
    return keys


def unpack_keys_torch(keys: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Device-agnostic unpack (used by the gloo tests; the CUDA path unpacks in fpv_p2p_min_unpack)."""
    d = (keys >> 32).to(torch.int32).view(torch.float32)
    i = (keys & KEY_IDX_MASK).to(torch.int64)
    return d, i


def pack_keys_torch(d: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    return (d.contiguous().view(torch.int32).to(torch.int64) << 32) | idx.to(torch.int64)


def _world(group) -> int:
    return dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1


class _ShardedChamferFn(torch.autograd.Function):
    """a [T,N,3] replicated; b_shard [1,Ms,3] this rank's scene range starting at global index idx_base.
    fused=True: the scene -> body output is the per-frame SUM over this rank's shard ([T]) instead of (d, i) [T,Ms]."""

    @staticmethod
    def forward(ctx, a, b_shard, idx_base: int, group, search: Optional[Callable], clip: bool, comm, opts, state,
                fused: bool):
        from . import chamfer, spatial
        a_c, b_c = a.contiguous(), b_shard.contiguous()
        T, N, _ = a_c.shape
        Ms = b_c.shape[1]
        ws = _world(group)
        ctx.fused = bool(fused)
        ctx.fix_shift = 0
        acc = None
        branch = None
        if search is not None:  # injected search (CPU oracle in the gloo tests): same combine logic, no CUDA
            if fused:
                raise RuntimeError("the fused scene->body form needs the CUDA path")
            d_loc, i_loc, d_b2a, i_b2a = search(a_c, b_c)
            keys = pack_keys_torch(d_loc, i_loc + idx_base)
            combine_keys(keys, group)
            d_a2b, i_a2b = unpack_keys_torch(keys)
        else:
            opts = opts or chamfer.DEFAULT_OPTIONS
            dev = a_c.device
            L = _lib.lib()
            use_spatial = opts.engine != "brute" and Ms >= opts.spatial_min_points
            if fused and not use_spatial:
                raise RuntimeError("the fused scene->body form needs the spatial engine (scene shard too small)")
            if use_spatial:
                scene = spatial.cached_scene(b_c)
                if state is None:
                    state = chamfer._default_state(scene, T, N)
                body = chamfer._body_cloud(a_c, scene, opts, state, clip, spheres=True)
                # body -> scene branch (search with the key epilogue, exchange, combine) on a second stream: it overlaps
                # the scene -> body search, and so does the wait for the slowest rank at its barrier
                use_box = comm is not None and ws > 1
                keys, d_a2b, i_a2b, stats, seed, seed_valid = chamfer._a2b_alloc(T, N, torch.int64, dev, opts, state,
                                                                                 want_keys=not use_box)
                branch = chamfer._SideBranch(state, dev, opts.overlap and (use_box or ws == 1))
                with branch:
                    if use_box:
                        # keys go from the search epilogue into slot `rank` of every mailbox; barrier; local min
                        spatial.culled_search_keys(body.sorted, T, scene, idx_base=idx_base, stats=stats, cand_orig=b_c,
                                                   seed=seed, seed_valid=seed_valid, keys=comm.key_slot(comm.rank, comm.rank),
                                                   push=comm.push_targets(), push_parity=comm.parity_keys,
                                                   push_half=comm.keys_half)
                        comm.combine_keys(T * N, N, body.perm_row(), out=(d_a2b, i_a2b))
                    else:
                        spatial.culled_search_keys(body.sorted, T, scene, idx_base=idx_base, stats=stats, cand_orig=b_c,
                                                   seed=seed, seed_valid=seed_valid, out=keys)
                        combine_keys(keys, group)
                        spatial.min_unpack(keys, 1, T * N, N, body.perm_row(), out=(d_a2b, i_a2b))
                d_a2b, i_a2b = d_a2b.view(T, N), i_a2b.view(T, N)
                state.stats["tiles_searched"] = stats
                if fused:
                    seed2, seed2_valid = state.seed_buffer("b2a", T, Ms, dev, True)
                    stats2 = torch.zeros(2, dtype=torch.int64, device=dev)
                    d_b2a = torch.empty(T, dtype=torch.float32, device=dev)          # per-frame sums over the shard
                    acc = torch.zeros(T, N, 4, dtype=torch.int64, device=dev)
                    ctx.fix_shift = scene.fix_shift()
                    with torch.cuda.device(dev):
                        wsb = _lib.workspace(L.fpv_nn_sphere_fused_workspace_bytes(T, Ms), dev)
                        pos, pos_shared = body.pos_table()
                        _lib.check(L.fpv_nn_sphere_fused(_lib.ptr(scene.sorted), T, Ms, _lib.ptr(body.planes),
                                                         _lib.ptr(body.boxes), _lib.ptr(body.oidx), _lib.ptr(pos), int(pos_shared),
                                                         _lib.ptr(seed2), int(seed2_valid), N, body.sphere_tile,
                                                         ctx.fix_shift, _lib.ptr(d_b2a), _lib.ptr(acc), _lib.ptr(stats2),
                                                         _lib.ptr(wsb), wsb.numel(), _lib.stream_ptr()),
                                   "fpv_nn_sphere_fused")
                    state.stats["tiles_searched_b2a"] = stats2
                    i_b2a = None
                else:
                    d_s2, i_s2 = chamfer._search_b2a(a_c, scene, body, torch.int64, opts, state)
                    if scene.identity:
                        d_b2a, i_b2a = d_s2, i_s2
                    else:
                        inv = scene.inv_perm[0]
                        d_b2a, i_b2a = d_s2.index_select(1, inv), i_s2.index_select(1, inv)
            else:
                planes_b = chamfer.pack_planes(b_c)
                keys = chamfer.nn_search(a_c, planes_b, Ms, ref_batches=1, idx_base=idx_base, want_keys=True)
                planes_a = chamfer.pack_planes(a_c)
                d_b2a, i_b2a = chamfer._nn_search_raw(b_c, True, T, Ms, planes_a, T, N, torch.int64)
                combine_keys(keys, group)
                d_a2b, i_a2b = chamfer.unpack_keys(keys, torch.int64)
        if branch is not None:
            branch.join()
        ctx.save_for_backward(a_c, b_c, i_b2a, i_a2b, acc)
        ctx.idx_base, ctx.world, ctx.search = idx_base, ws, search
        ctx.mark_non_differentiable(*[t for t in (i_b2a, i_a2b) if t is not None])
        ctx.set_materialize_grads(False)
        return d_b2a, d_a2b, i_b2a, i_a2b

    @staticmethod
    def backward(ctx, g_b2a, g_a2b, _1, _2):
        a, b, i_b2a, i_a2b, acc = ctx.saved_tensors
        none = (None,) * 10
        if not ctx.needs_input_grad[0] or (g_b2a is None and g_a2b is None):
            return none
        T, N, _ = a.shape
        Ms = b.shape[1]
        g2 = None
        i_loc = None
        if g_a2b is not None:
            # only the rank that owns the winning scene point back-propagates the body->scene term;
            # the replicated loss is scaled by 1/world on every rank, hence the factor `world` here.
            owned = (i_a2b >= ctx.idx_base) & (i_a2b < ctx.idx_base + Ms)
            g2 = torch.where(owned, g_a2b * float(ctx.world), torch.zeros_like(g_a2b)).contiguous()
            i_loc = torch.where(owned, i_a2b - ctx.idx_base, torch.zeros_like(i_a2b)).contiguous()
        if ctx.search is not None:  # CPU restatement for the gloo tests
            grad_a = torch.zeros_like(a)
            bb = b[0]
            if g2 is not None:
                grad_a += 2 * g2.unsqueeze(-1) * (a - bb[i_loc])
            if g_b2a is not None:
                g1 = g_b2a.contiguous()
                contrib = 2 * g1.unsqueeze(-1) * (torch.gather(a, 1, i_b2a.unsqueeze(-1).expand(-1, -1, 3)) - bb.unsqueeze(0))
                grad_a.scatter_add_(1, i_b2a.unsqueeze(-1).expand(-1, -1, 3), contrib)
            return (grad_a,) + (None,) * 9
        from . import chamfer
        L = _lib.lib()
        grad_a = torch.empty_like(a)
        dev = a.device
        if i_loc is None:
            i_loc = torch.zeros(1, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            if ctx.fused:
                have = g2 is not None
                if have:
                    wsb = _lib.workspace(L.fpv_chamfer_bwd_workspace_bytes(T, N, Ms, 1, 0), dev)
                    _lib.check(L.fpv_chamfer_bwd_bcast(_lib.ptr(a), _lib.ptr(b), T, N, Ms, 1, None, _lib.ptr(g2), 0,
                                                       _lib.ptr(i_loc), _lib.ptr(i_loc), 8, _lib.ptr(grad_a), None,
                                                       _lib.ptr(wsb), wsb.numel(), _lib.stream_ptr()), "fpv_chamfer_bwd")
                if g_b2a is not None:
                    gc = g_b2a.contiguous().float()
                    _lib.check(L.fpv_scene2body_grad(_lib.ptr(a), _lib.ptr(acc), ctx.fix_shift, _lib.ptr(gc), T, N,
                                                     _lib.ptr(grad_a), int(have), _lib.stream_ptr()), "fpv_scene2body_grad")
                elif not have:
                    grad_a.zero_()
            else:
                g1, bc1 = chamfer._weights(g_b2a)
                wsb = _lib.workspace(L.fpv_chamfer_bwd_workspace_bytes(T, N, Ms, 1, 0), dev)
                _lib.check(L.fpv_chamfer_bwd_bcast(_lib.ptr(a), _lib.ptr(b), T, N, Ms, 1, _lib.ptr(g1), _lib.ptr(g2), bc1,
                                                   _lib.ptr(i_b2a), _lib.ptr(i_loc), 8, _lib.ptr(grad_a), None,
                                                   _lib.ptr(wsb), wsb.numel(), _lib.stream_ptr()), "fpv_chamfer_bwd")
        return (grad_a,) + (None,) * 9


def distChamferSharded(a: torch.Tensor, b_shard: torch.Tensor, idx_base: int, group=None,
                       _search: Optional[Callable] = None, clip: bool = False, comm=None, options=None, state=None,
                       fused: bool = False):
    """distChamfer with the scene sharded over the ranks of `group`.

    Returns (d_b2a [T,Ms] for THIS rank's shard, d_a2b [T,N] combined over all shards,
             i_b2a [T,Ms] indices into a, i_a2b [T,N] GLOBAL scene indices).
    fused=True: d_b2a is instead the per-frame SUM of the shard's min distances ([T]; see chamfer.scene_to_body_sum)
    and i_b2a is None.  comm: a p2p.Mailbox moves the keys through peer memory from the search epilogue (kernels only,
    graph-capturable); None = NCCL all_reduce(MIN).
    Loss terms built on d_a2b (replicated on every rank) must be scaled by 1/world_size and terms on
    d_b2a normalised by the global count, so that sum-over-ranks of the local losses is the global loss
    and allreduce_grads() yields the global gradient.
    """
    if b_shard.dim() == 2:
        b_shard = b_shard.unsqueeze(0)
    return _ShardedChamferFn.apply(a, b_shard, int(idx_base), group, _search, bool(clip), comm, options, state, bool(fused))


def allreduce_grads(params, group=None, comm=None, extra: Optional[torch.Tensor] = None):
    """Sum the (small) parameter gradients across ranks: one flat exchange (p2p mailbox when comm is given, else one
    NCCL all_reduce).  extra: an optional tensor (e.g. the local loss) summed in the same exchange and returned."""
    if _world(group) == 1:
        return extra
    grads = [p.grad for p in params if p.grad is not None]
    parts = [g.reshape(-1) for g in grads] + ([extra.reshape(-1).float()] if extra is not None else [])
    if not parts:
        return extra
    flat = torch.cat(parts)
    if comm is not None:
        flat = comm.allreduce_sum(flat)
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    o = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[o:o + n].view_as(g))
        o += n
    if extra is not None:
        return flat[o:o + extra.numel()].view_as(extra).to(extra.dtype)
    return None
