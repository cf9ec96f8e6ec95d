"""Scene-sharded chamfer across the GPUs of one box.

The scene cloud is split into contiguous index ranges, one per rank, resident for the whole fit.
Per step and per rank:
  body -> scene : local search over the shard writes packed 64-bit keys
                  (float_bits(d) << 32 | GLOBAL index) straight from the kernel epilogue;
                  one all-reduce(MIN) over the keys combines the shards -- the integer minimum IS the
                  lexicographic (d, idx) minimum, so ties still resolve to the lowest global index;
  scene -> body : stays shard-local ([T, M/G] per rank), no communication;
  backward      : each rank back-propagates its own shard (the scatter of its scene points, and the
                  body->scene term only for the queries whose winner it owns); the caller sums the
                  small PARAMETER gradients across ranks (allreduce_grads) instead of the 37.7 MB
                  vertex gradient.
Runs on NCCL (GPU) and, for the host-side logic tests, on gloo with CPU tensors through the same
code path minus the CUDA kernels (see tests/test_sharded_gloo.py, which injects an oracle search).
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
    """Device-agnostic unpack (used by the gloo tests; the CUDA path uses fpv_nn_unpack_keys)."""
    d = (keys >> 32).to(torch.int32).view(torch.float32)
    i = (keys & KEY_IDX_MASK).to(torch.int64)
    return d, i


def pack_keys_torch(d: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    return (d.contiguous().view(torch.int32).to(torch.int64) << 32) | idx.to(torch.int64)


class _ShardedChamferFn(torch.autograd.Function):
    """a [T,N,3] replicated; b_shard [1,Ms,3] this rank's scene range starting at global index idx_base."""

    @staticmethod
    def forward(ctx, a, b_shard, idx_base: int, group, search: Optional[Callable], clip: bool = False):
        from . import chamfer
        a_c, b_c = a.contiguous(), b_shard.contiguous()
        T, N, _ = a_c.shape
        Ms = b_c.shape[1]
        ws = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        if search is None:
            if chamfer.ENGINE != "brute" and Ms >= chamfer.SPATIAL_MIN_POINTS:
                # shard-local search through the spatial index (global indices via idx_base), then the combine
                d_b2a, d_loc, i_b2a, i_glob = chamfer._forward_spatial(a_c, b_c, torch.int64, idx_base, clip=clip)
                chamfer.LAST_STATS.pop("sorted", None)
                keys = pack_keys_torch(d_loc, i_glob)
            else:
                planes_b = chamfer.pack_planes(b_c)
                keys = chamfer.nn_search(a_c, planes_b, Ms, ref_batches=1, idx_base=idx_base, want_keys=True)
                planes_a = chamfer.pack_planes(a_c)
                L = _lib.lib()
                d_b2a = torch.empty(T, Ms, dtype=torch.float32, device=a_c.device)
                i_b2a = torch.empty(T, Ms, dtype=torch.int64, device=a_c.device)
                with torch.cuda.device(a_c.device):
                    wsb = _lib.workspace(L.fpv_nn_search_workspace_bytes(T, Ms, N), a_c.device)
                    _lib.check(L.fpv_nn_search(_lib.ptr(b_c), 1, T, Ms, _lib.ptr(planes_a), T, N, 0, _lib.ptr(d_b2a),
                                               _lib.ptr(i_b2a), 8, None, _lib.ptr(wsb), wsb.numel(),
                                               _lib.stream_ptr()), "fpv_nn_search")
            combine_keys(keys, group)
            d_a2b, i_a2b = chamfer.unpack_keys(keys, torch.int64)
        else:  # injected search (CPU oracle in the gloo tests): same combine logic, no CUDA
            d_loc, i_loc, d_b2a, i_b2a = search(a_c, b_c)
            keys = pack_keys_torch(d_loc, i_loc + idx_base)
            combine_keys(keys, group)
            d_a2b, i_a2b = unpack_keys_torch(keys)
        ctx.save_for_backward(a_c, b_c, i_b2a, i_a2b)
        ctx.idx_base, ctx.world, ctx.search = idx_base, ws, search
        ctx.mark_non_differentiable(i_b2a, i_a2b)
        ctx.set_materialize_grads(False)
        return d_b2a, d_a2b, i_b2a, i_a2b

    @staticmethod
    def backward(ctx, g_b2a, g_a2b, _1, _2):
        a, b, i_b2a, i_a2b = ctx.saved_tensors
        if not ctx.needs_input_grad[0] or (g_b2a is None and g_a2b is None):
            return None, None, None, None, None, None
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
        if ctx.search is None:
            from . import chamfer
            g1, bc1 = chamfer._weights(g_b2a)
        else:
            g1, bc1 = (g_b2a.contiguous() if g_b2a is not None else None), 0
        if ctx.search is not None:  # CPU restatement for the gloo tests
            grad_a = torch.zeros_like(a)
            bb = b[0]
            if g2 is not None:
                grad_a += 2 * g2.unsqueeze(-1) * (a - bb[i_loc])
            if g1 is not None:
                contrib = 2 * g1.unsqueeze(-1) * (torch.gather(a, 1, i_b2a.unsqueeze(-1).expand(-1, -1, 3)) - bb.unsqueeze(0))
                grad_a.scatter_add_(1, i_b2a.unsqueeze(-1).expand(-1, -1, 3), contrib)
            return grad_a, None, None, None, None, None
        L = _lib.lib()
        grad_a = torch.empty_like(a)
        if i_loc is None:
            i_loc = torch.zeros(T, N, dtype=torch.int64, device=a.device)
        with torch.cuda.device(a.device):
            wsb = _lib.workspace(L.fpv_chamfer_bwd_workspace_bytes(T, N, Ms, 1, 0), a.device)
            _lib.check(L.fpv_chamfer_bwd_bcast(_lib.ptr(a), _lib.ptr(b), T, N, Ms, 1, _lib.ptr(g1), _lib.ptr(g2), bc1,
                                               _lib.ptr(i_b2a), _lib.ptr(i_loc), 8, _lib.ptr(grad_a), None,
                                               _lib.ptr(wsb), wsb.numel(), _lib.stream_ptr()), "fpv_chamfer_bwd")
        return grad_a, None, None, None, None, None


def distChamferSharded(a: torch.Tensor, b_shard: torch.Tensor, idx_base: int, group=None,
                       _search: Optional[Callable] = None, clip: bool = False):
    """distChamfer with the scene sharded over the ranks of `group`.

    Returns (d_b2a [T,Ms] for THIS rank's shard, d_a2b [T,N] combined over all shards,
             i_b2a [T,Ms] indices into a, i_a2b [T,N] GLOBAL scene indices).
    Loss terms built on d_a2b (replicated on every rank) must be scaled by 1/world_size and terms on
    d_b2a normalised by the global count, so that sum-over-ranks of the local losses is the global loss
    and allreduce_grads() yields the global gradient.
    """
    if b_shard.dim() == 2:
        b_shard = b_shard.unsqueeze(0)
    return _ShardedChamferFn.apply(a, b_shard, int(idx_base), group, _search, bool(clip))


def allreduce_grads(params, group=None) -> None:
    """Sum the (small) parameter gradients across ranks: one flat all-reduce."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    o = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[o:o + n].view_as(g))
        o += n
