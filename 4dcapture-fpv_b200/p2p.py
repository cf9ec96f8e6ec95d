"""Peer-memory mailboxes of the scene-sharded step -- host side of csrc/p2p.cu.

One process per GPU.  Every rank allocates one mailbox with the library (a plain cudaMalloc block, exportable through
CUDA IPC), the 64-byte handles travel through the process group's object all-gather (any backend: nccl under torchrun,
gloo in the tests), and every rank maps its peers' mailboxes.  From then on the data path is kernels only:

    keys   the body -> scene search stores packed (distance, global index) keys into slot `rank` of EVERY mailbox from
           its own epilogue (fpv_nn_culled_search_keys); barrier(); every rank takes the minimum over the slots of its
           own mailbox (fpv_p2p_min_unpack)
    floats small vectors (the parameter gradients, the loss) are pushed the same way and summed in rank order
           (fpv_p2p_push / fpv_p2p_sum): deterministic and bit-identical on every rank

No NCCL call, no allocation, no host synchronisation per step, so the sharded step captures into a CUDA graph.
Both channels are double-buffered on a device-side parity word that flips after every consume, which makes one barrier
per exchange sufficient (a rank two exchanges ahead has necessarily seen its peers finish the previous read).

Layout of a mailbox (bytes):  [0,256) flags[world] u32 | 256 epoch u32 | 260 error u32 | 264 parity_keys u32 |
268 parity_floats u32 | 1024.. keys: 2 halves x world slots x key_capacity u64 | floats: 2 x world x float_capacity f32
"""
from __future__ import annotations

import ctypes
from typing import List, Optional

import torch
import torch.distributed as dist

from . import _lib

_HDR = 1024
_OFF_EPOCH, _OFF_ERROR, _OFF_PAR_KEYS, _OFF_PAR_FLOATS = 256, 260, 264, 268


class Mailbox:
    def __init__(self, device, group=None, key_capacity: int = 0, float_capacity: int = 0, timeout_s: float = 20.0):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("p2p.Mailbox needs an initialised torch.distributed process group (any backend)")
        self.device = torch.device(device)
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > 8:
            raise RuntimeError("p2p.Mailbox: at most 8 ranks (one box)")
        self.key_cap = max(int(key_capacity), 1)
        self.float_cap = max((int(float_capacity) + 63) // 64 * 64, 64)
        self.timeout_s = float(timeout_s)
        self.keys_off = _HDR
        self.keys_half = self.world * self.key_cap                      # elements between the two halves
        self.floats_off = self.keys_off + 2 * self.keys_half * 8
        self.floats_half = self.world * self.float_cap
        self.nbytes = self.floats_off + 2 * self.floats_half * 4
        L = _lib.lib()
        with torch.cuda.device(self.device):
            p = ctypes.c_void_p()
            _lib.check(L.fpv_p2p_alloc(self.nbytes, ctypes.byref(p)), "fpv_p2p_alloc")
            self.base = int(p.value)
            handle = (ctypes.c_ubyte * 64)()
            _lib.check(L.fpv_p2p_export(ctypes.c_void_p(self.base), handle), "fpv_p2p_export")
        mine = (bytes(handle), self.key_cap, self.float_cap)
        gathered: List[Optional[tuple]] = [None] * self.world
        dist.all_gather_object(gathered, mine, group=group)
        self.peer_base: List[int] = []
        self._opened: List[int] = []
        with torch.cuda.device(self.device):
            for r, (h, kc, fc) in enumerate(gathered):
                if (kc, fc) != (self.key_cap, self.float_cap):
                    raise RuntimeError("p2p.Mailbox: ranks disagree on the mailbox capacities")
                if r == self.rank:
                    self.peer_base.append(self.base)
                    continue
                q = ctypes.c_void_p()
                buf = (ctypes.c_ubyte * 64).from_buffer_copy(h)
                _lib.check(L.fpv_p2p_open(buf, ctypes.byref(q)), "fpv_p2p_open")
                self.peer_base.append(int(q.value))
                self._opened.append(int(q.value))
        self._flags = (ctypes.c_void_p * self.world)(*[ctypes.c_void_p(b) for b in self.peer_base])
        dist.barrier(group=group)              # every mailbox is mapped (and zeroed) before anyone signals

    # ---- addresses ----
    def key_slot(self, owner_rank: int, slot: int) -> int:
        """Device address (first half) of key slot `slot` in rank `owner_rank`'s mailbox."""
        return self.peer_base[owner_rank] + self.keys_off + slot * self.key_cap * 8

    def float_slot(self, owner_rank: int, slot: int) -> int:
        return self.peer_base[owner_rank] + self.floats_off + slot * self.float_cap * 4

    @property
    def parity_keys(self) -> int:
        return self.base + _OFF_PAR_KEYS

    @property
    def parity_floats(self) -> int:
        return self.base + _OFF_PAR_FLOATS

    def push_targets(self) -> List[int]:
        """This rank's key slot in every OTHER rank's mailbox (what the search epilogue stores into)."""
        return [self.key_slot(r, self.rank) for r in range(self.world) if r != self.rank]

    # ---- operations (all enqueue on the current stream) ----
    def barrier(self) -> None:
        L = _lib.lib()
        with torch.cuda.device(self.device):
            _lib.check(L.fpv_p2p_barrier(self._flags, self.rank, self.world, ctypes.c_void_p(self.base + _OFF_EPOCH),
                                         ctypes.c_void_p(self.base + _OFF_ERROR), self.timeout_s, _lib.stream_ptr()),
                       "fpv_p2p_barrier")

    def combine_keys(self, n: int, row: int, perm_row, idx_dtype=torch.int32, out=None):
        """barrier, then min over the slots of the local mailbox -> (dist [n], idx [n]) un-permuted row by row."""
        from . import spatial
        if n > self.key_cap:
            raise RuntimeError(f"p2p.Mailbox: {n} keys exceed the capacity {self.key_cap}")
        self.barrier()
        return spatial.min_unpack(self.base + self.keys_off, self.world, n, row, perm_row, idx_dtype, device=self.device,
                                  slot_stride=self.key_cap, half_stride=self.keys_half, parity=self.parity_keys, flip=True,
                                  out=out)

    def allreduce_sum(self, flat: torch.Tensor) -> torch.Tensor:
        """Sum of a small float32 vector over the ranks, in rank order (same bits on every rank).  Returns a new tensor."""
        n = flat.numel()
        if n > self.float_cap:
            raise RuntimeError(f"p2p.Mailbox: {n} floats exceed the capacity {self.float_cap}")
        L = _lib.lib()
        src = flat.contiguous().float()
        out = torch.empty_like(src)
        dst = (ctypes.c_void_p * self.world)(*[ctypes.c_void_p(self.float_slot(r, self.rank)) for r in range(self.world)])
        with torch.cuda.device(self.device):
            _lib.check(L.fpv_p2p_push(_lib.ptr(src), n, dst, self.world, ctypes.c_void_p(self.parity_floats),
                                      self.floats_half, _lib.stream_ptr()), "fpv_p2p_push")
            self.barrier()
            _lib.check(L.fpv_p2p_sum(ctypes.c_void_p(self.base + self.floats_off), self.world, self.float_cap,
                                     self.floats_half, ctypes.c_void_p(self.parity_floats), 1, n, _lib.ptr(out),
                                     _lib.stream_ptr()), "fpv_p2p_sum")
        return out

    def all_gather(self, piece: torch.Tensor, offsets, counts, out: torch.Tensor) -> torch.Tensor:
        """out[offsets[r] : offsets[r] + counts[r]] = rank r's `piece` (flat float32), for every r.  Every rank pushes
        its piece into slot `rank` of every mailbox, one barrier, one read-out kernel."""
        L = _lib.lib()
        n = int(counts[self.rank])
        if max(counts) > self.float_cap or piece.numel() != n:
            raise RuntimeError("p2p.Mailbox.all_gather: piece size / capacity mismatch")
        src = piece.contiguous().float()
        dst = (ctypes.c_void_p * self.world)(*[ctypes.c_void_p(self.float_slot(r, self.rank)) for r in range(self.world)])
        off = (ctypes.c_int64 * self.world)(*[int(o) for o in offsets])
        cnt = (ctypes.c_int64 * self.world)(*[int(c) for c in counts])
        with torch.cuda.device(self.device):
            _lib.check(L.fpv_p2p_push(_lib.ptr(src), n, dst, self.world, ctypes.c_void_p(self.parity_floats),
                                      self.floats_half, _lib.stream_ptr()), "fpv_p2p_push")
            self.barrier()
            _lib.check(L.fpv_p2p_gather(ctypes.c_void_p(self.base + self.floats_off), self.world, self.float_cap,
                                        self.floats_half, ctypes.c_void_p(self.parity_floats), 1, off, cnt, _lib.ptr(out),
                                        _lib.stream_ptr()), "fpv_p2p_gather")
        return out

    def reduce_scatter(self, full: torch.Tensor, offsets, counts) -> torch.Tensor:
        """sum over the ranks of full[offsets[rank] : + counts[rank]] (flat float32), in rank order: every rank pushes
        each owner's slice of its `full` into slot `rank` of that owner's mailbox, one barrier, one sum kernel."""
        L = _lib.lib()
        if max(counts) > self.float_cap:
            raise RuntimeError("p2p.Mailbox.reduce_scatter: slice exceeds the capacity")
        src = full.contiguous().float().reshape(-1)
        n = int(counts[self.rank])
        out = torch.empty(n, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            for q in range(self.world):
                dst = (ctypes.c_void_p * 1)(ctypes.c_void_p(self.float_slot(q, self.rank)))
                piece = src[int(offsets[q]):int(offsets[q]) + int(counts[q])]
                _lib.check(L.fpv_p2p_push(_lib.ptr(piece), int(counts[q]), dst, 1, ctypes.c_void_p(self.parity_floats),
                                          self.floats_half, _lib.stream_ptr()), "fpv_p2p_push")
            self.barrier()
            _lib.check(L.fpv_p2p_sum(ctypes.c_void_p(self.base + self.floats_off), self.world, self.float_cap,
                                     self.floats_half, ctypes.c_void_p(self.parity_floats), 1, n, _lib.ptr(out),
                                     _lib.stream_ptr()), "fpv_p2p_sum")
        return out

    def check(self) -> None:
        """Host-side: raise if a barrier ever timed out (a peer died or fell out of step).  Synchronises."""
        torch.cuda.synchronize(self.device)
        word = torch.empty(1, dtype=torch.int32, device=self.device)
        ctypes_ptr = ctypes.c_void_p(self.base + _OFF_ERROR)
        # one 4-byte device-to-device copy through torch (no extra ABI entry needed)
        src = _as_tensor(ctypes_ptr.value, 1, torch.int32, self.device)
        word.copy_(src)
        if int(word.item()) != 0:
            raise RuntimeError("p2p.Mailbox: a barrier timed out -- a peer rank did not arrive")

    def close(self) -> None:
        L = _lib.lib()
        if getattr(self, "base", 0):
            torch.cuda.synchronize(self.device)
            try:
                dist.barrier(group=self.group)  # nobody unmaps while a peer may still store into it
            except Exception:
                pass
            with torch.cuda.device(self.device):
                for q in self._opened:
                    L.fpv_p2p_close(ctypes.c_void_p(q))
                L.fpv_p2p_free(ctypes.c_void_p(self.base))
            self._opened, self.base = [], 0


class _AllGatherRows(torch.autograd.Function):
    """rows [Tr, F] of this rank -> [T, F] of every rank (row ranges per rank); backward = reduce-scatter of the
    gradient: every rank holds a partial gradient for ALL rows, the owner of a row range gets the sum."""

    @staticmethod
    def forward(ctx, rows, box, row_ranges):
        F = rows.shape[1]
        offsets = [b * F for b, _ in row_ranges]
        counts = [(e - b) * F for b, e in row_ranges]
        out = torch.empty(row_ranges[-1][1], F, dtype=torch.float32, device=rows.device)
        box.all_gather(rows.reshape(-1), offsets, counts, out)
        ctx.box, ctx.offsets, ctx.counts, ctx.shape = box, offsets, counts, tuple(rows.shape)
        return out

    @staticmethod
    def backward(ctx, g):
        return ctx.box.reduce_scatter(g, ctx.offsets, ctx.counts).view(ctx.shape), None, None


def all_gather_rows(rows: torch.Tensor, box: "Mailbox", row_ranges) -> torch.Tensor:
    """Differentiable all-gather of per-rank row blocks through the mailbox (forward: push + barrier + read-out;
    backward: reduce-scatter in rank order)."""
    return _AllGatherRows.apply(rows, box, row_ranges)


def _as_tensor(addr: int, n: int, dtype, device) -> torch.Tensor:
    """A torch view of `n` elements of library-owned device memory (via the CUDA array interface)."""
    itemsize = torch.empty(0, dtype=dtype).element_size()
    typestr = {torch.int32: "<i4", torch.float32: "<f4", torch.int64: "<i8"}[dtype]

    class _Holder:
        __cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (addr, False), "version": 2,
                                    "strides": (itemsize,)}
    with torch.cuda.device(device):
        return torch.as_tensor(_Holder(), device=device)
