"""Peer-memory all-gather over NVLink / NVSwitch for the data-parallel InfoNCE (csrc/peer.cu).

One process per GPU.  Every rank allocates a "symmetric" buffer in the library, exports it through CUDA IPC,
maps every peer's buffer, and from then on ``allgather`` is ONE kernel launch per rank -- publish, flag, wait,
pull -- with no NCCL call and no host synchronisation, so it is captured into the step's CUDA graph like any
other kernel.  It carries the two exchanges that sit between the towers and the loss: the
normalised features (straight into ``img_all`` / ``txt_all``) and the row log-sum-exps + loss statistics.

``PeerExchange.create`` is collective over the group and returns ``None`` on EVERY rank if any rank cannot set the
exchange up (no peer access, IPC unavailable in the container, self-test mismatch); the caller then keeps using
NCCL.  The reference is single-GPU (CLIP/train.py:103): this is new plumbing, validated by ``tools/dist_check.py``.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.distributed as dist

from . import lib as L

_MIN_SLOT = 8 << 20


class PeerExchange:
    def __init__(self):
        self.world = self.rank = 0
        self.ctx = None
        self.own = 0
        self.peers: list[int] = []
        self.table = None
        self.slot_bytes = 0
        self.device = None

    # ------------------------------------------------------------------------------------ set-up
    @classmethod
    def create(cls, group, device: torch.device, slot_bytes: int = _MIN_SLOT):
        """Collective.  -> PeerExchange, or None (on every rank) when the exchange cannot be used."""
        if not (dist.is_available() and dist.is_initialized()) or device.type != "cuda":
            return None
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world < 2 or world > 16:
            return None
        self = cls()
        self.world, self.rank, self.device = world, rank, device
        self.slot_bytes = max(int(slot_bytes), _MIN_SLOT)
        lib = L.load()
        idx = device.index if device.index is not None else torch.cuda.current_device()
        handle = (C.c_ubyte * 64)()
        ok = True
        try:
            self.ctx = L.ctx(idx)
            total = lib.b200clip_peer_buffer_bytes(world, self.slot_bytes)
            ptr = C.c_void_p()
            L.check(lib.b200clip_peer_alloc(self.ctx, total, C.byref(ptr), handle), "peer_alloc")
            self.own = ptr.value
        except RuntimeError:
            ok = False
        infos = [None] * world
        dist.all_gather_object(infos, (ok, bytes(handle), os.getpid()), group=group)
        ok = all(i[0] for i in infos)
        self.peers = [0] * world
        if ok:
            try:
                for r, (_, h, _pid) in enumerate(infos):
                    if r == rank:
                        self.peers[r] = self.own
                        continue
                    p = C.c_void_p()
                    hb = (C.c_ubyte * 64).from_buffer_copy(h)
                    L.check(lib.b200clip_peer_open(self.ctx, hb, C.byref(p)), "peer_open")
                    self.peers[r] = p.value
            except RuntimeError:
                ok = False
        if not self._agree(ok, group):
            self.close()
            return None
        self.table = torch.tensor(self.peers, dtype=torch.int64, device=device)
        ok = self._self_test()
        if not self._agree(ok, group):
            self.close()
            return None
        return self

    def _agree(self, ok: bool, group) -> bool:
        t = torch.tensor([1 if ok else 0], device=self.device, dtype=torch.int32)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
        return bool(t.item())

    def _self_test(self) -> bool:
        """Three exchanges (both slots, odd sizes: the 4-byte path, then the 16-byte path) against the known answer."""
        try:
            for n in (37, 4096, 70000):
                src = (torch.arange(n, device=self.device, dtype=torch.float32) + 1000.0 * (self.rank + 1))
                src2 = -src[: max(4, n // 8 * 4)].clone()
                dst = torch.zeros((self.world, n), device=self.device, dtype=torch.float32)
                dst2 = torch.zeros((self.world, src2.numel()), device=self.device, dtype=torch.float32)
                self.allgather([src, src2], [dst, dst2])
                exp = torch.arange(n, device=self.device, dtype=torch.float32)[None] + \
                    1000.0 * (torch.arange(self.world, device=self.device, dtype=torch.float32)[:, None] + 1)
                if not (torch.equal(dst, exp) and torch.equal(dst2, -exp[:, : src2.numel()])):
                    return False
            return not self.timed_out()
        except RuntimeError:
            return False

    # ------------------------------------------------------------------------------------ the exchange
    def fits(self, tensors) -> bool:
        return sum(t.numel() * t.element_size() for t in tensors) <= self.slot_bytes

    def allgather(self, srcs, dsts) -> None:
        """dsts[s][r] <- rank r's srcs[s]; every rank passes the same shapes.  1..3 segments, contiguous tensors."""
        n = len(srcs)
        assert 1 <= n <= 3 and len(dsts) == n
        src_p, dst_p, nbytes = (C.c_void_p * 3)(), (C.c_void_p * 3)(), (C.c_int64 * 3)()
        for s, (a, b) in enumerate(zip(srcs, dsts)):
            if not (a.is_contiguous() and b.is_contiguous() and a.is_cuda and b.is_cuda):
                raise RuntimeError("peer allgather: contiguous CUDA tensors only")
            if b.numel() * b.element_size() != self.world * a.numel() * a.element_size():
                raise RuntimeError("peer allgather: destination must hold world x source")
            src_p[s], dst_p[s], nbytes[s] = a.data_ptr(), b.data_ptr(), a.numel() * a.element_size()
        st = torch.cuda.current_stream(self.device).cuda_stream
        L.check(L.load().b200clip_peer_allgather(self.ctx, self.table.data_ptr(), self.world, self.rank, self.slot_bytes,
                                                 n, src_p, dst_p, nbytes, st), "peer_allgather")

    def timed_out(self) -> bool:
        """True if a wait for a peer ever gave up (synchronises the device)."""
        out = (C.c_int64 * 2)()
        L.check(L.load().b200clip_peer_status(self.ctx, self.own, out), "peer_status")
        return bool(out[1])

    def close(self) -> None:
        lib = L.load()
        for r, p in enumerate(self.peers):
            if p and r != self.rank:
                lib.b200clip_peer_close(self.ctx, p)
        if self.own:
            torch.cuda.synchronize(self.device)
            lib.b200clip_peer_free(self.ctx, self.own)
        self.peers, self.own, self.table = [], 0, None


# one exchange per (process group, device), created on first use outside of a CUDA-graph capture
_cache: dict = {}


def mode() -> str:
    """B200CLIP_FEATURE_GATHER: peer (default) | allreduce | allgather."""
    return os.environ.get("B200CLIP_FEATURE_GATHER", "peer")


def active() -> list:
    """The exchanges this process has set up (bench.py / tools check their time-out flags)."""
    return [e for e in _cache.values() if e is not None]


def get(group, device: torch.device, need_bytes: int):
    """The group's exchange if it exists (or can be created now) and holds `need_bytes` per rank; else None."""
    if mode() != "peer" or device.type != "cuda" or not (dist.is_available() and dist.is_initialized()):
        return None
    key = (id(group) if group is not None else 0, device.index)
    if key not in _cache:
        if torch.cuda.is_current_stream_capturing():
            return None   # cannot set up inside a capture; every rank takes the same branch
        _cache[key] = PeerExchange.create(group, device, max(need_bytes, _MIN_SLOT))
    ex = _cache[key]
    if ex is None or need_bytes > ex.slot_bytes:
        return None
    return ex
