"""Symmetric peer-memory arena and exchange helpers for the frame-sharded forward (one process per GPU on one NVSwitch box).

Every rank cudaMalloc's one slab (fz_p2p_alloc), exports it with CUDA IPC and imports the slabs of all peers (handles travel through
torch.distributed once, at set-up): a buffer at offset `o` of the local slab is reachable on rank r at `peer_base[r] + o`.  Exchange
SITES (one per call site of the forward and tensor geometry: buffers + 32 flag words) are carved out of the slab by a bump allocator
that every rank runs in the same order, so offsets agree without any further communication.  Data then moves with fz_p2p_push (2-D
segments written into the peers by 16-byte NVLink stores, flag raised in the destination's slab by the last CTA) and fz_p2p_wait /
fz_gn_combine (spin on the local flags, clear them): plain kernels on the forward's stream, capturable in CUDA graphs.
torch.distributed is used for the one-time handle exchange only."""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import P2PSeg

FLAG_WORDS = 32          # flag words per site (one per source rank)
FLAG_REGION = 4 << 20    # bytes at the head of the slab reserved for flags: 32768 sites


class _Raw:
    """__cuda_array_interface__ view of raw device memory (so that arena buffers are ordinary torch tensors)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = dict(shape=(nbytes,), typestr="|u1", data=(ptr, False), version=3, strides=None)


class Site:
    __slots__ = ("offset", "nbytes", "flag_offset", "index")

    def __init__(self, offset, nbytes, flag_offset, index):
        self.offset, self.nbytes, self.flag_offset, self.index = offset, nbytes, flag_offset, index


class Arena:
    def __init__(self, rank: int, world: int, device: torch.device, group=None, nbytes: Optional[int] = None):
        import torch.distributed as dist
        self.rank, self.world, self.device, self.group = rank, world, device, group
        self.nbytes = int(nbytes or int(os.environ.get("FZ_P2P_ARENA_MB", "6144")) << 20)
        lib = _lib.load()
        base = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.fz_p2p_alloc(self.nbytes, C.byref(base)), "fz_p2p_alloc")
            handle = (C.c_ubyte * 64)()
            _lib.check(lib.fz_p2p_export(base, handle), "fz_p2p_export")
            handles: List[Optional[bytes]] = [None] * world
            dist.all_gather_object(handles, bytes(handle), group=group)
            self.peer_base: List[int] = []
            for r in range(world):
                if r == rank:
                    self.peer_base.append(base.value)
                else:
                    pp = C.c_void_p()
                    hb = (C.c_ubyte * 64).from_buffer_copy(handles[r])
                    _lib.check(lib.fz_p2p_import(hb, C.byref(pp)), "fz_p2p_import")
                    self.peer_base.append(pp.value)
        self.base = base.value
        self._mem = torch.as_tensor(_Raw(self.base, self.nbytes), device=device)  # uint8 view of the local slab (zero-filled)
        self._cursor = FLAG_REGION
        self._n_sites = 0
        self.sites: Dict[tuple, Site] = {}
        self.counters = torch.zeros(4096, dtype=torch.int32, device=device)  # local arrival counters of fz_p2p_push (16 per launch slot)
        self._counter_cursor = 0
        dist.barrier(group=group)

    # ---- allocation ---------------------------------------------------------------------------------------------------
    def site(self, key: tuple, nbytes: int) -> Site:
        """The site of `key` (allocated on first use; every rank must ask for sites in the same order)."""
        s = self.sites.get(key)
        if s is None:
            nbytes = (int(nbytes) + 255) // 256 * 256
            if self._cursor + nbytes > self.nbytes:
                raise RuntimeError(f"frame-shard arena exhausted ({self.nbytes >> 20} MiB; set FZ_P2P_ARENA_MB): site {key} needs {nbytes} bytes")
            if (self._n_sites + 1) * FLAG_WORDS * 4 > FLAG_REGION:
                raise RuntimeError("frame-shard arena: too many exchange sites")
            s = Site(self._cursor, nbytes, self._n_sites * FLAG_WORDS * 4, self._n_sites)
            self._cursor += nbytes
            self._n_sites += 1
            self.sites[key] = s
        elif s.nbytes < nbytes:
            raise RuntimeError(f"exchange site {key} was created with {s.nbytes} bytes, now asked for {nbytes}")
        return s

    def tensor(self, site: Site, byte_offset: int, shape: Sequence[int], dtype=torch.float16) -> torch.Tensor:
        n = 1
        for d in shape:
            n *= int(d)
        nb = n * torch.empty((), dtype=dtype).element_size()
        assert byte_offset + nb <= site.nbytes
        return self._mem[site.offset + byte_offset: site.offset + byte_offset + nb].view(dtype).view(*shape)

    def peer_ptr(self, rank: int, site: Site, byte_offset: int = 0) -> int:
        return self.peer_base[rank] + site.offset + byte_offset

    def flag_ptr(self, rank: int, site: Site, source: int) -> int:
        return self.peer_base[rank] + site.flag_offset + 4 * source

    def _counter_block(self, site: Site) -> int:
        # 16 counters per site, inside a persistent zero-initialised tensor (the kernels leave them zero)
        idx = (site.index * 16) % (self.counters.numel() - 16)
        return self.counters.data_ptr() + 4 * idx

    # ---- data movement ------------------------------------------------------------------------------------------------
    def exchange(self, site: Site, segs: Sequence[Tuple[int, int, int, int, int, int, int]], sources: Sequence[int], stream) -> None:
        """ONE launch: copy segs = (src_ptr, src_pitch, dst_rank, dst_ptr, dst_pitch, rows, row_bytes) into their destination ranks, raise
        this site's flag on every remote destination, then wait for the flags of `sources` (ranks this rank expects data from)."""
        dsts = sorted({s[2] for s in segs if s[2] != self.rank})
        mask = self.wait_mask(sources)
        if not segs:
            if mask:
                _lib.call("fz_p2p_wait", C.c_void_p(self.base + site.flag_offset), mask, stream)
            return
        slot = {r: i for i, r in enumerate(dsts)}
        arr = (P2PSeg * len(segs))()
        for i, (src, sp, r, dst, dp, rows, rb) in enumerate(segs):
            arr[i].src, arr[i].src_pitch, arr[i].dst, arr[i].dst_pitch = src, sp, dst, dp
            arr[i].rows, arr[i].row_bytes, arr[i].dst_slot = rows, rb, slot.get(r, -1)
        flags = (C.c_void_p * max(1, len(dsts)))(*[self.flag_ptr(r, site, self.rank) for r in dsts])
        _lib.call("fz_p2p_push", arr, len(segs), flags, C.c_void_p(self._counter_block(site)), len(dsts),
                  C.c_void_p(self.base + site.flag_offset) if mask else None, mask, stream)

    def wait_mask(self, sources: Sequence[int]) -> int:
        mask = 0
        for r in sources:
            if r != self.rank:
                mask |= 1 << r
        return mask
