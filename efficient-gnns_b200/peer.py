"""Peer-memory exchange group of the node-parallel engine.

One process per GPU (torch.distributed supplies rank/world and moves the 64-byte IPC handles once at start-up).
Every rank owns an *exchange arena* (cudaMalloc through the C ABI), carves identically laid out buffers from it
(same name -> same offset on every rank) and maps the peers' arenas with CUDA IPC.  An exchange step is then one
``b200gnn_peer_copy2d_f32`` launch whose destinations are the consumers' buffers (stores cross NVLink/NVSwitch
directly; no pack/unpack, no collective call) followed by ``b200gnn_peer_barrier``.  CUDA-graph capturable.

``TorchExchange`` implements the same three primitives with torch.distributed collectives (gloo on CPU for the host
logic tests, NCCL as the baseline the peer path is measured against).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import lib


class _RawCuda:
    """Expose raw device memory to torch through __cuda_array_interface__ (no copy, no ownership)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3,
                                         "strides": None}


def _wrap(ptr: int, nbytes: int, device: torch.device) -> torch.Tensor:
    return torch.as_tensor(_RawCuda(ptr, nbytes), device=device)


class PeerArena:
    """Symmetric exchange arena + peer mappings + the flag barrier."""

    FLAG_BYTES = 256            # 16 uint64 flag slots + epoch + error word, at the start of the arena

    def __init__(self, nbytes: int, group=None):
        assert dist.is_initialized()
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        assert self.world <= 16
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.nbytes = (int(nbytes) + self.FLAG_BYTES + 255) // 256 * 256
        L = lib.load()
        base = C.c_void_p()
        lib.check(L.b200gnn_arena_alloc(self.nbytes, C.byref(base)), "arena_alloc")
        self.base = int(base.value)
        handle = (C.c_ubyte * 64)()
        lib.check(L.b200gnn_ipc_get_handle(C.c_void_p(self.base), handle), "ipc_get_handle")
        mine = bytes(handle)
        handles: List[Optional[bytes]] = [None] * self.world
        dist.all_gather_object(handles, mine, group=group)
        self.peer_base: List[int] = []
        for q in range(self.world):
            if q == self.rank:
                self.peer_base.append(self.base)
                continue
            out = C.c_void_p()
            buf = (C.c_ubyte * 64).from_buffer_copy(handles[q])
            lib.check(L.b200gnn_ipc_open_handle(buf, C.byref(out)), "ipc_open_handle")
            self.peer_base.append(int(out.value))
        self._mem = _wrap(self.base, self.nbytes, self.device)
        self._off = self.FLAG_BYTES
        self._flag_ptrs = (C.c_void_p * self.world)(*[C.c_void_p(b) for b in self.peer_base])
        self._epoch_ptr = self.base + 128
        self._error_ptr = self.base + 136
        self._ticket_ptr = self.base + 144
        self.buffers: Dict[str, Tuple[int, Tuple[int, ...]]] = {}
        dist.barrier(group=group)

    # -- symmetric allocation: identical call sequence on every rank => identical offsets
    def alloc(self, name: str, shape: Sequence[int]) -> torch.Tensor:
        n = 1
        for d in shape:
            n *= int(d)
        nbytes = (n * 4 + 255) // 256 * 256
        if self._off + nbytes > self.nbytes:
            raise lib.B200GnnError(f"exchange arena exhausted allocating {name} {tuple(shape)}")
        off = self._off
        self._off += nbytes
        self.buffers[name] = (off, tuple(int(d) for d in shape))
        return self._mem[off:off + n * 4].view(torch.float32).view(*shape)

    def peer_ptr(self, name: str, q: int, elem_offset: int = 0) -> int:
        """Device address (in THIS process) of element `elem_offset` of rank q's buffer `name`."""
        return self.peer_base[q] + self.buffers[name][0] + 4 * int(elem_offset)

    def barrier(self):
        lib.check(lib.load().b200gnn_peer_barrier(self._flag_ptrs, self.rank, self.world, C.c_void_p(self._epoch_ptr),
                                                  C.c_void_p(self._error_ptr), lib.stream_ptr()), "peer_barrier")

    def exchange(self, copies, width: int):
        """The copies (dst_ptr, src_ptr, ld_dst, ld_src, rows) and the flag barrier in ONE launch (<= 16 copies)."""
        if len(copies) > 16:
            copy2d(copies, width)
            return self.barrier()
        arr = (lib.Copy2D * max(len(copies), 1))()
        for j, (d, s, ldd, lds, rows) in enumerate(copies):
            arr[j].dst, arr[j].src, arr[j].ld_dst, arr[j].ld_src, arr[j].rows = d, s, ldd, lds, rows
        lib.check(lib.load().b200gnn_peer_exchange_f32(arr, len(copies), int(width), self._flag_ptrs, self.rank, self.world,
                                                       C.c_void_p(self._epoch_ptr), C.c_void_p(self._error_ptr),
                                                       C.c_void_p(self._ticket_ptr), lib.stream_ptr()), "peer_exchange_f32")

    def error_flag(self) -> int:
        return int(self._mem[136:140].view(torch.int32).item())

    def close(self):
        L = lib.load()
        torch.cuda.synchronize()
        for q, b in enumerate(self.peer_base):
            if q != self.rank and b:
                L.b200gnn_ipc_close_handle(C.c_void_p(b))
        self.peer_base = []
        if self.base:
            L.b200gnn_arena_free(C.c_void_p(self.base))
            self.base = 0


def copy2d(copies: List[Tuple[int, int, int, int, int]], width: int):
    """copies: (dst_ptr, src_ptr, ld_dst, ld_src, rows) with raw device addresses; one launch."""
    for i in range(0, len(copies), 16):
        part = copies[i:i + 16]
        arr = (lib.Copy2D * len(part))()
        for j, (d, s, ldd, lds, rows) in enumerate(part):
            arr[j].dst, arr[j].src, arr[j].ld_dst, arr[j].ld_src, arr[j].rows = d, s, ldd, lds, rows
        lib.check(lib.load().b200gnn_peer_copy2d_f32(arr, len(part), int(width), lib.stream_ptr()), "peer_copy2d_f32")
