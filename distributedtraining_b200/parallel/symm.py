"""Symmetric peer-memory windows over NVLink 5 / NVSwitch (one process per GPU).

Each rank allocates one window with the C++ runtime (csrc/symm_runtime.cu: cudaMalloc + CUDA IPC), the handles are
exchanged through the ``torch.distributed`` control plane, and every rank maps every peer's window.  Kernels then load
from / store to the peer addresses directly (csrc/optim_avg.cu, csrc/sm100_gemm.cu), synchronised by release/acquire
round flags that live inside the windows.

This is the in-box replacement of the reference's tensor plane: ``torch.save`` -> git-lfs push -> ``hf_hub_download``
-> ``torch.load`` (reference hivetrain/hf_manager.py:91-136,186-197) and of its SHA polling (:151-159).
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from ..ops import _lib

FLAG_BYTES = 4096
FLAG_WORDS = FLAG_BYTES // 4
# flag word layout (uint32 indices); each block has one slot per source rank (<= 64 ranks)
F_DELTA = 0     # [F_DELTA + src]  = last delta round published by rank src
F_BASE = 64     # [F_BASE + src]   = last base round pushed by averager-shard src
F_BARRIER = 128  # [F_BARRIER + src] = barrier epoch
F_HEART = 192   # [F_HEART + src]  = averager's mixing-matrix epoch (w broadcast)
F_ALIVE = 256   # [F_ALIVE + src]  = liveness heartbeat counter (failure detection)
F_TB = 320      # [F_TB + src]     = meta tick whose theta_bar bf16 shard rank src has completed
F_G = 384       # [F_G + src]      = meta tick whose validation gradient rank src has completed (data-parallel mode)
F_GP = 448      # [F_GP + src]     = meta tick whose meta-gradient partial rank src has stored into my slot table
F_BAD = 512     # [F_BAD + src]    = round whose delta from miner src holds NaN/Inf (0 / older round = clean)


class _CudaBuffer:
    """Minimal ``__cuda_array_interface__`` carrier so torch can wrap a raw device pointer without copying."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3,
                                         "strides": None}


def tensor_from_ptr(ptr: int, nbytes: int, device: torch.device) -> torch.Tensor:
    return torch.as_tensor(_CudaBuffer(ptr, nbytes), device=device)


@dataclass
class Region:
    name: str
    offset: int
    nbytes: int


class SymmetricWindow:
    """One window per rank with named regions; identical layout on every rank (hence "symmetric")."""

    def __init__(self, regions: Dict[str, int], group: Optional[dist.ProcessGroup] = None, device: Optional[torch.device] = None):
        assert torch.cuda.is_available(), "SymmetricWindow needs CUDA"
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        L = _lib.lib()
        _lib.check(L.dtb_set_device(self.device.index), "dtb_set_device")
        self.regions: Dict[str, Region] = {}
        off = FLAG_BYTES
        self.regions["flags"] = Region("flags", 0, FLAG_BYTES)
        for name, nbytes in regions.items():
            nbytes = (nbytes + 1023) // 1024 * 1024
            self.regions[name] = Region(name, off, nbytes)
            off += nbytes
        self.nbytes = off
        self.peer_ptrs: List[int] = [0] * self.world
        self.p2p = True
        self.mc_ptr = 0  # multicast (NVLS) address of the windows; 0 = this window cannot be addressed through the switch
        self.backing = "ipc"
        want = os.environ.get("DTB200_SYMM", "vmm")
        if want == "vmm" and self._init_vmm(L, group):
            self.backing = "vmm"
            self._local = tensor_from_ptr(self.local_ptr, self.nbytes, self.device)
            self.error_flag = torch.zeros(1, dtype=torch.int32, device=self.device)
            self._epoch = 0
            return
        p = ctypes.c_void_p()
        _lib.check(L.dtb_symm_alloc(ctypes.c_size_t(self.nbytes), ctypes.byref(p)), "dtb_symm_alloc")
        self.local_ptr = p.value
        self._local = tensor_from_ptr(self.local_ptr, self.nbytes, self.device)
        self.peer_ptrs[self.rank] = self.local_ptr
        if self.world > 1:
            h = ctypes.create_string_buffer(64)
            _lib.check(L.dtb_ipc_get_handle(ctypes.c_void_p(self.local_ptr), h), "ipc_get_handle")
            handles: List[Optional[bytes]] = [None] * self.world
            dist.all_gather_object(handles, bytes(h.raw), group=group)
            for r, hr in enumerate(handles):
                if r == self.rank:
                    continue
                q = ctypes.c_void_p()
                rc = L.dtb_ipc_open_handle(ctypes.create_string_buffer(hr, 64), ctypes.byref(q))
                if rc != 0:
                    L.dtb_last_error()
                    self.p2p = False
                    raise RuntimeError(f"cudaIpcOpenMemHandle(rank {r}) failed with {rc}; peer windows unavailable")
                self.peer_ptrs[r] = q.value
            dist.barrier(group=group)
        self.error_flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._epoch = 0

    # -- VMM + NVLS multicast backing (csrc/symm_runtime.cu) ----------------------------------------------------------
    def _init_vmm(self, L, group) -> bool:
        """cuMemCreate window -> POSIX fd -> every peer maps it (unicast, NVLink P2P) and all windows are bound to ONE
        multicast object (``multimem.*`` through the NVSwitch).  The fds travel over unix sockets (SCM_RIGHTS).  Any failure
        (old driver, no fabric, sandboxed /tmp) makes every rank fall back to the cudaMalloc + CUDA-IPC backing."""
        import socket
        import tempfile
        import uuid
        dev = self.device.index
        if not hasattr(L, "dtb_vmm_granularity"):
            return False
        L.dtb_vmm_granularity.restype = ctypes.c_size_t
        gran = int(L.dtb_vmm_granularity(dev, self.world))
        ok = gran > 0
        size = (self.nbytes + gran - 1) // gran * gran if ok else 0
        handle, ptr, fd = ctypes.c_ulonglong(0), ctypes.c_void_p(), ctypes.c_int(-1)
        if ok:
            ok = L.dtb_vmm_alloc(ctypes.c_size_t(size), ctypes.c_size_t(gran), dev, ctypes.byref(handle), ctypes.byref(ptr),
                                 ctypes.byref(fd)) == 0

        def all_ok(flag: bool) -> bool:
            if self.world == 1:
                return flag
            t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=self.device)
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
            return bool(t.item())
        if not all_ok(ok):
            return False
        self.local_ptr = ptr.value
        self.peer_ptrs[self.rank] = self.local_ptr
        self.nbytes_mapped = size
        if self.world == 1:
            return True
        # ---- rank 0 creates the multicast object; fds are handed over rank by rank ----
        mc, mc_fd = ctypes.c_ulonglong(0), ctypes.c_int(-1)
        want_mc = bool(L.dtb_mc_supported(dev)) and os.environ.get("DTB200_NO_MULTICAST", "0") != "1"
        if self.rank == 0 and want_mc:
            want_mc = L.dtb_mc_create(self.world, ctypes.c_size_t(size), ctypes.byref(mc), ctypes.byref(mc_fd)) == 0
        want_mc = all_ok(want_mc)
        token = [uuid.uuid4().hex if self.rank == 0 else None]
        dist.broadcast_object_list(token, src=0, group=group)
        path = lambda r: os.path.join(tempfile.gettempdir(), f"dtb200_{token[0]}_{r}.sock")
        srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        srv.bind(path(self.rank))
        srv.listen(self.world)
        dist.barrier(group=group)
        peer_fds = {}
        try:
            for p in range(self.world):
                if p == self.rank:
                    for _ in range(self.world - 1):  # hand my window fd (and, from rank 0, the multicast fd) to every peer
                        c, _a = srv.accept()
                        fds = [fd.value] + ([mc_fd.value] if (self.rank == 0 and want_mc) else [])
                        socket.send_fds(c, [b"w"], fds)
                        c.close()
                else:
                    c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
                    c.connect(path(p))
                    _m, fds, _f, _ad = socket.recv_fds(c, 16, 2)
                    c.close()
                    peer_fds[p] = list(fds)
        finally:
            srv.close()
            try:
                os.unlink(path(self.rank))
            except OSError:
                pass
        good = True
        for p, fds in peer_fds.items():
            q = ctypes.c_void_p()
            rc = L.dtb_vmm_import(fds[0], ctypes.c_size_t(size), ctypes.c_size_t(gran), dev, ctypes.byref(q))
            good = good and rc == 0
            self.peer_ptrs[p] = q.value or 0
            os.close(fds[0])
            if p == 0 and want_mc and len(fds) > 1:
                good_mc = L.dtb_mc_import(fds[1], ctypes.byref(mc)) == 0
                os.close(fds[1])
                want_mc = want_mc and good_mc
        if not all_ok(good):
            raise RuntimeError("VMM peer mapping failed on some rank (set DTB200_SYMM=ipc to use CUDA IPC windows)")
        os.close(fd.value)
        if self.rank == 0 and mc_fd.value >= 0:
            os.close(mc_fd.value)
        want_mc = all_ok(want_mc)
        if want_mc:
            want_mc = all_ok(L.dtb_mc_add_device(mc, dev) == 0)
        if want_mc:  # every device has been added (the all-reduce above is the barrier the driver asks for): bind + map
            q = ctypes.c_void_p()
            rc = L.dtb_mc_bind_map(mc, handle, ctypes.c_size_t(size), ctypes.c_size_t(gran), dev, ctypes.byref(q))
            if all_ok(rc == 0):
                self.mc_ptr = q.value
        dist.barrier(group=group)
        return True

    def mc(self, region: str, byte_offset: int = 0) -> int:
        """Multicast address of ``region``: a ``multimem.st`` there lands in EVERY rank's window, a ``multimem.ld_reduce``
        sums over the ranks.  0 when the windows are not multicast-bound."""
        return 0 if not self.mc_ptr else self.mc_ptr + self.regions[region].offset + byte_offset

    # -- addressing ------------------------------------------------------------------------------------------------
    def ptr(self, region: str, rank: Optional[int] = None, byte_offset: int = 0) -> int:
        r = self.rank if rank is None else rank
        return self.peer_ptrs[r] + self.regions[region].offset + byte_offset

    def local(self, region: str, dtype=torch.uint8) -> torch.Tensor:
        reg = self.regions[region]
        return self._local[reg.offset:reg.offset + reg.nbytes].view(dtype)

    def peer(self, region: str, rank: int, dtype=torch.uint8) -> torch.Tensor:
        """Tensor view over a PEER's region (loads/stores by torch ops travel over NVLink)."""
        reg = self.regions[region]
        if rank == self.rank:
            return self.local(region, dtype)
        return tensor_from_ptr(self.peer_ptrs[rank] + reg.offset, reg.nbytes, self.device).view(dtype)

    def flag_ptr(self, word: int, rank: Optional[int] = None) -> int:
        return self.ptr("flags", rank, 4 * word)

    def flags(self) -> torch.Tensor:
        return self.local("flags", torch.int32)

    # -- synchronisation -------------------------------------------------------------------------------------------
    def publish(self, block: int, value: int, dst_ranks: Optional[List[int]] = None, cond: Optional[torch.Tensor] = None,
                multicast: bool = False) -> None:
        """After all prior work on the current stream: release-store ``value`` into slot [block + my_rank] of every
        destination rank's flag page (csrc/optim_avg.cu publish_flag_kernel).  ``cond`` (device int32[1]): store ``value``
        only if it is non-zero, else 0 (a verdict computed on the device travels without a host read)."""
        dst = list(range(self.world)) if dst_ranks is None else dst_ranks
        arr = (ctypes.c_void_p * len(dst))(*[self.flag_ptr(block + self.rank, r) for r in dst])
        # ``multicast``: flags that announce multimem.st DATA travel the same way -- one multimem.st.release to the multicast
        # address of the slot (all ranks at once), ordered behind the data by the release
        mc = self.mc("flags", 4 * (block + self.rank)) if (multicast and self.mc_ptr and dst_ranks is None) else 0
        _lib.check(_lib.lib().dtb_publish_flag(arr, len(dst), ctypes.c_uint32(value), _lib.stream_ptr(), _lib.ptr(cond),
                                               ctypes.c_void_p(mc)), "publish_flag")

    def wait(self, block: int, value: int, src_ranks: Optional[List[int]] = None) -> None:
        """Stream-ordered wait (device-side spin) until slot [block + src] >= value for every src."""
        src = list(range(self.world)) if src_ranks is None else src_ranks
        if src == list(range(src[0], src[0] + len(src))):
            _lib.check(_lib.lib().dtb_wait_flags(ctypes.c_void_p(self.flag_ptr(block + src[0])), len(src), 1,
                                                 ctypes.c_uint32(value), _lib.ptr(self.error_flag), _lib.stream_ptr()),
                       "wait_flags")
        else:
            for s in src:
                _lib.check(_lib.lib().dtb_wait_flags(ctypes.c_void_p(self.flag_ptr(block + s)), 1, 1, ctypes.c_uint32(value),
                                                     _lib.ptr(self.error_flag), _lib.stream_ptr()), "wait_flags")

    def device_barrier(self) -> None:
        """All ranks' streams rendezvous on the device (no host sync, no NCCL)."""
        self._epoch += 1
        self.publish(F_BARRIER, self._epoch)
        self.wait(F_BARRIER, self._epoch)

    # -- failure detection ---------------------------------------------------------------------------------------------
    def heartbeat(self) -> None:
        """Bump this rank's liveness counter in every peer's flag page (stream-ordered, one tiny kernel)."""
        self._beat = getattr(self, "_beat", 0) + 1
        self.publish(F_ALIVE, self._beat)

    def stale_ranks(self, min_beat: int) -> List[int]:
        """Ranks whose heartbeat counter is below ``min_beat`` (host read of the local flag page): a dead or stalled
        miner is then treated exactly like a failed download in the reference -- skipped for the round."""
        f = self.flags()[F_ALIVE:F_ALIVE + self.world].tolist()
        return [r for r, v in enumerate(f) if r != self.rank and v < min_beat]

    def check_errors(self) -> None:
        """Blocking check (one host sync): raise if any device-side flag wait of this rank timed out."""
        if int(self.error_flag.item()) != 0:
            raise RuntimeError("peer flag wait timed out (a rank stalled or died)")

    def poll_errors(self) -> None:
        """NON-blocking check for the hot path (once per round): the error flag is copied to pinned host memory behind the
        work queued so far; a copy that has completed by a later call is inspected then.  A timed-out wait therefore raises
        one round late instead of forcing a host sync into every stream-ordered round."""
        pend = getattr(self, "_err_pending", None)
        if pend is not None and pend[1].query():
            if int(pend[0].item()) != 0:
                raise RuntimeError("peer flag wait timed out (a rank stalled or died)")
            pend = None
        if pend is None:
            host = torch.zeros(1, dtype=torch.int32).pin_memory()
            host.copy_(self.error_flag, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            self._err_pending = (host, ev)

    def close(self) -> None:
        L = _lib.lib()
        if self.backing == "vmm":  # mappings are released with the process (windows live as long as the job)
            self.local_ptr = 0
            return
        for r, p in enumerate(self.peer_ptrs):
            if r != self.rank and p:
                L.dtb_ipc_close_handle(ctypes.c_void_p(p))
        if self.local_ptr:
            del self._local
            L.dtb_symm_free(ctypes.c_void_p(self.local_ptr))
            self.local_ptr = 0
