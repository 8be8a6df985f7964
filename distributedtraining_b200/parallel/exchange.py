"""Delta / base exchange planes.

=============  =====================================================================================================
``peer``       one-sided symmetric windows over NVLink/NVSwitch; the averaging is ONE fused kernel that pulls peer
               deltas, applies the learned weights, adds the base and pushes the result (csrc/optim_avg.cu).
``nvls``       uniform / pre-scaled mixers: ONE kernel per rank reduces its shard inside the NVSwitch (``multimem.ld_reduce``) and
               multicasts the new base to every rank (``multimem.st``) -- no per-peer loop, half the ingress of ``peer``.
``collective`` ``all_gather`` + torch weighted sum + ``broadcast`` over a process group (NCCL = the reference-style
               baseline this framework must beat; gloo = CPU plumbing).
``disk``       files in a shared directory (the analogue of the reference's ``LocalHFManager`` /
               ``LocalAverager`` fakes: reference hivetrain/hf_manager.py:200-241, averaging_logic.py:272-332).
=============  =====================================================================================================

All planes move *flat arenas* (see models/arena.py), never ``dict[str, Tensor]``.
"""
from __future__ import annotations

import hashlib
import os
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from .. import ops
from ..models.arena import Manifest

DELTA_DTYPES = {"fp32": torch.float32, "bf16": torch.bfloat16, "fp8": torch.uint8}


class Exchange:
    """Interface (round-based; ``round`` is a monotonically increasing integer replacing the reference's commit SHA)."""

    kind = "abstract"
    one_sided = True

    def publish_delta(self, trainer, round: int) -> None:
        raise NotImplementedError

    def fetch_delta(self, src: int, round: int) -> Optional[torch.Tensor]:
        """fp32-decodable flat delta of miner ``src`` for ``round`` or ``None`` if it has not been published."""
        raise NotImplementedError

    def delta_round(self, src: int) -> int:
        """Last round published by miner ``src`` (0 = nothing yet)."""
        raise NotImplementedError

    def publish_base(self, base: torch.Tensor, round: int) -> None:
        raise NotImplementedError

    def base_round(self) -> int:
        raise NotImplementedError

    def fetch_base(self, out: torch.Tensor) -> int:
        raise NotImplementedError


# ---------------------------------------------------------------------------------------------------------------------
# disk
# ---------------------------------------------------------------------------------------------------------------------
class DiskExchange(Exchange):
    kind = "disk"

    def __init__(self, root: str, rank: int, manifest: Manifest, delta_dtype: str = "fp32"):
        self.root, self.rank, self.man = root, rank, manifest
        self.delta_dtype = DELTA_DTYPES[delta_dtype]
        os.makedirs(os.path.join(root, "deltas"), exist_ok=True)
        os.makedirs(os.path.join(root, "base"), exist_ok=True)

    def _delta_path(self, src: int) -> str:
        return os.path.join(self.root, "deltas", f"weight_diff_{src}.pt")

    def _base_path(self) -> str:
        return os.path.join(self.root, "base", "averaged_model.pt")

    @staticmethod
    def _atomic_save(obj, path: str) -> None:
        tmp = f"{path}.tmp.{os.getpid()}"
        torch.save(obj, tmp)
        os.replace(tmp, path)  # readers never observe a half-written file (the reference sleeps 10 s instead)

    def publish_delta(self, trainer, round: int) -> None:
        d = torch.empty(self.man.total, dtype=self.delta_dtype if self.delta_dtype != torch.uint8 else torch.float32,
                        device=trainer.master.device)
        trainer.emit_delta(d)
        self._atomic_save({"round": round, "fingerprint": self.man.fingerprint(), "delta": d.cpu()}, self._delta_path(self.rank))
        with open(self._delta_path(self.rank) + ".round.tmp", "w") as f:
            f.write(str(round))
        os.replace(self._delta_path(self.rank) + ".round.tmp", self._delta_path(self.rank) + ".round")

    def fetch_delta(self, src: int, round: int) -> Optional[torch.Tensor]:
        p = self._delta_path(src)
        if not os.path.exists(p):
            return None
        try:
            blob = torch.load(p, map_location="cpu", weights_only=False)
        except Exception:
            return None
        if blob.get("fingerprint") != self.man.fingerprint() or blob["delta"].numel() != self.man.total:
            return None  # shape screen (reference averaging_logic.py:406-410)
        if round >= 0 and blob.get("round", -1) < round:
            return None
        return blob["delta"]

    def delta_round(self, src: int) -> int:
        p = self._delta_path(src) + ".round"
        try:
            return int(open(p).read().strip())
        except Exception:
            return 0

    def publish_base(self, base: torch.Tensor, round: int) -> None:
        self._atomic_save({"round": round, "fingerprint": self.man.fingerprint(), "base": base.detach().float().cpu()},
                          self._base_path())
        with open(self._base_path() + ".round.tmp", "w") as f:
            f.write(str(round))
        os.replace(self._base_path() + ".round.tmp", self._base_path() + ".round")

    def base_round(self) -> int:
        try:
            return int(open(self._base_path() + ".round").read().strip())
        except Exception:
            return 0

    def base_hash(self) -> Optional[str]:
        p = self._base_path()
        if not os.path.exists(p):
            return None
        h = hashlib.sha256()
        with open(p, "rb") as f:
            for chunk in iter(lambda: f.read(1 << 20), b""):
                h.update(chunk)
        return h.hexdigest()

    def fetch_base(self, out: torch.Tensor) -> int:
        blob = torch.load(self._base_path(), map_location="cpu", weights_only=False)
        out.copy_(blob["base"].to(out.device))
        return int(blob["round"])


# ---------------------------------------------------------------------------------------------------------------------
# collectives (baseline / CPU plumbing)
# ---------------------------------------------------------------------------------------------------------------------
class CollectiveExchange(Exchange):
    """Synchronous rounds over a process group: all ranks call :meth:`allgather_average` together."""

    kind = "collective"
    one_sided = False

    def __init__(self, manifest: Manifest, group=None, delta_dtype: str = "fp32"):
        self.man, self.group = manifest, group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.delta_dtype = DELTA_DTYPES[delta_dtype] if delta_dtype != "fp8" else torch.bfloat16
        self._gathered: Optional[torch.Tensor] = None
        self._round = 0

    def allgather_deltas(self, trainer) -> torch.Tensor:
        n = self.man.total
        dev = trainer.master.device
        mine = torch.empty(n, dtype=self.delta_dtype, device=dev)
        trainer.emit_delta(mine)
        if self._gathered is None or self._gathered.device != dev:
            self._gathered = torch.empty(self.world, n, dtype=self.delta_dtype, device=dev)
        dist.all_gather_into_tensor(self._gathered.view(-1), mine, group=self.group)
        return self._gathered

    def allgather_average(self, trainer, w: torch.Tensor, out: torch.Tensor, gathered: Optional[torch.Tensor] = None) -> torch.Tensor:
        """The NCCL(+ATen) baseline of fused kernel (a): all_gather, then N weighted axpys + base add.  ``gathered``: the
        deltas of THIS round if they were already collected (the learned mixer overwrites the averager's master copy while
        it evaluates candidate averages, so the delta must not be re-emitted afterwards)."""
        g = gathered if gathered is not None else self.allgather_deltas(trainer)
        tid = self.man.tensor_ids(out.device)
        _torch_weighted_avg(trainer.base, g, w, tid, out)
        self._round += 1
        return out

    def broadcast_base(self, base: torch.Tensor, src: int = 0) -> None:
        dist.broadcast(base, src=src, group=self.group)


def _torch_weighted_avg(base, gathered, w, tid, out):
    s = w.sum(0)
    out.copy_(base * s[tid])
    for i in range(gathered.shape[0]):
        out.addcmul_(gathered[i].float(), w[i][tid])
    return out


# ---------------------------------------------------------------------------------------------------------------------
# peer memory (the product path)
# ---------------------------------------------------------------------------------------------------------------------
class PeerExchange(Exchange):
    """Symmetric-window exchange; see :mod:`.symm`.  Windows: delta (x2, round parity), fp8 scales, base (fp32 landing
    buffer for the new averaged model), base16 (bf16 copy feeding the fused broadcast->GEMM path)."""

    kind = "peer"

    def __init__(self, manifest: Manifest, group=None, delta_dtype: str = "fp32", with_base16: bool = True,
                 with_meta: bool = True):
        from .symm import F_BAD, F_BASE, F_DELTA, SymmetricWindow

        self.man = manifest
        self.delta_dtype_name = delta_dtype
        self.delta_dtype = DELTA_DTYPES[delta_dtype]
        esz = torch.empty((), dtype=self.delta_dtype).element_size()
        n = manifest.total
        regions = {"delta0": n * esz, "delta1": n * esz, "base": n * 4}
        if delta_dtype == "fp8":
            regions.update({"scale0": n // 32 * 4, "scale1": n // 32 * 4})
        if with_base16:
            regions["base16"] = n * 2
        regions["w"] = 64 * len(manifest) * 4  # mixing matrix w[N<=64, P] shared by the averager
        if with_meta:
            # distributed learned mixer (parallel/meta.py): peer-readable validation gradient + per-rank meta-gradient tables
            world = dist.get_world_size(group) if dist.is_initialized() else 1
            self.slot_words = (world * len(manifest) + 1 + 63) // 64 * 64
            regions["g"] = n * 4
            regions["gslot"] = world * self.slot_words * 4
        self.with_meta = with_meta
        self.win = SymmetricWindow(regions, group)
        self.rank, self.world = self.win.rank, self.win.world
        self.F_DELTA, self.F_BASE, self.F_BAD = F_DELTA, F_BASE, F_BAD
        self.base_target = torch.zeros(1, dtype=torch.int32, device=self.win.device)  # base round the next forward has to see
        self._bad = torch.zeros(1, dtype=torch.int32, device=self.win.device)      # my delta holds NaN/Inf (set by the emit kernel)
        self.active = torch.ones(max(self.world, 1), dtype=torch.int32, device=self.win.device)  # per-miner mask of the round
        self.n_active = torch.zeros(1, dtype=torch.int32, device=self.win.device)
        self.nan_flags = torch.zeros(max(self.world, 1), dtype=torch.int32, device=self.win.device)
        self._base_round = 0
        self.with_base16 = with_base16
        self._snap = {}  # src rank -> local snapshot buffer of its delta (fetch_delta)

    # -- miner side -------------------------------------------------------------------------------------------------
    def delta_buf(self, round: int, rank: Optional[int] = None) -> torch.Tensor:
        name = f"delta{round & 1}"
        return self.win.local(name, self.delta_dtype) if rank is None else self.win.peer(name, rank, self.delta_dtype)

    def scale_buf(self, round: int, rank: Optional[int] = None) -> Optional[torch.Tensor]:
        if self.delta_dtype_name != "fp8":
            return None
        name = f"scale{round & 1}"
        return self.win.local(name, torch.float32) if rank is None else self.win.peer(name, rank, torch.float32)

    def publish_delta(self, trainer, round: int, dst_ranks: Optional[List[int]] = None) -> None:
        """Delta is written straight into this rank's window (no copies) while the emit kernel screens it for NaN/Inf;
        the verdict (F_BAD slot = round iff bad) and then the round flag are release-stored into every peer's flag page."""
        self._bad.zero_()
        try:
            trainer.emit_delta(self.delta_buf(round)[:self.man.total], self.scale_buf(round), self._bad)
        except TypeError:  # adapters with the two-argument signature
            trainer.emit_delta(self.delta_buf(round)[:self.man.total], self.scale_buf(round))
        self.win.publish(self.F_BAD, round, dst_ranks, cond=self._bad)
        self.win.publish(self.F_DELTA, round, dst_ranks)
        self.win.heartbeat()  # liveness counter next to the round flag (failure detection)

    def delta_round(self, src: int) -> int:
        return int(self.win.flags()[self.F_DELTA + src].item())

    def stale_miners(self, min_beat: int, miners: Optional[Sequence[int]] = None) -> List[int]:
        """Miners that published fewer than ``min_beat`` deltas so far (dead, stalled or late): the averager treats them
        like the reference treats a failed download -- skipped for the round -- and logs them."""
        stale = set(self.win.stale_ranks(min_beat))
        return [r for r in (range(self.world) if miners is None else miners) if r in stale]

    def fetch_delta(self, src: int, round: int, snapshot: bool = True) -> Optional[torch.Tensor]:
        """Delta of miner ``src`` (``round <= 0``: whatever it published last), ``None`` if nothing (new enough) is there.

        The window is a 2-deep round-parity ring that an asynchronous miner keeps overwriting, so by default the delta is
        SNAPSHOT into local HBM with one device copy and validated seqlock-style: the publish flag is re-read after the
        copy and the snapshot only counts if no newer round was published meanwhile (the buffer of round r is next written
        for round r+2, whose emit starts only after round r+1 was flagged) -- otherwise retry on the newer round.  The
        reference's download is atomic in the same sense (hf_manager.py:186-197).  ``snapshot=False`` returns the
        zero-copy peer view (synchronous rounds, where the producer cannot run ahead)."""
        n = self.man.total
        for _ in range(4):
            latest = int(self.win.flags()[self.F_DELTA + src].item())
            if latest < max(round, 1):
                return None  # stale flag / never published == failed download in the reference
            r = latest if round <= 0 else max(round, latest)  # always the newest complete publication
            d = self.delta_buf(r, src)[:n]
            sc = self.scale_buf(r, src)
            if not snapshot:
                return ops.dequant_fp8(d, sc) if self.delta_dtype_name == "fp8" else d
            if self.delta_dtype_name == "fp8":
                snap = ops.dequant_fp8(d, sc)  # materialises a local fp32 copy
            else:
                snap = self._snap.get(src)
                if snap is None or snap.numel() != n:
                    snap = self._snap[src] = torch.empty(n, dtype=self.delta_dtype, device=self.win.device)
                snap.copy_(d)
            after = int(self.win.flags()[self.F_DELTA + src].item())  # .item() orders the read after the copy
            if after == r:
                return snap
            round = after  # a newer round landed while we copied: the buffer may be torn, take the newer one
        return None

    # -- push mode: the trainer LIVES in the window and the windows are multicast-bound --------------------------------------
    def trainer_buffers(self) -> dict:
        """``Trainer(..., buffers=ex.trainer_buffers())``: theta_base (fp32) and the bf16 compute copy are this rank's ``base`` /
        ``base16`` window regions, so a peer's averaging kernel can land the new base in them with one ``multimem.st``."""
        out = {"base": self.win.local("base", torch.float32)}
        if self.with_base16:
            out["p16"] = self.win.local("base16", torch.bfloat16)
        return out

    def adopted(self, trainer) -> bool:
        return bool(self.with_base16 and getattr(trainer, "is_cuda", False) and trainer.base.data_ptr() == self.win.ptr("base")
                    and trainer.p16.data_ptr() == self.win.ptr("base16"))

    def can_push(self, trainer) -> bool:
        """The averaged-base broadcast can ride on the averaging kernel (NVLS multicast stores into every rank's arenas)."""
        return bool(self.win.mc_ptr) and self.adopted(trainer) and os.environ.get("DTB200_NO_PUSH", "0") != "1"

    def push_average(self, base: torch.Tensor, deltas, dscales, w: torch.Tensor, round: int, mode: int, active=None,
                     wait_flags=None) -> None:
        """Reduce-scatter + BROADCAST in one kernel: this rank's shard of ``s_j*base + sum_i w_ij delta_i`` is stored with
        ``multimem.st`` to the multicast addresses of the ``base`` (fp32) and ``base16`` (bf16) regions, i.e. it lands in
        every rank's theta_base and compute copy; then the base flag is published.  ``deltas``: peer window pointers (uniform
        mixer) or the local transposed shards of the learned mixer."""
        cs, _, _ = self.man.seg_table(base.device)
        nchunks = cs.numel()
        per = (nchunks + self.world - 1) // self.world
        c0, c1 = min(nchunks, self.rank * per), min(nchunks, (self.rank + 1) * per)
        self.nan_flags.zero_()
        ops.weighted_avg(base, deltas, w, self.man, [], None, dscales=dscales, nan_flags=self.nan_flags, wait_flags=wait_flags,
                         wait_value=round if wait_flags is not None else 0, error_flag=self.win.error_flag, chunk_range=(c0, c1),
                         mode=mode, active=active, mc_f32=self.win.mc("base"), mc_bf16=self.win.mc("base16"))
        self._base_round = round + 1
        self.win.publish(self.F_BASE, self._base_round, multicast=True)  # same path as the multimem.st data, release-ordered
        self.base_target.fill_(self._base_round)  # device-resident target of the in-kernel flag acquires (first forward GEMMs)

    def wait_base(self) -> None:
        """Stream-ordered wait until every shard owner has published the current base round (all shards have landed here)."""
        self.win.wait(self.F_BASE, self._base_round)

    def delta_is_bad(self, src: int, round: int) -> bool:
        """Host read of the NaN verdict miner ``src`` attached to its publish of ``round``."""
        return int(self.win.flags()[self.F_BAD + src].item()) == round

    # -- averager side ------------------------------------------------------------------------------------------------
    def prepare_round(self, round: int, miners: Sequence[int], w: Optional[torch.Tensor] = None, init_w: bool = False) -> torch.Tensor:
        """Stream-ordered start of an averaging round: wait for the miners' publish flags, collect their NaN verdicts into
        ``self.active`` (int32 [len(miners)], 1 = take part) and optionally reset ``w`` to 1/n_active (reference: a NaN or
        missing delta is skipped and ``w`` is rebuilt for the remaining miners, averaging_logic.py:396-430).  No host sync."""
        df = [self.win.flag_ptr(self.F_DELTA + r) for r in miners]
        bf = [self.win.flag_ptr(self.F_BAD + r) for r in miners]
        self.active = self.active[:len(miners)] if self.active.numel() >= len(miners) else torch.ones(
            len(miners), dtype=torch.int32, device=self.win.device)
        ops.round_prepare(df, bf, round, self.active, self.n_active, w, init_w, self.win.error_flag)
        return self.active

    def _delta_ptrs(self, round: int, miners: Sequence[int]) -> Tuple[List[int], Optional[List[int]]]:
        name, sname = f"delta{round & 1}", f"scale{round & 1}"
        d = [self.win.ptr(name, r) for r in miners]
        s = [self.win.ptr(sname, r) for r in miners] if self.delta_dtype_name == "fp8" else None
        return d, s

    def gather_average(self, base: torch.Tensor, w: torch.Tensor, round: int, miners: Sequence[int], out_f32,
                       out_bf16=None, wait: bool = True, active: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Pull form of fused kernel (a): this rank reads every miner's delta over NVLink and writes theta_new locally."""
        d, s = self._delta_ptrs(round, miners)
        wf = [self.win.flag_ptr(self.F_DELTA + r) for r in miners] if wait else None
        self.nan_flags.zero_()
        mode = {"fp32": 0, "bf16": 1, "fp8": 2}[self.delta_dtype_name]
        return ops.weighted_avg(base, d, w, self.man, [out_f32], [out_bf16] if out_bf16 is not None else None, dscales=s,
                                nan_flags=self.nan_flags, wait_flags=wf, wait_value=round, error_flag=self.win.error_flag,
                                mode=mode, active=active)

    def sharded_average_broadcast(self, base: torch.Tensor, w: torch.Tensor, round: int, miners: Sequence[int]) -> torch.Tensor:
        """Reduce-scatter + all-gather form: every rank reduces its shard of the arena from ALL miners' windows and
        stores the result into EVERY rank's ``base`` window (and bf16 ``base16``) -- one kernel per rank, no NCCL.
        Returns this rank's (complete, after the flag wait) fp32 base window."""
        cs, _, _ = self.man.seg_table(base.device)
        nchunks = cs.numel()
        per = (nchunks + self.world - 1) // self.world
        c0, c1 = min(nchunks, self.rank * per), min(nchunks, (self.rank + 1) * per)
        d, s = self._delta_ptrs(round, miners)
        wf = [self.win.flag_ptr(self.F_DELTA + r) for r in miners]
        outs_f32 = [self.win.ptr("base", r) for r in range(self.world)]
        outs_b16 = [self.win.ptr("base16", r) for r in range(self.world)] if self.with_base16 else None
        self.nan_flags.zero_()
        mode = {"fp32": 0, "bf16": 1, "fp8": 2}[self.delta_dtype_name]
        ops.weighted_avg(base, d, w, self.man, outs_f32, outs_b16, dscales=s, nan_flags=self.nan_flags, wait_flags=wf,
                         wait_value=round, error_flag=self.win.error_flag, chunk_range=(c0, c1), mode=mode)
        self._base_round = round + 1
        self.win.publish(self.F_BASE, self._base_round)
        self.win.wait(self.F_BASE, self._base_round)
        return self.win.local("base", torch.float32)[:self.man.total]

    def reduce_scatter_average(self, base: torch.Tensor, w: torch.Tensor, round: int, miners: Sequence[int],
                               active: Optional[torch.Tensor] = None) -> int:
        """Phase 1 of the pull-only round: this rank reduces ITS shard of the arena from all miners' windows (P2P loads)
        into its own base window and publishes the base flag.  Returns chunks_per_rank."""
        cs, _, _ = self.man.seg_table(base.device)
        nchunks = cs.numel()
        per = (nchunks + self.world - 1) // self.world
        c0, c1 = min(nchunks, self.rank * per), min(nchunks, (self.rank + 1) * per)
        d, s = self._delta_ptrs(round, miners)
        wf = [self.win.flag_ptr(self.F_DELTA + r) for r in miners]
        self.nan_flags.zero_()
        mode = {"fp32": 0, "bf16": 1, "fp8": 2}[self.delta_dtype_name]
        ops.weighted_avg(base, d, w, self.man, [self.win.ptr("base", self.rank)], None, dscales=s, nan_flags=self.nan_flags,
                         wait_flags=wf, wait_value=round, error_flag=self.win.error_flag, chunk_range=(c0, c1), mode=mode,
                         active=active)
        self._base_round = round + 1
        self.win.publish(self.F_BASE, self._base_round)
        return per

    def all_gather_reset(self, trainer, chunks_per_rank: int, reset_moments: bool = True) -> None:
        """Phase 2: pull every shard from its owner's window, fused with the optimizer/base reset of ``trainer``."""
        src = [self.win.ptr("base", r) for r in range(self.world)]
        wf = [self.win.flag_ptr(self.F_BASE + r) for r in range(self.world)]
        ops.shard_pull_reset(src, self.man, chunks_per_rank, trainer.base, trainer.master, trainer.p16 if trainer.is_cuda else None,
                             trainer.m, trainer.v, reset_moments=reset_moments, wait_flags=wf, wait_value=self._base_round,
                             error_flag=self.win.error_flag)
        # nobody may overwrite a shard (next round's phase 1) before every rank has pulled it
        self.win.device_barrier()

    def publish_base(self, base: torch.Tensor, round: int, dst_ranks: Optional[List[int]] = None) -> None:
        """Averager -> everyone: P2P stores of the fp32 base (+bf16 copy) into each rank's landing window, then flag."""
        dst = list(range(self.world)) if dst_ranks is None else dst_ranks
        n = self.man.total
        for r in dst:
            self.win.peer("base", r, torch.float32)[:n].copy_(base, non_blocking=True)
            if self.with_base16:
                self.win.peer("base16", r, torch.bfloat16)[:n].copy_(base, non_blocking=True)
        self._base_round = round
        self.win.publish(self.F_BASE, round, dst)

    def base_round(self, src: Optional[int] = None) -> int:
        """Newest base round visible in this rank's flag page.  ``SymmetricWindow.publish`` writes slot
        ``F_BASE + <publisher rank>``, and the averager may sit on ANY rank (``--roles miner:0-6,averager:7``), so the
        default is the maximum over all publisher slots; ``src`` reads one publisher's slot."""
        f = self.win.flags()
        if src is not None:
            return int(f[self.F_BASE + src].item())
        return int(f[self.F_BASE:self.F_BASE + max(self.world, 1)].max().item())

    def base_view(self) -> torch.Tensor:
        return self.win.local("base", torch.float32)[:self.man.total]

    def base16_ptr(self, rank: int) -> int:
        return self.win.ptr("base16", rank)

    def fetch_base(self, out: torch.Tensor) -> int:
        out.copy_(self.base_view())
        return self.base_round()


class NvlsExchange(Exchange):
    """Uniform / pre-scaled averaging through NVSwitch multicast objects (NVLS).

    One symmetric allocation per rank (``torch.distributed._symmetric_memory``: CUDA VMM + ``cuMulticast*``) holds this
    rank's fp32 delta and the landing buffer of the new base.  A round is: emit the delta locally -> device barrier ->
    ``nvls_avg`` on this rank's shard (``multimem.ld_reduce`` sums the N deltas in the switch, ``multimem.st`` lands the
    result on every rank) -> device barrier -> local optimizer reset from the landing buffer.
    The learned per-(miner, tensor) mixer needs the individual deltas on the averager and keeps using ``PeerExchange``;
    this plane serves ``--mixer uniform`` (reference DeltaAverager with equal weights, averaging_logic.py:200-269).
    """

    name = "nvls"

    def __init__(self, manifest: Manifest, group=None, device: Optional[torch.device] = None):
        import torch.distributed._symmetric_memory as symm_mem
        assert dist.is_initialized() and torch.cuda.is_available()
        self.man = manifest
        self.group = group or dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        n = manifest.total
        quantum = 4 * self.world * 32                      # float4 shards, 32 float4 per warp-row
        self.n, self.npad = n, (n + quantum - 1) // quantum * quantum
        self.buf = symm_mem.empty(2 * self.npad, dtype=torch.float32, device=self.device)
        self.handle = symm_mem.rendezvous(self.buf, self.group.group_name)
        if not self.handle.multicast_ptr:
            raise RuntimeError("this system exposes no NVLS multicast object (multicast_ptr == 0)")
        self.buf.zero_()
        self.delta = self.buf[:self.npad]
        self.landing = self.buf[self.npad:]
        self.mc_delta = int(self.handle.multicast_ptr)
        self.mc_landing = self.mc_delta + 4 * self.npad
        per4 = self.npad // 4 // self.world
        self.lo4, self.hi4 = self.rank * per4, (self.rank + 1) * per4
        self.handle.barrier()
        self._round = 0

    def publish_delta(self, trainer, round: int) -> None:
        trainer.emit_delta(self.delta[:self.n])

    def average_broadcast(self, base: torch.Tensor, scale: Optional[float] = None, base_scale: float = 1.0) -> torch.Tensor:
        """new_base (on EVERY rank, in ``self.landing``) = base_scale * base + scale * sum_i delta_i; default scale 1/N."""
        self.handle.barrier()  # every rank's delta is complete and visible
        ops.nvls_avg(self.mc_delta, self.mc_landing, base, self.lo4, self.hi4, scale if scale is not None else 1.0 / self.world,
                     base_scale)
        self.handle.barrier()  # every shard has landed everywhere
        self._round += 1
        return self.landing[:self.n]

    def delta_round(self, src: int) -> int:
        return self._round
