"""``hf_manager``-shaped push/pull of delta tensors and averaged models -- without the HuggingFace Hub.

Reference contract (hivetrain/hf_manager.py:12-197): ``push_changes``, ``push_to_hf_hub``,
``get_latest_commit_sha``, ``check_for_new_submissions``, ``pull_latest_model``, ``update_model``,
``receive_gradients``, ``get_local_*_directory``, ``clear_hf_cache``, ``git_prune_and_refresh``; plus the disk-backed
fake ``LocalHFManager`` (:200-241).

The same method names are served by an :class:`~distributedtraining_b200.parallel.exchange.Exchange`:

* ``peer``  -- a "repo" is a rank's symmetric window; ``push_changes`` = delta already written in place + release-store
  of the round flag; ``receive_gradients`` = *views* onto the peer window; the "commit SHA" is the round counter;
* ``disk``  -- files under a shared directory (atomic replace instead of git push + ``sleep(10)``).
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Union

import torch

from .models.arena import Manifest
from .models.transformer import pack_any
from .parallel.exchange import DiskExchange, Exchange
from .utils.logging import logger


def parse_repo(repo_id: Union[str, int, None]) -> Optional[int]:
    """'peer://3', 'disk://3', 'rank3', '3' or 3 -> 3."""
    if repo_id is None:
        return None
    if isinstance(repo_id, int):
        return repo_id
    s = str(repo_id)
    for pre in ("peer://", "disk://", "rank://", "rank"):
        if s.startswith(pre):
            s = s[len(pre):]
            break
    try:
        return int(s.split("/")[-1])
    except ValueError:
        return None


class HFManager:
    def __init__(self, local_dir: str = ".", hf_token: Optional[str] = None, my_repo_id=None, averaged_model_repo_id=None,
                 model_dir: Optional[str] = None, device: str = "cuda", exchange: Optional[Exchange] = None,
                 manifest: Optional[Manifest] = None, model_config=None):
        self.my_repo_id = my_repo_id
        self.model_repo_id = averaged_model_repo_id
        self.hf_token = hf_token  # accepted for signature compatibility; never used
        self.device = device
        self.exchange = exchange
        self.manifest = manifest if manifest is not None else getattr(exchange, "man", None)
        self.model_config = model_config  # ModelConfig: lets file-based pushes accept HF-layout (reference-format) dicts
        self.local_dir = local_dir
        self.local_gradient_dir = os.path.join(local_dir, str(my_repo_id).split("/")[-1]) if my_repo_id is not None else None
        self.model_dir = model_dir if model_dir else os.path.join(local_dir, str(averaged_model_repo_id).split("/")[-1])
        self.round = 0  # number of deltas published by this manager
        self._staged_base: Optional[torch.Tensor] = None
        self._published_round = 0  # last base round this manager published (averager side)
        # record the current version so that the first check is False until the averager publishes again
        # (reference hf_manager.py:54)
        self.latest_model_commit_sha = self.get_latest_commit_sha(self.model_repo_id)

    # -- versions ---------------------------------------------------------------------------------------------------
    def get_latest_commit_sha(self, repo_id=None) -> Optional[str]:
        try:
            return str(self.exchange.base_round())
        except Exception as e:
            logger.warning(f"Failed to fetch latest version: {e}")
            return None

    def check_for_new_submissions(self, repo_id=None) -> bool:
        current = self.get_latest_commit_sha(repo_id)
        if current is not None and current != self.latest_model_commit_sha:
            self.latest_model_commit_sha = current
            return True
        return False

    # -- miner side ---------------------------------------------------------------------------------------------------
    def push_changes(self, file_to_send: Union[str, None] = "weight_diff.pt", trainer=None) -> None:
        """Publish this miner's delta.  With ``trainer`` the delta is emitted straight into the exchange (zero copy on
        the peer plane); with a file name, ``<gradient_dir>/<file>`` (a ``dict[name, Tensor]``) is packed and published."""
        try:
            self.round += 1
            if trainer is None:
                path = os.path.join(self.get_local_gradient_directory(), file_to_send)
                sd = torch.load(path, map_location="cpu", weights_only=False)
                trainer = _DictDelta(self.manifest, sd, self.model_config)
            self.exchange.publish_delta(trainer, self.round)
        except Exception as e:  # best-effort, as in the reference (hf_manager.py:113-114)
            logger.warning(f"Failed to push changes: {e}")

    # -- averager / validator side --------------------------------------------------------------------------------------
    def receive_flat(self, miner_repo_id, min_round: int = 0) -> Optional[torch.Tensor]:
        src = parse_repo(miner_repo_id)
        if src is None:
            return None
        try:
            return self.exchange.fetch_delta(src, min_round)
        except Exception as e:
            logger.warning(f"Error receiving delta from {miner_repo_id}: {e}")
            return None

    def delta_round(self, miner_repo_id) -> int:
        src = parse_repo(miner_repo_id)
        try:
            return 0 if src is None else int(self.exchange.delta_round(src))
        except Exception:
            return 0

    def receive_gradients(self, miner_repo_id, weights_file_name: str = "weight_diff.pt") -> Optional[Dict[str, torch.Tensor]]:
        """``dict[name -> Tensor]`` *views* of the miner's delta, or ``None`` (same contract as hf_manager.py:186-197)."""
        flat = self.receive_flat(miner_repo_id)
        if flat is None:
            return None
        return self.manifest.views(flat)

    def push_to_hf_hub(self, path_to_model=None, commit_message: str = "Pushing model to Hub", base: Optional[torch.Tensor] = None,
                       round: Optional[int] = None) -> None:
        try:
            if base is None:
                blob = torch.load(path_to_model, map_location="cpu", weights_only=False)
                base = pack_any(self.manifest, blob, self.model_config)
            # the published round is tracked HERE as well: a peer-plane publisher only sees its own slot after the flag
            # kernel has run, and re-deriving "next" from a stale read would republish the same round forever
            nxt = (max(int(self.exchange.base_round()), self._published_round) + 1) if round is None else int(round)
            self.exchange.publish_base(base, nxt)
            self._published_round = nxt
            self.latest_model_commit_sha = str(nxt)
        except Exception as e:
            logger.warning(f"Failed to push model: {e}")

    # -- everybody: pull the averaged model ----------------------------------------------------------------------------
    def pull_latest_model(self) -> None:
        """Make the newest base available locally.  Peer plane: it already sits in this rank's landing window."""
        n = self.manifest.total
        if hasattr(self.exchange, "base_view"):
            self._staged_base = self.exchange.base_view()
        else:
            self._staged_base = torch.empty(n, dtype=torch.float32)
            self.exchange.fetch_base(self._staged_base)

    def update_model(self, model, model_file_name: str = "averaged_model.pt", lr: Optional[float] = None,
                     reset_optimizer: bool = True):
        """Load the pulled base into ``model`` (a Trainer / ModuleTrainer) -- theta = theta_base = averaged model."""
        if self._staged_base is None:
            self.pull_latest_model()
        base = self._staged_base
        model.load_base(base.to(model.master.device) if base.device != model.master.device else base, lr=lr,
                        reset_optimizer=reset_optimizer)
        return model

    # -- paths / GC ---------------------------------------------------------------------------------------------------
    def get_local_gradient_directory(self) -> Optional[str]:
        if self.local_gradient_dir:
            os.makedirs(self.local_gradient_dir, exist_ok=True)
        return self.local_gradient_dir

    def get_local_model_directory(self) -> str:
        os.makedirs(self.model_dir, exist_ok=True)
        return self.model_dir

    @staticmethod
    def clear_hf_cache() -> None:
        """Nothing is cached: peer deltas are views, disk deltas are read on demand."""

    @staticmethod
    def git_prune_and_refresh(repo_path: str) -> None:
        """Remove stale temp files left by crashed writers (the closest analogue of ``git lfs prune``)."""
        if repo_path and os.path.isdir(repo_path):
            for f in os.listdir(repo_path):
                if ".tmp." in f:
                    try:
                        os.remove(os.path.join(repo_path, f))
                    except OSError:
                        pass


class _DictDelta:
    """Adapter: a ``dict[name, Tensor]`` delta presented with the ``emit_delta`` interface of a trainer."""

    def __init__(self, manifest: Manifest, sd: Dict[str, torch.Tensor], cfg=None):
        self.flat = pack_any(manifest, sd, cfg)
        self.master = self.flat

    def emit_delta(self, out, scales=None, bad=None):
        out.copy_(self.flat.to(out.device, out.dtype))
        if bad is not None and not bool(torch.isfinite(self.flat).all()):
            bad.fill_(1)
        return out


class LocalHFManager(HFManager):
    """Shared-directory hub (reference hf_manager.py:200-241): "new submission" == sha256 of ``averaged_model.pt`` changed."""

    def __init__(self, my_repo_id=".", averaged_model_repo_id=".", device: str = "cpu", manifest: Optional[Manifest] = None,
                 rank: int = 0, delta_dtype: str = "fp32", model_config=None):
        self.root = str(averaged_model_repo_id)
        exchange = DiskExchange(self.root, rank, manifest, delta_dtype)
        self.last_known_hash: Optional[str] = None
        super().__init__(local_dir=str(my_repo_id), my_repo_id=rank, averaged_model_repo_id=averaged_model_repo_id,
                         model_dir=os.path.join(self.root, "base"), device=device, exchange=exchange, manifest=manifest,
                         model_config=model_config)
        self.local_gradient_dir = str(my_repo_id)
        self.last_known_hash = exchange.base_hash()

    def set_model_hash(self, model_hash: Optional[str]) -> None:
        self.last_known_hash = model_hash
        try:
            with open(os.path.join(self.get_local_model_directory(), "model_hash.txt"), "w") as f:
                f.write(str(model_hash))
        except OSError:
            pass

    def check_for_new_submissions(self, repo_id=None) -> bool:
        h = self.exchange.base_hash()
        if h is not None and h != self.last_known_hash:
            self.set_model_hash(h)
            return True
        return False
