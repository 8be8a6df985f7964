"""Offline stand-in for the ``bittensor`` package -- ONLY so that the unmodified reference (``baseline/_ref/hivetrain``)
can be imported and its own miner loop executed by ``bench.py --impl reference`` (bittensor==6.10.1 is not installable
without network access).  Nothing in the product imports this."""
import argparse
import logging as _pylog
import sys
import types

__version__ = "6.10.1-shim"

_log = _pylog.getLogger("bittensor-shim")
if not _log.handlers:
    _h = _pylog.StreamHandler(sys.stderr)
    _h.setFormatter(_pylog.Formatter("%(asctime)s | ref | %(message)s", "%H:%M:%S"))
    _log.addHandler(_h)
    _log.setLevel(_pylog.WARNING)
    _log.propagate = False


class _Logging:
    def __call__(self, *a, **k):
        return self

    @staticmethod
    def add_args(parser):
        parser.add_argument("--logging.debug", action="store_true")
        parser.add_argument("--logging.trace", action="store_true")
        parser.add_argument("--logging.logging_dir", type=str, default="~/.bittensor/miners")

    def enable_debug(self, *a, **k):
        pass

    enable_default = enable_trace = set_debug = set_trace = enable_debug

    def info(self, *a, **k):
        _log.info(" ".join(str(x) for x in a))

    def debug(self, *a, **k):
        _log.debug(" ".join(str(x) for x in a))

    def warning(self, *a, **k):
        _log.warning(" ".join(str(x) for x in a))

    def error(self, *a, **k):
        _log.error(" ".join(str(x) for x in a))

    success = trace = info


logging = _Logging()
btlogging = types.ModuleType("bittensor.btlogging")
btlogging.logging = logging
sys.modules["bittensor.btlogging"] = btlogging


class _Hotkey:
    def __init__(self, name):
        self.ss58_address = name


class wallet:
    def __init__(self, config=None, name="default", hotkey="default", **kw):
        w = getattr(config, "wallet", None)
        self.name = getattr(w, "name", name)
        self.hotkey_str = getattr(w, "hotkey", hotkey)
        self.hotkey = _Hotkey(self.hotkey_str)
        self.coldkeypub = _Hotkey(self.name)

    @staticmethod
    def add_args(parser):
        parser.add_argument("--wallet.name", type=str, default="default")
        parser.add_argument("--wallet.hotkey", type=str, default="default")
        parser.add_argument("--wallet.path", type=str, default="~/.bittensor/wallets")

    def __repr__(self):
        return f"wallet({self.name}, {self.hotkey_str})"


class _Metagraph:
    def __init__(self, netuid, n=1):
        import torch
        self.netuid, self.n = netuid, n
        self.hotkeys = ["default"]
        self.uids = torch.arange(n)
        self.S = torch.zeros(n)
        self.W = torch.zeros(n, n)
        self.last_update = torch.zeros(n)
        self.block = torch.tensor(0)

    def sync(self, *a, **k):
        pass


class subtensor:
    def __init__(self, config=None, **kw):
        self.network = "shim"
        self.chain_endpoint = "none"

    @staticmethod
    def add_args(parser):
        parser.add_argument("--subtensor.network", type=str, default="local")
        parser.add_argument("--subtensor.chain_endpoint", type=str, default="")

    def metagraph(self, netuid, lite=True):
        return _Metagraph(netuid)

    def is_hotkey_registered(self, netuid=None, hotkey_ss58=None):
        return True

    def get_current_block(self):
        return 0

    block = property(lambda self: 0)

    def commit(self, wallet, netuid, data):
        return True

    def set_weights(self, **kw):
        return True, "shim"


class axon:
    def __init__(self, *a, **k):
        pass

    @staticmethod
    def add_args(parser):
        parser.add_argument("--axon.port", type=int, default=8091)
        parser.add_argument("--axon.ip", type=str, default="[::]")


class _Cfg(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def config(parser: argparse.ArgumentParser, args=None):
    ns, _ = parser.parse_known_args(args)
    root = _Cfg()
    for key, val in vars(ns).items():
        node = root
        parts = key.split(".")
        for p in parts[:-1]:
            node = node.setdefault(p, _Cfg())
        node[parts[-1]] = val
    return root


metagraph = _Metagraph
extrinsics = types.SimpleNamespace(serving=types.SimpleNamespace(get_metadata=lambda *a, **k: None))
