"""distributedtraining_b200 -- a Blackwell-native local-SGD / weight-delta-averaging training framework.

Same roles and API surface as bit-current/DistributedTraining ("hivetrain"): miners train a private copy and emit a
weight delta (:mod:`.training_manager`), a validator scores deltas by loss drop (:mod:`.validation_logic`), an averager
learns a per-miner, per-tensor weighted parameter average into the next base model (:mod:`.averaging_logic`), all behind
an ``hf_manager``-shaped push/pull interface (:mod:`.hf_manager`) -- but co-located as ranks of one 8xB200 box with the
delta exchange in NVSwitch peer memory and the hot ops as hand-written sm_100a kernels (:mod:`.ops`, ``csrc/``).

Version scheme follows reference hivetrain/__init__.py:1-10 (``__spec_version__`` is the ``version_key`` of set_weights).
"""
__version__ = "0.1.0"
version_split = __version__.split(".")
__spec_version__ = (100 * int(version_split[0])) + (10 * int(version_split[1])) + (1 * int(version_split[2]))

# NOTE: unlike the reference, importing the package has NO side effects (the reference parses the CLI and connects to
# the chain at import time: hivetrain/training_manager.py:22-24).
