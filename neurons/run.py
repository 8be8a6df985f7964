#!/usr/bin/env python
"""Role dispatcher for multi-process launches: every rank runs the neuron its ``--roles`` entry names.

    torchrun --nproc-per-node 3 neurons/run.py --roles miner:0-1,averager:2 --device cpu --backend disk --model gpt2-tiny ...
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from distributedtraining_b200.config import Configurator  # noqa: E402
from distributedtraining_b200.parallel.launch import env_rank_world, parse_roles  # noqa: E402


def main(argv=None):
    cfg = Configurator.combine_configs(argv)
    rank, world, _ = env_rank_world(cfg)
    roles = parse_roles(cfg.roles, world)
    mine = [r for r, ranks in roles.items() if rank in ranks]
    role = "miner" if "miner" in mine else (mine[0] if mine else "miner")
    if len(mine) > 1:
        print(f"rank {rank}: several roles {mine}; running {role} (use bench.py / LocalSGDCoordinator for co-located roles)")
    if role == "miner":
        import miner as m
    elif role == "validator":
        import validator as m
    else:
        import averager as m
    return m.main(argv)


if __name__ == "__main__":
    main()
