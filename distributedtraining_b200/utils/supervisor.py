"""Process supervision (the pm2 half of the reference's run_miner.sh / run_validator.sh: ``min_uptime 5m``,
``max_restarts 5``, restart on auto-update: run_miner.sh:214-268).  Pure Python, no pm2/jq/curl.

    python -m distributedtraining_b200.utils.supervisor --max-restarts 5 --min-uptime 300 -- python neurons/miner.py ...
"""
from __future__ import annotations

import argparse
import signal
import subprocess
import sys
import time
from typing import List

from .auto_update import UPDATE_EXIT_CODE
from .logging import logger


def supervise(cmd: List[str], max_restarts: int = 5, min_uptime: float = 300.0, backoff: float = 2.0) -> int:
    restarts = 0
    child = None

    def forward(sig, frame):  # forward termination to the exact child PID we started
        if child and child.poll() is None:
            child.send_signal(sig)

    signal.signal(signal.SIGTERM, forward)
    signal.signal(signal.SIGINT, forward)
    while True:
        t0 = time.time()
        child = subprocess.Popen(cmd)
        rc = child.wait()
        up = time.time() - t0
        if rc == 0:
            logger.info("child exited cleanly")
            return 0
        if rc == UPDATE_EXIT_CODE:
            logger.info("child requested restart after update")
            continue
        if rc < 0 and -rc in (signal.SIGTERM, signal.SIGINT):
            return 128 - rc
        if up >= min_uptime:
            restarts = 0  # a long healthy run resets the crash budget (pm2 min_uptime semantics)
        restarts += 1
        if restarts > max_restarts:
            logger.error(f"child crashed {restarts} times within {min_uptime}s uptime; giving up (rc={rc})")
            return rc
        logger.warning(f"child exited rc={rc} after {up:.0f}s; restart {restarts}/{max_restarts}")
        time.sleep(backoff * restarts)


def main(argv=None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-restarts", type=int, default=5)
    ap.add_argument("--min-uptime", type=float, default=300.0)
    ap.add_argument("cmd", nargs=argparse.REMAINDER)
    a = ap.parse_args(argv)
    cmd = a.cmd[1:] if a.cmd and a.cmd[0] == "--" else a.cmd
    if not cmd:
        ap.error("no command given")
    return supervise(cmd, a.max_restarts, a.min_uptime)


if __name__ == "__main__":
    sys.exit(main())
