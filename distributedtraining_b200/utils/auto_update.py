"""Version-poll auto-updater (reference hivetrain/utils/auto_update.py:6-64 and run_miner.sh:129-155,233-268).

The reference polls GitHub for a ``__version__`` bump every 1800 s, then ``git reset --hard`` + ``pip install -e .`` +
``pm2 restart``.  There is no network here, so the "remote" is any git remote/path; the restart is delegated to the
supervisor (:mod:`.supervisor`) through its exit-code protocol.
"""
from __future__ import annotations

import os
import re
import subprocess
import time
from typing import Callable, Optional

from .logging import logger

UPDATE_EXIT_CODE = 75  # the supervisor restarts the child after pulling when it sees this code


def read_version_value(path: str) -> Optional[str]:
    try:
        m = re.search(r"__version__\s*=\s*['\"]([^'\"]+)['\"]", open(path).read())
        return m.group(1) if m else None
    except OSError:
        return None


def get_version_difference(a: str, b: str) -> int:
    """Signed distance between two dotted versions (major*10000 + minor*100 + patch), as in run_miner.sh:40-75."""
    def num(v):
        p = [int(x) for x in (v.split(".") + ["0", "0"])[:3]]
        return p[0] * 10000 + p[1] * 100 + p[2]
    return num(b) - num(a)


def check_variable_value_on_remote(repo_dir: str, remote: str = "origin", branch: str = "main",
                                   rel_path: str = "template/__init__.py") -> Optional[str]:
    try:
        subprocess.run(["git", "-C", repo_dir, "fetch", remote, branch], check=True, capture_output=True, timeout=60)
        out = subprocess.run(["git", "-C", repo_dir, "show", f"{remote}/{branch}:{rel_path}"], check=True, capture_output=True,
                             text=True, timeout=30).stdout
        m = re.search(r"__version__\s*=\s*['\"]([^'\"]+)['\"]", out)
        return m.group(1) if m else None
    except Exception as e:
        logger.debug(f"remote version check failed: {e}")
        return None


def update_checkout(repo_dir: str, remote: str = "origin", branch: str = "main") -> bool:
    try:
        subprocess.run(["git", "-C", repo_dir, "reset", "--hard", f"{remote}/{branch}"], check=True, capture_output=True, timeout=60)
        return True
    except Exception as e:
        logger.warning(f"update failed: {e}")
        return False


def monitor_repo(repo_dir: str, interval: float = 1800.0, on_update: Optional[Callable[[str, str], None]] = None,
                 max_checks: Optional[int] = None) -> Optional[str]:
    """Poll the remote; on a version bump pull and call ``on_update(old, new)`` (default: exit with UPDATE_EXIT_CODE)."""
    local_file = os.path.join(repo_dir, "template", "__init__.py")
    checks = 0
    while max_checks is None or checks < max_checks:
        checks += 1
        cur, new = read_version_value(local_file), check_variable_value_on_remote(repo_dir)
        if cur and new and get_version_difference(cur, new) > 0:
            logger.info(f"new version {new} (running {cur}); updating")
            if update_checkout(repo_dir):
                if on_update:
                    on_update(cur, new)
                    return new
                raise SystemExit(UPDATE_EXIT_CODE)
        if max_checks is None or checks < max_checks:
            time.sleep(interval)
    return None
