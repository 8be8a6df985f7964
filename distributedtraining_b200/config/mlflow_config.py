"""Optional mlflow sink (reference hivetrain/config/mlflow_config.py:1-3 hard-codes a tracking URL and keeps the sink off).

Here the three switches come from the environment so a deployment can turn the sink on without editing the package:
``DTB200_MLFLOW_ACTIVE=1``, ``DTB200_MLFLOW_URL=http://host:port``, ``DTB200_MODEL_NAME=...``.  The names the rest of the
code imports are unchanged.
"""
import os


def _flag(name: str, default: bool = False) -> bool:
    return os.environ.get(name, "1" if default else "0").strip().lower() in ("1", "true", "yes", "on")


MLFLOW_ACTIVE = _flag("DTB200_MLFLOW_ACTIVE")                       # off by default, as upstream
MLFLOW_UI_URL = os.environ.get("DTB200_MLFLOW_URL", "")             # no default tracking server
CURRENT_MODEL_NAME = os.environ.get("DTB200_MODEL_NAME", "openai-community/gpt2")
