"""Optional MLflow sink + system metrics (reference hivetrain/utils/mlflow_utils.py:15-180).

MLflow is not installed in the build image and is gated off by default (``MLFLOW_ACTIVE=False``, as in the reference);
when inactive, metrics go to the JSONL logger only.  The system-metric helpers are real (psutil / NVML / torch).
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch

from ..config.mlflow_config import CURRENT_MODEL_NAME, MLFLOW_ACTIVE, MLFLOW_UI_URL
from .logging import logger

try:  # optional
    import mlflow  # type: ignore
except Exception:  # pragma: no cover
    mlflow = None


def get_gpu_utilization(device: int = 0) -> float:
    try:
        return float(torch.cuda.utilization(device)) if torch.cuda.is_available() else 0.0
    except Exception:
        return 0.0


def get_cpu_utilization() -> float:
    import psutil
    return float(psutil.cpu_percent(interval=None))


def get_memory_usage() -> float:
    import psutil
    return float(psutil.virtual_memory().percent)


def get_network_bandwidth() -> Dict[str, float]:
    import psutil
    c = psutil.net_io_counters()
    return {"bytes_sent": float(c.bytes_sent), "bytes_recv": float(c.bytes_recv)}


def get_version_from_file(path: Optional[str] = None) -> str:
    from .. import __version__
    return __version__


VERSION = get_version_from_file()


def initialize_mlflow(role: str, device, version: Optional[str], mlflow_ui_url: str = MLFLOW_UI_URL,
                      current_model_name: str = CURRENT_MODEL_NAME, my_hotkey: Optional[str] = None,
                      learning_rate: Optional[float] = None, send_interval: Optional[float] = None,
                      check_update_interval: Optional[float] = None) -> bool:
    if not MLFLOW_ACTIVE or mlflow is None:
        return False
    try:
        os.environ["MLFLOW_ENABLE_SYSTEM_METRICS_LOGGING"] = "true"
        mlflow.set_tracking_uri(mlflow_ui_url)
        mlflow.set_experiment(current_model_name)
        run_name = "AVERAGER" if role == "averager" else f"{role}_{my_hotkey}"
        mlflow.start_run(run_name=run_name)
        mlflow.log_param("device", str(device))
        mlflow.log_param("Version of Code", version or VERSION)
        for k, v in (("learning_rate", learning_rate), ("send_interval", send_interval),
                     ("check_update_interval", check_update_interval)):
            if v is not None:
                mlflow.log_param(k, v)
        return True
    except Exception as e:
        logger.warning(f"mlflow init failed: {e}")
        return False


def log_model_metrics(step: int, **metrics) -> None:
    if not MLFLOW_ACTIVE or mlflow is None:
        return
    try:
        for k, v in metrics.items():
            mlflow.log_metric(k, float(v), step=step)
    except Exception as e:
        logger.warning(f"mlflow log failed: {e}")


def setup_mlflow_session(retries: int = 3, backoff: float = 0.5):
    """``requests`` session with retry adapters (reference :143-177)."""
    import requests
    from requests.adapters import HTTPAdapter
    from urllib3.util.retry import Retry

    s = requests.Session()
    r = Retry(total=retries, backoff_factor=backoff, status_forcelist=(500, 502, 503, 504))
    s.mount("http://", HTTPAdapter(max_retries=r))
    s.mount("https://", HTTPAdapter(max_retries=r))
    return s


create_mlflow_session = setup_mlflow_session
