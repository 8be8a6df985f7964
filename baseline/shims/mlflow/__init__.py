"""No-op stand-in for mlflow (MLFLOW_ACTIVE is False in the reference: hivetrain/config/mlflow_config.py:3)."""
import sys
import types


def _noop(*a, **k):
    return None


set_tracking_uri = set_experiment = start_run = end_run = log_param = log_params = log_metric = log_metrics = _noop
pytorch = types.ModuleType("mlflow.pytorch")
pytorch.log_model = _noop
sys.modules["mlflow.pytorch"] = pytorch
