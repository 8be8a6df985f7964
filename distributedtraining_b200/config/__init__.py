from .config import Config, Configurator  # noqa: F401
from .mlflow_config import CURRENT_MODEL_NAME, MLFLOW_ACTIVE, MLFLOW_UI_URL  # noqa: F401
