"""Reference hivetrain/config/mlflow_config.py:1-3 (mlflow is an optional sink; off by default)."""
MLFLOW_UI_URL = ""
CURRENT_MODEL_NAME = "openai-community/gpt2"
MLFLOW_ACTIVE = False
