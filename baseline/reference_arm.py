"""Reference arm of bench.py (placeholder until the shimmed reference install lands)."""


def run_reference(args):
    return {"impl": "reference", "unavailable": "reference install pending (bittensor/mlflow/hivemind not installable offline)"}
