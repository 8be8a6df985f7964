"""Bulk identity generation for test subnets (reference hivetrain/utils/generate_wallets.py:9-47 mass-creates, funds and
registers bittensor wallets).  Here an identity is a hotkey string + an HMAC secret stored under ``wallet.path``; the
"registration" is a ledger entry (and optional stake) -- enough to populate a LocalBittensorNetwork / JSON ledger."""
from __future__ import annotations

import json
import os
import secrets
from typing import Dict, List


def generate_multiple_wallets(n: int, path: str = "~/.dtb200/wallets", prefix: str = "test", ledger=None, stake: float = 10.0,
                              validators: int = 0, validator_stake: float = 10000.0) -> List[Dict[str, str]]:
    path = os.path.expanduser(path)
    os.makedirs(path, exist_ok=True)
    out = []
    for i in range(n):
        hk = f"{prefix}_hotkey_{i}"
        w = {"name": f"{prefix}_coldkey_{i}", "hotkey": hk, "secret": secrets.token_hex(16),
             "stake": validator_stake if i >= n - validators else stake}
        with open(os.path.join(path, f"{hk}.json"), "w") as f:
            json.dump(w, f)
        if ledger is not None:
            ledger.put(f"hotkey/{hk}", "1")
            ledger.put(f"stake/{hk}", str(w["stake"]))
        out.append(w)
    return out


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("-n", type=int, default=8)
    ap.add_argument("--path", default="~/.dtb200/wallets")
    a = ap.parse_args()
    print(json.dumps(generate_multiple_wallets(a.n, a.path), indent=1))
