"""Load generator for the bootstrap server (reference hivetrain/utils/bootstrap_stress.py:1-48: 100 requests x 50
concurrent x 300 s against /return_dht_address).  stdlib threads + urllib; returns latency statistics."""
from __future__ import annotations

import json
import statistics
import threading
import time
import urllib.request
from typing import Dict, List


def stress_test(url: str = "http://127.0.0.1:5000/return_dht_address", requests_per_worker: int = 100, concurrency: int = 50,
                duration: float = 300.0) -> Dict[str, float]:
    lat: List[float] = []
    errors = [0]
    lock = threading.Lock()
    deadline = time.time() + duration

    def worker():
        for _ in range(requests_per_worker):
            if time.time() > deadline:
                return
            t0 = time.time()
            try:
                with urllib.request.urlopen(url, timeout=5) as r:
                    json.loads(r.read())
                with lock:
                    lat.append(time.time() - t0)
            except Exception:
                with lock:
                    errors[0] += 1

    ts = [threading.Thread(target=worker) for _ in range(concurrency)]
    t0 = time.time()
    [t.start() for t in ts]
    [t.join() for t in ts]
    wall = time.time() - t0
    lat.sort()
    return {"requests": len(lat), "errors": errors[0], "wall_s": wall, "rps": len(lat) / max(wall, 1e-9),
            "p50_ms": 1e3 * statistics.median(lat) if lat else float("nan"),
            "p99_ms": 1e3 * lat[int(0.99 * (len(lat) - 1))] if lat else float("nan")}


if __name__ == "__main__":
    print(json.dumps(stress_test()))
