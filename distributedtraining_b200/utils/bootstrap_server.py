"""Rendezvous bootstrap service (the role of reference hivetrain/utils/bootstrap_server.py:1-115, which keeps a pool of
hivemind DHT nodes alive and serves ``initial_peers`` over HTTP -- a leftover the reference never calls).

Here it is the *real* rendezvous helper of the in-box job: a tiny stdlib HTTP server that owns a pool of
``torch.distributed.TCPStore`` masters (one per job id), health-checks and re-creates them, and returns
``{"store_address", "store_port", "world_size"}`` from ``/return_store_address`` (alias ``/return_dht_address``).
"""
from __future__ import annotations

import json
import threading
import time
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer
from typing import Dict
from urllib.parse import parse_qs, urlparse

from .logging import logger


class StorePool:
    def __init__(self, host: str = "127.0.0.1", base_port: int = 4999, max_jobs: int = 10):
        self.host, self.base_port, self.max_jobs = host, base_port, max_jobs
        self.jobs: Dict[str, dict] = {}
        self.lock = threading.Lock()

    def _create(self, job: str, world: int, port: int):
        import datetime
        from torch.distributed import TCPStore
        store = TCPStore(self.host, port, world_size=None, is_master=True, timeout=datetime.timedelta(seconds=30),
                         wait_for_workers=False)
        return {"store": store, "port": port, "world": world, "created": time.time(), "last_ok": time.time()}

    def get(self, job: str, world: int) -> dict:
        with self.lock:
            if job not in self.jobs:
                if len(self.jobs) >= self.max_jobs:
                    oldest = min(self.jobs, key=lambda j: self.jobs[j]["created"])
                    del self.jobs[oldest]
                used = {j["port"] for j in self.jobs.values()}
                port = next(p for p in range(self.base_port, self.base_port + 4 * self.max_jobs) if p not in used)
                self.jobs[job] = self._create(job, world, port)
            j = self.jobs[job]
            return {"store_address": self.host, "store_port": j["port"], "world_size": j["world"], "job": job}

    def check_and_manage(self) -> int:
        """Health-check every store (set/get round trip); re-create dead ones (reference check_and_manage_dhts :39-72)."""
        bad = 0
        with self.lock:
            for name, j in list(self.jobs.items()):
                try:
                    j["store"].set("__health__", str(time.time()))
                    j["store"].get("__health__")
                    j["last_ok"] = time.time()
                except Exception as e:
                    bad += 1
                    logger.warning(f"store for job {name} is unhealthy ({e}); re-creating")
                    try:
                        self.jobs[name] = self._create(name, j["world"], j["port"])
                    except Exception:
                        del self.jobs[name]
        return bad


def make_server(pool: StorePool, host: str = "127.0.0.1", port: int = 5000) -> ThreadingHTTPServer:
    class H(BaseHTTPRequestHandler):
        def log_message(self, *a):
            pass

        def do_GET(self):
            u = urlparse(self.path)
            if u.path in ("/return_store_address", "/return_dht_address"):
                q = parse_qs(u.query)
                body = json.dumps(pool.get(q.get("job", ["default"])[0], int(q.get("world_size", ["1"])[0]))).encode()
                self.send_response(200)
            elif u.path == "/health":
                body = json.dumps({"jobs": len(pool.jobs)}).encode()
                self.send_response(200)
            else:
                body = b"{}"
                self.send_response(404)
            self.send_header("Content-Type", "application/json")
            self.send_header("Content-Length", str(len(body)))
            self.end_headers()
            self.wfile.write(body)

    return ThreadingHTTPServer((host, port), H)


def serve(host: str = "127.0.0.1", port: int = 5000, base_port: int = 4999, check_every: float = 60.0) -> None:
    pool = StorePool(host, base_port)
    srv = make_server(pool, host, port)

    def janitor():
        while True:
            time.sleep(check_every)
            pool.check_and_manage()

    threading.Thread(target=janitor, daemon=True).start()
    logger.info(f"bootstrap server on {host}:{port}")
    srv.serve_forever()


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--host-address", default="127.0.0.1")
    ap.add_argument("--port", type=int, default=5000)
    ap.add_argument("--store-port", type=int, default=4999)
    a = ap.parse_args()
    serve(a.host_address, a.port, a.store_port)
