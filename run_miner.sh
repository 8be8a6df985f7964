#!/usr/bin/env bash
# Launch + supervise a miner (reference run_miner.sh: pm2 launcher with version-poll auto-update).
#   ./run_miner.sh [--gpus N] [--no-autoupdate] -- <neuron flags...>
# One process per GPU through torchrun when --gpus > 1; crashes are restarted (max 5 within 5 min of uptime); a version bump
# of template/__init__.py on the git remote triggers pull + restart (utils/auto_update.py).
set -euo pipefail
cd "$(dirname "$0")"
GPUS=1; AUTOUPDATE=1; PORT=${MASTER_PORT:-29555}
while [[ $# -gt 0 ]]; do
  case "$1" in
    --gpus) GPUS="$2"; shift 2;;
    --no-autoupdate) AUTOUPDATE=0; shift;;
    --) shift; break;;
    *) break;;
  esac
done
if [[ "$GPUS" -gt 1 ]]; then
  CMD=(python -m torch.distributed.run --nnodes=1 --nproc-per-node "$GPUS" --master-addr 127.0.0.1 --master-port "$PORT" neurons/miner.py "$@")
else
  CMD=(python neurons/miner.py "$@")
fi
if [[ "$AUTOUPDATE" -eq 1 ]]; then
  python - <<'PY' &
import os
from distributedtraining_b200.utils.auto_update import monitor_repo
try:
    monitor_repo(os.getcwd(), interval=1800.0, on_update=lambda old, new: os.kill(os.getppid(), 15))
except Exception:
    pass
PY
  UPDATER=$!
  trap 'kill $UPDATER 2>/dev/null || true' EXIT
fi
exec python -m distributedtraining_b200.utils.supervisor --max-restarts 5 --min-uptime 300 -- "${CMD[@]}"
