#!/bin/bash
# single-compute-stream host pipeline: parity, chunk sweep, stall hunt
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
T0=$(date +%s)
say() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
say "pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c5_pytest.log 2>&1
say "pytest rc=$? $(tail -1 gpurun_out/c5_pytest.log)"
grep -E "^(FAILED|ERROR)|Error" gpurun_out/c5_pytest.log | head -5
say "chunk sweep (e2e)"
for ch in 0 8 10 13 16 21 32; do
  DINOSEG_HOST_CHUNK=$ch timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra-configs --stall-limit 60 > gpurun_out/c5_tmp.log 2>gpurun_out/c5_tmp.err
  python - "$ch" gpurun_out/c5_tmp.log <<'PY'
import json, sys
cfg, path = sys.argv[1:3]
line = [l for l in open(path).read().splitlines() if l.startswith("{")]
d = json.loads(line[-1]) if line else {}
print("chunk", cfg, "value", round(d.get("value", 0), 1), "e2e", round((d.get("e2e") or {}).get("value", 0), 1), "e2e_u8", round((d.get("e2e_u8") or {}).get("value", 0), 1), d.get("error"))
PY
done
say "stall hunt, pair kernels on, 40 runs"
bash tools/r2_stall_hunt.sh 40 "DINOSEG_PAIR=1"
say done
