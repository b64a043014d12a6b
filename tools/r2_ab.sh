#!/bin/bash
# A/B runner: r2_ab.sh "ENV=.. ENV=.." ["ENV=.."...]  -> bench --kernels per config (twice), kernel checks first if CHECKS is set
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
T0=$(date +%s)
say() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
if [ -n "$CHECKS" ]; then
  say "kernel checks: $CHECKS"
  for c in $CHECKS; do timeout 120 python tools/gpu_check.py $c 2>&1 | grep -E "CHECK|Error" | cut -c1-300; done
fi
if [ -n "$PYTEST" ]; then
  say "pytest -m gpu"
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/ab_pytest.log 2>&1
  say "pytest rc=$? $(tail -1 gpurun_out/ab_pytest.log)"
  grep -E "^(FAILED|ERROR)|Error" gpurun_out/ab_pytest.log | head -5
fi
for rep in 1 2; do
for cfg in "$@"; do
  env $cfg timeout 300 python bench.py --steps 20 --warmup 3 --kernels --no-cpu-baseline --no-extra-configs --stall-limit 60 > gpurun_out/ab_tmp.log 2>gpurun_out/ab_tmp.err
  python - "$cfg" gpurun_out/ab_tmp.log <<'PY'
import json, sys
cfg, path = sys.argv[1:3]
line = [l for l in open(path).read().splitlines() if l.startswith("{")]
if not line:
    print(cfg, "NO JSON LINE; tail:", open(path.replace(".log", ".err")).read()[-800:])
else:
    d = json.loads(line[-1])
    if "error" in d:
        print(cfg, d)
    else:
        k = d.get("kernels", {})
        print(cfg, "value", round(d.get("value", 0), 1), "e2e", round((d.get("e2e") or {}).get("value", 0), 1), (d.get("clocks") or {}).get("sm_mhz"),
              {n: round(v["ms_per_step"], 3) for n, v in k.items()})
PY
done
done
say done
