#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
T0=$(date +%s)
say() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
say "kernel checks: fused-LN MLP"
for c in mlp_ln_ragged mlp_ln_multi mlp_ln_pair_ragged mlp_ln_pair_multi; do timeout 120 python tools/gpu_check.py $c 2>&1 | grep -E "CHECK|Error" | cut -c1-400; done
say "pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c4_pytest.log 2>&1
say "pytest rc=$? $(tail -1 gpurun_out/c4_pytest.log)"
grep -E "^(FAILED|ERROR)|Error" gpurun_out/c4_pytest.log | head -5
say "A/B fused LN"
for cfg in "DINOSEG_FUSE_LN=1" "DINOSEG_FUSE_LN=0" "DINOSEG_FUSE_LN=1" "DINOSEG_FUSE_LN=0"; do
  env $cfg timeout 300 python bench.py --steps 20 --warmup 3 --kernels --no-cpu-baseline --no-extra-configs > gpurun_out/c4_bench_tmp.log 2>&1
  python - "$cfg" gpurun_out/c4_bench_tmp.log <<'PY'
import json, sys
cfg, path = sys.argv[1:3]
line = [l for l in open(path).read().splitlines() if l.startswith("{")]
if not line:
    print(cfg, "NO JSON LINE; tail:", open(path).read()[-600:])
else:
    d = json.loads(line[-1])
    k = d.get("kernels", {})
    print(cfg, "value", round(d.get("value", 0), 1), "e2e", round((d.get("e2e") or {}).get("value", 0), 1), d["clocks"].get("sm_mhz"),
          {n: round(v["ms_per_step"], 3) for n, v in k.items()})
PY
done
say "soak with the fused LN (pair kernels)"
timeout 300 python tools/pair_soak.py --steps 3000 --pair 1 --host-calls 300 2>&1 | tail -1
say done
