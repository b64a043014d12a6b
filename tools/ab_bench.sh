#!/bin/bash
# A/B the whole step under different environment overrides: tools/ab_bench.sh "VAR=1" "VAR=0" ...
# (each config runs bench.py without the CPU baseline leg; prints value and the per-kind kernel times)
mkdir -p gpurun_out
i=0
for cfg in "$@"; do
  i=$((i+1))
  start=$(date +%s)
  env $cfg timeout 150 python bench.py --steps 10 --warmup 3 --kernels --no-cpu-baseline > gpurun_out/ab_$i.log 2>&1
  rc=$?
  python - "$cfg" "$rc" "$(( $(date +%s) - start ))" gpurun_out/ab_$i.log <<'PY'
import json, sys
cfg, rc, secs, path = sys.argv[1:5]
line = [l for l in open(path).read().splitlines() if l.startswith("{")]
if not line:
    print(cfg, "rc", rc, secs, "s: NO JSON LINE; tail:", open(path).read()[-400:])
else:
    d = json.loads(line[-1])
    k = d.get("kernels", {})
    print(cfg, "rc", rc, secs, "s value", round(d["value"], 1), {n: round(v["ms_per_step"], 3) for n, v in k.items() if n in ("gemm_qkv", "mlp_fused", "attention")})
PY
done
