#!/bin/bash
# Round-2 GPU call 3: parity suite with the new defaults (pair kernels on, 12.5 % polynomial share, barrier / TMEM
# addresses in registers), attention timing of the variants, in-step A/B.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
T0=$(date +%s)
say() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
say "pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c3_pytest.log 2>&1
say "pytest rc=$? $(tail -1 gpurun_out/r2c3_pytest.log)"
say "attention timing (B=64, N=3601, H=6), twice"
for rep in 1 2; do
  for v in dino_b200/lib/libdinoseg.so tools/ubench/libdinoseg_pm0.so tools/ubench/libdinoseg_pm1.so tools/ubench/libdinoseg_pm3.so tools/ubench/libdinoseg_pm4.so tools/ubench/libdinoseg_r112.so; do
    echo "smw=8 $v: $(DSG_TIMING_SO=$v DINOSEG_ATTN_SMW=8 timeout 120 python tools/attn_timing.py 2>&1 | tail -1)"
  done
done 2>&1 | tee gpurun_out/r2c3_attn_timing.log
echo "smw=4 default: $(DSG_TIMING_SO=dino_b200/lib/libdinoseg.so DINOSEG_ATTN_SMW=4 timeout 120 python tools/attn_timing.py 2>&1 | tail -1)" | tee -a gpurun_out/r2c3_attn_timing.log
say "ViT-B and 960 px attention shapes"
python - <<'PY' 2>&1 | tee gpurun_out/r2c3_attn_shapes.log
import ctypes as C, os, torch
for smw in ("8", "4"):
    pass
PY
say "in-step A/B"
for lib in dino_b200/lib/libdinoseg.so tools/ubench/libdinoseg_pm1.so tools/ubench/libdinoseg_pm3.so tools/ubench/libdinoseg_r112.so dino_b200/lib/libdinoseg.so; do
  DINOSEG_LIB=$lib timeout 300 python bench.py --steps 20 --warmup 3 --kernels --no-cpu-baseline --no-extra-configs > gpurun_out/r2c3_bench_tmp.log 2>&1
  python - "$lib" gpurun_out/r2c3_bench_tmp.log <<'PY'
import json, sys
cfg, path = sys.argv[1:3]
line = [l for l in open(path).read().splitlines() if l.startswith("{")]
if not line:
    print(cfg, "NO JSON LINE; tail:", open(path).read()[-600:])
else:
    d = json.loads(line[-1])
    k = d.get("kernels", {})
    print(cfg, "value", round(d.get("value", 0), 1), "e2e", round((d.get("e2e") or {}).get("value", 0), 1),
          "attn(timed region)", round(d["roofline"]["avg_launch_ms"], 4), d["clocks"],
          {n: round(v["ms_per_step"], 3) for n, v in k.items()})
PY
done
say "full bench line (default flags)"
timeout 600 python bench.py > gpurun_out/r2c3_bench_default.json 2> gpurun_out/r2c3_bench_default.err
say "rc=$?"; head -c 3000 gpurun_out/r2c3_bench_default.json
say "smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
say done
