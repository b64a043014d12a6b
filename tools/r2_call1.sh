#!/bin/bash
# Round-2 GPU call 1: parity suite, attention A/B (4 vs 8 softmax warps per query tile, polynomial share), CTA-pair
# soak (with the producer tail, then the round-1 build without it as the control), whole-step A/B.
# Everything is bounded by `timeout`; logs go to gpurun_out/.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
T0=$(date +%s)
say() { echo "[$(( $(date +%s) - T0 )) s] $*"; }

say "pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c1_pytest.log 2>&1
say "pytest rc=$? $(tail -1 gpurun_out/r2c1_pytest.log)"

say "attention checks, SMW=8 and SMW=4"
for smw in 8 4; do
  for c in attn_1tile attn_ragged_small attn_2tiles attn_901 attn_3601 attn_vitb_901; do
    DINOSEG_ATTN_SMW=$smw timeout 120 python tools/gpu_check.py $c 2>&1 | grep CHECK | sed "s/^/smw=$smw /"
  done
done > gpurun_out/r2c1_attn_checks.log 2>&1
grep -c '"ok": true' gpurun_out/r2c1_attn_checks.log; grep '"ok": false' gpurun_out/r2c1_attn_checks.log | head -5

say "attention timing (B=64, N=3601, H=6)"
{
  for smw in 4 8; do
    echo "smw=$smw mask=0x8888: $(DSG_TIMING_SO=dino_b200/lib/libdinoseg.so DINOSEG_ATTN_SMW=$smw timeout 120 python tools/attn_timing.py 2>&1 | tail -1)"
  done
  for v in pm0 pm2 pm5 pm6 pm8; do
    for smw in 8 4; do
      echo "smw=$smw $v: $(DSG_TIMING_SO=tools/ubench/libdinoseg_$v.so DINOSEG_ATTN_SMW=$smw timeout 120 python tools/attn_timing.py 2>&1 | tail -1)"
    done
  done
  echo "smw=8 mask=0x8888 again: $(DSG_TIMING_SO=dino_b200/lib/libdinoseg.so DINOSEG_ATTN_SMW=8 timeout 120 python tools/attn_timing.py 2>&1 | tail -1)"
} > gpurun_out/r2c1_attn_timing.log 2>&1
cat gpurun_out/r2c1_attn_timing.log

say "whole-step A/B (bench.py --steps 10 --kernels)"
for cfg in "DINOSEG_ATTN_SMW=8" "DINOSEG_ATTN_SMW=4" "DINOSEG_ATTN_SMW=8 DINOSEG_PAIR=1"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  env $cfg timeout 300 python bench.py --steps 10 --warmup 3 --kernels --no-cpu-baseline --no-extra-configs > gpurun_out/r2c1_bench_$tag.log 2>&1
  python - "$cfg" gpurun_out/r2c1_bench_$tag.log <<'PY'
import json, sys
cfg, path = sys.argv[1:3]
line = [l for l in open(path).read().splitlines() if l.startswith("{")]
if not line:
    print(cfg, "NO JSON LINE; tail:", open(path).read()[-600:])
else:
    d = json.loads(line[-1])
    k = d.get("kernels", {})
    print(cfg, "value", round(d.get("value", 0), 1), "e2e", round((d.get("e2e") or {}).get("value", 0), 1),
          {n: round(v["ms_per_step"], 3) for n, v in k.items()})
PY
done

say "pair soak: this build (producer tail), 6000 forwards"
timeout 300 python tools/pair_soak.py --steps 6000 --pair 1 > gpurun_out/r2c1_soak_tail.log 2>&1
say "rc=$? $(tail -1 gpurun_out/r2c1_soak_tail.log)"
say "pair soak, ViT-B (pair fc1 / fc2 GEMMs), 2000 forwards"
timeout 300 python tools/pair_soak.py --steps 2000 --pair 1 --arch vit_base --n-blocks 4 > gpurun_out/r2c1_soak_tail_vitb.log 2>&1
say "rc=$? $(tail -1 gpurun_out/r2c1_soak_tail_vitb.log)"
say "pair soak: control = round-1 build WITHOUT the producer tail, 6000 forwards"
timeout 300 python tools/pair_soak.py --lib tools/ubench/libdinoseg_r1_notail.so --steps 6000 --pair 1 > gpurun_out/r2c1_soak_notail.log 2>&1
say "rc=$? $(tail -1 gpurun_out/r2c1_soak_notail.log)"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv,noheader
say done
