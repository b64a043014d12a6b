#!/bin/bash
# isolated attention launch time (bench shape) for every library variant tools/ubench/libdinoseg_*.so, twice
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
if [ -n "$CHECKS" ]; then
  for c in $CHECKS; do timeout 120 python tools/gpu_check.py $c 2>&1 | grep -E "CHECK|Error" | cut -c1-300; done
fi
for rep in 1 2; do
  for so in tools/ubench/libdinoseg_*.so; do
    echo -n "$so: "; DSG_TIMING_SO=$so timeout 120 python tools/attn_timing.py 2>&1 | tail -1
  done
done
