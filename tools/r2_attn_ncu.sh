#!/bin/bash
# attention checks, isolated launch time, a short bench with the per-kernel breakdown, then ncu --set full (with source) of
# the unshifted attention kernel at the bench shape
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
for c in attn_1tile attn_ragged_small attn_2tiles attn_901 attn_3601 attn_vitb_901 attn_wide_901 attn_shifted_901 attn_spread8_901; do
  timeout 120 python tools/gpu_check.py $c 2>&1 | grep -E "CHECK|Error" | cut -c1-260; done
python tools/attn_timing.py | tail -1
DINOSEG_ATTN_UNSHIFTED=0 python tools/attn_timing.py | tail -1
bash tools/r2_ab.sh "A=1" 2>&1 | tail -4
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 6 -c 1 -o gpurun_out/r02_attn_u python tools/attn_timing.py > gpurun_out/r02_attn_u_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/r02_attn_u.ncu-rep
