#!/bin/bash
# Round-2 profile capture of the shipped configuration: full bench line, ncu launch list, ncu --set full of the
# attention kernel (with source) and of one whole step.  Outputs under gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
T0=$(date +%s)
say() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
say "bench (default flags + --kernels)"
timeout 900 python bench.py --kernels > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
say "rc=$?"
say "bench --impl reference"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null
say "rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-configs"
say "plain run of the profiled command"
$CMD > gpurun_out/r02_plain.log 2>&1 || { say "plain run failed"; tail -5 gpurun_out/r02_plain.log; exit 1; }
say "ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu1.log 2>&1
say "rc=$?"
say "ncu --set full: attention kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 4 -c 1 -o gpurun_out/r02_attn $CMD > gpurun_out/r02_ncu2.log 2>&1
say "rc=$?"
say "ncu --set full: one whole step"
NK=${STEP_KERNELS:-23}
timeout 1200 ncu --set full --clock-control none -k regex:'gemm_bf16|mlp_fused|layernorm|head_|im2col|attn_fwd|cls_row' -s $((3 * NK)) -c $NK -o gpurun_out/r02_step $CMD > gpurun_out/r02_ncu3.log 2>&1
say "rc=$?"
ls -la gpurun_out/r02_*.ncu-rep
say done
