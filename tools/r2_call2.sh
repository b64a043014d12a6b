#!/bin/bash
# Round-2 GPU call 2: CTA-pair soak incl. the concurrent-stream host path (tail build, then the round-1 control),
# in-step A/B of the attention polynomial share, ncu launch list + full captures.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
T0=$(date +%s)
say() { echo "[$(( $(date +%s) - T0 )) s] $*"; }

say "pair soak (tail build): 4000 forwards + 1500 host calls on 3 concurrent streams"
timeout 400 python tools/pair_soak.py --steps 4000 --pair 1 --host-calls 1500 > gpurun_out/r2c2_soak_tail.log 2>&1
say "rc=$? $(tail -1 gpurun_out/r2c2_soak_tail.log)"
say "pair soak (round-1 control, no tail): same"
timeout 400 python tools/pair_soak.py --lib tools/ubench/libdinoseg_r1_notail.so --steps 4000 --pair 1 --host-calls 1500 > gpurun_out/r2c2_soak_notail.log 2>&1
say "rc=$? $(tail -1 gpurun_out/r2c2_soak_notail.log)"

say "in-step A/B of the polynomial share (pair kernels on)"
for lib in dino_b200/lib/libdinoseg.so tools/ubench/libdinoseg_pm2.so tools/ubench/libdinoseg_pm0.so dino_b200/lib/libdinoseg.so tools/ubench/libdinoseg_pm2.so; do
  DINOSEG_LIB=$lib DINOSEG_PAIR=1 timeout 300 python bench.py --steps 20 --warmup 3 --kernels --no-cpu-baseline --no-extra-configs > gpurun_out/r2c2_bench_tmp.log 2>&1
  python - "$lib" gpurun_out/r2c2_bench_tmp.log <<'PY'
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

CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-configs"
say "plain run of the profiled command"
DINOSEG_PAIR=1 $CMD > gpurun_out/r2c2_plain.log 2>&1 || { say "plain run failed"; tail -5 gpurun_out/r2c2_plain.log; exit 1; }
say "ncu launch list"
DINOSEG_PAIR=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c2_launches.csv $CMD > gpurun_out/r2c2_ncu1.log 2>&1
say "rc=$?"
say "ncu --set full: attention kernel"
DINOSEG_PAIR=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 4 -c 1 -o gpurun_out/r2c2_attn $CMD > gpurun_out/r2c2_ncu2.log 2>&1
say "rc=$?"
say "ncu --set full: one whole step"
DINOSEG_PAIR=1 timeout 900 ncu --set full --clock-control none -k regex:'gemm_bf16|mlp_fused|layernorm|head_tail|im2col|attn_fwd|cls_row' -s 75 -c 25 -o gpurun_out/r2c2_step $CMD > gpurun_out/r2c2_ncu3.log 2>&1
say "rc=$?"
ls -la gpurun_out/*.ncu-rep
say done
