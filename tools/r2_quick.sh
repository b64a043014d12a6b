#!/bin/bash
# quick validation: GPU parity suite + default bench line + smoke
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
T0=$(date +%s)
say() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
say "pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/quick_pytest.log 2>&1
say "pytest rc=$? $(tail -1 gpurun_out/quick_pytest.log)"
grep -E "^(FAILED|ERROR)|Error" gpurun_out/quick_pytest.log | head -5
say "bench (default flags ${BENCH_FLAGS})"
timeout 600 python bench.py ${BENCH_FLAGS} > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err
say "rc=$?"
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/quick_bench.json") if l.startswith("{")][-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")})
    print("e2e", d["e2e"]["value"], "e2e_u8", (d.get("e2e_u8") or {}).get("value"))
    print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "frac_of_burst", "avg_launch_ms", "share_of_step", "traffic")})
    print("kernels", {k: round(v["ms_per_step"], 3) for k, v in (d.get("kernels") or {}).items()})
    print("extra", json.dumps(d.get("extra_configs"))[:1500])
    print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("no bench line:", e); print(open("gpurun_out/quick_bench.err").read()[-1500:])
PY
say "smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
say done
