#!/bin/bash
# One-shot GPU validation of the switchable schedulers/optimizer: GPU suite with everything on, then A/B bench lines.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export FT_ENC_STREAMS=1 FT_ENC_OVERLAP=1
timeout 170 python -m pytest tests -x -q -m gpu > gpurun_out/fv_tests.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/fv_tests.txt
run() { # name fused streams overlap
  FT_FUSED_OPT=$2 FT_ENC_STREAMS=$3 FT_ENC_OVERLAP=$4 timeout 80 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/fv_$1.json 2> gpurun_out/fv_$1.err
  python - "$1" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/fv_{n}.json").read().strip().splitlines()[-1])
    print(n, round(d["value"]), round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"], d["clocks"]["reasons"])
except Exception as e:
    print(n, "FAILED", e)
    print(open(f"gpurun_out/fv_{n}.err").read()[-1500:])
PY
}
run all 1 1 1
run base 0 0 0
run fused 1 0 0
run streams 0 1 0
run overlap 0 0 1
