#!/bin/bash
# One gpurun call of round 2: GPU test suite (all tests, errors recorded), then bench lines of the configurations under study.
# usage: tools/gpu_call.sh <tag> [sections...]   sections: tests train infer mel timeline ref
cd "$(dirname "$0")/.."
TAG=$1; shift
SECS="${@:-tests train infer mel ref}"
O=gpurun_out; mkdir -p $O
export FT_PARITY_LOG=${TAG}_parity.jsonl
rm -f $O/$FT_PARITY_LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,temperature.gpu --format=csv,noheader > $O/${TAG}_gpu.txt 2>&1
bench() { # name, env..., -- args
  local name=$1; shift
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python bench.py "$@" > $O/${TAG}_$name.json 2> $O/${TAG}_$name.err
  echo "== $name rc=$?"; python - "$O/${TAG}_$name.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print("  value", round(d.get("value") or 0, 1), d.get("unit"), "| ms/step", round(d.get("ms_per_step") or 0, 3), "| e2e", round((d.get("e2e") or {}).get("value") or 0, 1),
          "| graph", (d.get("config") or {}).get("cuda_graph"), (d.get("config") or {}).get("cuda_graph_note"), "| roof", r.get("kernel"), r.get("frac"), r.get("us_per_recurrent_step"),
          "| clocks", (d.get("clocks") or {}).get("sm_mhz"), (d.get("clocks") or {}).get("reasons"))
    k = d.get("kernels") or {}
    print("  kernels:", {n: round(v["ms_per_step"], 2) for n, v in sorted(k.items(), key=lambda kv: -kv[1]["ms_per_step"])[:8]})
    if d.get("cpu_baseline"): print("  cpu:", d["cpu_baseline"].get("value"), d["cpu_baseline"].get("cores"), d["cpu_baseline"].get("thread_sweep_s"))
except Exception as e:
    print("  FAILED", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-1200:])
PY
}
for s in $SECS; do case $s in
tests)
  timeout 900 python -m pytest tests -q -m gpu -rfE --tb=short -p no:cacheprovider > $O/${TAG}_tests.txt 2>&1; echo "pytest rc=$?"; tail -25 $O/${TAG}_tests.txt ;;
train)
  bench train_default X=1 -- --steps 5 --warmup 3 --no-cpu-baseline
  bench train_noxin FT_LSTM_XIN=0 -- --steps 5 --warmup 3 --no-cpu-baseline
  ;;
trace)
  timeout 120 python tools/trace_lstm.py 32 > $O/${TAG}_trace_lstm.txt 2>&1; echo "trace rc=$?"; cat $O/${TAG}_trace_lstm.txt
  ;;
driver)
  bench bench_default X=1 -- 
  bench bench_reference X=1 -- --impl reference ;;
cfg3)
  bench train_cfg3 X=1 -- --config 3 --steps 5 --warmup 3 --no-cpu-baseline ;;
tlgraph)
  timeout 300 python tools/timeline_graph.py > $O/${TAG}_timeline_graph.txt 2>&1; echo "timeline_graph rc=$?"; grep -v Warn $O/${TAG}_timeline_graph.txt | head -60 ;;
infer)
  bench infer_b1 X=1 -- --workload infer --batch 1 --steps 3 --warmup 3 --no-cpu-baseline
  bench infer_b16 X=1 -- --workload infer --batch 16 --steps 3 --warmup 3 --no-cpu-baseline ;;
mel)
  bench mel_f32 X=1 -- --workload mel --steps 5 --warmup 3 --no-cpu-baseline
  bench mel_s16 X=1 -- --workload mel --wav-int16 --steps 5 --warmup 3 --no-cpu-baseline ;;
timeline)
  FT_FUSED_OPT=1 timeout 200 python tools/timeline.py > $O/${TAG}_timeline_fused.txt 2>&1; echo "timeline rc=$?"; head -12 $O/${TAG}_timeline_fused.txt; tail -3 $O/${TAG}_timeline_fused.txt ;;
ref)
  bench ref_train X=1 -- --impl reference --steps 1 --warmup 0
  bench ref_infer X=1 -- --impl reference --workload infer --steps 1 --warmup 0
  bench ref_mel X=1 -- --impl reference --workload mel --steps 1 ;;
esac; done
echo "parity records:"; cat $O/$FT_PARITY_LOG 2>/dev/null | cut -c1-400
