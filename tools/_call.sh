O=gpurun_out; mkdir -p $O
export FT_PARITY_LOG=r2c21_parity.jsonl
timeout 300 python -m pytest tests/test_gpu_gemm.py -q -m gpu -x -rfE --tb=short -p no:cacheprovider > $O/r2c21_gemm_tests.txt 2>&1; echo "gemm pytest rc=$?"; tail -6 $O/r2c21_gemm_tests.txt
timeout 200 python tools/gemm_shapes.py 2>&1 | grep -v Warn | tee $O/r2c21_gemm_shapes.txt
bash tools/gpu_call.sh r2c21 tests | tail -8
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/r2c21_train.json 2> $O/r2c21_train.err; echo "train rc=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/r2c21_train.json").read().strip().splitlines()[-1])
print(" value",round(d["value"],1),"ms/step",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["value"],1), d.get("gpu_launches"))
print(" kernels:", {n: round(v["ms_per_step"], 2) for n, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"])[:10]})
PY
