O=gpurun_out; mkdir -p $O
export FT_PARITY_LOG=r2c12_parity.jsonl
timeout 600 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_infer.py tests/test_gpu_pipeline.py -q -m gpu -rfE --tb=line -p no:cacheprovider > $O/r2c12_tests.txt 2>&1; echo "pytest rc=$?"; tail -14 $O/r2c12_tests.txt
timeout 120 python tools/trace_infer.py 1 > $O/r2c12_trace_infer_b1.txt 2>&1; grep -v Warn $O/r2c12_trace_infer_b1.txt | head -3
timeout 120 python tools/trace_infer.py 16 > $O/r2c12_trace_infer_b16.txt 2>&1; grep -v Warn $O/r2c12_trace_infer_b16.txt
bash tools/gpu_call.sh r2c12 infer
for v in 0 128; do
FT_ATT_OVERLAP=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/r2c12_train_att$v.json 2> $O/r2c12_train_att$v.err; echo "train att=$v rc=$?"; python - <<PY
import json
d=json.loads(open("$O/r2c12_train_att$v.json").read().strip().splitlines()[-1])
print(" value",round(d["value"],1),"ms/step",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["value"],1))
PY
done
cat $O/r2c12_parity.jsonl | grep enc
