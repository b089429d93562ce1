O=gpurun_out; mkdir -p $O
bash tools/gpu_call.sh r2c13 tests infer
for v in 64 32; do
FT_MAX_KERNEL_BATCH=$v timeout 400 python bench.py --config 3 --steps 4 --warmup 3 --no-cpu-baseline > $O/r2c13_cfg3_mkb$v.json 2> $O/r2c13_cfg3_mkb$v.err; echo "cfg3 mkb=$v rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("$O/r2c13_cfg3_mkb$v.json").read().strip().splitlines()[-1])
    print(" value",round(d["value"],1),"ms/step",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["value"],1))
except Exception as e:
    print("FAILED", e); print(open("$O/r2c13_cfg3_mkb$v.err").read()[-1500:])
PY
done
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/r2c13_train.json 2> $O/r2c13_train.err; echo "train rc=$?"; python - <<PY
import json
d=json.loads(open("$O/r2c13_train.json").read().strip().splitlines()[-1])
print(" value",round(d["value"],1),"ms/step",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["value"],1), d.get("gpu_launches"))
print(" kernels:", {n: round(v["ms_per_step"], 2) for n, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"])[:10]})
PY
