O=gpurun_out; mkdir -p $O
export FT_PARITY_LOG=r2c8_parity.jsonl
timeout 600 python -m pytest tests/test_gpu_infer.py tests/test_gpu_baseline_shapes.py -q -m gpu -x -rfE --tb=short -p no:cacheprovider -k "infer" > $O/r2c8_tests.txt 2>&1; echo "pytest rc=$?"; tail -15 $O/r2c8_tests.txt
timeout 120 python tools/trace_infer.py 1 > $O/r2c8_trace_infer_b1.txt 2>&1; grep -v Warn $O/r2c8_trace_infer_b1.txt
timeout 120 python tools/trace_infer.py 16 > $O/r2c8_trace_infer_b16.txt 2>&1; grep -v Warn $O/r2c8_trace_infer_b16.txt
bash tools/gpu_call.sh r2c8 infer
