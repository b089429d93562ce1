bash tools/gpu_call.sh r2c7 tests infer
timeout 120 python tools/trace_infer.py 1 > gpurun_out/r2c7_trace_infer_b1.txt 2>&1; cat gpurun_out/r2c7_trace_infer_b1.txt | grep -v Warn
timeout 120 python tools/trace_infer.py 16 > gpurun_out/r2c7_trace_infer_b16.txt 2>&1; cat gpurun_out/r2c7_trace_infer_b16.txt | grep -v Warn
