O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_mel.py -q -m gpu -x -rfE --tb=short -p no:cacheprovider > $O/r2c16_gemm_tests.txt 2>&1; echo "gemm+mel pytest rc=$?"; tail -5 $O/r2c16_gemm_tests.txt
timeout 200 python tools/gemm_shapes.py > $O/r2c16_gemm_shapes.txt 2>&1; cat $O/r2c16_gemm_shapes.txt | grep -v Warn
for cfg in "single 4" "pair 19"; do set -- $cfg
timeout 300 ncu --clock-control none --set full -k regex:gemm2_kernel -s $2 -c 1 -f -o $O/r2c16_ncu_gemm_$1 python tools/gemm_shapes.py "fwd xproj0" > $O/r2c16_ncu_gemm_$1.log 2>&1; echo "ncu $1 rc=$?"
ncu -i $O/r2c16_ncu_gemm_$1.ncu-rep --page details > $O/r2c16_ncu_gemm_$1.txt 2>/dev/null
ncu -i $O/r2c16_ncu_gemm_$1.ncu-rep --page raw --csv > $O/r2c16_ncu_gemm_$1.csv 2>/dev/null
rm -f $O/r2c16_ncu_gemm_$1.ncu-rep
grep -E "^\s+(Duration|DRAM Throughput|L2 Cache Throughput|Compute \(SM\) Throughput|Executed Ipc Active|Registers Per Thread|Block Limit)" $O/r2c16_ncu_gemm_$1.txt | head -12
grep -iE "stall|warp cycles per issued|tensor" $O/r2c16_ncu_gemm_$1.txt | head -12
done
FT_MEL_OCC=2 timeout 200 python bench.py --workload mel --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mel occ2', round(d['value']/1e6,1), 'M frames/s')"
FT_MEL_OCC=3 timeout 200 python bench.py --workload mel --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mel occ3', round(d['value']/1e6,1), 'M frames/s')"
