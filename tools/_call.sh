O=gpurun_out; mkdir -p $O
for cfg in "single 4" "pair 19"; do set -- $cfg
timeout 300 ncu --clock-control none --set full -k regex:gemm2_kernel -s $2 -c 1 -f -o $O/r2c26_ncu_gemm_$1 python tools/gemm_shapes.py "fwd xproj0" > $O/r2c26_ncu_gemm_$1.log 2>&1; echo "ncu $1 rc=$?"
ncu -i $O/r2c26_ncu_gemm_$1.ncu-rep --page details > $O/r2c26_ncu_gemm_$1.txt 2>/dev/null
rm -f $O/r2c26_ncu_gemm_$1.ncu-rep
grep -E "^\s+(Duration|DRAM Throughput|L2 Cache Throughput|Compute \(SM\) Throughput|Executed Ipc Active|Registers Per Thread)" $O/r2c26_ncu_gemm_$1.txt | head -8
grep -E "gemm2_kernel<" $O/r2c26_ncu_gemm_$1.txt | head -1
done
bash tools/gpu_call.sh r2final driver infer mel
