O=gpurun_out; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > $O/r2c25_bench_n2.json 2> $O/r2c25_bench_n2.err; echo "n2 rc=$?"; tail -c 1800 $O/r2c25_bench_n2.json; tail -5 $O/r2c25_bench_n2.err
