#!/bin/bash
# ncu --set full captures (one launch each) of the hot kernels + a launch list of one training step.  Run under gpurun (1 GPU).
cd "$(dirname "$0")/.."
TAG=$1; O=gpurun_out; mkdir -p $O
NCU="ncu --clock-control none"
# launch list of one eager training step (cold-cache, serialised: shares only)
FT_GRAPH=0 FT_BWD_COOP=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 3200 --csv --log-file $O/${TAG}_launches.csv python bench.py --profile --steps 1 --warmup 1 > $O/${TAG}_launches.log 2>&1
echo "launch list rc=$?"
cap() { # name kernel-regex skip env... -- cmd
  local name=$1 regex=$2 skip=$3; shift 3
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 400 $NCU --set full -k "regex:$regex" -s $skip -c 1 -f -o $O/${TAG}_ncu_$name "$@" > $O/${TAG}_ncu_$name.log 2>&1
  echo "ncu $name rc=$?"
  ncu -i $O/${TAG}_ncu_$name.ncu-rep --page raw --csv > $O/${TAG}_ncu_$name.csv 2>/dev/null
  ncu -i $O/${TAG}_ncu_$name.ncu-rep --page details > $O/${TAG}_ncu_$name.txt 2>/dev/null
  rm -f $O/${TAG}_ncu_$name.ncu-rep          # keep the merged output small: the csv / details pages carry what is cited
}
cap lstm_fwd "lstm_fwd_kernel" 1 FT_GRAPH=0 FT_PIPE_FWD=0 FT_PIPE_BWD=0 -- python bench.py --profile --steps 1 --warmup 1
cap lstm_bwd4 "lstm_bwd4_kernel" 1 FT_GRAPH=0 FT_PIPE_FWD=0 FT_PIPE_BWD=0 FT_BWD_COOP=0 -- python bench.py --profile --steps 1 --warmup 1
cap gemm "gemm2_kernel" 12 FT_GRAPH=0 -- python bench.py --profile --steps 1 --warmup 1
cap attn_fwd "attn_fwd_kernel" 1 FT_GRAPH=0 -- python bench.py --profile --steps 1 --warmup 1
cap attn_bwd "attn_bwd_kernel" 1 FT_GRAPH=0 -- python bench.py --profile --steps 1 --warmup 1
cap enc_bilstm_fwd "enc_bilstm_fwd_kernel" 1 FT_GRAPH=0 -- python bench.py --profile --steps 1 --warmup 1
cap enc_bilstm_bwd "enc_bilstm_bwd_kernel" 1 FT_GRAPH=0 -- python bench.py --profile --steps 1 --warmup 1
cap mel_fused "mel_fused_kernel" 1 X=1 -- python bench.py --workload mel --utterances 2000 --profile --steps 1 --warmup 1
cap infer "infer_kernel" 1 X=1 -- python bench.py --workload infer --frames 100 --profile --steps 1 --warmup 1
ls -la $O/${TAG}_ncu_* | head -30
