#!/bin/bash
# 8-GPU evidence run (under `gpurun --gpus 8`): headline config with the fused exchange + flag barriers, the same with
# NCCL brackets (A/B of the barrier change), the NCCL all-reduce baseline, and BASELINE configs[4] (4M / 24 views / 4K);
# NVLink data counters of every GPU before/after the first run.
set -u
N=${1:-8}; TAG=${2:-r02}
OUT=gpurun_out; mkdir -p $OUT
PORT=29841
bench () {  # $1 exchange  $2 config  $3 steps  $4 suffix
  timeout 260 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus $N --steps $3 --warmup 10 --config $2 --exchange $1 --no-cpu-baseline \
      > $OUT/${TAG}_bench_n${N}_$4.json 2> $OUT/${TAG}_bench_n${N}_$4.err
  PORT=$((PORT+1))
  python -c "
import json,sys
try:
    d=json.load(open('$OUT/${TAG}_bench_n${N}_$4.json')); print('$4', d['value'], d['ms_per_step'], d['config'].get('exchange'), d['kernels'].get('adam'))
except Exception as e: print('$4 FAILED', e)"
}
nvidia-smi nvlink -gt d > $OUT/${TAG}_nvlink_before.txt 2>&1
bench fused_p2p 2 100 cfg2_fused
nvidia-smi nvlink -gt d > $OUT/${TAG}_nvlink_after.txt 2>&1
python profiles/nvlink_delta.py $OUT/${TAG}_nvlink_before.txt $OUT/${TAG}_nvlink_after.txt 212 > $OUT/${TAG}_nvlink_n${N}_fused.md 2>&1
cat $OUT/${TAG}_nvlink_n${N}_fused.md
bench fused_p2p_nccl 2 100 cfg2_fusednccl
bench allreduce 2 100 cfg2_allreduce
bench fused_p2p 4 40 cfg4_fused
