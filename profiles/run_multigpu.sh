#!/bin/bash
# Multi-GPU measurements of this repo (run under `gpurun --gpus N`):
#   bash profiles/run_multigpu.sh <N> <tag> [config] [steps]
# 1. the N-GPU parity test when N == 2 (bit-exact vs single-GPU accumulation, all exchange modes);
# 2. bench.py at N GPUs (default exchange = fused peer-memory kernel + flag barriers), with the NVLink data counters of
#    every GPU read before and after (`nvidia-smi nvlink -gt d`) -> bytes moved over NVLink per optimizer step;
# 3. the same bench with the NCCL all-reduce baseline and with the fused kernel + NCCL brackets (A/B).
set -u
N=${1:-2}; TAG=${2:-mgpu}; CFG=${3:-2}; STEPS=${4:-60}
OUT=gpurun_out; mkdir -p $OUT
PORT=29811
run_bench () {   # $1 = exchange mode, $2 = suffix
  nvidia-smi nvlink -gt d > $OUT/${TAG}_nvlink_before_$2.txt 2>&1
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus $N --steps $STEPS --warmup 10 --config $CFG --exchange $1 --no-cpu-baseline \
      > $OUT/${TAG}_bench_n${N}_$2.json 2> $OUT/${TAG}_bench_n${N}_$2.err
  nvidia-smi nvlink -gt d > $OUT/${TAG}_nvlink_after_$2.txt 2>&1
  PORT=$((PORT+1))
  tail -c 600 $OUT/${TAG}_bench_n${N}_$2.json; echo
}
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q > $OUT/${TAG}_mgpu_test.log 2>&1; tail -5 $OUT/${TAG}_mgpu_test.log
fi
run_bench fused_p2p fused
run_bench allreduce allreduce
run_bench fused_p2p_nccl fusednccl
python profiles/nvlink_delta.py $OUT/${TAG}_nvlink_before_fused.txt $OUT/${TAG}_nvlink_after_fused.txt $((2*STEPS+12)) > $OUT/${TAG}_nvlink_fused.md 2>&1
cat $OUT/${TAG}_nvlink_fused.md
