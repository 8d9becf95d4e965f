#!/bin/bash
# Multi-GPU visit: 2-GPU equality test, then strong-scaling bench lines at N = 1 .. $1 (default 2).
# Usage (under gpurun --gpus N):  bash scripts/gpu_multi.sh N tag [config]
N=${1:-2}; TAG=${2:-multi}; CFG=${3:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
( timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q -s 2>&1 | tail -40 ) > $OUT/pytest_multirank.txt
for n in 1 2 4 8; do
  [ $n -gt $N ] && break
  if [ $n -eq 1 ]; then
    ( timeout 600 python bench.py --config $CFG --no-cpu-baseline 2> $OUT/bench_n$n.err | tail -1 ) > $OUT/bench_cfg${CFG}_n$n.json
  else
    ( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --config $CFG --gpus $n 2> $OUT/bench_n$n.err | tail -1 ) > $OUT/bench_cfg${CFG}_n$n.json
  fi
  tail -5 $OUT/bench_n$n.err
done
ls -la $OUT
echo done
