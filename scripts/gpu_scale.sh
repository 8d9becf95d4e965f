#!/bin/bash
# Strong-scaling bench lines at N = 1, 2, 4, 8 for the given configs (under gpurun --gpus 8).
# Usage: bash scripts/gpu_scale.sh tag "2 4"
TAG=${1:-scale}; CFGS=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
for CFG in $CFGS; do
  for n in 1 2 4 8; do
    if [ $n -eq 1 ]; then
      ( timeout 300 python bench.py --config $CFG --no-cpu-baseline --no-kernel-breakdown 2> $OUT/err_cfg${CFG}_n$n.txt | tail -1 ) > $OUT/bench_cfg${CFG}_n$n.json
    else
      ( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
          bench.py --config $CFG --gpus $n 2> $OUT/err_cfg${CFG}_n$n.txt | tail -1 ) > $OUT/bench_cfg${CFG}_n$n.json
    fi
  done
done
ls -la $OUT
