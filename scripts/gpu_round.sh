#!/bin/bash
# One GPU-box visit: parity tests, bench line, per-layer breakdown, ncu launch list and full-set capture.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_round.sh <tag>
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $OUT/pytest_gpu.txt
( timeout 600 python bench.py 2> $OUT/bench.err | tail -1 ) > $OUT/bench.json
( timeout 300 python scripts/step_breakdown.py 2>&1 ) > $OUT/breakdown.txt
if [ -z "$NO_NCU" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c ${NCU_LAUNCHES:-600} --csv --log-file $OUT/launches.csv \
    python scripts/profile_step.py --inner-steps ${NCU_INNER_STEPS:-1} > $OUT/ncu_launches.log 2>&1
fi
if [ -n "$NCU_FULL" ]; then
# full-set capture of the kernels named in $NCU_FULL (regex), 2 launches each; the .ncu-rep stays on the box
# (gpurun_out/ is capped at 64 MiB) -- only the raw CSV page comes back
timeout 900 ncu --set full --clock-control none -k "regex:$NCU_FULL" -c ${NCU_COUNT:-16} -o /tmp/full \
    python scripts/profile_step.py --inner-steps 1 > $OUT/ncu_full.log 2>&1
ncu -i /tmp/full.ncu-rep --page raw --csv > $OUT/full_raw.csv 2>/dev/null
fi
ls -la $OUT
echo done
