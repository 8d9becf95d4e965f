#!/bin/bash
# ncu --set full of one launch of every distinct call of the config-2 program; only the CSV pages come back.
TAG=${1:-ncu}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 ncu --set full --clock-control none --profile-from-start off -o /tmp/full \
    python scripts/profile_calls.py $PROFILE_ARGS > $OUT/ncu_full.log 2>&1
ncu -i /tmp/full.ncu-rep --page raw --csv > $OUT/full_raw.csv 2>/dev/null
ls -la /tmp/full.ncu-rep $OUT
tail -3 $OUT/ncu_full.log
