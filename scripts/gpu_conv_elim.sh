#!/bin/bash
# Elimination experiments on conv_tc (time only, results meaningless): which role's work sets the tile period?
# (the table in profiles/r02_conv_tc_elimination.txt also had -DXM_TC_NOSTATS lines: that switch's stub sent the accumulators
# to local memory and was removed with the shuffle statistics it replaced)
OUT=gpurun_out/${1:-elim}
mkdir -p $OUT
for flags in "" "-DXM_TC_NOMMA" "-DXM_TC_NOPROD" "-DXM_TC_NOSTORE" "-DXM_TC_NOMMA -DXM_TC_NOPROD" "-DXM_TC_NOPROD -DXM_TC_NOSTORE" "-DXM_TC_NOMMA -DXM_TC_NOSTORE"; do
  XM_NVCC_EXTRA="$flags" python -m exploring_meta_b200.build --force > $OUT/build.log 2>&1 || { tail -5 $OUT/build.log; continue; }
  echo "== flags: $flags" >> $OUT/elim.txt
  XM_TIMING_REPS=20 timeout 60 python scripts/conv_timing.py 42 1 2>&1 | tail -1 >> $OUT/elim.txt
done
python -m exploring_meta_b200.build --force > $OUT/build.log 2>&1
cat $OUT/elim.txt
