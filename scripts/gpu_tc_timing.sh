#!/bin/bash
# Role-level timing of the tcgen05 conv kernel (XM_TC_TIMING printf instrumentation), built on the GPU box.
TAG=${1:-tct}
OUT=gpurun_out/$TAG
mkdir -p $OUT
XM_NVCC_EXTRA=-DXM_TC_TIMING python -m exploring_meta_b200.build --force > $OUT/build.log 2>&1
python scripts/profile_calls.py --only "xm_conv cin32 42x42 fwd,xm_conv cin32 42x42 dgrad" > $OUT/tc_timing.txt 2>&1
tail -40 $OUT/tc_timing.txt
