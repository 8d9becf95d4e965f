#!/bin/bash
# GPU visit for the fused image block: its kernel tests (+ a memcheck pass on the small cases), then the usual round.
TAG=${1:-img}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 600 python -m pytest tests/test_gpu_img_block.py -x -q 2>&1 | tail -40 ) > $OUT/pytest_img.txt
( timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_img_block.py -q -k "case0 or case3 or no_base" 2>&1 | tail -40 ) > $OUT/memcheck_img.txt
NO_NCU=${NO_NCU:-} bash scripts/gpu_round.sh $TAG
