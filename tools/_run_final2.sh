#!/bin/bash
cd $GRAFT_REPO_ROOT
OUT=gpurun_out/final2; mkdir -p $OUT
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; tail -c 600 $OUT/bench.json; tail -2 $OUT/bench.err
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dmvr or unpadded or bdof or non_power" > $OUT/sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -6 $OUT/sanitizer.log | cut -c1-300
