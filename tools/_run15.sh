#!/bin/bash
cd $GRAFT_REPO_ROOT
OUT=gpurun_out/r15; mkdir -p $OUT
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; tail -c 1500 $OUT/bench.json; tail -3 $OUT/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; tail -c 600 $OUT/bench_ref.json
