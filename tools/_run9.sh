#!/bin/bash
cd $GRAFT_REPO_ROOT
OUT=gpurun_out/r9; mkdir -p $OUT
timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -25 > $OUT/pytest.log; cat $OUT/pytest.log
python tools/mc_jobs_time.py 2>&1 | tail -2 | tee $OUT/mc_jobs.txt
