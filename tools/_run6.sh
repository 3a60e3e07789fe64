#!/bin/bash
cd $GRAFT_REPO_ROOT
OUT=gpurun_out/r6; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_reference or tan_double or oob" 2>&1 | tail -5 > $OUT/pytest.log; cat $OUT/pytest.log
python tools/multiref_time.py 4 2>&1 | tee $OUT/multiref.txt
python tools/mc_jobs_time.py 2>&1 | tee $OUT/mc_jobs.txt
python tools/mc_batch_time.py 2>&1 | tee -a $OUT/mc_jobs.txt
python tools/sweep.py $OUT/sweep_r02.csv > $OUT/sweep.log 2>&1; tail -4 $OUT/sweep.log
