#!/bin/bash
cd $GRAFT_REPO_ROOT
OUT=gpurun_out/r14; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dmvr or alt or bdof" 2>&1 | tail -15 > $OUT/pytest_dmvr.log; cat $OUT/pytest_dmvr.log
timeout 900 python -m pytest tests/test_e2e_reference_apps.py -m gpu -x -q -s -k "decoder_app" 2>&1 | tail -25 | cut -c1-600 > $OUT/pytest_dec.log; cat $OUT/pytest_dec.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; tail -c 3000 $OUT/bench.json; tail -3 $OUT/bench.err
MMGPU_E2E_RA_ENCODE=1 timeout 1500 python -m pytest tests/test_e2e_reference_apps.py -m gpu -q -s -k "random_access_bitstream" 2>&1 | tail -30 | cut -c1-600 > $OUT/ra_encode.log; cat $OUT/ra_encode.log
