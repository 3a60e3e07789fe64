#!/bin/bash
cd $GRAFT_REPO_ROOT
OUT=gpurun_out/r17; mkdir -p $OUT
timeout 300 python -m pytest tests/test_e2e_reference_apps.py -m gpu -x -q -s -k "decoder_app" 2>&1 | tail -25 | cut -c1-700 > $OUT/pytest_dec.log; cat $OUT/pytest_dec.log
MMGPU_E2E_VERIFY_ENCODE=1 timeout 1500 python -m pytest tests/test_e2e_reference_apps.py -m gpu -q -s -k "verify_mode_every" 2>&1 | tail -30 | cut -c1-700 > $OUT/verify_encode.log; cat $OUT/verify_encode.log
