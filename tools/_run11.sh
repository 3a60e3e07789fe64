#!/bin/bash
cd $GRAFT_REPO_ROOT
OUT=gpurun_out/r11; mkdir -p $OUT
MMGPU_E2E_RA_ENCODE=1 timeout 2400 python -m pytest tests/test_e2e_reference_apps.py -m gpu -q -s -k "random_access_bitstream or non_power" 2>&1 | tail -30 > $OUT/ra_encode.log; cat $OUT/ra_encode.log | cut -c1-400
