#!/bin/bash
cd $GRAFT_REPO_ROOT
OUT=gpurun_out/r3; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_e2e_reference_apps.py -m gpu -x -q -k "not app_with" 2>&1 | tail -8 > $OUT/pytest.log; cat $OUT/pytest.log
bash tools/variant_bench.sh build/variants/v1.so vvc-extension-mm_b200/libmmgpu.so build/variants/w4c6.so build/variants/w4c8.so build/variants/w2c14.so build/variants/w8c3.so build/variants/w1c28.so 2>&1 | tee $OUT/variants.txt
