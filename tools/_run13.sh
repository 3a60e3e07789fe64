#!/bin/bash
cd $GRAFT_REPO_ROOT
OUT=gpurun_out/r13; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $OUT/pytest.log; cat $OUT/pytest.log
timeout 600 bash tools/variant_bench.sh vvc-extension-mm_b200/libmmgpu.so build/variants/w2c18.so build/variants/w4c9.so build/variants/w3c12.so build/variants/w2c20.so build/variants/w4c10.so vvc-extension-mm_b200/libmmgpu.so 2>&1 | tee $OUT/variants.txt
