#!/bin/bash
cd $GRAFT_REPO_ROOT
OUT=gpurun_out/r4; mkdir -p $OUT
bash tools/variant_bench.sh build/variants/notail.so build/variants/oneshot.so build/variants/oneshot_w1c28.so 2>&1 | tee $OUT/variants.txt
MMGPU_LIB=$PWD/build/variants/v1.so python tools/step_time.py > $OUT/plain_v1.log 2>&1 && MMGPU_LIB=$PWD/build/variants/v1.so ncu --set full --clock-control none --import-source on -k regex:k_fir -s 20 -c 1 -o $OUT/prof_v1 python tools/step_time.py > $OUT/ncu_v1.log 2>&1
python tools/step_time.py > $OUT/plain_v2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_fir -s 20 -c 1 -o $OUT/prof_v2 python tools/step_time.py > $OUT/ncu_v2.log 2>&1
cat $OUT/plain_v1.log $OUT/plain_v2.log; ls -la $OUT
