#!/bin/bash
cd $GRAFT_REPO_ROOT
OUT=gpurun_out/r5; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_e2e_reference_apps.py -m gpu -x -q -k "not app_with" 2>&1 | tail -5 > $OUT/pytest.log; cat $OUT/pytest.log
bash tools/variant_bench.sh build/variants/v1.so vvc-extension-mm_b200/libmmgpu.so build/variants/v3_w2.so 2>&1 | tee $OUT/variants.txt
python bench.py --steps 20 --warmup 5 --cpu-budget 6 > $OUT/bench.json 2> $OUT/bench.err; tail -c 3000 $OUT/bench.json; tail -3 $OUT/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2>> $OUT/bench.err; cut -c1-300 $OUT/bench_ref.json
