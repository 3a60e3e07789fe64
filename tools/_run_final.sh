#!/bin/bash
# final session of the round: tests, bench (both arms), ncu launch list and one full capture of k_map / k_fir (each after
# the same command ran clean without ncu)
cd $GRAFT_REPO_ROOT
OUT=gpurun_out/final; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $OUT/pytest.log; cat $OUT/pytest.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; tail -c 800 $OUT/bench.json; tail -2 $OUT/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; tail -c 300 $OUT/bench_ref.json
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity-check --sustain 0.01 > $OUT/plain_short.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity-check --sustain 0.01 > $OUT/ncu_launches.log 2>&1
tail -2 $OUT/ncu_launches.log | cut -c1-300
timeout 300 python tools/step_time.py > $OUT/plain_step.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_map|k_fir' -s 20 -c 2 -o $OUT/prof_kmap_kfir python tools/step_time.py > $OUT/ncu_full.log 2>&1
tail -2 $OUT/ncu_full.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
ls -la $OUT
