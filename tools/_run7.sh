#!/bin/bash
cd $GRAFT_REPO_ROOT
OUT=gpurun_out/r7; mkdir -p $OUT
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 2> $OUT/err_$1.log | tail -1 > $OUT/bench_n2_$1.json; python -c "
import json,sys
d=json.load(open('$OUT/bench_n2_$1.json'))
print('$1', 'value %.1f G/s  ms %.4f  e2e %.1f G/s  k_map %.4f k_fir %.4f sustained %.4f ms' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['roofline']['path']['k_map_ms'], d['roofline']['kernel_ms'], d['sustained']['ms_per_step']))"; }
run default
NCCL_MAX_CTAS=4 run ctas4
NCCL_MAX_CTAS=8 run ctas8
tail -3 $OUT/err_default.log
