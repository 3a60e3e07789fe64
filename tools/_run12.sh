#!/bin/bash
cd $GRAFT_REPO_ROOT
N=$1
OUT=gpurun_out/r12; mkdir -p $OUT
run() { timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 2> $OUT/err_n${N}_$1.log | tail -1 > $OUT/bench_n${N}_$1.json; python -c "
import json,sys
d=json.load(open('$OUT/bench_n${N}_$1.json'))
print('N=$N $1', 'value %.1f G/s  ms %.4f  e2e %.1f G/s  k_map %.4f k_fir %.4f sustained %.4f ms' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['roofline']['path']['k_map_ms'], d['roofline']['kernel_ms'], d['sustained']['ms_per_step']))"; }
run unpadded
MMGPU_EXCHANGE_PADDED=1 run padded
