#!/bin/bash
cd $GRAFT_REPO_ROOT
OUT=gpurun_out/r8; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "unpadded or decoder_shape or merge_satd" 2>&1 | tail -4 > $OUT/pytest.log; cat $OUT/pytest.log
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 2> $OUT/err_$1.log | tail -1 > $OUT/bench_n2_$1.json; python -c "
import json,sys
d=json.load(open('$OUT/bench_n2_$1.json'))
print('$1', 'value %.1f G/s  ms %.4f  e2e %.1f G/s  k_map %.4f k_fir %.4f sustained %.4f ms' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['roofline']['path']['k_map_ms'], d['roofline']['kernel_ms'], d['sustained']['ms_per_step']))"; }
run unpadded
MMGPU_EXCHANGE_PADDED=1 run padded
MMGPU_EXCHANGE=p2p run unpadded_p2p
python tools/mc_jobs_time.py 2>&1 | tail -2
