#!/bin/bash
cd $GRAFT_REPO_ROOT
N=$1; shift
OUT=gpurun_out/r18; mkdir -p $OUT
run() { tag=$1; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 2> $OUT/err_n${N}_$tag.log | tail -1 > $OUT/bench_n${N}_$tag.json; python -c "
import json,sys
d=json.load(open('$OUT/bench_n${N}_$tag.json'))
print('N=$N $tag', 'value %.1f G/s  ms %.4f  e2e %.1f G/s  k_map %.4f k_fir %.4f sustained %.4f ms' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['roofline']['path']['k_map_ms'], d['roofline']['kernel_ms'], d['sustained']['ms_per_step']))" || tail -5 $OUT/err_n${N}_$tag.log; }
for v in "$@"; do
  case $v in
    nccl) MMGPU_EXCHANGE=nccl MMGPU_EXCHANGE_PADDED=0 run nccl_unpadded ;;
    padded) MMGPU_EXCHANGE=nccl MMGPU_EXCHANGE_PADDED=1 run nccl_padded ;;
    default) run default ;;
    ring2) MMGPU_RING=2 run ring2 ;;
    pull*) MMGPU_EXCHANGE=pull MMGPU_PULL_CTAS=${v#pull} run $v ;;
    push*) MMGPU_EXCHANGE=push MMGPU_PUSH_CTAS=${v#push} run $v ;;
  esac
done
