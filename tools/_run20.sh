#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r20
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "unpadded" 2>&1 | tail -3
for st in 48; do
echo "== TMA bulk pull, $st stages"
MMGPU_PULL_STAGES=$st timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/pull_time.py 7 2> gpurun_out/r20/err_tma$st.log | tee gpurun_out/r20/pull_time_tma$st.txt
tail -2 gpurun_out/r20/err_tma$st.log | cut -c1-300
done
