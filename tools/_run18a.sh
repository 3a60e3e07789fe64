#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r18
bash tools/_run18.sh 2 pull24 pull12 nccl
MMGPU_COMM_PRIORITY=0 bash tools/_run18.sh 2 nccl
