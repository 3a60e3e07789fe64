#!/bin/bash
cd $GRAFT_REPO_ROOT
OUT=gpurun_out/r10; mkdir -p $OUT
timeout 300 oracle/_ref/vtm/DecoderApp_gpu -b tests/golden/e2e/mm_ldp_q37.bin -o /tmp/d.yuv > $OUT/dec.log 2>&1; echo rc=$?; tail -5 $OUT/dec.log
MMGPU_CLASSIC=0 timeout 300 oracle/_ref/vtm/DecoderApp_gpu -b tests/golden/e2e/mm_ldp_q37.bin -o /tmp/d.yuv > $OUT/dec_noclassic.log 2>&1; echo rc=$?; tail -3 $OUT/dec_noclassic.log
MMGPU_MVP=0 timeout 300 oracle/_ref/vtm/DecoderApp_gpu -b tests/golden/e2e/mm_ldp_q37.bin -o /tmp/d.yuv > $OUT/dec_nomvp.log 2>&1; echo rc=$?; tail -3 $OUT/dec_nomvp.log
