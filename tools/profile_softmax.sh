#!/bin/bash
set -x
A="python bench.py --only-stage-a"
$A > gpurun_out/plain_a3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 8 -c 4 -o gpurun_out/prof_gemm_latent_v3 $A > gpurun_out/ncu_a3.log 2>&1
