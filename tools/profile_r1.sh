#!/bin/bash
# ncu passes for round 1 (one GPU).  Each ncu run follows a plain run of the same command.
set -x
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-stage-a --no-cpu-baseline --no-e2e --impressions 1200000"
A="python bench.py --only-stage-a"
$B > gpurun_out/plain_b.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv $B > gpurun_out/ncu_b.log 2>&1
$B > gpurun_out/plain_b2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:score_rank_kernel -s 3 -c 1 -o gpurun_out/prof_score_rank $B > gpurun_out/ncu_b2.log 2>&1
$B > gpurun_out/plain_b3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 160 -c 5 -o gpurun_out/prof_gemm_fa $B > gpurun_out/ncu_b3.log 2>&1
$A > gpurun_out/plain_a.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_stage_a.csv $A > gpurun_out/ncu_a.log 2>&1
$A > gpurun_out/plain_a2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 8 -c 4 -o gpurun_out/prof_gemm_latent $A > gpurun_out/ncu_a2.log 2>&1
ls -la gpurun_out
