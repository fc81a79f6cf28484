#!/bin/bash
# ncu passes, round 1 (after GEMM v2).  Each ncu run follows a plain run of the same command.
set -x
mkdir -p gpurun_out
A="python bench.py --only-stage-a"
B="python bench.py --steps 2 --warmup 1 --no-stage-a --no-cpu-baseline --no-e2e"
$A > gpurun_out/plain_a.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_stage_a_v2.csv $A > gpurun_out/ncu_a.log 2>&1
$A > gpurun_out/plain_a2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 8 -c 4 -o gpurun_out/prof_gemm_latent_v2 $A > gpurun_out/ncu_a2.log 2>&1
$B > gpurun_out/plain_b.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_v2.csv $B > gpurun_out/ncu_b.log 2>&1
$B > gpurun_out/plain_b2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:score_rank_kernel -s 3 -c 1 -o gpurun_out/prof_score_rank_v2 $B > gpurun_out/ncu_b2.log 2>&1
ls -la gpurun_out | tail -12
