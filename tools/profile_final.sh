#!/bin/bash
# Round-1 final evidence: plain bench (both arms), then ncu launch list + full capture of the dominant kernel.
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_reference.json 2> gpurun_out/bench_final_reference.err
B="python bench.py --steps 2 --warmup 1 --no-stage-a --no-cpu-baseline --no-e2e"
A="python bench.py --only-stage-a"
$B > gpurun_out/plain_b.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_final.csv $B > gpurun_out/ncu_b.log 2>&1
$B > gpurun_out/plain_b2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:score_rank_kernel -s 3 -c 1 -o gpurun_out/prof_score_rank_final $B > gpurun_out/ncu_b2.log 2>&1
$A > gpurun_out/plain_a.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_stage_a_final.csv $A > gpurun_out/ncu_a.log 2>&1
$A > gpurun_out/plain_a2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 8 -c 4 -o gpurun_out/prof_gemm_latent_final $A > gpurun_out/ncu_a2.log 2>&1
tail -c 400 gpurun_out/bench_final.json
