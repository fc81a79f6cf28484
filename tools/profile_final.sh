#!/bin/bash
# Round-1 final evidence, one profiler pass per invocation (each after a plain run of the same command):
#   tools/profile_final.sh bench       both bench arms, no profiler
#   tools/profile_final.sh launches    ncu launch list of the bench step
#   tools/profile_final.sh score       ncu --set full of the dominant kernel (score_rank_kernel)
#   tools/profile_final.sh launches_a  ncu launch list of stage A
#   tools/profile_final.sh gemm        ncu --set full of the four stage-A GEMMs
set -x
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-stage-a --no-cpu-baseline --no-e2e"
A="python bench.py --only-stage-a"
case "$1" in
  bench)
    python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
    python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_reference.json 2> gpurun_out/bench_final_reference.err
    tail -c 600 gpurun_out/bench_final.json; tail -c 300 gpurun_out/bench_final_reference.json ;;
  launches)
    $B > gpurun_out/plain_b.log 2>&1 && \
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_final.csv $B > gpurun_out/ncu_b.log 2>&1 ;;
  score)
    $B > gpurun_out/plain_b2.log 2>&1 && \
    ncu --set full --clock-control none --import-source on -k regex:score_rank_kernel -s 3 -c 1 -o gpurun_out/prof_score_rank_final $B > gpurun_out/ncu_b2.log 2>&1 ;;
  launches_a)
    $A > gpurun_out/plain_a.log 2>&1 && \
    ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_stage_a_final.csv $A > gpurun_out/ncu_a.log 2>&1 ;;
  gemm)
    $A > gpurun_out/plain_a2.log 2>&1 && \
    ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 8 -c 4 -o gpurun_out/prof_gemm_latent_final $A > gpurun_out/ncu_a2.log 2>&1 ;;
  *) echo "usage: $0 bench|launches|score|launches_a|gemm"; exit 2 ;;
esac
