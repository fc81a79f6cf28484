#!/bin/bash
# Round-2 evidence in ONE gpurun call (1 GPU): every profiler pass runs after a plain run of the same command.
#   gpurun --timeout 1500 -- 'bash tools/profile_r02.sh'
# Outputs (gpurun_out/): launch lists (csv) and .ncu-rep files; summarise here with tools/summarize_r02.sh.
set -x
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-stage-a --no-cpu-baseline --no-e2e --no-extras"
A="python bench.py --only-stage-a"
NCU="ncu --clock-control none"
# 1. bench step: plain run (its JSON carries the algorithmic bytes the traffic file is keyed by), launch list,
#    full capture of the dominant kernel and of the five FinalAttention transform GEMMs at the bench shape
$B > gpurun_out/r02_plain_bench.json 2> gpurun_out/r02_plain_bench.err || exit 1
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv $B > gpurun_out/r02_ncu_b.log 2>&1
$NCU --set full --import-source on -k regex:score_rank_kernel -s 3 -c 1 -o gpurun_out/r02_prof_score_rank -f $B > gpurun_out/r02_ncu_b2.log 2>&1
$NCU --set full --import-source on -k regex:gemm_tc_kernel -s 15 -c 5 -o gpurun_out/r02_prof_gemm_final_attention -f $B > gpurun_out/r02_ncu_b3.log 2>&1
# 2. stage A: plain run, launch list, full capture of one chunk's kernels (GEMMs + LayerNorm + pooling)
$A > gpurun_out/r02_plain_stage_a.json 2> gpurun_out/r02_plain_stage_a.err || exit 1
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r02_launches_stage_a.csv $A > gpurun_out/r02_ncu_a.log 2>&1
$NCU --set full --import-source on -k regex:'gemm_tc_kernel|layer_norm|pool_items' -s 40 -c 8 -o gpurun_out/r02_prof_stage_a -f $A > gpurun_out/r02_ncu_a2.log 2>&1
ls -la gpurun_out/*.ncu-rep
