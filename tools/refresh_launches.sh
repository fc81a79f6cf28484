#!/bin/bash
# ncu launch lists of the bench step and of the stage-A leg with the current build (each after a plain run).
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-stage-a --no-cpu-baseline --no-e2e --no-extras"
A="python bench.py --only-stage-a"
$B > gpurun_out/final_plain_bench.json 2> gpurun_out/final_plain_bench.err || exit 1
ncu --clock-control none --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/final_launches_bench.csv $B > gpurun_out/final_ncu_b.log 2>&1
$A > gpurun_out/final_plain_stage_a.json 2> gpurun_out/final_plain_stage_a.err || exit 1
ncu --clock-control none --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/final_launches_stage_a.csv $A > gpurun_out/final_ncu_a.log 2>&1
python tools/summarize_ncu.py launches gpurun_out/final_launches_bench.csv gpurun_out/final_launches_bench.txt
python tools/summarize_ncu.py launches gpurun_out/final_launches_stage_a.csv gpurun_out/final_launches_stage_a.txt
head -n 12 gpurun_out/final_launches_bench.txt; head -n 14 gpurun_out/final_launches_stage_a.txt
