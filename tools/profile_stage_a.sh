#!/bin/bash
A="python bench.py --only-stage-a"
$A > gpurun_out/plain_a.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/launches_stage_a_x.csv $A > gpurun_out/ncu_a.log 2>&1
