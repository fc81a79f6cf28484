#!/bin/bash
# per-kernel durations of the stage-A leg with and without riders (ncu launch list; cold-cache, serialised)
for r in 0 1; do
  NRB200_RIDERS=$r timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file gpurun_out/rider_launches_$r.csv python bench.py --only-stage-a > gpurun_out/rider_launches_$r.log 2>&1
  python tools/summarize_ncu.py launches gpurun_out/rider_launches_$r.csv gpurun_out/rider_launches_$r.txt || true
  head -n 14 gpurun_out/rider_launches_$r.txt
done
