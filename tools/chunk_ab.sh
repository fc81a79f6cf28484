#!/bin/bash
# stage-A leg as a function of the latent chunk size (token slots per internal chunk), alternating, same box
mkdir -p gpurun_out
for i in 1 2 3; do
  for t in ${@:-524288 1048576}; do
    NRB200_BENCH_STAGE_A_TOKENS=$t timeout 300 python bench.py --only-stage-a 2>> gpurun_out/chunk_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])['stage_a']
print('tokens', $t, 'ms', d['ms'], 'frac_executed', d['roofline']['frac_executed'], 'launches', d['kernel_launches_per_call'])"
  done
done
