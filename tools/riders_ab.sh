#!/bin/bash
# A/B of the rider warps (LayerNorm / pooling passes inside the tcgen05 GEMM kernels) on the stage-A leg.
# usage: tools/riders_ab.sh [rounds]      -> gpurun_out/riders_ab_{0,1}.jsonl
mkdir -p gpurun_out
rm -f gpurun_out/riders_ab_0.jsonl gpurun_out/riders_ab_1.jsonl
for i in $(seq 1 ${1:-3}); do
  for r in 0 1; do
    NRB200_RIDERS=$r timeout 300 python bench.py --only-stage-a >> gpurun_out/riders_ab_$r.jsonl 2>> gpurun_out/riders_ab.err
  done
done
python - <<'PY'
import json
for r in (0, 1):
    rows = [json.loads(l)["stage_a"] for l in open(f"gpurun_out/riders_ab_{r}.jsonl") if l.strip().startswith("{")]
    print("riders", r, [(x["ms"], x["roofline"]["frac_executed"], x["kernel_launches_per_call"]) for x in rows])
PY
