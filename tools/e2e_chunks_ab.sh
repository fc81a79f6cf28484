#!/bin/bash
# e2e (host buffers in / out) as a function of the number of impression chunks of the H2D | score | D2H pipeline
mkdir -p gpurun_out
for c in ${@:-8 16 32}; do
  NRB200_E2E_CHUNKS=$c timeout 600 python bench.py --no-extras --no-stage-a --no-cpu-baseline --steps 5 \
      > gpurun_out/e2e_chunks_$c.json 2> gpurun_out/e2e_chunks_$c.err
  python - "$c" <<'PY'
import json, sys
c = sys.argv[1]
d = json.loads(open(f"gpurun_out/e2e_chunks_{c}.json").read().strip().splitlines()[-1])
print("chunks", c, "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "cold", d["e2e"]["cold"]["value"])
PY
done
