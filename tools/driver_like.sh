#!/bin/bash
# What the driver runs at round end on one GPU, in its order, with wall-clock times.
mkdir -p gpurun_out
t0=$(date +%s); timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/driver_pytest.log 2>&1; s1=$?; t1=$(date +%s)
tail -n 3 gpurun_out/driver_pytest.log; echo "== pytest -m gpu exit $s1 in $((t1-t0)) s"
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/driver_smoke.log 2>&1; s2=$?; t2=$(date +%s)
tail -n 2 gpurun_out/driver_smoke.log; echo "== smoke exit $s2 in $((t2-t1)) s"
timeout 900 python bench.py --impl reference > gpurun_out/driver_bench_reference.json 2> gpurun_out/driver_bench_reference.err; s3=$?; t3=$(date +%s)
tail -c 400 gpurun_out/driver_bench_reference.json; echo "== bench --impl reference exit $s3 in $((t3-t2)) s"
timeout 1500 python bench.py > gpurun_out/driver_bench.json 2> gpurun_out/driver_bench.err; s4=$?; t4=$(date +%s)
python - <<'PY'
import json
d = json.loads(open("gpurun_out/driver_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"],
      "roofline", d["roofline"]["frac"], "stage_a", d.get("stage_a", {}).get("roofline", {}).get("frac_executed"),
      "cfg3", d.get("cfg3", {}).get("roofline", {}).get("frac_executed"), "clocks", d["clocks"])
PY
echo "== bench exit $s4 in $((t4-t3)) s"
[ $s1 -eq 0 ] && [ $s2 -eq 0 ] && [ $s3 -eq 0 ] && [ $s4 -eq 0 ]
