#!/bin/bash
# End-of-round check on one GPU: full parity suite, smoke(), single-GPU run of the row-sharded workload.
bash tools/gpu_suite.sh; s1=$?
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; s2=$?
tail -n 2 gpurun_out/smoke.log
timeout 600 python bench.py --workload cfg5 --table-rows 1250000 --steps 2 --warmup 1 > gpurun_out/cfg5_n1.json 2> gpurun_out/cfg5_n1.err; s3=$?
tail -n 1 gpurun_out/cfg5_n1.json | cut -c1-300; tail -n 3 gpurun_out/cfg5_n1.err
echo "== suite $s1 smoke $s2 cfg5_n1 $s3"
[ $s1 -eq 0 ] && [ $s2 -eq 0 ] && [ $s3 -eq 0 ]
