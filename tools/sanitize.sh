#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over small cases of the dense / latent / score kernels.
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh'
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
T="python -m pytest -q -m gpu -x --no-header -p no:cacheprovider"
for tool in memcheck racecheck; do
  timeout -k 10 420 $CS --tool $tool --error-exitcode 9 $T tests/test_gpu_dense.py -k "(tcgen05 and 129-64-64) or (tcgen05 and 300-768-1024) or (softmax and 300-512-128-256) or (softmax and 700-1024)" > gpurun_out/sanitize_${tool}_dense.log 2>&1
  echo "== $tool dense exit $?"; grep -E "ERROR SUMMARY|passed|failed|Error" gpurun_out/sanitize_${tool}_dense.log | tail -4
  timeout -k 10 420 $CS --tool $tool --error-exitcode 9 $T tests/test_gpu_latent.py -k "mask_semantics or packed_all_empty" > gpurun_out/sanitize_${tool}_latent.log 2>&1
  echo "== $tool latent exit $?"; grep -E "ERROR SUMMARY|passed|failed|Error" gpurun_out/sanitize_${tool}_latent.log | tail -4
  timeout -k 10 420 $CS --tool $tool --error-exitcode 9 $T tests/test_gpu_score_rank.py -k "dense_rank or int16 or float64" > gpurun_out/sanitize_${tool}_score.log 2>&1
  echo "== $tool score exit $?"; grep -E "ERROR SUMMARY|passed|failed|Error" gpurun_out/sanitize_${tool}_score.log | tail -4
done
