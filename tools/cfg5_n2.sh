#!/bin/bash
# 2-GPU check of the row-sharded path: parity test, then cfg5 at 2.5 M rows for each all-gather variant.
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_multi.py -q -m gpu --no-header -p no:cacheprovider > gpurun_out/test_gpu_multi.log 2>&1
echo "== multi exit $?"; tail -n 5 gpurun_out/test_gpu_multi.log
run() {  # name, extra args
  timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
      bench.py --gpus 2 --steps 3 --warmup 1 --workload cfg5 --table-rows 2500000 $2 > gpurun_out/cfg5_n2_$1.json 2> gpurun_out/cfg5_n2_$1.err
  echo "== $1 exit $?"; tail -n 1 gpurun_out/cfg5_n2_$1.json | cut -c1-120; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/cfg5_n2_$1.err | tail -n 5
}
run p2p "--gather p2p"
run dma "--gather dma"
run nccl "--gather nccl"
run nvls "--gather nvls"
