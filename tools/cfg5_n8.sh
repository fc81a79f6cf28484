#!/bin/bash
# BASELINE configs[4] at full size on 8 GPUs: 10 M-row table, row-sharded, all-gather overlapped with the transform.
# usage: tools/cfg5_n8.sh [gather ...]   (default: nvls dma; nvls = the all-gather riding inside the GEMM kernels)
mkdir -p gpurun_out
for G in ${@:-nvls dma}; do
  timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 \
      bench.py --gpus 8 --steps 5 --warmup 2 --workload cfg5 --table-rows 10000000 --gather $G > gpurun_out/cfg5_n8_$G.json 2> gpurun_out/cfg5_n8_$G.err
  echo "== $G exit $?"; tail -n 1 gpurun_out/cfg5_n8_$G.json; grep -v "NCCL\|OMP_NUM\|\*\*\*" gpurun_out/cfg5_n8_$G.err | tail -n 5
done
