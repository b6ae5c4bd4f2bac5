#!/bin/bash
# round 2, GPU call 28 (8 GPUs): default bench at N=8 on the final library
mkdir -p gpurun_out/r02_28 && O=gpurun_out/r02_28
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_g8.json 2> $O/bench_g8.err
echo done > $O/done
