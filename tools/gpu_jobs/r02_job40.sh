#!/bin/bash
# round 2, GPU call 40 (4 GPUs): default bench at N=4 on the final library
mkdir -p gpurun_out/r02_40 && O=gpurun_out/r02_40
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 20 --warmup 5 > $O/bench_g4.json 2> $O/bench_g4.err
echo done > $O/done
