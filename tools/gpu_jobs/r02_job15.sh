#!/bin/bash
# round 2, GPU call 15 (4 GPUs): the N=4 point of the 1/2/4/8 table
mkdir -p gpurun_out/r02_15 && O=gpurun_out/r02_15
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 4 --steps 20 --warmup 5 > $O/bench_g4.json 2> $O/bench_g4.err
echo done > $O/done
