#!/bin/bash
# round 2, GPU call 3 (8 GPUs): the sharded commit at N=8 -- default bench (Ligero + Brakedown at 2^24), then config 5
# (Brakedown/Ft127 2^28) with the sampled oracle check
mkdir -p gpurun_out/r02_3 && O=gpurun_out/r02_3
nvidia-smi -L > $O/smi.txt 2>&1
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_g8.json 2> $O/bench_g8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 10 --warmup 3 --workload brakedown --lgl 28 > $O/bench_g8_brakedown28.json 2> $O/bench_g8_brakedown28.err
echo done > $O/done
