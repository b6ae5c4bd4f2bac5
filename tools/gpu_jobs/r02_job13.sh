#!/bin/bash
# round 2, GPU call 13 (8 GPUs): final 8-GPU record (chunked encode + pipelined commits by default) and the pipeline A/B
mkdir -p gpurun_out/r02_13 && O=gpurun_out/r02_13
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 500 $R --master-port 29561 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_g8.json 2> $O/bench_g8.err
LCPC_B200_SHARD_PIPELINE=0 timeout 500 $R --master-port 29562 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_g8_pipeline0.json 2> $O/bench_g8_pipeline0.err
echo done > $O/done
