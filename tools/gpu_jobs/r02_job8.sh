#!/bin/bash
# round 2, GPU call 8 (2 GPUs): the whole GPU suite with two devices (IPC shard worker, cross-device MultiCommit, C host program), bench N=2
mkdir -p gpurun_out/r02_8 && O=gpurun_out/r02_8
export LCPC_B200_SHARD_TIMEOUT_MS=10000
( timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 ) > $O/pytest_gpu.txt
unset LCPC_B200_SHARD_TIMEOUT_MS
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_g2.json 2> $O/bench_g2.err
echo done > $O/done
