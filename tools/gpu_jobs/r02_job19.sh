#!/bin/bash
# round 2, GPU call 19 (2 GPUs): final validation of the whole GPU suite with the final build (clean rebuild) + bench N=2
mkdir -p gpurun_out/r02_19 && O=gpurun_out/r02_19
export LCPC_B200_SHARD_TIMEOUT_MS=10000
( timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 ) > $O/pytest_gpu.txt
unset LCPC_B200_SHARD_TIMEOUT_MS
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_g2.json 2> $O/bench_g2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29582 bench.py --gpus 2 --steps 3 --warmup 1 --impl reference > $O/bench_g2_reference.json 2> $O/bench_g2_reference.err
echo done > $O/done
