#!/bin/bash
# round 2, GPU call 10 (8 GPUs): final 8-GPU records + the two-stream chunked encode A/B + sweep points
mkdir -p gpurun_out/r02_10 && O=gpurun_out/r02_10
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 500 $R --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_g8.json 2> $O/bench_g8.err
LCPC_B200_SHARD_ENC_CHUNKS=2 timeout 300 $R --master-port 29542 bench.py --gpus 8 --steps 20 --warmup 5 --workload ligero > $O/bench_g8_chunks2.json 2> $O/bench_g8_chunks2.err
LCPC_B200_SHARD_ENC_CHUNKS=4 timeout 300 $R --master-port 29543 bench.py --gpus 8 --steps 20 --warmup 5 --workload ligero > $O/bench_g8_chunks4.json 2> $O/bench_g8_chunks4.err
rm -f gpurun_out/sweep_g8.jsonl
timeout 900 python tools/sweep.py --gpus 8 --steps 10 --lgls 20 22 26 28 > $O/sweep_g8.log 2>&1
cp gpurun_out/sweep_g8.jsonl $O/sweep_g8.jsonl
echo done > $O/done
