#!/bin/bash
# round 2, GPU call 11 (2 GPUs): pipelined hash stream -- parity tests, A/B at N=2
mkdir -p gpurun_out/r02_11 && O=gpurun_out/r02_11
export LCPC_B200_SHARD_TIMEOUT_MS=10000
( timeout 900 python -m pytest tests/test_shard.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -15 ) > $O/pytest.txt
unset LCPC_B200_SHARD_TIMEOUT_MS
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for pl in 0 1; do
  LCPC_B200_SHARD_PIPELINE=$pl timeout 400 $R --master-port 2955$pl bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_g2_pipeline$pl.json 2> $O/bench_g2_pipeline$pl.err
done
echo done > $O/done
