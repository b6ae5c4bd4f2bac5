#!/bin/bash
# round 2, GPU call 2 (2 GPUs): the sharded commit behind the C ABI -- tests, then bench at N=2 (new path vs the torch path)
mkdir -p gpurun_out/r02_2 && O=gpurun_out/r02_2
nvidia-smi -L > $O/smi.txt 2>&1
nvidia-smi topo -m >> $O/smi.txt 2>&1
export LCPC_B200_SHARD_TIMEOUT_MS=8000
( timeout 900 python -m pytest tests/test_shard.py -m gpu -x -q 2>&1 | tail -25 ) > $O/pytest_shard.txt
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 ) > $O/pytest_gpu.txt
unset LCPC_B200_SHARD_TIMEOUT_MS
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_g2.json 2> $O/bench_g2.err
LCPC_B200_TRANSPORT=torch timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --workload ligero > $O/bench_g2_torch.json 2> $O/bench_g2_torch.err
echo done > $O/done
