#!/bin/bash
# round 2, GPU call 27 (2 GPUs): the whole GPU suite on the final library, then the default bench at N=2
mkdir -p gpurun_out/r02_27 && O=gpurun_out/r02_27
( timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 ) > $O/pytest.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_g2.json 2> $O/bench_g2.err
echo done > $O/done
