#!/bin/bash
# round 2, GPU call 22 (1 GPU): Ligero host-route commit with a half-size chunk before the quarter-size tail
mkdir -p gpurun_out/r02_22 && O=gpurun_out/r02_22
( timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -8 ) > $O/pytest.txt
LCPC_B200_H2D_PRE_TAIL=0 timeout 600 python bench.py --steps 20 --warmup 5 --workload ligero > $O/bench_ligero_pre0.json 2> $O/bench_ligero_pre0.err
LCPC_B200_H2D_PRE_TAIL=1 timeout 600 python bench.py --steps 20 --warmup 5 --workload ligero > $O/bench_ligero_pre1.json 2> $O/bench_ligero_pre1.err
echo done > $O/done
