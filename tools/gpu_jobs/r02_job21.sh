#!/bin/bash
# round 2, GPU call 21 (1 GPU): Brakedown host-route commit with the tail cut at the leaf-chunk boundary
mkdir -p gpurun_out/r02_21 && O=gpurun_out/r02_21
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -k "brakedown or sdig or tail" 2>&1 | tail -8 ) > $O/pytest.txt
LCPC_B200_H2D_TAIL_SDIG=0 timeout 600 python bench.py --steps 20 --warmup 5 --workload brakedown > $O/bench_brakedown_tail0.json 2> $O/bench_brakedown_tail0.err
LCPC_B200_H2D_TAIL_SDIG=1 timeout 600 python bench.py --steps 20 --warmup 5 --workload brakedown > $O/bench_brakedown_tail1.json 2> $O/bench_brakedown_tail1.err
echo done > $O/done
