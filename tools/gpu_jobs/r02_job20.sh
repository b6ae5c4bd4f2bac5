#!/bin/bash
# round 2, GPU call 20 (1 GPU): Brakedown without the final transpose (leaves and openings from the column-major work buffer)
mkdir -p gpurun_out/r02_20 && O=gpurun_out/r02_20
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_protocol.py -m gpu -q 2>&1 | tail -8 ) > $O/pytest.txt
timeout 300 python tools/ab_sweep.py --steps 20 brakedown SDIG_LAZY_COMM=0,1 > $O/ab_lazy.jsonl 2> $O/ab_lazy.err
timeout 300 python tools/ab_sweep.py --steps 20 --lgl 28 brakedown SDIG_LAZY_COMM=0,1 >> $O/ab_lazy.jsonl 2>> $O/ab_lazy.err
timeout 300 python tools/ab_sweep.py --steps 20 --lgl 20 brakedown SDIG_LAZY_COMM=0,1 >> $O/ab_lazy.jsonl 2>> $O/ab_lazy.err
timeout 600 python bench.py --steps 20 --warmup 5 --workload brakedown > $O/bench_brakedown.json 2> $O/bench_brakedown.err
echo done > $O/done
