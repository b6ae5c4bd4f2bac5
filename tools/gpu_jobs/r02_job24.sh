#!/bin/bash
# round 2, GPU call 24 (1 GPU): Brakedown device commit that leaves the coefficient rows in the work buffer as well
mkdir -p gpurun_out/r02_24 && O=gpurun_out/r02_24
( timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 ) > $O/pytest.txt
timeout 300 python tools/ab_sweep.py --steps 20 brakedown SDIG_LAZY_COMM=0,1 > $O/ab_lazy.jsonl 2> $O/ab_lazy.err
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err
echo done > $O/done
