#!/bin/bash
# round 2, GPU call 42 (1 GPU): prove / verify timing repeated (one sample per bench run)
mkdir -p gpurun_out/r02_42 && O=gpurun_out/r02_42
timeout 120 python bench.py --steps 5 --warmup 3 --workload ligero > $O/bench1.json 2> $O/bench1.err
echo done > $O/done
