#!/bin/bash
# round 2, GPU call 16 (1 GPU): short tail chunk of the host-fed encode: parity of the host routes, e2e A/B
mkdir -p gpurun_out/r02_16 && O=gpurun_out/r02_16
( timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py tests/test_gpu_protocol.py -m gpu -q 2>&1 | tail -6 ) > $O/pytest.txt
LCPC_B200_H2D_TAIL_DIV=0 timeout 600 python bench.py --steps 20 --warmup 5 --workload ligero --no-cpu-baseline > $O/bench_tail0.json 2> $O/bench_tail0.err
for d in 4 8; do
LCPC_B200_H2D_TAIL_DIV=$d timeout 600 python bench.py --steps 20 --warmup 5 --workload ligero --no-cpu-baseline > $O/bench_tail$d.json 2> $O/bench_tail$d.err
done
echo done > $O/done
