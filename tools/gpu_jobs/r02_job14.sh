#!/bin/bash
# round 2, GPU call 14 (1 GPU): final single-GPU records: smoke, full suite, default bench, ncu --set full of the dominant kernels
mkdir -p gpurun_out/r02_14 && O=gpurun_out/r02_14
( timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > $O/smoke.txt
( timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 ) > $O/pytest_gpu.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ntt_pass|leaf_chunk" -s 3 -c 3 -o $O/prof_ligero24 python tools/ab_sweep.py --steps 1 ligero > $O/ncu_full_l.log 2>&1
LCPC_B200_MATGEN=host timeout 600 ncu --set full --clock-control none -k regex:"spmm_kernel|leaf_chunk|transpose_kernel|fused_levels" -s 10 -c 10 -o $O/prof_brakedown24 python tools/ab_sweep.py --steps 1 brakedown > $O/ncu_full_b.log 2>&1
ls -la $O > $O/ls.txt
echo done > $O/done
