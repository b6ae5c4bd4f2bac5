#!/bin/bash
# round 2, GPU call 31 (1 GPU): final library: parity, A/B of the split thresholds, ncu --set full of the dominant kernels of both workloads
mkdir -p gpurun_out/r02_31 && O=gpurun_out/r02_31
( timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 ) > $O/pytest.txt
timeout 300 python tools/ab_sweep.py --steps 20 brakedown SPMM_SPLIT=0,1 > $O/ab_split2.jsonl 2> $O/ab_split2.err
timeout 300 python tools/ab_sweep.py --steps 20 --lgl 20 brakedown SPMM_SPLIT=0,1 >> $O/ab_split2.jsonl 2>> $O/ab_split2.err
timeout 300 python tools/ab_sweep.py --steps 20 --rows 9 brakedown SPMM_SPLIT=0,1 >> $O/ab_split2.jsonl 2>> $O/ab_split2.err
timeout 300 python tools/ab_sweep.py --steps 20 --lgl 22 brakedown SPMM_SPLIT=0,1 >> $O/ab_split2.jsonl 2>> $O/ab_split2.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ntt_pass|leaf_chunk" -s 3 -c 3 -o $O/prof_ligero24_final python tools/ab_sweep.py --steps 1 ligero > $O/ncu_full_l.log 2>&1
LCPC_B200_MATGEN=host timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spmm_sum|leaf_chunk|transpose_kernel|fused_levels" -s 9 -c 9 -o $O/prof_brakedown24_final python tools/ab_sweep.py --steps 1 brakedown > $O/ncu_full_b.log 2>&1
ls -la $O > $O/ls.txt
echo done > $O/done
