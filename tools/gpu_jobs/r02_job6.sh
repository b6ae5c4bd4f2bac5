#!/bin/bash
# round 2, GPU call 6 (1 GPU): early hash of the systematic columns (ragged-safe) A/B, device matgen timing, sdig parity
mkdir -p gpurun_out/r02_6 && O=gpurun_out/r02_6
( timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_protocol.py tests/test_cpp_mirror.py -m gpu -q 2>&1 | tail -15 ) > $O/pytest_subset.txt
timeout 300 python tools/ab_sweep.py brakedown --steps 20 SDIG_EARLY_HASH=0,1 LEAF_SMEM_PAD_KB=0,24,47 > $O/ab_early_hash.jsonl 2> $O/ab_early_hash.err
timeout 300 python tools/ab_sweep.py brakedown --steps 20 SDIG_EARLY_HASH=1 LEAF_SMEM_PAD_KB=47 SPMM_SMEM_PAD_KB=0,60 >> $O/ab_early_hash.jsonl 2>> $O/ab_early_hash.err
( timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "headline_shape_setup_time" 2>&1 | tail -5 ) > $O/matgen_time.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_brakedown.csv python tools/ab_sweep.py brakedown --steps 1 > $O/ncu_launches.log 2>&1
echo done > $O/done
