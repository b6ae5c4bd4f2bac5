#!/bin/bash
# round 2, GPU call 5 (1 GPU): bulk-copy (TMA) gather kernel + early hash of the systematic columns: parity, A/B, ncu
mkdir -p gpurun_out/r02_5 && O=gpurun_out/r02_5
( timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_protocol.py -m gpu -q -k "sdig or brakedown or matgen" 2>&1 | tail -25 ) > $O/pytest_sdig.txt
timeout 300 python tools/ab_sweep.py brakedown SDIG_EARLY_HASH=0,1 SPMM_SMEM_PAD_KB=0,60 SPMM_BULK=0 > $O/ab_early_hash.jsonl 2> $O/ab_early_hash.err
timeout 300 python tools/ab_sweep.py brakedown SDIG_EARLY_HASH=0,1 SPMM_BULK=1 > $O/ab_bulk.jsonl 2> $O/ab_bulk.err
M=gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,lts__t_bytes.sum,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active
LCPC_B200_SPMM_BULK=1 timeout 300 ncu --metrics $M --clock-control none -k regex:spmm -c 12 --csv --log-file $O/ncu_spmm_bulk.csv python tools/ab_sweep.py brakedown --steps 1 > $O/ncu_bulk.log 2>&1
( LCPC_B200_SPMM_BULK=1 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_sdig_encode_vs_oracle or test_commit_brakedown_vs_oracle" 2>&1 | tail -12 ) > $O/sanitizer_bulk.txt
timeout 600 python bench.py --steps 20 --warmup 5 --workload brakedown > $O/bench_brakedown.json 2> $O/bench_brakedown.err
echo done > $O/done
