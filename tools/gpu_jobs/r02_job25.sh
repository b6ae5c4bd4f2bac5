#!/bin/bash
# round 2, GPU call 25 (1 GPU): sparse products with the carry-counting accumulator (SPMM_MAC)
mkdir -p gpurun_out/r02_25 && O=gpurun_out/r02_25
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -k "brakedown or sdig or expander or tail" 2>&1 | tail -8 ) > $O/pytest.txt
timeout 300 python tools/ab_sweep.py --steps 20 brakedown SPMM_MAC=0,1 > $O/ab_mac.jsonl 2> $O/ab_mac.err
timeout 300 python tools/ab_sweep.py --steps 20 --lgl 20 brakedown SPMM_MAC=0,1 >> $O/ab_mac.jsonl 2>> $O/ab_mac.err
timeout 300 python tools/ab_sweep.py --steps 10 --lgl 28 brakedown SPMM_MAC=0,1 >> $O/ab_mac.jsonl 2>> $O/ab_mac.err
LCPC_B200_MATGEN=host timeout 600 ncu --metrics gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,dram__bytes_read.sum --clock-control none -k regex:"spmm|transpose|fused|leaf" -c 40 --csv --log-file $O/ncu_mac.csv python tools/ab_sweep.py --steps 2 brakedown SPMM_MAC=1 > $O/ncu_run.log 2>&1
echo done > $O/done
