#!/bin/bash
# round 2, GPU call 17 (1 GPU): what a rank's share of the 8-GPU Brakedown commit costs, kernel by kernel (9 rows of 72)
mkdir -p gpurun_out/r02_17 && O=gpurun_out/r02_17
timeout 300 python tools/ab_sweep.py --steps 20 --rows 9 brakedown > $O/ab_rows9.jsonl 2> $O/ab_rows9.err
timeout 300 python tools/ab_sweep.py --steps 20 --rows 18 brakedown >> $O/ab_rows9.jsonl 2>> $O/ab_rows9.err
timeout 300 python tools/ab_sweep.py --steps 20 --rows 32 ligero >> $O/ab_rows9.jsonl 2>> $O/ab_rows9.err
M=gpu__time_duration.sum,dram__bytes_read.sum,lts__t_bytes.sum
LCPC_B200_MATGEN=host timeout 600 ncu --metrics $M --clock-control none --cache-control none -c 80 --csv --log-file $O/launches_brakedown_rows9.csv python tools/ab_sweep.py --steps 1 --rows 9 brakedown > $O/ncu_b.log 2>&1
echo done > $O/done
