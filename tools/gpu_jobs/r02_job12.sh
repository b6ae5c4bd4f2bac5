#!/bin/bash
# round 2, GPU call 12 (1 GPU): BLAKE3 add placement A/B (2 / 3 / 4 adds per G on the multiplier pipe), Brakedown launch list
mkdir -p gpurun_out/r02_12 && O=gpurun_out/r02_12
for v in "" _b3fma3 _b3fma4; do
  LCPC_B200_LIB=$PWD/lcpc_b200/lib/liblcpc_b200$v.so timeout 300 python tools/ab_sweep.py --steps 20 ligero > $O/ab_ligero$v.jsonl 2> $O/ab$v.err
  LCPC_B200_LIB=$PWD/lcpc_b200/lib/liblcpc_b200$v.so timeout 300 python tools/ab_sweep.py --steps 20 brakedown > $O/ab_brakedown$v.jsonl 2>> $O/ab$v.err
done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum
LCPC_B200_MATGEN=host timeout 600 ncu --metrics $M --clock-control none -c 80 --csv --log-file $O/launches_brakedown.csv python tools/ab_sweep.py --steps 1 brakedown > $O/ncu_b.log 2>&1
echo done > $O/done
