#!/bin/bash
# round 2, GPU call 9 (1 GPU): final single-GPU records: suite, default bench, sweep 2^20..2^28, rho = 1/4, launch lists, dram traffic
mkdir -p gpurun_out/r02_9 && O=gpurun_out/r02_9
( timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 ) > $O/pytest_gpu.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err
timeout 600 python bench.py --steps 20 --warmup 5 --workload ligero --rho-den 4 --no-cpu-baseline > $O/bench_ligero_rho4.json 2> $O/bench_ligero_rho4.err
rm -f gpurun_out/sweep_g1.jsonl
timeout 1500 python tools/sweep.py --gpus 1 --steps 10 > $O/sweep_g1.log 2>&1
cp gpurun_out/sweep_g1.jsonl $O/sweep_g1.jsonl
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum
timeout 600 ncu --metrics $M --clock-control none -c 80 --csv --log-file $O/launches_brakedown.csv python tools/ab_sweep.py --steps 1 brakedown > $O/ncu_b.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -c 60 --csv --log-file $O/launches_ligero.csv python tools/ab_sweep.py --steps 1 ligero > $O/ncu_l.log 2>&1
echo done > $O/done
