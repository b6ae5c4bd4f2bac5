#!/bin/bash
# round 2, GPU call 1: parity suite, Brakedown schedule sweep (+ ncu dram evidence), Ligero knob sweep, default bench line
mkdir -p gpurun_out/r02_1 && O=gpurun_out/r02_1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $O/pytest_gpu.txt
timeout 600 python tools/ab_sweep.py brakedown SPMM_HINTS=0,1 SPMM_WINDOW_KB=0,24576,32768,49152,65536,98304 > $O/ab_brakedown_chunks.jsonl 2> $O/ab_brakedown_chunks.err
timeout 300 python tools/ab_sweep.py brakedown SPMM_HINTS=0,1 SPMM_WINDOW_KB=0 SPMM_SLICE_KB=32768,49152,98304 > $O/ab_brakedown_slices.jsonl 2>> $O/ab_brakedown_chunks.err
LCPC_B200_L2_PERSIST_MB=64 timeout 300 python tools/ab_sweep.py brakedown SPMM_HINTS=1 SPMM_WINDOW_KB=0,32768,49152,65536 > $O/ab_brakedown_persist64.jsonl 2>> $O/ab_brakedown_chunks.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_bytes.sum
LCPC_B200_SPMM_HINTS=0 LCPC_B200_SPMM_WINDOW_KB=0 timeout 600 ncu --metrics $M --clock-control none -k regex:spmm_kernel -c 24 --csv --log-file $O/ncu_spmm_base.csv python tools/ab_sweep.py brakedown --steps 1 > $O/ncu_base.log 2>&1
LCPC_B200_SPMM_HINTS=1 LCPC_B200_SPMM_WINDOW_KB=49152 timeout 600 ncu --metrics $M --clock-control none -k regex:spmm_kernel -c 36 --csv --log-file $O/ncu_spmm_chunk48.csv python tools/ab_sweep.py brakedown --steps 1 > $O/ncu_chunk.log 2>&1
LCPC_B200_SPMM_HINTS=1 LCPC_B200_SPMM_WINDOW_KB=0 timeout 600 ncu --metrics $M --clock-control none -k regex:spmm_kernel -c 24 --csv --log-file $O/ncu_spmm_hints.csv python tools/ab_sweep.py brakedown --steps 1 > $O/ncu_hints.log 2>&1
for v in "" _o1 _b3 _b3o1; do
  LCPC_B200_LIB=$PWD/lcpc_b200/lib/liblcpc_b200$v.so timeout 300 python tools/ab_sweep.py ligero DEV_CHUNKS=1,4,8 NTT_SMEM_PAD_KB=0,28 > $O/ab_ligero$v.jsonl 2> $O/ab_ligero$v.err
done
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err
echo done > $O/done
