#!/bin/bash
# tools/profile.sh -- the ncu captures behind profiles/ (run on the GPU box: gpurun -- bash tools/profile.sh).
# Summaries are produced HERE afterwards (no GPU needed):
#   python tools/ncu_summary.py gpurun_out/prof_ligero24.ncu-rep > profiles/rNN_ncu_full_ligero_2_24_summary.csv
#   python tools/ncu_opmix.py  gpurun_out/prof_ligero24.ncu-rep ntt_pass 1 > profiles/rNN_ncu_opmix_ntt_pass1.txt
set -u
mkdir -p gpurun_out
# every launch of the bench command with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_ligero24.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 17 -c 17 --csv --log-file gpurun_out/launches_brakedown24.csv \
    python bench.py --workload brakedown --steps 2 --warmup 1 --no-cpu-baseline >> gpurun_out/ncu_launches.log 2>&1
# full sets of the dominant kernels (skip the first commit's launches: -s)
ncu --set full --clock-control none --import-source on -k regex:"ntt_pass|leaf_chunk" -s 3 -c 3 -o gpurun_out/prof_ligero24 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none -k regex:"spmm_kernel|leaf_chunk|transpose_kernel|fused_levels" -s 11 -c 11 \
    -o gpurun_out/prof_brakedown24 python bench.py --workload brakedown --steps 1 --warmup 1 --no-cpu-baseline >> gpurun_out/ncu_full.log 2>&1
# prove / verify kernels (collapse, check_columns, dot, expand_tensor, gathers) of one proof at the bench size
ncu --set full --clock-control none -k regex:"collapse_kernel|check_columns|dot_kernel|expand_tensor|gather_|transpose_columns" -c 24 \
    -o gpurun_out/prof_prove_verify python bench.py --steps 1 --warmup 1 --no-cpu-baseline >> gpurun_out/ncu_full.log 2>&1
# host side of prove / verify: the transcript absorb per Keccak build on this box's CPU
python tools/time_transcript.py > gpurun_out/time_transcript.txt 2>&1
# pipe ceilings the design rests on
for m in microbench microbench_fp64; do
  [ -x tools/$m ] || nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/$m tools/$m.cu
  ./tools/$m > gpurun_out/$m.txt 2>&1
done
ls -la gpurun_out
