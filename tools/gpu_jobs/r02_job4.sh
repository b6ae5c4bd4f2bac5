#!/bin/bash
# round 2, GPU call 4 (1 GPU): full parity suite with the device code generator as default, pipelined-gather spmm A/B
# (+ ncu dram/L2 evidence), sanitizer on the new kernels, default bench line
mkdir -p gpurun_out/r02_4 && O=gpurun_out/r02_4
export LCPC_B200_SHARD_TIMEOUT_MS=8000
( timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > $O/pytest_gpu.txt
( timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "headline_shape_setup_time" 2>&1 | tail -8 ) > $O/matgen_time.txt
( LCPC_B200_MATGEN=host timeout 300 python -c "
import time, lcpc_b200 as P
ctx = P.Context(0)
t0 = time.perf_counter(); enc = P.SdigEncoding(P.FT127, 1 << 24, seed=0, ctx=ctx); ctx.synchronize()
print('host matgen + upload 2^24: %.1f ms' % ((time.perf_counter() - t0) * 1e3))
" 2>&1 | tail -3 ) >> $O/matgen_time.txt
unset LCPC_B200_SHARD_TIMEOUT_MS
timeout 600 python tools/ab_sweep.py brakedown SPMM_PIPE=0,1 SPMM_WINDOW_KB=0,49152,65536,98304 > $O/ab_brakedown_pipe.jsonl 2> $O/ab_brakedown_pipe.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_bytes.sum,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active
LCPC_B200_SPMM_PIPE=1 timeout 600 ncu --metrics $M --clock-control none -k regex:spmm -c 12 --csv --log-file $O/ncu_spmm_pipe.csv python tools/ab_sweep.py brakedown --steps 1 > $O/ncu_pipe.log 2>&1
LCPC_B200_SPMM_PIPE=1 LCPC_B200_SPMM_WINDOW_KB=65536 timeout 600 ncu --metrics $M --clock-control none -k regex:spmm -c 24 --csv --log-file $O/ncu_spmm_pipe_chunk64.csv python tools/ab_sweep.py brakedown --steps 1 > $O/ncu_pipe_chunk.log 2>&1
( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_shard.py -m gpu -x -q -k "device_matgen_equals or schedules_are_result_neutral or multi_commit_and_prove_equal_the_oracle and one" 2>&1 | tail -15 ) > $O/sanitizer_memcheck_new.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err
echo done > $O/done
