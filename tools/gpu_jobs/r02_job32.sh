#!/bin/bash
# round 2, GPU call 32 (1 GPU): R mod p as a compile-time constant in the fused inner levels
mkdir -p gpurun_out/r02_32 && O=gpurun_out/r02_32
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_cpp_mirror.py -m gpu -q -x -k "brakedown or sdig or expander or tail or mirror" 2>&1 | tail -6 ) > $O/pytest.txt
timeout 300 python tools/ab_sweep.py --steps 20 brakedown SPMM_SPLIT=1 > $O/ab.jsonl 2> $O/ab.err
timeout 300 python tools/ab_sweep.py --steps 20 --lgl 20 brakedown SPMM_SPLIT=1 >> $O/ab.jsonl 2>> $O/ab.err
timeout 300 python tools/ab_sweep.py --steps 20 --rows 9 brakedown SPMM_SPLIT=1 >> $O/ab.jsonl 2>> $O/ab.err
LCPC_B200_MATGEN=host timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fused_levels" -s 1 -c 1 -o $O/prof_fused python tools/ab_sweep.py --steps 1 brakedown > $O/ncu_full_f.log 2>&1
echo done > $O/done
