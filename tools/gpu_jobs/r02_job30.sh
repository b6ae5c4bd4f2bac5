#!/bin/bash
# round 2, GPU call 30 (1 GPU): small expander levels with every output's non-zeros split over 4 / 8 lanes (SPMM_SPLIT)
mkdir -p gpurun_out/r02_30 && O=gpurun_out/r02_30
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_cpp_mirror.py -m gpu -q -x -k "brakedown or sdig or expander or tail or mirror" 2>&1 | tail -6 ) > $O/pytest.txt
timeout 300 python tools/ab_sweep.py --steps 20 brakedown SPMM_SPLIT=0,1 > $O/ab_split.jsonl 2> $O/ab_split.err
timeout 300 python tools/ab_sweep.py --steps 20 --lgl 20 brakedown SPMM_SPLIT=0,1,4,8 >> $O/ab_split.jsonl 2>> $O/ab_split.err
timeout 300 python tools/ab_sweep.py --steps 20 --rows 9 brakedown SPMM_SPLIT=0,1 >> $O/ab_split.jsonl 2>> $O/ab_split.err
timeout 300 python tools/ab_sweep.py --steps 10 --lgl 28 brakedown SPMM_SPLIT=0,1 >> $O/ab_split.jsonl 2>> $O/ab_split.err
echo done > $O/done
