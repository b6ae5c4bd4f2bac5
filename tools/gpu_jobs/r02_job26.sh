#!/bin/bash
# round 2, GPU call 26 (1 GPU): deeper gather batches in the sparse products; the leaf hash of one rank's column block
mkdir -p gpurun_out/r02_26 && O=gpurun_out/r02_26
timeout 300 python tools/ab_sweep.py --steps 20 brakedown SPMM_SUM_DEEP=0,1 > $O/ab_deep.jsonl 2> $O/ab_deep.err
timeout 300 python tools/ab_sweep.py --steps 10 --lgl 28 brakedown SPMM_SUM_DEEP=0,1 >> $O/ab_deep.jsonl 2>> $O/ab_deep.err
timeout 300 python tools/ab_sweep.py --steps 20 --rows 9 brakedown SPMM_SUM_DEEP=0,1 SPMM_MAC=1 >> $O/ab_deep.jsonl 2>> $O/ab_deep.err
timeout 300 python tools/ab_sweep.py --steps 20 --rows 9 brakedown SPMM_MAC=0 >> $O/ab_deep.jsonl 2>> $O/ab_deep.err
timeout 300 python tools/ab_sweep.py --steps 20 --per-row 8192 --rows 256 ligero LEAF_SMEM_PAD_KB=0 > $O/ab_leaf_block.jsonl 2> $O/ab_leaf_block.err
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "brakedown or sdig or expander" 2>&1 | tail -4 ) > $O/pytest.txt
echo done > $O/done
