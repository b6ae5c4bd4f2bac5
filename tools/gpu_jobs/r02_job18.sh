#!/bin/bash
# round 2, GPU call 18 (1 GPU): wider fused window of the expander's inner levels, 72 rows and a 9-row share
mkdir -p gpurun_out/r02_18 && O=gpurun_out/r02_18
for kb in 40 100 200; do
  LCPC_B200_FUSED_SMEM_KB=$kb timeout 300 python tools/ab_sweep.py --steps 20 brakedown >> $O/ab_fused.jsonl 2>> $O/ab_fused.err
  LCPC_B200_FUSED_SMEM_KB=$kb timeout 300 python tools/ab_sweep.py --steps 20 --rows 9 brakedown >> $O/ab_fused.jsonl 2>> $O/ab_fused.err
done
( LCPC_B200_FUSED_SMEM_KB=200 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "sdig or brakedown or expander" 2>&1 | tail -4 ) > $O/pytest_fused200.txt
echo done > $O/done
