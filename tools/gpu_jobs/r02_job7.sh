#!/bin/bash
mkdir -p gpurun_out/r02_7 && O=gpurun_out/r02_7
timeout 300 python tools/ab_sweep.py --steps 20 brakedown SDIG_EARLY_HASH=0,1 LEAF_SMEM_PAD_KB=0,24,47 > $O/ab_early_hash.jsonl 2> $O/ab_early_hash.err
timeout 300 python tools/ab_sweep.py --steps 20 brakedown SDIG_EARLY_HASH=1 LEAF_SMEM_PAD_KB=47 SPMM_SMEM_PAD_KB=0,60 >> $O/ab_early_hash.jsonl 2>> $O/ab_early_hash.err
echo done > $O/done
