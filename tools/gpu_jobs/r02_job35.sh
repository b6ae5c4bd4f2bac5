#!/bin/bash
# round 2, GPU call 35 (1 GPU): leaf hashing with the next block's elements requested before the current block is compressed
mkdir -p gpurun_out/r02_35 && O=gpurun_out/r02_35
for v in "" pf1 "" pf1; do
  if [ -z "$v" ]; then L=lcpc_b200/lib/liblcpc_b200.so; else L=lcpc_b200/lib/liblcpc_b200_$v.so; fi
  echo "{\"lib\": \"$L\"}" >> $O/ab_leaf.jsonl
  LCPC_B200_LIB=$PWD/$L timeout 300 python tools/ab_sweep.py --steps 20 ligero >> $O/ab_leaf.jsonl 2>> $O/ab_leaf.err
  LCPC_B200_LIB=$PWD/$L timeout 300 python tools/ab_sweep.py --steps 20 brakedown >> $O/ab_leaf.jsonl 2>> $O/ab_leaf.err
  LCPC_B200_LIB=$PWD/$L timeout 300 python tools/ab_sweep.py --steps 20 --per-row 8192 --rows 256 ligero >> $O/ab_leaf.jsonl 2>> $O/ab_leaf.err
done
LCPC_B200_LIB=$PWD/lcpc_b200/lib/liblcpc_b200_pf1.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2 > $O/pytest_pf1.txt
echo done > $O/done
