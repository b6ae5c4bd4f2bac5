#!/bin/bash
# round 2, GPU call 34 (1 GPU): NTT store phase batched (4 / 8 granules read from shared memory before the stores); load batch 2 / 6 beside 4
mkdir -p gpurun_out/r02_34 && O=gpurun_out/r02_34
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -k "ligero or ntt or encode" 2>&1 | tail -4 ) > $O/pytest.txt
for v in "" sb4 sb8 lb2 lb6 "" sb4; do
  if [ -z "$v" ]; then L=lcpc_b200/lib/liblcpc_b200.so; else L=lcpc_b200/lib/liblcpc_b200_$v.so; fi
  echo "{\"lib\": \"$L\"}" >> $O/ab_ntt.jsonl
  LCPC_B200_LIB=$PWD/$L timeout 300 python tools/ab_sweep.py --steps 20 ligero >> $O/ab_ntt.jsonl 2>> $O/ab_ntt.err
done
LCPC_B200_LIB=$PWD/lcpc_b200/lib/liblcpc_b200_sb4.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ligero or ntt or encode" 2>&1 | tail -2 > $O/pytest_sb4.txt
echo done > $O/done
