#!/bin/bash
# round 2, GPU call 33 (1 GPU): NTT tile staging with 4 / 8 / 16 loads in flight per thread against the plain loop
mkdir -p gpurun_out/r02_33 && O=gpurun_out/r02_33
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ligero or ntt or encode" 2>&1 | tail -4 ) > $O/pytest.txt
for v in lb0 "" lb4 lb16 lb0 ""; do
  if [ -z "$v" ]; then L=lcpc_b200/lib/liblcpc_b200.so; else L=lcpc_b200/lib/liblcpc_b200_$v.so; fi
  echo "{\"lib\": \"$L\"}" >> $O/ab_ntt.jsonl
  LCPC_B200_LIB=$PWD/$L timeout 300 python tools/ab_sweep.py --steps 20 ligero >> $O/ab_ntt.jsonl 2>> $O/ab_ntt.err
done
LCPC_B200_LIB=$PWD/lcpc_b200/lib/liblcpc_b200_lb0.so timeout 300 python tools/ab_sweep.py --steps 20 --lgl 20 ligero >> $O/ab_ntt.jsonl 2>> $O/ab_ntt.err
timeout 300 python tools/ab_sweep.py --steps 20 --lgl 20 ligero >> $O/ab_ntt.jsonl 2>> $O/ab_ntt.err
echo done > $O/done
