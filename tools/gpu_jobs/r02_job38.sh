#!/bin/bash
# round 2, GPU call 38 (1 GPU): Merkle layers: wide one-thread-per-node launches against subtrees only
mkdir -p gpurun_out/r02_38 && O=gpurun_out/r02_38
timeout 300 python tools/ab_sweep.py --steps 20 brakedown MERKLE_WIDE_MIN=37888,0,300000 > $O/ab_merkle.jsonl 2> $O/ab_merkle.err
timeout 300 python tools/ab_sweep.py --steps 20 --lgl 20 brakedown MERKLE_WIDE_MIN=37888,0 >> $O/ab_merkle.jsonl 2>> $O/ab_merkle.err
timeout 300 python tools/ab_sweep.py --steps 20 ligero MERKLE_WIDE_MIN=37888,0 >> $O/ab_merkle.jsonl 2>> $O/ab_merkle.err
timeout 300 python tools/ab_sweep.py --steps 10 --lgl 28 brakedown MERKLE_WIDE_MIN=37888,0 >> $O/ab_merkle.jsonl 2>> $O/ab_merkle.err
( LCPC_B200_MERKLE_WIDE_MIN=0 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "merkle or commit" 2>&1 | tail -3 ) > $O/pytest.txt
echo done > $O/done
