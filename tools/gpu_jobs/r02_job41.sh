#!/bin/bash
# round 2, GPU call 41 (1 GPU): the faster host transcript (plane-wise Keccak, merged STROBE absorbs) in prove / verify
mkdir -p gpurun_out/r02_41 && O=gpurun_out/r02_41
timeout 120 python tools/time_transcript.py > $O/time_transcript.txt 2>&1
( timeout 300 python -m pytest tests/test_gpu_protocol.py -m gpu -q -x 2>&1 | tail -3 ) > $O/pytest.txt
timeout 300 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err
echo done > $O/done
