#!/bin/bash
# round 2, GPU call 43 (1 GPU): sharded prove / protocol tests on the library with the faster host transcript
mkdir -p gpurun_out/r02_43 && O=gpurun_out/r02_43
( timeout 100 python -m pytest tests/test_shard.py tests/test_gpu_protocol.py tests/test_cpp_mirror.py -m gpu -q -x 2>&1 | tail -3 ) > $O/pytest.txt
echo done > $O/done
