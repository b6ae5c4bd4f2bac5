#!/bin/bash
# round 2, GPU call 39 (1 GPU): sanity of the clean rebuild of the final sources: smoke + the parity suite
mkdir -p gpurun_out/r02_39 && O=gpurun_out/r02_39
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 ) > $O/smoke.txt
( timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 ) > $O/pytest.txt
echo done > $O/done
