#!/bin/bash
# round 2, GPU call 36 (1 GPU): final library: whole GPU suite, smoke, default bench, launch list of the bench command
mkdir -p gpurun_out/r02_36 && O=gpurun_out/r02_36
( timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 ) > $O/pytest.txt
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 ) > $O/smoke.txt
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 --workload ligero > $O/bench_under_ncu.log 2>&1
echo done > $O/done
