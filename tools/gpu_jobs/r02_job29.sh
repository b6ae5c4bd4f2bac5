#!/bin/bash
# round 2, GPU call 29 (1 GPU): whole GPU suite + smoke + default bench + reference arm on the final library; launch list with DRAM bytes
mkdir -p gpurun_out/r02_29 && O=gpurun_out/r02_29
( timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 ) > $O/pytest.txt
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 ) > $O/smoke.txt
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err
timeout 600 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err
LCPC_B200_MATGEN=host timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:"spmm|transpose|fused|leaf|merkle" -c 36 --csv --log-file $O/launches_brakedown.csv python tools/ab_sweep.py --steps 1 brakedown > $O/ncu_run.log 2>&1
echo done > $O/done
