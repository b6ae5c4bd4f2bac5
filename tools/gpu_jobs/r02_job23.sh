#!/bin/bash
# round 2, GPU call 23 (1 GPU): what the chunking of the host->device copy costs by itself; fewer, larger row-chunks
mkdir -p gpurun_out/r02_23 && O=gpurun_out/r02_23
timeout 300 python tools/h2d_chunks.py 512 > $O/h2d_chunks.jsonl 2> $O/h2d_chunks.err
for n in 6 10 15; do
LCPC_B200_H2D_MAX_CHUNKS=$n timeout 600 python bench.py --steps 20 --warmup 5 --workload ligero > $O/bench_ligero_chunks$n.json 2> $O/bench_ligero_chunks$n.err
done
echo done > $O/done
