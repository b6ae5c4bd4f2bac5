#!/usr/bin/env python
"""How much does cutting one pinned host->device copy into row-chunks cost by itself?  (The commit's host route copies
the coefficient rows in up to 16 chunks so that the encode can trail the copy; this measures the copy alone.)

  python tools/h2d_chunks.py [MiB=512]   -> one JSON line per chunk count
"""
import json
import sys
import time

import torch

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n = mib << 20
host = torch.empty(n, dtype=torch.uint8).pin_memory()
host.random_(0, 255)
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
st = torch.cuda.Stream()
for chunks in (1, 2, 4, 8, 16, 32, 64):
    cuts = [k * n // chunks for k in range(chunks + 1)]
    wall = []
    for it in range(12):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(st):
            for k in range(chunks):
                dev[cuts[k]:cuts[k + 1]].copy_(host[cuts[k]:cuts[k + 1]], non_blocking=True)
        st.synchronize()
        wall.append((time.perf_counter() - t0) * 1e3)
    wall = sorted(wall[2:])
    print(json.dumps({"mib": mib, "chunks": chunks, "ms_median": wall[len(wall) // 2], "ms_min": wall[0],
                      "gbs": n / wall[len(wall) // 2] / 1e6}))
