#!/usr/bin/env python
"""Where a kernel's warp-stall samples fall, read from an .ncu-rep without a GPU (ncu --page source --csv):
samples grouped by the execution count of the SASS instructions (= by loop nest) and the hottest instructions.

  python tools/ncu_hot.py gpurun_out/x.ncu-rep KERNEL_REGEX [launch_index=0]
"""
import collections
import csv
import io
import subprocess
import sys


def main(path, kernel, which=0):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", f"regex:{kernel}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    if not starts:
        sys.exit("no such kernel in the report")
    s = starts[min(which, len(starts) - 1)]
    e = starts[starts.index(s) + 1] if starts.index(s) + 1 < len(starts) else len(rows)
    print(rows[s][1][:150])
    h = rows[s + 1]
    si, src, ie = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
    data = [r for r in rows[s + 2:e] if len(r) > max(si, ie) and r[si].isdigit()]
    tot = sum(int(r[si]) for r in data) or 1
    groups = collections.OrderedDict()
    for k, r in enumerate(data):
        g = groups.setdefault(r[ie], [0, 0, k, k])
        g[0] += int(r[si]); g[1] += 1; g[3] = k
    print(f"{tot} samples over {len(data)} SASS instructions; by execution count (loop nest):")
    for key, (n, cnt, a, b) in groups.items():
        if n * 100 >= tot:
            print(f"  executed {key:>12}: {100 * n / tot:5.1f} %  ({cnt} instructions, SASS index {a}..{b})")
    print("hottest instructions:")
    for r in sorted(data, key=lambda r: -int(r[si]))[:14]:
        print(f"  {100 * int(r[si]) / tot:5.1f} %  {r[src].strip()[:100]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
