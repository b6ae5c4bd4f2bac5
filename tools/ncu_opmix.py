#!/usr/bin/env python
"""Dynamic opcode mix of one captured kernel from an .ncu-rep source page: executed warp instructions and
stall samples per SASS opcode.  Usage: python tools/ncu_opmix.py rep.ncu-rep <kernel regex> [launch index]"""
import collections
import csv
import io
import subprocess
import sys


def main(path, regex, idx="1"):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-id", f"::regex:{regex}:{idx}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    i_src, i_ex, i_smp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    ex, smp = collections.Counter(), collections.Counter()
    for r in rows[2:]:
        if r and r[0] == "Kernel Name":  # the regex matched a further launch: keep the first only
            break
        if len(r) <= i_ex or not r[i_ex].isdigit():
            continue
        toks = r[i_src].split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") else toks[0]
        ex[op] += int(r[i_ex] or 0)
        smp[op] += int(r[i_smp] or 0)
    tot, tots = sum(ex.values()), sum(smp.values())
    print(f"{rows[0][1][:90]}\n total warp instructions {tot}, stall samples {tots}")
    for op, n in ex.most_common(24):
        print(f"{op:24s} {n:14d} {100.0 * n / tot:6.2f}%   samples {100.0 * smp[op] / max(1, tots):6.2f}%")


if __name__ == "__main__":
    main(*sys.argv[1:])
