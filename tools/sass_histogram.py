#!/usr/bin/env python
"""Opcode histogram of the built library's SASS (per kernel family and total): the evidence of WHAT the SMs execute --
wide integer multiply-adds for the field arithmetic, ALU ops for BLAKE3, UBLKCP / SYNCS for the bulk-copy path.

  python tools/sass_histogram.py [lcpc_b200/lib/liblcpc_b200.so] > profiles/r02_sass_histogram.txt
"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "lcpc_b200/lib/liblcpc_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn, per_fn, total = None, collections.defaultdict(collections.Counter), collections.Counter()
ins = re.compile(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)")
for line in out.splitlines():
    if "Function :" in line:
        fn = line.split("Function :")[1].strip()
        continue
    m = ins.match(line)
    if m and fn:
        op = m.group(1)
        per_fn[fn][op] += 1
        total[op] += 1


def family(name):
    r = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    r = re.sub(r"\(.*", "", r)
    return r.replace("void ", "").replace("lcpc::", "")


print(f"# {lib}: {len(per_fn)} kernels, {sum(total.values())} SASS instructions (static counts)")
print("## whole library, top 40 opcodes")
for op, c in total.most_common(40):
    print(f"{c:9d}  {op}")
print("## Blackwell / Hopper+ specific")
for op, c in sorted(total.items()):
    if op.startswith(("UBLKCP", "UTMA", "SYNCS", "UTC", "LDTM", "STTM", "LDGSTS")):
        print(f"{c:9d}  {op}")
print("## per kernel: instructions, wide multiply-adds, ALU logic, top opcodes")
for name, cnt in sorted(per_fn.items(), key=lambda kv: -sum(kv[1].values())):
    n = sum(cnt.values())
    wide = sum(c for o, c in cnt.items() if o.startswith("IMAD.WIDE"))
    alu = sum(c for o, c in cnt.items() if o.split(".")[0] in ("LOP3", "SHF", "IADD3", "PRMT", "SEL"))
    top = ", ".join(f"{o} {c}" for o, c in cnt.most_common(4))
    print(f"{n:7d}  wide {wide:5d}  alu {alu:6d}  {family(name)[:70]:70s}  {top}")
