#!/usr/bin/env python
"""tools/ab_sweep.py -- A/B sweep of schedule knobs (lcpc_b200_set_tunable) on one device-resident commit.

  python tools/ab_sweep.py brakedown [--lgl 24] [--steps 10] -- KNOB=v[,v..] KNOB=v[,v..] ...
  python tools/ab_sweep.py ligero ...

Every combination of the listed knob values is timed with CUDA events on the engine stream (the library's own
phase events: pad/copy, encode, leaf hash, merkle) over `steps` commits after 2 warm-ups; the LcRoot of every
combination must equal the first one's (knobs change schedules, never results).  One JSON line per combination.
"""
import argparse
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", choices=["ligero", "brakedown"])
    ap.add_argument("--lgl", type=int, default=24)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--rows", type=int, default=0,
                    help="commit only this many rows under the 2^lgl encoding (the share of one rank of a sharded commit)")
    ap.add_argument("--per-row", type=int, default=0,
                    help="ligero: an encoding of this many coefficients per row (rho 1/2) instead of the 2^lgl one, with "
                         "--rows rows: e.g. --per-row 8192 --rows 256 has the leaf-hash shape of one rank's column block at 8 GPUs")
    ap.add_argument("knobs", nargs="*")
    args = ap.parse_args()
    import torch

    import bench as B
    import lcpc_b200 as P
    from lcpc_b200 import _cabi

    lib = _cabi.lib()
    field = 4 if args.workload == "ligero" else 2
    n = 1 << args.lgl
    ctx = P.Context(0)
    enc = P.LigeroEncoding(field, n, ctx=ctx) if args.workload == "ligero" else P.SdigEncoding(field, n, seed=0, ctx=ctx)
    if args.per_row:
        enc = P.LigeroEncoding.new_from_dims(field, args.per_row, 2 * args.per_row, ctx=ctx)
    if args.rows:
        n = args.rows * enc.n_per_row
    x = B.synthetic_coeffs(field, n, seed=0)
    dev = torch.from_numpy(x.view(np.int64)).cuda()
    names, values = [], []
    for k in args.knobs:
        name, vs = k.split("=")
        names.append(name)
        values.append([int(v) for v in vs.split(",")])
    try:
        from cuda.bindings import runtime as rt
        attrs = {}
        for nm in ("cudaDevAttrL2CacheSize", "cudaDevAttrMaxPersistingL2CacheSize", "cudaDevAttrMaxAccessPolicyWindowSize"):
            err, v = rt.cudaDeviceGetAttribute(getattr(rt.cudaDeviceAttr, nm), 0)
            attrs[nm] = int(v)
        err, lim = rt.cudaDeviceGetLimit(rt.cudaLimit.cudaLimitPersistingL2CacheSize)
        attrs["cudaLimitPersistingL2CacheSize"] = int(lim)
        print(json.dumps({"device_attrs": attrs}), flush=True)
    except Exception as e:  # informational only
        print(json.dumps({"device_attrs_error": repr(e)}), flush=True)
    commit = P.LcCommit.commit_device(dev.data_ptr(), n, enc)
    root0 = None
    for combo in itertools.product(*values) if names else [()]:
        for name, v in zip(names, combo):
            lib.lcpc_b200_set_tunable(name.encode(), v)
        for _ in range(2):
            commit.rerun_device(dev.data_ptr(), n)
        ctx.synchronize()
        ph = np.zeros(4)
        for _ in range(args.steps):
            commit.rerun_device(dev.data_ptr(), n)
            ms, nl = commit.phase_times()
            ph += np.array(ms)
        ph /= args.steps
        root = commit.get_root().root.hex()
        root0 = root0 or root
        print(json.dumps({"workload": args.workload, "lgl": args.lgl, "knobs": dict(zip(names, combo)),
                          "encode_ms": round(float(ph[1]), 4), "leaf_ms": round(float(ph[2]), 4),
                          "merkle_ms": round(float(ph[3]), 4), "commit_ms": round(float(ph.sum()), 4),
                          "launches": nl, "root_same": root == root0}), flush=True)
        assert root == root0, "a schedule knob changed the LcRoot"


if __name__ == "__main__":
    main()
