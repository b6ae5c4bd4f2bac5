#!/usr/bin/env python
"""Roofline sweep (BASELINE.json config 5): bench.py over 2^20..2^28 coefficients for both encodings, at one GPU
count per invocation.  Appends one JSON line per point to gpurun_out/sweep_g<N>.jsonl.

  python tools/sweep.py --gpus 1 [--lgls 20 22 24 26 28] [--workloads ligero brakedown] [--steps 10]
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--lgls", type=int, nargs="+", default=[20, 22, 24, 26, 28])
    ap.add_argument("--workloads", nargs="+", default=["ligero", "brakedown"])
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cpu-baseline", action="store_true")
    a = ap.parse_args()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out = os.path.join(ROOT, "gpurun_out", f"sweep_g{a.gpus}.jsonl")
    for wl in a.workloads:
        for lgl in a.lgls:
            cmd = [sys.executable]
            if a.gpus > 1:
                cmd += ["-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}", "--master-addr",
                        "127.0.0.1", "--master-port", str(29600 + lgl)]
            cmd += [os.path.join(ROOT, "bench.py"), "--gpus", str(a.gpus), "--steps", str(a.steps), "--warmup",
                    str(a.warmup), "--workload", wl, "--lgl", str(lgl)]
            if not a.cpu_baseline:
                cmd.append("--no-cpu-baseline")
            r = subprocess.run(cmd, capture_output=True, text=True)
            lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
            rec = json.loads(lines[-1]) if lines else {"error": r.stderr[-800:], "workload": wl, "lgl": lgl}
            rec["lgl"] = lgl
            with open(out, "a") as f:
                f.write(json.dumps(rec) + "\n")
            print(wl, lgl, rec.get("value"), rec.get("ms_per_step"), (rec.get("e2e") or {}).get("ms_per_step"),
                  rec.get("error", "")[:300], flush=True)


if __name__ == "__main__":
    main()
