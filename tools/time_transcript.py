"""Host-only timing of the Fiat-Shamir transcript absorb (one p_random / p_eval vector of the 2^24 Ligero commit) for
each Keccak build the CPU supports.  Usage: python tools/time_transcript.py"""
import os
import subprocess
import sys

CODE = r"""
import time, numpy as np, lcpc_b200 as P
r = (np.arange(65536 * 32) % 251).astype(np.uint8).reshape(65536, 32)
best = 1e9
for _ in range(7):
    t = P.Transcript(b"x"); t0 = time.perf_counter(); t.append_reprs(b"$l//PR", r); best = min(best, time.perf_counter() - t0)
print(round(best * 1e3, 3))
"""
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for which in ("base", "bmi_table", "bmi", "avx512"):
    env = dict(os.environ, LCPC_B200_KECCAK=which, PYTHONPATH=root)
    out = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True)
    print(which, out.stdout.strip() or out.stderr.strip()[-200:], "ms per 65536 x 32-byte absorbs", flush=True)
flags = [l for l in open("/proc/cpuinfo") if l.startswith("flags")][0].split()
print("cpu:", [l.split(":")[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0],
      "| avx512f" if "avx512f" in flags else "| no avx512f", "| bmi2" if "bmi2" in flags else "| no bmi2")
