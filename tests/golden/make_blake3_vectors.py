#!/usr/bin/env python
"""Generate tests/golden/blake3_vectors.json with the Python `blake3` package.

`blake3` (PyPI 1.0.8, present in this image) is a binding of the Rust `blake3` crate -- the same
crate the reference hashes with (`blake3 = "1"`, traits-preview; lcpc-2d/src/tests.rs:12,
lcpc-ligero-pc/src/bench.rs:12).  The vectors pin, for both the oracle and the CUDA kernels:
  * raw BLAKE3 over the lengths the column hash meets (multi-chunk tree included),
  * the leaf rule  leaf = D(0^32 || repr(col[0]) || ... )   lcpc-2d/src/lib.rs:719-735,
  * the node rule  node = D(left || right)                 lcpc-2d/src/lib.rs:770-775.
Run from the repo root:  python tests/golden/make_blake3_vectors.py
"""
import json
import os

import blake3

HERE = os.path.dirname(os.path.abspath(__file__))


def pattern(n):
    # the input pattern of the official BLAKE3 test vectors: byte i = i mod 251
    return bytes(i % 251 for i in range(n))


def main():
    out = {"_generator": "tests/golden/make_blake3_vectors.py", "_blake3_py": blake3.__version__}
    lengths = [0, 1, 2, 3, 31, 32, 33, 63, 64, 65, 96, 127, 128, 129, 320, 1023, 1024, 1025, 1184,
               2048, 2049, 2080, 3072, 3073, 4096, 4097, 4608, 5120, 5121, 6144, 6145, 7168, 7169,
               8192, 8193, 8224, 16384, 31744, 32800, 102400]
    out["raw"] = [{"len": n, "hash": blake3.blake3(pattern(n)).hexdigest()} for n in lengths]
    # leaf rule on canonical little-endian reprs: column values 1..n_rows, element width B bytes
    leaves = []
    for B in (8, 16, 24, 32):
        for n_rows in (1, 2, 3, 18, 31, 32, 64, 72, 256, 286, 1024):
            data = bytes(32) + b"".join((r + 1).to_bytes(B, "little") for r in range(n_rows))
            leaves.append({"elem_bytes": B, "n_rows": n_rows, "hash": blake3.blake3(data).hexdigest()})
    out["leaf_of_1_to_n"] = leaves
    z = bytes(32)
    out["node_zero_zero"] = blake3.blake3(z + z).hexdigest()
    a = blake3.blake3(b"left").digest()
    b = blake3.blake3(b"right").digest()
    out["node_left_right"] = {"left": a.hex(), "right": b.hex(), "hash": blake3.blake3(a + b).hexdigest()}
    with open(os.path.join(HERE, "blake3_vectors.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote blake3_vectors.json:", len(out["raw"]), "raw,", len(leaves), "leaf vectors")


if __name__ == "__main__":
    main()
