#!/usr/bin/env python
"""Generate tests/golden/expander_vectors.npz from the reference's own Python spec.

The reference ships a Python design-time statement of the Brakedown expander code,
/root/reference/doc/encoding.py (recursive form: encode(x) = x || encode(x*precode) || z*postcode,
Reed-Solomon base case at the points 1..m).  This script imports that file's functions in THIS
container (the file is executed up to its "# example" driver, which would otherwise run a 2^13
demo at import), feeds them the sparse matrices produced for the Rust dimension rules
(lcpc-brakedown-pc/src/matgen.rs:56-111) and records input -> codeword pairs.  The fixture pins,
independently of our C restatement:
  * the recursion / flat codeword layout of lcpc-brakedown-pc/src/encode.rs:36-94,
  * the orientation of the sparse product (doc/encoding.py `multiply`: y[j] += x[i]*A[i][j]
    == sprs CSC `M.dot(x)` with M[j,i] = A[i][j]),
  * the Vandermonde base code of encode.rs:97-110.
It cannot travel to the GPU box (no /root/reference there): the .npz is committed.

Run from the repo root:  python tests/golden/make_expander_vectors.py
"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle as O  # noqa: E402  (only for matgen dims + matrices + Montgomery conversion)

REF = "/root/reference/doc/encoding.py"


def load_reference_spec():
    src = open(REF).read()
    cut = src.index("# example")
    ns = {}
    exec(compile(src[:cut], REF, "exec"), ns)
    return ns


def main():
    ref = load_reference_spec()
    SparseMatrix = ref["SparseMatrix"]
    out = {}
    cases = [("ft63_n400", O.FT63, 400, 11), ("ft127_n150", O.FT127, 150, 12), ("ft255_n64", O.FT255, 64, 13)]
    for name, field, n, seed in cases:
        info = O.field_info(field)
        p = info["modulus"]
        enc = O.Encoding.sdig_from_dims(field, n, 0, seed=seed, code=3)
        pre, post = enc.matrices()
        # reference globals: field size and the rate of SdigCode3 (codespec.rs:191-199)
        ref["p"] = p
        ref["r"] = 1.521

        def to_ref(M):
            # CSC column j of the Rust matrix (m x n) == row j of the Python matrix (n x m)
            S = SparseMatrix(M["n"], M["m"])
            vals = O.from_mont(field, M["data"])
            for j in range(M["n"]):
                for k in range(int(M["ptrs"][j]), int(M["ptrs"][j + 1])):
                    S.add(j, int(M["idxs"][k]), vals[k])
            return S

        code = ([to_ref(M) for M in pre], [to_ref(M) for M in post])
        x_mont = O.random_elems(field, n, seed=seed + 100)
        x = O.from_mont(field, x_mont)
        # the spec's base case triggers on len <= 20 and emits ceil(r*len) symbols; the Rust dims use
        # ceil_muldiv(len, 1521, 1000) (matgen.rs:96) -- assert the two agree for this case
        last_m = pre[-1]["m"]
        assert last_m <= 20 and math.ceil(1.521 * last_m) == post[-1]["n"], (last_m, post[-1]["n"])
        cw = [v % p for v in ref["encode"](x, code)]
        assert len(cw) == enc.n_cols, (len(cw), enc.n_cols)
        out[name + "_field"] = np.array([field])
        out[name + "_n_levels"] = np.array([len(pre)])
        out[name + "_input_mont"] = x_mont
        out[name + "_codeword_mont"] = O.to_mont(field, cw)
        for tag, mats in (("pre", pre), ("post", post)):
            for i, M in enumerate(mats):
                out[f"{name}_{tag}{i}_shape"] = np.array([M["m"], M["n"]], dtype=np.uint64)
                out[f"{name}_{tag}{i}_ptrs"] = M["ptrs"]
                out[f"{name}_{tag}{i}_idxs"] = M["idxs"].astype(np.uint32)
                out[f"{name}_{tag}{i}_data"] = M["data"]
        # cross-check our restatement while we are here
        row = np.zeros((enc.n_cols, enc.L), np.uint64)
        row[:n] = x_mont
        assert (enc.encode(row) == out[name + "_codeword_mont"]).all(), name
        print(name, "levels", len(pre), "codeword", len(cw), "ok")
    path = os.path.join(HERE, "expander_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
