#!/usr/bin/env python
"""Write the seeded inputs of tools/rust_golden (canonical little-endian bytes, what F::from_repr parses) and
the roots THIS repository's oracle computes for them.

  python tools/rust_golden/make_inputs.py <out dir>

<out dir>/<case>.bin are the coefficient files for the Rust program; <out dir>/oracle_roots.json holds our
roots.  After running the Rust program on another machine, save its lines as tests/golden/rust_roots.json
(one JSON object per line): tests/test_rust_golden.py compares the two and turns "parity unpinned" into a
checked fact -- or shows exactly which convention (field repr, NTT order / root of unity, matgen stream)
differs.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402

CASES = [  # name, kind, field, length, seed of the expander code
    ("ligero_ft255_2_10", "ligero", O.FT255, 1 << 10, None),
    ("ligero_ft255_2_14", "ligero", O.FT255, 1 << 14, None),
    ("ligero_ft127_2_12", "ligero", O.FT127, 1 << 12, None),
    ("ligero_ft63_1000", "ligero", O.FT63, 1000, None),
    ("brakedown_ft127_2_12_seed0", "sdig", O.FT127, 1 << 12, 0),
    ("brakedown_ft255_3000_seed1", "sdig", O.FT255, 3000, 1),
    ("brakedown_ft63_2_13_seed7", "sdig", O.FT63, 1 << 13, 7),
]


def case_coeffs(name, field, length):
    return O.random_elems(field, length, seed=sum(name.encode()) % 997)


def case_proof(field, enc, c, x):
    """What the Rust program prints about prove/verify: the wire image of the proof made on
    Transcript::new(b"rust golden") with outer = first n_rows coefficients, and the evaluation verify() returns with
    inner = first n_per_row coefficients."""
    from oracle import protocol as PR
    from oracle.transcript import Transcript
    outer, inner = x[:c["n_rows"]], x[:c["n_per_row"]]
    proof = PR.prove(field, c, outer, enc.get_n_degree_tests(), enc.get_n_col_opens(), Transcript(b"rust golden"))
    wire = PR.wire_proof(proof)
    ev = PR.verify(field, enc, c["root"], outer, inner, proof, Transcript(b"rust golden"))
    return {"proof_len": len(wire), "proof_blake3": O.blake3(wire).hex(),
            "eval": O.to_repr(field, ev.reshape(1, -1)).tobytes().hex()}


def main(out):
    os.makedirs(out, exist_ok=True)
    roots = []
    for name, kind, field, length, seed in CASES:
        x = case_coeffs(name, field, length)
        open(os.path.join(out, name + ".bin"), "wb").write(O.to_repr(field, x).tobytes())
        enc = O.Encoding.ligero(field, length) if kind == "ligero" else O.Encoding.sdig(field, length, seed=seed)
        c = enc.commit(x)
        n_rows, n_per_row, n_cols = enc.get_dims(length)
        rec = {"case": name, "root": c["root"].hex(), "n_rows": n_rows, "n_per_row": n_per_row, "n_cols": n_cols}
        rec.update(case_proof(field, enc, c, x))
        roots.append(rec)
    json.dump(roots, open(os.path.join(out, "oracle_roots.json"), "w"), indent=1)
    print(f"wrote {len(CASES)} input files and oracle_roots.json to {out}")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "rust_golden_inputs")
