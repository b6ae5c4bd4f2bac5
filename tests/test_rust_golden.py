"""Parity pin against the REAL reference, when its roots are available.

tools/rust_golden/ runs conroi/lcpc itself (cargo) on seeded inputs; this image has no Rust toolchain, so the
file tests/golden/rust_roots.json may not exist yet -- then the test is reported as XFAIL (an open gap, not a silent
skip) and the parity status of DESIGN.md section 7 stands: ff_derive's random / to_repr and fffft's output order are
held to published descriptions only, everything else to published vectors.  Once the file is there,
the oracle has to reproduce every root and the dims the reference chose.
"""
import importlib.util
import json
import os

import pytest

import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "rust_roots.json")


def _cases():
    spec = importlib.util.spec_from_file_location("make_inputs", os.path.join(ROOT, "tools", "rust_golden", "make_inputs.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_tool_inputs_are_canonical_and_reproducible(tmp_path):
    """the generator is deterministic and writes what F::from_repr accepts (canonical LE bytes below p)"""
    mod = _cases()
    mod.main(str(tmp_path))
    roots = json.load(open(tmp_path / "oracle_roots.json"))
    assert [r["case"] for r in roots] == [c[0] for c in mod.CASES]
    name, _, field, length, _ = mod.CASES[0]
    raw = open(tmp_path / (name + ".bin"), "rb").read()
    nb = 8 * O.FIELD_LIMBS[field]
    assert len(raw) == length * nb
    p = O.field_info(field)["modulus"]
    assert all(int.from_bytes(raw[i:i + nb], "little") < p for i in range(0, 64 * nb, nb))


def test_oracle_reproduces_reference_roots():
    if not os.path.exists(GOLD):
        pytest.xfail("tests/golden/rust_roots.json absent: no Rust toolchain was available to run the reference; parity "
                     "against the real crates stays unpinned for ff_derive's random/to_repr and fffft's output order")
    mod = _cases()
    want = {}
    for line in open(GOLD):
        line = line.strip()
        if line.startswith("{"):
            d = json.loads(line)
            want[d["case"]] = d
    assert want, "rust_roots.json holds no records"
    for name, kind, field, length, seed in mod.CASES:
        if name not in want:
            continue
        x = mod.case_coeffs(name, field, length)
        enc = O.Encoding.ligero(field, length) if kind == "ligero" else O.Encoding.sdig(field, length, seed=seed)
        assert list(enc.get_dims(length)) == [want[name][k] for k in ("n_rows", "n_per_row", "n_cols")], name
        c = enc.commit(x)
        assert c["root"].hex() == want[name]["root"], name
        if "proof_blake3" in want[name]:  # prove / verify / bincode wire image of the real reference
            got = mod.case_proof(field, enc, c, x)
            assert got == {k: want[name][k] for k in ("proof_len", "proof_blake3", "eval")}, name
