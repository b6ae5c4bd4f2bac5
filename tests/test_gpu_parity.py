"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle and the golden
fixtures.  Bit-exact everywhere (integer prime-field arithmetic and hashing; no tolerance).

Sizes: seeded inputs at sizes the oracle finishes in seconds; BASELINE.json's full sizes
(2^20 / 2^24) are covered in test_gpu_fullsize.py through whole-row, sampled-column and Merkle-tree
checks.
"""
import json
import os

import numpy as np
import pytest

import oracle as O
import lcpc_b200 as P
from lcpc_b200 import _cabi

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIELDS = [P.FT63, P.FT127, P.FT191, P.FT255]


def rows_from(field, n_rows, n_cols, seed):
    return O.random_elems(field, n_rows * n_cols, seed=seed)


# ------------------------------------------------------------------ field arithmetic
@pytest.mark.parametrize("field", FIELDS)
def test_field_ops(field):
    n = 100_000
    a, b = O.random_elems(field, n, seed=11), O.random_elems(field, n, seed=12)
    p = O.field_info(field)["modulus"]
    nl = O.FIELD_LIMBS[field]
    edge = O.ints_to_elems([0, 1, p - 1, p - 2, (1 << (64 * nl - 1)) % p, 0xffffffff, 1 << 32, p - 0xffffffff], field)
    a[:8], b[:8] = edge, edge[::-1]
    a[8:16], b[8:16] = edge, edge
    for op in ("add", "sub", "mul"):
        assert (P.field_op(field, op, a, b) == O.field_op(field, op, a, b)).all(), op
    assert (P.field_op(field, "from_mont", a) == O.field_op(field, "from_mont", a)).all()
    # full product + separate reduction, and double-width sums of 37 products reduced once
    assert (P.field_op(field, "mul_sos", a, b) == O.field_op(field, "mul", a, b)).all()
    assert (P.field_op(field, "mul_karatsuba", a, b) == O.field_op(field, "mul", a, b)).all()
    m = 5000
    a2, b2 = a[:m].copy(), b[:m].copy()
    a2[100:164] = O.ints_to_elems([p - 1] * 64, field)   # runs of (p-1)^2 terms: worst case of the fold bound
    b2[700:1200] = O.ints_to_elems([p - 1] * 500, field)
    idx = np.arange(m)
    want = O.ints_to_elems([0] * m, field)
    for k in range(37):
        want = O.field_op(field, "add", want, O.field_op(field, "mul", a2[(idx + k) % m], b2[(idx * 7 + k) % m]))
    assert (P.field_op(field, "lazy_sum37", a2, b2) == want).all()


# ------------------------------------------------------------------ hashing / Merkle
def test_leaf_rule_golden_vectors():
    """leaf = D(0^32 || repr(col[0]) || ...) against the fixtures made with the Rust crate's binding."""
    g = json.load(open(os.path.join(GOLD, "blake3_vectors.json")))
    fld = {8: P.FT63, 16: P.FT127, 24: P.FT191, 32: P.FT255}
    for v in g["leaf_of_1_to_n"]:
        field, n_rows = fld[v["elem_bytes"]], v["n_rows"]
        col = O.to_mont(field, list(range(1, n_rows + 1)))
        comm = np.repeat(col[:, None, :], 3, axis=1).reshape(-1, O.FIELD_LIMBS[field])
        h = P.merkleize(field, comm, n_rows, 3)
        assert h.shape == (7, 32)
        for c in range(3):
            assert h[c].tobytes().hex() == v["hash"], (v["elem_bytes"], n_rows)
        assert not h[3].any()  # padding leaf stays Output::default()
        assert h[4].tobytes() == O.blake3(h[0].tobytes() + h[1].tobytes())
        assert h[5].tobytes().hex() != g["node_zero_zero"]
        assert h[6].tobytes() == O.blake3(h[4].tobytes() + h[5].tobytes())


def test_node_rule_golden():
    g = json.load(open(os.path.join(GOLD, "blake3_vectors.json")))
    # 1 real column + 3 padding leaves: node over two zero leaves must be the golden constant
    comm = O.to_mont(P.FT63, [5])
    h = P.merkleize(P.FT63, np.tile(comm, (5, 1)), 1, 5)  # 5 columns -> np2 = 8
    assert h[8 + 3].tobytes().hex() == g["node_zero_zero"]  # layer-1 node over leaves 6,7


@pytest.mark.parametrize("field,n_rows,n_cols", [
    (P.FT255, 1, 1), (P.FT255, 2, 1024), (P.FT255, 31, 100), (P.FT255, 32, 257), (P.FT255, 33, 64),
    (P.FT255, 64, 300), (P.FT255, 256, 96), (P.FT255, 1024, 40),
    (P.FT127, 1, 7), (P.FT127, 18, 1000), (P.FT127, 62, 129), (P.FT127, 63, 128), (P.FT127, 72, 513),
    (P.FT127, 286, 70),
    (P.FT63, 3, 50), (P.FT63, 124, 33), (P.FT63, 125, 32), (P.FT63, 500, 20),
    (P.FT191, 1, 5), (P.FT191, 41, 77), (P.FT191, 42, 64), (P.FT191, 200, 31),
])
def test_merkleize_vs_oracle(field, n_rows, n_cols):
    comm = rows_from(field, n_rows, n_cols, seed=n_rows * 1000 + n_cols)
    got = P.merkleize(field, comm, n_rows, n_cols)
    want = O.merkleize(field, comm, n_rows, n_cols)
    assert (got == want).all()
    assert (got == O.merkleize(field, comm, n_rows, n_cols, serial=True)).all()  # lcpc-2d tests.rs:136-149


# ------------------------------------------------------------------ Ligero encode
@pytest.mark.parametrize("field", FIELDS)
@pytest.mark.parametrize("log_n", [1, 2, 3, 5, 8, 10, 11, 12, 13, 15])
def test_ligero_encode_vs_oracle(field, log_n):
    n_cols = 1 << log_n
    n_per_row = n_cols // 2
    enc = P.LigeroEncoding.new_from_dims(field, n_per_row, n_cols)
    oenc = O.Encoding.ligero_from_dims(field, n_per_row, n_cols)
    n_rows = 3 if log_n < 13 else 2
    rows = np.zeros((n_rows, n_cols, enc.L), np.uint64)
    rows[:, :n_per_row] = O.random_elems(field, n_rows * n_per_row, seed=log_n).reshape(n_rows, n_per_row, -1)
    got = enc.encode(rows)
    for r in range(n_rows):
        assert (got[r] == oenc.encode(rows[r])).all(), r
    # the whole row is input (verifier use, lcpc-2d/src/lib.rs:886): a dense row must work too
    dense = O.random_elems(field, n_cols, seed=99)
    assert (enc.encode(dense) == oenc.encode(dense)).all()


@pytest.mark.parametrize("field,log_n", [(P.FT255, 19), (P.FT63, 20), (P.FT127, 21)])
def test_ligero_encode_long_rows(field, log_n):
    """row lengths of the 2^26..2^28 sweep points: 2^19 and 2^20 points are two passes of 10 + 9 / 10 + 10
    stages, 2^21 three passes"""
    n_cols = 1 << log_n
    n_per_row = n_cols // 2
    enc = P.LigeroEncoding.new_from_dims(field, n_per_row, n_cols)
    oenc = O.Encoding.ligero_from_dims(field, n_per_row, n_cols)
    row = np.zeros((n_cols, enc.L), np.uint64)
    row[:n_per_row] = O.random_elems(field, n_per_row, seed=log_n)
    assert (enc.encode(row) == oenc.encode(row)).all()


@pytest.mark.parametrize("rho,n_cols", [((1, 4), 1 << 12), ((39, 40), 1 << 13), ((1, 2), 1 << 16)])
def test_ligero_commit_other_rates(rho, n_cols):
    field = P.FT127
    n_per_row = n_cols * rho[0] // rho[1]
    enc = P.LigeroEncoding.new_from_dims(field, n_per_row, n_cols, rho=rho)
    oenc = O.Encoding.ligero_from_dims(field, n_per_row, n_cols, rho=rho)
    length = 3 * n_per_row - 17
    x = O.random_elems(field, length, seed=5)
    c, oc = P.LcCommit.commit(x, enc), oenc.commit(x)
    assert c.get_root().root == oc["root"]
    assert (c.comm == oc["comm"]).all() and (c.coeffs == oc["coeffs"]).all() and (c.hashes == oc["hashes"]).all()


def test_ligero_errors():
    with pytest.raises(P.LcpcError) as e:
        P.LigeroEncoding.new_from_dims(P.FT255, 8, 12)  # not a power of two: dims_ok fails
    assert e.value.code == _cabi.ERR_BAD_ARG
    with pytest.raises(P.LcpcError):
        P.LigeroEncoding.new_from_dims(P.FT255, 16, 16)  # n_per_row must be < n_cols
    enc = P.LigeroEncoding.new_from_dims(P.FT63, 8, 16)
    with pytest.raises(P.LcpcError):
        enc.encode(np.zeros((15, 1), np.uint64))
    with pytest.raises(P.LcpcError):
        P.LcCommit.commit(np.zeros((0, 1), np.uint64), enc)
    assert enc.dims_ok(8, 16) and not enc.dims_ok(4, 16) and not enc.dims_ok(8, 32)
    assert enc.get_dims(17) == (3, 8, 16)


# ------------------------------------------------------------------ Ligero commit (config 1 of BASELINE.json)
def test_commit_ligero_ft255_2_10_bit_exact_root():
    """lcpc-ligero-pc commit, Ft255, 2^10 coefficients: every LcCommit field and the LcRoot."""
    field, length = P.FT255, 1 << 10
    enc = P.LigeroEncoding(field, length)
    oenc = O.Encoding.ligero(field, length)
    assert (enc.n_per_row, enc.n_cols) == (oenc.n_per_row, oenc.n_cols) == (512, 1024)
    assert enc.get_n_col_opens() == oenc.get_n_col_opens() == 309
    assert enc.get_n_degree_tests() == oenc.get_n_degree_tests() == 1
    for seed in (0, 1, 2):
        x = O.random_elems(field, length, seed=seed)
        c, oc = P.LcCommit.commit(x, enc), oenc.commit(x)
        assert (c.n_rows, c.n_per_row, c.n_cols) == (2, 512, 1024)
        assert c.get_root() == P.LcRoot(oc["root"])
        assert (c.coeffs == oc["coeffs"]).all()
        assert (c.comm == oc["comm"]).all()
        assert (c.hashes == oc["hashes"]).all()


@pytest.mark.parametrize("field,lgl", [(P.FT255, 14), (P.FT255, 16), (P.FT127, 15), (P.FT63, 13), (P.FT191, 12)])
def test_commit_ligero_vs_oracle(field, lgl):
    length = (1 << lgl) - 3  # ragged: last row zero-padded (lcpc-2d/src/lib.rs:640-645)
    enc = P.LigeroEncoding(field, length)
    oenc = O.Encoding.ligero(field, length)
    x = O.random_elems(field, length, seed=lgl)
    c, oc = P.LcCommit.commit(x, enc), oenc.commit(x)
    assert c.get_root().root == oc["root"]
    assert (c.comm == oc["comm"]).all() and (c.coeffs == oc["coeffs"]).all() and (c.hashes == oc["hashes"]).all()
    # the one-shot host form fills caller-owned arrays identically
    comm, coeffs, hashes = (np.empty_like(oc[k]) for k in ("comm", "coeffs", "hashes"))
    rc = _cabi.lib().lcpc_b200_commit_to_host(enc._h, x.ctypes.data, length, comm.ctypes.data, coeffs.ctypes.data,
                                             hashes.ctypes.data)
    assert rc == 0 and (comm == oc["comm"]).all() and (coeffs == oc["coeffs"]).all() and (hashes == oc["hashes"]).all()


# ------------------------------------------------------------------ Brakedown
def _npz_case(z, name):
    field = int(z[name + "_field"][0])
    nlev = int(z[name + "_n_levels"][0])

    def mats(tag):
        out = []
        for i in range(nlev):
            m, n = (int(v) for v in z[f"{name}_{tag}{i}_shape"])
            out.append(dict(m=m, n=n, ptrs=z[f"{name}_{tag}{i}_ptrs"], idxs=z[f"{name}_{tag}{i}_idxs"].astype(np.uint64),
                            data=z[f"{name}_{tag}{i}_data"]))
        return out
    return field, mats("pre"), mats("post"), z[name + "_input_mont"], z[name + "_codeword_mont"]


@pytest.mark.parametrize("name", ["ft63_n400", "ft127_n150", "ft255_n64"])
def test_expander_encode_vs_reference_python_spec(name):
    """Codewords produced by the reference's own doc/encoding.py (tests/golden/make_expander_vectors.py)."""
    z = np.load(os.path.join(GOLD, "expander_vectors.npz"))
    field, pre, post, x, want = _npz_case(z, name)
    enc = P.SdigEncoding.from_matrices(field, pre, post)
    assert enc.n_per_row == x.shape[0] and enc.n_cols == want.shape[0]
    row = np.zeros((enc.n_cols, enc.L), np.uint64)
    row[:x.shape[0]] = x
    assert (enc.encode(row) == want).all()
    # batch of identical and of different rows
    rows = np.stack([row, np.zeros_like(row), row])
    got = enc.encode(rows)
    assert (got[0] == want).all() and not got[1].any() and (got[2] == want).all()


@pytest.mark.parametrize("field,n,seed", [(P.FT63, 256, 0), (P.FT127, 1500, 1), (P.FT255, 4351, 2), (P.FT191, 333, 3)])
def test_sdig_encode_vs_oracle(field, n, seed):
    enc = P.SdigEncoding.new_from_dims(field, n, seed=seed)
    oenc = O.Encoding.sdig_from_dims(field, n, seed=seed)
    assert enc.n_cols == oenc.n_cols
    n_rows = 5
    rows = np.zeros((n_rows, enc.n_cols, enc.L), np.uint64)
    rows[:, :n] = O.random_elems(field, n_rows * n, seed=seed + 50).reshape(n_rows, n, -1)
    rows[:, n:] = 7  # junk beyond n_per_row must be ignored and overwritten (encode.rs:46-90)
    got = enc.encode(rows)
    for r in range(n_rows):
        assert (got[r] == oenc.encode(rows[r])).all()


@pytest.mark.parametrize("hints,window_kb,slice_kb,pipe", [(1, 8, 0, 0), (0, 8, 0, 0), (1, 64, 0, 0), (1, 0, 0, 0), (0, 0, 16, 0),
                                                          (1, 1 << 20, 0, 0), (0, 0, 0, 1), (0, 8, 0, 1), (0, 64, 0, 1), (0, 0, 16, 1),
                                                          (0, 0, 0, 2)])  # pipe == 2: the bulk-copy (TMA) gather kernel
@pytest.mark.parametrize("field,n,seed", [(P.FT127, 4000, 4), (P.FT255, 1500, 5), (P.FT63, 9000, 6), (P.FT191, 700, 7)])
def test_sdig_encode_schedules_are_result_neutral(field, n, seed, hints, window_kb, slice_kb, pipe):
    """The sparse products' schedule knobs (L2 eviction hints, column chunks with accumulation onto y, batch-row
    slices) must not change a single limb: windows of 8 KB force up to 16 column chunks on these small codes."""
    from lcpc_b200 import _cabi
    lib = _cabi.lib()
    knobs = {b"SPMM_HINTS": hints, b"SPMM_WINDOW_KB": window_kb, b"SPMM_SLICE_KB": slice_kb, b"SPMM_PIPE": int(pipe == 1),
             b"SPMM_BULK": int(pipe == 2)}
    try:
        for k, v in knobs.items():
            lib.lcpc_b200_set_tunable(k, v)
        enc = P.SdigEncoding.new_from_dims(field, n, seed=seed)
        oenc = O.Encoding.sdig_from_dims(field, n, seed=seed)
        n_rows = 7
        rows = np.zeros((n_rows, enc.n_cols, enc.L), np.uint64)
        rows[:, :n] = O.random_elems(field, n_rows * n, seed=seed + 50).reshape(n_rows, n, -1)
        got = enc.encode(rows)
        again = enc.encode(rows)  # the cached cut points are reused
        for r in range(n_rows):
            want = oenc.encode(rows[r])
            assert (got[r] == want).all() and (again[r] == want).all()
    finally:
        lib.lcpc_b200_set_tunable(b"SPMM_HINTS", 0)
        lib.lcpc_b200_set_tunable(b"SPMM_WINDOW_KB", 0)
        lib.lcpc_b200_set_tunable(b"SPMM_SLICE_KB", 0)
        lib.lcpc_b200_set_tunable(b"SPMM_PIPE", 0)
        lib.lcpc_b200_set_tunable(b"SPMM_BULK", 0)


@pytest.mark.parametrize("field,length,seed", [(P.FT127, 1 << 14, 0), (P.FT127, (1 << 16) - 11, 1), (P.FT255, 1 << 13, 0),
                                               (P.FT63, 5000, 1)])
def test_commit_brakedown_vs_oracle(field, length, seed):
    enc = P.SdigEncoding(field, length, seed=seed)
    oenc = O.Encoding.sdig(field, length, seed=seed)
    assert (enc.n_per_row, enc.n_cols) == (oenc.n_per_row, oenc.n_cols)
    assert enc.get_n_col_opens() == oenc.get_n_col_opens() == 6593
    assert enc.get_n_degree_tests() == oenc.get_n_degree_tests()
    x = O.random_elems(field, length, seed=seed + 7)
    c, oc = P.LcCommit.commit(x, enc), oenc.commit(x)
    assert c.get_root().root == oc["root"]
    assert (c.comm == oc["comm"]).all() and (c.coeffs == oc["coeffs"]).all() and (c.hashes == oc["hashes"]).all()
    np2 = 1 << (c.n_cols - 1).bit_length()
    assert not c.hashes[c.n_cols:np2].any()


def test_sdig_rejects_inconsistent_matrices():
    pre, post, _ = P.host.generate_sdig_code(P.FT63, 300, 0)
    bad = [dict(m) for m in post]
    bad[0] = dict(bad[0], n=bad[0]["n"] + 1, ptrs=np.append(bad[0]["ptrs"], bad[0]["ptrs"][-1]))
    with pytest.raises(P.LcpcError):
        P.SdigEncoding.from_matrices(P.FT63, pre, bad)
    bad2 = [dict(m) for m in pre]
    idx = bad2[0]["idxs"].copy()
    idx[0] = bad2[0]["m"]  # row index out of range
    bad2[0] = dict(bad2[0], idxs=idx)
    with pytest.raises(P.LcpcError):
        P.SdigEncoding.from_matrices(P.FT63, bad2, post)


# ------------------------------------------------------------------ prove pieces
@pytest.mark.parametrize("field,n_rows,n_per_row", [(P.FT255, 2, 512), (P.FT255, 64, 1000), (P.FT127, 7, 100),
                                                    (P.FT127, 72, 4099), (P.FT63, 300, 33), (P.FT191, 9, 65), (P.FT255, 1, 1)])
def test_collapse_vs_oracle(field, n_rows, n_per_row):
    coeffs = O.random_elems(field, n_rows * n_per_row, seed=n_rows)
    tensor = O.random_elems(field, n_rows, seed=n_per_row)
    got = P.collapse_columns(field, coeffs, tensor, n_rows, n_per_row)
    assert (got == O.collapse(field, coeffs, tensor, n_rows, n_per_row)).all()
    assert (got == O.collapse(field, coeffs, tensor, n_rows, n_per_row, serial=True)).all()  # tests.rs:151-165


def test_collapse_edge_tensors():
    field, n_rows, n_per_row = P.FT255, 16, 200
    p = O.field_info(field)["modulus"]
    coeffs = O.to_mont(field, [p - 1] * (n_rows * n_per_row))
    tensor = O.to_mont(field, [p - 1] * n_rows)
    assert (P.collapse_columns(field, coeffs, tensor, n_rows, n_per_row) == O.collapse(field, coeffs, tensor, n_rows, n_per_row)).all()
    zero = np.zeros((n_rows, 4), np.uint64)
    assert not P.collapse_columns(field, coeffs, zero, n_rows, n_per_row).any()


def test_commit_collapse_and_open_columns():
    field, length = P.FT255, 1 << 14
    enc = P.LigeroEncoding(field, length)
    oenc = O.Encoding.ligero(field, length)
    x = O.random_elems(field, length, seed=3)
    c, oc = P.LcCommit.commit(x, enc), oenc.commit(x)
    tensor = O.random_elems(field, c.n_rows, seed=4)
    assert (c.collapse(tensor) == O.collapse(field, oc["coeffs"], tensor, c.n_rows, c.n_per_row)).all()
    with pytest.raises(P.LcpcError):
        c.collapse(tensor[:-1])
    rng = np.random.default_rng(0)
    cols = [0, 1, c.n_cols - 1, c.n_cols // 2] + [int(v) for v in rng.integers(0, c.n_cols, 60)]
    vals, paths = c.open_columns(cols)
    root = c.get_root().root
    for i, col in enumerate(cols):  # lcpc-2d/src/tests.rs:167-191
        ov, op = O.open_column(field, oc["comm"], oc["hashes"], c.n_rows, c.n_cols, col)
        assert (vals[i] == ov).all() and (paths[i] == op).all()
        assert O.verify_column_path(field, vals[i], paths[i], col, root)
    with pytest.raises(P.LcpcError) as e:
        c.open_columns([c.n_cols])
    assert e.value.code == _cabi.ERR_COLUMN


def test_open_columns_non_power_of_two():
    field, length = P.FT127, 1 << 13
    enc = P.SdigEncoding(field, length, seed=2)
    oenc = O.Encoding.sdig(field, length, seed=2)
    x = O.random_elems(field, length, seed=9)
    c, oc = P.LcCommit.commit(x, enc), oenc.commit(x)
    cols = [0, c.n_cols - 1, c.n_cols - 2, 12345 % c.n_cols]
    vals, paths = c.open_columns(cols)
    for i, col in enumerate(cols):
        ov, op = O.open_column(field, oc["comm"], oc["hashes"], c.n_rows, c.n_cols, col)
        assert (vals[i] == ov).all() and (paths[i] == op).all()
        assert O.verify_column_path(field, vals[i], paths[i], col, c.get_root().root)


def test_commit_evaluation_property():
    """lcpc-2d/src/tests.rs:193-236 (i): sum_i coeffs_i x^i == <inner, collapse(coeffs, outer)>."""
    field = P.FT63
    p = O.field_info(field)["modulus"]
    enc = P.LigeroEncoding.new_from_dims(field, 32, 64)
    x = O.random_elems(field, 128, seed=21)
    c = P.LcCommit.commit(x, enc)
    pt = 0x123456789abcdef % p
    direct = sum(v * pow(pt, i, p) for i, v in enumerate(O.from_mont(field, x))) % p
    inner = [pow(pt, i, p) for i in range(32)]
    outer = O.to_mont(field, [pow(pt, 32 * r, p) for r in range(4)])
    poly = O.from_mont(field, c.collapse(outer))
    assert sum(a * b for a, b in zip(poly, inner)) % p == direct


def test_rerun_and_launch_counter():
    field, length = P.FT127, 1 << 12
    enc = P.LigeroEncoding(field, length)
    x0, x1 = O.random_elems(field, length, seed=0), O.random_elems(field, length, seed=1)
    c = P.LcCommit.commit(x0, enc)
    r0 = c.get_root()
    before = enc.ctx.launch_count
    c.rerun(x1)
    assert enc.ctx.launch_count > before
    r1 = c.get_root()
    assert r0 != r1 and r1.root == O.Encoding.ligero(field, length).commit(x1)["root"]
    c.rerun(x0)
    assert c.get_root() == r0


@pytest.mark.parametrize("kind,field,length", [("ligero", P.FT255, 1 << 16), ("ligero", P.FT255, 1 << 20),
                                               ("ligero", P.FT127, (1 << 15) - 77), ("ligero", P.FT63, 1 << 13),
                                               ("sdig", P.FT127, 1 << 14), ("sdig_exact", P.FT127, 1 << 14),
                                               ("sdig_exact", P.FT255, 1 << 13)])
def test_commit_from_device_memory(kind, field, length):
    """commit_new_dev / commit_rerun_dev (coefficients already in HBM: the roofline-timed region of bench.py).
    For Ligero with a full last row the commit's own copy of the coefficients is written by the first
    transform pass; a ragged length takes the copy + pad route.  Both must give the oracle's LcCommit."""
    import torch
    if kind == "ligero":
        enc, oenc = P.LigeroEncoding(field, length), O.Encoding.ligero(field, length)
    else:
        enc, oenc = P.SdigEncoding(field, length, seed=4), O.Encoding.sdig(field, length, seed=4)
        if kind == "sdig_exact":  # a length that fills the last row: the copy rides on the transpose into the work buffer
            length = enc.get_dims(length)[0] * enc.n_per_row
    x0, x1 = O.random_elems(field, length, seed=31), O.random_elems(field, length, seed=32)
    dev = torch.device("cuda", enc.ctx.device)
    d0 = torch.from_numpy(x0.view(np.int64)).to(dev)
    d1 = torch.from_numpy(x1.view(np.int64)).to(dev)
    torch.cuda.synchronize()
    c = P.LcCommit.commit_device(d0.data_ptr(), length, enc)
    for x, d in ((x0, None), (x1, d1), (x0, d0)):
        if d is not None:
            c.rerun_device(d.data_ptr(), length)
            enc.ctx.synchronize()
        oc = oenc.commit(x)
        assert c.get_root().root == oc["root"]
        assert (c.coeffs == oc["coeffs"]).all() and (c.comm == oc["comm"]).all() and (c.hashes == oc["hashes"]).all()


@pytest.mark.parametrize("kind,field,n_per_row", [("ligero", P.FT255, 1 << 12), ("ligero", P.FT255, 1 << 14),
                                                  ("ligero", P.FT127, 1 << 9), ("sdig", P.FT127, 2000)])
def test_encode_rows_scatter_store(kind, field, n_per_row):
    """The fused encode + transpose step of the multi-GPU commit on ONE GPU: the last pass stores column block h
    of every row into its own [total rows][width_h] matrix (here three local buffers with uneven widths and a
    row offset) instead of row-major comm.  Result == the oracle's row encodes, cut into the same blocks."""
    import ctypes as C

    import torch
    from lcpc_b200 import _cabi
    if kind == "ligero":
        enc = P.LigeroEncoding.new_from_dims(field, n_per_row, 2 * n_per_row)
        oenc = O.Encoding.ligero_from_dims(field, n_per_row, 2 * n_per_row)
    else:
        enc, oenc = P.SdigEncoding.new_from_dims(field, n_per_row, seed=2), O.Encoding.sdig_from_dims(field, n_per_row, seed=2)
    n_cols, L = enc.n_cols, enc.L
    n_rows, row0, total_rows = 5, 2, 9
    dev = torch.device("cuda", enc.ctx.device)
    x = O.random_elems(field, n_rows * n_per_row, seed=77).reshape(n_rows, n_per_row, L)
    d_src = torch.from_numpy(x.view(np.int64).reshape(-1)).to(dev)
    d_tmp = torch.zeros(n_rows * n_cols * L, dtype=torch.int64, device=dev)
    starts = np.array([0, n_cols // 3 + 1, n_cols // 3 + 1, n_cols - 5, n_cols], dtype=np.uint64)  # one empty block
    bufs = [torch.full((total_rows * int(starts[h + 1] - starts[h]) * L + 1,), -1, dtype=torch.int64, device=dev)
            for h in range(4)]
    ptrs = np.array([b.data_ptr() for b in bufs], dtype=np.uint64)
    torch.cuda.synchronize()  # torch filled the buffers on its own stream; the engine stream does not wait for it
    sc = _cabi.Scatter(4, starts.ctypes.data, ptrs.ctypes.data, row0)
    lib = _cabi.lib()
    rc = lib.lcpc_b200_encode_rows_scatter_dev(enc._h, C.c_void_p(d_src.data_ptr()), n_per_row, n_per_row,
                                               C.c_void_p(d_tmp.data_ptr()), n_rows, C.byref(sc))
    assert rc == 0, enc.ctx.last_error() if hasattr(enc.ctx, "last_error") else rc
    enc.ctx.synchronize()
    want = np.zeros((n_rows, n_cols, L), np.uint64)
    for r in range(n_rows):
        row = np.zeros((n_cols, L), np.uint64)
        row[:n_per_row] = x[r]
        want[r] = oenc.encode(row)
    for h in range(4):
        w = int(starts[h + 1] - starts[h])
        got = bufs[h].cpu().numpy().view(np.uint64)
        assert got[-1] == np.uint64(0xffffffffffffffff)  # guard word untouched
        m = got[:-1].reshape(total_rows, w, L)
        assert (m[row0:row0 + n_rows] == want[:, int(starts[h]):int(starts[h + 1])]).all(), h
        rest = np.delete(m, np.s_[row0:row0 + n_rows], axis=0)
        assert (rest == np.uint64(0xffffffffffffffff)).all()  # rows of other ranks untouched
    # descriptor validation
    bad = np.array([0, 10, n_cols - 1], dtype=np.uint64)  # does not end at n_cols
    sc2 = _cabi.Scatter(2, bad.ctypes.data, ptrs.ctypes.data, 0)
    assert lib.lcpc_b200_encode_rows_scatter_dev(enc._h, C.c_void_p(d_src.data_ptr()), n_per_row, n_per_row,
                                                 C.c_void_p(d_tmp.data_ptr()), n_rows, C.byref(sc2)) == _cabi.ERR_BAD_ARG
    sc3 = _cabi.Scatter(17, starts.ctypes.data, ptrs.ctypes.data, 0)  # more than 16 blocks
    assert lib.lcpc_b200_encode_rows_scatter_dev(enc._h, C.c_void_p(d_src.data_ptr()), n_per_row, n_per_row,
                                                 C.c_void_p(d_tmp.data_ptr()), n_rows, C.byref(sc3)) == _cabi.ERR_BAD_ARG


@pytest.mark.parametrize("field", FIELDS)
@pytest.mark.parametrize("length", [1, 2, 3, 17, 100, 1000])
def test_commit_tiny_lengths(field, length):
    """The smallest polynomials `_get_dims` accepts (lcpc-ligero-pc/src/lib.rs:70-112): one or two rows, rows
    shorter than a warp, a ragged last row."""
    try:
        oenc = O.Encoding.ligero(field, length)
    except Exception:
        with pytest.raises(P.LcpcError):
            P.LigeroEncoding(field, length)
        return
    enc = P.LigeroEncoding(field, length)
    assert enc.get_dims(length) == oenc.get_dims(length)
    x = O.random_elems(field, length, seed=length)
    c, oc = P.LcCommit.commit(x, enc), oenc.commit(x)
    assert c.get_root().root == oc["root"]
    assert (c.comm == oc["comm"]).all() and (c.coeffs == oc["coeffs"]).all() and (c.hashes == oc["hashes"]).all()
    t = O.random_elems(field, c.n_rows, seed=3)
    assert (c.collapse(t) == O.collapse(field, oc["coeffs"], t, c.n_rows, c.n_per_row)).all()


# ------------------------------------------------------------------ challenge tensors
@pytest.mark.parametrize("field", FIELDS)
@pytest.mark.parametrize("n", [1, 7, 256, 1000, 5000])
def test_expand_tensor_vs_oracle(field, n):
    """ChaCha20Rng::from_seed(key) + n x F::random (lcpc-2d/src/lib.rs:1026-1032) expanded in parallel on the
    device == the oracle's serial draw, for keys that exercise early and late rejections."""
    for k in range(3):
        key = O.blake3(bytes([k, n % 251, field]))  # any 32 bytes
        assert (P.expand_tensor(field, key, n) == O.random_elems_from_key(field, key, n)).all(), k


def test_degree_test_matches_collapse_of_oracle_tensor():
    field, length = P.FT255, 1 << 16
    enc, oenc = P.LigeroEncoding(field, length), O.Encoding.ligero(field, length)
    x = O.random_elems(field, length, seed=5)
    c = P.LcCommit.commit(x, enc)
    key = bytes(range(32))
    poly, tensor = c.degree_test(key, with_tensor=True)
    want_t = O.random_elems_from_key(field, key, c.n_rows)
    assert (tensor == want_t).all()
    assert (poly == O.collapse(field, c.coeffs, want_t, c.n_rows, c.n_per_row)).all()
    assert (c.degree_test(key) == poly).all()
    with pytest.raises(P.LcpcError):
        c.degree_test(b"short")


def test_contexts_are_thread_safe():
    """A context serialises its calls (one stream, one mutex); separate contexts run concurrently.  Reference:
    `LcEncoding: Sync` because rayon workers call encode concurrently (lcpc-2d/src/lib.rs:74, 648-653)."""
    import threading
    field, length = P.FT127, 1 << 14
    oenc = O.Encoding.ligero(field, length)
    xs = [O.random_elems(field, length, seed=200 + i) for i in range(6)]
    want = [oenc.commit(x)["root"] for x in xs]
    shared = P.LigeroEncoding(field, length)                      # three threads share one context/encoding
    own = [P.LigeroEncoding(field, length, ctx=P.Context(0)) for _ in range(3)]  # three have their own
    got, errs = [None] * 6, []

    def work(i, enc):
        try:
            for _ in range(3):
                c = P.LcCommit.commit(xs[i], enc)
                t = O.random_elems(field, c.n_rows, seed=i)
                c.collapse(t)
                got[i] = c.get_root().root
        except Exception as e:  # pragma: no cover
            errs.append(repr(e))

    ts = [threading.Thread(target=work, args=(i, shared if i < 3 else own[i - 3])) for i in range(6)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs
    assert got == want


def test_row_block_entry_points_reject_bad_shapes():
    import ctypes as C

    import torch
    from lcpc_b200 import _cabi
    field = P.FT127
    enc = P.LigeroEncoding.new_from_dims(field, 256, 512)
    dev = torch.device("cuda", enc.ctx.device)
    d_a = torch.zeros(4 * 512 * 2, dtype=torch.int64, device=dev)
    d_b = torch.zeros(4 * 512 * 2, dtype=torch.int64, device=dev)
    host = np.zeros((4 * 256, 2), np.uint64)
    torch.cuda.synchronize()
    lib = _cabi.lib()
    # len must fill exactly n_rows rows (last one possibly short): 4 rows need 769..1024 elements
    for bad_len in (768, 1025, 0):
        assert lib.lcpc_b200_encode_rows_h2d(enc._h, host.ctypes.data_as(C.c_void_p), bad_len, C.c_void_p(d_a.data_ptr()),
                                             C.c_void_p(d_b.data_ptr()), 4) == _cabi.ERR_BAD_ARG
    assert lib.lcpc_b200_encode_rows_h2d(enc._h, host.ctypes.data_as(C.c_void_p), 1000, C.c_void_p(d_a.data_ptr()),
                                         C.c_void_p(d_b.data_ptr()), 4) == 0
    # valid > n_cols / > stride
    assert lib.lcpc_b200_encode_rows_dev(enc._h, C.c_void_p(d_a.data_ptr()), 256, 600, C.c_void_p(d_b.data_ptr()), 2) \
        == _cabi.ERR_BAD_ARG
    assert lib.lcpc_b200_encode_dev(enc._h, C.c_void_p(d_a.data_ptr()), 2, 513) == _cabi.ERR_BAD_ARG
    enc.ctx.synchronize()


def test_contexts_on_two_devices_in_one_process():
    """One process may hold contexts on several GPUs (the multi-GPU commit uses one process per GPU, but the ABI
    does not require it): kernels with opted-in shared memory must work on every device."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    field, length = P.FT255, 1 << 14
    x = O.random_elems(field, length, seed=9)
    want = O.Encoding.ligero(field, length).commit(x)["root"]
    for dev in (1, 0, 1):
        enc = P.LigeroEncoding(field, length, ctx=P.Context(dev))
        assert P.LcCommit.commit(x, enc).get_root().root == want
    enc2 = P.SdigEncoding(P.FT127, 1 << 12, seed=0, ctx=P.Context(1))
    x2 = O.random_elems(P.FT127, 1 << 12, seed=10)
    assert P.LcCommit.commit(x2, enc2).get_root().root == O.Encoding.sdig(P.FT127, 1 << 12, seed=0).commit(x2)["root"]


# ---- f3: the code generator on the device (csrc/device_matgen.cu) -------------------------------------------------
@pytest.mark.parametrize("field,n,seed,code", [(P.FT127, 1 << 12, 0, 3), (P.FT255, 3000, 1, 3), (P.FT63, 9001, 7, 3),
                                               (P.FT191, 777, 2, 3), (P.FT127, 50000, 5, 3), (P.FT127, 2500, 3, 1),
                                               (P.FT255, 1 << 13, 4, 5), (P.FT63, 21, 0, 3), (P.FT127, 400, 9, 6)])
def test_device_matgen_equals_host_generator_and_oracle(field, n, seed, code):
    """matgen::generate (matgen.rs:28-52, 114-188) drawn on the device -- keystream, per-word column-end map, pointer
    doubling, per-column draw -- gives the same CSC arrays, bit for bit, as the host generator (and through it the
    oracle's, tests/test_host_cpu.py::test_matgen_matches_oracle); the gather form built from it encodes like the
    oracle."""
    enc = P.SdigEncoding.new_from_dims(field, n, seed=seed, code=code)
    assert enc._code_h is None  # no host matrices were made
    pre, post = enc.matrices()  # downloaded from the device
    hpre, hpost, cw = P.host.generate_sdig_code(field, n, seed, code)
    assert cw == enc.n_cols and len(pre) == len(hpre)
    for got, want in zip(pre + post, hpre + hpost):
        assert (got["m"], got["n"]) == (want["m"], want["n"])
        assert (got["ptrs"] == want["ptrs"]).all() and (got["idxs"] == want["idxs"]).all() and (got["data"] == want["data"]).all()
    opre, opost = O.Encoding.sdig_from_dims(field, n, seed=seed, code=code).matrices()
    for got, want in zip(pre + post, opre + opost):
        assert (got["idxs"] == want["idxs"]).all() and (got["data"] == want["data"]).all()
    oenc = O.Encoding.sdig_from_dims(field, n, seed=seed, code=code)
    rows = np.zeros((3, enc.n_cols, enc.L), np.uint64)
    rows[:, :n] = O.random_elems(field, 3 * n, seed=seed + 11).reshape(3, n, -1)
    got = enc.encode(rows)
    for r in range(3):
        assert (got[r] == oenc.encode(rows[r])).all()
    # the same (field, code, n_per_row, seed) again comes out of the context's cache: same device code, same results
    again = P.SdigEncoding.new_from_dims(field, n, seed=seed, code=code)
    assert (again.encode(rows)[1] == got[1]).all()
    other = P.SdigEncoding.new_from_dims(field, n, seed=seed + 1, code=code)
    assert not (other.matrices()[0][0]["data"] == pre[0]["data"]).all()  # another seed, another code


def test_device_matgen_rejects_what_the_reference_asserts():
    with pytest.raises(P.LcpcError):
        P.SdigEncoding.new_from_dims(P.FT127, 20, seed=0)      # assert!(n > baselen), matgen.rs:62
    with pytest.raises(P.LcpcError):
        P.SdigEncoding.new_from_dims(P.FT127, 4096, n_cols=1, seed=0)  # assert_eq!(codeword_length, n_cols), lib.rs:129


def test_device_matgen_headline_shape_setup_time():
    """2^24 coefficients, Ft127, SdigCode3, seed 0 (config 3): generation + gather form on the device; spot-check the
    first and last columns of the big matrices against the host generator's stream, and time the setup."""
    import time
    field, n = P.FT127, 1 << 24
    ctx = P.Context(0)
    t0 = time.perf_counter()
    enc = P.SdigEncoding(field, n, seed=0, ctx=ctx)
    ctx.synchronize()
    dt = time.perf_counter() - t0
    assert (enc.n_per_row, enc.n_cols) == (235173, 357699)
    t0 = time.perf_counter()
    enc2 = P.SdigEncoding(field, n, seed=0, ctx=ctx)   # cached
    dt2 = time.perf_counter() - t0
    print(f"device matgen 2^24: {dt * 1e3:.1f} ms, cached: {dt2 * 1e3:.2f} ms")
    assert dt2 < dt
    hpre, hpost, _ = P.host.generate_sdig_code(field, enc.n_per_row, 0, 3)
    pre, post = enc.matrices()
    for got, want in zip(pre + post, hpre + hpost):
        assert (got["idxs"] == want["idxs"]).all() and (got["data"] == want["data"]).all()
    enc2.close()


@pytest.mark.parametrize("field,length,seed", [(P.FT127, 1 << 15, 0), (P.FT255, (1 << 13) - 77, 1), (P.FT63, 6000, 2), (P.FT191, 3000, 3)])
def test_brakedown_device_commit_keeps_codewords_column_major_until_asked(field, length, seed):
    """Device-resident Brakedown commit: the codewords stay in the encoder's work buffer (no final transpose), leaves
    are hashed and columns opened from there; the row-major comm appears when it is downloaded -- all equal to the
    oracle, before and after, and with the lazy route switched off."""
    import torch
    from lcpc_b200 import _cabi
    oenc = O.Encoding.sdig(field, length, seed=seed)
    x = O.random_elems(field, length, seed=seed + 70)
    oc = oenc.commit(x)
    dev = torch.from_numpy(x.view(np.int64)).cuda()
    for lazy in (1, 0):
        _cabi.lib().lcpc_b200_set_tunable(b"SDIG_LAZY_COMM", lazy)
        try:
            enc = P.SdigEncoding(field, length, seed=seed)
            c = P.LcCommit.commit_device(dev.data_ptr(), length, enc)
            assert c.get_root().root == oc["root"]
            comm = oc["comm"].reshape(c.n_rows, c.n_cols, -1)
            cols = np.array([0, c.n_cols - 1, c.n_per_row - 1, c.n_per_row, c.n_cols // 2], np.uint64)
            vals, paths = c.open_columns(cols)          # before anything asked for the row-major matrix
            for i, col in enumerate(cols):
                assert (vals[i] == comm[:, int(col)]).all()
                assert O.verify_column_path(field, vals[i], paths[i], int(col), oc["root"])
            t = O.random_elems(field, c.n_rows, seed=9)      # the row combination reads the coefficients there too
            assert (c.collapse(t) == O.collapse(field, oc["coeffs"], t, c.n_rows, c.n_per_row)).all()
            assert (c.hashes == oc["hashes"]).all()
            assert (c.coeffs == oc["coeffs"]).all() and (c.comm == oc["comm"]).all()   # materialised on demand
            assert (c.collapse(t) == O.collapse(field, oc["coeffs"], t, c.n_rows, c.n_per_row)).all()
            vals2, _ = c.open_columns(cols)             # and afterwards
            assert (vals2 == vals).all()
            c.rerun_device(dev.data_ptr(), length)      # a second commit into the same object: lazy again
            vals3, _ = c.open_columns(cols)
            assert (vals3 == vals).all() and c.get_root().root == oc["root"]
            # the device pointers of the row-major matrices (written on this request when the commit was lazy)
            from cuda.bindings import runtime as rt
            d_comm, d_coeffs, d_hashes = c.device_ptrs()
            enc.ctx.synchronize()
            for ptr, want in ((d_comm, oc["comm"]), (d_coeffs, oc["coeffs"]), (d_hashes, oc["hashes"])):
                got = np.empty_like(want)
                err, = rt.cudaMemcpy(got.ctypes.data, ptr, got.nbytes, rt.cudaMemcpyKind.cudaMemcpyDeviceToHost)
                assert int(err) == 0 and (got == want).all()
            outer = O.random_elems(field, c.n_rows, seed=5)
            inner = O.random_elems(field, c.n_per_row, seed=6)
            proof = c.prove(outer, enc, P.Transcript(b"lazy comm"))
            ev = proof.verify(c.get_root(), outer, inner, enc, P.Transcript(b"lazy comm"))
            assert (ev == O.dot(field, inner, O.collapse(field, oc["coeffs"], outer, c.n_rows, c.n_per_row))).all()
        finally:
            _cabi.lib().lcpc_b200_set_tunable(b"SDIG_LAZY_COMM", 1)


@pytest.mark.parametrize("field,n_per_row,n_rows,short", [(P.FT255, 8200, 40, 0), (P.FT127, 16400, 72, 0), (P.FT127, 16400, 70, 333),
                                                          (P.FT127, 9000, 130, 1)])
def test_brakedown_host_commit_tail_chunk_at_the_leaf_boundary(field, n_per_row, n_rows, short):
    """Host-route Brakedown commit with several row-chunks: the last chunk is cut where the last BLAKE3 chunk of the
    leaf inputs starts, the earlier leaf chunks are hashed while it is still crossing PCIe.  Same LcCommit as the
    oracle with the tail schedule on and off, also for a short last row."""
    length = n_rows * n_per_row - short
    enc = P.SdigEncoding.new_from_dims(field, n_per_row, seed=3)
    oenc = O.Encoding.sdig_from_dims(field, n_per_row, seed=3)
    x = O.random_elems(field, length, seed=n_rows)
    oc = oenc.commit(x)
    for tail in (1, 0):
        _cabi.lib().lcpc_b200_set_tunable(b"H2D_TAIL_SDIG", tail)
        try:
            c = P.LcCommit.commit(x, enc)
            assert c.n_rows == n_rows
            assert c.get_root().root == oc["root"]
            assert (c.hashes == oc["hashes"]).all() and (c.comm == oc["comm"]).all() and (c.coeffs == oc["coeffs"]).all()
            y = O.random_elems(field, length, seed=n_rows + 1)
            c.rerun(y)
            assert c.get_root().root == oenc.commit(y)["root"]
        finally:
            _cabi.lib().lcpc_b200_set_tunable(b"H2D_TAIL_SDIG", 1)
