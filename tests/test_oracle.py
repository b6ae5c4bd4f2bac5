"""Pins the CPU oracle (oracle/) before anything is compared against it.

Sources of truth, in order of strength:
  * tests/golden/blake3_vectors.json -- produced by the Python binding of the Rust `blake3` crate the
    reference hashes with: raw hashes, the leaf rule (lcpc-2d/src/lib.rs:719-735) and the node rule
    (:770-775);
  * tests/golden/expander_vectors.npz -- produced by running the reference's own
    doc/encoding.py: pins lcpc-brakedown-pc/src/encode.rs:36-110 (recursion, layout, orientation);
  * Python big-int arithmetic + SURVEY.md App. A constants -- pins the field/NTT restatement
    (the reference holds no known-answer vector for those: "parity unpinned");
  * the reference's own differential/property tests (lcpc-2d/src/tests.rs:127-236) restated.
"""
import json
import os

import numpy as np
import pytest

import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
FIELDS = [O.FT63, O.FT127, O.FT191, O.FT255]

# SURVEY.md Appendix A (derived with sympy from lcpc-test-fields/src/lib.rs:19-20,31-32,43-44,55-56)
CONSTS = {
    O.FT63: dict(p=0x46d0760000000001, gen=10, s=41, inv=0x46d075ffffffffff),
    O.FT127: dict(p=0x6e754097ba20e0bf7f2bd90000000001, gen=3, s=40, inv=0x7f2bd8ffffffffff),
    O.FT191: dict(p=0x453708aa3fbc8dda936888270ceecbcdd246820000000001, gen=5, s=41, inv=0xd24681ffffffffff),
    O.FT255: dict(p=0x663c799b6e4d2900fda9df04b9575969ef73c79086595f3002a4f20000000001, gen=5, s=41,
                  inv=0x02a4f1ffffffffff),
}


def pattern(n):
    return bytes(i % 251 for i in range(n))


def rand_ints(rng, p, n):
    return [int.from_bytes(rng.bytes(40), "little") % p for _ in range(n)]


# ------------------------------------------------------------------ fields
@pytest.mark.parametrize("field", FIELDS)
def test_field_constants(field):
    info, c = O.field_info(field), CONSTS[field]
    L = info["limbs"]
    R = 1 << (64 * L)
    assert info["modulus"] == c["p"] and info["s"] == c["s"] and info["inv"] == c["inv"]
    assert info["num_bits"] == c["p"].bit_length()
    assert info["r"] == R % c["p"] and info["r2"] == R * R % c["p"]
    t = (c["p"] - 1) >> c["s"]
    assert info["rou_mont"] == pow(c["gen"], t, c["p"]) * R % c["p"]
    assert (c["inv"] * c["p"] + 1) % (1 << 64) == 0


def test_montgomery_convention_on_a_published_instance():
    """The formulas test_field_constants holds the oracle to -- R = 2^(64 L) mod p with L the least limb count such
    that 2p <= 2^(64 L) (ff_derive's rule), INV = -p^-1 mod 2^64, ROOT_OF_UNITY = gen^((p-1)/2^S) in Montgomery form,
    limbs little-endian -- reproduce the constants the `bls12_381` crate publishes for its scalar field (same
    `ff::PrimeField` convention; src/scalar.rs: MODULUS, R, R2, INV, GENERATOR = 7, S = 32, ROOT_OF_UNITY)."""
    q = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
    L = 1
    while (1 << (64 * L)) < 2 * q:
        L += 1
    assert L == 4
    R = (1 << (64 * L)) % q

    def limbs(v):
        return [(v >> (64 * i)) & ((1 << 64) - 1) for i in range(L)]

    assert limbs(R) == [0x00000001fffffffe, 0x5884b7fa00034802, 0x998c4fefecbc4ff5, 0x1824b159acc5056f]
    assert limbs(R * R % q) == [0xc999e990f3f29c6d, 0x2b6cedcb87925c23, 0x05d314967254398f, 0x0748d9d99f59ff11]
    assert (-pow(q, -1, 1 << 64)) % (1 << 64) == 0xfffffffeffffffff
    rou = pow(7, (q - 1) >> 32, q)
    assert limbs(rou * R % q) == [0xb9b58d8c5f0e466a, 0x5b1b4c801819d7ec, 0x0af53ae352a31e64, 0x5bf3adda19e9b27b]
    # the four test fields follow the same limb rule (and the reference declares exactly these array lengths,
    # lcpc-test-fields/src/lib.rs:22,34,46,58)
    for field, c in CONSTS.items():
        n = 1
        while (1 << (64 * n)) < 2 * c["p"]:
            n += 1
        assert n == O.FIELD_LIMBS[field]


@pytest.mark.parametrize("field", FIELDS)
def test_field_ops_vs_bigint(field):
    p = CONSTS[field]["p"]
    L = O.FIELD_LIMBS[field]
    R = 1 << (64 * L)
    rng = np.random.default_rng(field)
    a = rand_ints(rng, p, 300) + [0, 1, p - 1, p - 1, 0]
    b = rand_ints(rng, p, 300) + [0, p - 1, p - 1, 1, p - 1]
    am, bm = O.to_mont(field, a), O.to_mont(field, b)
    assert O.elems_to_ints(am) == [x * R % p for x in a]
    assert O.from_mont(field, am) == a
    assert O.from_mont(field, O.field_op(field, "add", am, bm)) == [(x + y) % p for x, y in zip(a, b)]
    assert O.from_mont(field, O.field_op(field, "sub", am, bm)) == [(x - y) % p for x, y in zip(a, b)]
    assert O.from_mont(field, O.field_op(field, "mul", am, bm)) == [(x * y) % p for x, y in zip(a, b)]
    nz = [x for x in a if x]
    inv = O.from_mont(field, O.field_op(field, "inv", O.to_mont(field, nz)))
    assert all(x * y % p == 1 for x, y in zip(nz, inv))
    # to_repr = canonical little-endian bytes (PrimeFieldReprEndianness = "little")
    rep = O.to_repr(field, am)
    assert [int.from_bytes(r.tobytes(), "little") for r in rep] == a


@pytest.mark.parametrize("field", FIELDS)
def test_random_elems_in_range(field):
    x = O.random_elems(field, 500, seed=3, stream=1)
    assert all(v < CONSTS[field]["p"] for v in O.elems_to_ints(x))
    assert (O.random_elems(field, 500, seed=3, stream=1) == x).all()
    assert not (O.random_elems(field, 500, seed=3, stream=2) == x).all()


# ------------------------------------------------------------------ hash / rng
def test_blake3_golden_raw():
    g = json.load(open(os.path.join(GOLD, "blake3_vectors.json")))
    for v in g["raw"]:
        assert O.blake3(pattern(v["len"])).hex() == v["hash"], v["len"]
    assert O.blake3(bytes(64)).hex() == g["node_zero_zero"]
    n = g["node_left_right"]
    assert O.blake3(bytes.fromhex(n["left"]) + bytes.fromhex(n["right"])).hex() == n["hash"]


def test_blake3_vs_python_binding():
    blake3 = pytest.importorskip("blake3")
    rng = np.random.default_rng(7)
    for n in [0, 1, 64, 65, 1024, 1025, 2047, 2048, 2049, 3072, 5000, 8224, 40000]:
        data = rng.bytes(n)
        assert O.blake3(data) == blake3.blake3(data).digest()


def test_leaf_rule_golden():
    """leaf = D(0^32 || repr(col[0]) || ..), lcpc-2d/src/lib.rs:719-735, through merkleize."""
    g = json.load(open(os.path.join(GOLD, "blake3_vectors.json")))
    fld = {8: O.FT63, 16: O.FT127, 24: O.FT191, 32: O.FT255}
    for v in g["leaf_of_1_to_n"]:
        field, n_rows = fld[v["elem_bytes"]], v["n_rows"]
        col = O.to_mont(field, list(range(1, n_rows + 1)))
        # a 2-column matrix whose columns are both 1..n_rows
        comm = np.repeat(col[:, None, :], 2, axis=1).reshape(-1, O.FIELD_LIMBS[field])
        for serial in (False, True):
            h = O.merkleize(field, comm, n_rows, 2, serial=serial)
            assert h[0].tobytes().hex() == v["hash"] and h[1].tobytes().hex() == v["hash"]
            assert h[2].tobytes() == O.blake3(h[0].tobytes() + h[1].tobytes())


def test_chacha_rfc8439_block():
    # RFC 8439 2.3.2 with the IETF nonce folded into (counter, stream) words 12..15
    key = np.frombuffer(bytes(range(32)), dtype="<u4")
    counter = 1 | (0x09000000 << 32)
    stream = 0x4A000000 | (0x00000000 << 32)
    out = O.chacha_block(key, counter, stream)
    assert out[0] == 0xE4E7F110 and out[15] == 0x4E3C50A2 and out[4] == 0xC7F4D1C7 and out[7] == 0x4E6CD4C3


# ------------------------------------------------------------------ NTT
def direct_dft_bitrev(x, w, p):
    n = len(x)
    lg = n.bit_length() - 1
    out = [0] * n
    for i in range(n):
        k = int(format(i, f"0{lg}b")[::-1], 2) if lg else 0
        out[i] = sum(xj * pow(w, j * k, p) for j, xj in enumerate(x)) % p
    return out


@pytest.mark.parametrize("field", FIELDS)
@pytest.mark.parametrize("n", [1, 2, 8, 64])
def test_fft_io_is_bitreversed_dft(field, n):
    """fffft fft_io: position i holds X[bitrev(i)], X[k] = sum_j x_j w^(jk), w = rou^(2^(S-log n))."""
    c = CONSTS[field]
    p = c["p"]
    rng = np.random.default_rng(n)
    x = rand_ints(rng, p, n)
    w = pow(pow(c["gen"], (p - 1) >> c["s"], p), 1 << (c["s"] - (n.bit_length() - 1)), p)
    assert O.from_mont(field, O.root_of_unity(field, n).reshape(1, -1))[0] == w
    got = O.from_mont(field, O.fft_io(field, O.to_mont(field, x)))
    assert got == direct_dft_bitrev(x, w, p)
    back = O.from_mont(field, O.ifft_oi(field, O.fft_io(field, O.to_mont(field, x))))
    assert back == x


def test_fft_io_survey_pin():
    # Ft63 fft_io([1,2,3,4,0,0,0,0]) canonical values
    got = O.from_mont(O.FT63, O.fft_io(O.FT63, O.to_mont(O.FT63, [1, 2, 3, 4, 0, 0, 0, 0])))
    assert got == [0xa, 0x46d075ffffffffff, 0x3f1b208b6be814ac, 0x07b555749417eb51, 0x3978407655337b20,
                   0x247835e7671446dc, 0x1208be327d1cea3e, 0x1da7b76fc69b53cc]


def test_fft_errors():
    with pytest.raises(ValueError):
        O.fft_io(O.FT63, O.to_mont(O.FT63, [1, 2, 3]))


# ------------------------------------------------------------------ dims / parameters
def test_log2_and_degree_tests():
    # lcpc-2d/src/tests.rs:127-134 and lib.rs:613-616
    assert O.n_degree_tests(128, 1 << 17, 254) == 1
    assert O.n_degree_tests(128, 357699, 126) == 2
    assert O.n_degree_tests(128, 1 << 10, 62) == 3


def test_ligero_dims_survey_table():
    # SURVEY.md 8 config shapes (restating lcpc-ligero-pc/src/lib.rs:70-112)
    assert O.ligero_get_dims(O.FT255, 1 << 10) == (2, 512, 1024)
    assert O.ligero_get_dims(O.FT255, 1 << 20) == (64, 16384, 32768)
    assert O.ligero_get_dims(O.FT255, 1 << 24) == (256, 65536, 131072)
    assert O.ligero_get_dims(O.FT255, 1 << 28) == (1024, 262144, 524288)
    assert O.Encoding.ligero(O.FT255, 1 << 20).get_n_col_opens() == 309


def test_ligero_get_dims_property():
    # lcpc-ligero-pc/src/tests.rs:22-41
    rng = np.random.default_rng(0)
    for lgl in range(2, 30, 3):
        for _ in range(8):
            length = int(rng.integers(1 << lgl, 2 << lgl))
            n_rows, n_per_row, n_cols = O.ligero_get_dims(O.FT255, length)
            assert n_rows * n_per_row >= length and (n_rows - 1) * n_per_row < length
            assert n_cols & (n_cols - 1) == 0 and n_per_row * 2 == n_cols


def test_sdig_dims_survey_table():
    pre, post = O.sdig_level_dims(O.FT127, 3, 235173)
    assert [(n, m, d) for n, m, d in pre] == [(235173, 41861, 8), (41861, 7452, 8), (7452, 1327, 8),
                                              (1327, 237, 9), (237, 43, 14), (43, 8, 7)]
    assert [(n, m, d) for n, m, d in post][::-1] == [(13, 10, 8), (66, 58, 31), (361, 331, 28),
                                                     (2019, 1864, 24), (11335, 10475, 23), (63671, 58855, 23)]
    assert O.lib().lcpc_oracle_sdig_n_col_opens(3) == 6593


# ------------------------------------------------------------------ brakedown
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
    z = np.load(os.path.join(GOLD, "expander_vectors.npz"))
    field, pre, post, x, want = _npz_case(z, name)
    enc = O.Encoding.sdig_from_matrices(field, pre, post)
    row = np.zeros((enc.n_cols, enc.L), np.uint64)
    row[:x.shape[0]] = x
    assert (enc.encode(row) == want).all()


def test_matgen_shape_contract():
    """matgen.rs:114-188: every CSC column has exactly d sorted distinct rows and non-zero values;
    brakedown tests.rs:77-93: generate + encode run through the offset asserts."""
    for n, seed in [(256, 0), (1000, 1), (4351, 2)]:
        enc = O.Encoding.sdig_from_dims(O.FT63, n, seed=seed)
        pre_d, post_d = O.sdig_level_dims(O.FT63, 3, n)
        pre, post = enc.matrices()
        for M, (nn, mm, d) in list(zip(pre, pre_d)) + list(zip(post, post_d)):
            assert (M["n"], M["m"]) == (nn, mm)
            assert (np.diff(M["ptrs"].astype(np.int64)) == d).all()
            idx = M["idxs"].reshape(nn, d).astype(np.int64)
            assert (np.diff(idx, axis=1) > 0).all() and idx.max() < mm
            assert M["data"].any(axis=1).all()
        row = np.zeros((enc.n_cols, 1), np.uint64)
        row[:n] = O.random_elems(O.FT63, n, seed=9)
        enc.encode(row)
        # same seed -> same code; other seed -> other code
        again = O.Encoding.sdig_from_dims(O.FT63, n, seed=seed).matrices()[0][0]
        assert (again["idxs"] == pre[0]["idxs"]).all() and (again["data"] == pre[0]["data"]).all()


def test_encode_is_linear():
    # lcpc-2d/src/tests.rs:193-236 pins linearity of encode for the commit check
    for enc in (O.Encoding.ligero_from_dims(O.FT127, 32, 64), O.Encoding.sdig_from_dims(O.FT127, 300, seed=4)):
        f, n = enc.field, enc.n_per_row
        a = np.zeros((enc.n_cols, enc.L), np.uint64)
        b = a.copy()
        a[:n], b[:n] = O.random_elems(f, n, seed=1), O.random_elems(f, n, seed=2)
        s = a.copy()
        s[:n] = O.field_op(f, "add", a[:n], b[:n])
        assert (enc.encode(s) == O.field_op(f, "add", enc.encode(a), enc.encode(b))).all()


# ------------------------------------------------------------------ commit / prove pieces
@pytest.mark.parametrize("mk", [lambda: O.Encoding.ligero(O.FT255, 1 << 10),
                                lambda: O.Encoding.ligero_from_dims(O.FT63, 16, 64, rho=(1, 4)),
                                lambda: O.Encoding.sdig(O.FT127, 3000, seed=1)])
def test_commit_structure(mk):
    enc = mk()
    f = enc.field
    length = enc.n_per_row * 3 - 5  # ragged last row
    x = O.random_elems(f, length, seed=11)
    c = enc.commit(x)
    nr, npr, nc = c["n_rows"], c["n_per_row"], c["n_cols"]
    assert nr == 3
    coeffs = c["coeffs"].reshape(nr, npr, enc.L)
    comm = c["comm"].reshape(nr, nc, enc.L)
    assert (coeffs.reshape(-1, enc.L)[:length] == x).all() and not coeffs.reshape(-1, enc.L)[length:].any()
    for r in range(nr):
        row = np.zeros((nc, enc.L), np.uint64)
        row[:npr] = coeffs[r]
        assert (enc.encode(row) == comm[r]).all()
    # merkleize == merkleize_ser (lcpc-2d/src/tests.rs:136-149)
    assert (O.merkleize(f, c["comm"], nr, nc, serial=True) == c["hashes"]).all()
    np2 = 1 << (nc - 1).bit_length()
    assert not c["hashes"][nc:np2].any()
    # leaf rule straight from the definition
    col = 5
    data = bytes(32) + b"".join(O.to_repr(f, comm[:, col]).tobytes()[i * 8 * enc.L:(i + 1) * 8 * enc.L] for i in range(nr))
    assert c["hashes"][col].tobytes() == O.blake3(data)
    # open_column verifies against the root (lcpc-2d/src/tests.rs:167-191)
    for column in [0, 1, nc - 1, nc // 2]:
        cv, path = O.open_column(f, c["comm"], c["hashes"], nr, nc, column)
        assert (cv == comm[:, column]).all() and len(path) == (nc - 1).bit_length()
        assert O.verify_column_path(f, cv, path, column, c["root"])
        assert not O.verify_column_path(f, cv, path, column ^ 1, c["root"])
    with pytest.raises(IndexError):
        O.open_column(f, c["comm"], c["hashes"], nr, nc, nc)


def test_collapse_matches_serial_and_bigint():
    # lcpc-2d/src/tests.rs:151-165 (eval_outer vs eval_outer_ser)
    f, nr, npr = O.FT127, 7, 100
    p = CONSTS[f]["p"]
    coeffs = O.random_elems(f, nr * npr, seed=5)
    tensor = O.random_elems(f, nr, seed=6)
    a = O.collapse(f, coeffs, tensor, nr, npr)
    assert (a == O.collapse(f, coeffs, tensor, nr, npr, serial=True)).all()
    ci, ti = O.from_mont(f, coeffs), O.from_mont(f, tensor)
    want = [sum(ti[r] * ci[r * npr + c] for r in range(nr)) % p for c in range(npr)]
    assert O.from_mont(f, a) == want


def test_commit_evaluation_property():
    """lcpc-2d/src/tests.rs:193-236 (i): sum coeffs*x^i equals <inner, collapse(coeffs, outer)>."""
    f = O.FT63
    p = CONSTS[f]["p"]
    enc = O.Encoding.ligero_from_dims(f, 32, 64)
    length = 32 * 4
    x = O.random_elems(f, length, seed=21)
    c = enc.commit(x)
    pt = 0x123456789abcdef % p
    xi = O.from_mont(f, x)
    direct = sum(v * pow(pt, i, p) for i, v in enumerate(xi)) % p
    inner = [pow(pt, i, p) for i in range(32)]
    outer = [pow(pt, 32 * r, p) for r in range(4)]
    poly = O.from_mont(f, O.collapse(f, c["coeffs"], O.to_mont(f, outer), 4, 32))
    assert sum(a * b for a, b in zip(poly, inner)) % p == direct


def test_random_elems_from_key_is_the_serial_rejection_draw():
    """from_seed(key) + Field::random restated twice: the C oracle vs a direct Python walk over the ChaCha20
    word stream (two u32 per u64, low word first; mask the top limb; reject >= p)."""
    key = bytes((7 * i + 3) % 256 for i in range(32))
    kw = np.frombuffer(key, dtype="<u4").astype(np.uint32)
    for field in (O.FT63, O.FT127, O.FT191, O.FT255):
        info = O.field_info(field)
        p, nl, bits = info["modulus"], O.FIELD_LIMBS[field], info["num_bits"]
        words, blk = [], 0
        while len(words) < 2 * nl * 64:
            words += [int(w) for w in O.chacha_block(kw, blk, 0)]
            blk += 1
        got, pos = [], 0
        while len(got) < 20:
            limbs = [words[pos + 2 * i] | (words[pos + 2 * i + 1] << 32) for i in range(nl)]
            pos += 2 * nl
            v = sum(l << (64 * i) for i, l in enumerate(limbs)) & ((1 << bits) - 1)
            if v < p:
                got.append(v)
        want = O.random_elems_from_key(field, key, 20)
        assert [sum(int(want[i, j]) << (64 * j) for j in range(nl)) for i in range(20)] == got


# ------------------------------------------------------------------ third-party conventions pinned to PUBLISHED vectors
# The crates that hold these conventions are not vendored in the reference and no Rust toolchain exists here, so the
# oracle's restatements are held to the vectors those crates' own test suites assert (quoted from their sources).
CHACHA_ZERO_KEY_BLOCK0 = [0xade0b876, 0x903df1a0, 0xe56a5d40, 0x28bd8653, 0xb819d2bd, 0x1aed8da0, 0xccef36a8, 0xc70d778b,
                          0x7c5941da, 0x8d485751, 0x3fe02477, 0x374ad8b8, 0xf4b8436a, 0x1ca11815, 0x69b687c3, 0x8665eeb2]
CHACHA_ZERO_KEY_BLOCK1 = [0xbee7079f, 0x7a385155, 0x7c97ba98, 0x0d082d73, 0xa0290fcb, 0x6965e348, 0x3e53c612, 0xed7aee32,
                          0x7621b729, 0x434ee69c, 0xb03371d5, 0xd539d874, 0x281fed31, 0x45fb0a51, 0x1f0ae1ac, 0x6f4d794b]
CHACHA_ZERO_KEY_NONCE2 = [0x374dc6c2, 0x3736d58c, 0xb904e24a, 0xcd3f93ef, 0x88228b1a, 0x96a4dfb3, 0x5b76ab72, 0xc727ee54,
                          0x0e0e978a, 0xf3145c95, 0x1b748ea8, 0xf786c297, 0x99c28f5f, 0x628314e8, 0x398a19fa, 0x6ded1b53]


def test_chacha20rng_true_values_of_rand_chacha():
    """rand_chacha 0.3 src/chacha.rs `test_chacha_true_values_a` (IETF draft vectors 1 and 2: zero key, blocks 0 and 1
    of one stream => the 64-bit block counter sits in state words 12-13 and output words are consumed in order) and
    `test_chacha_nonce` (zero key, `set_stream(2u64 << (24 + 32))` => the stream id sits in words 14-15)."""
    zero = np.zeros(8, np.uint32)
    assert [int(w) for w in O.chacha_block(zero, 0, 0)] == CHACHA_ZERO_KEY_BLOCK0
    assert [int(w) for w in O.chacha_block(zero, 1, 0)] == CHACHA_ZERO_KEY_BLOCK1
    assert [int(w) for w in O.chacha_block(zero, 0, 2 << 56)] == CHACHA_ZERO_KEY_NONCE2


def test_chacha20rng_construction_vector_of_rand_chacha():
    """rand_chacha 0.3 `test_chacha_construction`: from_seed([0,0,0,0,0,0,0,0, 1,0,..., 2,0,..., 3,0,...]) then
    next_u32() == 137206642 -- the 32 seed bytes are the key as little-endian words."""
    seed = bytes([0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 2, 0, 0, 0, 0, 0, 0, 0, 3, 0, 0, 0, 0, 0, 0, 0])
    key = np.frombuffer(seed, dtype="<u4")
    assert int(O.chacha_block(key, 0, 0)[0]) == 137206642


def _pcg32_seed_bytes(state, n_bytes):
    """rand_core 0.6 SeedableRng::seed_from_u64: PCG32 (XSH RR) steps, one little-endian u32 per 4 seed bytes."""
    M, out = (1 << 64) - 1, b""
    for _ in range(n_bytes // 4):
        state = (state * 6364136223846793005 + 11634580027462260723) & M
        xs, rot = (((state >> 18) ^ state) >> 27) & 0xFFFFFFFF, state >> 59
        out += (((xs >> rot) | (xs << ((32 - rot) & 31))) & 0xFFFFFFFF).to_bytes(4, "little")
    return out


def test_seed_from_u64_value_of_rand_core_and_the_oracle_draw():
    """rand_core 0.6 `test_seed_from_u64` asserts seed_from_u64(0) of an 8-byte-seed RNG == 5029875928683246316; the
    oracle's seeded element draw (matgen's and random_coeffs' shape: ChaCha20Rng::seed_from_u64(seed), set_stream,
    F::random) must be the rejection sampling of exactly the ChaCha20 stream keyed by that expansion."""
    assert int.from_bytes(_pcg32_seed_bytes(0, 8), "little") == 5029875928683246316
    for seed, stream in ((0, 0), (7, 3), ((1 << 64) - 1, 1)):
        key = np.frombuffer(_pcg32_seed_bytes(seed, 32), dtype="<u4")
        words = [int(w) for c in range(40) for w in O.chacha_block(key, c, stream)]
        u64s = [words[2 * i] | (words[2 * i + 1] << 32) for i in range(len(words) // 2)]
        p = O.field_info(O.FT63)["modulus"]
        want = [v & ((1 << 63) - 1) for v in u64s if (v & ((1 << 63) - 1)) < p][:100]  # ff_derive random: mask, reject >= p
        got = [int(v) for v in O.random_elems(O.FT63, 100, seed=seed, stream=stream)[:, 0]]
        assert got == want


def test_commit_property_on_encoded_rows_at_the_headline_row_length():
    """lcpc-2d/src/tests.rs:193-236 (ii) at config 4's row shape (Ft255, 65536 coefficients -> 2^17 points): the ENCODED
    rows combined with the outer tensor (`eval_outer_fft`, lib.rs:1229-1249) and taken back through `ifft_oi` give the
    combined coefficient row -- high half exactly zero, low half equal to collapse_columns of the coefficients -- and
    the same evaluation as the direct sum.  This is everything the reference's own test pins about `fft_io_pc`: it is
    linear and inverted by `ifft_oi` (which order / which root it uses stays a property of the fffft crate)."""
    f = O.FT255
    p = CONSTS[f]["p"]
    n_rows, npr, nc = 4, 65536, 131072
    enc = O.Encoding.ligero_from_dims(f, npr, nc)
    x = O.random_elems(f, n_rows * npr, seed=17)
    c = enc.commit(x)
    pt = 0x1234567890abcdef1234567890abcdef % p
    outer_int = [pow(pt, npr * r, p) for r in range(n_rows)]
    outer = O.to_mont(f, outer_int)
    comm = c["comm"].reshape(n_rows, nc, -1)
    combined = np.zeros((nc, comm.shape[2]), np.uint64)
    for r in range(n_rows):
        t = np.repeat(outer[r:r + 1], nc, axis=0)
        combined = O.field_op(f, "add", combined, O.field_op(f, "mul", np.ascontiguousarray(comm[r]), t))
    back = O.ifft_oi(f, combined)
    assert not back[npr:].any()                       # "high coefficients zero" (:226-231)
    want = O.collapse(f, c["coeffs"], outer, n_rows, npr)
    assert (back[:npr] == want).all()
    # and the combination itself against big-int arithmetic on the first coefficients (collapse_columns' definition)
    head = [sum(O.from_mont(f, x[r * npr + i:r * npr + i + 1])[0] * outer_int[r] for r in range(n_rows)) % p for i in range(8)]
    assert O.from_mont(f, want[:8]) == head
