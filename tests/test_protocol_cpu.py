"""CPU tests of the host-side protocol layer: the Fiat-Shamir transcript, the column challenge, the oracle's
prove/verify restatement and the bincode wire format.  Nothing here needs a GPU (no compute entry point of the C ABI
is called); the GPU side of prove/verify is in tests/test_gpu_protocol.py.
"""
import hashlib
import os
import random

import numpy as np
import pytest

import oracle as O
import lcpc_b200 as P
from oracle import protocol as PR
from oracle import transcript as T

# merlin's conformance vector (merlin/src/transcript.rs tests; the same constant is used by its Go/JS/C ports):
#   Transcript::new(b"test protocol"); append_message(b"some label", b"some data"); challenge_bytes(b"challenge", 32)
MERLIN_VECTOR = "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"


def test_keccak_permutation_matches_hashlib_sha3():
    rng = random.Random(7)
    for n in (0, 1, 55, 135, 136, 137, 167, 168, 169, 1000):
        d = bytes(rng.getrandbits(8) for _ in range(n))
        assert T.sponge(136, 0x06, d, 32) == hashlib.sha3_256(d).digest()
        assert T.sponge(168, 0x1F, d, 400) == hashlib.shake_128(d).digest(400)


def test_oracle_transcript_reproduces_merlin_vector():
    t = T.Transcript(b"test protocol")
    t.append_message(b"some label", b"some data")
    assert t.challenge_bytes(b"challenge", 32).hex() == MERLIN_VECTOR


def test_product_transcript_reproduces_merlin_vector():
    t = P.Transcript(b"test protocol")
    t.append_message(b"some label", b"some data")
    assert t.challenge_bytes(b"challenge", 32).hex() == MERLIN_VECTOR


@pytest.mark.parametrize("impl", ["base", "bmi", "bmi_table", "avx512"])
def test_every_keccak_build_reproduces_merlin_vector_and_agrees(impl):
    """The transcript's permutation has four builds (portable, BMI plane-wise and table-driven, AVX-512) chosen at run time; each one the CPU
    supports is forced in a fresh process and must give the merlin vector and the same digest of a long absorb."""
    import subprocess
    import sys
    code = ("import numpy as np, lcpc_b200 as P\n"
            "t = P.Transcript(b'test protocol'); t.append_message(b'some label', b'some data')\n"
            "print(t.challenge_bytes(b'challenge', 32).hex())\n"
            "t.append_reprs(b'$l//PR', (np.arange(5000 * 32) % 251).astype(np.uint8).reshape(5000, 32))\n"
            "print(t.challenge_bytes(b'x', 64).hex())\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = {}
    for which in ("base", impl):
        env = dict(os.environ, LCPC_B200_KECCAK=which, PYTHONPATH=root)
        outs[which] = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True).stdout.split()
    assert outs[impl][0] == MERLIN_VECTOR and outs[impl] == outs["base"]


def test_product_transcript_equals_oracle_on_random_operation_sequences():
    rng = random.Random(11)
    for trial in range(4):
        label = bytes(rng.getrandbits(8) for _ in range(rng.randint(0, 20)))
        a, b = P.Transcript(label), T.Transcript(label)
        for _ in range(250):
            lab = bytes(rng.getrandbits(8) for _ in range(rng.randint(0, 12)))
            kind = rng.random()
            if kind < 0.55:  # message sizes around the STROBE rate (166) exercise the mid-operation permutation
                m = os.urandom(rng.choice([0, 1, 8, 31, 32, 160, 163, 164, 165, 166, 167, 332, 1000]))
                a.append_message(lab, m)
                b.append_message(lab, m)
            elif kind < 0.7:
                x = rng.getrandbits(64)
                a.append_u64(lab, x)
                b.append_u64(lab, x)
            else:
                n = rng.choice([0, 1, 32, 64, 165, 166, 167, 500])
                assert a.challenge_bytes(lab, n) == b.challenge_bytes(lab, n)
        assert a.challenge_bytes(b"end", 32) == b.challenge_bytes(b"end", 32)


def test_transcript_clone_and_batched_reprs():
    a = P.Transcript(b"x")
    reprs = np.frombuffer(os.urandom(32 * 50), dtype=np.uint8).reshape(50, 32)
    b = a.clone()
    a.append_reprs(b"$l//PR", reprs)
    for r in reprs:
        b.append_message(b"$l//PR", r.tobytes())
    assert a.challenge_bytes(b"c", 32) == b.challenge_bytes(b"c", 32)
    c = a.clone()
    assert a.challenge_bytes(b"d", 16) == c.challenge_bytes(b"d", 16)


@pytest.mark.parametrize("n_cols", [1, 2, 3, 1000, 1024, 131072, 357699, (1 << 40) + 7, (1 << 63) + 1])
def test_sample_columns_matches_oracle(n_cols):
    """ChaCha20Rng::from_seed(key) + Uniform::new(0usize, n_cols): product C++ vs the oracle's restatement."""
    for key in (bytes(32), bytes(range(32)), hashlib.sha256(b"k").digest()):
        got = [int(v) for v in P.sample_columns(key, n_cols, 700)]
        assert got == PR.sample_columns(key, n_cols, 700)
        assert all(0 <= v < n_cols for v in got)


def test_sample_columns_rejects_bad_arguments():
    with pytest.raises(P.LcpcError):
        P.sample_columns(bytes(31), 10, 5)
    with pytest.raises(P.LcpcError):
        P.sample_columns(bytes(32), 0, 5)


# ------------------------------------------------------------------ oracle prove / verify (the checker itself)
def _oracle_case(field, enc, length, seed):
    x = O.random_elems(field, length, seed=seed)
    c = enc.commit(x)
    outer = O.random_elems(field, c["n_rows"], seed=seed + 1)
    inner = O.random_elems(field, c["n_per_row"], seed=seed + 2)
    proof = PR.prove(field, c, outer, enc.get_n_degree_tests(), enc.get_n_col_opens(), T.Transcript(b"test transcript"))
    return c, outer, inner, proof


@pytest.mark.parametrize("kind,field,length", [("ligero", O.FT255, 1 << 10), ("ligero", O.FT63, 3000),
                                               ("sdig", O.FT127, 1 << 12)])
def test_oracle_prove_verify_round_trip(kind, field, length):
    """The reference's own end-to-end property (lcpc-2d/src/tests.rs `end_to_end`, `end_to_end_two_proofs`): a proof
    verifies against the root and yields the evaluation <outer (x) inner, coeffs>."""
    enc = O.Encoding.ligero(field, length) if kind == "ligero" else O.Encoding.sdig(field, length, seed=0)
    c, outer, inner, proof = _oracle_case(field, enc, length, seed=3)
    ev = PR.verify(field, enc, c["root"], outer, inner, proof, T.Transcript(b"test transcript"))
    want = O.dot(field, inner, O.collapse(field, c["coeffs"], outer, c["n_rows"], c["n_per_row"]))
    assert (ev == want).all()
    # a different transcript label changes every challenge: the proof must not verify
    with pytest.raises(PR.VerifierError):
        PR.verify(field, enc, c["root"], outer, inner, proof, T.Transcript(b"another transcript"))


def test_oracle_verify_rejects_tampering():
    field, length = O.FT127, 1 << 11
    enc = O.Encoding.ligero(field, length)
    c, outer, inner, proof = _oracle_case(field, enc, length, seed=9)
    one = O.to_mont(field, [1])[0]

    def run(p, root=c["root"], o=outer, i=inner):
        return PR.verify(field, enc, root, o, i, p, T.Transcript(b"test transcript"))

    bad = dict(proof, columns=proof["columns"][:-1])
    with pytest.raises(PR.VerifierError) as e:
        run(bad)
    assert e.value.kind == "NumColOpens"
    with pytest.raises(PR.VerifierError) as e:
        run(proof, i=inner[:-1])
    assert e.value.kind == "InnerTensor"
    with pytest.raises(PR.VerifierError) as e:
        run(proof, o=np.concatenate([outer, outer[:1]]))
    assert e.value.kind == "OuterTensor"
    cols = [(col.copy(), path.copy()) for col, path in proof["columns"]]
    cols[5][0][0] = O.field_op(field, "add", cols[5][0][:1], one[None, :])[0]
    with pytest.raises(PR.VerifierError) as e:
        run(dict(proof, columns=cols))
    assert e.value.kind in ("ColumnDegree", "ColumnEval")
    cols = [(col.copy(), path.copy()) for col, path in proof["columns"]]
    cols[7][1][2, 3] ^= 1
    with pytest.raises(PR.VerifierError) as e:
        run(dict(proof, columns=cols))
    assert e.value.kind == "ColumnPath"
    with pytest.raises(PR.VerifierError) as e:
        run(proof, root=bytes(32))
    assert e.value.kind == "ColumnPath"


# ------------------------------------------------------------------ wire format
def _as_product_proof(field, proof):
    cols = np.stack([c for c, _ in proof["columns"]])
    paths = np.stack([p for _, p in proof["columns"]])
    return P.LcEvalProof(field, proof["n_cols"], proof["p_eval"], np.stack(proof["p_random_vec"]), cols, paths)


@pytest.mark.parametrize("kind,field,length", [("ligero", O.FT255, 1 << 10), ("sdig", O.FT127, 1 << 12),
                                               ("ligero", O.FT191, 700), ("ligero", O.FT63, 2000)])
def test_wire_proof_matches_oracle_writer_and_round_trips(kind, field, length):
    enc = O.Encoding.ligero(field, length) if kind == "ligero" else O.Encoding.sdig(field, length, seed=0)
    c, outer, inner, proof = _oracle_case(field, enc, length, seed=21)
    pp = _as_product_proof(field, proof)
    blob = P.serialize_proof(pp)
    assert blob == PR.wire_proof(proof)
    back = P.deserialize_proof(blob, field)
    assert back.n_cols == pp.n_cols and (back.p_eval == pp.p_eval).all() and (back.p_random_vec == pp.p_random_vec).all()
    assert (back.cols == pp.cols).all() and (back.paths == pp.paths).all()
    assert P.serialize_proof(back) == blob
    # size formula of the bincode layout: every Vec costs 8 bytes of length, every digest 8 + 32
    L, ndt = O.FIELD_LIMBS[field], len(proof["p_random_vec"])
    n_open, n_rows, plen = pp.cols.shape[0], pp.cols.shape[1], pp.paths.shape[1]
    want = 8 + (8 + pp.p_eval.shape[0] * 8 * L) * (1 + ndt) + 8 + 8 + n_open * (8 + n_rows * 8 * L + 8 + plen * 40)
    assert len(blob) == want


def test_wire_root_and_commit_fields():
    field, length = O.FT255, 1 << 10
    enc = O.Encoding.ligero(field, length)
    c = enc.commit(O.random_elems(field, length, seed=1))
    assert P.serialize_root(P.LcRoot(c["root"])) == PR.wire_root(c["root"]) == (32).to_bytes(8, "little") + c["root"]
    assert P.deserialize_root(PR.wire_root(c["root"])).root == c["root"]
    blob = PR.wire_commit(c)

    class HostCommit:  # the attributes serialize_commit reads from an LcCommit
        comm, coeffs, hashes = c["comm"], c["coeffs"], c["hashes"]
        n_rows, n_cols, n_per_row = c["n_rows"], c["n_cols"], c["n_per_row"]

    assert P.serialize_commit(HostCommit) == blob
    f = P.deserialize_commit_fields(blob, field)
    assert (f["comm"] == c["comm"]).all() and (f["coeffs"] == c["coeffs"]).all() and (f["hashes"] == c["hashes"]).all()
    assert (f["n_rows"], f["n_cols"], f["n_per_row"]) == (c["n_rows"], c["n_cols"], c["n_per_row"])


def test_wire_rejects_malformed_input():
    field = O.FT63
    with pytest.raises(P.LcpcError):
        P.deserialize_root(b"\x20" + bytes(7) + bytes(31))  # truncated digest
    with pytest.raises(P.LcpcError):
        P.deserialize_root((31).to_bytes(8, "little") + bytes(31))  # wrong digest length
    with pytest.raises(P.LcpcError):
        P.deserialize_root((32).to_bytes(8, "little") + bytes(33))  # trailing bytes
    with pytest.raises(P.LcpcError):
        P.deserialize_proof((1024).to_bytes(8, "little") + (1 << 60).to_bytes(8, "little"), field)  # absurd length
    enc = O.Encoding.ligero(field, 512)
    c = enc.commit(O.random_elems(field, 512, seed=1))
    blob = bytearray(PR.wire_commit(c))
    off = 8 + c["comm"].nbytes + 8 + c["coeffs"].nbytes
    blob[off:off + 8] = (c["n_rows"] + 1).to_bytes(8, "little")  # n_rows inconsistent with comm.len()
    with pytest.raises(P.LcpcError):
        P.deserialize_commit_fields(bytes(blob), field)


# ------------------------------------------------------------------ pins against the real reference's output
def _reference_sizes():
    import json
    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_proof_sizes.json")))


def _proof_shape(kind, field, lgl, rho=None, code=None):
    """(n_rows, n_per_row, n_cols, n_col_opens, n_degree_tests) from the PRODUCT's host-side setup functions."""
    import ctypes as C
    from lcpc_b200 import _cabi, host
    lib, n = _cabi.lib(), 1 << lgl
    if kind == "ligero":
        n_rows, n_per_row, n_cols = host.ligero_get_dims(field, n, tuple(rho))
        n_open = lib.lcpc_b200_ligero_n_col_opens(rho[0], rho[1])
    else:
        npr = C.c_size_t()
        assert lib.lcpc_b200_sdig_choose_n_per_row(field, code, n, C.byref(npr)) == 0
        n_per_row = npr.value
        pre, post = O.sdig_level_dims(field, code, n_per_row)  # level shapes only; codeword_length, encode.rs:18-33
        n_cols = pre[0][0] + post[-1][0] + sum(p[1] for p in pre[:-1]) + sum(p[1] for p in post)
        n_rows = (n + n_per_row - 1) // n_per_row
        n_open = lib.lcpc_b200_sdig_n_col_opens(code)
    ndt = host.n_degree_tests(128, n_cols, lib.lcpc_b200_field_flog2(field))
    return n_rows, n_per_row, n_cols, n_open, ndt


@pytest.mark.parametrize("name", ["ligero_rho_1_4", "ligero_rho_1_2", "sdig_code3"])
def test_proof_sizes_equal_the_reference_runs(name):
    """The reference's own benchmark logs hold `bincode::serialize(&pf).len()` for Ft255 proofs at 2^13 .. 2^29
    coefficients.  Our dimension choosers, n_col_opens / n_degree_tests and the bincode layout together must give
    exactly those byte counts -- a pin against real Rust output (not against our restatement)."""
    g = _reference_sizes()
    case, field = g[name], P.FT255
    for lgl, want in zip(g["lgl"], case["proof_bytes"]):
        kind = "ligero" if "rho" in case else "sdig"
        n_rows, n_per_row, n_cols, n_open, ndt = _proof_shape(kind, field, lgl, case.get("rho"), case.get("code"))
        path_len = (n_cols - 1).bit_length()
        size = 8 + (8 + n_per_row * 32) + 8 + ndt * (8 + n_per_row * 32) + 8 + n_open * (8 + n_rows * 32 + 8 + path_len * 40)
        assert size == want, (name, lgl, size, want)
        if lgl <= 15:  # and the serializer itself, on a proof of that shape
            pf = P.LcEvalProof(field, n_cols, np.zeros((n_per_row, 4), np.uint64), np.zeros((ndt, n_per_row, 4), np.uint64),
                               np.zeros((n_open, n_rows, 4), np.uint64), np.zeros((n_open, path_len, 32), np.uint8))
            assert len(P.serialize_proof(pf)) == want


def test_oracle_dims_agree_with_the_reference_runs():
    """Same pin for the ORACLE's dimension functions (they feed every parity test)."""
    g = _reference_sizes()
    for lgl, want in zip(g["lgl"][:4], g["ligero_rho_1_2"]["proof_bytes"][:4]):
        enc = O.Encoding.ligero(P.FT255, 1 << lgl)
        n_rows, n_per_row, n_cols = enc.get_dims(1 << lgl)
        n_open, ndt, path_len = enc.get_n_col_opens(), enc.get_n_degree_tests(), (n_cols - 1).bit_length()
        assert 8 + (8 + n_per_row * 32) * (1 + ndt) + 16 + n_open * (16 + n_rows * 32 + path_len * 40) == want
    for lgl, want in zip(g["lgl"][:3], g["sdig_code3"]["proof_bytes"][:3]):
        enc = O.Encoding.sdig(P.FT255, 1 << lgl, seed=0)
        n_rows, n_per_row, n_cols = enc.get_dims(1 << lgl)
        n_open, ndt, path_len = enc.get_n_col_opens(), enc.get_n_degree_tests(), (n_cols - 1).bit_length()
        assert 8 + (8 + n_per_row * 32) * (1 + ndt) + 16 + n_open * (16 + n_rows * 32 + path_len * 40) == want


def test_wire_round_trip_on_random_shapes():
    """Property test: any proof shape survives serialize -> deserialize, and the numpy writer equals the oracle's
    element-by-element writer (hypothesis drives field, dims, number of degree tests / columns, path length)."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=40, deadline=None)
    @given(st.sampled_from([O.FT63, O.FT127, O.FT191, O.FT255]), st.integers(1, 9), st.integers(0, 3), st.integers(1, 6),
           st.integers(1, 5), st.integers(0, 4), st.integers(0, 2**32 - 1))
    def run(field, n_per_row, ndt, n_open, n_rows, path_len, seed):
        rng = np.random.default_rng(seed)
        L = O.FIELD_LIMBS[field]
        u64 = lambda *shape: rng.integers(0, 2**64, size=shape, dtype=np.uint64)
        p_eval, p_rand, cols = u64(n_per_row, L), u64(ndt, n_per_row, L), u64(n_open, n_rows, L)
        paths = rng.integers(0, 256, size=(n_open, path_len, 32), dtype=np.uint8)
        pf = P.LcEvalProof(field, int(rng.integers(1, 2**40)), p_eval, p_rand, cols, paths)
        blob = P.serialize_proof(pf)
        oracle_form = dict(n_cols=pf.n_cols, p_eval=p_eval, p_random_vec=list(p_rand), columns=[(cols[j], paths[j]) for j in range(n_open)])
        assert blob == PR.wire_proof(oracle_form)
        back = P.deserialize_proof(blob, field)
        assert back.n_cols == pf.n_cols and (back.p_eval == p_eval).all() and (back.cols == cols).all()
        assert back.p_random_vec.shape == p_rand.shape and (back.p_random_vec == p_rand).all()
        assert back.paths.shape == paths.shape and (back.paths == paths).all()
        for cut in (1, 8, len(blob) // 2):
            if cut < len(blob):
                with pytest.raises(P.LcpcError):
                    P.deserialize_proof(blob[:-cut], field)

    run()


def test_product_chacha20rng_reproduces_rand_chacha_vectors():
    """The PRODUCT's host ChaCha20Rng (csrc/host_chacha.h, behind lcpc_b200_sample_columns) against rand_chacha 0.3's own
    test vectors.  Uniform::new(0, 2^64 - 1) exposes the raw stream: range = 2^64 - 1 gives zone = 2^64 - 2 and
    (hi, lo) = v * range = (v - 1, 2^64 - v), accepted iff v >= 2, so every draw is next_u64() - 1."""
    from test_oracle import CHACHA_ZERO_KEY_BLOCK0, CHACHA_ZERO_KEY_BLOCK1
    words = CHACHA_ZERO_KEY_BLOCK0 + CHACHA_ZERO_KEY_BLOCK1
    want = [(words[2 * i] | (words[2 * i + 1] << 32)) - 1 for i in range(16)]
    got = [int(v) for v in P.sample_columns(bytes(32), (1 << 64) - 1, 16)]
    assert got == want
    seed = bytes([0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 2, 0, 0, 0, 0, 0, 0, 0, 3, 0, 0, 0, 0, 0, 0, 0])
    first = int(P.sample_columns(seed, (1 << 64) - 1, 1)[0]) + 1
    assert first & 0xFFFFFFFF == 137206642  # test_chacha_construction: next_u32() of this seed
    # and the oracle's Uniform draw agrees on the same degenerate range
    assert [int(v) for v in PR.sample_columns(bytes(32), (1 << 64) - 1, 16)] == want


class _Pcg32:
    """rand_pcg 0.3 Lcg64Xsh32 (what rand 0.8's own tests use as `crate::test::rng(seed)`)."""
    MUL, M = 6364136223846793005, (1 << 64) - 1

    def __init__(self, state, stream):
        self.inc = ((stream << 1) | 1) & self.M
        self.state = (state + self.inc) & self.M
        self._step()

    def _step(self):
        self.state = (self.state * self.MUL + self.inc) & self.M

    def next_u32(self):
        st = self.state
        self._step()
        rot, xsh = st >> 59, (((st >> 18) ^ st) >> 27) & 0xFFFFFFFF
        return ((xsh >> rot) | (xsh << ((32 - rot) & 31))) & 0xFFFFFFFF


def _uniform_new_sample(next_word, low, high, bits):
    """rand 0.8 `uniform_int_impl!`: Uniform::new(low, high).sample() for an unsigned type of `bits` bits."""
    mx = (1 << bits) - 1
    rng_range = (high - low) & mx
    zone = mx - (mx - rng_range + 1) % rng_range
    while True:
        m = next_word() * rng_range
        if m & mx <= zone:
            return low + (m >> bits)


def _uniform_sample_single(next_word, low, high, bits):
    rng_range = high - low
    zone = ((rng_range << (bits - rng_range.bit_length())) - 1) & ((1 << bits) - 1)
    while True:
        m = next_word() * rng_range
        if m & ((1 << bits) - 1) <= zone:
            return low + (m >> bits)


def test_uniform_rejection_zone_is_rand_08s():
    """rand 0.8 src/distributions/uniform.rs `value_stability`: with rng = Pcg32::new(897, 11634580027462260723),
    three `sample_single(11u32, 219)` give [17, 66, 214] and then three `Uniform::new(11u32, 219)` samples give
    [181, 93, 165].  The same macro instantiated for usize (64-bit words) is what prove()/verify() draw column numbers
    with (lcpc-2d/src/lib.rs:1077-1080) and what matgen draws row indices with (matgen.rs:119,147-158): both the
    oracle's and the product's draw must equal this restatement fed with the ChaCha20 stream."""
    rng = _Pcg32(897, 11634580027462260723)
    assert [_uniform_sample_single(rng.next_u32, 11, 219, 32) for _ in range(3)] == [17, 66, 214]
    assert [_uniform_new_sample(rng.next_u32, 11, 219, 32) for _ in range(3)] == [181, 93, 165]
    key = bytes(range(32))
    kw = np.frombuffer(key, dtype="<u4")
    for n_cols in (1, 2, 3, 1000, 131072, 357699, (1 << 63) + 12345, (1 << 64) - 1):
        words = iter(int(w) for c in range(64) for w in O.chacha_block(kw, c, 0))
        next_u64 = lambda: next(words) | (next(words) << 32)  # noqa: E731  (rand_core: next_u64 = lo word, then hi word)
        want = [_uniform_new_sample(next_u64, 0, n_cols, 64) for _ in range(200)]
        assert [int(v) for v in P.sample_columns(key, n_cols, 200)] == want
        assert [int(v) for v in PR.sample_columns(key, n_cols, 200)] == want
