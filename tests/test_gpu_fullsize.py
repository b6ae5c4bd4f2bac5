"""Full-size parity (BASELINE.json configs 2-4) through checks whose cost does not grow with the
whole matrix: complete rows against the oracle's single-row encode, sampled columns against the leaf
rule, the whole Merkle tree rebuilt by the oracle from the GPU's leaves, linearity of the encoding,
and collapse against the oracle on the padded coefficient matrix.
"""
import numpy as np
import pytest

import oracle as O
import lcpc_b200 as P
from oracle import protocol as PR
from oracle.transcript import Transcript as OTranscript

pytestmark = pytest.mark.gpu


def check_prove_verify(c, enc, field, oc=None, oenc=None, seed=0):
    """Whole prove() + verify() at the full size.  With the oracle's commit at hand the proof is compared field by
    field; always: the proof verifies on the device against the commit's root, survives the wire, and the returned
    evaluation is <inner, collapse(coeffs, outer)> (lcpc-2d/src/tests.rs end_to_end)."""
    outer = O.random_elems(field, c.n_rows, seed=seed + 1)
    inner = O.random_elems(field, c.n_per_row, seed=seed + 2)
    proof = c.prove(outer, enc, P.Transcript(b"full size"))
    assert proof.cols.shape[:2] == (enc.get_n_col_opens(), c.n_rows)
    assert proof.p_random_vec.shape[0] == enc.get_n_degree_tests()
    if oc is not None:
        oproof = PR.prove(field, oc, outer, oenc.get_n_degree_tests(), oenc.get_n_col_opens(), OTranscript(b"full size"))
        assert P.serialize_proof(proof) == PR.wire_proof(oproof)
        assert [int(v) for v in proof.col_idx] == oproof["cols_to_open"]
    back = P.deserialize_proof(P.serialize_proof(proof), field)
    ev = back.verify(c.get_root(), outer, inner, enc, P.Transcript(b"full size"))
    want = O.dot(field, inner, O.collapse(field, c.coeffs, outer, c.n_rows, c.n_per_row))
    assert (ev == want).all()
    # every opened column is the committed column and its path leads to the root
    comm = c.comm.reshape(c.n_rows, c.n_cols, -1)
    for j in (0, proof.cols.shape[0] // 2, proof.cols.shape[0] - 1):
        col = int(proof.col_idx[j])
        assert (proof.cols[j] == comm[:, col]).all()
        assert O.verify_column_path(field, proof.cols[j], proof.paths[j], col, c.get_root().root)
    bad = proof.cols.copy()
    bad[-1, -1, 0] ^= 1
    with pytest.raises(P.LcpcError):
        P.LcEvalProof(field, proof.n_cols, proof.p_eval, proof.p_random_vec, bad, proof.paths).verify(
            c.get_root(), outer, inner, enc, P.Transcript(b"full size"))


def check_commit_samples(c, oenc, x, field, rows, cols, full_oracle=False):
    L = c.enc.L
    comm = c.comm.reshape(c.n_rows, c.n_cols, L)
    coeffs = c.coeffs.reshape(c.n_rows, c.n_per_row, L)
    flat = coeffs.reshape(-1, L)
    assert (flat[:x.shape[0]] == x).all() and not flat[x.shape[0]:].any()
    for r in rows:  # whole rows: encode parity at the real transform length
        row = np.zeros((c.n_cols, L), np.uint64)
        row[:c.n_per_row] = coeffs[r]
        assert (oenc.encode(row) == comm[r]).all(), r
    hashes = c.hashes
    for col in cols:  # leaf rule on sampled columns
        data = bytes(32) + O.to_repr(field, comm[:, col]).tobytes()
        assert hashes[col].tobytes() == O.blake3(data), col
    np2 = 1 << (c.n_cols - 1).bit_length()
    assert not hashes[c.n_cols:np2].any()
    assert (O.merkle_tree(hashes[:np2]) == hashes).all()  # every inner node and the root
    assert c.get_root().root == hashes[-1].tobytes()
    if full_oracle:
        oc = oenc.commit(x)
        assert oc["root"] == c.get_root().root and (oc["hashes"] == hashes).all() and (oc["comm"] == c.comm).all()


def test_ligero_ft255_2_20_full_oracle():
    """config 2: lcpc-ligero-pc commit, Ft255, 2^20 coefficients -- complete comparison."""
    field, length = P.FT255, 1 << 20
    enc, oenc = P.LigeroEncoding(field, length), O.Encoding.ligero(field, length)
    assert (enc.n_per_row, enc.n_cols) == (16384, 32768)
    x = O.random_elems(field, length, seed=20)
    c = P.LcCommit.commit(x, enc)
    assert c.n_rows == 64
    check_commit_samples(c, oenc, x, field, rows=[0, 63], cols=[0, 1, 32767, 12345], full_oracle=True)
    check_prove_verify(c, enc, field, oc=oenc.commit(x), oenc=oenc, seed=20)


def test_ligero_ft255_2_24_sampled():
    """config 4 on one GPU: 256 x 65536 -> 131072, 2^17-point transforms, 9-chunk leaves."""
    field, length = P.FT255, 1 << 24
    enc, oenc = P.LigeroEncoding(field, length), O.Encoding.ligero(field, length)
    assert (enc.n_per_row, enc.n_cols) == (65536, 131072)
    x = O.random_elems(field, length, seed=24)
    c = P.LcCommit.commit(x, enc)
    assert c.n_rows == 256
    check_commit_samples(c, oenc, x, field, rows=[0, 1, 128, 255], cols=[0, 1, 2, 65535, 65536, 131071, 99999, 31337])
    # linearity of the committed encoding (lcpc-2d/src/tests.rs:193-236): enc(a)+enc(b) == enc(a+b)
    comm = c.comm.reshape(c.n_rows, c.n_cols, 4)
    coeffs = c.coeffs.reshape(c.n_rows, c.n_per_row, 4)
    s = np.zeros((c.n_cols, 4), np.uint64)
    s[:c.n_per_row] = O.field_op(field, "add", coeffs[3], coeffs[200])
    assert (enc.encode(s) == O.field_op(field, "add", comm[3], comm[200])).all()
    # collapse on the full coefficient matrix
    tensor = O.random_elems(field, c.n_rows, seed=5)
    assert (c.collapse(tensor) == O.collapse(field, c.coeffs, tensor, c.n_rows, c.n_per_row)).all()
    vals, paths = c.open_columns([7, 131071])
    for i, col in enumerate([7, 131071]):
        assert (vals[i] == comm[:, col]).all()
        assert O.verify_column_path(field, vals[i], paths[i], col, c.get_root().root)
    check_prove_verify(c, enc, field, seed=24)  # config 4's "commit + prove"


def test_brakedown_ft127_2_24_sampled():
    """config 3: lcpc-brakedown-pc commit, Ft127, 2^24 coefficients (72 x 235173 -> 357699)."""
    field, length = P.FT127, 1 << 24
    enc = P.SdigEncoding(field, length, seed=0)
    assert (enc.n_per_row, enc.n_cols) == (235173, 357699)
    pre, post = enc.matrices()
    oenc = O.Encoding.sdig_from_matrices(field, pre, post)  # same code, no second matgen run
    x = O.random_elems(field, length, seed=3)
    c = P.LcCommit.commit(x, enc)
    assert c.n_rows == 72
    check_commit_samples(c, oenc, x, field, rows=[0, 35, 71], cols=[0, 1, 235172, 235173, 357698, 300000, 41861])
    tensor = O.random_elems(field, c.n_rows, seed=6)
    assert (c.collapse(tensor) == O.collapse(field, c.coeffs, tensor, c.n_rows, c.n_per_row)).all()
    check_prove_verify(c, enc, field, seed=3)  # 6593 column openings, 2 degree tests


def test_ligero_ragged_length_chunked_host_copy():
    """A length that leaves the last row short, large enough that the host->device copy runs in several
    row-chunks overlapped with the encode: the zero padding (lcpc-2d/src/lib.rs:636-645) must land in the
    last chunk only."""
    field, length = P.FT255, (1 << 20) - 12345
    enc = P.LigeroEncoding.new_from_dims(field, 16384, 32768)
    oenc = O.Encoding.ligero_from_dims(field, 16384, 32768)
    x = O.random_elems(field, length, seed=21)
    c = P.LcCommit.commit(x, enc)
    assert c.n_rows == 64
    oc = oenc.commit(x)
    assert oc["root"] == c.get_root().root and (oc["hashes"] == c.hashes).all() and (oc["comm"] == c.comm).all()
    assert (oc["coeffs"] == c.coeffs).all()
    # same object, re-run from the host with different data: no stale rows
    y = O.random_elems(field, length, seed=22)
    c.rerun(y)
    assert oenc.commit(y)["root"] == c.get_root().root


def test_ligero_ft255_2_26_sampled():
    """Beyond the headline size: 512 x 131072 -> 262144 (2^18-point rows, 17-chunk leaves, 2 GiB of coefficients
    crossing PCIe in 16 row-chunks with hashing trailing the encode)."""
    field, length = P.FT255, 1 << 26
    enc, oenc = P.LigeroEncoding(field, length), O.Encoding.ligero(field, length)
    assert (enc.n_per_row, enc.n_cols) == (131072, 262144)
    x = O.random_elems(field, length, seed=26)
    c = P.LcCommit.commit(x, enc)
    assert c.n_rows == 512
    check_commit_samples(c, oenc, x, field, rows=[0, 300, 511], cols=[0, 1, 131071, 131072, 262143, 77777])
    vals, paths = c.open_columns([5, 262143])
    comm = c.comm.reshape(c.n_rows, c.n_cols, 4)
    for i, col in enumerate([5, 262143]):
        assert (vals[i] == comm[:, col]).all()
        assert O.verify_column_path(field, vals[i], paths[i], col, c.get_root().root)


def test_eager_commit_download_overlapped_with_upload():
    """commit_rerun_to_host / commit_to_host: comm and coeffs row-chunks come back on a second copy stream while
    later chunks are uploaded and encoded; the host-visible LcCommit must equal the oracle's (pinned and
    pageable destinations, chunked and single-chunk sizes, ragged last row)."""
    import ctypes as C

    import torch
    from lcpc_b200 import _cabi
    for field, length, pinned in ((P.FT255, (1 << 20) - 999, True), (P.FT255, 1 << 20, False), (P.FT127, 1 << 12, True),
                                  (P.FT63, 100, False)):
        enc, oenc = P.LigeroEncoding(field, length), O.Encoding.ligero(field, length)
        x = O.random_elems(field, length, seed=length % 1000)
        oc = oenc.commit(x)
        n_rows, n_per_row, n_cols = enc.get_dims(length)
        L = enc.L
        n_hashes = 2 * (1 << (n_cols - 1).bit_length()) - 1
        mk = (lambda *shape, dt=torch.int64: torch.empty(shape, dtype=dt, pin_memory=pinned))
        h_comm, h_coef, h_hash = mk(n_rows * n_cols, L), mk(n_rows * n_per_row, L), mk(n_hashes, 32, dt=torch.uint8)
        h_comm.fill_(-1), h_coef.fill_(-1)
        rc = _cabi.lib().lcpc_b200_commit_to_host(enc._h, x.ctypes.data_as(C.c_void_p), length,
                                                  C.c_void_p(h_comm.data_ptr()), C.c_void_p(h_coef.data_ptr()),
                                                  C.c_void_p(h_hash.data_ptr()))
        assert rc == 0
        assert (h_comm.numpy().view(np.uint64) == oc["comm"]).all()
        assert (h_coef.numpy().view(np.uint64) == oc["coeffs"]).all()
        assert (h_hash.numpy() == oc["hashes"]).all()
        # into an existing object, only comm + hashes requested
        c = P.LcCommit.commit(O.random_elems(field, length, seed=1), enc)
        h_comm.fill_(-1)
        c.rerun_to_host(x, comm=h_comm.numpy().view(np.uint64), hashes=h_hash.numpy())
        assert (h_comm.numpy().view(np.uint64) == oc["comm"]).all() and c.get_root().root == oc["root"]


@pytest.mark.gpu
@pytest.mark.parametrize("pre_tail", [1, 0])
def test_ligero_host_copy_chunk_schedule_with_short_last_chunks(pre_tail):
    """The host->device copy of a Ligero commit ends with a half-size and a quarter-size row-chunk (what is left
    after the last byte is pure latency).  Forced here at 2^20 by lowering the CTA thresholds: 8 body chunks + 4 + 2
    rows, ragged last row; same LcCommit as the oracle."""
    from lcpc_b200 import _cabi
    field, length = P.FT255, (1 << 20) - 4321
    enc = P.LigeroEncoding.new_from_dims(field, 16384, 32768)
    oenc = O.Encoding.ligero_from_dims(field, 16384, 32768)
    x = O.random_elems(field, length, seed=31)
    knobs = {b"H2D_MIN_CHUNK_CTAS": (32, 1184), b"H2D_PRE_TAIL_MIN_CTAS": (1, 512), b"H2D_PRE_TAIL": (pre_tail, 1)}
    for k, (v, _) in knobs.items():
        _cabi.lib().lcpc_b200_set_tunable(k, v)
    try:
        c = P.LcCommit.commit(x, enc)
        oc = oenc.commit(x)
        assert oc["root"] == c.get_root().root and (oc["hashes"] == c.hashes).all() and (oc["comm"] == c.comm).all()
        assert (oc["coeffs"] == c.coeffs).all()
        y = O.random_elems(field, length, seed=32)
        c.rerun(y)
        assert oenc.commit(y)["root"] == c.get_root().root
    finally:
        for k, (_, d) in knobs.items():
            _cabi.lib().lcpc_b200_set_tunable(k, d)
