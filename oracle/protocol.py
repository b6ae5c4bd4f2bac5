"""oracle/protocol.py -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU restatement of the reference's prove() and verify() control flow, composed from the oracle's own pieces
(collapse, open_column, verify_column_path, dot, encode, ChaCha20 draws) and the pure-Python transcript:

    prove   lcpc-2d/src/lib.rs:1004-1093
    verify  lcpc-2d/src/lib.rs:832-952  (verify_column_path :954-989, verify_column_value :992-1000)
    wire    bincode 1.x default options over the Wrapped* structs (:186-197, :353-357, :430-437, :551-560)

The wire writer here is deliberately element-by-element (struct.pack per integer), independent of the numpy-based
writer in lcpc_b200/proof.py it is compared with.
"""
from __future__ import annotations

import struct

import numpy as np

import oracle as O
from oracle.transcript import Transcript

# def_labels! (lcpc-2d/src/macros.rs:28-36) does not substitute `$l` inside byte-string literals
LABEL_DT, LABEL_PR, LABEL_PE, LABEL_CO = b"$l//DT", b"$l//PR", b"$l//PE", b"$l//CO"


class VerifierError(Exception):
    """lcpc-2d/src/lib.rs:141-170; ``kind`` is the variant name."""

    def __init__(self, kind):
        super().__init__(kind)
        self.kind = kind


def _key_words(key: bytes):
    return np.frombuffer(key, dtype="<u4")


def sample_columns(key: bytes, n_cols: int, n: int):
    """ChaCha20Rng::from_seed(key), then n x Uniform::new(0usize, n_cols).sample (rand 0.8 UniformInt<usize>:
    64-bit draw v, (hi, lo) = v * range, accept iff lo <= zone, zone = u64::MAX - (2^64 - range) % range)."""
    words, counter = [], 0

    def next_u64():
        nonlocal words, counter
        while len(words) < 2:
            words += [int(w) for w in O.chacha_block(_key_words(key), counter, 0)]
            counter += 1
        lo, hi = words[0], words[1]
        words = words[2:]
        return lo | (hi << 32)

    zone = (1 << 64) - 1 - ((1 << 64) - n_cols) % n_cols
    out = []
    while len(out) < n:
        wide = next_u64() * n_cols
        if wide & ((1 << 64) - 1) <= zone:
            out.append(wide >> 64)
    return out


def _absorb(tr: Transcript, label: bytes, field, elems):
    for r in O.to_repr(field, elems):  # FieldHash::transcript_update, :46-49
        tr.append_message(label, r.tobytes())


def prove(field, commit: dict, outer_tensor, n_degree_tests: int, n_col_opens: int, tr: Transcript) -> dict:
    """`commit` is the dict Encoding.commit returns (comm, coeffs, hashes, n_rows, n_per_row, n_cols)."""
    n_rows, n_per_row, n_cols = commit["n_rows"], commit["n_per_row"], commit["n_cols"]
    assert len(outer_tensor) == n_rows  # ProverError::OuterTensor
    p_random_vec = []
    for _ in range(n_degree_tests):  # :1025-1048
        key = tr.challenge_bytes(LABEL_DT, 32)
        rand_tensor = O.random_elems_from_key(field, key, n_rows)
        p_random = O.collapse(field, commit["coeffs"], rand_tensor, n_rows, n_per_row)
        _absorb(tr, LABEL_PR, field, p_random)
        p_random_vec.append(p_random)
    p_eval = O.collapse(field, commit["coeffs"], outer_tensor, n_rows, n_per_row)  # :1051-1063
    _absorb(tr, LABEL_PE, field, p_eval)
    key = tr.challenge_bytes(LABEL_CO, 32)  # :1066-1085
    cols_to_open = sample_columns(key, n_cols, n_col_opens)
    columns = [O.open_column(field, commit["comm"], commit["hashes"], n_rows, n_cols, c) for c in cols_to_open]
    return dict(n_cols=n_cols, p_eval=p_eval, p_random_vec=p_random_vec, columns=columns, cols_to_open=cols_to_open)


def verify(field, enc, root: bytes, outer_tensor, inner_tensor, proof: dict, tr: Transcript):
    """`enc` is an oracle Encoding.  Returns the evaluation (L limbs) or raises VerifierError."""
    L = O.FIELD_LIMBS[field]
    n_col_opens = enc.get_n_col_opens()
    if n_col_opens != len(proof["columns"]) or n_col_opens == 0:
        raise VerifierError("NumColOpens")
    n_rows = len(proof["columns"][0][0])
    n_cols, n_per_row = proof["n_cols"], len(proof["p_eval"])
    if len(inner_tensor) != n_per_row:
        raise VerifierError("InnerTensor")
    if len(outer_tensor) != n_rows:
        raise VerifierError("OuterTensor")
    if not enc.dims_ok(n_per_row, n_cols):
        raise VerifierError("EncodingDims")

    def encode_padded(v):
        tmp = np.zeros((n_cols, L), np.uint64)
        tmp[:n_per_row] = v
        return enc.encode(tmp)

    rand_tensor_vec, p_random_fft = [], []
    n_degree_tests = enc.get_n_degree_tests()
    for i in range(n_degree_tests):  # :866-898
        key = tr.challenge_bytes(LABEL_DT, 32)
        rand_tensor_vec.append(O.random_elems_from_key(field, key, n_rows))
        p_random_fft.append(encode_padded(proof["p_random_vec"][i]))
        _absorb(tr, LABEL_PR, field, proof["p_random_vec"][i])
    _absorb(tr, LABEL_PE, field, proof["p_eval"])
    key = tr.challenge_bytes(LABEL_CO, 32)  # :901-911
    cols_to_open = sample_columns(key, n_cols, n_col_opens)
    p_eval_fft = encode_padded(proof["p_eval"])  # :914-921
    for col_num, (col, path) in zip(cols_to_open, proof["columns"]):  # :926-942
        rand = True
        for i in range(n_degree_tests):
            rand &= bool((O.dot(field, rand_tensor_vec[i], col) == p_random_fft[i][col_num]).all())
        ev = bool((O.dot(field, outer_tensor, col) == p_eval_fft[col_num]).all())
        pth = O.verify_column_path(field, col, path, col_num, root)
        if not rand:
            raise VerifierError("ColumnDegree")
        if not ev:
            raise VerifierError("ColumnEval")
        if not pth:
            raise VerifierError("ColumnPath")
    return O.dot(field, inner_tensor, proof["p_eval"])  # :944-951


# ------------------------------------------------------------------ wire format, element by element
def _u64(x):
    return struct.pack("<Q", int(x))


def _elem(e):
    return b"".join(_u64(limb) for limb in e)  # struct FtNNN([u64; L]): a tuple of L u64, no length


def _vec(items, f):
    return _u64(len(items)) + b"".join(f(x) for x in items)


def _digest(h):
    b = bytes(bytearray(h))
    return _u64(len(b)) + b  # WrappedOutput { #[serde(with = "serde_bytes")] bytes: Vec<u8> }


def wire_root(root: bytes) -> bytes:
    return _digest(root)


def wire_commit(commit: dict) -> bytes:
    return (_vec(commit["comm"], _elem) + _vec(commit["coeffs"], _elem) + _u64(commit["n_rows"]) +
            _u64(commit["n_cols"]) + _u64(commit["n_per_row"]) + _vec(commit["hashes"], _digest))


def wire_proof(proof: dict) -> bytes:
    def column(cp):
        col, path = cp
        return _vec(col, _elem) + _vec(path, _digest)

    return (_u64(proof["n_cols"]) + _vec(proof["p_eval"], _elem) +
            _vec(proof["p_random_vec"], lambda v: _vec(v, _elem)) + _vec(proof["columns"], column))
