"""prove() / verify() and the wire format: the host-side mirror of ``LcEvalProof`` over the C ABI.

=========================================  ====================================================
reference (lcpc-2d/src/lib.rs)             here
=========================================  ====================================================
``merlin::Transcript`` (:16)               ``Transcript`` (C++ STROBE-128/Keccak, host side)
``LcCommit::prove`` (:304-311, :1004-1093) ``prove(commit, outer_tensor, enc, tr)``
``LcEvalProof`` (:490-500)                 ``LcEvalProof`` (flat numpy arrays)
``LcEvalProof::verify`` (:518-527,         ``LcEvalProof.verify(root, outer, inner, enc, tr)``
:832-952)
``Serialize``/``Deserialize`` of           ``serialize_*`` / ``deserialize_*`` (bincode 1.x default
``LcRoot``/``LcCommit``/``LcEvalProof``    options: little-endian fixed-width integers, u64 lengths)
(:186-268, :353-371, :430-487, :551-609)
=========================================  ====================================================

All field arithmetic, encoding and hashing of prove/verify runs on the GPU through ``lcpc_b200_commit_prove`` /
``lcpc_b200_verify``; the transcript is sequential host work (C++), and the wire format is byte shuffling (numpy).
"""
from __future__ import annotations

import ctypes as C
import struct

import numpy as np

from . import _cabi
from ._cabi import LcpcError
from .host import FIELD_LIMBS, LcCommit, LcEncoding, LcRoot, _check, _elems, _ptr


class Transcript:
    """``merlin::Transcript``: ``new`` / ``append_message`` / ``append_u64`` / ``challenge_bytes``."""

    def __init__(self, label: bytes, _handle=None):
        if _handle is not None:
            self._h = _handle
            return
        self._h = C.c_void_p()
        _check(_cabi.lib().lcpc_b200_transcript_new(bytes(label), len(label), C.byref(self._h)))

    def clone(self) -> "Transcript":
        h = C.c_void_p()
        _check(_cabi.lib().lcpc_b200_transcript_clone(self._h, C.byref(h)))
        return Transcript(b"", _handle=h)

    def append_message(self, label: bytes, message: bytes):
        _check(_cabi.lib().lcpc_b200_transcript_append_message(self._h, bytes(label), len(label), bytes(message), len(message)))

    def append_u64(self, label: bytes, x: int):
        _check(_cabi.lib().lcpc_b200_transcript_append_u64(self._h, bytes(label), len(label), int(x)))

    def challenge_bytes(self, label: bytes, n: int) -> bytes:
        out = np.empty(max(n, 1), np.uint8)
        _check(_cabi.lib().lcpc_b200_transcript_challenge_bytes(self._h, bytes(label), len(label), _ptr(out), n))
        return out[:n].tobytes()

    def append_reprs(self, label: bytes, reprs: np.ndarray):
        """FieldHash::transcript_update (:46-49) for every row of ``reprs`` (canonical little-endian bytes)."""
        a = np.ascontiguousarray(reprs, dtype=np.uint8)
        _check(_cabi.lib().lcpc_b200_transcript_append_reprs(self._h, bytes(label), len(label), _ptr(a), a.shape[1], a.shape[0]))

    def close(self):
        if getattr(self, "_h", None):
            _cabi.lib().lcpc_b200_transcript_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def sample_columns(key: bytes, n_cols: int, n: int) -> np.ndarray:
    """The column challenge (:1073-1080): n x Uniform::new(0usize, n_cols) from ChaCha20Rng::from_seed(key)."""
    if len(key) != 32:
        raise LcpcError(_cabi.ERR_BAD_ARG, "key must be 32 bytes")
    kb = np.frombuffer(bytes(key), dtype=np.uint8).copy()
    out = np.empty(max(n, 1), np.uint64)
    _check(_cabi.lib().lcpc_b200_sample_columns(_ptr(kb), n_cols, n, _ptr(out)))
    return out[:n]


def _labels(enc):
    lb = _cabi.Labels(enc.LABEL_DT, enc.LABEL_PR, enc.LABEL_PE, enc.LABEL_CO, len(enc.LABEL_DT), len(enc.LABEL_PR),
                      len(enc.LABEL_PE), len(enc.LABEL_CO))
    return lb


class LcEvalProof:
    """``LcEvalProof<D, E>`` (:490-500): ``n_cols``, ``p_eval`` (n_per_row, L), ``p_random_vec``
    (n_degree_tests, n_per_row, L), and the opened columns as ``cols`` (n_col_opens, n_rows, L) with their Merkle
    ``paths`` (n_col_opens, path_len, 32) -- ``columns[i] = LcColumn{col: cols[i], path: paths[i]}``."""

    def __init__(self, field, n_cols, p_eval, p_random_vec, cols, paths, col_idx=None):
        self.field, self.n_cols = field, int(n_cols)
        self.p_eval, self.p_random_vec, self.cols, self.paths = p_eval, p_random_vec, cols, paths
        self.col_idx = col_idx  # the prover's view of which columns were opened (not part of the proof)

    def get_n_cols(self) -> int:
        """:507-509."""
        return self.n_cols

    def get_n_per_row(self) -> int:
        """:512-514."""
        return self.p_eval.shape[0]

    def verify(self, root, outer_tensor, inner_tensor, enc: LcEncoding, tr: Transcript) -> np.ndarray:
        """LcEvalProof::verify (:518-527 -> :832-952): returns the evaluation (one element, Montgomery limbs) or
        raises ``LcpcError`` carrying the ``VerifierError`` variant's code."""
        root_b = root.root if isinstance(root, LcRoot) else bytes(root)
        if len(root_b) != 32:
            raise LcpcError(_cabi.ERR_BAD_ARG, "root must be 32 bytes")
        outer, inner = _elems(outer_tensor, enc.field), _elems(inner_tensor, enc.field)
        L = FIELD_LIMBS[enc.field]
        # the C side reads n_open * n_rows * 8L bytes of columns etc.: a proof deserialized for another field (or
        # ragged arrays) must fail here, not as an out-of-bounds read
        if self.field != enc.field:
            raise LcpcError(_cabi.ERR_BAD_ARG, f"proof is over field {self.field}, the encoding over {enc.field}")
        p_eval = np.ascontiguousarray(self.p_eval, np.uint64)
        cols = np.ascontiguousarray(self.cols, np.uint64)
        paths = np.ascontiguousarray(self.paths, np.uint8)
        p_rand = np.ascontiguousarray(self.p_random_vec, np.uint64)
        if p_eval.ndim != 2 or p_eval.shape[1] != L or cols.ndim != 3 or cols.shape[2] != L:
            raise LcpcError(_cabi.ERR_BAD_ARG, "proof arrays do not have the field's limb count")
        if p_rand.size == 0:
            p_rand = p_rand.reshape(0, p_eval.shape[0], L)
        if p_rand.ndim != 3 or p_rand.shape[1:] != (p_eval.shape[0], L):
            raise LcpcError(_cabi.ERR_BAD_ARG, "p_random_vec rows must have n_per_row elements")
        if paths.ndim != 3 or paths.shape[2] != 32 or paths.shape[0] != cols.shape[0]:
            raise LcpcError(_cabi.ERR_BAD_ARG, "paths must be (n_col_opens, path_len, 32)")
        n_columns = cols.shape[0]
        n_rows = cols.shape[1]
        pr = _cabi.Proof(self.n_cols, p_eval.shape[0], p_rand.shape[0], n_columns, n_rows,
                         paths.shape[1], p_eval.ctypes.data, p_rand.ctypes.data,
                         cols.ctypes.data, paths.ctypes.data)
        rb = np.frombuffer(root_b, dtype=np.uint8).copy()
        out = np.empty((1, L), np.uint64)
        lb = _labels(enc)
        _check(_cabi.lib().lcpc_b200_verify(enc._h, tr._h, C.byref(lb), _ptr(rb), _ptr(outer), outer.shape[0],
                                            _ptr(inner), inner.shape[0], enc.get_n_col_opens(),
                                            enc.get_n_degree_tests(), C.byref(pr), _ptr(out)), enc.ctx)
        return out[0]


def prove(commit: LcCommit, outer_tensor, enc: LcEncoding, tr: Transcript) -> LcEvalProof:
    """LcCommit::prove (:304-311 -> :1004-1093)."""
    if commit.enc is not enc and not (enc.n_per_row == commit.n_per_row and enc.n_cols == commit.n_cols and
                                      enc.field == commit.enc.field):
        raise LcpcError(_cabi.ERR_BAD_ARG, "inconsistent commitment fields")  # check_comm, ProverError::Commit (:673-688)
    outer = _elems(outer_tensor, enc.field)
    L = FIELD_LIMBS[enc.field]
    ndt, nco = enc.get_n_degree_tests(), enc.get_n_col_opens()
    path_len = (commit.n_cols - 1).bit_length()
    p_eval = np.empty((commit.n_per_row, L), np.uint64)
    p_rand = np.empty((ndt, commit.n_per_row, L), np.uint64)
    idx = np.empty(nco, np.uint64)
    cols = np.empty((nco, commit.n_rows, L), np.uint64)
    paths = np.empty((nco, path_len, 32), np.uint8)
    lb = _labels(enc)
    _check(_cabi.lib().lcpc_b200_commit_prove(commit._h, tr._h, C.byref(lb), _ptr(outer), outer.shape[0], ndt, nco,
                                              _ptr(p_eval), _ptr(p_rand), _ptr(idx), _ptr(cols), _ptr(paths)), enc.ctx)
    return LcEvalProof(enc.field, commit.n_cols, p_eval, p_rand, cols, paths, col_idx=idx)


# ------------------------------------------------------------------ wire format (bincode 1.x, default options)
# usize -> u64 little-endian; Vec<T> -> u64 length + items; struct -> fields in declaration order; a field element is
# `struct FtNNN([u64; L])` with #[derive(Serialize)] (lcpc-test-fields/src/lib.rs:18,30,42,54): a newtype around a
# fixed-size array = L little-endian u64 MONTGOMERY limbs, no length prefix; a digest is WrappedOutput{bytes} with
# serde_bytes (:353-357) = u64 length + raw bytes.
def _u64(x):
    return struct.pack("<Q", int(x))


def _vec_elems(a: np.ndarray) -> bytes:
    a = np.ascontiguousarray(a, dtype="<u8")
    return _u64(a.shape[0]) + a.tobytes()


def _vec_digests(h: np.ndarray) -> bytes:
    h = np.ascontiguousarray(h, dtype=np.uint8).reshape(-1, 32)
    rec = np.empty((h.shape[0], 40), np.uint8)
    rec[:, :8] = np.frombuffer(_u64(32), dtype=np.uint8)
    rec[:, 8:] = h
    return _u64(h.shape[0]) + rec.tobytes()


def serialize_root(root) -> bytes:
    """LcRoot -> WrappedOutput (:325-331, :353-357)."""
    b = root.root if isinstance(root, LcRoot) else bytes(root)
    return _u64(len(b)) + b


def serialize_commit(c: LcCommit) -> bytes:
    """LcCommit -> WrappedLcCommit{comm, coeffs, n_rows, n_cols, n_per_row, hashes} (:186-197, :226-240)."""
    return (_vec_elems(c.comm) + _vec_elems(c.coeffs) + _u64(c.n_rows) + _u64(c.n_cols) + _u64(c.n_per_row) +
            _vec_digests(c.hashes))


def serialize_proof(p: LcEvalProof) -> bytes:
    """LcEvalProof -> WrappedLcEvalProof{n_cols, p_eval, p_random_vec, columns[{col, path}]} (:551-560, :430-437)."""
    out = [_u64(p.n_cols), _vec_elems(p.p_eval), _u64(p.p_random_vec.shape[0])]
    out += [_vec_elems(v) for v in p.p_random_vec]
    out.append(_u64(p.cols.shape[0]))
    for col, path in zip(p.cols, p.paths):
        out.append(_vec_elems(col))
        out.append(_vec_digests(path))
    return b"".join(out)


class _Reader:
    def __init__(self, data: bytes):
        self.b, self.o = memoryview(data), 0

    def u64(self) -> int:
        if self.o + 8 > len(self.b):
            raise LcpcError(_cabi.ERR_BAD_ARG, "wire: truncated")
        v = struct.unpack_from("<Q", self.b, self.o)[0]
        self.o += 8
        return v

    def take(self, n: int) -> memoryview:
        if n < 0 or self.o + n > len(self.b):
            raise LcpcError(_cabi.ERR_BAD_ARG, "wire: truncated")
        v = self.b[self.o:self.o + n]
        self.o += n
        return v

    def elems(self, L: int) -> np.ndarray:
        n = self.u64()
        return np.frombuffer(self.take(n * 8 * L), dtype="<u8").reshape(n, L).astype(np.uint64)

    def digest(self) -> np.ndarray:
        n = self.u64()
        if n != 32:
            raise LcpcError(_cabi.ERR_BAD_ARG, f"wire: digest of {n} bytes (BLAKE3 output is 32)")
        return np.frombuffer(self.take(32), dtype=np.uint8)

    def digests(self) -> np.ndarray:
        n = self.u64()
        rec = np.frombuffer(self.take(n * 40), dtype=np.uint8).reshape(n, 40)
        if n and not (rec[:, :8] == np.frombuffer(_u64(32), dtype=np.uint8)).all():
            raise LcpcError(_cabi.ERR_BAD_ARG, "wire: digest length != 32")
        return rec[:, 8:].copy()

    def done(self):
        if self.o != len(self.b):
            raise LcpcError(_cabi.ERR_BAD_ARG, f"wire: {len(self.b) - self.o} trailing bytes")


def deserialize_root(data: bytes) -> LcRoot:
    r = _Reader(data)
    root = LcRoot(r.digest().tobytes())
    r.done()
    return root


def deserialize_commit_fields(data: bytes, field: int) -> dict:
    """The LcCommit fields as host arrays (a device-resident LcCommit is rebuilt by committing ``coeffs`` again)."""
    r, L = _Reader(data), FIELD_LIMBS[field]
    comm, coeffs = r.elems(L), r.elems(L)
    n_rows, n_cols, n_per_row = r.u64(), r.u64(), r.u64()
    hashes = r.digests()
    r.done()
    if comm.shape[0] != n_rows * n_cols or coeffs.shape[0] != n_rows * n_per_row:
        raise LcpcError(_cabi.ERR_BAD_ARG, "wire: inconsistent commitment fields")  # check_comm (:673-688)
    return dict(comm=comm, coeffs=coeffs, n_rows=n_rows, n_cols=n_cols, n_per_row=n_per_row, hashes=hashes)


def deserialize_commit(data: bytes, enc: LcEncoding) -> LcCommit:
    """``bincode::deserialize::<LcCommit<D, E>>``: the wire image back into a device-resident commit for ``enc``."""
    f = deserialize_commit_fields(data, enc.field)
    if (f["n_per_row"], f["n_cols"]) != (enc.n_per_row, enc.n_cols):
        raise LcpcError(_cabi.ERR_BAD_ARG, "wire: commitment dimensions do not match the encoding")
    return LcCommit.from_fields(enc, f["comm"], f["coeffs"], f["hashes"], f["n_rows"])


def deserialize_proof(data: bytes, field: int) -> LcEvalProof:
    r, L = _Reader(data), FIELD_LIMBS[field]
    n_cols = r.u64()
    p_eval = r.elems(L)
    p_rand = [r.elems(L) for _ in range(r.u64())]
    cols, paths = [], []
    for _ in range(r.u64()):
        cols.append(r.elems(L))
        paths.append(r.digests())
    r.done()
    if any(v.shape != p_eval.shape for v in p_rand) or any(c.shape != cols[0].shape for c in cols) or \
            any(p.shape != paths[0].shape for p in paths):
        raise LcpcError(_cabi.ERR_BAD_ARG, "wire: ragged proof")
    p_rand_a = np.stack(p_rand) if p_rand else np.empty((0,) + p_eval.shape, np.uint64)
    cols_a = np.stack(cols) if cols else np.empty((0, 0, L), np.uint64)
    paths_a = np.stack(paths) if paths else np.empty((0, 0, 32), np.uint8)
    return LcEvalProof(field, n_cols, p_eval, p_rand_a, cols_a, paths_a)
