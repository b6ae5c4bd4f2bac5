"""oracle -- TEST INFRASTRUCTURE ONLY.

ctypes binding of ``oracle/liblcpc_oracle.so`` (the CPU restatement of the reference's
commit/prove hot path, see ``oracle/lcpc_oracle.h``).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this package; the
product package ``lcpc_b200`` never does.

Field elements travel as ``numpy.uint64`` arrays of shape ``(n, L)``: L little-endian limbs in
Montgomery form, i.e. the in-memory image of the reference's ``struct FtNNN([u64; L])``
(lcpc-test-fields/src/lib.rs:22,34,46,58).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liblcpc_oracle.so")

FT63, FT127, FT191, FT255 = 1, 2, 3, 4
ENC_LIGERO, ENC_SDIG = 1, 2
FIELD_LIMBS = {FT63: 1, FT127: 2, FT191: 3, FT255: 4}
FIELD_NAMES = {FT63: "Ft63", FT127: "Ft127", FT191: "Ft191", FT255: "Ft255"}


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc + OpenMP)."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []))
    return _LIB_PATH


_lib = None

_u64p = C.POINTER(C.c_uint64)
_u8p = C.POINTER(C.c_uint8)
_szp = C.POINTER(C.c_size_t)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.lcpc_oracle_ligero_new_from_dims.restype = C.c_void_p
        L.lcpc_oracle_ligero_new.restype = C.c_void_p
        L.lcpc_oracle_sdig_new.restype = C.c_void_p
        L.lcpc_oracle_sdig_new_from_dims.restype = C.c_void_p
        L.lcpc_oracle_sdig_from_matrices.restype = C.c_void_p
        for name in ("lcpc_oracle_n_degree_tests", "lcpc_oracle_ligero_n_col_opens",
                     "lcpc_oracle_sdig_n_col_opens", "lcpc_oracle_enc_n_col_opens",
                     "lcpc_oracle_enc_n_degree_tests", "lcpc_oracle_sdig_n_levels"):
            getattr(L, name).restype = C.c_size_t
        _lib = L
    return _lib


def _p64(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_u64p)


def _p8(a: np.ndarray):
    assert a.dtype == np.uint8 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_u8p)


def _sz(v):
    return C.c_size_t(int(v))


def _elems(a, field) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    L = FIELD_LIMBS[field]
    if a.ndim == 1:
        a = a.reshape(-1, L)
    assert a.shape[-1] == L, (a.shape, L)
    return a


# ----------------------------------------------------------------------------- fields
def field_info(field):
    L = FIELD_LIMBS[field]
    nb, s, inv = C.c_uint32(), C.c_uint32(), C.c_uint64()
    mod, r, r2, rou = (np.zeros(L, np.uint64) for _ in range(4))
    rc = lib().lcpc_oracle_field_info(field, C.byref(nb), C.byref(s), _p64(mod), _p64(r), _p64(r2),
                                      C.byref(inv), _p64(rou))
    assert rc == 0
    return dict(limbs=L, num_bits=nb.value, s=s.value, modulus=limbs_to_int(mod), r=limbs_to_int(r),
                r2=limbs_to_int(r2), inv=inv.value, rou_mont=limbs_to_int(rou))


def limbs_to_int(limbs) -> int:
    v = 0
    for i, l in enumerate(np.asarray(limbs, dtype=np.uint64).tolist()):
        v |= int(l) << (64 * i)
    return v


def int_to_limbs(v: int, L: int) -> np.ndarray:
    return np.array([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(L)], dtype=np.uint64)


def ints_to_elems(vals, field) -> np.ndarray:
    """Plain integers -> (n, L) limb array (no Montgomery conversion)."""
    L = FIELD_LIMBS[field]
    out = np.zeros((len(vals), L), np.uint64)
    for i, v in enumerate(vals):
        out[i] = int_to_limbs(int(v), L)
    return out


def elems_to_ints(a) -> list:
    return [limbs_to_int(row) for row in np.asarray(a, dtype=np.uint64)]


_OPS = {"add": 0, "sub": 1, "mul": 2, "to_mont": 3, "from_mont": 4, "inv": 5}


def field_op(field, op, a, b=None) -> np.ndarray:
    a = _elems(a, field)
    out = np.empty_like(a)
    bp = _p64(_elems(b, field)) if b is not None else None
    rc = lib().lcpc_oracle_field_op(field, _OPS[op], _p64(out), _p64(a), bp, _sz(a.shape[0]))
    assert rc == 0, rc
    return out


def to_mont(field, ints) -> np.ndarray:
    return field_op(field, "to_mont", ints_to_elems(ints, field))


def from_mont(field, a) -> list:
    return elems_to_ints(field_op(field, "from_mont", a))


def to_repr(field, a) -> np.ndarray:
    a = _elems(a, field)
    out = np.empty((a.shape[0], 8 * FIELD_LIMBS[field]), np.uint8)
    assert lib().lcpc_oracle_to_repr(field, _p8(out), _p64(a), _sz(a.shape[0])) == 0
    return out


def random_elems(field, n, seed=0, stream=0) -> np.ndarray:
    out = np.empty((n, FIELD_LIMBS[field]), np.uint64)
    rc = lib().lcpc_oracle_random_elems(field, C.c_uint64(seed), C.c_uint64(stream), _p64(out), _sz(n))
    assert rc == 0
    return out


def random_elems_from_key(field, key: bytes, n) -> np.ndarray:
    """n x Field::random from ChaCha20Rng::from_seed(key) (lcpc-2d/src/lib.rs:1026-1032)."""
    assert len(key) == 32
    out = np.empty((n, FIELD_LIMBS[field]), np.uint64)
    kb = np.frombuffer(bytes(key), dtype=np.uint8).copy()
    assert lib().lcpc_oracle_random_elems_from_key(field, _p8(kb), _p64(out), _sz(n)) == 0
    return out


# ----------------------------------------------------------------------------- hash / rng
def blake3(data: bytes) -> bytes:
    buf = np.frombuffer(bytes(data), dtype=np.uint8) if len(data) else np.zeros(0, np.uint8)
    buf = np.ascontiguousarray(buf)
    out = np.empty(32, np.uint8)
    lib().lcpc_oracle_blake3(_p8(buf) if len(data) else None, _sz(len(data)), _p8(out))
    return out.tobytes()


def chacha_block(key_words, counter, stream) -> np.ndarray:
    key = np.ascontiguousarray(key_words, dtype=np.uint32)
    out = np.empty(16, np.uint32)
    lib().lcpc_oracle_chacha_block(key.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_uint64(counter),
                                   C.c_uint64(stream), out.ctypes.data_as(C.POINTER(C.c_uint32)))
    return out


# ----------------------------------------------------------------------------- NTT
def fft_io(field, x) -> np.ndarray:
    x = _elems(x, field).copy()
    rc = lib().lcpc_oracle_fft_io(field, _p64(x), _sz(x.shape[0]))
    if rc:
        raise ValueError(f"FFTError rc={rc}")
    return x


def ifft_oi(field, x) -> np.ndarray:
    x = _elems(x, field).copy()
    rc = lib().lcpc_oracle_ifft_oi(field, _p64(x), _sz(x.shape[0]))
    if rc:
        raise ValueError(f"FFTError rc={rc}")
    return x


def root_of_unity(field, length) -> np.ndarray:
    w = np.empty(FIELD_LIMBS[field], np.uint64)
    rc = lib().lcpc_oracle_root_of_unity(field, _sz(length), _p64(w))
    if rc:
        raise ValueError(f"FFTError rc={rc}")
    return w


# ----------------------------------------------------------------------------- parameters
def n_degree_tests(lam, length, flog2) -> int:
    return lib().lcpc_oracle_n_degree_tests(_sz(lam), _sz(length), _sz(flog2))


def ligero_get_dims(field, length, rho=(1, 2)):
    nr, np_, nc = C.c_size_t(), C.c_size_t(), C.c_size_t()
    rc = lib().lcpc_oracle_ligero_get_dims(field, _sz(length), _sz(rho[0]), _sz(rho[1]), C.byref(nr),
                                           C.byref(np_), C.byref(nc))
    if rc:
        raise ValueError(f"ligero_get_dims rc={rc}")
    return nr.value, np_.value, nc.value


def sdig_level_dims(field, code, n):
    pre = ((C.c_size_t * 3) * 32)()
    post = ((C.c_size_t * 3) * 32)()
    lv = lib().lcpc_oracle_sdig_level_dims(field, code, _sz(n), _sz(32), pre, post)
    if lv <= 0:
        raise ValueError(f"sdig_level_dims rc={lv}")
    return [tuple(pre[i]) for i in range(lv)], [tuple(post[i]) for i in range(lv)]


# ----------------------------------------------------------------------------- encodings
class Encoding:
    """Oracle twin of an ``impl LcEncoding`` (lcpc-2d/src/lib.rs:74-104)."""

    def __init__(self, handle):
        if not handle:
            raise ValueError("oracle: could not construct encoding (bad dims?)")
        self._h = C.c_void_p(handle)
        self.field = lib().lcpc_oracle_enc_field(self._h)
        self.kind = lib().lcpc_oracle_enc_kind(self._h)
        self.L = FIELD_LIMBS[self.field]
        _, self.n_per_row, self.n_cols = self.get_dims(1)

    def __del__(self):
        try:
            if self._h:
                lib().lcpc_oracle_enc_free(self._h)
                self._h = None
        except Exception:
            pass

    # constructors
    @classmethod
    def ligero(cls, field, length, rho=(1, 2)):
        return cls(lib().lcpc_oracle_ligero_new(field, _sz(length), _sz(rho[0]), _sz(rho[1])))

    @classmethod
    def ligero_from_dims(cls, field, n_per_row, n_cols, rho=(1, 2)):
        return cls(lib().lcpc_oracle_ligero_new_from_dims(field, _sz(n_per_row), _sz(n_cols),
                                                          _sz(rho[0]), _sz(rho[1])))

    @classmethod
    def sdig(cls, field, length, seed=0, code=3):
        return cls(lib().lcpc_oracle_sdig_new(field, code, _sz(length), C.c_uint64(seed)))

    @classmethod
    def sdig_from_dims(cls, field, n_per_row, n_cols=0, seed=0, code=3):
        return cls(lib().lcpc_oracle_sdig_new_from_dims(field, code, _sz(n_per_row), _sz(n_cols),
                                                        C.c_uint64(seed)))

    @classmethod
    def sdig_from_matrices(cls, field, pre, post, code=3):
        """pre/post: lists of dicts with m, n, ptrs, idxs, data ((nnz, L) uint64)."""
        nlev = len(pre)
        keep = []

        def pack(mats):
            ms = (C.c_size_t * nlev)(*[int(M["m"]) for M in mats])
            ns = (C.c_size_t * nlev)(*[int(M["n"]) for M in mats])
            arrs = []
            for key in ("ptrs", "idxs", "data"):
                col = [np.ascontiguousarray(M[key], dtype=np.uint64) for M in mats]
                keep.extend(col)
                arrs.append((_u64p * nlev)(*[_p64(a.reshape(-1)) for a in col]))
            return ms, ns, arrs

        pm, pn, pa = pack(pre)
        qm, qn, qa = pack(post)
        h = lib().lcpc_oracle_sdig_from_matrices(field, code, _sz(nlev), pm, pn, pa[0], pa[1], pa[2],
                                                 qm, qn, qa[0], qa[1], qa[2])
        return cls(h)

    # LcEncoding
    def get_dims(self, length):
        nr, np_, nc = C.c_size_t(), C.c_size_t(), C.c_size_t()
        lib().lcpc_oracle_enc_get_dims(self._h, _sz(length), C.byref(nr), C.byref(np_), C.byref(nc))
        return nr.value, np_.value, nc.value

    def dims_ok(self, n_per_row, n_cols) -> bool:
        return bool(lib().lcpc_oracle_enc_dims_ok(self._h, _sz(n_per_row), _sz(n_cols)))

    def get_n_col_opens(self) -> int:
        return lib().lcpc_oracle_enc_n_col_opens(self._h)

    def get_n_degree_tests(self) -> int:
        return lib().lcpc_oracle_enc_n_degree_tests(self._h)

    def encode(self, row) -> np.ndarray:
        row = _elems(row, self.field).copy()
        assert row.shape[0] == self.n_cols
        rc = lib().lcpc_oracle_encode(self._h, _p64(row))
        if rc:
            raise ValueError(f"encode rc={rc}")
        return row

    # brakedown matrices
    @property
    def n_levels(self) -> int:
        return lib().lcpc_oracle_sdig_n_levels(self._h)

    def matrix(self, level, is_post):
        m, n, nnz = C.c_size_t(), C.c_size_t(), C.c_size_t()
        ptrs, idxs, data = _u64p(), _u64p(), _u64p()
        rc = lib().lcpc_oracle_sdig_matrix(self._h, _sz(level), int(is_post), C.byref(m), C.byref(n),
                                           C.byref(nnz), C.byref(ptrs), C.byref(idxs), C.byref(data))
        assert rc == 0
        return dict(m=m.value, n=n.value,
                    ptrs=np.ctypeslib.as_array(ptrs, shape=(n.value + 1,)).copy(),
                    idxs=np.ctypeslib.as_array(idxs, shape=(max(nnz.value, 1),))[:nnz.value].copy(),
                    data=np.ctypeslib.as_array(data, shape=(max(nnz.value, 1), self.L))[:nnz.value].copy())

    def matrices(self):
        nl = self.n_levels
        return ([self.matrix(i, False) for i in range(nl)], [self.matrix(i, True) for i in range(nl)])

    # commit
    def commit(self, coeffs_in, threads=0):
        """LcCommit::commit (lcpc-2d/src/lib.rs:299-301, 622-671) -> dict of LcCommit fields."""
        coeffs_in = _elems(coeffs_in, self.field)
        length = coeffs_in.shape[0]
        n_rows, n_per_row, n_cols = self.get_dims(length)
        np2 = 1 << (n_cols - 1).bit_length()
        comm = np.empty((n_rows * n_cols, self.L), np.uint64)
        coeffs = np.empty((n_rows * n_per_row, self.L), np.uint64)
        hashes = np.empty((2 * np2 - 1, 32), np.uint8)
        rc = lib().lcpc_oracle_commit(self._h, _p64(coeffs_in), _sz(length), _p64(comm), _p64(coeffs),
                                      _p8(hashes), int(threads))
        if rc:
            raise ValueError(f"commit rc={rc}")
        return dict(comm=comm, coeffs=coeffs, hashes=hashes, n_rows=n_rows, n_per_row=n_per_row,
                    n_cols=n_cols, root=hashes[-1].tobytes())


def merkleize(field, comm, n_rows, n_cols, serial=False, threads=0) -> np.ndarray:
    comm = _elems(comm, field)
    np2 = 1 << (n_cols - 1).bit_length()
    hashes = np.empty((2 * np2 - 1, 32), np.uint8)
    rc = lib().lcpc_oracle_merkleize(field, _p64(comm), _sz(n_rows), _sz(n_cols), _p8(hashes),
                                     int(serial), int(threads))
    assert rc == 0, rc
    return hashes


def merkle_tree(leaves) -> np.ndarray:
    leaves = np.ascontiguousarray(leaves, dtype=np.uint8).reshape(-1, 32)
    np2 = leaves.shape[0]
    assert np2 & (np2 - 1) == 0
    hashes = np.zeros((2 * np2 - 1, 32), np.uint8)
    hashes[:np2] = leaves
    lib().lcpc_oracle_merkle_tree(_p8(hashes), _sz(np2), 0)
    return hashes


def collapse(field, coeffs, tensor, n_rows, n_per_row, serial=False, threads=0) -> np.ndarray:
    coeffs = _elems(coeffs, field)
    tensor = _elems(tensor, field)
    assert coeffs.shape[0] == n_rows * n_per_row and tensor.shape[0] == n_rows
    poly = np.empty((n_per_row, FIELD_LIMBS[field]), np.uint64)
    rc = lib().lcpc_oracle_collapse(field, _p64(coeffs), _p64(tensor), _p64(poly), _sz(n_rows),
                                    _sz(n_per_row), int(serial), int(threads))
    assert rc == 0, rc
    return poly


def open_column(field, comm, hashes, n_rows, n_cols, column):
    comm = _elems(comm, field)
    hashes = np.ascontiguousarray(hashes, dtype=np.uint8)
    col = np.empty((n_rows, FIELD_LIMBS[field]), np.uint64)
    path = np.empty((64, 32), np.uint8)
    rc = lib().lcpc_oracle_open_column(field, _p64(comm), _p8(hashes), _sz(n_rows), _sz(n_cols),
                                       _sz(column), _p64(col), _p8(path))
    if rc < 0:
        raise IndexError("ProverError::ColumnNumber")
    return col, path[:rc].copy()


def verify_column_path(field, col, path, col_num, root: bytes) -> bool:
    col = _elems(col, field)
    path = np.ascontiguousarray(path, dtype=np.uint8).reshape(-1, 32)
    rootb = np.frombuffer(root, dtype=np.uint8).copy()
    rc = lib().lcpc_oracle_verify_column_path(field, _p64(col), _sz(col.shape[0]), _p8(path),
                                              _sz(path.shape[0]), _sz(col_num), _p8(rootb))
    assert rc >= 0
    return bool(rc)


def dot(field, a, b) -> np.ndarray:
    a, b = _elems(a, field), _elems(b, field)
    out = np.empty(FIELD_LIMBS[field], np.uint64)
    assert lib().lcpc_oracle_dot(field, _p64(a), _p64(b), _sz(a.shape[0]), _p64(out)) == 0
    return out


def max_threads() -> int:
    return lib().lcpc_oracle_max_threads()
