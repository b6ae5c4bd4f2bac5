"""Host layer: Python mirror of the reference's LcEncoding / LcCommit interface over the C ABI.

Every method cites the reference item it stands for (paths relative to the reference repo).  Error
behaviour follows the reference: what is a ``ProverError`` / ``assert!`` there is an ``LcpcError``
(with the C status code) here.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _cabi
from ._cabi import Csc, LcpcError

FT63, FT127, FT191, FT255 = 1, 2, 3, 4
FIELD_LIMBS = {FT63: 1, FT127: 2, FT191: 3, FT255: 4}
LAMBDA = 128  # lcpc-ligero-pc/src/lib.rs:45, lcpc-brakedown-pc/src/lib.rs:54


def _check(rc, ctx=None):
    if rc != _cabi.OK:
        msg = ""
        if ctx is not None and ctx._h:
            msg = (_cabi.lib().lcpc_b200_last_error(ctx._h) or b"").decode()
        raise LcpcError(rc, msg)


def _elems(a, field) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    L = FIELD_LIMBS[field]
    if a.ndim == 1:
        a = a.reshape(-1, L)
    if a.shape[-1] != L:
        raise ValueError(f"expected (n, {L}) limbs, got {a.shape}")
    return a


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Context:
    """One CUDA device + stream (``lcpc_b200_ctx``)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        _check(_cabi.lib().lcpc_b200_ctx_create(int(device), C.byref(self._h)))
        self.device = device

    def close(self):
        if self._h:
            _cabi.lib().lcpc_b200_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self) -> int:
        """Raw ``cudaStream_t`` the context enqueues on (for CUDA-event timing)."""
        return _cabi.lib().lcpc_b200_ctx_stream(self._h) or 0

    def synchronize(self):
        _check(_cabi.lib().lcpc_b200_ctx_synchronize(self._h), self)

    @property
    def launch_count(self) -> int:
        return _cabi.lib().lcpc_b200_ctx_launch_count(self._h)


_default_ctx = {}


def default_context(device: int = 0) -> Context:
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


# ------------------------------------------------------------------ protocol parameters (host)
def n_degree_tests(lam: int, length: int, flog2: int) -> int:
    """lcpc-2d/src/lib.rs:613-616."""
    return _cabi.lib().lcpc_b200_n_degree_tests(lam, length, flog2)


def ligero_get_dims(field: int, length: int, rho=(1, 2)):
    """LigeroEncodingRho::_get_dims, lcpc-ligero-pc/src/lib.rs:70-112."""
    nr, npr, nc = C.c_size_t(), C.c_size_t(), C.c_size_t()
    _check(_cabi.lib().lcpc_b200_ligero_get_dims(field, length, rho[0], rho[1], C.byref(nr), C.byref(npr), C.byref(nc)))
    return nr.value, npr.value, nc.value


# ------------------------------------------------------------------ encodings
class LcEncoding:
    """``trait LcEncoding`` (lcpc-2d/src/lib.rs:74-104) over a device-side ``lcpc_b200_enc``."""

    # def_labels! does not substitute `$l` inside the byte-string literal (lcpc-2d/src/macros.rs:28-36):
    # every encoding's labels are literally these
    LABEL_DT, LABEL_PR, LABEL_PE, LABEL_CO = b"$l//DT", b"$l//PR", b"$l//PE", b"$l//CO"

    def __init__(self, ctx: Context, handle, field: int):
        self.ctx, self._h, self.field = ctx, handle, field
        self.L = FIELD_LIMBS[field]
        _, self.n_per_row, self.n_cols = self.get_dims(1)

    def close(self):
        if getattr(self, "_h", None):
            _cabi.lib().lcpc_b200_enc_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def get_dims(self, length: int):
        """LcEncoding::get_dims (:94)."""
        nr, npr, nc = C.c_size_t(), C.c_size_t(), C.c_size_t()
        _check(_cabi.lib().lcpc_b200_enc_get_dims(self._h, length, C.byref(nr), C.byref(npr), C.byref(nc)), self.ctx)
        return nr.value, npr.value, nc.value

    def dims_ok(self, n_per_row: int, n_cols: int) -> bool:
        """LcEncoding::dims_ok (:97)."""
        return bool(_cabi.lib().lcpc_b200_enc_dims_ok(self._h, n_per_row, n_cols))

    def get_n_degree_tests(self) -> int:
        """LcEncoding::get_n_degree_tests (:103)."""
        return n_degree_tests(LAMBDA, self.n_cols, _cabi.lib().lcpc_b200_field_flog2(self.field))

    def get_n_col_opens(self) -> int:  # pragma: no cover - overridden
        raise NotImplementedError

    def encode(self, inp: np.ndarray) -> np.ndarray:
        """LcEncoding::encode (:91), in place semantics returned as a new array.  ``inp`` is one row
        ``(n_cols, L)`` or a batch ``(n_rows, n_cols, L)``."""
        a = np.ascontiguousarray(inp, dtype=np.uint64)
        single = a.ndim == 2
        rows = a.reshape(1, *a.shape) if single else a
        if rows.ndim != 3 or rows.shape[1] != self.n_cols or rows.shape[2] != self.L:
            raise LcpcError(_cabi.ERR_ENCODE, f"encode: row shape {a.shape} does not match n_cols={self.n_cols}")
        out = rows.copy()
        _check(_cabi.lib().lcpc_b200_encode(self._h, _ptr(out), out.shape[0]), self.ctx)
        return out[0] if single else out


class LigeroEncoding(LcEncoding):
    """``LigeroEncodingRho<Ft, Rn, Rd>`` (lcpc-ligero-pc/src/lib.rs:31-37); rho defaults to 1/2 like
    the public alias ``LigeroEncoding<F>`` (:189)."""

    def __init__(self, field: int, length: int, rho=(1, 2), ctx: Context | None = None):
        """LigeroEncodingRho::new (:121-124)."""
        _, n_per_row, n_cols = ligero_get_dims(field, length, rho)
        self._init_dims(field, n_per_row, n_cols, rho, ctx)

    @classmethod
    def new_ml(cls, field: int, n_vars: int, rho=(1, 2), ctx: Context | None = None):
        """LigeroEncodingRho::new_ml (:126-135): a multilinear polynomial with 2^n_vars monomials."""
        n_rows, n_per_row, n_cols = ligero_get_dims(field, 1 << n_vars, rho)
        if n_rows & (n_rows - 1) or n_per_row & (n_per_row - 1) or n_rows * n_per_row != 1 << n_vars:  # asserts :131-133
            raise LcpcError(_cabi.ERR_BAD_ARG, "new_ml: dimensions are not powers of two")
        return cls.new_from_dims(field, n_per_row, n_cols, rho, ctx)

    @classmethod
    def new_from_dims(cls, field: int, n_per_row: int, n_cols: int, rho=(1, 2), ctx: Context | None = None):
        """LigeroEncodingRho::new_from_dims (:138-148)."""
        self = cls.__new__(cls)
        self._init_dims(field, n_per_row, n_cols, rho, ctx)
        return self

    def _init_dims(self, field, n_per_row, n_cols, rho, ctx):
        ctx = ctx or default_context()
        h = C.c_void_p()
        _check(_cabi.lib().lcpc_b200_ligero_new(ctx._h, field, n_per_row, n_cols, C.byref(h)), ctx)
        self.rho = rho
        LcEncoding.__init__(self, ctx, h, field)

    def get_n_col_opens(self) -> int:
        """lcpc-ligero-pc/src/lib.rs:61-64, 179-181."""
        return _cabi.lib().lcpc_b200_ligero_n_col_opens(self.rho[0], self.rho[1])


class SdigEncoding(LcEncoding):
    """``SdigEncodingS<Ft, S>`` (lcpc-brakedown-pc/src/lib.rs:40-47); code 3 = ``SdigCode3``, the
    default alias ``SdigEncoding<F>`` (:19, :179)."""

    def __init__(self, field: int, length: int, seed: int = 0, code: int = 3, ctx: Context | None = None):
        """SdigEncodingS::new (:103-110)."""
        npr = C.c_size_t()
        _check(_cabi.lib().lcpc_b200_sdig_choose_n_per_row(field, code, length, C.byref(npr)))
        self._init_dims(field, npr.value, 0, seed, code, ctx)

    @classmethod
    def new_ml(cls, field: int, n_vars: int, seed: int = 0, code: int = 3, ctx: Context | None = None):
        """SdigEncodingS::new_ml (:114-124): 2^n_vars monomials, n_per_row a power of two."""
        npr = C.c_size_t()
        _check(_cabi.lib().lcpc_b200_sdig_choose_n_per_row_ml(field, code, n_vars, C.byref(npr)))
        return cls.new_from_dims(field, npr.value, 0, seed, code, ctx)

    @classmethod
    def new_from_dims(cls, field: int, n_per_row: int, n_cols: int = 0, seed: int = 0, code: int = 3,
                      ctx: Context | None = None):
        """SdigEncodingS::new_from_dims (:126-137); n_cols = 0 skips the codeword-length assert."""
        self = cls.__new__(cls)
        self._init_dims(field, n_per_row, n_cols, seed, code, ctx)
        return self

    @classmethod
    def from_matrices(cls, field: int, pre, post, code: int = 3, ctx: Context | None = None):
        """Around caller-supplied CSC matrices (what a Rust host passes after its own matgen):
        lists of dicts with m, n, ptrs, idxs, data."""
        self = cls.__new__(cls)
        ctx = ctx or default_context()
        keep = []

        def pack(mats):
            arr = (Csc * len(mats))()
            for i, M in enumerate(mats):
                cols = [np.ascontiguousarray(M[k], dtype=np.uint64).reshape(-1) for k in ("ptrs", "idxs", "data")]
                keep.extend(cols)
                arr[i] = Csc(int(M["m"]), int(M["n"]), *[c.ctypes.data for c in cols])
            return arr

        h = C.c_void_p()
        _check(_cabi.lib().lcpc_b200_sdig_new(ctx._h, field, len(pre), pack(pre), pack(post), C.byref(h)), ctx)
        self.code, self.seed, self._code_h = code, None, None
        # keep the caller's matrices: matrices() of an encoding built this way returns exactly what it was given
        self._given = ([dict(M) for M in pre], [dict(M) for M in post])
        LcEncoding.__init__(self, ctx, h, field)
        return self

    def _init_dims(self, field, n_per_row, n_cols, seed, code, ctx):
        ctx = ctx or default_context()
        if _device_matgen_enabled():
            # matgen::generate on the device (csrc/device_matgen.cu), cached per context: no host matrices at all
            h = C.c_void_p()
            _check(_cabi.lib().lcpc_b200_sdig_new_seeded(ctx._h, field, code, n_per_row, seed, C.byref(h)), ctx)
            self.code, self.seed, self._code_h, self._given = code, seed, None, None
            LcEncoding.__init__(self, ctx, h, field)
            if n_cols and self.n_cols != n_cols:  # assert_eq! at :129
                cw = self.n_cols
                self.close()
                raise LcpcError(_cabi.ERR_BAD_ARG, f"codeword length {cw} != n_cols {n_cols}")
            return
        ch = C.c_void_p()
        _check(_cabi.lib().lcpc_b200_sdig_code_generate(field, code, n_per_row, seed, C.byref(ch)))
        self._code_h = ch
        cw = _cabi.lib().lcpc_b200_sdig_code_codeword_length(ch)
        if n_cols and cw != n_cols:  # assert_eq! at :129
            raise LcpcError(_cabi.ERR_BAD_ARG, f"codeword length {cw} != n_cols {n_cols}")
        h = C.c_void_p()
        _check(_cabi.lib().lcpc_b200_sdig_new_from_code(ctx._h, ch, C.byref(h)), ctx)
        self.code, self.seed = code, seed
        LcEncoding.__init__(self, ctx, h, field)

    def close(self):
        LcEncoding.close(self)
        if getattr(self, "_code_h", None):
            _cabi.lib().lcpc_b200_sdig_code_free(self._code_h)
            self._code_h = None

    def get_n_col_opens(self) -> int:
        """lcpc-brakedown-pc/src/lib.rs:57-61, 169-171."""
        return _cabi.lib().lcpc_b200_sdig_n_col_opens(self.code)

    def matrices(self):
        """(precodes, postcodes) as lists of dicts -- the CSC arrays matgen::generate produced."""
        if getattr(self, "_given", None) is not None:
            return self._given
        if not self._code_h:
            if not self._h:
                raise LcpcError(_cabi.ERR_BAD_ARG, "encoding is closed: no code matrices")
            return device_code_matrices(self, self.L)
        return host_code_matrices(self._code_h, self.L)


def _device_matgen_enabled() -> bool:
    """LCPC_B200_MATGEN=host keeps the round-1 path (host generator + upload) for A/B runs and for the tests that
    compare the two generators; default: on the device."""
    import os
    return os.environ.get("LCPC_B200_MATGEN", "device") != "host"


def device_code_matrices(enc, L):
    """The matrices of a device-generated code, downloaded in the reference's CSC form (matgen.rs:187)."""
    lib = _cabi.lib()
    t = lib.lcpc_b200_enc_sdig_levels(enc._h)
    out = ([], [])
    for is_post in (0, 1):
        for i in range(t):
            m, n, d = C.c_size_t(), C.c_size_t(), C.c_size_t()
            _check(lib.lcpc_b200_enc_sdig_matrix(enc._h, i, is_post, C.byref(m), C.byref(n), C.byref(d), None, None), enc.ctx)
            nnz = n.value * d.value
            idxs, data = np.empty(nnz, np.uint64), np.empty((nnz, L), np.uint64)
            _check(lib.lcpc_b200_enc_sdig_matrix(enc._h, i, is_post, None, None, None, _ptr(idxs), _ptr(data)), enc.ctx)
            ptrs = np.arange(n.value + 1, dtype=np.uint64) * np.uint64(d.value)
            out[is_post].append(dict(m=m.value, n=n.value, ptrs=ptrs, idxs=idxs, data=data))
    return out


def host_code_matrices(code_h, L):
    lib = _cabi.lib()
    t = lib.lcpc_b200_sdig_code_levels(code_h)
    out = ([], [])
    for is_post in (0, 1):
        for i in range(t):
            m = Csc()
            _check(lib.lcpc_b200_sdig_code_matrix(code_h, i, is_post, C.byref(m)))
            ptrs = np.ctypeslib.as_array(C.cast(m.ptrs, C.POINTER(C.c_uint64)), shape=(m.n + 1,)).copy()
            nnz = int(ptrs[-1])
            if nnz == 0 or not m.idxs or not m.data:  # an empty level may hand back NULL data pointers
                idxs, data = np.zeros(0, np.uint64), np.zeros((0, L), np.uint64)
            else:
                idxs = np.ctypeslib.as_array(C.cast(m.idxs, C.POINTER(C.c_uint64)), shape=(nnz,)).copy()
                data = np.ctypeslib.as_array(C.cast(m.data, C.POINTER(C.c_uint64)), shape=(nnz * L,)).reshape(nnz, L).copy()
            out[is_post].append(dict(m=m.m, n=m.n, ptrs=ptrs, idxs=idxs, data=data))
    return out


def generate_sdig_code(field: int, n_per_row: int, seed: int = 0, code: int = 3):
    """matgen::generate (lcpc-brakedown-pc/src/matgen.rs:28-52) on the host; no device needed."""
    ch = C.c_void_p()
    _check(_cabi.lib().lcpc_b200_sdig_code_generate(field, code, n_per_row, seed, C.byref(ch)))
    try:
        pre, post = host_code_matrices(ch, FIELD_LIMBS[field])
        return pre, post, _cabi.lib().lcpc_b200_sdig_code_codeword_length(ch)
    finally:
        _cabi.lib().lcpc_b200_sdig_code_free(ch)


# ------------------------------------------------------------------ commit
class LcRoot:
    """``LcRoot<D, E>`` (lcpc-2d/src/lib.rs:315-323): the Merkle root digest."""

    def __init__(self, root: bytes):
        self.root = bytes(root)

    def into_raw(self) -> bytes:
        """LcRoot::into_raw (:337-339)."""
        return self.root

    def __eq__(self, other):
        return isinstance(other, LcRoot) and self.root == other.root

    def __repr__(self):
        return f"LcRoot({self.root.hex()})"


class LcCommit:
    """``LcCommit<D, E>`` (lcpc-2d/src/lib.rs:172-184), device-resident.

    ``comm``, ``coeffs`` and ``hashes`` (the struct's fields, :178-183) are materialised on the host
    lazily, on first access; ``get_root``, ``collapse`` and ``open_columns`` never need them.
    """

    def __init__(self, enc: LcEncoding, handle):
        self.enc, self._h = enc, handle
        nr, npr, nc, nh = (C.c_size_t() for _ in range(4))
        _check(_cabi.lib().lcpc_b200_commit_dims(handle, C.byref(nr), C.byref(npr), C.byref(nc), C.byref(nh)))
        self.n_rows, self.n_per_row, self.n_cols, self.n_hashes = nr.value, npr.value, nc.value, nh.value
        self._comm = self._coeffs = self._hashes = None

    @classmethod
    def commit(cls, coeffs_in, enc: LcEncoding) -> "LcCommit":
        """LcCommit::commit (:299-301 -> :622-671)."""
        a = _elems(coeffs_in, enc.field)
        h = C.c_void_p()
        _check(_cabi.lib().lcpc_b200_commit_new(enc._h, _ptr(a), a.shape[0], C.byref(h)), enc.ctx)
        return cls(enc, h)

    @classmethod
    def commit_device(cls, d_ptr: int, length: int, enc: LcEncoding) -> "LcCommit":
        """Same with the coefficients already in device memory (raw pointer)."""
        h = C.c_void_p()
        _check(_cabi.lib().lcpc_b200_commit_new_dev(enc._h, C.c_void_p(d_ptr), length, C.byref(h)), enc.ctx)
        return cls(enc, h)

    @classmethod
    def from_fields(cls, enc: LcEncoding, comm, coeffs, hashes, n_rows: int) -> "LcCommit":
        """``Deserialize for LcCommit`` (:256-268): a device-resident commit from host-side fields (e.g. the output of
        ``deserialize_commit_fields``); sizes are checked like ``check_comm`` (:672-688), nothing is recomputed."""
        a, k = _elems(comm, enc.field), _elems(coeffs, enc.field)
        hs = np.ascontiguousarray(hashes, dtype=np.uint8).reshape(-1, 32)
        h = C.c_void_p()
        _check(_cabi.lib().lcpc_b200_commit_from_host(enc._h, _ptr(a), a.shape[0], _ptr(k), k.shape[0], _ptr(hs),
                                                      hs.shape[0], int(n_rows), C.byref(h)), enc.ctx)
        return cls(enc, h)

    def rerun_device(self, d_ptr: int, length: int):
        """Enqueue the commit again into this object's buffers (no allocation, no sync)."""
        _check(_cabi.lib().lcpc_b200_commit_rerun_dev(self._h, C.c_void_p(d_ptr), length), self.enc.ctx)
        self._comm = self._coeffs = self._hashes = None

    def rerun(self, coeffs_in):
        a = _elems(coeffs_in, self.enc.field)
        _check(_cabi.lib().lcpc_b200_commit_rerun(self._h, _ptr(a), a.shape[0]), self.enc.ctx)
        self._comm = self._coeffs = self._hashes = None

    def close(self):
        if getattr(self, "_h", None):
            _cabi.lib().lcpc_b200_commit_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def get_n_per_row(self) -> int:
        """:284-286."""
        return self.n_per_row

    def get_n_cols(self) -> int:
        """:289-291."""
        return self.n_cols

    def get_n_rows(self) -> int:
        """:294-296."""
        return self.n_rows

    def get_root(self) -> LcRoot:
        """LcCommit::get_root (:276-281)."""
        out = np.empty(32, np.uint8)
        _check(_cabi.lib().lcpc_b200_commit_root(self._h, _ptr(out)), self.enc.ctx)
        return LcRoot(out.tobytes())

    def _download(self, comm=False, coeffs=False, hashes=False):
        L = self.enc.L
        c = np.empty((self.n_rows * self.n_cols, L), np.uint64) if comm else None
        k = np.empty((self.n_rows * self.n_per_row, L), np.uint64) if coeffs else None
        h = np.empty((self.n_hashes, 32), np.uint8) if hashes else None
        _check(_cabi.lib().lcpc_b200_commit_download(self._h, _ptr(c), _ptr(k), _ptr(h)), self.enc.ctx)
        return c, k, h

    def rerun_to_host(self, coeffs_in, comm=None, coeffs=None, hashes=None):
        """commit() into this object with the fields copied out as they become final (download overlapped with
        the upload and the encode); arrays should be page-locked for the overlap to be real."""
        a = _elems(coeffs_in, self.enc.field)
        self._comm = self._coeffs = self._hashes = None
        _check(_cabi.lib().lcpc_b200_commit_rerun_to_host(self._h, _ptr(a), a.shape[0], _ptr(comm), _ptr(coeffs),
                                                          _ptr(hashes)), self.enc.ctx)

    def download_into(self, comm=None, coeffs=None, hashes=None):
        """Copy the LcCommit fields (lcpc-2d/src/lib.rs:178-183) into caller-owned arrays (any may be None); with
        page-locked arrays this is the eager host-visible commit() of INTEGRATION.md at full PCIe rate."""
        _check(_cabi.lib().lcpc_b200_commit_download(self._h, _ptr(comm), _ptr(coeffs), _ptr(hashes)), self.enc.ctx)

    @property
    def comm(self) -> np.ndarray:
        if self._comm is None:
            self._comm = self._download(comm=True)[0]
        return self._comm

    @property
    def coeffs(self) -> np.ndarray:
        if self._coeffs is None:
            self._coeffs = self._download(coeffs=True)[1]
        return self._coeffs

    @property
    def hashes(self) -> np.ndarray:
        if self._hashes is None:
            self._hashes = self._download(hashes=True)[2]
        return self._hashes

    def phase_times(self):
        """Device ms of (pad/copy, encode, leaf hash, merkle) of the last run + kernel launches per phase."""
        ms = (C.c_float * 4)()
        nl = (C.c_int * 3)()
        _check(_cabi.lib().lcpc_b200_commit_phase_times(self._h, ms, nl), self.enc.ctx)
        return list(ms), list(nl)

    def device_ptrs(self):
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _check(_cabi.lib().lcpc_b200_commit_device_ptrs(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def collapse(self, tensor) -> np.ndarray:
        """collapse_columns (:1095-1123) against this commit's coefficient matrix."""
        t = _elems(tensor, self.enc.field)
        if t.shape[0] != self.n_rows:  # ProverError::OuterTensor (:1016-1018)
            raise LcpcError(_cabi.ERR_BAD_ARG, "outer tensor length != n_rows")
        poly = np.empty((self.n_per_row, self.enc.L), np.uint64)
        _check(_cabi.lib().lcpc_b200_commit_collapse(self._h, _ptr(t), _ptr(poly)), self.enc.ctx)
        return poly

    def degree_test(self, key: bytes, with_tensor: bool = False):
        """One degree test of prove() (:1026-1041): the 32-byte transcript challenge `key` is expanded into the
        random tensor on the device (ChaCha20Rng::from_seed + F::random) and collapsed against the coefficients."""
        if len(key) != 32:
            raise LcpcError(_cabi.ERR_BAD_ARG, "degree-test key must be 32 bytes")
        kb = np.frombuffer(bytes(key), dtype=np.uint8).copy()
        poly = np.empty((self.n_per_row, self.enc.L), np.uint64)
        tensor = np.empty((self.n_rows, self.enc.L), np.uint64) if with_tensor else None
        _check(_cabi.lib().lcpc_b200_commit_degree_test(self._h, _ptr(kb), _ptr(poly), _ptr(tensor)), self.enc.ctx)
        return (poly, tensor) if with_tensor else poly

    def prove(self, outer_tensor, enc: "LcEncoding", tr):
        """LcCommit::prove (:304-311): the evaluation proof for ``outer_tensor`` under transcript ``tr``."""
        from .proof import prove
        return prove(self, outer_tensor, enc, tr)

    def open_columns(self, cols):
        """open_column (:788-825) for every index in ``cols``: (values (n, n_rows, L), paths (n, path_len, 32))."""
        idx = np.ascontiguousarray(cols, dtype=np.uint64)
        n = idx.shape[0]
        path_len = (self.n_cols - 1).bit_length()
        vals = np.empty((n, self.n_rows, self.enc.L), np.uint64)
        paths = np.empty((n, path_len, 32), np.uint8)
        _check(_cabi.lib().lcpc_b200_commit_open_columns(self._h, _ptr(idx), n, _ptr(vals), _ptr(paths)), self.enc.ctx)
        return vals, paths


# ------------------------------------------------------------------ standalone pieces
def merkleize(field: int, comm, n_rows: int, n_cols: int, ctx: Context | None = None) -> np.ndarray:
    """merkleize (lcpc-2d/src/lib.rs:690-704) of a host matrix -> hashes (2*np2-1, 32)."""
    ctx = ctx or default_context()
    a = _elems(comm, field)
    assert a.shape[0] == n_rows * n_cols
    np2 = 1 << (n_cols - 1).bit_length()
    out = np.empty((2 * np2 - 1, 32), np.uint8)
    _check(_cabi.lib().lcpc_b200_merkleize(ctx._h, field, _ptr(a), n_rows, n_cols, _ptr(out)), ctx)
    return out


def collapse_columns(field: int, coeffs, tensor, n_rows: int, n_per_row: int, ctx: Context | None = None) -> np.ndarray:
    """collapse_columns (lcpc-2d/src/lib.rs:1095-1123) on host arrays."""
    ctx = ctx or default_context()
    a, t = _elems(coeffs, field), _elems(tensor, field)
    assert a.shape[0] == n_rows * n_per_row and t.shape[0] == n_rows
    out = np.empty((n_per_row, FIELD_LIMBS[field]), np.uint64)
    _check(_cabi.lib().lcpc_b200_collapse(ctx._h, field, _ptr(a), _ptr(t), _ptr(out), n_rows, n_per_row), ctx)
    return out


def expand_tensor(field: int, key: bytes, n: int, ctx: Context | None = None) -> np.ndarray:
    """n x F::random from ChaCha20Rng::from_seed(key) on the device (lcpc-2d/src/lib.rs:1028-1032, 868-877)."""
    ctx = ctx or default_context()
    if len(key) != 32:
        raise LcpcError(_cabi.ERR_BAD_ARG, "key must be 32 bytes")
    kb = np.frombuffer(bytes(key), dtype=np.uint8).copy()
    out = np.empty((n, FIELD_LIMBS[field]), np.uint64)
    _check(_cabi.lib().lcpc_b200_expand_tensor(ctx._h, field, _ptr(kb), n, _ptr(out)), ctx)
    return out


def field_one(field: int) -> np.ndarray:
    """Field::one() in its stored (Montgomery) form: R mod p as L limbs.  Host-only."""
    out = np.empty(FIELD_LIMBS[field], np.uint64)
    _check(_cabi.lib().lcpc_b200_field_one(field, _ptr(out)))
    return out


_OPS = {"add": 0, "sub": 1, "mul": 2, "from_mont": 4, "mul_sos": 5, "lazy_sum37": 6, "mul_karatsuba": 7}


def field_op(field: int, op: str, a, b=None, ctx: Context | None = None) -> np.ndarray:
    """Element-wise device field arithmetic (parity tests of lcpc_b200/csrc/field.cuh)."""
    ctx = ctx or default_context()
    a = _elems(a, field)
    bb = _elems(b, field) if b is not None else None
    out = np.empty_like(a)
    _check(_cabi.lib().lcpc_b200_field_op(ctx._h, field, _OPS[op], _ptr(out), _ptr(a), _ptr(bb), a.shape[0]), ctx)
    return out
