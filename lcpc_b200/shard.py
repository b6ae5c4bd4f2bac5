"""Thin callers of the sharded (multi-GPU) commit of the C ABI (include/lcpc_b200.h, lcpc_b200_shard_* / _multi_*).

Everything on the data path -- row-block encode with peer stores into the column owners' memory, the "tiles have
landed" flags, column hashing, the exchange of subtree roots, the top tree, the prover's partial row combinations and
column openings -- happens inside the library (csrc/shard.cu).  What is left here is what a host has to do anyway:

* ``MultiCommit``: ONE process drives several GPUs, like the reference's single-process ``commit()``
  (lcpc-2d/src/lib.rs:622-671): one encoding per device, one call.
* ``ShardedCommit``: one process per GPU (how bench.py is launched under torchrun); the only thing exchanged
  through ``torch.distributed`` is the 64-byte CUDA IPC handle of each rank's window, once, at construction --
  plumbing, not data.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _cabi
from .host import FIELD_LIMBS, LcRoot, _check, _elems, _ptr
from .proof import LcEvalProof, _labels


def shard_plan(n_rows: int, n_per_row: int, n_cols: int, world: int) -> dict:
    """The partition the library uses (host arithmetic only, no device needed)."""
    arr = lambda: (C.c_size_t * (world + 1))()  # noqa: E731
    row_lo, col_lo, sub_lo = arr(), arr(), arr()
    T, S = C.c_size_t(), C.c_size_t()
    _check(_cabi.lib().lcpc_b200_shard_plan(n_rows, n_per_row, n_cols, world, row_lo, col_lo, sub_lo, C.byref(T), C.byref(S)))
    return dict(row_lo=list(row_lo), col_lo=list(col_lo), sub_lo=list(sub_lo), sub_leaves=T.value, n_sub=S.value)


class _ProveMixin:
    """prove()-side plumbing shared by both shapes; subclasses supply _call_prove."""

    def _proof_buffers(self, enc):
        L = FIELD_LIMBS[enc.field]
        ndt, nco = enc.get_n_degree_tests(), enc.get_n_col_opens()
        path_len = (self.n_cols - 1).bit_length()
        return (ndt, nco, np.empty((self.n_per_row, L), np.uint64), np.empty((ndt, self.n_per_row, L), np.uint64),
                np.empty(nco, np.uint64), np.empty((nco, self.n_rows, L), np.uint64), np.empty((nco, path_len, 32), np.uint8))


class Shard(_ProveMixin):
    """One GPU's part of a sharded LcCommit (``lcpc_b200_shard``)."""

    def __init__(self, enc, n_coeffs: int, world: int, rank: int, max_open: int | None = None):
        self.enc, self.world, self.rank = enc, world, rank
        self._h = C.c_void_p()
        max_open = enc.get_n_col_opens() if max_open is None else max_open
        _check(_cabi.lib().lcpc_b200_shard_new(enc._h, n_coeffs, world, rank, max_open, C.byref(self._h)), enc.ctx)
        v = [C.c_size_t() for _ in range(8)]
        _check(_cabi.lib().lcpc_b200_shard_dims(self._h, *[C.byref(x) for x in v]))
        (self.n_rows, self.n_per_row, self.n_cols, self.row_lo, self.row_hi, self.col_lo, self.col_hi,
         self.n_elems) = [x.value for x in v]

    @classmethod
    def _view(cls, enc, handle, world, rank):
        """A non-owning view of a shard that belongs to a MultiCommit (inspection only)."""
        self = cls.__new__(cls)
        self.enc, self.world, self.rank, self._h, self._borrowed = enc, world, rank, C.c_void_p(handle), True
        v = [C.c_size_t() for _ in range(8)]
        _check(_cabi.lib().lcpc_b200_shard_dims(self._h, *[C.byref(x) for x in v]))
        (self.n_rows, self.n_per_row, self.n_cols, self.row_lo, self.row_hi, self.col_lo, self.col_hi,
         self.n_elems) = [x.value for x in v]
        return self

    def close(self):
        if getattr(self, "_h", None) and not getattr(self, "_borrowed", False):
            _cabi.lib().lcpc_b200_shard_free(self._h)
        self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def window(self):
        """(device pointer, bytes, 64-byte CUDA IPC handle) of this rank's exported window."""
        ptr, n = C.c_void_p(), C.c_size_t()
        handle = (C.c_uint8 * 64)()
        _check(_cabi.lib().lcpc_b200_shard_window(self._h, C.byref(ptr), C.byref(n), handle), self.enc.ctx)
        return ptr.value, n.value, bytes(handle)

    def connect(self, peer_ptrs=None, ipc_handles=None):
        pp = None
        if peer_ptrs is not None:
            pp = (C.c_void_p * self.world)(*[C.c_void_p(p) if p else None for p in peer_ptrs])
        hh = None
        if ipc_handles is not None:
            blob = b"".join(ipc_handles)
            assert len(blob) == 64 * self.world
            hh = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        _check(_cabi.lib().lcpc_b200_shard_connect(self._h, pp, hh), self.enc.ctx)

    def load_rows(self, rows):
        a = _elems(rows, self.enc.field) if self.n_elems else np.zeros((0, FIELD_LIMBS[self.enc.field]), np.uint64)
        _check(_cabi.lib().lcpc_b200_shard_load_rows(self._h, _ptr(a), a.shape[0]), self.enc.ctx)

    def commit_host_ptr(self, ptr: int, n_elems: int):
        """Enqueue a commit fed from (pinned) host memory at `ptr`."""
        _check(_cabi.lib().lcpc_b200_shard_commit(self._h, C.c_void_p(ptr), n_elems), self.enc.ctx)

    def commit(self, rows=None):
        """Enqueue one commit: from host rows, or (None) from the rows `load_rows` stored on the device."""
        if rows is None:
            _check(_cabi.lib().lcpc_b200_shard_commit_dev(self._h, None, 0), self.enc.ctx)
        else:
            a = _elems(rows, self.enc.field)
            _check(_cabi.lib().lcpc_b200_shard_commit(self._h, _ptr(a), a.shape[0]), self.enc.ctx)
            self.enc.ctx.synchronize()  # `a` may be a temporary

    def get_root(self) -> LcRoot:
        out = np.empty(32, np.uint8)
        _check(_cabi.lib().lcpc_b200_shard_root(self._h, _ptr(out)), self.enc.ctx)
        return LcRoot(out.tobytes())

    def root_enqueue(self, host_ptr: int):
        """Enqueue the D2H of the LcRoot into page-locked host memory at `host_ptr` (no synchronisation)."""
        _check(_cabi.lib().lcpc_b200_shard_root_enqueue(self._h, C.c_void_p(host_ptr)), self.enc.ctx)

    def join(self):
        """Make the context's stream wait for the shard's hash stream (pipelined mode); enqueue only."""
        _check(_cabi.lib().lcpc_b200_shard_join(self._h), self.enc.ctx)

    def phase_times(self):
        ms = (C.c_float * 3)()
        _check(_cabi.lib().lcpc_b200_shard_phase_times(self._h, ms), self.enc.ctx)
        return list(ms)

    def collapse(self, tensor=None, key: bytes | None = None, want_repr=False):
        L = FIELD_LIMBS[self.enc.field]
        t = _elems(tensor, self.enc.field) if tensor is not None else None
        if t is not None and t.shape[0] != self.n_rows:
            raise _cabi.LcpcError(_cabi.ERR_OUTER_TENSOR, "tensor length != n_rows")
        kb = (C.c_uint8 * 32).from_buffer_copy(key) if key is not None else None
        _check(_cabi.lib().lcpc_b200_shard_collapse_begin(self._h, _ptr(t), kb), self.enc.ctx)
        poly = np.empty((self.n_per_row, L), np.uint64)
        repr_ = np.empty((self.n_per_row, 8 * L), np.uint8) if want_repr else None
        _check(_cabi.lib().lcpc_b200_shard_collapse_finish(self._h, _ptr(poly), _ptr(repr_)), self.enc.ctx)
        return (poly, repr_) if want_repr else poly

    def open_columns(self, cols):
        cols = np.ascontiguousarray(cols, dtype=np.uint64)
        L = FIELD_LIMBS[self.enc.field]
        n, path_len = cols.shape[0], (self.n_cols - 1).bit_length()
        _check(_cabi.lib().lcpc_b200_shard_open_begin(self._h, _ptr(cols), n), self.enc.ctx)
        vals, paths = np.empty((n, self.n_rows, L), np.uint64), np.empty((n, path_len, 32), np.uint8)
        _check(_cabi.lib().lcpc_b200_shard_open_finish(self._h, _ptr(vals), _ptr(paths)), self.enc.ctx)
        return vals, paths

    def prove(self, outer_tensor, tr, enc=None) -> LcEvalProof:
        """LcCommit::prove (lcpc-2d/src/lib.rs:1004-1093); every rank passes an identical transcript."""
        enc = enc or self.enc
        outer = _elems(outer_tensor, enc.field)
        ndt, nco, p_eval, p_rand, idx, cols, paths = self._proof_buffers(enc)
        lb = _labels(enc)
        _check(_cabi.lib().lcpc_b200_shard_prove(self._h, tr._h, C.byref(lb), _ptr(outer), outer.shape[0], ndt, nco,
                                                 _ptr(p_eval), _ptr(p_rand), _ptr(idx), _ptr(cols), _ptr(paths)), enc.ctx)
        return LcEvalProof(enc.field, self.n_cols, p_eval, p_rand, cols, paths, col_idx=idx)

    # inspection helpers (tests)
    def local_columns(self) -> np.ndarray:
        """This rank's column block of comm as (n_rows, my_cols, L), copied to the host."""
        L, my_cols = FIELD_LIMBS[self.enc.field], self.col_hi - self.col_lo
        out = np.empty((self.n_rows, my_cols, L), np.uint64)
        _check(_cabi.lib().lcpc_b200_shard_download(self._h, _ptr(out), None), self.enc.ctx)
        return out

    def local_leaves(self) -> np.ndarray:
        out = np.empty((self.col_hi - self.col_lo, 32), np.uint8)
        _check(_cabi.lib().lcpc_b200_shard_download(self._h, None, _ptr(out)), self.enc.ctx)
        return out


class ShardedCommit(Shard):
    """One process per GPU: this rank's shard, connected to its peers' windows through CUDA IPC handles exchanged
    once over `torch.distributed` (any backend: the handles are 64 bytes of host data)."""

    def __init__(self, enc, n_coeffs: int, group=None, max_open: int | None = None):
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        super().__init__(enc, n_coeffs, world, rank, max_open)
        _, _, handle = self.window()
        handles = [None] * world
        dist.all_gather_object(handles, handle, group=group)
        self.connect(ipc_handles=handles)
        dist.barrier(group=group)  # every window is mapped before anyone stores into it
        self.transport = "p2p (peer stores from the encode kernel, C ABI)"


class MultiCommit(_ProveMixin):
    """ONE process, several GPUs (``lcpc_b200_multi``): `encs` = the same encoding built on one context per device."""

    def __init__(self, encs, coeffs_in, max_open: int | None = None):
        self.encs = list(encs)
        enc = self.enc = self.encs[0]
        self._a = _elems(coeffs_in, enc.field)
        n = self._a.shape[0]
        self.n_rows, self.n_per_row, self.n_cols = enc.get_dims(n)
        arr = (C.c_void_p * len(self.encs))(*[e._h for e in self.encs])
        self._h = C.c_void_p()
        max_open = enc.get_n_col_opens() if max_open is None else max_open
        _check(_cabi.lib().lcpc_b200_commit_new_multi(arr, len(self.encs), _ptr(self._a), n, max_open, C.byref(self._h)), enc.ctx)

    @classmethod
    def commit(cls, coeffs_in, encs):
        """LcCommit::commit (lcpc-2d/src/lib.rs:299-301) over len(encs) GPUs."""
        return cls(encs, coeffs_in)

    def close(self):
        if getattr(self, "_h", None):
            _cabi.lib().lcpc_b200_multi_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def shard(self, g: int) -> Shard:
        """Rank g's shard (a view; owned by this object)."""
        h = _cabi.lib().lcpc_b200_multi_shard(self._h, g)
        if not h:
            raise IndexError(g)
        return Shard._view(self.encs[g], h, len(self.encs), g)

    def rerun(self, coeffs_in=None):
        if coeffs_in is not None:
            self._a = _elems(coeffs_in, self.enc.field)
        _check(_cabi.lib().lcpc_b200_multi_rerun(self._h, _ptr(self._a), self._a.shape[0]), self.enc.ctx)

    def get_root(self) -> LcRoot:
        out = np.empty(32, np.uint8)
        _check(_cabi.lib().lcpc_b200_multi_root(self._h, _ptr(out)), self.enc.ctx)
        return LcRoot(out.tobytes())

    def collapse(self, tensor=None, key: bytes | None = None) -> np.ndarray:
        L = FIELD_LIMBS[self.enc.field]
        t = _elems(tensor, self.enc.field) if tensor is not None else None
        if t is not None and t.shape[0] != self.n_rows:
            raise _cabi.LcpcError(_cabi.ERR_OUTER_TENSOR, "tensor length != n_rows")
        kb = (C.c_uint8 * 32).from_buffer_copy(key) if key is not None else None
        poly = np.empty((self.n_per_row, L), np.uint64)
        _check(_cabi.lib().lcpc_b200_multi_collapse(self._h, _ptr(t), kb, _ptr(poly), None), self.enc.ctx)
        return poly

    def open_columns(self, cols):
        cols = np.ascontiguousarray(cols, dtype=np.uint64)
        L = FIELD_LIMBS[self.enc.field]
        n, path_len = cols.shape[0], (self.n_cols - 1).bit_length()
        vals, paths = np.empty((n, self.n_rows, L), np.uint64), np.empty((n, path_len, 32), np.uint8)
        _check(_cabi.lib().lcpc_b200_multi_open_columns(self._h, _ptr(cols), n, _ptr(vals), _ptr(paths)), self.enc.ctx)
        return vals, paths

    def prove(self, outer_tensor, enc, tr) -> LcEvalProof:
        outer = _elems(outer_tensor, enc.field)
        ndt, nco, p_eval, p_rand, idx, cols, paths = self._proof_buffers(enc)
        lb = _labels(enc)
        _check(_cabi.lib().lcpc_b200_multi_prove(self._h, tr._h, C.byref(lb), _ptr(outer), outer.shape[0], ndt, nco,
                                                 _ptr(p_eval), _ptr(p_rand), _ptr(idx), _ptr(cols), _ptr(paths)), enc.ctx)
        return LcEvalProof(enc.field, self.n_cols, p_eval, p_rand, cols, paths, col_idx=idx)
