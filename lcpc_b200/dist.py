"""Multi-GPU commit: row-block encode -> one all-to-all -> column-block hashing (one process per GPU).

The reference is a single process (rayon threads); its two data-parallel loops are row-parallel
(`enc.encode` per row, lcpc-2d/src/lib.rs:648-653) and column-parallel (`hash_columns`, :706-745).
Between them sits a transpose, which across GPUs is exactly one all-to-all:

  rank g   : encodes rows [row_lo_g, row_hi_g)            -> comm_rows[rows_g][n_cols]      (device)
           : packs per-destination tiles                   -> send[h] = comm_rows[:, cols_h] (device)
  all ranks: torch.distributed.all_to_all_single over NCCL/NVLink (the only bulk collective)
  rank h   : holds comm[:, cols_h] as [n_rows][width_h]    -> leaf digests of its columns
           : reduces its aligned Merkle subtrees on device -> subtree roots (32 B each)
  all ranks: all_gather of the subtree roots (<= 2 KiB; commitment assembly, not bulk data), then the
             top log2(S) layers on every rank -> every rank holds the same LcRoot.

Column blocks are unions of ALIGNED Merkle subtrees (np2 / S leaves each, S = 8 * world), dealt
contiguously over the subtrees that contain real columns, so that non-power-of-two `n_cols`
(Brakedown) still balances: the all-padding subtrees have a constant root that is computed once.

`torch.distributed` is plumbing only: the process group, the NCCL call, and (for tests on CPU) the
gloo backend for the planning logic.  All field / hash work goes through the C ABI.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _cabi
from .host import FIELD_LIMBS, LcRoot, _check


@dataclass
class Plan:
    """Static partition of one commit over `world` ranks."""
    world: int
    n_rows: int
    n_per_row: int
    n_cols: int
    np2: int
    sub_leaves: int          # leaves per aligned subtree (T)
    n_sub: int               # S = np2 / T
    n_real_sub: int          # subtrees containing at least one real column
    row_lo: list             # world + 1 row boundaries
    sub_lo: list             # world + 1 subtree boundaries (over the real subtrees)
    col_lo: list             # world + 1 column boundaries

    def rows(self, r):
        return self.row_lo[r], self.row_lo[r + 1]

    def cols(self, r):
        return self.col_lo[r], self.col_lo[r + 1]

    def subs(self, r):
        return self.sub_lo[r], self.sub_lo[r + 1]


def make_plan(n_rows: int, n_per_row: int, n_cols: int, world: int, sub_per_rank: int = 8) -> Plan:
    np2 = 1 << (n_cols - 1).bit_length() if n_cols > 1 else 1
    n_sub = min(np2, 1 << ((world * sub_per_rank) - 1).bit_length())
    T = np2 // n_sub
    n_real = (n_cols + T - 1) // T
    row_lo = [(g * n_rows) // world for g in range(world + 1)]
    sub_lo = [(g * n_real) // world for g in range(world + 1)]
    col_lo = [min(s * T, n_cols) for s in sub_lo]
    col_lo[-1] = n_cols
    return Plan(world, n_rows, n_per_row, n_cols, np2, T, n_sub, n_real, row_lo, sub_lo, col_lo)


def _sz(v):
    return C.c_size_t(int(v))


class CudaOps:
    """Compute steps of the distributed commit on this rank's GPU, through the C ABI."""

    def __init__(self, enc):
        import torch
        self.torch, self.enc, self.ctx = torch, enc, enc.ctx
        self.field, self.L = enc.field, enc.L
        self.device = torch.device("cuda", self.ctx.device)
        self.stream = torch.cuda.ExternalStream(self.ctx.stream, device=self.device)

    def on_stream(self):
        return self.torch.cuda.stream(self.stream)

    def synchronize(self):
        self.ctx.synchronize()

    def encode_rows(self, coeffs, comm_rows, n_rows, n_per_row):
        _check(_cabi.lib().lcpc_b200_encode_rows_dev(self.enc._h, C.c_void_p(coeffs.data_ptr()), _sz(n_per_row),
                                                     _sz(n_per_row), C.c_void_p(comm_rows.data_ptr()), _sz(n_rows)), self.ctx)

    def encode_rows_h2d(self, host, n_elems, coeffs, comm_rows, n_rows):
        """Pinned host rows -> device coefficient rows -> encoded rows, PCIe copy overlapped with the encode."""
        _check(_cabi.lib().lcpc_b200_encode_rows_h2d(self.enc._h, C.c_void_p(host.data_ptr()), _sz(n_elems),
                                                     C.c_void_p(coeffs.data_ptr()), C.c_void_p(comm_rows.data_ptr()),
                                                     _sz(n_rows)), self.ctx)

    def encode_rows_scatter(self, src, n_rows, n_per_row, tmp, starts, ptrs, row0, coeffs=None, n_elems=None):
        """Row-block encode whose last pass stores per column block: block h -> the matrix at device
        address ptrs[h] (local or peer-mapped).  `src` is a device tensor, or -- with `coeffs` (device
        staging) and `n_elems` -- a pinned host tensor whose PCIe copy is overlapped with the encode."""
        starts = np.ascontiguousarray(starts, dtype=np.uint64)
        ptrs = np.ascontiguousarray(ptrs, dtype=np.uint64)
        sc = _cabi.Scatter(len(ptrs), starts.ctypes.data, ptrs.ctypes.data, int(row0))
        lib = _cabi.lib()
        if coeffs is None:
            _check(lib.lcpc_b200_encode_rows_scatter_dev(self.enc._h, C.c_void_p(src.data_ptr()), _sz(n_per_row),
                                                         _sz(n_per_row), C.c_void_p(tmp.data_ptr()), _sz(n_rows),
                                                         C.byref(sc)), self.ctx)
        else:
            _check(lib.lcpc_b200_encode_rows_scatter_h2d(self.enc._h, C.c_void_p(src.data_ptr()), _sz(n_elems),
                                                         C.c_void_p(coeffs.data_ptr()), C.c_void_p(tmp.data_ptr()),
                                                         _sz(n_rows), C.byref(sc)), self.ctx)

    def pack(self, comm_rows, n_rows, n_cols, n_blocks, starts, send):
        _check(_cabi.lib().lcpc_b200_pack_column_blocks_dev(self.ctx._h, self.field, C.c_void_p(comm_rows.data_ptr()),
                                                            _sz(n_rows), _sz(n_cols), _sz(n_blocks),
                                                            C.c_void_p(starts.data_ptr()), C.c_void_p(send.data_ptr())), self.ctx)

    def hash_columns(self, cols, n_rows, n_cols, leaves):
        _check(_cabi.lib().lcpc_b200_hash_columns_dev(self.ctx._h, self.field, C.c_void_p(cols.data_ptr()), _sz(n_rows),
                                                      _sz(n_cols), _sz(n_cols), C.c_void_p(leaves.data_ptr())), self.ctx)

    def merkle_layers(self, nodes, n_leaves, n_layers):
        _check(_cabi.lib().lcpc_b200_merkle_layers_dev(self.ctx._h, C.c_void_p(nodes.data_ptr()), _sz(n_leaves), n_layers), self.ctx)

    # ---- prove side ----
    def collapse_rows(self, coeffs, row_stride, tensor, poly, n_rows, n_per_row):
        """poly[c] = sum_r tensor[r] * coeffs[r][c] over device tensors (collapse_columns, lcpc-2d/src/lib.rs:1095-1123)."""
        _check(_cabi.lib().lcpc_b200_collapse_dev(self.ctx._h, self.field, C.c_void_p(coeffs.data_ptr()), _sz(row_stride),
                                                  C.c_void_p(tensor.data_ptr()), C.c_void_p(poly.data_ptr()), _sz(n_rows),
                                                  _sz(n_per_row)), self.ctx)

    def expand_tensor(self, key: bytes, n: int) -> np.ndarray:
        from .host import expand_tensor
        return expand_tensor(self.field, key, n, ctx=self.ctx)

    def to_repr(self, elems: np.ndarray) -> np.ndarray:
        """Canonical little-endian bytes of Montgomery-form elements (to_repr): (n, L) u64 -> (n, 8L) u8."""
        from .host import field_op
        return field_op(self.field, "from_mont", elems, ctx=self.ctx).view(np.uint8).reshape(elems.shape[0], -1)

    def one(self) -> np.ndarray:
        from .host import field_one
        return field_one(self.field)


class DistributedCommit:
    """Device-resident, column-sharded LcCommit over a torch.distributed process group.

    `ops` supplies the per-rank compute steps (default: CudaOps, the C ABI).  The orchestration --
    partition, split sizes, the all-to-all, root assembly -- is backend-independent, which is how the
    world_size-2 gloo tests exercise it on CPU with a checker backend.
    """

    def __init__(self, enc, n_coeffs: int, group=None, ops=None, transport=None):
        import os

        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.enc, self.group = enc, group
        self.ops = ops if ops is not None else CudaOps(enc)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.field, self.L = enc.field, enc.L
        n_rows, n_per_row, n_cols = enc.get_dims(n_coeffs)
        self.n_coeffs = n_coeffs
        self.plan = p = make_plan(n_rows, n_per_row, n_cols, self.world)
        dev = self.ops.device
        L = self.L
        r0, r1 = p.rows(self.rank)
        c0, c1 = p.cols(self.rank)
        s0, s1 = p.subs(self.rank)
        self.my_rows, self.my_cols, self.my_subs = r1 - r0, c1 - c0, s1 - s0
        i64 = torch.int64
        self.d_coeffs = torch.zeros(max(1, self.my_rows * n_per_row * L), dtype=i64, device=dev)
        self.d_comm_rows = torch.empty(max(1, self.my_rows * n_cols * L), dtype=i64, device=dev)
        self.d_send = torch.empty(max(1, self.my_rows * n_cols * L), dtype=i64, device=dev)
        # how encoded row blocks reach the column owners:
        #   "p2p"    the encode's last pass stores straight into every owner's receive matrix through
        #            peer-mapped (symmetric) memory over NVLink -- compute and exchange are one kernel
        #   "nccl"   the last pass stores per-destination tiles into a local send buffer (fused pack), then one
        #            all_to_all_single
        #   "staged" row-major encode + pack kernel + all_to_all_single (backends without the scatter store)
        want = transport or os.environ.get("LCPC_B200_TRANSPORT", "auto")
        self.max_cols = max(p.col_lo[h + 1] - p.col_lo[h] for h in range(self.world))
        self.symm = None
        self.transport = "staged"
        if hasattr(self.ops, "encode_rows_scatter"):
            self.transport = "nccl"
            if want in ("auto", "p2p") and self.world > 1:
                try:
                    import torch.distributed._symmetric_memory as symm_mem
                    g = group if group is not None else dist.group.WORLD
                    # two receive matrices used alternately: a rank can only be two commits ahead of an owner
                    # after passing the "stores have landed" barrier of the commit in between, which that owner
                    # reaches after hashing -- so no second barrier is needed before overwriting
                    half = max(1, n_rows * self.max_cols * L)
                    buf = symm_mem.empty(2 * half, dtype=i64, device=dev)
                    self.symm = symm_mem.rendezvous(buf, g.group_name)
                    self.recv_bufs = [buf[:half], buf[half:]]
                    self.d_recv = self.recv_bufs[0]
                    self.peer_ptrs = [[int(x) + b * half * 8 for x in self.symm.buffer_ptrs] for b in (0, 1)]
                    self.parity = 1
                    self.transport = "p2p"
                except Exception as e:  # no peer access / allocator unavailable: NCCL moves the tiles instead
                    if want == "p2p":
                        raise
                    self.symm_error = repr(e)
            # every rank must take the same path (the collectives differ)
            flag = torch.tensor([1 if self.transport == "p2p" else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag.item()) == 0:
                self.transport, self.symm = "nccl", None
        if self.transport != "p2p":
            self.d_recv = torch.empty(max(1, n_rows * self.my_cols * L), dtype=i64, device=dev)
        self.in_splits = [self.my_rows * (p.col_lo[h + 1] - p.col_lo[h]) * L for h in range(self.world)]
        self.out_splits = [(p.row_lo[g + 1] - p.row_lo[g]) * self.my_cols * L for g in range(self.world)]
        self.d_starts = torch.tensor(p.col_lo, dtype=i64, device=dev)
        # local forest: my_subs aligned subtrees side by side, [leaves | layer 1 | ... | roots]
        T = p.sub_leaves
        self.sub_layers = T.bit_length() - 1
        n_leaves = self.my_subs * T
        self.forest_nodes = sum(n_leaves >> l for l in range(self.sub_layers + 1))
        self.d_forest = torch.zeros(max(1, self.forest_nodes) * 32, dtype=torch.uint8, device=dev)
        self.roots_off = sum(n_leaves >> l for l in range(self.sub_layers)) * 32
        self.max_subs = max(p.sub_lo[g + 1] - p.sub_lo[g] for g in range(self.world))
        self.d_my_roots = torch.zeros(max(1, self.max_subs) * 32, dtype=torch.uint8, device=dev)
        self.d_all_roots = torch.zeros(self.world * max(1, self.max_subs) * 32, dtype=torch.uint8, device=dev)
        self.d_top = torch.zeros((2 * p.n_sub - 1) * 32, dtype=torch.uint8, device=dev)
        self.uniform_subs = self.my_subs > 0 and all(
            p.sub_lo[g + 1] - p.sub_lo[g] == self.my_subs for g in range(self.world))
        self.zero_root = self._zero_subtree_root()
        if p.n_real_sub < p.n_sub:
            pad = torch.from_numpy(np.tile(np.frombuffer(self.zero_root, np.uint8), p.n_sub - p.n_real_sub).copy()).to(dev)
            with self.ops.on_stream():
                self.d_top[p.n_real_sub * 32:p.n_sub * 32] = pad
        self.ops.synchronize()
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)  # buffers above were zero-filled on torch's stream, not the engine stream

    # root of an all-padding subtree: T zero leaves hashed up (lcpc-2d/src/lib.rs:665,696 leave them zero)
    def _zero_subtree_root(self) -> bytes:
        torch = self.torch
        dev = self.ops.device
        buf = torch.zeros(3 * 32, dtype=torch.uint8, device=dev)
        with self.ops.on_stream():
            for _ in range(self.sub_layers):
                self.ops.merkle_layers(buf, 2, 1)
                buf[:32] = buf[64:96]
                buf[32:64] = buf[64:96]
            out = buf[:32].cpu()
        self.ops.synchronize()
        return out.numpy().tobytes()

    def load_rows_from_host(self, host_rows):
        """H2D of this rank's row block (a pinned int64 torch tensor or numpy array of limbs)."""
        torch = self.torch
        t = host_rows if isinstance(host_rows, torch.Tensor) else torch.from_numpy(
            np.ascontiguousarray(host_rows).view(np.int64).reshape(-1))
        with self.ops.on_stream():
            self.d_coeffs[:t.numel()].copy_(t.reshape(-1), non_blocking=True)
            if t.numel() < self.d_coeffs.numel():
                self.d_coeffs[t.numel():].zero_()

    def run(self, host_rows=None):
        """Enqueue one distributed commit on the engine stream (no host synchronisation).

        `host_rows` (a pinned int64 tensor holding this rank's coefficient rows) makes the row-block step
        start from host memory, its PCIe copy overlapped with the encode; without it the rows loaded by
        `load_rows_from_host` are encoded."""
        dist, p, ops = self.dist, self.plan, self.ops
        ev = getattr(self, "phase_events", None)  # optional (start, encode done, exchange done, end) CUDA events
        with ops.on_stream():
            if ev:
                ev[0].record(ops.stream)
            if self.transport == "staged":
                if self.my_rows:
                    if host_rows is not None:
                        self.load_rows_from_host(host_rows)
                    ops.encode_rows(self.d_coeffs, self.d_comm_rows, self.my_rows, p.n_per_row)
                    ops.pack(self.d_comm_rows, self.my_rows, p.n_cols, self.world, self.d_starts, self.d_send)
            else:
                if self.transport == "p2p":
                    self.parity ^= 1
                    self.d_recv = self.recv_bufs[self.parity]
                    ptrs, row0 = self.peer_ptrs[self.parity], p.row_lo[self.rank]
                else:  # tile h of the send buffer = [my_rows][width_h] at element offset my_rows * col_lo[h]
                    base = self.d_send.data_ptr()
                    ptrs = [base + self.my_rows * p.col_lo[h] * self.L * 8 for h in range(self.world)]
                    row0 = 0
                if self.my_rows:
                    if host_rows is not None:
                        ops.encode_rows_scatter(host_rows, self.my_rows, p.n_per_row, self.d_comm_rows, p.col_lo, ptrs,
                                                row0, coeffs=self.d_coeffs, n_elems=host_rows.numel() // self.L)
                    else:
                        ops.encode_rows_scatter(self.d_coeffs, self.my_rows, p.n_per_row, self.d_comm_rows, p.col_lo,
                                                ptrs, row0)
            if ev:
                ev[1].record(ops.stream)
            if self.transport == "p2p":
                self.symm.barrier(channel=1)  # every rank's stores into this rank's matrix have landed
            else:
                n_send, n_recv = sum(self.in_splits), sum(self.out_splits)
                dist.all_to_all_single(self.d_recv[:n_recv], self.d_send[:n_send], self.out_splits, self.in_splits,
                                       group=self.group)
            if ev:
                ev[2].record(ops.stream)
            if self.my_cols:
                ops.hash_columns(self.d_recv, p.n_rows, self.my_cols, self.d_forest)
                if self.sub_layers:
                    ops.merkle_layers(self.d_forest, self.my_subs * p.sub_leaves, self.sub_layers)
            if self.uniform_subs:
                # every rank owns the same number of subtrees: gather their roots straight into the top tree
                dist.all_gather_into_tensor(self.d_top[:p.n_real_sub * 32],
                                            self.d_forest[self.roots_off:self.roots_off + self.my_subs * 32], group=self.group)
            else:
                if self.my_cols:
                    self.d_my_roots[:self.my_subs * 32] = self.d_forest[self.roots_off:self.roots_off + self.my_subs * 32]
                dist.all_gather_into_tensor(self.d_all_roots, self.d_my_roots, group=self.group)
                stride = max(1, self.max_subs) * 32
                for g in range(self.world):
                    s0, s1 = p.subs(g)
                    if s1 > s0:
                        self.d_top[s0 * 32:s1 * 32] = self.d_all_roots[g * stride:g * stride + (s1 - s0) * 32]
            if p.n_sub > 1:
                ops.merkle_layers(self.d_top, p.n_sub, p.n_sub.bit_length() - 1)
            if ev:
                ev[3].record(ops.stream)

    def get_root(self) -> LcRoot:
        with self.ops.on_stream():
            root = self.d_top[-32:].cpu()
        self.ops.synchronize()
        return LcRoot(root.numpy().tobytes())

    # ---- prove() over the sharded commit (row-sharded coeffs, column-sharded comm) ----
    def collapse(self, tensor) -> np.ndarray:
        """collapse_columns (lcpc-2d/src/lib.rs:1095-1123) with the coefficient rows sharded by row block: every rank
        combines its own rows with its slice of `tensor`, the partial vectors are all-gathered (world x n_per_row
        elements) and summed on the device.  Every rank returns the same (n_per_row, L) array."""
        torch, dist, p, ops, L = self.torch, self.dist, self.plan, self.ops, self.L
        dev = ops.device
        t = np.ascontiguousarray(tensor, dtype=np.uint64).reshape(-1, L)
        if t.shape[0] != p.n_rows:
            raise _cabi.LcpcError(_cabi.ERR_OUTER_TENSOR, "tensor length != n_rows")
        if not hasattr(self, "d_part"):
            i64 = torch.int64
            self.d_part = torch.zeros(p.n_per_row * L, dtype=i64, device=dev)
            self.d_all_parts = torch.zeros(self.world * p.n_per_row * L, dtype=i64, device=dev)
            self.d_poly = torch.zeros(p.n_per_row * L, dtype=i64, device=dev)
            self.d_tensor = torch.zeros(p.n_rows * L, dtype=i64, device=dev)
            ones = np.tile(ops.one().reshape(1, L), (self.world, 1))
            self.d_ones = torch.from_numpy(ones.view(np.int64).reshape(-1).copy()).to(dev)
            if dev.type == "cuda":
                torch.cuda.synchronize(dev)  # filled on torch's stream, read on the engine stream
        r0, r1 = p.rows(self.rank)
        with ops.on_stream():
            self.d_tensor.copy_(torch.from_numpy(t.view(np.int64).reshape(-1)))
            if self.my_rows:
                ops.collapse_rows(self.d_coeffs, p.n_per_row, self.d_tensor[r0 * L:r1 * L], self.d_part, self.my_rows,
                                  p.n_per_row)
            else:
                self.d_part.zero_()  # the zero element is all-zero limbs
            dist.all_gather_into_tensor(self.d_all_parts, self.d_part, group=self.group)
            ops.collapse_rows(self.d_all_parts, p.n_per_row, self.d_ones, self.d_poly, self.world, p.n_per_row)
            out = self.d_poly.cpu()
        ops.synchronize()
        return out.numpy().view(np.uint64).reshape(p.n_per_row, L).copy()

    def open_columns(self, cols):
        """open_column (lcpc-2d/src/lib.rs:788-825) for every index in `cols` over the column-sharded commit: the
        owner of a column gathers its values and the siblings inside its own subtrees, the siblings above come from
        the top tree; one all-reduce over disjoint supports hands every rank all openings.  Pure data movement."""
        torch, dist, p, L = self.torch, self.dist, self.plan, self.L
        dev = self.ops.device
        cols = np.ascontiguousarray(cols, dtype=np.uint64).astype(np.int64)
        if cols.size and (cols.min() < 0 or cols.max() >= p.n_cols):
            raise _cabi.LcpcError(_cabi.ERR_COLUMN, "bad column number")  # ProverError::ColumnNumber (:797-799)
        n, path_len = cols.shape[0], (p.np2 - 1).bit_length()
        c0, c1 = p.cols(self.rank)
        with self.ops.on_stream():
            vals = torch.zeros((n, p.n_rows, L), dtype=torch.int64, device=dev)
            paths = torch.zeros((n, path_len, 32), dtype=torch.uint8, device=dev)
            mine = np.nonzero((cols >= c0) & (cols < c1))[0]
            if mine.size:
                where = torch.from_numpy(mine).to(dev)
                local = torch.from_numpy(cols[mine] - c0).to(dev)
                recv = self.d_recv[:p.n_rows * self.my_cols * L].view(p.n_rows, self.my_cols, L)
                vals[where] = recv[:, local, :].permute(1, 0, 2)
                forest, off, ln = self.d_forest.view(-1, 32), 0, self.my_subs * p.sub_leaves
                for l in range(self.sub_layers):  # siblings inside this rank's aligned subtrees
                    paths[where, l] = forest[off + ((local >> l) ^ 1)]
                    off, ln = off + ln, ln >> 1
                top, off, ln = self.d_top.view(-1, 32), 0, p.n_sub
                sub = torch.from_numpy(cols[mine] // p.sub_leaves).to(dev)
                for t in range(p.n_sub.bit_length() - 1):  # siblings in the tree over the subtree roots
                    paths[where, self.sub_layers + t] = top[off + ((sub >> t) ^ 1)]
                    off, ln = off + ln, ln >> 1
            dist.all_reduce(vals, group=self.group)
            dist.all_reduce(paths, group=self.group)
            vals_h, paths_h = vals.cpu(), paths.cpu()
        self.ops.synchronize()
        return vals_h.numpy().view(np.uint64), paths_h.numpy()

    def prove(self, outer_tensor, tr):
        """LcCommit::prove (lcpc-2d/src/lib.rs:1004-1093) on the sharded commit; every rank drives an identical
        transcript `tr` (lcpc_b200.Transcript) and returns the same LcEvalProof."""
        from .proof import LcEvalProof, sample_columns
        enc, p, ops = self.enc, self.plan, self.ops
        p_random = []
        for _ in range(enc.get_n_degree_tests()):
            key = tr.challenge_bytes(enc.LABEL_DT, 32)
            pr = self.collapse(ops.expand_tensor(key, p.n_rows))
            tr.append_reprs(enc.LABEL_PR, ops.to_repr(pr))
            p_random.append(pr)
        p_eval = self.collapse(outer_tensor)
        tr.append_reprs(enc.LABEL_PE, ops.to_repr(p_eval))
        key = tr.challenge_bytes(enc.LABEL_CO, 32)
        cols = sample_columns(key, p.n_cols, enc.get_n_col_opens())
        vals, paths = self.open_columns(cols)
        p_rand = np.stack(p_random) if p_random else np.empty((0, p.n_per_row, self.L), np.uint64)
        return LcEvalProof(self.field, p.n_cols, p_eval, p_rand, vals, paths, col_idx=cols)

    # ---- inspection helpers (tests) ----
    def local_columns(self) -> np.ndarray:
        """This rank's column block of comm as (n_rows, my_cols, L)."""
        self.ops.synchronize()
        n = self.plan.n_rows * self.my_cols * self.L
        return self.d_recv[:n].cpu().numpy().view(np.uint64).reshape(self.plan.n_rows, self.my_cols, self.L)

    def local_leaves(self) -> np.ndarray:
        self.ops.synchronize()
        return self.d_forest[:self.my_cols * 32].cpu().numpy().reshape(self.my_cols, 32)


def bench_distributed(args, ctx, enc, field, n, synthetic_coeffs):
    """The N>1 arm of bench.py: one commit of `n` coefficients sharded over all ranks."""
    import time

    import torch
    import torch.distributed as dist

    import bench as B  # ClockSampler

    world, rank = dist.get_world_size(), dist.get_rank()
    L = FIELD_LIMBS[field]
    dc = DistributedCommit(enc, n)
    p = dc.plan
    r0, r1 = p.rows(rank)
    # this rank's rows of a synthetic polynomial (uniform field elements, seeded per rank: only the slice a rank
    # owns is ever materialised, so 2^28 coefficients do not cost every rank 8 GiB of host memory)
    lo, hi = r0 * p.n_per_row, min(r1 * p.n_per_row, n)
    x = synthetic_coeffs(field, max(hi - lo, 0), seed=1000 + rank)
    host = torch.from_numpy(np.ascontiguousarray(x).view(np.int64).reshape(-1)).pin_memory()
    dc.load_rows_from_host(host)
    ctx.synchronize()
    for _ in range(args.warmup):
        dc.run()
    ctx.synchronize()
    root0 = dc.get_root()
    sampler = B.ClockSampler(torch.cuda.current_device())
    launches0 = ctx.launch_count
    dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(dc.ops.stream)
    for _ in range(args.steps):
        dc.run()
    ev1.record(dc.ops.stream)
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    launches = torch.tensor([ctx.launch_count - launches0], device="cuda")
    dist.all_reduce(launches, op=dist.ReduceOp.SUM)
    assert dc.get_root() == root0
    # per-phase device times of this rank (separate loop; events on the engine stream)
    dc.phase_events = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ph = np.zeros(3)
    for _ in range(args.steps):
        dc.run()
        torch.cuda.synchronize()
        ph += np.array([dc.phase_events[i].elapsed_time(dc.phase_events[i + 1]) for i in range(3)])
    dc.phase_events = None
    ph /= args.steps
    pht = torch.tensor(ph, device="cuda")
    dist.all_reduce(pht, op=dist.ReduceOp.MAX)
    ph = pht.cpu().numpy()
    # end to end: pinned host rows -> H2D -> commit -> D2H root, wall clock bracketed by barriers
    for _ in range(2):
        dc.run(host)
        dc.get_root()
    dist.barrier()
    torch.cuda.synchronize()
    e2e_steps = max(3, args.steps // 2)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        dc.run(host)
        r = dc.get_root()
    torch.cuda.synchronize()
    dist.barrier()
    e2e = torch.tensor([(time.perf_counter() - t0) / e2e_steps], device="cuda")
    dist.all_reduce(e2e, op=dist.ReduceOp.MAX)
    clocks = sampler.stop()
    assert r == root0
    ms_per_step = float(ms.item()) / args.steps
    e2e_s = float(e2e.item())
    # parity at the benchmarked size: the sharded commit's LcRoot == the single-GPU commit of the same polynomial ==
    # the oracle's commit of it (bench.py's checker leg; the product never imports the oracle).  Rank 0 rebuilds
    # every rank's slice; beyond 2^24 the full-size comparison is skipped (host memory / minutes of CPU time).
    root_check = "skipped (size)"
    if n <= (1 << 24):
        if rank == 0:
            from .host import LcCommit
            parts = []
            for g in range(world):
                g0, g1 = p.rows(g)
                parts.append(synthetic_coeffs(field, max(min(g1 * p.n_per_row, n) - g0 * p.n_per_row, 0), seed=1000 + g))
            full = np.concatenate(parts)
            single = LcCommit.commit(full, enc)
            same_single = single.get_root() == root0
            single.close()
            root_check = "equals the single-GPU commit" if same_single else "MISMATCH"
            oracle_root = getattr(B, "oracle_root", None)
            if same_single and oracle_root is not None:
                ok = oracle_root(enc, field, full) == root0.root
                root_check = "equals the oracle's LcRoot and the single-GPU commit (same coefficients)" if ok else "MISMATCH"
            assert root_check != "MISMATCH", "distributed LcRoot differs from the single-GPU commit / the oracle"
        dist.barrier()
    # prove() over the sharded commit (config 4 of BASELINE.json is commit + prove): wall clock, max over ranks;
    # rank 0 then verifies the proof on its own GPU against the sharded commit's root
    from .proof import Transcript
    outer = synthetic_coeffs(field, p.n_rows, seed=7)
    dc.prove(outer, Transcript(b"bench"))
    dist.barrier()
    t0 = time.perf_counter()
    proof = dc.prove(outer, Transcript(b"bench"))
    t_prove = torch.tensor([time.perf_counter() - t0], device="cuda")
    dist.all_reduce(t_prove, op=dist.ReduceOp.MAX)
    prove = {"prove_ms": float(t_prove.item()) * 1e3, "n_degree_tests": int(proof.p_random_vec.shape[0]),
             "n_col_opens": int(proof.cols.shape[0]),
             "note": "LcCommit::prove over row-sharded coefficients and column-sharded comm: per-rank partial row "
                     "combinations all-gathered and summed on the device, openings gathered from the column owners"}
    if rank == 0:
        inner = synthetic_coeffs(field, p.n_per_row, seed=8)
        t0 = time.perf_counter()
        proof.verify(root0, outer, inner, enc, Transcript(b"bench"))
        prove["verify_ms_rank0"] = (time.perf_counter() - t0) * 1e3
    dist.barrier()
    return dict(value=n / (ms_per_step * 1e-3), root_check=root_check, prove=prove, ms_per_step=ms_per_step, gpu_launches=int(launches.item()), clocks=clocks,
                root=root0.root.hex(), transport=dc.transport,
                phases_ms={"encode_and_scatter": float(ph[0]), "exchange_wait": float(ph[1]), "hash_merkle_root": float(ph[2])},
                dominant=_dominant_multi(enc, field, p, dc, float(ph[0])),
                e2e={"value": n / e2e_s, "unit": "field-elts/s", "h2d_bytes_per_step": int(n * 8 * L),
                     "d2h_bytes_per_step": 32 * world, "ms_per_step": e2e_s * 1e3,
                     "mode": "row blocks from pinned host memory on every rank; every rank reads back the LcRoot"})


def _dominant_multi(enc, field, p, dc, encode_ms):
    """Per-GPU roofline entry of the encode step at N > 1 (max over ranks of the event-timed phase): Ligero's two
    transform passes over this rank's rows; algorithmic bytes as in the single-GPU case, per rank."""
    if enc.__class__.__name__ != "LigeroEncoding" or dc.my_rows == 0:
        return None
    B = 8 * FIELD_LIMBS[field]
    log_n = p.n_cols.bit_length() - 1
    n_pass = 1 if log_n <= 10 else -(-log_n // 10)
    moved = B * dc.my_rows * (p.n_per_row + p.n_cols) + 2 * B * dc.my_rows * p.n_cols * (n_pass - 1)
    return dict(kernel="ntt_pass_kernel (per GPU; last pass stores into the column owners' memory)",
                launches_per_step=n_pass, phase_ms=encode_ms, moved_bytes_per_launch=moved / n_pass, traffic=None)
