"""Worker for the multi-process tests: one rank of a DistributedCommit.

backend "gloo": CPU tensors, the per-rank compute steps are done by the ORACLE (tests only) so that
the orchestration of lcpc_b200/dist.py -- partition, split sizes, all-to-all, root assembly -- is
exercised without a GPU.  backend "nccl": the product path (CudaOps) on one GPU per rank.
Prints one JSON line per rank.
"""
import contextlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle as O  # noqa: E402
from lcpc_b200 import dist as D  # noqa: E402


class OracleEnc:
    """Duck-typed encoding for the CPU run: dims from the oracle's encoding."""

    def __init__(self, oenc):
        self.o, self.field, self.L = oenc, oenc.field, oenc.L

    LABEL_DT, LABEL_PR, LABEL_PE, LABEL_CO = b"$l//DT", b"$l//PR", b"$l//PE", b"$l//CO"

    def get_dims(self, n):
        return self.o.get_dims(n)

    def get_n_degree_tests(self):
        return self.o.get_n_degree_tests()

    def get_n_col_opens(self):
        return self.o.get_n_col_opens()


class OracleOps:
    """Checker backend: same step interface as lcpc_b200.dist.CudaOps, computed by the oracle on CPU."""

    def __init__(self, enc):
        self.enc, self.field, self.L = enc, enc.field, enc.L
        self.device = torch.device("cpu")

    def on_stream(self):
        return contextlib.nullcontext()

    def synchronize(self):
        pass

    def _u64(self, t):
        return t.numpy().view(np.uint64)

    def encode_rows(self, coeffs, comm_rows, n_rows, n_per_row):
        n_cols = self.enc.o.n_cols
        src = self._u64(coeffs).reshape(-1, self.L)[:n_rows * n_per_row].reshape(n_rows, n_per_row, self.L)
        dst = self._u64(comm_rows)[:n_rows * n_cols * self.L].reshape(n_rows, n_cols, self.L)
        for r in range(n_rows):
            row = np.zeros((n_cols, self.L), np.uint64)
            row[:n_per_row] = src[r]
            dst[r] = self.enc.o.encode(row)

    def pack(self, comm_rows, n_rows, n_cols, n_blocks, starts, send):
        src = self._u64(comm_rows)[:n_rows * n_cols * self.L].reshape(n_rows, n_cols, self.L)
        out = self._u64(send)
        st = starts.tolist()
        for h in range(n_blocks):
            tile = src[:, st[h]:st[h + 1]].reshape(-1)
            out[n_rows * st[h] * self.L:n_rows * st[h] * self.L + tile.size] = tile

    def hash_columns(self, cols, n_rows, n_cols, leaves):
        comm = self._u64(cols)[:n_rows * n_cols * self.L].reshape(n_rows * n_cols, self.L)
        h = O.merkleize(self.field, comm, n_rows, n_cols)
        leaves.numpy()[:n_cols * 32] = h[:n_cols].reshape(-1)

    def merkle_layers(self, nodes, n_leaves, n_layers):
        buf = nodes.numpy()
        off, ln = 0, n_leaves
        for _ in range(n_layers):
            for i in range(ln // 2):
                buf[(off + ln + i) * 32:(off + ln + i + 1) * 32] = np.frombuffer(
                    O.blake3(buf[(off + 2 * i) * 32:(off + 2 * i + 2) * 32].tobytes()), np.uint8)
            off, ln = off + ln, ln // 2


    # ---- prove side ----
    def collapse_rows(self, coeffs, row_stride, tensor, poly, n_rows, n_per_row):
        c = self._u64(coeffs)[:n_rows * row_stride * self.L].reshape(n_rows, row_stride, self.L)[:, :n_per_row]
        t = self._u64(tensor)[:n_rows * self.L].reshape(n_rows, self.L)
        out = O.collapse(self.field, np.ascontiguousarray(c).reshape(-1, self.L), t, n_rows, n_per_row)
        self._u64(poly)[:n_per_row * self.L] = out.reshape(-1)

    def expand_tensor(self, key, n):
        return O.random_elems_from_key(self.field, key, n)

    def to_repr(self, elems):
        return O.to_repr(self.field, elems)

    def one(self):
        return O.to_mont(self.field, [1])[0]


def main():
    backend, kind, field, n, seed = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    transport = sys.argv[6] if len(sys.argv) > 6 else None
    rank = int(os.environ["RANK"])
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if backend == "nccl":
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        dist.init_process_group("gloo")
    oenc = O.Encoding.ligero(field, n) if kind == "ligero" else O.Encoding.sdig(field, n, seed=seed)
    x = O.random_elems(field, n, seed=seed + 100)
    if backend == "nccl":
        import lcpc_b200 as P
        ctx = P.Context(local_rank)
        enc = P.LigeroEncoding(field, n, ctx=ctx) if kind == "ligero" else P.SdigEncoding(field, n, seed=seed, ctx=ctx)
        dc = D.DistributedCommit(enc, n, transport=transport)
    else:
        enc = OracleEnc(oenc)
        dc = D.DistributedCommit(enc, n, ops=OracleOps(enc))
    p = dc.plan
    r0, r1 = p.rows(rank)
    dc.load_rows_from_host(x[r0 * p.n_per_row:min(r1 * p.n_per_row, n)])
    dc.run()
    root = dc.get_root().root
    oc = oenc.commit(x)
    c0, c1 = p.cols(rank)
    comm = oc["comm"].reshape(p.n_rows, p.n_cols, -1)
    ok_cols = bool((dc.local_columns() == comm[:, c0:c1]).all())
    ok_leaves = bool((dc.local_leaves() == oc["hashes"][c0:c1]).all())
    if backend == "nccl":  # second run straight from pinned host rows (PCIe copy overlapped with the encode)
        rows = x[r0 * p.n_per_row:min(r1 * p.n_per_row, n)]
        host = torch.from_numpy(np.ascontiguousarray(rows).view(np.int64).reshape(-1)).pin_memory()
        dc.run(host if host.numel() else None)
    else:
        dc.run()  # a second run into the same buffers must reproduce the root
    again = dc.get_root().root
    # prove() over the sharded commit == the oracle's single-process prove, on every rank
    import lcpc_b200 as P
    from oracle import protocol as PR
    from oracle.transcript import Transcript as OTranscript
    outer = O.random_elems(field, p.n_rows, seed=seed + 200)
    proof = dc.prove(outer, P.Transcript(b"dist"))
    oproof = PR.prove(field, oc, outer, oenc.get_n_degree_tests(), oenc.get_n_col_opens(), OTranscript(b"dist"))
    ok_prove = bool((proof.p_eval == oproof["p_eval"]).all()) and len(oproof["p_random_vec"]) == proof.p_random_vec.shape[0]
    ok_prove = ok_prove and all(bool((a == b).all()) for a, b in zip(proof.p_random_vec, oproof["p_random_vec"]))
    ok_prove = ok_prove and [int(v) for v in proof.col_idx] == oproof["cols_to_open"]
    ok_prove = ok_prove and all(bool((proof.cols[j] == col).all()) and bool((proof.paths[j] == path).all())
                                for j, (col, path) in enumerate(oproof["columns"]))
    ok_verify = True
    if backend == "nccl":  # the verifier is one party: every rank checks the sharded proof on its own GPU
        inner = O.random_elems(field, p.n_per_row, seed=seed + 300)
        ev = proof.verify(root, outer, inner, enc, P.Transcript(b"dist"))
        ok_verify = bool((ev == O.dot(field, inner, oproof["p_eval"])).all())
    print(json.dumps(dict(rank=rank, root=root.hex(), want=oc["root"].hex(), ok_cols=ok_cols, ok_leaves=ok_leaves,
                          again=again.hex(), rows=[r0, r1], cols=[c0, c1], transport=dc.transport,
                          ok_prove=ok_prove, ok_verify=ok_verify)), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
