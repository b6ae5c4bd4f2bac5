"""One rank of a ShardedCommit (one process per GPU; tests/test_shard.py launches it under torchrun).  gloo is used
for the rendezvous and for nothing else: the 64-byte IPC handles of the windows travel through it once."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle as O  # noqa: E402
import lcpc_b200 as P  # noqa: E402
from oracle import protocol as PR  # noqa: E402
from oracle.transcript import Transcript as OTranscript  # noqa: E402


def main():
    kind, field, length = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = int(os.environ.get("LOCAL_RANK", "0")) % torch.cuda.device_count()
    ctx = P.Context(dev)
    if kind == "ligero":
        enc, oenc = P.LigeroEncoding(field, length, ctx=ctx), O.Encoding.ligero(field, length)
    else:
        enc, oenc = P.SdigEncoding(field, length, seed=0, ctx=ctx), O.Encoding.sdig(field, length, seed=0)
    x = O.random_elems(field, length, seed=3)
    oc = oenc.commit(x)
    sc = P.ShardedCommit(enc, length)
    mine = x[sc.row_lo * sc.n_per_row:sc.row_lo * sc.n_per_row + sc.n_elems]
    sc.commit(mine)      # from host rows (copy overlapped with the encode)
    sc.load_rows(mine)
    sc.commit()          # from device-resident rows
    root = sc.get_root()
    comm = oc["comm"].reshape(sc.n_rows, sc.n_cols, -1)
    cols_ok = bool((sc.local_columns() == comm[:, sc.col_lo:sc.col_hi]).all())
    outer, inner = O.random_elems(field, sc.n_rows, seed=4), O.random_elems(field, sc.n_per_row, seed=5)
    proof = sc.prove(outer, P.Transcript(b"ipc"))
    oproof = PR.prove(field, oc, outer, oenc.get_n_degree_tests(), oenc.get_n_col_opens(), OTranscript(b"ipc"))
    prove_ok = P.serialize_proof(proof) == PR.wire_proof(oproof)
    verify_ok = True
    if rank == 0:
        ev = proof.verify(root, outer, inner, enc, P.Transcript(b"ipc"))
        verify_ok = bool((ev == O.dot(field, inner, oproof["p_eval"])).all())
    dist.barrier()
    print(json.dumps({"rank": rank, "root_ok": root.root == oc["root"], "cols_ok": cols_ok, "prove_ok": prove_ok,
                      "verify_ok": verify_ok}), flush=True)
    dist.barrier()
    sc.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
