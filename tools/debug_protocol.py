"""Step-by-step prove/verify on one small case with a flushed print per stage (GPU debugging aid)."""
import faulthandler
import sys

faulthandler.enable()
sys.path.insert(0, ".")
import numpy as np  # noqa: E402

import lcpc_b200 as P  # noqa: E402
import oracle as O  # noqa: E402


def say(*a):
    print(*a, flush=True)


for kind, field, length in [("ligero", P.FT255, 1 << 10), ("sdig", P.FT127, 1 << 12)]:
    say("case", kind, field, length)
    enc = P.LigeroEncoding(field, length) if kind == "ligero" else P.SdigEncoding(field, length, seed=0)
    x = O.random_elems(field, length, seed=1)
    c = P.LcCommit.commit(x, enc)
    say("commit ok", c.n_rows, c.n_per_row, c.n_cols, enc.get_n_degree_tests(), enc.get_n_col_opens())
    outer = O.random_elems(field, c.n_rows, seed=2)
    inner = O.random_elems(field, c.n_per_row, seed=3)
    tr = P.Transcript(b"t")
    say("transcript ok")
    proof = c.prove(outer, enc, tr)
    say("prove ok", proof.p_eval.shape, proof.p_random_vec.shape, proof.cols.shape, proof.paths.shape)
    ev = proof.verify(c.get_root(), outer, inner, enc, P.Transcript(b"t"))
    say("verify ok", ev)
say("ALLOK")
