"""GPU parity tests of whole prove() / verify() through the C ABI (lcpc_b200_commit_prove, lcpc_b200_verify) against
the oracle's restatement of lcpc-2d/src/lib.rs:1004-1093 and :832-952.  Bit-exact: every proof field, the opened
column numbers, the evaluation, and the state the transcript is left in.
"""
import numpy as np
import pytest

import oracle as O
import lcpc_b200 as P
from lcpc_b200 import _cabi
from oracle import protocol as PR
from oracle import transcript as T

pytestmark = pytest.mark.gpu

LABEL = b"test transcript"


def _encodings(kind, field, length):
    if kind == "ligero":
        return P.LigeroEncoding(field, length), O.Encoding.ligero(field, length)
    return P.SdigEncoding(field, length, seed=0), O.Encoding.sdig(field, length, seed=0)


def _case(kind, field, length, seed):
    enc, oenc = _encodings(kind, field, length)
    x = O.random_elems(field, length, seed=seed)
    c, oc = P.LcCommit.commit(x, enc), oenc.commit(x)
    assert c.get_root().root == oc["root"]
    outer = O.random_elems(field, c.n_rows, seed=seed + 1)
    inner = O.random_elems(field, c.n_per_row, seed=seed + 2)
    return enc, oenc, c, oc, outer, inner


CASES = [("ligero", P.FT255, 1 << 10), ("ligero", P.FT255, 1 << 14), ("ligero", P.FT63, 3000),
         ("ligero", P.FT191, 5000), ("ligero", P.FT127, (1 << 13) - 5), ("sdig", P.FT127, 1 << 12),
         ("sdig", P.FT255, 3000), ("sdig", P.FT63, 1 << 13)]


@pytest.mark.parametrize("kind,field,length", CASES)
def test_prove_matches_oracle_and_verifies(kind, field, length):
    enc, oenc, c, oc, outer, inner = _case(kind, field, length, seed=31)
    assert enc.get_n_degree_tests() == oenc.get_n_degree_tests() and enc.get_n_col_opens() == oenc.get_n_col_opens()
    tr, otr = P.Transcript(LABEL), T.Transcript(LABEL)
    proof = c.prove(outer, enc, tr)
    oproof = PR.prove(field, oc, outer, oenc.get_n_degree_tests(), oenc.get_n_col_opens(), otr)
    assert proof.n_cols == oproof["n_cols"]
    assert (proof.p_eval == oproof["p_eval"]).all()
    assert proof.p_random_vec.shape[0] == len(oproof["p_random_vec"])
    for a, b in zip(proof.p_random_vec, oproof["p_random_vec"]):
        assert (a == b).all()
    assert [int(v) for v in proof.col_idx] == oproof["cols_to_open"]
    for j, (col, path) in enumerate(oproof["columns"]):
        assert (proof.cols[j] == col).all(), j
        assert (proof.paths[j] == path).all(), j
    # both transcripts absorbed the same bytes
    assert tr.challenge_bytes(b"after", 32) == otr.challenge_bytes(b"after", 32)
    # wire image == the oracle's element-by-element writer
    assert P.serialize_proof(proof) == PR.wire_proof(oproof)

    # verify on the device: accepts, returns the evaluation, leaves the transcript where the oracle's verifier does
    vtr, votr = P.Transcript(LABEL), T.Transcript(LABEL)
    ev = proof.verify(c.get_root(), outer, inner, enc, vtr)
    oev = PR.verify(field, oenc, oc["root"], outer, inner, oproof, votr)
    assert (ev == oev).all()
    want = O.dot(field, inner, O.collapse(field, oc["coeffs"], outer, c.n_rows, c.n_per_row))
    assert (ev == want).all()
    assert vtr.challenge_bytes(b"after", 32) == votr.challenge_bytes(b"after", 32)
    # the proof survives the wire
    back = P.deserialize_proof(P.serialize_proof(proof), field)
    assert (back.verify(oc["root"], outer, inner, enc, P.Transcript(LABEL)) == ev).all()


def test_two_proofs_on_one_transcript():
    """lcpc-2d/src/tests.rs `end_to_end_two_proofs`: two evaluations proved on the same transcript, verified in order."""
    field, length = P.FT255, 1 << 12
    enc, oenc, c, oc, outer, inner = _case("ligero", field, length, seed=5)
    outer2 = O.random_elems(field, c.n_rows, seed=77)
    inner2 = O.random_elems(field, c.n_per_row, seed=78)
    tr = P.Transcript(LABEL)
    tr.append_message(b"polycommit", c.get_root().root)
    tr.append_message(b"rate", b"0.25")
    pf1 = c.prove(outer, enc, tr)
    pf2 = c.prove(outer2, enc, tr)
    assert not (pf1.col_idx == pf2.col_idx).all()
    vtr = P.Transcript(LABEL)
    vtr.append_message(b"polycommit", c.get_root().root)
    vtr.append_message(b"rate", b"0.25")
    e1 = pf1.verify(c.get_root(), outer, inner, enc, vtr)
    e2 = pf2.verify(c.get_root(), outer2, inner2, enc, vtr)
    assert (e1 == O.dot(field, inner, O.collapse(field, oc["coeffs"], outer, c.n_rows, c.n_per_row))).all()
    assert (e2 == O.dot(field, inner2, O.collapse(field, oc["coeffs"], outer2, c.n_rows, c.n_per_row))).all()
    # verifying in the wrong order desynchronises the challenges
    vtr = P.Transcript(LABEL)
    vtr.append_message(b"polycommit", c.get_root().root)
    vtr.append_message(b"rate", b"0.25")
    with pytest.raises(P.LcpcError):
        pf2.verify(c.get_root(), outer2, inner2, enc, vtr)


def _expect(code, fn):
    with pytest.raises(P.LcpcError) as e:
        fn()
    assert e.value.code == code, (e.value.code, str(e.value))


@pytest.mark.parametrize("kind,field,length", [("ligero", P.FT127, 1 << 11), ("sdig", P.FT127, 1 << 12)])
def test_verify_error_variants(kind, field, length):
    """Every VerifierError the reference can return (:141-170), same precedence as the match at :937-942."""
    enc, oenc, c, oc, outer, inner = _case(kind, field, length, seed=9)
    proof = c.prove(outer, enc, P.Transcript(LABEL))
    root = c.get_root()
    one = O.to_mont(field, [1])

    def run(p=proof, r=root, o=outer, i=inner, e=enc, label=LABEL):
        return p.verify(r, o, i, e, P.Transcript(label))

    run()
    fewer = P.LcEvalProof(field, proof.n_cols, proof.p_eval, proof.p_random_vec, proof.cols[:-1], proof.paths[:-1])
    _expect(_cabi.VERR_NUM_COL_OPENS, lambda: run(p=fewer))
    _expect(_cabi.VERR_INNER_TENSOR, lambda: run(i=inner[:-1]))
    _expect(_cabi.VERR_OUTER_TENSOR, lambda: run(o=np.concatenate([outer, outer[:1]])))
    wrong_dims = P.LcEvalProof(field, proof.n_cols * 2, proof.p_eval, proof.p_random_vec, proof.cols, proof.paths)
    _expect(_cabi.VERR_ENCODING_DIMS, lambda: run(p=wrong_dims))
    # a corrupted column value breaks the degree test first (and the path): ColumnDegree has precedence
    cols = proof.cols.copy()
    cols[5, 0] = O.field_op(field, "add", cols[5, :1], one)[0]
    _expect(_cabi.VERR_COLUMN_DEGREE, lambda: run(p=P.LcEvalProof(field, proof.n_cols, proof.p_eval, proof.p_random_vec, cols, proof.paths)))
    # a corrupted p_eval changes the PE absorb -> other columns get opened: the dot products still match the proof's
    # own columns' degree test only by luck, so any of the three column errors is legitimate
    pe = proof.p_eval.copy()
    pe[3] = O.field_op(field, "add", pe[3:4], one)[0]
    with pytest.raises(P.LcpcError) as e:
        run(p=P.LcEvalProof(field, proof.n_cols, pe, proof.p_random_vec, proof.cols, proof.paths))
    assert e.value.code in (_cabi.VERR_COLUMN_DEGREE, _cabi.VERR_COLUMN_EVAL, _cabi.VERR_COLUMN_PATH)
    # wrong outer tensor with everything else intact: degree tests pass, the evaluation check fails
    other = O.random_elems(field, c.n_rows, seed=1234)
    _expect(_cabi.VERR_COLUMN_EVAL, lambda: run(o=other))
    # a flipped path byte, a wrong root: ColumnPath
    paths = proof.paths.copy()
    paths[7, 2, 3] ^= 1
    _expect(_cabi.VERR_COLUMN_PATH, lambda: run(p=P.LcEvalProof(field, proof.n_cols, proof.p_eval, proof.p_random_vec, proof.cols, paths)))
    _expect(_cabi.VERR_COLUMN_PATH, lambda: run(r=bytes(32)))
    # the same tampered proofs are rejected by the oracle's verifier with the same variant
    oproof = dict(n_cols=proof.n_cols, p_eval=proof.p_eval, p_random_vec=list(proof.p_random_vec),
                  columns=[(cols[j], proof.paths[j]) for j in range(cols.shape[0])])
    with pytest.raises(PR.VerifierError) as oe:
        PR.verify(field, oenc, oc["root"], outer, inner, oproof, T.Transcript(LABEL))
    assert oe.value.kind == "ColumnDegree"
    # a different transcript label changes all challenges
    with pytest.raises(P.LcpcError):
        run(label=b"another transcript")


@pytest.mark.parametrize("kind,field,length", [("ligero", P.FT255, 1 << 10), ("sdig", P.FT127, 1 << 12)])
def test_verify_rejects_malformed_proofs(kind, field, length):
    """The verifier's trust boundary: a non-canonical element (v + p has the same transcript bytes as v), a Merkle
    path of the wrong length and a proof over another field are refused before anything reaches the device."""
    enc, oenc, c, oc, outer, inner = _case(kind, field, length, seed=11)
    proof = c.prove(outer, enc, P.Transcript(LABEL))
    root = c.get_root()
    proof.verify(root, outer, inner, enc, P.Transcript(LABEL))
    L = enc.L
    p_limbs = O.int_to_limbs(O.field_info(field)["modulus"], L)

    def plus_p(elem):  # the same residue, not canonical: v + p < 2^(64L) for every v < p
        return O.int_to_limbs(O.limbs_to_int(elem) + O.limbs_to_int(p_limbs), L)

    for which in ("p_eval", "p_random_vec", "cols"):
        arrs = dict(p_eval=proof.p_eval.copy(), p_random_vec=proof.p_random_vec.copy(), cols=proof.cols.copy())
        a = arrs[which].reshape(-1, L)
        a[a.shape[0] // 2] = plus_p(a[a.shape[0] // 2])
        bad = P.LcEvalProof(field, proof.n_cols, arrs["p_eval"], arrs["p_random_vec"], arrs["cols"], proof.paths)
        with pytest.raises(P.LcpcError) as e:
            bad.verify(root, outer, inner, enc, P.Transcript(LABEL))
        assert e.value.code == _cabi.ERR_BAD_ARG and "non-canonical" in str(e.value), (which, str(e.value))
    # p itself (the residue 0 in non-canonical form)
    pe = proof.p_eval.copy()
    pe[0] = p_limbs
    _expect(_cabi.ERR_BAD_ARG, lambda: P.LcEvalProof(field, proof.n_cols, pe, proof.p_random_vec, proof.cols, proof.paths)
            .verify(root, outer, inner, enc, P.Transcript(LABEL)))
    # path one sibling short / one long
    short = P.LcEvalProof(field, proof.n_cols, proof.p_eval, proof.p_random_vec, proof.cols, proof.paths[:, :-1].copy())
    _expect(_cabi.VERR_COLUMN_PATH, lambda: short.verify(root, outer, inner, enc, P.Transcript(LABEL)))
    longer = np.concatenate([proof.paths, proof.paths[:, :1]], axis=1)
    _expect(_cabi.VERR_COLUMN_PATH, lambda: P.LcEvalProof(field, proof.n_cols, proof.p_eval, proof.p_random_vec, proof.cols, longer)
            .verify(root, outer, inner, enc, P.Transcript(LABEL)))
    # a proof deserialized for a smaller field: refused by shape, never handed to C
    other = P.FT63 if field != P.FT63 else P.FT127
    wire = P.serialize_proof(proof)
    try:
        small = P.deserialize_proof(wire, other)
    except Exception:
        small = None
    if small is not None:
        _expect(_cabi.ERR_BAD_ARG, lambda: small.verify(root, outer, inner, enc, P.Transcript(LABEL)))
    wrong = P.LcEvalProof(other, proof.n_cols, proof.p_eval, proof.p_random_vec, proof.cols, proof.paths)
    _expect(_cabi.ERR_BAD_ARG, lambda: wrong.verify(root, outer, inner, enc, P.Transcript(LABEL)))


def test_prove_rejects_wrong_outer_tensor_length():
    field, length = P.FT63, 2000
    enc, oenc, c, oc, outer, inner = _case("ligero", field, length, seed=2)
    _expect(_cabi.ERR_OUTER_TENSOR, lambda: c.prove(outer[:-1], enc, P.Transcript(LABEL)))


def test_wire_commit_from_device_resident_commit():
    field, length = P.FT127, 3000
    enc, oenc, c, oc, outer, inner = _case("ligero", field, length, seed=4)
    blob = P.serialize_commit(c)
    assert blob == PR.wire_commit(oc)
    f = P.deserialize_commit_fields(blob, field)
    again = P.LcCommit.commit(f["coeffs"][:length], enc)  # a device-resident commit is rebuilt from its coefficients
    assert again.get_root().root == oc["root"] and (again.hashes == f["hashes"]).all()
    assert P.serialize_root(c.get_root()) == PR.wire_root(oc["root"])
    # Deserialize for LcCommit: the wire image becomes a device-resident commit that proves like the original
    back = P.deserialize_commit(blob, enc)
    assert back.get_root().root == oc["root"] and (back.comm == oc["comm"]).all() and (back.coeffs == oc["coeffs"]).all()
    p1 = c.prove(outer, enc, P.Transcript(LABEL))
    p2 = back.prove(outer, enc, P.Transcript(LABEL))
    assert P.serialize_proof(p1) == P.serialize_proof(p2)
    assert (p2.verify(oc["root"], outer, inner, enc, P.Transcript(LABEL)) ==
            O.dot(field, inner, O.collapse(field, oc["coeffs"], outer, c.n_rows, c.n_per_row))).all()
    with pytest.raises(P.LcpcError):  # check_comm: inconsistent sizes -> ProverError::Commit
        P.LcCommit.from_fields(enc, oc["comm"][:-1], oc["coeffs"], oc["hashes"], c.n_rows)
    with pytest.raises(P.LcpcError):
        P.LcCommit.from_fields(enc, oc["comm"], oc["coeffs"], oc["hashes"][:-1], c.n_rows)


def test_handles_may_be_freed_in_any_order():
    """An encoding outlives its Python wrapper while commits made with it are alive (reference counts in the C
    ABI): finalisers of a garbage-collected cycle run in arbitrary order."""
    ctx = P.Context(0)
    field, length = P.FT63, 1 << 10
    enc = P.LigeroEncoding(field, length, ctx=ctx)
    oenc = O.Encoding.ligero(field, length)
    x = O.random_elems(field, length, seed=8)
    c = P.LcCommit.commit(x, enc)
    want = oenc.commit(x)
    enc.close()   # encoding handle released first ...
    ctx.close()   # ... and the context too
    assert c.get_root().root == want["root"]          # the commit still works
    t = O.random_elems(field, c.n_rows, seed=9)
    assert (c.collapse(t) == O.collapse(field, want["coeffs"], t, c.n_rows, c.n_per_row)).all()
    c.close()


def test_new_ml_encodings_commit_like_the_oracle():
    """LigeroEncodingRho::new_ml / SdigEncodingS::new_ml (multilinear sizes): power-of-two row lengths, same commit."""
    field, n_vars = P.FT127, 12
    x = O.random_elems(field, 1 << n_vars, seed=12)
    enc = P.LigeroEncoding.new_ml(field, n_vars)
    oenc = O.Encoding.ligero_from_dims(field, enc.n_per_row, enc.n_cols)
    c = P.LcCommit.commit(x, enc)
    assert c.get_n_rows() * c.get_n_per_row() == 1 << n_vars and c.get_root().into_raw() == oenc.commit(x)["root"]
    senc = P.SdigEncoding.new_ml(field, n_vars, seed=3)
    assert senc.n_per_row & (senc.n_per_row - 1) == 0
    soenc = O.Encoding.sdig_from_dims(field, senc.n_per_row, seed=3)
    assert soenc.n_cols == senc.n_cols
    sc = P.LcCommit.commit(x, senc)
    assert sc.get_n_cols() == senc.n_cols and sc.get_root().root == soenc.commit(x)["root"]
