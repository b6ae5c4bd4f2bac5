"""The sharded (multi-GPU) commit behind the C ABI: lcpc_b200_shard_* / lcpc_b200_multi_* (csrc/shard.cu).

CPU (not gpu): the library's partition equals the Python planner the gloo tests exercise (tests/test_dist.py).
GPU: a whole commit + prove through N shards.  Shards only need distinct CONTEXTS, not distinct devices -- peer
windows on the same device are ordinary pointers -- so the complete protocol (peer stores from the encode, epoch
flags, root exchange, partial row combinations, distributed openings) runs on a one-GPU box too; with >= 2 GPUs the
same tests also run across devices, and one test drives one process per GPU over CUDA IPC handles.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import lcpc_b200 as P
from lcpc_b200.dist import make_plan

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("shape", [(256, 65536, 131072, 8), (72, 235173, 357699, 8), (72, 235173, 357699, 4),
                                   (2, 512, 1024, 2), (64, 16384, 32768, 1), (286, 940690, 1430790, 8), (3, 8, 16, 8),
                                   (1, 300, 457, 2), (5, 1, 2, 4), (1024, 262144, 524288, 16), (18, 58794, 89426, 3)])
def test_library_plan_equals_python_planner(shape):
    n_rows, n_per_row, n_cols, world = shape
    p, q = make_plan(n_rows, n_per_row, n_cols, world), P.shard_plan(n_rows, n_per_row, n_cols, world)
    assert q["row_lo"] == p.row_lo and q["col_lo"] == p.col_lo and q["sub_lo"] == p.sub_lo
    assert q["sub_leaves"] == p.sub_leaves and q["n_sub"] == p.n_sub


def test_plan_rejects_bad_worlds():
    with pytest.raises(P.LcpcError):
        P.shard_plan(4, 4, 8, 0)
    with pytest.raises(P.LcpcError):
        P.shard_plan(4, 4, 8, 17)


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _encodings(kind, field, length, devices):
    ctxs = [P.Context(d) for d in devices]
    if kind == "ligero":
        return [P.LigeroEncoding(field, length, ctx=c) for c in ctxs]
    return [P.SdigEncoding(field, length, seed=0, ctx=c) for c in ctxs]


def _oracle(kind, field, length):
    import oracle as O
    return O.Encoding.ligero(field, length) if kind == "ligero" else O.Encoding.sdig(field, length, seed=0)


CASES = [("ligero", P.FT255, 1 << 14, 2), ("ligero", P.FT255, 1 << 16, 3), ("ligero", P.FT127, (1 << 13) - 7, 4),
         ("sdig", P.FT127, 1 << 14, 2), ("sdig", P.FT255, 5000, 3), ("ligero", P.FT255, 1 << 10, 1),
         ("ligero", P.FT255, 1 << 12, 8), ("sdig", P.FT127, 1 << 13, 5)]


@pytest.mark.gpu
@pytest.mark.parametrize("spread", ["one device", "all devices"])
@pytest.mark.parametrize("kind,field,length,world", CASES)
def test_multi_commit_and_prove_equal_the_oracle(kind, field, length, world, spread):
    """lcpc_b200_commit_new_multi / _multi_prove: LcRoot, every proof field and the wire image equal the oracle's
    single-process commit + prove; the proof verifies on one GPU."""
    import oracle as O
    from oracle import protocol as PR
    from oracle.transcript import Transcript as OTranscript
    n_dev = _n_gpus()
    if spread == "all devices" and n_dev < 2:
        pytest.skip("needs >= 2 GPUs")
    devices = [0] * world if spread == "one device" else [g % n_dev for g in range(world)]
    encs = _encodings(kind, field, length, devices)
    oenc = _oracle(kind, field, length)
    x = O.random_elems(field, length, seed=length % 97)
    oc = oenc.commit(x)
    mc = P.MultiCommit.commit(x, encs)
    assert mc.get_root().root == oc["root"]
    # a second run into the same object with other data, then back: no stale tiles, flags keep counting
    y = O.random_elems(field, length, seed=5)
    mc.rerun(y)
    assert mc.get_root().root == oenc.commit(y)["root"]
    mc.rerun(x)
    assert mc.get_root().root == oc["root"]
    # prove pieces
    t = O.random_elems(field, mc.n_rows, seed=1)
    assert (mc.collapse(t) == O.collapse(field, oc["coeffs"], t, mc.n_rows, mc.n_per_row)).all()
    cols = np.array([0, mc.n_cols - 1, mc.n_cols // 2, 1, mc.n_cols // 3], np.uint64)
    vals, paths = mc.open_columns(cols)
    comm = oc["comm"].reshape(mc.n_rows, mc.n_cols, -1)
    for i, c in enumerate(cols):
        assert (vals[i] == comm[:, int(c)]).all()
        assert O.verify_column_path(field, vals[i], paths[i], int(c), oc["root"])
    with pytest.raises(P.LcpcError):
        mc.open_columns(np.array([mc.n_cols], np.uint64))
    # whole prove(): equal to the oracle's proof, verifies on a single GPU
    outer, inner = O.random_elems(field, mc.n_rows, seed=2), O.random_elems(field, mc.n_per_row, seed=3)
    proof = mc.prove(outer, encs[0], P.Transcript(b"shard test"))
    oproof = PR.prove(field, oc, outer, oenc.get_n_degree_tests(), oenc.get_n_col_opens(), OTranscript(b"shard test"))
    assert P.serialize_proof(proof) == PR.wire_proof(oproof)
    ev = proof.verify(mc.get_root(), outer, inner, encs[0], P.Transcript(b"shard test"))
    assert (ev == O.dot(field, inner, oproof["p_eval"])).all()
    with pytest.raises(P.LcpcError):
        mc.prove(outer[:-1], encs[0], P.Transcript(b"shard test"))
    mc.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind,field,length,world", [("ligero", P.FT255, 1 << 15, 3), ("sdig", P.FT127, 1 << 14, 4)])
def test_every_shard_holds_its_column_block_and_the_root(kind, field, length, world):
    """After a commit, shard g's receive matrix is columns [col_lo[g], col_lo[g+1]) of the oracle's comm, written there
    by all ranks' encode kernels; every shard ends with the same LcRoot.  Repeated commits alternate buffers."""
    import oracle as O
    encs = _encodings(kind, field, length, [0] * world)
    oenc = _oracle(kind, field, length)
    x = O.random_elems(field, length, seed=9)
    oc = oenc.commit(x)
    mc = P.MultiCommit.commit(x, encs)
    for _ in range(3):  # epochs advance, buffers alternate
        mc.rerun()
    assert mc.get_root().root == oc["root"]
    comm = oc["comm"].reshape(mc.n_rows, mc.n_cols, -1)
    covered = 0
    for g in range(world):
        s = mc.shard(g)
        assert s.get_root().root == oc["root"]
        assert (s.local_columns() == comm[:, s.col_lo:s.col_hi]).all()
        assert (s.local_leaves() == oc["hashes"][s.col_lo:s.col_hi]).all()
        assert len(s.phase_times()) == 3
        covered += s.col_hi - s.col_lo
    assert covered == mc.n_cols
    mc.close()


@pytest.mark.gpu
def test_unconnected_shard_refuses_to_commit():
    enc = P.LigeroEncoding(P.FT127, 1 << 10)
    s = P.Shard(enc, 1 << 10, 2, 0)
    with pytest.raises(P.LcpcError):
        s.commit()
    s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind,field,length", [("ligero", P.FT255, 1 << 16), ("sdig", P.FT127, 1 << 14)])
def test_one_process_per_gpu_over_ipc_handles(kind, field, length):
    """ShardedCommit under torchrun: windows mapped through CUDA IPC handles (gloo carries the 64-byte handles);
    with one GPU both processes share it -- the mapping path is the same."""
    n_dev = _n_gpus()
    if n_dev < 1:
        pytest.skip("needs a GPU")
    world = min(max(n_dev, 2), 8)
    from test_dist import free_port
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(HERE, "shard_worker.py"), kind, str(field), str(length)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    recs = []
    dec = json.JSONDecoder()
    for l in out.stdout.splitlines():
        pos = l.find("{")
        while pos >= 0:
            try:
                obj, end = dec.raw_decode(l, pos)
            except json.JSONDecodeError:
                break
            recs.append(obj)
            pos = l.find("{", end)
    assert out.returncode == 0 and len(recs) == world, out.stderr[-3000:]
    for d in recs:
        assert d["root_ok"] and d["cols_ok"] and d["prove_ok"] and d["verify_ok"], d


# ---- the same through the header alone: a plain C program, one process, N GPUs ----------------------------------
ROOT = os.path.dirname(HERE)
C_SRC = os.path.join(HERE, "host", "multi_gpu_check.c")
C_EXE = os.path.join(HERE, "host", "multi_gpu_check")
LIBDIR = os.path.join(ROOT, "lcpc_b200", "lib")


@pytest.fixture(scope="module")
def c_exe():
    deps = [C_SRC, os.path.join(ROOT, "include", "lcpc_b200.h"), os.path.join(ROOT, "include", "lcpc_b200_host.h"),
            os.path.join(LIBDIR, "liblcpc_b200.so")]
    if not os.path.exists(C_EXE) or os.path.getmtime(C_EXE) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["gcc", "-std=c11", "-O1", "-Wall", "-Wextra", "-Werror", "-o", C_EXE, C_SRC, "-L" + LIBDIR,
                               "-llcpc_b200", "-Wl,-rpath," + LIBDIR])
    return C_EXE


def test_header_is_plain_c_and_there_is_no_cpu_fallback(c_exe):
    """include/lcpc_b200.h compiles as C11 (what bindgen / cgo / a Rust `extern "C"` block consume); without a device
    the program stops at context creation."""
    if _n_gpus() > 0:
        pytest.skip("a device is present")
    out = subprocess.run([c_exe, "2", "ligero", "12"], capture_output=True, text=True)
    assert out.returncode == 4 and "no device" in out.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("n_gpus,kind,lg", [(2, "ligero", 16), (4, "sdig", 14), (8, "ligero", 18), (3, "sdig", 16)])
def test_c_host_drives_n_gpus_through_the_header_alone(c_exe, n_gpus, kind, lg):
    n_dev = _n_gpus()
    for devices in ([0] * n_gpus, [g % n_dev for g in range(n_gpus)]):
        out = subprocess.run([c_exe, str(n_gpus), kind, str(lg), ",".join(map(str, devices))], capture_output=True,
                             text=True, timeout=600)
        assert out.returncode == 0 and out.stdout.strip().endswith("ok"), (devices, out.stdout, out.stderr)
        if n_dev < 2:
            break


@pytest.mark.gpu
@pytest.mark.parametrize("chunks", [2, 3, 5])
def test_two_stream_chunked_encode_is_result_neutral(chunks):
    """SHARD_ENC_CHUNKS: a shard's row block encoded in chunks that alternate between two streams must leave the same
    column blocks and LcRoot (rows of a chunk go to their own rows of the owners' matrices)."""
    import oracle as O
    from lcpc_b200 import _cabi
    field, length, world = P.FT255, 1 << 16, 2
    lib = _cabi.lib()
    try:
        lib.lcpc_b200_set_tunable(b"SHARD_ENC_CHUNKS", chunks)
        encs = _encodings("ligero", field, length, [0] * world)
        x = O.random_elems(field, length, seed=chunks)
        oc = O.Encoding.ligero(field, length).commit(x)
        mc = P.MultiCommit.commit(x, encs)
        for g in range(world):
            s = mc.shard(g)
            s.load_rows(x[s.row_lo * s.n_per_row:s.row_lo * s.n_per_row + s.n_elems])
        # device-resident rows take the chunked route
        for g in range(world):
            mc.shard(g).enc.ctx.synchronize()
        shards = [mc.shard(g) for g in range(world)]
        for step in (1, 2, 3):
            for s in shards:
                assert _cabi.lib().lcpc_b200_shard_commit_step(s._h, step) == 0
        comm = oc["comm"].reshape(mc.n_rows, mc.n_cols, -1)
        for s in shards:
            assert s.get_root().root == oc["root"]
            assert (s.local_columns() == comm[:, s.col_lo:s.col_hi]).all()
        mc.close()
    finally:
        lib.lcpc_b200_set_tunable(b"SHARD_ENC_CHUNKS", 4)


@pytest.mark.gpu
@pytest.mark.parametrize("pipeline", [1, 0])
@pytest.mark.parametrize("kind,field,length,world", [("ligero", P.FT255, 1 << 16, 2), ("sdig", P.FT127, 1 << 14, 3), ("ligero", P.FT127, 5000, 4)])
def test_pipelined_hash_stream_is_result_neutral(kind, field, length, world, pipeline):
    """SHARD_PIPELINE=1: exchange wait, hashing and the tree of commit k run on a second stream under the encode of
    commit k+1.  Several commits back to back with alternating inputs (so that a receive matrix overwritten too early,
    or a root read too early, would show), then prove pieces, then a whole prove()."""
    import oracle as O
    from oracle import protocol as PR
    from oracle.transcript import Transcript as OTranscript
    from lcpc_b200 import _cabi
    lib = _cabi.lib()
    try:
        lib.lcpc_b200_set_tunable(b"SHARD_PIPELINE", pipeline)
        encs = _encodings(kind, field, length, [0] * world)
        oenc = _oracle(kind, field, length)
        xs = [O.random_elems(field, length, seed=40 + i) for i in range(2)]
        ocs = [oenc.commit(x) for x in xs]
        mc = P.MultiCommit.commit(xs[0], encs)
        assert mc.get_root().root == ocs[0]["root"]
        for i in range(1, 6):  # enqueue five commits without looking at any result in between
            mc.rerun(xs[i % 2])
        assert mc.get_root().root == ocs[1]["root"]
        comm = ocs[1]["comm"].reshape(mc.n_rows, mc.n_cols, -1)
        for g in range(world):
            s = mc.shard(g)
            assert (s.local_columns() == comm[:, s.col_lo:s.col_hi]).all()
            assert (s.local_leaves() == ocs[1]["hashes"][s.col_lo:s.col_hi]).all()
        outer = O.random_elems(field, mc.n_rows, seed=2)
        proof = mc.prove(outer, encs[0], P.Transcript(b"pipelined"))
        oproof = PR.prove(field, ocs[1], outer, oenc.get_n_degree_tests(), oenc.get_n_col_opens(), OTranscript(b"pipelined"))
        assert P.serialize_proof(proof) == PR.wire_proof(oproof)
        mc.rerun(xs[0])  # a commit after a prove: the streams are joined again
        assert mc.get_root().root == ocs[0]["root"]
        mc.close()
    finally:
        lib.lcpc_b200_set_tunable(b"SHARD_PIPELINE", 1)
