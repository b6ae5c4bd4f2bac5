"""The C++ host-side mirror (include/lcpc_b200.hpp) against the in-tree library: compiles with the system g++, its
host-only parts run on CPU, the full commit / prove / verify flow on the GPU is compared with the Python path."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host", "cpp_mirror_check.cpp")
EXE = os.path.join(ROOT, "tests", "host", "cpp_mirror_check")
LIBDIR = os.path.join(ROOT, "lcpc_b200", "lib")
MERLIN_VECTOR = "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"


@pytest.fixture(scope="module")
def exe():
    deps = [SRC, os.path.join(ROOT, "include", "lcpc_b200.hpp"), os.path.join(ROOT, "include", "lcpc_b200.h"),
            os.path.join(ROOT, "include", "lcpc_b200_host.h"), os.path.join(LIBDIR, "liblcpc_b200.so")]
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-o", EXE, SRC, "-L" + LIBDIR, "-llcpc_b200",
                               "-Wl,-rpath," + LIBDIR])
    return EXE


def test_cpp_mirror_compiles_and_host_parts_work(exe):
    out = subprocess.run([exe, "host"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.split("\n")
    assert lines[0] == "merlin " + MERLIN_VECTOR
    import torch
    if not torch.cuda.is_available():
        assert lines[1] == "no device: code -4"  # LCPC_B200_ERR_CUDA, never a CPU fallback


def test_cpp_mirror_wire_format_equals_python_writer(exe, tmp_path):
    """serialize / deserialize_proof of the C++ mirror: byte-identical to lcpc_b200.proof (which is held to the oracle's
    element-by-element writer and to the reference's logged proof sizes)."""
    import lcpc_b200 as P
    out_file = str(tmp_path / "wire.bin")
    out = subprocess.run([exe, "wire", out_file], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stderr)
    n_proof, n_root = (int(v) for v in out.stdout.split()[1:3])
    blob = open(out_file, "rb").read()
    assert len(blob) == n_proof + n_root
    L, npr, ndt, ncol, nrows, plen = 4, 512, 2, 7, 3, 10
    M = (1 << 64) - 1
    p_eval = np.array([(0x0101010101010101 * (i % 251) + i) & M for i in range(npr * L)], np.uint64).reshape(npr, L)
    p_rand = np.array([(i * 0x9E3779B97F4A7C15) & M for i in range(ndt * npr * L)], np.uint64).reshape(ndt, npr, L)
    cols = np.array([((~i & M) * 3) & M for i in range(ncol * nrows * L)], np.uint64).reshape(ncol, nrows, L)
    paths = np.array([(i * 7 + 1) & 255 for i in range(ncol * plen * 32)], np.uint8).reshape(ncol, plen, 32)
    want = P.serialize_proof(P.LcEvalProof(P.FT255, 1024, p_eval, p_rand, cols, paths))
    assert blob[:n_proof] == want
    assert blob[n_proof:] == P.serialize_root(P.LcRoot(bytes(200 - i for i in range(32))))


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["ligero", "sdig"])
def test_cpp_mirror_commit_prove_verify_equal_the_python_path(exe, kind):
    import lcpc_b200 as P
    out = subprocess.run([exe, "gpu", kind], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    got = dict(l.split(" ", 1) for l in out.stdout.strip().split("\n") if " " in l)
    assert out.stdout.strip().endswith("ok")
    field, length, L = P.FT127, 3000, 2

    def small(n, mul, add):
        v = np.zeros((n, L), np.uint64)
        v[:, 0] = np.arange(n, dtype=np.uint64) * np.uint64(mul) + np.uint64(add)
        return v

    enc = P.LigeroEncoding(field, length) if kind == "ligero" else P.SdigEncoding(field, length, seed=5)
    c = P.LcCommit.commit(small(length, 7, 1), enc)
    assert got["root"] == c.get_root().root.hex()
    outer, inner = small(c.n_rows, 11, 3), small(c.n_per_row, 13, 2)
    proof = c.prove(outer, enc, P.Transcript(b"cpp mirror"))
    ev = proof.verify(c.get_root(), outer, inner, enc, P.Transcript(b"cpp mirror"))
    assert got["eval"] == ev.tobytes().hex()
    assert got["cols"].split() == [str(int(v)) for v in proof.col_idx[:3]]
