"""CPU-side tests of the product's host layer and of the C-ABI library as a binary artefact:
the library loads, exports every symbol the headers declare, its host-side setup mirrors
(dimension choosers, matgen) agree with the oracle, and compute entry points fail loudly -- not fall
back -- when there is no CUDA device.  No kernels run here.
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import oracle as O
import lcpc_b200 as P
from lcpc_b200 import _cabi, host

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for hdr in ("lcpc_b200.h", "lcpc_b200_host.h"):
        src = open(os.path.join(ROOT, "include", hdr)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(lcpc_b200_[a-z0-9_]+)\s*\(", src))
    return names


def test_library_exports_every_declared_symbol():
    lib = _cabi.lib()
    declared = _declared_symbols()
    assert len(declared) >= 45
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    # and the Python binding table covers exactly the declared set
    assert set(_cabi.SIGNATURES) == declared
    assert b"sm_100a" in lib.lcpc_b200_version()


def test_library_holds_sm100a_code_only():
    out = subprocess.run(["cuobjdump", "--list-elf", _cabi.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(P.LcpcError) as ei:
        P.Context(0)
    assert ei.value.code == _cabi.ERR_CUDA
    with pytest.raises(P.LcpcError):
        P.LigeroEncoding(P.FT255, 1 << 10)


def test_product_does_not_touch_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "lcpc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in src and "oracle/" not in src and "lcpc_oracle" not in src, f


def test_field_limbs():
    lib = _cabi.lib()
    assert [lib.lcpc_b200_field_limbs(f) for f in (1, 2, 3, 4, 9)] == [1, 2, 3, 4, -1]
    assert [lib.lcpc_b200_field_flog2(f) for f in (1, 2, 3, 4)] == [62, 126, 190, 254]


@pytest.mark.parametrize("field", [P.FT63, P.FT127, P.FT255])
def test_ligero_dims_match_oracle(field):
    rng = np.random.default_rng(field)
    for lgl in range(2, 30, 2):
        for rho in [(1, 2), (1, 4), (39, 40)]:
            length = int(rng.integers(1 << lgl, 2 << lgl))
            assert P.ligero_get_dims(field, length, rho) == O.ligero_get_dims(field, length, rho)
    assert P.ligero_get_dims(P.FT255, 1 << 24) == (256, 65536, 131072)
    assert _cabi.lib().lcpc_b200_ligero_n_col_opens(1, 2) == 309 == O.lib().lcpc_oracle_ligero_n_col_opens(1, 2)
    for args in [(128, 1 << 17, 254), (128, 357699, 126), (128, 1024, 62)]:
        assert P.n_degree_tests(*args) == O.n_degree_tests(*args)


def test_sdig_parameters_match_oracle():
    lib = _cabi.lib()
    for code in range(1, 7):
        assert lib.lcpc_b200_sdig_n_col_opens(code) == O.lib().lcpc_oracle_sdig_n_col_opens(code)
    assert lib.lcpc_b200_sdig_n_col_opens(3) == 6593
    for field, length in [(P.FT127, 1 << 24), (P.FT127, 1 << 20), (P.FT255, 1 << 16), (P.FT63, 5000)]:
        npr = C.c_size_t()
        assert lib.lcpc_b200_sdig_choose_n_per_row(field, 3, length, C.byref(npr)) == 0
        want = O.Encoding.sdig(field, length, seed=0).n_per_row if length <= (1 << 16) else None
        if want is not None:
            assert npr.value == want
    npr = C.c_size_t()
    lib.lcpc_b200_sdig_choose_n_per_row(P.FT127, 3, 1 << 24, C.byref(npr))
    assert npr.value == 235173


@pytest.mark.parametrize("field,n,seed,code", [(P.FT63, 400, 3, 3), (P.FT127, 3000, 1, 3), (P.FT255, 700, 7, 3),
                                               (P.FT191, 300, 2, 3), (P.FT127, 2000, 5, 1), (P.FT63, 2500, 0, 6)])
def test_matgen_matches_oracle(field, n, seed, code):
    """matgen::generate restated twice (oracle C, product C++): same seeded CSC arrays."""
    pre, post, cw = host.generate_sdig_code(field, n, seed, code)
    enc = O.Encoding.sdig_from_dims(field, n, seed=seed, code=code)
    opre, opost = enc.matrices()
    assert cw == enc.n_cols and len(pre) == len(opre)
    for a, b in list(zip(pre, opre)) + list(zip(post, opost)):
        assert (a["m"], a["n"]) == (b["m"], b["n"])
        for k in ("ptrs", "idxs", "data"):
            assert (a[k] == b[k]).all()


def test_matgen_rejects_tiny_rows():
    with pytest.raises(P.LcpcError):
        host.generate_sdig_code(P.FT127, 20, 0)  # assert!(n > baselen), matgen.rs:62


@pytest.fixture(scope="module")
def hostcheck():
    src = os.path.join(ROOT, "tests", "host", "field_host_check.cu")
    out = os.path.join(ROOT, "tests", "host", "libfield_host_check.so")
    if not os.path.exists(out) or os.path.getmtime(out) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(ROOT, "lcpc_b200", "csrc", "field.cuh"))):
        subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
                               "-Wno-deprecated-gpu-targets", "-o", out, src])
    return C.CDLL(out)


@pytest.mark.parametrize("field", [O.FT63, O.FT127, O.FT191, O.FT255])
def test_field_carry_chains_on_host_emulation(hostcheck, field):
    """field.cuh's interleaved Montgomery product / add / sub / from_mont, run through the header's
    host emulation of the PTX carry flag, equal the oracle's u64-limb arithmetic."""
    n = 20000
    a, b = O.random_elems(field, n, seed=1), O.random_elems(field, n, seed=2)
    p = O.field_info(field)["modulus"]
    nl = O.FIELD_LIMBS[field]
    edge = O.ints_to_elems([0, 1, p - 1, p - 2, (1 << (64 * nl - 1)) % p, 0xffffffff, 1 << 32, p - 0xffffffff], field)
    a[:8], b[:8] = edge, edge[::-1]
    a[8:16], b[8:16] = edge, edge
    for op, name in [(0, "add"), (1, "sub"), (2, "mul"), (4, "from_mont")]:
        r = np.empty_like(a)
        rc = hostcheck.hostcheck_field_op(field, op, r.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p),
                                          b.ctypes.data_as(C.c_void_p), C.c_size_t(n))
        assert rc == 0
        assert (r == O.field_op(field, name, a, b if op < 3 else None)).all(), name
    # product and reduction done separately (mul_full + redc) == the interleaved product
    r = np.empty_like(a)
    assert hostcheck.hostcheck_field_op(field, 5, r.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p),
                                        b.ctypes.data_as(C.c_void_p), C.c_size_t(n)) == 0
    assert (r == O.field_op(field, "mul", a, b)).all()
    # one-level subtractive Karatsuba product + reduction == the interleaved product
    r = np.empty_like(a)
    assert hostcheck.hostcheck_field_op(field, 7, r.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p),
                                        b.ctypes.data_as(C.c_void_p), C.c_size_t(n)) == 0
    assert (r == O.field_op(field, "mul", a, b)).all()
    # 37-term sums of products accumulated double-width and reduced once == reduce-every-product
    m = 3000
    r = np.empty_like(a[:m])
    assert hostcheck.hostcheck_field_op(field, 6, r.ctypes.data_as(C.c_void_p), a[:m].ctypes.data_as(C.c_void_p),
                                        b[:m].ctypes.data_as(C.c_void_p), C.c_size_t(m)) == 0
    want = O.ints_to_elems([0] * m, field)
    idx = np.arange(m)
    for k in range(37):
        want = O.field_op(field, "add", want, O.field_op(field, "mul", a[:m][(idx + k) % m], b[:m][(idx * 7 + k) % m]))
    assert (r == want).all()
    # R mod p folded at compile time (Field::one) == 2^(64 L) mod p, the value SURVEY.md's appendix A lists
    r = np.empty_like(a[:1])
    assert hostcheck.hostcheck_field_op(field, 9, r.ctypes.data_as(C.c_void_p), a[:1].ctypes.data_as(C.c_void_p),
                                        b[:1].ctypes.data_as(C.c_void_p), C.c_size_t(1)) == 0
    assert (r == O.ints_to_elems([(1 << (64 * nl)) % p], field)).all()
    assert (r == O.ints_to_elems([{O.FT63: 0x2b8e9dfffffffffd, O.FT127: 0x23157ed08bbe3e8101a84dfffffffffe,
                                   O.FT191: 0x305ae60140ca567045c6678ad9339c96892c79fffffffffd,
                                   O.FT255: 0x33870cc92365adfe04ac41f68d514d2c211870def34d419ffab61bfffffffffe}[field]], field)).all()
    # the same sums through the carry-counting accumulator (Field::Sum: sum_mac / sum_reduce), 1..300 terms
    for terms in (1, 2, 37, 300):
        r = np.empty_like(a[:m])
        assert hostcheck.hostcheck_field_op(field, 8 | (terms << 8), r.ctypes.data_as(C.c_void_p), a[:m].ctypes.data_as(C.c_void_p),
                                            b[:m].ctypes.data_as(C.c_void_p), C.c_size_t(m)) == 0
        want = O.ints_to_elems([0] * m, field)
        for k in range(terms):
            want = O.field_op(field, "add", want, O.field_op(field, "mul", a[:m][(idx + k) % m], b[:m][(idx * 7 + k) % m]))
        assert (r == want).all(), terms
    # ... split over four accumulators whose collected sums are added (the lanes of spmm_sum_split_kernel)
    for terms in (3, 45, 301):
        r = np.empty_like(a[:m])
        assert hostcheck.hostcheck_field_op(field, 10 | (terms << 8), r.ctypes.data_as(C.c_void_p), a[:m].ctypes.data_as(C.c_void_p),
                                            b[:m].ctypes.data_as(C.c_void_p), C.c_size_t(m)) == 0
        want = O.ints_to_elems([0] * m, field)
        for k in range(terms):
            want = O.field_op(field, "add", want, O.field_op(field, "mul", a[:m][(idx + k) % m], b[:m][(idx * 7 + k) % m]))
        assert (r == want).all(), terms
    # ... and its worst case: 100 000 terms of (p-1)^2 (the carry counters and the quotient estimate of sum_reduce)
    R = 1 << (64 * nl)
    big = O.ints_to_elems([p - 1] * 8, field)
    for terms in (1, 37, 100000):
        r = np.empty_like(big)
        assert hostcheck.hostcheck_field_op(field, 8 | (terms << 8), r.ctypes.data_as(C.c_void_p), big.ctypes.data_as(C.c_void_p),
                                            big.ctypes.data_as(C.c_void_p), C.c_size_t(8)) == 0
        assert (r == O.ints_to_elems([terms * (p - 1) * (p - 1) * pow(R, -1, p) % p] * 8, field)).all(), terms
    # worst case for the fold bound: every term is (p-1)^2
    big = O.ints_to_elems([p - 1] * 64, field)
    r = np.empty_like(big)
    assert hostcheck.hostcheck_field_op(field, 6, r.ctypes.data_as(C.c_void_p), big.ctypes.data_as(C.c_void_p),
                                        big.ctypes.data_as(C.c_void_p), C.c_size_t(64)) == 0
    sq = O.field_op(field, "mul", big, big)
    want = O.ints_to_elems([0] * 64, field)
    for k in range(37):
        want = O.field_op(field, "add", want, sq)
    assert (r == want).all()


def test_sdig_new_ml_chooses_a_power_of_two_first_guess():
    """SdigEncodingS::new_ml (lcpc-brakedown-pc/src/lib.rs:114-124) restated here in Python: the first guess is rounded
    up to a power of two, then _new_from_np1 (:69-99) keeps it or halves it."""
    import ctypes as C
    import math
    lib = _cabi.lib()
    for field, flog2 in ((P.FT127, 126), (P.FT255, 254)):
        for n_vars in (12, 16, 20, 24, 28):
            n = 1 << n_vars
            opens = lib.lcpc_b200_sdig_n_col_opens(3)
            lncf = float(opens * n)
            ndt = host.n_degree_tests(128, math.ceil(math.sqrt(lncf)) * 2, flog2)
            np1 = 1 << (math.ceil(math.sqrt(lncf / ndt)) - 1).bit_length()
            np1 = min(np1, n)
            np2 = np1 // 2
            sz = lambda npr: opens * -(-n // npr) + (1 + host.n_degree_tests(128, npr * 2, flog2)) * npr
            want = np1 if sz(np1) < sz(np2) else np2
            got = C.c_size_t()
            assert lib.lcpc_b200_sdig_choose_n_per_row_ml(field, 3, n_vars, C.byref(got)) == 0
            assert got.value == want and want & (want - 1) == 0
