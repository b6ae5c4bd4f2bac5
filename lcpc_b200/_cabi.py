"""ctypes binding of lcpc_b200/lib/liblcpc_b200.so (the C ABI of include/lcpc_b200.h).

Loads the in-tree CUDA library and nothing else: there is no fallback implementation.  Importing this
module works without a GPU (the library only needs the CUDA runtime it links statically); creating
a context without one raises ``LcpcError``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# LCPC_B200_LIB selects another build of the same library (A/B tuning builds); default: the in-tree one
LIB_PATH = os.environ.get("LCPC_B200_LIB") or os.path.join(_HERE, "lib", "liblcpc_b200.so")

OK = 0
ERR_BAD_ARG, ERR_TOO_BIG, ERR_ENCODE, ERR_CUDA, ERR_OOM, ERR_COLUMN, ERR_UNSUPPORTED = -1, -2, -3, -4, -5, -6, -7
ERR_OUTER_TENSOR = -8
# VerifierError (lcpc-2d/src/lib.rs:141-170)
(VERR_NUM_COL_OPENS, VERR_COLUMN_PATH, VERR_COLUMN_EVAL, VERR_COLUMN_DEGREE, VERR_OUTER_TENSOR, VERR_INNER_TENSOR,
 VERR_ENCODING_DIMS) = -20, -21, -22, -23, -24, -25, -26
_ERR_NAMES = {ERR_BAD_ARG: "BAD_ARG", ERR_TOO_BIG: "TOO_BIG", ERR_ENCODE: "ENCODE", ERR_CUDA: "CUDA",
              ERR_OOM: "OOM", ERR_COLUMN: "COLUMN", ERR_UNSUPPORTED: "UNSUPPORTED", ERR_OUTER_TENSOR: "OUTER_TENSOR",
              VERR_NUM_COL_OPENS: "VerifierError::NumColOpens", VERR_COLUMN_PATH: "VerifierError::ColumnPath",
              VERR_COLUMN_EVAL: "VerifierError::ColumnEval", VERR_COLUMN_DEGREE: "VerifierError::ColumnDegree",
              VERR_OUTER_TENSOR: "VerifierError::OuterTensor", VERR_INNER_TENSOR: "VerifierError::InnerTensor",
              VERR_ENCODING_DIMS: "VerifierError::EncodingDims"}


class LcpcError(RuntimeError):
    """A non-zero status from the C ABI (the Rust shim maps these onto ProverError)."""

    def __init__(self, code, message=""):
        self.code = code
        super().__init__(f"lcpc_b200: {_ERR_NAMES.get(code, code)} {message}".rstrip())


class Csc(C.Structure):
    _fields_ = [("m", C.c_size_t), ("n", C.c_size_t), ("ptrs", C.c_void_p), ("idxs", C.c_void_p),
                ("data", C.c_void_p)]


class Scatter(C.Structure):
    _fields_ = [("n_blocks", C.c_size_t), ("starts", C.c_void_p), ("dst", C.c_void_p), ("row0", C.c_size_t)]


class Labels(C.Structure):
    _fields_ = [("dt", C.c_char_p), ("pr", C.c_char_p), ("pe", C.c_char_p), ("co", C.c_char_p),
                ("dt_len", C.c_size_t), ("pr_len", C.c_size_t), ("pe_len", C.c_size_t), ("co_len", C.c_size_t)]


class Proof(C.Structure):
    _fields_ = [("n_cols", C.c_size_t), ("n_per_row", C.c_size_t), ("n_degree_tests", C.c_size_t),
                ("n_columns", C.c_size_t), ("n_rows", C.c_size_t), ("path_len", C.c_size_t),
                ("p_eval", C.c_void_p), ("p_random", C.c_void_p), ("cols", C.c_void_p), ("paths", C.c_void_p)]


_vp, _sz, _i, _u64 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint64
_pvp, _psz = C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)

# name -> (restype, argtypes); mirrors include/lcpc_b200.h and include/lcpc_b200_host.h one to one
SIGNATURES = {
    "lcpc_b200_version": (C.c_char_p, []),
    "lcpc_b200_set_tunable": (_i, [C.c_char_p, C.c_long]),
    "lcpc_b200_get_tunable": (C.c_long, [C.c_char_p, C.c_long]),
    "lcpc_b200_field_limbs": (_i, [_i]),
    "lcpc_b200_field_one": (_i, [_i, _vp]),
    "lcpc_b200_ctx_create": (_i, [_i, _pvp]),
    "lcpc_b200_ctx_destroy": (None, [_vp]),
    "lcpc_b200_last_error": (C.c_char_p, [_vp]),
    "lcpc_b200_ctx_device": (_i, [_vp]),
    "lcpc_b200_ctx_stream": (_vp, [_vp]),
    "lcpc_b200_ctx_synchronize": (_i, [_vp]),
    "lcpc_b200_ctx_launch_count": (_u64, [_vp]),
    "lcpc_b200_host_alloc": (_i, [_sz, _pvp]),
    "lcpc_b200_host_free": (None, [_vp]),
    "lcpc_b200_host_register": (_i, [_vp, _sz]),
    "lcpc_b200_host_unregister": (_i, [_vp]),
    "lcpc_b200_ligero_new": (_i, [_vp, _i, _sz, _sz, _pvp]),
    "lcpc_b200_sdig_new": (_i, [_vp, _i, _sz, C.POINTER(Csc), C.POINTER(Csc), _pvp]),
    "lcpc_b200_enc_free": (None, [_vp]),
    "lcpc_b200_enc_kind": (_i, [_vp]),
    "lcpc_b200_enc_field": (_i, [_vp]),
    "lcpc_b200_enc_get_dims": (_i, [_vp, _sz, _psz, _psz, _psz]),
    "lcpc_b200_enc_dims_ok": (_i, [_vp, _sz, _sz]),
    "lcpc_b200_encode": (_i, [_vp, _vp, _sz]),
    "lcpc_b200_encode_dev": (_i, [_vp, _vp, _sz, _sz]),
    "lcpc_b200_commit_new": (_i, [_vp, _vp, _sz, _pvp]),
    "lcpc_b200_commit_new_dev": (_i, [_vp, _vp, _sz, _pvp]),
    "lcpc_b200_commit_rerun_dev": (_i, [_vp, _vp, _sz]),
    "lcpc_b200_commit_rerun": (_i, [_vp, _vp, _sz]),
    "lcpc_b200_commit_from_host": (_i, [_vp, _vp, _sz, _vp, _sz, _vp, _sz, _sz, _pvp]),
    "lcpc_b200_commit_free": (None, [_vp]),
    "lcpc_b200_commit_dims": (_i, [_vp, _psz, _psz, _psz, _psz]),
    "lcpc_b200_commit_root": (_i, [_vp, _vp]),
    "lcpc_b200_commit_download": (_i, [_vp, _vp, _vp, _vp]),
    "lcpc_b200_commit_phase_times": (_i, [_vp, C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "lcpc_b200_encode_rows_dev": (_i, [_vp, _vp, _sz, _sz, _vp, _sz]),
    "lcpc_b200_encode_rows_h2d": (_i, [_vp, _vp, _sz, _vp, _vp, _sz]),
    "lcpc_b200_encode_rows_scatter_dev": (_i, [_vp, _vp, _sz, _sz, _vp, _sz, C.POINTER(Scatter)]),
    "lcpc_b200_encode_rows_scatter_h2d": (_i, [_vp, _vp, _sz, _vp, _vp, _sz, C.POINTER(Scatter)]),
    "lcpc_b200_commit_device_ptrs": (_i, [_vp, _pvp, _pvp, _pvp]),
    "lcpc_b200_commit_rerun_to_host": (_i, [_vp, _vp, _sz, _vp, _vp, _vp]),
    "lcpc_b200_commit_to_host": (_i, [_vp, _vp, _sz, _vp, _vp, _vp]),
    "lcpc_b200_commit_collapse": (_i, [_vp, _vp, _vp]),
    "lcpc_b200_commit_degree_test": (_i, [_vp, _vp, _vp, _vp]),
    "lcpc_b200_expand_tensor": (_i, [_vp, _i, _vp, _sz, _vp]),
    "lcpc_b200_collapse": (_i, [_vp, _i, _vp, _vp, _vp, _sz, _sz]),
    "lcpc_b200_commit_open_columns": (_i, [_vp, _vp, _sz, _vp, _vp]),
    "lcpc_b200_commit_prove": (_i, [_vp, _vp, C.POINTER(Labels), _vp, _sz, _sz, _sz, _vp, _vp, _vp, _vp, _vp]),
    "lcpc_b200_verify": (_i, [_vp, _vp, C.POINTER(Labels), _vp, _vp, _sz, _vp, _sz, _sz, _sz, C.POINTER(Proof), _vp]),
    "lcpc_b200_shard_plan": (_i, [_sz, _sz, _sz, C.c_uint, _vp, _vp, _vp, _psz, _psz]),
    "lcpc_b200_shard_new": (_i, [_vp, _sz, C.c_uint, C.c_uint, _sz, _pvp]),
    "lcpc_b200_shard_free": (None, [_vp]),
    "lcpc_b200_shard_window": (_i, [_vp, _pvp, _psz, _vp]),
    "lcpc_b200_shard_connect": (_i, [_vp, _vp, _vp]),
    "lcpc_b200_shard_dims": (_i, [_vp, _psz, _psz, _psz, _psz, _psz, _psz, _psz, _psz]),
    "lcpc_b200_shard_commit": (_i, [_vp, _vp, _sz]),
    "lcpc_b200_shard_commit_dev": (_i, [_vp, _vp, _sz]),
    "lcpc_b200_shard_load_rows": (_i, [_vp, _vp, _sz]),
    "lcpc_b200_shard_commit_step": (_i, [_vp, _i]),
    "lcpc_b200_shard_root": (_i, [_vp, _vp]),
    "lcpc_b200_shard_root_enqueue": (_i, [_vp, _vp]),
    "lcpc_b200_shard_join": (_i, [_vp]),
    "lcpc_b200_shard_phase_times": (_i, [_vp, _vp]),
    "lcpc_b200_shard_device_ptrs": (_i, [_vp, _pvp, _pvp, _pvp, _pvp]),
    "lcpc_b200_shard_download": (_i, [_vp, _vp, _vp]),
    "lcpc_b200_shard_collapse_begin": (_i, [_vp, _vp, _vp]),
    "lcpc_b200_shard_collapse_finish": (_i, [_vp, _vp, _vp]),
    "lcpc_b200_shard_open_begin": (_i, [_vp, _vp, _sz]),
    "lcpc_b200_shard_open_finish": (_i, [_vp, _vp, _vp]),
    "lcpc_b200_shard_prove": (_i, [_vp, _vp, C.POINTER(Labels), _vp, _sz, _sz, _sz, _vp, _vp, _vp, _vp, _vp]),
    "lcpc_b200_commit_new_multi": (_i, [_vp, _sz, _vp, _sz, _sz, _pvp]),
    "lcpc_b200_multi_rerun": (_i, [_vp, _vp, _sz]),
    "lcpc_b200_multi_root": (_i, [_vp, _vp]),
    "lcpc_b200_multi_collapse": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "lcpc_b200_multi_open_columns": (_i, [_vp, _vp, _sz, _vp, _vp]),
    "lcpc_b200_multi_prove": (_i, [_vp, _vp, C.POINTER(Labels), _vp, _sz, _sz, _sz, _vp, _vp, _vp, _vp, _vp]),
    "lcpc_b200_multi_n_shards": (_sz, [_vp]),
    "lcpc_b200_multi_shard": (_vp, [_vp, _sz]),
    "lcpc_b200_multi_free": (None, [_vp]),
    "lcpc_b200_merkleize": (_i, [_vp, _i, _vp, _sz, _sz, _vp]),
    "lcpc_b200_hash_columns_dev": (_i, [_vp, _i, _vp, _sz, _sz, _sz, _vp]),
    "lcpc_b200_merkle_tree_dev": (_i, [_vp, _vp, _sz]),
    "lcpc_b200_merkle_layers_dev": (_i, [_vp, _vp, _sz, C.c_uint]),
    "lcpc_b200_pack_column_blocks_dev": (_i, [_vp, _i, _vp, _sz, _sz, _sz, _vp, _vp]),
    "lcpc_b200_collapse_dev": (_i, [_vp, _i, _vp, _sz, _vp, _vp, _sz, _sz]),
    "lcpc_b200_field_op": (_i, [_vp, _i, _i, _vp, _vp, _vp, _sz]),
    # host-side setup (include/lcpc_b200_host.h)
    "lcpc_b200_n_degree_tests": (_sz, [_sz, _sz, _sz]),
    "lcpc_b200_field_flog2": (C.c_uint, [_i]),
    "lcpc_b200_ligero_n_col_opens": (_sz, [_sz, _sz]),
    "lcpc_b200_ligero_get_dims": (_i, [_i, _sz, _sz, _sz, _psz, _psz, _psz]),
    "lcpc_b200_sdig_n_col_opens": (_sz, [_i]),
    "lcpc_b200_sdig_choose_n_per_row": (_i, [_i, _i, _sz, _psz]),
    "lcpc_b200_sdig_choose_n_per_row_ml": (_i, [_i, _i, _sz, _psz]),
    "lcpc_b200_sdig_code_generate": (_i, [_i, _i, _sz, _u64, _pvp]),
    "lcpc_b200_sdig_code_free": (None, [_vp]),
    "lcpc_b200_sdig_code_levels": (_sz, [_vp]),
    "lcpc_b200_sdig_code_n_per_row": (_sz, [_vp]),
    "lcpc_b200_sdig_code_codeword_length": (_sz, [_vp]),
    "lcpc_b200_sdig_code_matrix": (_i, [_vp, _sz, _i, C.POINTER(Csc)]),
    "lcpc_b200_sdig_new_from_code": (_i, [_vp, _vp, _pvp]),
    "lcpc_b200_sdig_new_seeded": (_i, [_vp, _i, _i, _sz, _u64, _pvp]),
    "lcpc_b200_enc_sdig_levels": (_sz, [_vp]),
    "lcpc_b200_enc_sdig_matrix": (_i, [_vp, _sz, _i, _psz, _psz, _psz, _vp, _vp]),
    "lcpc_b200_transcript_new": (_i, [C.c_char_p, _sz, _pvp]),
    "lcpc_b200_transcript_clone": (_i, [_vp, _pvp]),
    "lcpc_b200_transcript_free": (None, [_vp]),
    "lcpc_b200_transcript_append_message": (_i, [_vp, C.c_char_p, _sz, C.c_char_p, _sz]),
    "lcpc_b200_transcript_append_u64": (_i, [_vp, C.c_char_p, _sz, _u64]),
    "lcpc_b200_transcript_challenge_bytes": (_i, [_vp, C.c_char_p, _sz, _vp, _sz]),
    "lcpc_b200_transcript_append_reprs": (_i, [_vp, C.c_char_p, _sz, _vp, _sz, _sz]),
    "lcpc_b200_sample_columns": (_i, [_vp, _sz, _sz, _vp]),
}

_lib = None


def lib():
    """The loaded C-ABI library; raises if it has not been built (run __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LcpcError(ERR_CUDA, f"{LIB_PATH} is missing: build it with `make -C lcpc_b200/csrc` "
                                      "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib
