"""lcpc_b200 -- B200-native commit/prove engine for lcpc-2d's data-parallel hot path.

Host-side mirror of the reference's operator interface, over the C ABI in ``include/lcpc_b200.h``:

=====================================  ==========================================================
reference (Rust)                       here
=====================================  ==========================================================
``trait LcEncoding`` (lcpc-2d lib.rs   ``LcEncoding`` base class: ``encode``, ``get_dims``,
:74-104)                               ``dims_ok``, ``get_n_col_opens``, ``get_n_degree_tests``
``LigeroEncoding<F>`` (ligero lib.rs   ``LigeroEncoding(field, len)`` / ``.new_from_dims``
:189)
``SdigEncoding<F>`` (brakedown lib.rs  ``SdigEncoding(field, len, seed)`` / ``.new_from_dims``
:179)
``LcCommit::commit / get_root /        ``LcCommit.commit(coeffs, enc)``, ``.get_root()``,
prove`` (lib.rs:276-311)               ``.collapse(tensor)``, ``.open_columns(cols)``
``LcRoot`` (lib.rs:315-323)            ``LcRoot`` (32-byte digest)
``LcCommit::prove`` (lib.rs:304-311)   ``LcCommit.prove(outer_tensor, enc, tr)`` -> ``LcEvalProof``
``LcEvalProof::verify`` (:518-527)     ``LcEvalProof.verify(root, outer, inner, enc, tr)``
``merlin::Transcript``                 ``Transcript(label)``
serde / bincode of the three types     ``serialize_*`` / ``deserialize_*`` (``lcpc_b200.proof``)
=====================================  ==========================================================

Field elements are ``numpy.uint64`` arrays of shape ``(n, L)``: the in-memory image of the reference's
``struct FtNNN([u64; L])`` (Montgomery limbs, little-endian).  All compute happens on the GPU through
the C ABI; there is no CPU implementation in this package.
"""
from .host import (FT63, FT127, FT191, FT255, FIELD_LIMBS, Context, LcCommit, LcEncoding, LcRoot,  # noqa: F401
                   LigeroEncoding, SdigEncoding, default_context, field_op, merkleize, collapse_columns,
                   ligero_get_dims, n_degree_tests, expand_tensor)
from ._cabi import LcpcError, LIB_PATH  # noqa: F401
from .proof import (LcEvalProof, Transcript, prove, sample_columns, serialize_root, serialize_commit,  # noqa: F401
                    serialize_proof, deserialize_root, deserialize_commit_fields, deserialize_commit, deserialize_proof)

from .shard import MultiCommit, Shard, ShardedCommit, shard_plan  # noqa: F401,E402

__all__ = ["MultiCommit", "Shard", "ShardedCommit", "shard_plan", "FT63", "FT127", "FT191", "FT255", "FIELD_LIMBS", "Context", "LcCommit", "LcEncoding", "LcRoot",
           "LigeroEncoding", "SdigEncoding", "LcpcError", "default_context", "field_op", "merkleize",
           "collapse_columns", "ligero_get_dims", "n_degree_tests", "expand_tensor", "LcEvalProof", "Transcript", "prove",
           "sample_columns", "serialize_root", "serialize_commit", "serialize_proof", "deserialize_root",
           "deserialize_commit_fields", "deserialize_commit", "deserialize_proof"]
