/*
 * include/lcpc_b200.h -- C ABI of the B200-native commit/prove engine for lcpc-2d's hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  A Rust `lcpc-2d`
 * built with the shim shown in INTEGRATION.md binds exactly these symbols (bindgen/extern "C").
 * Each entry point names the reference item it replaces (paths relative to the reference repo).
 *
 * Field elements cross the boundary as the in-memory image of the reference's
 * `struct FtNNN([u64; L])` (lcpc-test-fields/src/lib.rs:22,34,46,58): L little-endian u64 limbs in
 * Montgomery form, so a Rust `&[F]` can be passed as `*const u64` without conversion.
 * Digests are 32-byte BLAKE3 outputs (`Output<D>` with D = blake3::Hasher).
 *
 * Every function returns LCPC_B200_OK (0) or a negative status; nothing unwinds across the boundary.
 * `lcpc_b200_last_error(ctx)` returns a human-readable description of the last failure on `ctx`.
 * There is NO CPU fallback: without a CUDA device every compute entry point fails with
 * LCPC_B200_ERR_CUDA.
 *
 * Lifetimes: an encoding holds a reference on its context and a commit on its encoding, so lcpc_b200_ctx_destroy /
 * lcpc_b200_enc_free / lcpc_b200_commit_free may be called in any order (the object is released when its last
 * dependant is); a handle must not be USED after its own free call.
 *
 * Threading: a context serialises its calls internally (one stream, one mutex); use one context per
 * host thread for concurrency.  `LcEncoding::encode` is called from many rayon workers in the
 * reference (lcpc-2d/src/lib.rs:648-653); the batched entry points here replace that loop.
 */
#ifndef LCPC_B200_H
#define LCPC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes; the Rust shim maps them onto ProverError (lcpc-2d/src/lib.rs:111-139) */
#define LCPC_B200_OK 0
#define LCPC_B200_ERR_BAD_ARG (-1)     /* failed `assert!`s of commit()/prove(); ProverError::Commit (:683-684) */
#define LCPC_B200_ERR_TOO_BIG (-2)     /* ProverError::TooBig (:656-658); FFTError::TooBig via precomp_fft */
#define LCPC_B200_ERR_ENCODE (-3)      /* ProverError::Encode(E::Err) (:120-121) */
#define LCPC_B200_ERR_CUDA (-4)        /* CUDA runtime failure or no device; message in last_error */
#define LCPC_B200_ERR_OOM (-5)         /* device or pinned-host allocation failed */
#define LCPC_B200_ERR_COLUMN (-6)      /* ProverError::ColumnNumber (:797-799) */
#define LCPC_B200_ERR_UNSUPPORTED (-7) /* valid request this build does not implement */
#define LCPC_B200_ERR_OUTER_TENSOR (-8) /* ProverError::OuterTensor (:1016-1018) */
/* VerifierError (lcpc-2d/src/lib.rs:141-170); VerifierError::Encode is LCPC_B200_ERR_ENCODE */
#define LCPC_B200_VERR_NUM_COL_OPENS (-20) /* :845-847 */
#define LCPC_B200_VERR_COLUMN_PATH (-21)   /* :940 */
#define LCPC_B200_VERR_COLUMN_EVAL (-22)   /* :939 */
#define LCPC_B200_VERR_COLUMN_DEGREE (-23) /* :938 */
#define LCPC_B200_VERR_OUTER_TENSOR (-24)  /* :854-856 */
#define LCPC_B200_VERR_INNER_TENSOR (-25)  /* :851-853 */
#define LCPC_B200_VERR_ENCODING_DIMS (-26) /* :857-859 */

/* field ids: lcpc-test-fields/src/lib.rs:13-59 */
enum { LCPC_B200_FT63 = 1, LCPC_B200_FT127 = 2, LCPC_B200_FT191 = 3, LCPC_B200_FT255 = 4 };
/* encoding kinds */
enum { LCPC_B200_ENC_LIGERO = 1, LCPC_B200_ENC_SDIG = 2 };

typedef struct lcpc_b200_ctx lcpc_b200_ctx;       /* one CUDA device + stream + scratch */
typedef struct lcpc_b200_enc lcpc_b200_enc;       /* device side of an `impl LcEncoding` */
typedef struct lcpc_b200_commit lcpc_b200_commit; /* device-resident LcCommit (lcpc-2d/src/lib.rs:172-184) */

/* ---- library / context ---- */
const char *lcpc_b200_version(void);
/* Schedule knobs for A/B measurements (never change results): `name` without the LCPC_B200_ prefix of the
 * equivalent environment variable, e.g. "SPMM_WINDOW_KB".  A set value wins over the environment. */
int lcpc_b200_set_tunable(const char *name, long value);
long lcpc_b200_get_tunable(const char *name, long dflt);
/* number of u64 limbs of a field (1,2,3,4) or -1 */
int lcpc_b200_field_limbs(int field);
/* Field::one() as stored: R mod p, L u64 limbs (ff_derive's R constant); host-only, no device needed */
int lcpc_b200_field_one(int field, uint64_t *out);
int lcpc_b200_ctx_create(int device, lcpc_b200_ctx **out);
void lcpc_b200_ctx_destroy(lcpc_b200_ctx *ctx);
const char *lcpc_b200_last_error(const lcpc_b200_ctx *ctx);
int lcpc_b200_ctx_device(const lcpc_b200_ctx *ctx);
/* the cudaStream_t all work of this context is enqueued on (for event timing by the caller) */
void *lcpc_b200_ctx_stream(const lcpc_b200_ctx *ctx);
int lcpc_b200_ctx_synchronize(lcpc_b200_ctx *ctx);
/* kernels launched by this context since creation (bench.py's `gpu_launches`) */
uint64_t lcpc_b200_ctx_launch_count(const lcpc_b200_ctx *ctx);

/* page-locked host buffers (no reference analogue: the reference never leaves host memory).  commit()
 * accepts any host pointer; from a buffer obtained here (or registered in place, e.g. a Rust Vec<F>'s
 * allocation) the coefficient rows cross PCIe at full rate while earlier rows are already being encoded. */
int lcpc_b200_host_alloc(size_t bytes, void **out);
void lcpc_b200_host_free(void *p);
int lcpc_b200_host_register(void *p, size_t bytes);
int lcpc_b200_host_unregister(void *p);

/* ---- encodings (impl LcEncoding, lcpc-2d/src/lib.rs:74-104) ----
 * Dimension choosing (`_get_dims`, lcpc-ligero-pc/src/lib.rs:70-112; `_new_from_np1`,
 * lcpc-brakedown-pc/src/lib.rs:69-99) and Brakedown code generation (matgen.rs) stay on the host
 * side of the boundary; the device object is built from their results. */

/* LigeroEncodingRho::new_from_dims (lcpc-ligero-pc/src/lib.rs:138-148): checks dims_ok (:114-118),
 * builds the root table precomp_fft(n_cols) would (:140).  ERR_BAD_ARG when !dims_ok, ERR_TOO_BIG when
 * log2(n_cols) exceeds the field's two-adicity. */
int lcpc_b200_ligero_new(lcpc_b200_ctx *ctx, int field, size_t n_per_row, size_t n_cols, lcpc_b200_enc **out);

/* One Brakedown code matrix exactly as matgen::gen_code builds it (matgen.rs:114-188, :187):
 * CsMat::new_csc((m, n), ptrs, idxs, data) -- m outputs, n inputs, column-compressed. */
typedef struct {
  size_t m, n;
  const uint64_t *ptrs; /* n + 1 column starts */
  const uint64_t *idxs; /* nnz row indices */
  const uint64_t *data; /* nnz elements, Montgomery limbs */
} lcpc_b200_csc;
/* SdigEncodingS from its `precodes` / `postcodes` vectors (lcpc-brakedown-pc/src/lib.rs:40-47);
 * n_per_row = pre[0].n, n_cols = codeword_length (encode.rs:18-33). */
int lcpc_b200_sdig_new(lcpc_b200_ctx *ctx, int field, size_t n_levels, const lcpc_b200_csc *pre,
                       const lcpc_b200_csc *post, lcpc_b200_enc **out);
void lcpc_b200_enc_free(lcpc_b200_enc *enc);
int lcpc_b200_enc_kind(const lcpc_b200_enc *enc);
int lcpc_b200_enc_field(const lcpc_b200_enc *enc);
/* LcEncoding::get_dims (lcpc-ligero-pc/src/lib.rs:166-169, lcpc-brakedown-pc/src/lib.rs:155-158) */
int lcpc_b200_enc_get_dims(const lcpc_b200_enc *enc, size_t len, size_t *n_rows, size_t *n_per_row,
                           size_t *n_cols);
/* LcEncoding::dims_ok (lcpc-ligero-pc/src/lib.rs:171-177, lcpc-brakedown-pc/src/lib.rs:160-167): 1 / 0 */
int lcpc_b200_enc_dims_ok(const lcpc_b200_enc *enc, size_t n_per_row, size_t n_cols);

/* LcEncoding::encode (trait :91; lcpc-ligero-pc/src/lib.rs:162-164, lcpc-brakedown-pc/src/lib.rs:150-153),
 * batched: `rows` holds n_rows rows of n_cols elements (host memory), each encoded in place; like the
 * reference, the whole row is the input (callers zero the tail, lcpc-2d/src/lib.rs:648-653, :886, :918). */
int lcpc_b200_encode(lcpc_b200_enc *enc, uint64_t *rows, size_t n_rows);
/* same on DEVICE memory; `valid` leading elements of each row are read, the rest taken as zero */
int lcpc_b200_encode_dev(lcpc_b200_enc *enc, uint64_t *d_rows, size_t n_rows, size_t valid);

/* out of place on DEVICE memory: source rows are src_stride elements apart with `valid` leading
 * elements each; destination rows are n_cols apart (the row-block step of the multi-GPU commit) */
int lcpc_b200_encode_rows_dev(lcpc_b200_enc *enc, const uint64_t *d_src, size_t src_stride, size_t valid,
                              uint64_t *d_dst, size_t n_rows);

/* the same fed from HOST memory: `len` coefficients (the rank's row block; the last row may be short and is
 * zero-padded) are copied into d_coeffs[n_rows][n_per_row] in row-chunks on a copy stream while the engine
 * stream encodes the chunks that have landed into d_dst[n_rows][n_cols].  Enqueues only. */
int lcpc_b200_encode_rows_h2d(lcpc_b200_enc *enc, const uint64_t *rows, size_t len, uint64_t *d_coeffs,
                              uint64_t *d_dst, size_t n_rows);

/* Row-block encode whose result is stored per COLUMN BLOCK instead of row-major: the fused "encode +
 * transpose" step of the multi-GPU commit (no reference analogue).  Column block h = columns
 * [starts[h], starts[h+1]) of every encoded row lands in the matrix dst[h][all rows][width_h]; dst[h] is
 * device memory of this GPU or a peer-mapped buffer of another GPU (written over NVLink by the encode's last
 * pass itself).  This call's rows are rows row0 .. row0+n_rows-1 of those matrices.  d_tmp: n_rows*n_cols
 * elements of row-major scratch for the passes before the last.  Enqueues only. */
typedef struct {
  size_t n_blocks;        /* 1..16 */
  const uint64_t *starts; /* n_blocks + 1 column boundaries, starts[0] = 0, starts[n_blocks] = n_cols (host) */
  uint64_t *const *dst;   /* n_blocks device pointers (host array) */
  size_t row0;
} lcpc_b200_scatter;
int lcpc_b200_encode_rows_scatter_dev(lcpc_b200_enc *enc, const uint64_t *d_src, size_t src_stride, size_t valid,
                                      uint64_t *d_tmp, size_t n_rows, const lcpc_b200_scatter *scatter);
/* the same fed from host rows (see lcpc_b200_encode_rows_h2d) */
int lcpc_b200_encode_rows_scatter_h2d(lcpc_b200_enc *enc, const uint64_t *rows, size_t len, uint64_t *d_coeffs,
                                      uint64_t *d_tmp, size_t n_rows, const lcpc_b200_scatter *scatter);

/* ---- commit (LcCommit::commit, lcpc-2d/src/lib.rs:299-301 -> :622-671) ----
 * coeffs_in: `len` elements on the host.  Pads to n_rows x n_per_row (:636-645), encodes every row
 * (:648-653), hashes columns and builds the Merkle tree (:656-668).  The result stays on the device. */
int lcpc_b200_commit_new(lcpc_b200_enc *enc, const uint64_t *coeffs_in, size_t len, lcpc_b200_commit **out);
/* same with coeffs_in already in device memory (the roofline-timed region of bench.py) */
int lcpc_b200_commit_new_dev(lcpc_b200_enc *enc, const uint64_t *d_coeffs_in, size_t len, lcpc_b200_commit **out);
/* re-run the commit into an existing object of the same shape (no allocation; for timing loops) */
int lcpc_b200_commit_rerun_dev(lcpc_b200_commit *c, const uint64_t *d_coeffs_in, size_t len);
int lcpc_b200_commit_rerun(lcpc_b200_commit *c, const uint64_t *coeffs_in, size_t len);
/* Deserialize for LcCommit (:256-268): a device-resident commit from the fields of one made elsewhere (host arrays:
 * comm_len = n_rows*n_cols elements, coeffs_len = n_rows*n_per_row, n_hashes = 2*np2-1 digests).  Nothing is recomputed;
 * sizes are checked like check_comm (:672-688), ERR_BAD_ARG standing for ProverError::Commit. */
int lcpc_b200_commit_from_host(lcpc_b200_enc *enc, const uint64_t *comm, size_t comm_len, const uint64_t *coeffs,
                               size_t coeffs_len, const uint8_t *hashes, size_t n_hashes, size_t n_rows,
                               lcpc_b200_commit **out);
void lcpc_b200_commit_free(lcpc_b200_commit *c);
/* n_hashes = 2 * next_power_of_two(n_cols) - 1 (:656-666) */
int lcpc_b200_commit_dims(const lcpc_b200_commit *c, size_t *n_rows, size_t *n_per_row, size_t *n_cols,
                          size_t *n_hashes);
/* LcCommit::get_root (:276-281): the last entry of `hashes` */
int lcpc_b200_commit_root(lcpc_b200_commit *c, uint8_t root[32]);
/* copy out the LcCommit fields (:178-183); any pointer may be NULL to skip that field.
 * comm: n_rows*n_cols elements, coeffs: n_rows*n_per_row elements, hashes: n_hashes*32 bytes. */
int lcpc_b200_commit_download(lcpc_b200_commit *c, uint64_t *comm, uint64_t *coeffs, uint8_t *hashes);
/* device time of the phases of the last (re)run on this object, from events recorded on the context's
 * stream: ms[0] pad/copy, ms[1] row encode, ms[2] column leaf hashing, ms[3] Merkle layers; launches[]
 * (optional) = kernels launched by the encode / leaf-hash / Merkle phases.  Synchronises. */
int lcpc_b200_commit_phase_times(lcpc_b200_commit *c, float ms[4], int launches[3]);
/* device pointers of the same three arrays (owned by the commit).  A Brakedown commit made from device memory keeps
 * its codewords and coefficient rows column-major in the encoder's work buffer (columns are hashed, opened and combined
 * from there); the row-major `comm` / `coeffs` the reference's LcCommit holds are written when they are first asked
 * for, here or in lcpc_b200_commit_download: ask only for what you need (pass NULL for the rest).  The pointers stay
 * valid until the commit is freed, their CONTENTS until the next rerun on this object. */
int lcpc_b200_commit_device_ptrs(lcpc_b200_commit *c, uint64_t **d_comm, uint64_t **d_coeffs, uint8_t **d_hashes);
/* commit into an existing object AND copy the LcCommit fields out (any of comm / coeffs / hashes may be NULL):
 * row-chunks of comm and coeffs travel back on a second copy stream while later chunks are still being
 * uploaded and encoded (PCIe is full duplex), so the eager, host-visible commit() costs about max(upload,
 * download) instead of their sum.  Page-locked buffers (lcpc_b200_host_alloc / _register) are needed for the
 * overlap; pageable ones work at the driver's staging rate. */
int lcpc_b200_commit_rerun_to_host(lcpc_b200_commit *c, const uint64_t *coeffs_in, size_t len, uint64_t *comm,
                                   uint64_t *coeffs, uint8_t *hashes);
/* one-shot form with host outputs, the exact shape of commit(): allocate + the call above + free */
int lcpc_b200_commit_to_host(lcpc_b200_enc *enc, const uint64_t *coeffs_in, size_t len, uint64_t *comm,
                             uint64_t *coeffs, uint8_t *hashes);

/* ---- prove pieces ---- */
/* collapse_columns (lcpc-2d/src/lib.rs:1095-1123; call sites :1034, :1055):
 * poly[c] = sum_r tensor[r] * coeffs[r*n_per_row + c]; tensor: n_rows elements, poly: n_per_row (host) */
int lcpc_b200_commit_collapse(lcpc_b200_commit *c, const uint64_t *tensor, uint64_t *poly);
/* One degree test of prove() (lcpc-2d/src/lib.rs:1026-1041) without the tensor crossing PCIe: `key` is the 32
 * bytes the transcript yields for LABEL_DT (:1026-1027); the device expands ChaCha20Rng::from_seed(key) into
 * n_rows x F::random (:1028-1032, the step the reference marks "could expand seed in parallel") and collapses
 * the coefficient matrix with it (:1033-1041).  poly: n_per_row elements; tensor_out (optional): the n_rows
 * tensor elements, for callers that want to cross-check. */
int lcpc_b200_commit_degree_test(lcpc_b200_commit *c, const uint8_t key[32], uint64_t *poly, uint64_t *tensor_out);
/* the expansion alone: out[0..n) = n x F::random from ChaCha20Rng::from_seed(key) (also what verify() draws, :868-877) */
int lcpc_b200_expand_tensor(lcpc_b200_ctx *ctx, int field, const uint8_t key[32], size_t n, uint64_t *out);
/* stateless form on host arrays */
int lcpc_b200_collapse(lcpc_b200_ctx *ctx, int field, const uint64_t *coeffs, const uint64_t *tensor,
                       uint64_t *poly, size_t n_rows, size_t n_per_row);
/* open_column (lcpc-2d/src/lib.rs:788-825) for `n` columns at once (the par_iter at :1081-1084):
 * cols_out: n * n_rows elements (column i contiguous), paths_out: n * path_len * 32 bytes with
 * path_len = log2(next_power_of_two(n_cols)).  ERR_COLUMN if any index >= n_cols (:797-799). */
int lcpc_b200_commit_open_columns(lcpc_b200_commit *c, const uint64_t *cols, size_t n, uint64_t *cols_out,
                                  uint8_t *paths_out);

/* ---- whole prove() / verify() (lcpc-2d/src/lib.rs:1004-1093, :832-952) ----
 * The Fiat-Shamir transcript (merlin::Transcript) is sequential host work and lives on the host side of the
 * boundary (include/lcpc_b200_host.h: lcpc_b200_transcript_*); a Rust host passes challenges itself through the
 * piecewise entry points above, a host without merlin hands a transcript handle to the two calls below. */
typedef struct lcpc_b200_transcript lcpc_b200_transcript;
/* domain-separation labels (LcEncoding::LABEL_DT/PR/PE/CO, :78-85).  NULL selects what def_labels! actually
 * produces for every encoding of the reference, the literal byte strings "$l//DT", "$l//PR", "$l//PE", "$l//CO"
 * (lcpc-2d/src/macros.rs:28-36: `$l` inside a byte-string literal is not substituted). */
typedef struct {
  const uint8_t *dt, *pr, *pe, *co;
  size_t dt_len, pr_len, pe_len, co_len;
} lcpc_b200_labels;
/* prove (:1004-1093): n_degree_tests x { key <- tr.challenge(LABEL_DT); tensor <- ChaCha20(key) on the device;
 * p_random[i] = collapse(coeffs, tensor); tr <- p_random[i] }, p_eval = collapse(coeffs, outer_tensor), tr <- p_eval,
 * key <- tr.challenge(LABEL_CO), n_col_opens column numbers <- Uniform(0, n_cols) over ChaCha20(key), open_column each.
 * Outputs (host): p_eval n_per_row elements; p_random n_degree_tests x n_per_row; col_idx (optional) the opened column
 * numbers; cols_out n_col_opens x n_rows elements; paths_out n_col_opens x path_len x 32 bytes.
 * ERR_OUTER_TENSOR when outer_len != n_rows. */
int lcpc_b200_commit_prove(lcpc_b200_commit *c, lcpc_b200_transcript *tr, const lcpc_b200_labels *labels,
                           const uint64_t *outer_tensor, size_t outer_len, size_t n_degree_tests, size_t n_col_opens,
                           uint64_t *p_eval, uint64_t *p_random, uint64_t *col_idx, uint64_t *cols_out,
                           uint8_t *paths_out);
/* the fields of an LcEvalProof (:490-500) as flat host arrays */
typedef struct {
  size_t n_cols;            /* LcEvalProof::n_cols */
  size_t n_per_row;         /* p_eval.len() */
  size_t n_degree_tests;    /* p_random_vec.len() */
  size_t n_columns;         /* columns.len() */
  size_t n_rows;            /* columns[0].col.len() */
  size_t path_len;          /* columns[i].path.len() */
  const uint64_t *p_eval;   /* n_per_row elements */
  const uint64_t *p_random; /* n_degree_tests x n_per_row */
  const uint64_t *cols;     /* n_columns x n_rows (each LcColumn::col contiguous) */
  const uint8_t *paths;     /* n_columns x path_len x 32 */
} lcpc_b200_proof;
/* verify (:832-952): argument checks (:845-859), challenges re-derived from `tr` (:866-911), p_random / p_eval
 * rows encoded on the device (:883-888, :914-921), every opened column checked in one batch (:926-942: dot
 * products against all tensors, leaf digest, Merkle path), then <inner_tensor, p_eval> (:944-951) -> eval_out.
 * Returns LCPC_B200_OK, an LCPC_B200_VERR_* code, or LCPC_B200_ERR_*. */
int lcpc_b200_verify(lcpc_b200_enc *enc, lcpc_b200_transcript *tr, const lcpc_b200_labels *labels,
                     const uint8_t root[32], const uint64_t *outer_tensor, size_t outer_len,
                     const uint64_t *inner_tensor, size_t inner_len, size_t n_col_opens, size_t n_degree_tests,
                     const lcpc_b200_proof *proof, uint64_t *eval_out);

/* ---- commit / prove sharded over the GPUs of one box (no reference analogue: the reference is one process on CPU
 * threads; what is replaced is still LcCommit::commit, lcpc-2d/src/lib.rs:622-671, and ::prove, :1004-1093) ----
 * GPU g encodes the row block [row_lo[g], row_lo[g+1]) and hashes the column block [col_lo[g], col_lo[g+1]); the
 * transpose between the two is fused into the encode, whose last pass stores every tile straight into the owner's
 * memory over NVLink (peer-mapped windows), ordered by flags in those windows -- no collective library on the path.
 *
 * Two shapes:
 *  (1) ONE PROCESS, several GPUs (a Rust host; the reference's own shape): lcpc_b200_commit_new_multi and the
 *      lcpc_b200_multi_* calls below -- one encoding per GPU (each on its own context), everything else inside.
 *  (2) one process per GPU (bench.py under torchrun): every rank creates its lcpc_b200_shard, publishes the 64-byte
 *      CUDA IPC handle of its window by whatever channel the host has, connects, and calls the same operations;
 *      ranks stay in lock-step through the windows' flags.  A rank that never arrives makes the others fail with
 *      LCPC_B200_ERR_CUDA after a timeout (tunable SHARD_TIMEOUT_MS) instead of hanging the device. */
typedef struct lcpc_b200_shard lcpc_b200_shard; /* one GPU's part of a sharded LcCommit */
typedef struct lcpc_b200_multi lcpc_b200_multi; /* a whole sharded LcCommit driven from one process */

/* the partition (pure host arithmetic, the same on every rank): world + 1 boundaries each; column blocks are unions
 * of aligned Merkle subtrees of *sub_leaves leaves, dealt over the subtrees that contain real columns */
int lcpc_b200_shard_plan(size_t n_rows, size_t n_per_row, size_t n_cols, unsigned world, size_t *row_lo, size_t *col_lo,
                         size_t *sub_lo, size_t *sub_leaves, size_t *n_sub);
/* rank `rank` of `world` for a commit of `len` coefficients under `enc` (this GPU's instance of the encoding);
 * max_open = the most columns one prove() will open (LcEncoding::get_n_col_opens) */
int lcpc_b200_shard_new(lcpc_b200_enc *enc, size_t len, unsigned world, unsigned rank, size_t max_open,
                        lcpc_b200_shard **out);
void lcpc_b200_shard_free(lcpc_b200_shard *s);
/* this rank's window: device pointer (same-process peers) and/or its CUDA IPC handle (other processes) */
int lcpc_b200_shard_window(lcpc_b200_shard *s, void **d_ptr, size_t *bytes, uint8_t ipc_handle[64]);
/* map every peer's window: peer_ptrs[h] (same process; peer access is enabled here) or, where that is NULL,
 * ipc_handles + 64*h (another process).  Entry `rank` is ignored. */
int lcpc_b200_shard_connect(lcpc_b200_shard *s, void *const *peer_ptrs, const uint8_t *ipc_handles);
int lcpc_b200_shard_dims(const lcpc_b200_shard *s, size_t *n_rows, size_t *n_per_row, size_t *n_cols, size_t *row_lo,
                         size_t *row_hi, size_t *col_lo, size_t *col_hi, size_t *n_elems);
/* enqueue one commit: `rows` = this rank's n_elems coefficients (row block, the polynomial's last row may be
 * short) in HOST memory, copied in row-chunks under the encode; _dev: in device memory, or NULL to commit the rows
 * lcpc_b200_shard_load_rows stored.  No host synchronisation. */
int lcpc_b200_shard_commit(lcpc_b200_shard *s, const uint64_t *rows, size_t n_elems);
int lcpc_b200_shard_commit_dev(lcpc_b200_shard *s, const uint64_t *d_rows, size_t n_elems);
int lcpc_b200_shard_load_rows(lcpc_b200_shard *s, const uint64_t *rows, size_t n_elems);
/* one commit of the loaded rows in its three enqueue steps (1: encode + peer stores + "tiles landed"; 2: wait, hash,
 * subtree roots to all, "roots landed"; 3: wait, top tree).  A single process that drives several shards itself must
 * run step k on EVERY shard before step k+1 on any (lcpc_b200_multi_rerun does exactly that), so that no wait is
 * enqueued in front of a signal it depends on. */
int lcpc_b200_shard_commit_step(lcpc_b200_shard *s, int step);
/* LcCommit::get_root (:276-281): synchronises this rank's stream; every rank returns the same digest */
int lcpc_b200_shard_root(lcpc_b200_shard *s, uint8_t root[32]);
/* the same without the synchronisation: the 32 bytes land in `root` (page-locked host memory) when the stream
 * gets there; a pipelined caller enqueues the next commit right behind it and reads the roots later */
int lcpc_b200_shard_root_enqueue(lcpc_b200_shard *s, uint8_t *root);
/* Tunable SHARD_PIPELINE=1 (read at lcpc_b200_shard_new) runs a commit's exchange wait, hashing and tree on a second
 * stream so that the next commit's encode follows this one's directly.  lcpc_b200_shard_join makes the context's
 * stream (lcpc_b200_ctx_stream, what callers time and synchronise) wait for that second stream; every call that
 * reads a commit's results joins by itself.  Enqueue only. */
int lcpc_b200_shard_join(lcpc_b200_shard *s);
/* device times of the last commit on this rank: ms[0] encode + peer stores, ms[1] wait for the peers' tiles,
 * ms[2] column hashing, subtree roots, root exchange and top tree */
int lcpc_b200_shard_phase_times(lcpc_b200_shard *s, float ms[3]);
/* this rank's pieces of the commit (tests): column block [n_rows][my_cols], row block of coeffs, leaf digests, top tree */
int lcpc_b200_shard_device_ptrs(lcpc_b200_shard *s, uint64_t **d_recv, uint64_t **d_coeffs, uint8_t **d_leaves, uint8_t **d_top);
/* copy this rank's column block of comm ([n_rows][col_hi - col_lo] elements) and its leaf digests to the host */
int lcpc_b200_shard_download(lcpc_b200_shard *s, uint64_t *cols_out, uint8_t *leaves_out);
/* collapse_columns (:1095-1123) over row-sharded coefficients, split in two so that a caller can overlap host work:
 * begin = this rank's partial combination (tensor: n_rows elements on the host, or key: the 32 challenge bytes the
 * device expands, :1026-1032) stored into every peer's window; finish = wait for all partials, sum them on the
 * device, poly (n_per_row elements) and optionally their canonical bytes (to_repr, for the transcript) to the host. */
int lcpc_b200_shard_collapse_begin(lcpc_b200_shard *s, const uint64_t *tensor, const uint8_t key[32]);
int lcpc_b200_shard_collapse_finish(lcpc_b200_shard *s, uint64_t *poly, uint8_t *repr);
/* open_column (:788-825) for n columns: the owner of a column supplies values and path; all ranks receive all */
int lcpc_b200_shard_open_begin(lcpc_b200_shard *s, const uint64_t *cols, size_t n);
int lcpc_b200_shard_open_finish(lcpc_b200_shard *s, uint64_t *cols_out, uint8_t *paths_out);
/* LcCommit::prove (:1004-1093); arguments as lcpc_b200_commit_prove.  Shape (2): every rank calls it with an
 * identical transcript and receives the same proof. */
int lcpc_b200_shard_prove(lcpc_b200_shard *s, lcpc_b200_transcript *tr, const lcpc_b200_labels *labels,
                          const uint64_t *outer_tensor, size_t outer_len, size_t n_degree_tests, size_t n_col_opens,
                          uint64_t *p_eval, uint64_t *p_random, uint64_t *col_idx, uint64_t *cols_out, uint8_t *paths_out);

/* shape (1): LcCommit::commit over encs[0..n_gpus) (the same encoding built on n_gpus contexts); coeffs_in: `len`
 * elements on the host (page-locked memory gives full-rate copies on every GPU's own PCIe link) */
int lcpc_b200_commit_new_multi(lcpc_b200_enc *const *encs, size_t n_gpus, const uint64_t *coeffs_in, size_t len,
                               size_t max_open, lcpc_b200_multi **out);
int lcpc_b200_multi_rerun(lcpc_b200_multi *m, const uint64_t *coeffs_in, size_t len); /* enqueue only */
int lcpc_b200_multi_root(lcpc_b200_multi *m, uint8_t root[32]);
int lcpc_b200_multi_collapse(lcpc_b200_multi *m, const uint64_t *tensor, const uint8_t key[32], uint64_t *poly, uint8_t *repr);
int lcpc_b200_multi_open_columns(lcpc_b200_multi *m, const uint64_t *cols, size_t n, uint64_t *cols_out, uint8_t *paths_out);
int lcpc_b200_multi_prove(lcpc_b200_multi *m, lcpc_b200_transcript *tr, const lcpc_b200_labels *labels,
                          const uint64_t *outer_tensor, size_t outer_len, size_t n_degree_tests, size_t n_col_opens,
                          uint64_t *p_eval, uint64_t *p_random, uint64_t *col_idx, uint64_t *cols_out, uint8_t *paths_out);
size_t lcpc_b200_multi_n_shards(const lcpc_b200_multi *m);
lcpc_b200_shard *lcpc_b200_multi_shard(lcpc_b200_multi *m, size_t g);
void lcpc_b200_multi_free(lcpc_b200_multi *m);

/* ---- standalone pieces (tests, verifier-side use, multi-GPU pipeline) ---- */
/* merkleize (lcpc-2d/src/lib.rs:690-704) of a host row-major comm into host hashes[2*np2-1][32] */
int lcpc_b200_merkleize(lcpc_b200_ctx *ctx, int field, const uint64_t *comm, size_t n_rows, size_t n_cols,
                        uint8_t *hashes);
/* device building blocks: column leaf digests of a device matrix whose element (r,c) is at
 * d_comm[(r*row_stride + c)*L]; d_leaves: n_cols*32 bytes */
int lcpc_b200_hash_columns_dev(lcpc_b200_ctx *ctx, int field, const uint64_t *d_comm, size_t n_rows,
                               size_t n_cols, size_t row_stride, uint8_t *d_leaves);
/* merkle_tree (lcpc-2d/src/lib.rs:747-760) in place on device: d_hashes[0..np2) given */
int lcpc_b200_merkle_tree_dev(lcpc_b200_ctx *ctx, uint8_t *d_hashes, size_t np2);
/* n_layers Merkle layers over n_leaves nodes laid out [nodes | layer 1 | .. | layer n_layers]; n_leaves a
 * multiple of 2^n_layers (several equal aligned subtrees side by side are reduced together) */
int lcpc_b200_merkle_layers_dev(lcpc_b200_ctx *ctx, uint8_t *d_hashes, size_t n_leaves, unsigned n_layers);
/* multi-GPU transpose step (no reference analogue; the reference is one process): split the device
 * row-block d_rows[n_rows][n_cols] into n_blocks column-block tiles; tile h = columns
 * [starts[h], starts[h+1]) stored as [n_rows][width_h] at element offset n_rows*starts[h] of d_out,
 * i.e. the send buffer of one all-to-all.  d_starts: n_blocks+1 u64 in DEVICE memory. */
int lcpc_b200_pack_column_blocks_dev(lcpc_b200_ctx *ctx, int field, const uint64_t *d_rows, size_t n_rows,
                                     size_t n_cols, size_t n_blocks, const uint64_t *d_starts, uint64_t *d_out);
int lcpc_b200_collapse_dev(lcpc_b200_ctx *ctx, int field, const uint64_t *d_coeffs, size_t row_stride,
                           const uint64_t *d_tensor, uint64_t *d_poly, size_t n_rows, size_t n_per_row);
/* element-wise field arithmetic on host arrays (parity tests of the device arithmetic):
 * op 0 add, 1 sub, 2 mul, 4 from_mont (b ignored), 5 mul as full product + separate reduction,
 * 6 r[i] = sum_{k<37} a[(i+k)%n] * b[(7i+k)%n] accumulated double-width and reduced once,
 * 7 mul as a one-level Karatsuba product + reduction (fields with 2 or 4 u64 limbs; plain mul otherwise) */
int lcpc_b200_field_op(lcpc_b200_ctx *ctx, int field, int op, uint64_t *r, const uint64_t *a,
                       const uint64_t *b, size_t n);

#ifdef __cplusplus
}
#endif
#endif /* LCPC_B200_H */
