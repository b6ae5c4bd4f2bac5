/*
 * oracle/lcpc_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C + OpenMP) of the reference's commit/prove hot path, used ONLY as the
 * parity checker in tests/, in __graft_entry__.smoke(), and as the timed CPU baseline in bench.py
 * (`cpu_baseline` leg and `--impl reference`). The product (lcpc_b200/) never imports, links or
 * executes anything under oracle/.
 *
 * PARITY STATUS (see DESIGN.md "Oracle"): the real reference is Rust and cannot be built in this
 * environment (no cargo/rustc; the crates holding the arithmetic -- ff_derive 0.12, fffft 0.4,
 * sprs 0.10, blake3 1, rand_chacha 0.3 -- are not vendored), and its tests hold no known-answer
 * vectors (all inputs come from thread_rng). Hence:
 *   - BLAKE3 / Merkle wiring: pinned against the Python `blake3` binding of the same Rust crate.
 *   - Brakedown encode recursion + flat layout: pinned against the reference's own Python
 *     spec /root/reference/doc/encoding.py (fixtures in tests/golden/, generator committed).
 *   - field arithmetic, NTT ordering/root choice, matgen RNG streams: PARITY UNPINNED against the
 *     reference; pinned only against the published algorithms (big-int Python, RFC 8439).
 *
 * Elements are NL little-endian u64 limbs in MONTGOMERY form (the in-memory form of the
 * reference's `struct FtNNN([u64; L])`, lcpc-test-fields/src/lib.rs:22,34,46,58).
 */
#ifndef LCPC_ORACLE_H
#define LCPC_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { LCPC_FT63 = 1, LCPC_FT127 = 2, LCPC_FT191 = 3, LCPC_FT255 = 4 };
enum { LCPC_ENC_LIGERO = 1, LCPC_ENC_SDIG = 2 };

/* ---- fields (lcpc-test-fields/src/lib.rs:13-59) ---- */
int lcpc_oracle_field_limbs(int field);
/* out arrays hold NL limbs each; any may be NULL */
int lcpc_oracle_field_info(int field, uint32_t *num_bits, uint32_t *s, uint64_t *modulus,
                           uint64_t *r, uint64_t *r2, uint64_t *inv, uint64_t *rou);
/* op: 0 add, 1 sub, 2 mul, 3 to_mont (b ignored), 4 from_mont (b ignored), 5 inverse (b ignored) */
int lcpc_oracle_field_op(int field, int op, uint64_t *r, const uint64_t *a, const uint64_t *b,
                         size_t n);
int lcpc_oracle_to_repr(int field, uint8_t *out, const uint64_t *a, size_t n);
/* n uniform elements the way `Field::random` draws them from ChaCha20Rng::seed_from_u64(seed)
 * with set_stream(stream) */
int lcpc_oracle_random_elems(int field, uint64_t seed, uint64_t stream, uint64_t *out, size_t n);
/* n elements from ChaCha20Rng::from_seed(key): the challenge tensors of prove()/verify()
 * (lcpc-2d/src/lib.rs:1026-1032, 868-877) */
int lcpc_oracle_random_elems_from_key(int field, const uint8_t key[32], uint64_t *out, size_t n);

/* ---- hash ---- */
void lcpc_oracle_blake3(const uint8_t *in, size_t len, uint8_t out[32]);
void lcpc_oracle_chacha_block(const uint32_t key[8], uint64_t counter, uint64_t stream,
                              uint32_t out[16]);

/* ---- NTT (fffft fft_io / ifft_oi), in place, len a power of two ---- */
int lcpc_oracle_fft_io(int field, uint64_t *x, size_t len);
int lcpc_oracle_ifft_oi(int field, uint64_t *x, size_t len);
/* w = root_of_unity()^(2^(S - log2 len)), Montgomery form */
int lcpc_oracle_root_of_unity(int field, size_t len, uint64_t *w);

/* ---- n_degree_tests (lcpc-2d/src/lib.rs:613-616) and per-code parameters ---- */
size_t lcpc_oracle_n_degree_tests(size_t lambda, size_t len, size_t flog2);
size_t lcpc_oracle_ligero_n_col_opens(size_t rho_num, size_t rho_den);
/* LigeroEncodingRho::_get_dims (lcpc-ligero-pc/src/lib.rs:70-112); returns 0 on success */
int lcpc_oracle_ligero_get_dims(int field, size_t len, size_t rho_num, size_t rho_den,
                                size_t *n_rows, size_t *n_per_row, size_t *n_cols);
size_t lcpc_oracle_sdig_n_col_opens(int code);
/* matgen::get_dims (matgen.rs:56-111): dims[i] = {n, m, d}; returns number of levels */
int lcpc_oracle_sdig_level_dims(int field, int code, size_t n, size_t max_levels,
                                size_t (*pre_dims)[3], size_t (*post_dims)[3]);

/* ---- encodings (impl LcEncoding) ---- */
typedef struct lcpc_oracle_enc lcpc_oracle_enc;
lcpc_oracle_enc *lcpc_oracle_ligero_new_from_dims(int field, size_t n_per_row, size_t n_cols,
                                                  size_t rho_num, size_t rho_den);
lcpc_oracle_enc *lcpc_oracle_ligero_new(int field, size_t len, size_t rho_num, size_t rho_den);
/* SdigEncodingS::new (lcpc-brakedown-pc/src/lib.rs:103-110) and ::new_from_dims (:126-137);
 * code = 1..6 selects SdigCode1..6 (codespec.rs:169-232) */
lcpc_oracle_enc *lcpc_oracle_sdig_new(int field, int code, size_t len, uint64_t seed);
lcpc_oracle_enc *lcpc_oracle_sdig_new_from_dims(int field, int code, size_t n_per_row, size_t n_cols,
                                                uint64_t seed);
/* an SdigEncoding around caller-supplied CSC matrices (copied) */
lcpc_oracle_enc *lcpc_oracle_sdig_from_matrices(int field, int code, size_t n_levels,
                                                const size_t *pre_m, const size_t *pre_n,
                                                const uint64_t *const *pre_ptrs,
                                                const uint64_t *const *pre_idxs,
                                                const uint64_t *const *pre_data,
                                                const size_t *post_m, const size_t *post_n,
                                                const uint64_t *const *post_ptrs,
                                                const uint64_t *const *post_idxs,
                                                const uint64_t *const *post_data);
void lcpc_oracle_enc_free(lcpc_oracle_enc *e);
int lcpc_oracle_enc_field(const lcpc_oracle_enc *e);
int lcpc_oracle_enc_kind(const lcpc_oracle_enc *e);
void lcpc_oracle_enc_get_dims(const lcpc_oracle_enc *e, size_t len, size_t *n_rows,
                              size_t *n_per_row, size_t *n_cols);
int lcpc_oracle_enc_dims_ok(const lcpc_oracle_enc *e, size_t n_per_row, size_t n_cols);
size_t lcpc_oracle_enc_n_col_opens(const lcpc_oracle_enc *e);
size_t lcpc_oracle_enc_n_degree_tests(const lcpc_oracle_enc *e);
size_t lcpc_oracle_sdig_n_levels(const lcpc_oracle_enc *e);
/* borrow one code matrix (CSC: m rows, n cols); pointers stay valid until enc_free */
int lcpc_oracle_sdig_matrix(const lcpc_oracle_enc *e, size_t level, int is_post, size_t *m,
                            size_t *n, size_t *nnz, const uint64_t **ptrs, const uint64_t **idxs,
                            const uint64_t **data);
/* LcEncoding::encode: one n_cols-long row, in place */
int lcpc_oracle_encode(const lcpc_oracle_enc *e, uint64_t *row);

/* ---- commit path (lcpc-2d/src/lib.rs:622-785) ----
 * comm_out: n_rows*n_cols elements; coeffs_out: n_rows*n_per_row; hashes_out: 32*(2*np2-1) bytes.
 * threads <= 0 means "all". Returns 0, or <0 (-1 TooBig, -2 Commit, -3 Encode, -4 alloc). */
int lcpc_oracle_commit(const lcpc_oracle_enc *e, const uint64_t *coeffs_in, size_t len,
                       uint64_t *comm_out, uint64_t *coeffs_out, uint8_t *hashes_out, int threads);
/* merkleize (:690-704) / merkleize_ser (:1128-1158) on a row-major comm */
int lcpc_oracle_merkleize(int field, const uint64_t *comm, size_t n_rows, size_t n_cols,
                          uint8_t *hashes, int serial, int threads);
/* merkle_tree on caller-filled leaves: hashes[0..np2) given, rest computed (:747-785) */
void lcpc_oracle_merkle_tree(uint8_t *hashes, size_t np2, int threads);

/* ---- prove path ---- */
/* collapse_columns (:1095-1123); poly is zeroed here like the call sites do (:1033,1054) */
int lcpc_oracle_collapse(int field, const uint64_t *coeffs, const uint64_t *tensor, uint64_t *poly,
                         size_t n_rows, size_t n_per_row, int serial, int threads);
/* open_column (:788-825): col_out n_rows elements, path_out 32*log2(n_cols) bytes;
 * returns path length, or -1 for ProverError::ColumnNumber */
int lcpc_oracle_open_column(int field, const uint64_t *comm, const uint8_t *hashes, size_t n_rows,
                            size_t n_cols, size_t column, uint64_t *col_out, uint8_t *path_out);
/* verify_column_path (:955-982); returns 1 if the path hashes to root */
int lcpc_oracle_verify_column_path(int field, const uint64_t *col, size_t n_rows,
                                   const uint8_t *path, size_t path_len, size_t col_num,
                                   const uint8_t root[32]);
/* verify_column_value (:985-1000): returns <tensor, col> */
int lcpc_oracle_dot(int field, const uint64_t *a, const uint64_t *b, size_t n, uint64_t *out);

int lcpc_oracle_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
