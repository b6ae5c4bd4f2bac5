/*
 * include/lcpc_b200_host.h -- host-side setup helpers that accompany include/lcpc_b200.h.
 *
 * In a Rust integration these stay in Rust (they ARE the reference's code: `_get_dims`, `matgen`);
 * they are exported here so that hosts without the reference crates (this repo's C++/Python host
 * layer, tests, bench.py) can construct the same encodings.  No device work happens in this header.
 */
#ifndef LCPC_B200_HOST_H
#define LCPC_B200_HOST_H

#include "lcpc_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* n_degree_tests (lcpc-2d/src/lib.rs:613-616) and SizedField::FLOG2 (:61-71) */
size_t lcpc_b200_n_degree_tests(size_t lambda, size_t len, size_t flog2);
unsigned lcpc_b200_field_flog2(int field);
/* LigeroEncodingRho::_n_col_opens / _get_dims (lcpc-ligero-pc/src/lib.rs:61-64, 70-112) */
size_t lcpc_b200_ligero_n_col_opens(size_t rho_num, size_t rho_den);
int lcpc_b200_ligero_get_dims(int field, size_t len, size_t rho_num, size_t rho_den, size_t *n_rows,
                              size_t *n_per_row, size_t *n_cols);
/* SdigEncodingS::_n_col_opens and the n_per_row choice of ::new (lcpc-brakedown-pc/src/lib.rs:57-61, 69-110);
 * code = 1..6 selects SdigCode1..6 (codespec.rs:169-232; the default alias is SdigCode3, lib.rs:19) */
size_t lcpc_b200_sdig_n_col_opens(int code);
int lcpc_b200_sdig_choose_n_per_row(int field, int code, size_t len, size_t *n_per_row);
/* the same for SdigEncodingS::new_ml (:114-124): 2^n_vars monomials, first guess rounded up to a power of two */
int lcpc_b200_sdig_choose_n_per_row_ml(int field, int code, size_t n_vars, size_t *n_per_row);

/* matgen::generate (lcpc-brakedown-pc/src/matgen.rs:28-52): the seeded precodes/postcodes, host memory */
typedef struct lcpc_b200_sdig_code lcpc_b200_sdig_code;
int lcpc_b200_sdig_code_generate(int field, int code, size_t n_per_row, uint64_t seed, lcpc_b200_sdig_code **out);
void lcpc_b200_sdig_code_free(lcpc_b200_sdig_code *c);
size_t lcpc_b200_sdig_code_levels(const lcpc_b200_sdig_code *c);
size_t lcpc_b200_sdig_code_n_per_row(const lcpc_b200_sdig_code *c);
size_t lcpc_b200_sdig_code_codeword_length(const lcpc_b200_sdig_code *c); /* encode.rs:18-33 */
/* borrow one matrix; pointers stay valid until lcpc_b200_sdig_code_free */
int lcpc_b200_sdig_code_matrix(const lcpc_b200_sdig_code *c, size_t level, int is_post, lcpc_b200_csc *out);
/* convenience: lcpc_b200_sdig_new on every matrix of a generated code */
int lcpc_b200_sdig_new_from_code(lcpc_b200_ctx *ctx, const lcpc_b200_sdig_code *c, lcpc_b200_enc **out);

/* SdigEncodingS::new (lcpc-brakedown-pc/src/lib.rs:103-110) without host matrices: matgen::generate (matgen.rs:28-52,
 * 114-188) runs ON THE DEVICE -- the sequential ChaCha20 draw of every level is resolved in parallel (keystream, "where
 * would a column starting at word s end" for every s, pointer doubling over that map) -- straight into the gather form
 * the encoder reads, bit-identical to lcpc_b200_sdig_code_generate.  Codes are cached per context and
 * (field, code, n_per_row, seed).  n_per_row: as chosen by lcpc_b200_sdig_choose_n_per_row. */
int lcpc_b200_sdig_new_seeded(lcpc_b200_ctx *ctx, int field, int code, size_t n_per_row, uint64_t seed, lcpc_b200_enc **out);
/* levels of a Brakedown encoding; one matrix of a device-generated code back in the reference's CSC form: exactly *d
 * sorted row indices per input column (ptrs[c] = c * d); idxs / data: n * d entries each, either may be NULL */
size_t lcpc_b200_enc_sdig_levels(const lcpc_b200_enc *enc);
int lcpc_b200_enc_sdig_matrix(lcpc_b200_enc *enc, size_t level, int is_post, size_t *m, size_t *n, size_t *d, uint64_t *idxs,
                              uint64_t *data);

/* merlin::Transcript (merlin 2.0: STROBE-128 over Keccak-f[1600]), the Fiat-Shamir transcript of prove()/verify()
 * (lcpc-2d/src/lib.rs:16, :304-311, :518-527).  Sequential host work; in a Rust integration this stays merlin. */
int lcpc_b200_transcript_new(const uint8_t *label, size_t n, lcpc_b200_transcript **out); /* Transcript::new */
int lcpc_b200_transcript_clone(const lcpc_b200_transcript *tr, lcpc_b200_transcript **out);
void lcpc_b200_transcript_free(lcpc_b200_transcript *tr);
int lcpc_b200_transcript_append_message(lcpc_b200_transcript *tr, const uint8_t *label, size_t nl, const uint8_t *msg,
                                        size_t n);
int lcpc_b200_transcript_append_u64(lcpc_b200_transcript *tr, const uint8_t *label, size_t nl, uint64_t x);
int lcpc_b200_transcript_challenge_bytes(lcpc_b200_transcript *tr, const uint8_t *label, size_t nl, uint8_t *out,
                                         size_t n);
/* FieldHash::transcript_update (lcpc-2d/src/lib.rs:46-49) for `count` elements: one append_message(label, repr_i)
 * per element; repr = canonical little-endian bytes (to_repr), elem_bytes each */
int lcpc_b200_transcript_append_reprs(lcpc_b200_transcript *tr, const uint8_t *label, size_t nl, const uint8_t *repr,
                                      size_t elem_bytes, size_t count);
/* the column challenge of prove()/verify() (:1073-1080, :903-911): n draws of Uniform::new(0usize, n_cols) from
 * ChaCha20Rng::from_seed(key) */
int lcpc_b200_sample_columns(const uint8_t key[32], size_t n_cols, size_t n, uint64_t *out);

#ifdef __cplusplus
}
#endif
#endif /* LCPC_B200_HOST_H */
