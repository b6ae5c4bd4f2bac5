/*
 * oracle/blake3_ref.h -- TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
 *
 * Portable BLAKE3 (default hash mode, 32-byte output), written from the published BLAKE3
 * specification. The reference hashes with the un-vendored crate `blake3 = "1"` through the
 * `digest::Digest` trait (lcpc-2d/Cargo.toml dev-deps; lcpc-ligero-pc/src/bench.rs:12), using only
 * new / update / finalize / finalize_reset (lcpc-2d/src/lib.rs:719-735, 770-775).
 * Pinned in tests/ against the Python `blake3` package (a binding of that same Rust crate) and
 * against fixtures under tests/golden/ generated with it.
 */
#ifndef LCPC_ORACLE_BLAKE3_REF_H
#define LCPC_ORACLE_BLAKE3_REF_H

#include <stddef.h>
#include <stdint.h>

#define B3_BLOCK_LEN 64
#define B3_CHUNK_LEN 1024
#define B3_OUT_LEN 32
#define B3_MAX_DEPTH 54

typedef struct {
  uint32_t cv[8];
  uint64_t chunk_counter;
  uint8_t block[B3_BLOCK_LEN];
  uint8_t block_len;
  uint8_t blocks_compressed;
} b3_chunk_state;

typedef struct {
  b3_chunk_state chunk;
  uint32_t cv_stack[B3_MAX_DEPTH][8];
  uint8_t cv_stack_len;
} b3_hasher;

void b3_init(b3_hasher *h);                                   /* Digest::new            */
void b3_update(b3_hasher *h, const void *input, size_t len);  /* Digest::update         */
void b3_finalize(const b3_hasher *h, uint8_t out[B3_OUT_LEN]); /* Digest::finalize       */
void b3_hash(const void *input, size_t len, uint8_t out[B3_OUT_LEN]);

#endif
