/*
 * oracle/chacha_rng.h -- TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
 *
 * Restates the random streams the reference draws from un-vendored crates
 * (SURVEY.md App. B4/B5 -- published behaviour of rand_core 0.6, rand_chacha 0.3, rand 0.8):
 *   - ChaCha20Rng::seed_from_u64 + set_stream   (lcpc-brakedown-pc/src/matgen.rs:43-44)
 *   - ChaCha20Rng::from_seed(key)               (lcpc-2d/src/lib.rs:1026-1028, 1073-1075)
 *   - Uniform::new(0usize, m).sample(rng)       (matgen.rs:119,147-158; lib.rs:1077-1080)
 * PARITY UNPINNED: the reference holds no known-answer vector for these streams.
 * The ChaCha20 block function itself is pinned against RFC 8439 section 2.3.2 in tests/.
 */
#ifndef LCPC_ORACLE_CHACHA_RNG_H
#define LCPC_ORACLE_CHACHA_RNG_H

#include <stdint.h>

typedef struct {
  uint32_t key[8];
  uint64_t counter; /* 64-bit block counter, state words 12..13 */
  uint64_t stream;  /* 64-bit stream id,     state words 14..15 */
  uint32_t buf[16];
  int idx; /* next unread word in buf; 16 = empty */
} chacha_rng;

void chacha_block(const uint32_t key[8], uint64_t counter, uint64_t stream, uint32_t out[16]);
void chacha_from_seed(chacha_rng *r, const uint8_t seed[32]);
void chacha_seed_from_u64(chacha_rng *r, uint64_t state);
void chacha_set_stream(chacha_rng *r, uint64_t stream);
uint32_t chacha_next_u32(chacha_rng *r);
uint64_t chacha_next_u64(void *r); /* void* so it can serve as the field-template word source */
/* rand 0.8 UniformInt<usize>::sample for the half-open range [0, range) */
uint64_t chacha_uniform(chacha_rng *r, uint64_t range);

#endif
