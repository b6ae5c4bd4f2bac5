/*
 * oracle/chacha_rng.c -- TEST INFRASTRUCTURE ONLY. See chacha_rng.h.
 */
#include "chacha_rng.h"

#include <string.h>

static inline uint32_t rotl32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }

#define QR(a, b, c, d) \
  do {                 \
    a += b;            \
    d ^= a;            \
    d = rotl32(d, 16); \
    c += d;            \
    b ^= c;            \
    b = rotl32(b, 12); \
    a += b;            \
    d ^= a;            \
    d = rotl32(d, 8);  \
    c += d;            \
    b ^= c;            \
    b = rotl32(b, 7);  \
  } while (0)

void chacha_block(const uint32_t key[8], uint64_t counter, uint64_t stream, uint32_t out[16]) {
  uint32_t s[16], x[16];
  s[0] = 0x61707865u;
  s[1] = 0x3320646eu;
  s[2] = 0x79622d32u;
  s[3] = 0x6b206574u;
  for (int i = 0; i < 8; i++) s[4 + i] = key[i];
  s[12] = (uint32_t)counter;
  s[13] = (uint32_t)(counter >> 32);
  s[14] = (uint32_t)stream;
  s[15] = (uint32_t)(stream >> 32);
  memcpy(x, s, sizeof x);
  for (int i = 0; i < 10; i++) {
    QR(x[0], x[4], x[8], x[12]);
    QR(x[1], x[5], x[9], x[13]);
    QR(x[2], x[6], x[10], x[14]);
    QR(x[3], x[7], x[11], x[15]);
    QR(x[0], x[5], x[10], x[15]);
    QR(x[1], x[6], x[11], x[12]);
    QR(x[2], x[7], x[8], x[13]);
    QR(x[3], x[4], x[9], x[14]);
  }
  for (int i = 0; i < 16; i++) out[i] = x[i] + s[i];
}

void chacha_from_seed(chacha_rng *r, const uint8_t seed[32]) {
  for (int i = 0; i < 8; i++)
    r->key[i] = (uint32_t)seed[4 * i] | ((uint32_t)seed[4 * i + 1] << 8) |
                ((uint32_t)seed[4 * i + 2] << 16) | ((uint32_t)seed[4 * i + 3] << 24);
  r->counter = 0;
  r->stream = 0;
  r->idx = 16;
}

/* rand_core 0.6 SeedableRng::seed_from_u64: PCG32 expands the u64 into the 32-byte seed */
void chacha_seed_from_u64(chacha_rng *r, uint64_t state) {
  const uint64_t MUL = 6364136223846793005ULL;
  const uint64_t INC = 11634580027462260723ULL;
  uint8_t seed[32];
  for (int i = 0; i < 8; i++) {
    state = state * MUL + INC;
    uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
    uint32_t rot = (uint32_t)(state >> 59);
    uint32_t x = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
    seed[4 * i] = (uint8_t)x;
    seed[4 * i + 1] = (uint8_t)(x >> 8);
    seed[4 * i + 2] = (uint8_t)(x >> 16);
    seed[4 * i + 3] = (uint8_t)(x >> 24);
  }
  chacha_from_seed(r, seed);
}

void chacha_set_stream(chacha_rng *r, uint64_t stream) {
  r->stream = stream;
  /* only ever called on a fresh generator (matgen.rs:43-44): nothing buffered to redo */
  r->idx = 16;
}

uint32_t chacha_next_u32(chacha_rng *r) {
  if (r->idx >= 16) {
    chacha_block(r->key, r->counter, r->stream, r->buf);
    r->counter++;
    r->idx = 0;
  }
  return r->buf[r->idx++];
}

uint64_t chacha_next_u64(void *vr) {
  chacha_rng *r = (chacha_rng *)vr;
  /* BlockRng::next_u64: two consecutive words, low word first */
  uint64_t lo = chacha_next_u32(r);
  uint64_t hi = chacha_next_u32(r);
  return lo | (hi << 32);
}

uint64_t chacha_uniform(chacha_rng *r, uint64_t range) {
  /* UniformInt::new(0, range) -> new_inclusive(0, range-1); sample() with the widening-multiply zone */
  uint64_t ints_to_reject = (UINT64_MAX - range + 1) % range;
  uint64_t zone = UINT64_MAX - ints_to_reject;
  for (;;) {
    uint64_t v = chacha_next_u64(r);
    unsigned __int128 m = (unsigned __int128)v * range;
    uint64_t hi = (uint64_t)(m >> 64), lo = (uint64_t)m;
    if (lo <= zone) return hi;
  }
}
