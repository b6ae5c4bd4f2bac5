/*
 * oracle/blake3_ref.c -- TEST INFRASTRUCTURE ONLY. See blake3_ref.h.
 */
#include "blake3_ref.h"

#include <string.h>

enum { CHUNK_START = 1, CHUNK_END = 2, PARENT = 4, ROOT = 8 };

static const uint32_t B3_IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                                  0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};

static const uint8_t MSG_PERM[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};

static inline uint32_t rotr32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

static inline void g(uint32_t *s, int a, int b, int c, int d, uint32_t mx, uint32_t my) {
  s[a] = s[a] + s[b] + mx;
  s[d] = rotr32(s[d] ^ s[a], 16);
  s[c] = s[c] + s[d];
  s[b] = rotr32(s[b] ^ s[c], 12);
  s[a] = s[a] + s[b] + my;
  s[d] = rotr32(s[d] ^ s[a], 8);
  s[c] = s[c] + s[d];
  s[b] = rotr32(s[b] ^ s[c], 7);
}

/* full 16-word compression output */
static void compress(const uint32_t cv[8], const uint32_t block_words[16], uint64_t counter,
                     uint32_t block_len, uint32_t flags, uint32_t out[16]) {
  uint32_t s[16], m[16], t[16];
  for (int i = 0; i < 8; i++) s[i] = cv[i];
  for (int i = 0; i < 4; i++) s[8 + i] = B3_IV[i];
  s[12] = (uint32_t)counter;
  s[13] = (uint32_t)(counter >> 32);
  s[14] = block_len;
  s[15] = flags;
  for (int i = 0; i < 16; i++) m[i] = block_words[i];
  for (int r = 0; r < 7; r++) {
    g(s, 0, 4, 8, 12, m[0], m[1]);
    g(s, 1, 5, 9, 13, m[2], m[3]);
    g(s, 2, 6, 10, 14, m[4], m[5]);
    g(s, 3, 7, 11, 15, m[6], m[7]);
    g(s, 0, 5, 10, 15, m[8], m[9]);
    g(s, 1, 6, 11, 12, m[10], m[11]);
    g(s, 2, 7, 8, 13, m[12], m[13]);
    g(s, 3, 4, 9, 14, m[14], m[15]);
    for (int i = 0; i < 16; i++) t[i] = m[MSG_PERM[i]];
    for (int i = 0; i < 16; i++) m[i] = t[i];
  }
  for (int i = 0; i < 8; i++) {
    out[i] = s[i] ^ s[i + 8];
    out[i + 8] = s[i + 8] ^ cv[i];
  }
}

static void words_from_le(const uint8_t *b, uint32_t *w, int nwords) {
  for (int i = 0; i < nwords; i++)
    w[i] = (uint32_t)b[4 * i] | ((uint32_t)b[4 * i + 1] << 8) | ((uint32_t)b[4 * i + 2] << 16) |
           ((uint32_t)b[4 * i + 3] << 24);
}

/* "Output" of the spec: the not-yet-compressed last node, so ROOT can be added late */
typedef struct {
  uint32_t input_cv[8];
  uint32_t block_words[16];
  uint64_t counter;
  uint32_t block_len;
  uint32_t flags;
} b3_output;

static void output_cv(const b3_output *o, uint32_t cv[8]) {
  uint32_t out[16];
  compress(o->input_cv, o->block_words, o->counter, o->block_len, o->flags, out);
  memcpy(cv, out, 32);
}

static void output_root_bytes(const b3_output *o, uint8_t out32[B3_OUT_LEN]) {
  uint32_t out[16];
  compress(o->input_cv, o->block_words, 0, o->block_len, o->flags | ROOT, out);
  for (int i = 0; i < 8; i++) {
    out32[4 * i] = (uint8_t)out[i];
    out32[4 * i + 1] = (uint8_t)(out[i] >> 8);
    out32[4 * i + 2] = (uint8_t)(out[i] >> 16);
    out32[4 * i + 3] = (uint8_t)(out[i] >> 24);
  }
}

static void chunk_init(b3_chunk_state *c, uint64_t counter) {
  memcpy(c->cv, B3_IV, 32);
  c->chunk_counter = counter;
  memset(c->block, 0, B3_BLOCK_LEN);
  c->block_len = 0;
  c->blocks_compressed = 0;
}

static size_t chunk_len(const b3_chunk_state *c) {
  return (size_t)B3_BLOCK_LEN * c->blocks_compressed + c->block_len;
}

static uint32_t chunk_start_flag(const b3_chunk_state *c) {
  return c->blocks_compressed == 0 ? CHUNK_START : 0;
}

static void chunk_update(b3_chunk_state *c, const uint8_t *in, size_t len) {
  while (len > 0) {
    if (c->block_len == B3_BLOCK_LEN) {
      uint32_t w[16], out[16];
      words_from_le(c->block, w, 16);
      compress(c->cv, w, c->chunk_counter, B3_BLOCK_LEN, chunk_start_flag(c), out);
      memcpy(c->cv, out, 32);
      c->blocks_compressed++;
      memset(c->block, 0, B3_BLOCK_LEN);
      c->block_len = 0;
    }
    size_t want = B3_BLOCK_LEN - c->block_len;
    size_t take = len < want ? len : want;
    memcpy(c->block + c->block_len, in, take);
    c->block_len += (uint8_t)take;
    in += take;
    len -= take;
  }
}

static void chunk_output(const b3_chunk_state *c, b3_output *o) {
  memcpy(o->input_cv, c->cv, 32);
  words_from_le(c->block, o->block_words, 16);
  o->counter = c->chunk_counter;
  o->block_len = c->block_len;
  o->flags = chunk_start_flag(c) | CHUNK_END;
}

static void parent_output(const uint32_t left[8], const uint32_t right[8], b3_output *o) {
  memcpy(o->input_cv, B3_IV, 32);
  memcpy(o->block_words, left, 32);
  memcpy(o->block_words + 8, right, 32);
  o->counter = 0;
  o->block_len = B3_BLOCK_LEN;
  o->flags = PARENT;
}

void b3_init(b3_hasher *h) {
  chunk_init(&h->chunk, 0);
  h->cv_stack_len = 0;
}

static void add_chunk_cv(b3_hasher *h, uint32_t new_cv[8], uint64_t total_chunks) {
  while ((total_chunks & 1) == 0) {
    b3_output o;
    h->cv_stack_len--;
    parent_output(h->cv_stack[h->cv_stack_len], new_cv, &o);
    output_cv(&o, new_cv);
    total_chunks >>= 1;
  }
  memcpy(h->cv_stack[h->cv_stack_len], new_cv, 32);
  h->cv_stack_len++;
}

void b3_update(b3_hasher *h, const void *input, size_t len) {
  const uint8_t *in = (const uint8_t *)input;
  while (len > 0) {
    if (chunk_len(&h->chunk) == B3_CHUNK_LEN) {
      b3_output o;
      uint32_t cv[8];
      chunk_output(&h->chunk, &o);
      output_cv(&o, cv);
      uint64_t total = h->chunk.chunk_counter + 1;
      add_chunk_cv(h, cv, total);
      chunk_init(&h->chunk, total);
    }
    size_t want = B3_CHUNK_LEN - chunk_len(&h->chunk);
    size_t take = len < want ? len : want;
    chunk_update(&h->chunk, in, take);
    in += take;
    len -= take;
  }
}

void b3_finalize(const b3_hasher *h, uint8_t out[B3_OUT_LEN]) {
  b3_output o;
  chunk_output(&h->chunk, &o);
  int remaining = h->cv_stack_len;
  while (remaining > 0) {
    uint32_t cv[8];
    remaining--;
    output_cv(&o, cv);
    parent_output(h->cv_stack[remaining], cv, &o);
  }
  output_root_bytes(&o, out);
}

void b3_hash(const void *input, size_t len, uint8_t out[B3_OUT_LEN]) {
  b3_hasher h;
  b3_init(&h);
  b3_update(&h, input, len);
  b3_finalize(&h, out);
}
