// lcpc_b200/csrc/host_transcript.cpp -- merlin 2.0 transcript (STROBE-128 over Keccak-f[1600]); see the header.
#include "host_transcript.h"

#include <cstring>
#include <new>

#include "../../include/lcpc_b200_host.h"

namespace lcpc {
namespace host {

static inline uint64_t rotl64(uint64_t x, unsigned n) { return (x << n) | (x >> (64 - n)); }

// FIPS 202 Keccak-p[1600, 24].  Lane (x, y) is st[x + 5 y].  theta, then rho and pi fused into one table-driven move
// (lane i goes to PI_DST[i] rotated by RHO[i]), then chi and iota; the fixed-trip loops are unrolled by the compiler.
// The prover absorbs every coefficient of p_random / p_eval through this permutation (lcpc-2d/src/lib.rs:1043-1045,
// :1061-1063: ~18 700 permutations per 65 536 Ft255 elements), which makes it the prover's host-side hot spot.
// Two builds of the same rounds, chosen once at run time: with BMI1/BMI2 (andn for chi, rorx for rho) the permutation
// takes about half the time of the baseline x86-64 build.
static inline __attribute__((always_inline)) void keccak_rounds(uint64_t st[25]) {
  static const uint64_t RC[24] = {
      0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull,
      0x000000000000808bull, 0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull,
      0x000000000000008aull, 0x0000000000000088ull, 0x0000000080008009ull, 0x000000008000000aull,
      0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull, 0x8000000000008003ull,
      0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
      0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
  // rho offsets r[x][y] at index x + 5 y, and pi: (x, y) -> (y, 2x + 3y)
  static constexpr unsigned RHO[25] = {0,  1,  62, 28, 27, 36, 44, 6,  55, 20, 3,  10, 43,
                                       25, 39, 41, 45, 15, 21, 8,  18, 2,  61, 56, 14};
  static constexpr unsigned PI_DST[25] = {0, 10, 20, 5, 15, 16, 1, 11, 21, 6, 7, 17, 2, 12, 22, 23, 8, 18, 3, 13, 14, 24, 9, 19, 4};
  uint64_t a[25], b[25], c[5], d[5];
  for (int i = 0; i < 25; i++) a[i] = st[i];
  for (int round = 0; round < 24; round++) {
#pragma GCC unroll 5
    for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
#pragma GCC unroll 5
    for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rotl64(c[(x + 1) % 5], 1);
#pragma GCC unroll 25
    for (int i = 0; i < 25; i++) {
      const uint64_t v = a[i] ^ d[i % 5];
      b[PI_DST[i]] = RHO[i] ? rotl64(v, RHO[i]) : v;
    }
#pragma GCC unroll 25
    for (int i = 0; i < 25; i++) {
      const int y5 = (i / 5) * 5, x = i % 5;
      a[i] = b[i] ^ (~b[y5 + (x + 1) % 5] & b[y5 + (x + 2) % 5]);
    }
    a[0] ^= RC[round];
  }
  for (int i = 0; i < 25; i++) st[i] = a[i];
}

#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("bmi,bmi2"))) static void keccak_f1600_bmi(uint64_t st[25]) { keccak_rounds(st); }
#endif
static void keccak_f1600_base(uint64_t st[25]) { keccak_rounds(st); }

void keccak_f1600(uint64_t st[25]) {
#if defined(__x86_64__) && defined(__GNUC__)
  static void (*const impl)(uint64_t *) =
      (__builtin_cpu_supports("bmi") && __builtin_cpu_supports("bmi2")) ? keccak_f1600_bmi : keccak_f1600_base;
  impl(st);
#else
  keccak_f1600_base(st);
#endif
}

// ---- STROBE-128 as merlin instantiates it (merlin/src/strobe.rs) --------------------------------
Strobe128::Strobe128(const uint8_t *protocol_label, size_t n) {
  memset(st_, 0, sizeof st_);
  const uint8_t head[6] = {1, (uint8_t)(R + 2), 1, 0, 1, 96};
  memcpy(st_, head, 6);
  memcpy(st_ + 6, "STROBEv1.0.2", 12);
  uint64_t lanes[25];
  memcpy(lanes, st_, 200);  // lanes are little-endian; so is every host this builds for
  keccak_f1600(lanes);
  memcpy(st_, lanes, 200);
  meta_ad(protocol_label, n, false);
}

void Strobe128::run_f() {
  st_[pos_] ^= pos_begin_;
  st_[pos_ + 1] ^= 0x04;
  st_[R + 1] ^= 0x80;
  uint64_t lanes[25];
  memcpy(lanes, st_, 200);
  keccak_f1600(lanes);
  memcpy(st_, lanes, 200);
  pos_ = 0, pos_begin_ = 0;
}

void Strobe128::absorb(const uint8_t *data, size_t n) {
  while (n) {
    size_t run = (size_t)(R - pos_) < n ? (size_t)(R - pos_) : n;
    for (size_t i = 0; i < run; i++) st_[pos_ + i] ^= data[i];
    pos_ = (uint8_t)(pos_ + run), data += run, n -= run;
    if (pos_ == R) run_f();
  }
}

void Strobe128::overwrite(const uint8_t *data, size_t n) {
  for (size_t i = 0; i < n; i++) {
    st_[pos_++] = data[i];
    if (pos_ == R) run_f();
  }
}

void Strobe128::squeeze(uint8_t *data, size_t n) {
  for (size_t i = 0; i < n; i++) {
    data[i] = st_[pos_];
    st_[pos_++] = 0;
    if (pos_ == R) run_f();
  }
}

void Strobe128::begin_op(uint8_t flags, bool more) {
  if (more) return;  // continuation of the current operation (merlin asserts flags == cur_flags)
  const uint8_t old_begin = pos_begin_;
  pos_begin_ = (uint8_t)(pos_ + 1);
  cur_flags_ = flags;
  const uint8_t hdr[2] = {old_begin, flags};
  absorb(hdr, 2);
  const bool force_f = (flags & (FLAG_C | FLAG_K)) != 0;
  if (force_f && pos_ != 0) run_f();
}

void Strobe128::meta_ad(const uint8_t *data, size_t n, bool more) {
  begin_op(FLAG_M | FLAG_A, more);
  absorb(data, n);
}
void Strobe128::ad(const uint8_t *data, size_t n, bool more) {
  begin_op(FLAG_A, more);
  absorb(data, n);
}
void Strobe128::prf(uint8_t *data, size_t n, bool more) {
  begin_op(FLAG_I | FLAG_A | FLAG_C, more);
  squeeze(data, n);
}
void Strobe128::key(const uint8_t *data, size_t n, bool more) {
  begin_op(FLAG_A | FLAG_C, more);
  overwrite(data, n);
}

// ---- merlin::Transcript (merlin/src/transcript.rs) -----------------------------------------------
static const uint8_t kMerlinLabel[] = "Merlin v1.0";
static const uint8_t kDomSep[] = "dom-sep";

Transcript::Transcript(const uint8_t *label, size_t n) : strobe_(kMerlinLabel, sizeof kMerlinLabel - 1) {
  append_message(kDomSep, sizeof kDomSep - 1, label, n);
}

static inline void le32(uint8_t out[4], size_t n) {
  out[0] = (uint8_t)n, out[1] = (uint8_t)(n >> 8), out[2] = (uint8_t)(n >> 16), out[3] = (uint8_t)(n >> 24);
}

void Transcript::append_message(const uint8_t *label, size_t nl, const uint8_t *msg, size_t n) {
  uint8_t len[4];
  le32(len, n);
  strobe_.meta_ad(label, nl, false);
  strobe_.meta_ad(len, 4, true);
  strobe_.ad(msg, n, false);
}

void Transcript::append_u64(const uint8_t *label, size_t nl, uint64_t x) {
  uint8_t b[8];
  for (int i = 0; i < 8; i++) b[i] = (uint8_t)(x >> (8 * i));
  append_message(label, nl, b, 8);
}

void Transcript::challenge_bytes(const uint8_t *label, size_t nl, uint8_t *out, size_t n) {
  uint8_t len[4];
  le32(len, n);
  strobe_.meta_ad(label, nl, false);
  strobe_.meta_ad(len, 4, true);
  strobe_.prf(out, n, false);
}

void Transcript::append_elems(const uint8_t *label, size_t nl, const uint8_t *repr, size_t elem_bytes, size_t count) {
  for (size_t i = 0; i < count; i++) append_message(label, nl, repr + i * elem_bytes, elem_bytes);
}

}  // namespace host
}  // namespace lcpc

// ---- C ABI (include/lcpc_b200_host.h) ------------------------------------------------------------
extern "C" {

int lcpc_b200_transcript_new(const uint8_t *label, size_t n, lcpc_b200_transcript **out) {
  if (!out || (!label && n)) return LCPC_B200_ERR_BAD_ARG;
  auto *t = new (std::nothrow) lcpc_b200_transcript(label, n);
  if (!t) return LCPC_B200_ERR_OOM;
  *out = t;
  return LCPC_B200_OK;
}
int lcpc_b200_transcript_clone(const lcpc_b200_transcript *tr, lcpc_b200_transcript **out) {
  if (!tr || !out) return LCPC_B200_ERR_BAD_ARG;
  auto *t = new (std::nothrow) lcpc_b200_transcript(*tr);
  if (!t) return LCPC_B200_ERR_OOM;
  *out = t;
  return LCPC_B200_OK;
}
void lcpc_b200_transcript_free(lcpc_b200_transcript *tr) { delete tr; }
int lcpc_b200_transcript_append_message(lcpc_b200_transcript *tr, const uint8_t *label, size_t nl, const uint8_t *msg,
                                        size_t n) {
  if (!tr || (!label && nl) || (!msg && n) || n > 0xffffffffull) return LCPC_B200_ERR_BAD_ARG;
  tr->tr.append_message(label, nl, msg, n);
  return LCPC_B200_OK;
}
int lcpc_b200_transcript_append_u64(lcpc_b200_transcript *tr, const uint8_t *label, size_t nl, uint64_t x) {
  if (!tr || (!label && nl)) return LCPC_B200_ERR_BAD_ARG;
  tr->tr.append_u64(label, nl, x);
  return LCPC_B200_OK;
}
int lcpc_b200_transcript_challenge_bytes(lcpc_b200_transcript *tr, const uint8_t *label, size_t nl, uint8_t *out,
                                         size_t n) {
  if (!tr || (!label && nl) || (!out && n) || n > 0xffffffffull) return LCPC_B200_ERR_BAD_ARG;
  tr->tr.challenge_bytes(label, nl, out, n);
  return LCPC_B200_OK;
}
int lcpc_b200_transcript_append_reprs(lcpc_b200_transcript *tr, const uint8_t *label, size_t nl, const uint8_t *repr,
                                      size_t elem_bytes, size_t count) {
  if (!tr || (!label && nl) || (!repr && count) || !elem_bytes) return LCPC_B200_ERR_BAD_ARG;
  tr->tr.append_elems(label, nl, repr, elem_bytes, count);
  return LCPC_B200_OK;
}

}  // extern "C"
