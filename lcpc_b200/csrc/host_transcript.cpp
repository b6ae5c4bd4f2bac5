// lcpc_b200/csrc/host_transcript.cpp -- merlin 2.0 transcript (STROBE-128 over Keccak-f[1600]); see the header.
#include "host_transcript.h"

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

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

// The same permutation produced one output plane at a time, ping-ponging between two state arrays, with the next
// round's column parities accumulated while the outputs are written (the order of the well-known 64-bit
// implementations): every loop has a constant trip count and is unrolled, the arrays become registers.  With
// andn / rorx this is the fastest single-state form on the hosts measured (tools/time_transcript.py); without BMI
// the table-driven form above is.
static inline __attribute__((always_inline)) uint64_t rol_or_id(uint64_t x, unsigned n) { return n ? rotl64(x, n) : x; }
static inline __attribute__((always_inline)) void keccak_round_planes(const uint64_t *S, uint64_t *D, uint64_t *c, uint64_t rc) {
  constexpr unsigned RHO[25] = {0,  1,  62, 28, 27, 36, 44, 6,  55, 20, 3,  10, 43,
                                25, 39, 41, 45, 15, 21, 8,  18, 2,  61, 56, 14};  // r[x][y] at x + 5 y
  uint64_t d[5], b[5];
#pragma GCC unroll 5
  for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rotl64(c[(x + 1) % 5], 1);
#pragma GCC unroll 5
  for (int Y = 0; Y < 5; Y++) {
#pragma GCC unroll 5
    for (int X = 0; X < 5; X++) {
      // pi sends (x, y) to (X, Y) = (y, 2x + 3y): lane (X, Y) comes from y = X, x = 3 (Y - 3 X) mod 5
      const int y = X, x = (3 * (Y + 15 - 3 * X)) % 5;
      b[X] = rol_or_id(S[x + 5 * y] ^ d[x], RHO[x + 5 * y]);
    }
#pragma GCC unroll 5
    for (int X = 0; X < 5; X++) {
      uint64_t v = b[X] ^ (~b[(X + 1) % 5] & b[(X + 2) % 5]);
      if (X == 0 && Y == 0) v ^= rc;
      D[X + 5 * Y] = v;
      c[X] = Y == 0 ? v : c[X] ^ v;
    }
  }
}
static inline __attribute__((always_inline)) void keccak_rounds_planes(uint64_t st[25]) {
  static const uint64_t RC[24] = {
      0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull,
      0x000000000000808bull, 0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull,
      0x000000000000008aull, 0x0000000000000088ull, 0x0000000080008009ull, 0x000000008000000aull,
      0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull, 0x8000000000008003ull,
      0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
      0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
  uint64_t a[25], e[25], c[5];
#pragma GCC unroll 25
  for (int i = 0; i < 25; i++) a[i] = st[i];
#pragma GCC unroll 5
  for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
#pragma GCC unroll 12
  for (int r = 0; r < 24; r += 2) {
    keccak_round_planes(a, e, c, RC[r]);
    keccak_round_planes(e, a, c, RC[r + 1]);
  }
#pragma GCC unroll 25
  for (int i = 0; i < 25; i++) st[i] = a[i];
}

#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("bmi,bmi2"))) static void keccak_f1600_bmi(uint64_t st[25]) { keccak_rounds_planes(st); }
__attribute__((target("bmi,bmi2"))) static void keccak_f1600_bmi_table(uint64_t st[25]) { keccak_rounds(st); }
#endif
static void keccak_f1600_base(uint64_t st[25]) { keccak_rounds(st); }

#if defined(__x86_64__) && defined(__GNUC__)
}  // namespace host
}  // namespace lcpc
#include <immintrin.h>
namespace lcpc {
namespace host {
// AVX-512 build: one register per plane (the five lanes that share y in elements 0..4).  theta is two 3-way xors, two
// in-register lane rotations and a rotate; rho is one variable rotate per plane; pi moves plane y into the register of
// column x' = y with ONE in-register permute (B[y][2x+3y] = A[x][y] reads only plane y), which leaves the state
// transposed -- exactly the form in which chi is five 3-operand logic ops across registers with no shuffles -- and a
// 5x5 transpose (14 two-source shuffles) brings it back to planes for the next round.
__attribute__((target("avx512f,avx512vl"))) static void keccak_f1600_avx512(uint64_t st[25]) {
  static const uint64_t RC[24] = {
      0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull,
      0x000000000000808bull, 0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull,
      0x000000000000008aull, 0x0000000000000088ull, 0x0000000080008009ull, 0x000000008000000aull,
      0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull, 0x8000000000008003ull,
      0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
      0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
  const __mmask8 m5 = 0x1f;
  __m512i p0 = _mm512_maskz_loadu_epi64(m5, st), p1 = _mm512_maskz_loadu_epi64(m5, st + 5),
          p2 = _mm512_maskz_loadu_epi64(m5, st + 10), p3 = _mm512_maskz_loadu_epi64(m5, st + 15),
          p4 = _mm512_maskz_loadu_epi64(m5, st + 20);
  const __m512i i_prev = _mm512_setr_epi64(4, 0, 1, 2, 3, 5, 6, 7), i_next = _mm512_setr_epi64(1, 2, 3, 4, 0, 5, 6, 7);
  // rho offsets r[x][y] of plane y in elements x
  const __m512i r0 = _mm512_setr_epi64(0, 1, 62, 28, 27, 0, 0, 0), r1 = _mm512_setr_epi64(36, 44, 6, 55, 20, 0, 0, 0),
                r2 = _mm512_setr_epi64(3, 10, 43, 25, 39, 0, 0, 0), r3 = _mm512_setr_epi64(41, 45, 15, 21, 8, 0, 0, 0),
                r4 = _mm512_setr_epi64(18, 2, 61, 56, 14, 0, 0, 0);
  // pi: register x' (= source plane y) element y' takes the source's lane (3 y' + x') mod 5
  const __m512i q0 = _mm512_setr_epi64(0, 3, 1, 4, 2, 5, 6, 7), q1 = _mm512_setr_epi64(1, 4, 2, 0, 3, 5, 6, 7),
                q2 = _mm512_setr_epi64(2, 0, 3, 1, 4, 5, 6, 7), q3 = _mm512_setr_epi64(3, 1, 4, 2, 0, 5, 6, 7),
                q4 = _mm512_setr_epi64(4, 2, 0, 3, 1, 5, 6, 7);
  // transpose helpers
  const __m512i t_pairs = _mm512_setr_epi64(0, 8, 1, 9, 2, 10, 3, 11), t_last = _mm512_setr_epi64(4, 12, 4, 12, 4, 12, 4, 12);
  const __m512i u0 = _mm512_setr_epi64(0, 1, 8, 9, 0, 0, 0, 0), u1 = _mm512_setr_epi64(2, 3, 10, 11, 0, 0, 0, 0),
                u2 = _mm512_setr_epi64(4, 5, 12, 13, 0, 0, 0, 0), u3 = _mm512_setr_epi64(6, 7, 14, 15, 0, 0, 0, 0);
  const __m512i e0 = _mm512_set1_epi64(0), e1 = _mm512_set1_epi64(1), e2 = _mm512_set1_epi64(2), e3 = _mm512_set1_epi64(3),
                e4 = _mm512_set1_epi64(4);
  const __mmask8 k4 = 0x10;
  for (int round = 0; round < 24; round++) {
    // theta
    __m512i c = _mm512_ternarylogic_epi64(_mm512_ternarylogic_epi64(p0, p1, p2, 0x96), p3, p4, 0x96);
    __m512i d = _mm512_xor_si512(_mm512_permutexvar_epi64(i_prev, c), _mm512_rol_epi64(_mm512_permutexvar_epi64(i_next, c), 1));
    // rho, then pi into the transposed form (register = column x', element = row y')
    __m512i x0 = _mm512_permutexvar_epi64(q0, _mm512_rolv_epi64(_mm512_xor_si512(p0, d), r0));
    __m512i x1 = _mm512_permutexvar_epi64(q1, _mm512_rolv_epi64(_mm512_xor_si512(p1, d), r1));
    __m512i x2 = _mm512_permutexvar_epi64(q2, _mm512_rolv_epi64(_mm512_xor_si512(p2, d), r2));
    __m512i x3 = _mm512_permutexvar_epi64(q3, _mm512_rolv_epi64(_mm512_xor_si512(p3, d), r3));
    __m512i x4 = _mm512_permutexvar_epi64(q4, _mm512_rolv_epi64(_mm512_xor_si512(p4, d), r4));
    // chi: a ^ (~b & c) across columns, iota into lane (0, 0)
    __m512i y0 = _mm512_ternarylogic_epi64(x0, x1, x2, 0xD2), y1 = _mm512_ternarylogic_epi64(x1, x2, x3, 0xD2),
            y2 = _mm512_ternarylogic_epi64(x2, x3, x4, 0xD2), y3 = _mm512_ternarylogic_epi64(x3, x4, x0, 0xD2),
            y4 = _mm512_ternarylogic_epi64(x4, x0, x1, 0xD2);
    y0 = _mm512_xor_si512(y0, _mm512_maskz_set1_epi64(0x01, (long long)RC[round]));
    // back to planes: p_j[x'] = y_{x'}[j]
    __m512i a = _mm512_permutex2var_epi64(y0, t_pairs, y1), a4 = _mm512_permutex2var_epi64(y0, t_last, y1);
    __m512i b = _mm512_permutex2var_epi64(y2, t_pairs, y3), b4 = _mm512_permutex2var_epi64(y2, t_last, y3);
    p0 = _mm512_mask_permutexvar_epi64(_mm512_permutex2var_epi64(a, u0, b), k4, e0, y4);
    p1 = _mm512_mask_permutexvar_epi64(_mm512_permutex2var_epi64(a, u1, b), k4, e1, y4);
    p2 = _mm512_mask_permutexvar_epi64(_mm512_permutex2var_epi64(a, u2, b), k4, e2, y4);
    p3 = _mm512_mask_permutexvar_epi64(_mm512_permutex2var_epi64(a, u3, b), k4, e3, y4);
    p4 = _mm512_mask_permutexvar_epi64(_mm512_permutex2var_epi64(a4, u0, b4), k4, e4, y4);
  }
  _mm512_mask_storeu_epi64(st, m5, p0);
  _mm512_mask_storeu_epi64(st + 5, m5, p1);
  _mm512_mask_storeu_epi64(st + 10, m5, p2);
  _mm512_mask_storeu_epi64(st + 15, m5, p3);
  _mm512_mask_storeu_epi64(st + 20, m5, p4);
}
#endif

#if defined(__x86_64__) && defined(__GNUC__)
static void (*keccak_pick())(uint64_t *) {
  static void (*const impl)(uint64_t *) = [] {
    // LCPC_B200_KECCAK = base | bmi | bmi_table | avx512 forces a build the CPU supports (tests compare all of them);
    const char *want = getenv("LCPC_B200_KECCAK");
    const std::string w = want ? want : "";
    const bool has_bmi = __builtin_cpu_supports("bmi") && __builtin_cpu_supports("bmi2");
    const bool has_512 = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512vl");
    typedef void (*Fn)(uint64_t *);
    if (w == "base") return (Fn)keccak_f1600_base;
    if (w == "bmi" && has_bmi) return (Fn)keccak_f1600_bmi;
    if (w == "bmi_table" && has_bmi) return (Fn)keccak_f1600_bmi_table;
    if (w == "avx512" && has_512) return (Fn)keccak_f1600_avx512;
    // Preference measured end to end (tools/time_transcript.py, ms per 65 536 absorbed Ft255 coefficients, this
    // image's Xeon: plane-wise scalar with BMI 7.9, AVX-512 in place 8.7, portable 10.7; the bare permutations are
    // within 2 % of each other -- the vector build pays for entering and leaving 512-bit code every 166 bytes).
    // A timed selection at start-up was tried and dropped: on a shared host a 0.1 ms probe picks a different build
    // from run to run.
    if (has_bmi) return (Fn)keccak_f1600_bmi;
    if (has_512) return (Fn)keccak_f1600_avx512;
    return (Fn)keccak_f1600_base;
  }();
  return impl;
}
#endif

void keccak_f1600(uint64_t st[25]) {
#if defined(__x86_64__) && defined(__GNUC__)
  keccak_pick()(st);
#else
  keccak_f1600_base(st);
#endif
}

// The duplex state is written byte-wise by STROBE right before and read byte-wise right after every permutation.  The
// vector build loads and stores it with five masked 512-bit moves and is fastest working on it in place; the scalar
// builds' 25 64-bit loads would each wait for the narrower stores still in flight (store forwarding cannot merge
// them: measured 16.2 against 8.0 ms per 65 536 absorbs), so they work on a copy, as memcpy's wide moves do not.
bool keccak_in_place() {
#if defined(__x86_64__) && defined(__GNUC__)
  static const bool v = keccak_pick() == keccak_f1600_avx512;
  return v;
#else
  return false;
#endif
}

// ---- STROBE-128 as merlin instantiates it (merlin/src/strobe.rs) --------------------------------
Strobe128::Strobe128(const uint8_t *protocol_label, size_t n) {
  memset(lanes_, 0, sizeof lanes_);
  const uint8_t head[6] = {1, (uint8_t)(R + 2), 1, 0, 1, 96};
  memcpy(st(), head, 6);
  memcpy(st() + 6, "STROBEv1.0.2", 12);
  keccak_f1600(lanes_);  // lanes are little-endian; so is every host this builds for
  meta_ad(protocol_label, n, false);
}

void Strobe128::run_f() {
  uint8_t *s = st();
  s[pos_] ^= pos_begin_;
  s[pos_ + 1] ^= 0x04;
  s[R + 1] ^= 0x80;
  if (keccak_in_place()) {
    keccak_f1600(lanes_);  // the byte view and the lane view are the same storage
  } else {
    uint64_t lanes[25];
    memcpy(lanes, lanes_, 200);
    keccak_f1600(lanes);
    memcpy(lanes_, lanes, 200);
  }
  pos_ = 0, pos_begin_ = 0;
}

void Strobe128::absorb(const uint8_t *data, size_t n) {
  while (n) {
    size_t run = (size_t)(R - pos_) < n ? (size_t)(R - pos_) : n;
    size_t i = 0;
    for (; i + 8 <= run; i += 8) {  // eight bytes at a time (unaligned access through memcpy)
      uint64_t a, b;
      memcpy(&a, st() + pos_ + i, 8);
      memcpy(&b, data + i, 8);
      a ^= b;
      memcpy(st() + pos_ + i, &a, 8);
    }
    for (; i < run; i++) st()[pos_ + i] ^= data[i];
    pos_ = (uint8_t)(pos_ + run), data += run, n -= run;
    if (pos_ == R) run_f();
  }
}

void Strobe128::overwrite(const uint8_t *data, size_t n) {
  for (size_t i = 0; i < n; i++) {
    st()[pos_++] = data[i];
    if (pos_ == R) run_f();
  }
}

void Strobe128::squeeze(uint8_t *data, size_t n) {
  for (size_t i = 0; i < n; i++) {
    data[i] = st()[pos_];
    st()[pos_++] = 0;
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

// meta_ad(label, false) followed by meta_ad(len, true), the pair every merlin message and challenge starts with: the
// operation header, the label and the length are consecutive duplex input, so they are absorbed in ONE pass (the
// prover appends 65 536 messages per vector; five small absorb calls per message were a third of the transcript time)
void Strobe128::meta_ad_label_len(const uint8_t *label, size_t nl, const uint8_t len[4]) {
  if (nl > 56) {
    meta_ad(label, nl, false);
    meta_ad(len, 4, true);
    return;
  }
  uint8_t buf[64];
  buf[0] = pos_begin_, buf[1] = FLAG_M | FLAG_A;  // begin_op: the header names where the previous operation began
  pos_begin_ = (uint8_t)(pos_ + 1);
  cur_flags_ = FLAG_M | FLAG_A;
  memcpy(buf + 2, label, nl);
  memcpy(buf + 2 + nl, len, 4);
  absorb(buf, 2 + nl + 4);
}
void Strobe128::meta_ad(const uint8_t *data, size_t n, bool more) {
  begin_op(FLAG_M | FLAG_A, more);
  absorb(data, n);
}
void Strobe128::ad(const uint8_t *data, size_t n, bool more) {
  if (!more && n <= 62) {  // header and a short message in one pass (a field element is 8..32 bytes)
    uint8_t buf[64];
    buf[0] = pos_begin_, buf[1] = FLAG_A;
    pos_begin_ = (uint8_t)(pos_ + 1);
    cur_flags_ = FLAG_A;
    memcpy(buf + 2, data, n);
    absorb(buf, 2 + n);
    return;
  }
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
  strobe_.meta_ad_label_len(label, nl, len);
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
  strobe_.meta_ad_label_len(label, nl, len);
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
