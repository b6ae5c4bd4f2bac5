/*
 * oracle/field_tmpl.h -- TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
 *
 * Prime-field arithmetic "template": include once per field with
 *     #define FT   ft255          (symbol prefix)
 *     #define NL   4              (64-bit limbs)
 * and the instantiating file provides  FT_modulus / FT_generator  tables.
 *
 * Restates the arithmetic that `#[derive(PrimeField)]` generates for
 *   lcpc-test-fields/src/lib.rs:18-22 (Ft63), :30-34 (Ft127), :42-46 (Ft191), :54-58 (Ft255)
 * i.e. the published behaviour of the un-vendored crate ff_derive 0.12:
 *   - an element is NL little-endian u64 limbs holding the MONTGOMERY image x*R mod p, R = 2^(64*NL)
 *   - mul = Montgomery product, add/sub with one conditional correction, results always in [0,p)
 *   - to_repr = Montgomery-reduce to the canonical integer, bytes little-endian
 *     (PrimeFieldReprEndianness = "little")
 *   - Field::random = fill NL limbs from next_u64 in limb order, mask the top limb to NUM_BITS,
 *     accept iff < p; the accepted limbs ARE the Montgomery image
 *   - root_of_unity = generator^t, t = (p-1)/2^S
 * All arithmetic is exact mod p, so any correct algorithm yields the same limbs as ff_derive's.
 * PARITY UNPINNED by reference tests (the reference holds no known-answer vector);
 * pinned instead against Python big-int arithmetic and SURVEY.md App. A constants in tests/.
 */

#ifndef CAT_
#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#endif
#define FN(name) CAT(FT, name)

typedef unsigned __int128 FN(u128);

/* runtime constants, derived from modulus + generator at init */
static uint64_t FN(P)[NL];
static uint64_t FN(R)[NL];   /* R mod p  == Montgomery one */
static uint64_t FN(R2)[NL];  /* R^2 mod p */
static uint64_t FN(INV);     /* -p^{-1} mod 2^64 */
static uint64_t FN(ROU)[NL]; /* 2^S-th root of unity, Montgomery form */
static uint32_t FN(S);
static uint32_t FN(NBITS);

static inline int FN(geq_p)(const uint64_t *a) {
  for (int i = NL - 1; i >= 0; i--) {
    if (a[i] > FN(P)[i]) return 1;
    if (a[i] < FN(P)[i]) return 0;
  }
  return 1;
}

static inline int FN(is_zero)(const uint64_t *a) {
  uint64_t acc = 0;
  for (int i = 0; i < NL; i++) acc |= a[i];
  return acc == 0;
}

static inline int FN(eq)(const uint64_t *a, const uint64_t *b) {
  uint64_t acc = 0;
  for (int i = 0; i < NL; i++) acc |= a[i] ^ b[i];
  return acc == 0;
}

static inline void FN(sub_p)(uint64_t *a) {
  uint64_t borrow = 0;
  for (int i = 0; i < NL; i++) {
    FN(u128) d = (FN(u128))a[i] - FN(P)[i] - borrow;
    a[i] = (uint64_t)d;
    borrow = (uint64_t)(d >> 64) & 1;
  }
}

static inline void FN(add)(uint64_t *r, const uint64_t *a, const uint64_t *b) {
  uint64_t t[NL];
  uint64_t carry = 0;
  for (int i = 0; i < NL; i++) {
    FN(u128) s = (FN(u128))a[i] + b[i] + carry;
    t[i] = (uint64_t)s;
    carry = (uint64_t)(s >> 64);
  }
  /* 2p < 2^(64 NL) for all four fields, so carry == 0 */
  if (carry || FN(geq_p)(t)) FN(sub_p)(t);
  for (int i = 0; i < NL; i++) r[i] = t[i];
}

static inline void FN(sub)(uint64_t *r, const uint64_t *a, const uint64_t *b) {
  uint64_t t[NL];
  uint64_t borrow = 0;
  for (int i = 0; i < NL; i++) {
    FN(u128) d = (FN(u128))a[i] - b[i] - borrow;
    t[i] = (uint64_t)d;
    borrow = (uint64_t)(d >> 64) & 1;
  }
  if (borrow) {
    uint64_t carry = 0;
    for (int i = 0; i < NL; i++) {
      FN(u128) s = (FN(u128))t[i] + FN(P)[i] + carry;
      t[i] = (uint64_t)s;
      carry = (uint64_t)(s >> 64);
    }
  }
  for (int i = 0; i < NL; i++) r[i] = t[i];
}

/* Montgomery product (CIOS): r = a*b*R^{-1} mod p, r in [0,p) */
static inline void FN(mul)(uint64_t *r, const uint64_t *a, const uint64_t *b) {
  uint64_t t[NL + 2];
  for (int i = 0; i < NL + 2; i++) t[i] = 0;
  for (int i = 0; i < NL; i++) {
    FN(u128) c = 0;
    for (int j = 0; j < NL; j++) {
      c += (FN(u128))a[j] * b[i] + t[j];
      t[j] = (uint64_t)c;
      c >>= 64;
    }
    c += t[NL];
    t[NL] = (uint64_t)c;
    t[NL + 1] = (uint64_t)(c >> 64);
    uint64_t m = t[0] * FN(INV);
    c = (FN(u128))m * FN(P)[0] + t[0];
    c >>= 64;
    for (int j = 1; j < NL; j++) {
      c += (FN(u128))m * FN(P)[j] + t[j];
      t[j - 1] = (uint64_t)c;
      c >>= 64;
    }
    c += t[NL];
    t[NL - 1] = (uint64_t)c;
    t[NL] = t[NL + 1] + (uint64_t)(c >> 64);
  }
  if (t[NL] || FN(geq_p)(t)) FN(sub_p)(t);
  for (int i = 0; i < NL; i++) r[i] = t[i];
}

/* canonical integer of a Montgomery-form element (to_repr minus the byte dump) */
static inline void FN(from_mont)(uint64_t *r, const uint64_t *a) {
  uint64_t one[NL];
  for (int i = 0; i < NL; i++) one[i] = 0;
  one[0] = 1;
  FN(mul)(r, a, one);
}

static inline void FN(to_mont)(uint64_t *r, const uint64_t *a) { FN(mul)(r, a, FN(R2)); }

/* to_repr: little-endian canonical bytes (lcpc-2d/src/lib.rs:52-57 via PrimeField::to_repr) */
static inline void FN(to_repr)(uint8_t *out, const uint64_t *a) {
  uint64_t c[NL];
  FN(from_mont)(c, a);
  for (int i = 0; i < NL; i++)
    for (int k = 0; k < 8; k++) out[8 * i + k] = (uint8_t)(c[i] >> (8 * k));
}

/* r = base^e, e given as ne little-endian u64 limbs (plain integer) */
static void FN(pow)(uint64_t *r, const uint64_t *base, const uint64_t *e, int ne) {
  uint64_t acc[NL], b[NL];
  for (int i = 0; i < NL; i++) {
    acc[i] = FN(R)[i];
    b[i] = base[i];
  }
  for (int i = 0; i < ne; i++) {
    for (int k = 0; k < 64; k++) {
      if ((e[i] >> k) & 1) FN(mul)(acc, acc, b);
      FN(mul)(b, b, b);
    }
  }
  for (int i = 0; i < NL; i++) r[i] = acc[i];
}

/* r = a^{-1} = a^{p-2} */
static void FN(inv)(uint64_t *r, const uint64_t *a) {
  uint64_t e[NL];
  for (int i = 0; i < NL; i++) e[i] = FN(P)[i];
  /* p is odd and > 2: subtract 2 from the low limb (low limb of every modulus here is ...0001) */
  uint64_t borrow = 2;
  for (int i = 0; i < NL && borrow; i++) {
    uint64_t before = e[i];
    e[i] = before - borrow;
    borrow = before < borrow ? 1 : 0;
  }
  FN(pow)(r, a, e, NL);
}

/* Field::random (ff_derive): see header comment. next_u64 is the RNG word source. */
static void FN(random)(uint64_t *r, uint64_t (*next_u64)(void *), void *rng) {
  const int shave = 64 * NL - (int)FN(NBITS);
  for (;;) {
    for (int i = 0; i < NL; i++) r[i] = next_u64(rng);
    r[NL - 1] &= 0xffffffffffffffffULL >> shave;
    if (!FN(geq_p)(r)) return;
  }
}

static void FN(init)(const uint64_t *modulus, uint64_t generator) {
  for (int i = 0; i < NL; i++) FN(P)[i] = modulus[i];
  /* bit length (top limb of every modulus is non-zero) */
  FN(NBITS) = (uint32_t)(64 * NL - __builtin_clzll(FN(P)[NL - 1]));
  /* INV = -p^{-1} mod 2^64 by Newton iteration */
  uint64_t inv = 1;
  for (int i = 0; i < 6; i++) inv *= 2 - FN(P)[0] * inv;
  FN(INV) = (uint64_t)0 - inv;
  /* R = 2^(64 NL) mod p by doubling 1, 64*NL times */
  uint64_t x[NL];
  for (int i = 0; i < NL; i++) x[i] = 0;
  x[0] = 1;
  for (int k = 0; k < 64 * NL; k++) FN(add)(x, x, x);
  for (int i = 0; i < NL; i++) FN(R)[i] = x[i];
  for (int k = 0; k < 64 * NL; k++) FN(add)(x, x, x);
  for (int i = 0; i < NL; i++) FN(R2)[i] = x[i];
  /* S and t: p - 1 = 2^S * t */
  uint64_t t[NL];
  for (int i = 0; i < NL; i++) t[i] = FN(P)[i];
  t[0] -= 1; /* p odd */
  uint32_t s = 0;
  while (!(t[0] & 1)) {
    for (int i = 0; i < NL; i++) {
      uint64_t hi = (i + 1 < NL) ? t[i + 1] : 0;
      t[i] = (t[i] >> 1) | (hi << 63);
    }
    s++;
  }
  FN(S) = s;
  uint64_t g[NL], gm[NL];
  for (int i = 0; i < NL; i++) g[i] = 0;
  g[0] = generator;
  FN(to_mont)(gm, g);
  FN(pow)(FN(ROU), gm, t, NL);
}

#undef FN
