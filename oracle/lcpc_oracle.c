/*
 * oracle/lcpc_oracle.c -- TEST INFRASTRUCTURE ONLY. See lcpc_oracle.h for scope and parity status.
 */
#include "lcpc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "blake3_ref.h"
#include "chacha_rng.h"

/* parallelization limit when working on columns: lcpc-2d/src/lib.rs:619 */
#define LCPC_LOG_MIN_NCOLS 5
#define LCPC_LAMBDA 128 /* lcpc-ligero-pc/src/lib.rs:45, lcpc-brakedown-pc/src/lib.rs:54 */

/* CSC sparse matrix: m rows (outputs), n columns (inputs) -- CsMat::new_csc((m,n),..), matgen.rs:187 */
typedef struct {
  size_t m, n;
  uint64_t *ptrs; /* n+1 */
  uint64_t *idxs; /* nnz row indices, sorted within a column */
  uint64_t *data; /* nnz * NL Montgomery limbs */
} lcpc_csc;

/* codeword_length: lcpc-brakedown-pc/src/encode.rs:18-33 */
static size_t lcpc_codeword_length(const lcpc_csc *pre, const lcpc_csc *post, size_t n_levels) {
  size_t len = pre[0].n + post[n_levels - 1].n;
  for (size_t i = 0; i + 1 < n_levels; i++) len += pre[i].m;
  for (size_t i = 0; i < n_levels; i++) len += post[i].m;
  return len;
}

/* ---- field instantiations ---- */
#define FT ft63
#define NL 1
#include "field_tmpl.h"
#include "algo_tmpl.h"
#undef FT
#undef NL
#define FT ft127
#define NL 2
#include "field_tmpl.h"
#include "algo_tmpl.h"
#undef FT
#undef NL
#define FT ft191
#define NL 3
#include "field_tmpl.h"
#include "algo_tmpl.h"
#undef FT
#undef NL
#define FT ft255
#define NL 4
#include "field_tmpl.h"
#include "algo_tmpl.h"
#undef FT
#undef NL

/* moduli and generators: lcpc-test-fields/src/lib.rs:19-20, 31-32, 43-44, 55-56 (decimal there) */
static const uint64_t MOD63[1] = {0x46d0760000000001ULL};
static const uint64_t MOD127[2] = {0x7f2bd90000000001ULL, 0x6e754097ba20e0bfULL};
static const uint64_t MOD191[3] = {0xd246820000000001ULL, 0x936888270ceecbcdULL,
                                   0x453708aa3fbc8ddaULL};
static const uint64_t MOD255[4] = {0x02a4f20000000001ULL, 0xef73c79086595f30ULL,
                                   0xfda9df04b9575969ULL, 0x663c799b6e4d2900ULL};

__attribute__((constructor)) static void lcpc_oracle_init(void) {
  ft63_init(MOD63, 10);
  ft127_init(MOD127, 3);
  ft191_init(MOD191, 5);
  ft255_init(MOD255, 5);
}

#define DISPATCH(field, CALL)   \
  switch (field) {              \
    case LCPC_FT63: {           \
      enum { NLV = 1 };         \
      CALL(ft63);               \
    } break;                    \
    case LCPC_FT127: {          \
      enum { NLV = 2 };         \
      CALL(ft127);              \
    } break;                    \
    case LCPC_FT191: {          \
      enum { NLV = 3 };         \
      CALL(ft191);              \
    } break;                    \
    case LCPC_FT255: {          \
      enum { NLV = 4 };         \
      CALL(ft255);              \
    } break;                    \
    default:                    \
      return -100;              \
  }

static void set_threads(int threads) {
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
  else omp_set_num_threads(omp_get_num_procs());
#else
  (void)threads;
#endif
}

int lcpc_oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_num_procs();
#else
  return 1;
#endif
}

int lcpc_oracle_field_limbs(int field) {
  switch (field) {
    case LCPC_FT63: return 1;
    case LCPC_FT127: return 2;
    case LCPC_FT191: return 3;
    case LCPC_FT255: return 4;
    default: return -1;
  }
}

static uint32_t field_flog2(int field) { /* SizedField::FLOG2 = NUM_BITS - 1, lcpc-2d/src/lib.rs:68-71 */
  switch (field) {
    case LCPC_FT63: return ft63_NBITS - 1;
    case LCPC_FT127: return ft127_NBITS - 1;
    case LCPC_FT191: return ft191_NBITS - 1;
    default: return ft255_NBITS - 1;
  }
}

static uint32_t field_s(int field) {
  switch (field) {
    case LCPC_FT63: return ft63_S;
    case LCPC_FT127: return ft127_S;
    case LCPC_FT191: return ft191_S;
    default: return ft255_S;
  }
}

int lcpc_oracle_field_info(int field, uint32_t *num_bits, uint32_t *s, uint64_t *modulus,
                           uint64_t *r, uint64_t *r2, uint64_t *inv, uint64_t *rou) {
#define CALL(F)                                              \
  if (num_bits) *num_bits = F##_NBITS;                       \
  if (s) *s = F##_S;                                         \
  if (modulus) memcpy(modulus, F##_P, NLV * 8);              \
  if (r) memcpy(r, F##_R, NLV * 8);                          \
  if (r2) memcpy(r2, F##_R2, NLV * 8);                       \
  if (inv) *inv = F##_INV;                                   \
  if (rou) memcpy(rou, F##_ROU, NLV * 8);
  DISPATCH(field, CALL)
#undef CALL
  return 0;
}

int lcpc_oracle_field_op(int field, int op, uint64_t *r, const uint64_t *a, const uint64_t *b,
                         size_t n) {
#define CALL(F)                                                      \
  for (size_t i = 0; i < n; i++) {                                   \
    uint64_t *ri = r + i * NLV;                                      \
    const uint64_t *ai = a + i * NLV;                                \
    const uint64_t *bi = b ? b + i * NLV : NULL;                     \
    switch (op) {                                                    \
      case 0: F##_add(ri, ai, bi); break;                            \
      case 1: F##_sub(ri, ai, bi); break;                            \
      case 2: F##_mul(ri, ai, bi); break;                            \
      case 3: F##_to_mont(ri, ai); break;                            \
      case 4: F##_from_mont(ri, ai); break;                          \
      case 5: F##_inv(ri, ai); break;                                \
      default: return -101;                                          \
    }                                                                \
  }
  DISPATCH(field, CALL)
#undef CALL
  return 0;
}

int lcpc_oracle_to_repr(int field, uint8_t *out, const uint64_t *a, size_t n) {
#define CALL(F) \
  for (size_t i = 0; i < n; i++) F##_to_repr(out + i * NLV * 8, a + i * NLV);
  DISPATCH(field, CALL)
#undef CALL
  return 0;
}

int lcpc_oracle_random_elems(int field, uint64_t seed, uint64_t stream, uint64_t *out, size_t n) {
  chacha_rng rng;
  chacha_seed_from_u64(&rng, seed);
  chacha_set_stream(&rng, stream);
#define CALL(F) \
  for (size_t i = 0; i < n; i++) F##_random(out + i * NLV, chacha_next_u64, &rng);
  DISPATCH(field, CALL)
#undef CALL
  return 0;
}

/* the prover's / verifier's challenge tensors (lcpc-2d/src/lib.rs:1026-1032, 868-877):
 * ChaCha20Rng::from_seed(key), then n x Field::random */
int lcpc_oracle_random_elems_from_key(int field, const uint8_t key[32], uint64_t *out, size_t n) {
  chacha_rng rng;
  chacha_from_seed(&rng, key);
#define CALL(F) \
  for (size_t i = 0; i < n; i++) F##_random(out + i * NLV, chacha_next_u64, &rng);
  DISPATCH(field, CALL)
#undef CALL
  return 0;
}

void lcpc_oracle_blake3(const uint8_t *in, size_t len, uint8_t out[32]) { b3_hash(in, len, out); }

void lcpc_oracle_chacha_block(const uint32_t key[8], uint64_t counter, uint64_t stream,
                              uint32_t out[16]) {
  chacha_block(key, counter, stream, out);
}

/* log2: lcpc-2d/src/lib.rs:827-829 (ceil log2 via next_power_of_two) */
static size_t lcpc_log2(size_t v) {
  size_t l = 0;
  while (((size_t)1 << l) < v) l++;
  return l;
}

static size_t next_pow2(size_t v) { return (size_t)1 << lcpc_log2(v); }

static int is_pow2(size_t v) { return v && !(v & (v - 1)); }

/* ---- NTT ---- */
static int fft_dispatch(int field, uint64_t *x, size_t len, int inverse) {
  if (!is_pow2(len)) return -1; /* FFTError::NotPowerOfTwo */
  uint32_t log_len = (uint32_t)lcpc_log2(len);
  if (log_len > field_s(field)) return -2; /* FFTError::TooBig */
  if (log_len == 0) return 0;
  int nl = lcpc_oracle_field_limbs(field);
  if (nl < 0) return -100;
  uint64_t *roots = (uint64_t *)malloc((len / 2) * nl * 8);
  if (!roots) return -4;
#define CALL(F)                                    \
  F##_roots_of_unity(roots, log_len, inverse);     \
  if (inverse) F##_ifft_oi(x, log_len, roots);     \
  else F##_fft_io(x, log_len, roots);
  DISPATCH(field, CALL)
#undef CALL
  free(roots);
  return 0;
}

int lcpc_oracle_fft_io(int field, uint64_t *x, size_t len) { return fft_dispatch(field, x, len, 0); }
int lcpc_oracle_ifft_oi(int field, uint64_t *x, size_t len) { return fft_dispatch(field, x, len, 1); }

int lcpc_oracle_root_of_unity(int field, size_t len, uint64_t *w) {
  if (!is_pow2(len)) return -1;
  uint32_t log_len = (uint32_t)lcpc_log2(len);
  if (log_len > field_s(field)) return -2;
#define CALL(F)                                                    \
  memcpy(w, F##_ROU, NLV * 8);                                     \
  for (uint32_t i = 0; i < F##_S - log_len; i++) F##_mul(w, w, w);
  DISPATCH(field, CALL)
#undef CALL
  return 0;
}

/* ---- protocol parameters ---- */
size_t lcpc_oracle_n_degree_tests(size_t lambda, size_t len, size_t flog2) {
  size_t den = flog2 - lcpc_log2(len); /* lcpc-2d/src/lib.rs:613-616 */
  return (lambda + den - 1) / den;
}

size_t lcpc_oracle_ligero_n_col_opens(size_t rho_num, size_t rho_den) {
  /* lcpc-ligero-pc/src/lib.rs:61-64 */
  double rho = (double)rho_num / (double)rho_den;
  double den = log2((1.0 + rho) / 2.0);
  return (size_t)ceil(-(double)LCPC_LAMBDA / den);
}

int lcpc_oracle_ligero_get_dims(int field, size_t len, size_t rho_num, size_t rho_den,
                                size_t *n_rows, size_t *n_per_row, size_t *n_cols) {
  /* lcpc-ligero-pc/src/lib.rs:70-112 */
  if (lcpc_oracle_field_limbs(field) < 0 || rho_num >= rho_den || len == 0) return -1;
  size_t flog2 = field_flog2(field);
  double rho = (double)rho_num / (double)rho_den;
  size_t n_col_opens = lcpc_oracle_ligero_n_col_opens(rho_num, rho_den);
  double lncf = (double)(n_col_opens * len);
  double ndt =
      (double)lcpc_oracle_n_degree_tests(LCPC_LAMBDA, (size_t)ceil(sqrt(lncf) / rho), flog2);
  size_t nc1 = next_pow2((size_t)ceil(sqrt(lncf / ndt) / rho));
  if (lcpc_log2(nc1) > field_s(field)) return -2;
  size_t np1 = nc1 * rho_num / rho_den;
  size_t nr1 = (len + np1 - 1) / np1;
  size_t nd1 = lcpc_oracle_n_degree_tests(LCPC_LAMBDA, nc1, flog2);
  size_t nc2 = nc1 / 2;
  size_t np2 = np1 / 2;
  if (np2 == 0) return -3;
  size_t nr2 = (len + np2 - 1) / np2;
  size_t nd2 = lcpc_oracle_n_degree_tests(LCPC_LAMBDA, nc2, flog2);
  size_t sz1 = n_col_opens * nr1 + (1 + nd1) * np1;
  size_t sz2 = n_col_opens * nr2 + (1 + nd2) * np2;
  if (sz1 < sz2) {
    *n_rows = nr1, *n_per_row = np1, *n_cols = nc1;
  } else {
    *n_rows = nr2, *n_per_row = np2, *n_cols = nc2;
  }
  return 0;
}

/* codespec.rs:169-232: {alpha_num, alpha_den, beta_num, beta_den, r_num, r_den}; baselen 20 */
static const size_t SDIG_CODES[6][6] = {
    {239, 2000, 71, 2500, 71, 50},   {69, 500, 111, 2500, 147, 100}, {89, 500, 61, 1000, 1521, 1000},
    {1, 5, 41, 500, 41, 25},         {211, 1000, 97, 1000, 202, 125}, {119, 500, 241, 2000, 43, 25},
};
#define SDIG_BASELEN 20

typedef struct {
  size_t an, ad, bn, bd, rn, rd;
  double alpha, beta, r;
} sdig_spec;

static int sdig_spec_get(int code, sdig_spec *s) {
  if (code < 1 || code > 6) return -1;
  const size_t *c = SDIG_CODES[code - 1];
  s->an = c[0], s->ad = c[1], s->bn = c[2], s->bd = c[3], s->rn = c[4], s->rd = c[5];
  s->alpha = (double)s->an / (double)s->ad;
  s->beta = (double)s->bn / (double)s->bd;
  s->r = (double)s->rn / (double)s->rd;
  return 0;
}

static double ent(double z) { /* codespec.rs:17-21 */
  double m = 1.0 - z;
  return -z * log2(z) - m * log2(m);
}

static size_t ceil_muldiv(size_t n, size_t num, size_t den) { return (n * num + den - 1) / den; }

size_t lcpc_oracle_sdig_n_col_opens(int code) {
  /* lcpc-brakedown-pc/src/lib.rs:57-61; dist = beta / r (codespec.rs:44-47) */
  sdig_spec s;
  if (sdig_spec_get(code, &s)) return 0;
  double dist = (double)(s.bn * s.rd) / (double)(s.bd * s.rn);
  double den = log2(1.0 - dist / 3.0);
  return (size_t)ceil(-(double)LCPC_LAMBDA / den);
}

int lcpc_oracle_sdig_level_dims(int field, int code, size_t n, size_t max_levels,
                                size_t (*pre_dims)[3], size_t (*post_dims)[3]) {
  /* matgen.rs:56-111 */
  sdig_spec s;
  if (sdig_spec_get(code, &s) || lcpc_oracle_field_limbs(field) < 0) return -1;
  if (n <= SDIG_BASELEN) return -2; /* assert at :62 */
  double log2p = (double)field_flog2(field);
  double mu = s.r - 1.0 - s.r * s.alpha;                                  /* codespec.rs:100-102 */
  double nu = s.beta + s.alpha * s.beta + 0.03;                           /* :105-107 */
  double cn1 = ent(s.beta) + s.alpha * ent(1.28 * s.beta / s.alpha);      /* :110-112 */
  double cn2 = s.beta * log2(s.alpha / (1.28 * s.beta));                  /* :115-117 */
  double dn1 = s.r * s.alpha * ent(s.beta / s.r) + mu * ent(nu / mu);     /* :120-123 */
  double dn2 = s.alpha * s.beta * log2(mu / nu);                          /* :126-128 */
  /* the chain n, ceil(alpha n), ... while > baselen, plus the first value <= baselen (:66-73) */
  size_t chain[128];
  size_t nchain = 0;
  size_t ni = n;
  while (ni > SDIG_BASELEN) {
    if (nchain >= 126) return -3;
    chain[nchain++] = ni;
    ni = ceil_muldiv(ni, s.an, s.ad);
  }
  chain[nchain++] = ni;
  size_t levels = nchain - 1;
  if (levels > max_levels) return -3;
  for (size_t i = 0; i < levels; i++) {
    size_t a = chain[i], mi = chain[i + 1];
    size_t c1 = ceil_muldiv(a, 32 * s.bn, 25 * s.bd);
    size_t c2 = 4 + ceil_muldiv(a, s.bn, s.bd);
    size_t cmax = c1 > c2 ? c1 : c2;
    size_t c3 = (size_t)ceil((110.0 / (double)a + cn1) / cn2);
    size_t cn = cmax < c3 ? cmax : c3;
    if (cn > mi) cn = mi;
    pre_dims[i][0] = a, pre_dims[i][1] = mi, pre_dims[i][2] = cn;
    size_t niprime = ceil_muldiv(mi, s.rn, s.rd);
    size_t miprime = ceil_muldiv(a, s.rn, s.rd) - a - niprime;
    size_t t1 = ceil_muldiv(a, 2 * s.bn, s.bd);
    size_t t2 = ceil_muldiv(a, s.rn, s.rd) - a + 110;
    size_t d1 = t1 + (size_t)ceil((double)t2 / log2p);
    size_t d2 = (size_t)ceil((110.0 / (double)a + dn1) / dn2);
    size_t dn = d1 < d2 ? d1 : d2;
    if (dn > miprime) dn = miprime;
    post_dims[i][0] = niprime, post_dims[i][1] = miprime, post_dims[i][2] = dn;
  }
  return (int)levels;
}

/* ---- encodings ---- */
#define SDIG_MAX_LEVELS 32
struct lcpc_oracle_enc {
  int kind;
  int field;
  size_t n_per_row, n_cols;
  /* ligero */
  size_t rho_num, rho_den;
  uint64_t *roots; /* FFTPrecomp: w^0 .. w^(n_cols/2 - 1) */
  /* sdig */
  int code;
  size_t n_levels;
  lcpc_csc pre[SDIG_MAX_LEVELS], post[SDIG_MAX_LEVELS];
};

static void csc_free(lcpc_csc *m) {
  free(m->ptrs);
  free(m->idxs);
  free(m->data);
  memset(m, 0, sizeof *m);
}

void lcpc_oracle_enc_free(lcpc_oracle_enc *e) {
  if (!e) return;
  free(e->roots);
  for (size_t i = 0; i < e->n_levels; i++) {
    csc_free(&e->pre[i]);
    csc_free(&e->post[i]);
  }
  free(e);
}

static int ligero_dims_ok(size_t n_per_row, size_t n_cols) { /* lcpc-ligero-pc/src/lib.rs:114-118 */
  return n_per_row < n_cols && is_pow2(n_cols);
}

lcpc_oracle_enc *lcpc_oracle_ligero_new_from_dims(int field, size_t n_per_row, size_t n_cols,
                                                  size_t rho_num, size_t rho_den) {
  /* lcpc-ligero-pc/src/lib.rs:138-148 */
  int nl = lcpc_oracle_field_limbs(field);
  if (nl < 0 || !ligero_dims_ok(n_per_row, n_cols)) return NULL;
  uint32_t log_len = (uint32_t)lcpc_log2(n_cols);
  if (log_len > field_s(field)) return NULL; /* precomp_fft -> FFTError::TooBig */
  lcpc_oracle_enc *e = (lcpc_oracle_enc *)calloc(1, sizeof *e);
  if (!e) return NULL;
  e->kind = LCPC_ENC_LIGERO;
  e->field = field;
  e->n_per_row = n_per_row;
  e->n_cols = n_cols;
  e->rho_num = rho_num;
  e->rho_den = rho_den;
  e->roots = (uint64_t *)malloc((n_cols / 2 + 1) * nl * 8);
  if (!e->roots) {
    free(e);
    return NULL;
  }
  switch (field) {
    case LCPC_FT63: ft63_roots_of_unity(e->roots, log_len, 0); break;
    case LCPC_FT127: ft127_roots_of_unity(e->roots, log_len, 0); break;
    case LCPC_FT191: ft191_roots_of_unity(e->roots, log_len, 0); break;
    default: ft255_roots_of_unity(e->roots, log_len, 0); break;
  }
  return e;
}

lcpc_oracle_enc *lcpc_oracle_ligero_new(int field, size_t len, size_t rho_num, size_t rho_den) {
  /* lcpc-ligero-pc/src/lib.rs:121-124 */
  size_t nr, np, nc;
  if (lcpc_oracle_ligero_get_dims(field, len, rho_num, rho_den, &nr, &np, &nc)) return NULL;
  return lcpc_oracle_ligero_new_from_dims(field, np, nc, rho_num, rho_den);
}

static int sdig_generate(lcpc_oracle_enc *e, size_t n, uint64_t seed) {
  /* matgen::generate, matgen.rs:28-52: one ChaCha20 stream per level, precode then postcode */
  size_t pre_dims[SDIG_MAX_LEVELS][3], post_dims[SDIG_MAX_LEVELS][3];
  int levels = lcpc_oracle_sdig_level_dims(e->field, e->code, n, SDIG_MAX_LEVELS, pre_dims, post_dims);
  if (levels <= 0) return -1;
  e->n_levels = (size_t)levels;
  int rc = 0;
#pragma omp parallel for schedule(dynamic, 1)
  for (int i = 0; i < levels; i++) {
    chacha_rng rng;
    chacha_seed_from_u64(&rng, seed);
    chacha_set_stream(&rng, (uint64_t)i);
    int r1 = 0, r2 = 0;
    switch (e->field) {
      case LCPC_FT63:
        r1 = ft63_gen_code(&e->pre[i], pre_dims[i][0], pre_dims[i][1], pre_dims[i][2], &rng);
        r2 = ft63_gen_code(&e->post[i], post_dims[i][0], post_dims[i][1], post_dims[i][2], &rng);
        break;
      case LCPC_FT127:
        r1 = ft127_gen_code(&e->pre[i], pre_dims[i][0], pre_dims[i][1], pre_dims[i][2], &rng);
        r2 = ft127_gen_code(&e->post[i], post_dims[i][0], post_dims[i][1], post_dims[i][2], &rng);
        break;
      case LCPC_FT191:
        r1 = ft191_gen_code(&e->pre[i], pre_dims[i][0], pre_dims[i][1], pre_dims[i][2], &rng);
        r2 = ft191_gen_code(&e->post[i], post_dims[i][0], post_dims[i][1], post_dims[i][2], &rng);
        break;
      default:
        r1 = ft255_gen_code(&e->pre[i], pre_dims[i][0], pre_dims[i][1], pre_dims[i][2], &rng);
        r2 = ft255_gen_code(&e->post[i], post_dims[i][0], post_dims[i][1], post_dims[i][2], &rng);
        break;
    }
    if (r1 || r2) {
#pragma omp atomic write
      rc = -1;
    }
  }
  return rc;
}

lcpc_oracle_enc *lcpc_oracle_sdig_new_from_dims(int field, int code, size_t n_per_row, size_t n_cols,
                                                uint64_t seed) {
  /* lcpc-brakedown-pc/src/lib.rs:126-137; n_cols == 0 means "whatever codeword_length gives" */
  if (lcpc_oracle_field_limbs(field) < 0) return NULL;
  lcpc_oracle_enc *e = (lcpc_oracle_enc *)calloc(1, sizeof *e);
  if (!e) return NULL;
  e->kind = LCPC_ENC_SDIG;
  e->field = field;
  e->code = code;
  if (sdig_generate(e, n_per_row, seed)) {
    lcpc_oracle_enc_free(e);
    return NULL;
  }
  e->n_per_row = n_per_row;
  e->n_cols = lcpc_codeword_length(e->pre, e->post, e->n_levels);
  if (e->pre[0].n != n_per_row || (n_cols && n_cols != e->n_cols)) { /* asserts :128-129 */
    lcpc_oracle_enc_free(e);
    return NULL;
  }
  return e;
}

lcpc_oracle_enc *lcpc_oracle_sdig_new(int field, int code, size_t len, uint64_t seed) {
  /* SdigEncodingS::new (:103-110) + _new_from_np1 (:69-99) */
  if (lcpc_oracle_field_limbs(field) < 0 || len == 0) return NULL;
  size_t flog2 = field_flog2(field);
  size_t n_col_opens = lcpc_oracle_sdig_n_col_opens(code);
  if (!n_col_opens) return NULL;
  double lncf = (double)(n_col_opens * len);
  double ndt = (double)lcpc_oracle_n_degree_tests(LCPC_LAMBDA, (size_t)ceil(sqrt(lncf)) * 2, flog2);
  size_t np1 = (size_t)ceil(sqrt(lncf / ndt));
  if (np1 > len) np1 = len;
  size_t nr1 = (len + np1 - 1) / np1;
  size_t nd1 = lcpc_oracle_n_degree_tests(LCPC_LAMBDA, np1 * 2, flog2);
  size_t np2 = np1 / 2;
  if (np2 == 0) return NULL;
  size_t nr2 = (len + np2 - 1) / np2;
  size_t nd2 = lcpc_oracle_n_degree_tests(LCPC_LAMBDA, np2 * 2, flog2);
  size_t sz1 = n_col_opens * nr1 + (1 + nd1) * np1;
  size_t sz2 = n_col_opens * nr2 + (1 + nd2) * np2;
  size_t n_per_row = sz1 < sz2 ? np1 : np2;
  return lcpc_oracle_sdig_new_from_dims(field, code, n_per_row, 0, seed);
}

static int csc_copy(lcpc_csc *dst, size_t m, size_t n, const uint64_t *ptrs, const uint64_t *idxs,
                    const uint64_t *data, int nl) {
  size_t nnz = ptrs[n];
  dst->m = m;
  dst->n = n;
  dst->ptrs = (uint64_t *)malloc((n + 1) * 8);
  dst->idxs = (uint64_t *)malloc((nnz + 1) * 8);
  dst->data = (uint64_t *)malloc((nnz + 1) * nl * 8);
  if (!dst->ptrs || !dst->idxs || !dst->data) return -1;
  memcpy(dst->ptrs, ptrs, (n + 1) * 8);
  memcpy(dst->idxs, idxs, nnz * 8);
  memcpy(dst->data, data, nnz * nl * 8);
  return 0;
}

lcpc_oracle_enc *lcpc_oracle_sdig_from_matrices(int field, int code, size_t n_levels,
                                                const size_t *pre_m, const size_t *pre_n,
                                                const uint64_t *const *pre_ptrs,
                                                const uint64_t *const *pre_idxs,
                                                const uint64_t *const *pre_data,
                                                const size_t *post_m, const size_t *post_n,
                                                const uint64_t *const *post_ptrs,
                                                const uint64_t *const *post_idxs,
                                                const uint64_t *const *post_data) {
  int nl = lcpc_oracle_field_limbs(field);
  if (nl < 0 || n_levels == 0 || n_levels > SDIG_MAX_LEVELS) return NULL;
  lcpc_oracle_enc *e = (lcpc_oracle_enc *)calloc(1, sizeof *e);
  if (!e) return NULL;
  e->kind = LCPC_ENC_SDIG;
  e->field = field;
  e->code = code;
  e->n_levels = n_levels;
  for (size_t i = 0; i < n_levels; i++) {
    if (csc_copy(&e->pre[i], pre_m[i], pre_n[i], pre_ptrs[i], pre_idxs[i], pre_data[i], nl) ||
        csc_copy(&e->post[i], post_m[i], post_n[i], post_ptrs[i], post_idxs[i], post_data[i], nl)) {
      lcpc_oracle_enc_free(e);
      return NULL;
    }
  }
  e->n_per_row = e->pre[0].n;
  e->n_cols = lcpc_codeword_length(e->pre, e->post, n_levels);
  return e;
}

int lcpc_oracle_enc_field(const lcpc_oracle_enc *e) { return e->field; }
int lcpc_oracle_enc_kind(const lcpc_oracle_enc *e) { return e->kind; }

void lcpc_oracle_enc_get_dims(const lcpc_oracle_enc *e, size_t len, size_t *n_rows,
                              size_t *n_per_row, size_t *n_cols) {
  /* lcpc-ligero-pc/src/lib.rs:166-169, lcpc-brakedown-pc/src/lib.rs:155-158 */
  *n_rows = (len + e->n_per_row - 1) / e->n_per_row;
  *n_per_row = e->n_per_row;
  *n_cols = e->n_cols;
}

int lcpc_oracle_enc_dims_ok(const lcpc_oracle_enc *e, size_t n_per_row, size_t n_cols) {
  if (e->kind == LCPC_ENC_LIGERO) /* lcpc-ligero-pc/src/lib.rs:171-177 */
    return ligero_dims_ok(n_per_row, n_cols) && n_per_row == e->n_per_row && n_cols == e->n_cols;
  /* lcpc-brakedown-pc/src/lib.rs:160-167 */
  return n_per_row < n_cols && n_per_row == e->n_per_row && n_per_row == e->pre[0].n &&
         n_cols == e->n_cols && n_cols == lcpc_codeword_length(e->pre, e->post, e->n_levels);
}

size_t lcpc_oracle_enc_n_col_opens(const lcpc_oracle_enc *e) {
  return e->kind == LCPC_ENC_LIGERO ? lcpc_oracle_ligero_n_col_opens(e->rho_num, e->rho_den)
                                    : lcpc_oracle_sdig_n_col_opens(e->code);
}

size_t lcpc_oracle_enc_n_degree_tests(const lcpc_oracle_enc *e) {
  /* lcpc-ligero-pc/src/lib.rs:183-185, lcpc-brakedown-pc/src/lib.rs:173-175 */
  return lcpc_oracle_n_degree_tests(LCPC_LAMBDA, e->n_cols, field_flog2(e->field));
}

size_t lcpc_oracle_sdig_n_levels(const lcpc_oracle_enc *e) { return e->n_levels; }

int lcpc_oracle_sdig_matrix(const lcpc_oracle_enc *e, size_t level, int is_post, size_t *m,
                            size_t *n, size_t *nnz, const uint64_t **ptrs, const uint64_t **idxs,
                            const uint64_t **data) {
  if (e->kind != LCPC_ENC_SDIG || level >= e->n_levels) return -1;
  const lcpc_csc *M = is_post ? &e->post[level] : &e->pre[level];
  *m = M->m, *n = M->n, *nnz = M->ptrs[M->n];
  *ptrs = M->ptrs, *idxs = M->idxs, *data = M->data;
  return 0;
}

int lcpc_oracle_encode(const lcpc_oracle_enc *e, uint64_t *row) {
  if (e->kind == LCPC_ENC_LIGERO) {
    uint32_t log_len = (uint32_t)lcpc_log2(e->n_cols);
    switch (e->field) {
      case LCPC_FT63: ft63_fft_io(row, log_len, e->roots); break;
      case LCPC_FT127: ft127_fft_io(row, log_len, e->roots); break;
      case LCPC_FT191: ft191_fft_io(row, log_len, e->roots); break;
      default: ft255_fft_io(row, log_len, e->roots); break;
    }
    return 0;
  }
  switch (e->field) {
    case LCPC_FT63: return ft63_sdig_encode(row, e->n_cols, e->pre, e->post, e->n_levels);
    case LCPC_FT127: return ft127_sdig_encode(row, e->n_cols, e->pre, e->post, e->n_levels);
    case LCPC_FT191: return ft191_sdig_encode(row, e->n_cols, e->pre, e->post, e->n_levels);
    default: return ft255_sdig_encode(row, e->n_cols, e->pre, e->post, e->n_levels);
  }
}

/* ---- Merkle ---- */
/* merkle_layer base case: lcpc-2d/src/lib.rs:768-775 -- node = D(left || right) */
static void merkle_layer(const uint8_t *ins, uint8_t *outs, size_t n_out) {
#pragma omp parallel for schedule(static) if (n_out > 64)
  for (size_t i = 0; i < n_out; i++) b3_hash(ins + 64 * i, 64, outs + 32 * i);
}

void lcpc_oracle_merkle_tree(uint8_t *hashes, size_t np2, int threads) {
  /* merkle_tree: lcpc-2d/src/lib.rs:747-760 over hashes = [leaves | layer 1 | ... | root] */
  set_threads(threads);
  uint8_t *ins = hashes;
  size_t n_in = np2;
  while (n_in > 1) {
    uint8_t *outs = ins + 32 * n_in;
    merkle_layer(ins, outs, n_in / 2);
    ins = outs;
    n_in /= 2;
  }
}

int lcpc_oracle_merkleize(int field, const uint64_t *comm, size_t n_rows, size_t n_cols,
                          uint8_t *hashes, int serial, int threads) {
  size_t np2 = next_pow2(n_cols);
  set_threads(threads);
  /* hashes beyond n_cols stay Output::default() = zeros (lcpc-2d/src/lib.rs:665,696) */
  memset(hashes, 0, 32 * (2 * np2 - 1));
  if (serial) {
    /* merkleize_ser: lcpc-2d/src/lib.rs:1128-1158 */
    static const uint8_t zeros[32] = {0};
    int nl = lcpc_oracle_field_limbs(field);
    if (nl < 0) return -100;
    for (size_t col = 0; col < n_cols; col++) {
      b3_hasher d;
      b3_init(&d);
      b3_update(&d, zeros, 32);
      for (size_t row = 0; row < n_rows; row++) {
        uint8_t repr[32];
        lcpc_oracle_to_repr(field, repr, comm + (row * n_cols + col) * nl, 1);
        b3_update(&d, repr, 8 * (size_t)nl);
      }
      b3_finalize(&d, hashes + 32 * col);
    }
    uint8_t *ins = hashes;
    size_t n_in = np2;
    while (n_in > 1) {
      uint8_t *outs = ins + 32 * n_in;
      for (size_t i = 0; i < n_in / 2; i++) {
        b3_hasher d;
        b3_init(&d);
        b3_update(&d, ins + 64 * i, 32);
        b3_update(&d, ins + 64 * i + 32, 32);
        b3_finalize(&d, outs + 32 * i);
      }
      ins = outs;
      n_in /= 2;
    }
    return 0;
  }
  /* merkleize: lcpc-2d/src/lib.rs:690-704 */
#define CALL(F)                                            \
  _Pragma("omp parallel") _Pragma("omp single")            \
      F##_hash_columns(comm, hashes, n_rows, n_cols, 0, n_cols);
  DISPATCH(field, CALL)
#undef CALL
  lcpc_oracle_merkle_tree(hashes, np2, threads);
  return 0;
}

/* ---- commit: lcpc-2d/src/lib.rs:622-671 ---- */
int lcpc_oracle_commit(const lcpc_oracle_enc *e, const uint64_t *coeffs_in, size_t len,
                       uint64_t *comm, uint64_t *coeffs, uint8_t *hashes, int threads) {
  size_t n_rows, n_per_row, n_cols;
  lcpc_oracle_enc_get_dims(e, len, &n_rows, &n_per_row, &n_cols);
  if (len == 0) return -2;
  /* asserts :630-632 */
  if (!(n_rows * n_per_row >= len) || !((n_rows - 1) * n_per_row < len) ||
      !lcpc_oracle_enc_dims_ok(e, n_per_row, n_cols))
    return -2;
  size_t nl = (size_t)lcpc_oracle_field_limbs(e->field);
  set_threads(threads);
  /* :636-645 zero-init + padded copy */
  memset(coeffs, 0, n_rows * n_per_row * nl * 8);
  memset(comm, 0, n_rows * n_cols * nl * 8);
  memcpy(coeffs, coeffs_in, len * nl * 8);
  /* :648-653 per-row copy + encode */
  int rc = 0;
#pragma omp parallel for schedule(dynamic, 1)
  for (size_t r = 0; r < n_rows; r++) {
    uint64_t *row = comm + r * n_cols * nl;
    memcpy(row, coeffs + r * n_per_row * nl, n_per_row * nl * 8);
    int err = lcpc_oracle_encode(e, row);
    if (err) {
#pragma omp atomic write
      rc = -3;
    }
  }
  if (rc) return rc;
  /* :656-668 */
  return lcpc_oracle_merkleize(e->field, comm, n_rows, n_cols, hashes, 0, threads);
}

/* ---- prove pieces ---- */
int lcpc_oracle_collapse(int field, const uint64_t *coeffs, const uint64_t *tensor, uint64_t *poly,
                         size_t n_rows, size_t n_per_row, int serial, int threads) {
  int nl = lcpc_oracle_field_limbs(field);
  if (nl < 0) return -100;
  set_threads(threads);
  memset(poly, 0, n_per_row * (size_t)nl * 8);
#define CALL(F)                                                              \
  if (serial) F##_collapse_columns_ser(coeffs, tensor, poly, n_rows, n_per_row); \
  else {                                                                     \
    _Pragma("omp parallel") _Pragma("omp single")                            \
        F##_collapse_columns(coeffs, tensor, poly, n_rows, n_per_row, 0, n_per_row); \
  }
  DISPATCH(field, CALL)
#undef CALL
  return 0;
}

int lcpc_oracle_open_column(int field, const uint64_t *comm, const uint8_t *hashes, size_t n_rows,
                            size_t n_cols, size_t column, uint64_t *col_out, uint8_t *path_out) {
  /* lcpc-2d/src/lib.rs:788-825 */
  int nl = lcpc_oracle_field_limbs(field);
  if (nl < 0) return -100;
  if (column >= n_cols) return -1;
  for (size_t r = 0; r < n_rows; r++)
    memcpy(col_out + r * nl, comm + (r * n_cols + column) * nl, (size_t)nl * 8);
  size_t path_len = lcpc_log2(n_cols);
  size_t layer_len = 2 * next_pow2(n_cols) - 1; /* hashes.len() */
  const uint8_t *layer = hashes;
  for (size_t i = 0; i < path_len; i++) {
    size_t other = column ^ 1;
    memcpy(path_out + 32 * i, layer + 32 * other, 32);
    size_t skip = (layer_len + 1) / 2; /* split_at((len+1)/2), :818 */
    layer += 32 * skip;
    layer_len -= skip;
    column >>= 1;
  }
  return (int)path_len;
}

int lcpc_oracle_verify_column_path(int field, const uint64_t *col, size_t n_rows,
                                   const uint8_t *path, size_t path_len, size_t col_num,
                                   const uint8_t root[32]) {
  /* lcpc-2d/src/lib.rs:955-982 */
  static const uint8_t zeros[32] = {0};
  int nl = lcpc_oracle_field_limbs(field);
  if (nl < 0) return -100;
  b3_hasher d;
  b3_init(&d);
  b3_update(&d, zeros, 32);
  for (size_t r = 0; r < n_rows; r++) {
    uint8_t repr[32];
    lcpc_oracle_to_repr(field, repr, col + r * nl, 1);
    b3_update(&d, repr, 8 * (size_t)nl);
  }
  uint8_t hash[32];
  b3_finalize(&d, hash);
  for (size_t i = 0; i < path_len; i++) {
    b3_init(&d);
    if (col_num % 2 == 0) {
      b3_update(&d, hash, 32);
      b3_update(&d, path + 32 * i, 32);
    } else {
      b3_update(&d, path + 32 * i, 32);
      b3_update(&d, hash, 32);
    }
    b3_finalize(&d, hash);
    col_num >>= 1;
  }
  return memcmp(hash, root, 32) == 0;
}

int lcpc_oracle_dot(int field, const uint64_t *a, const uint64_t *b, size_t n, uint64_t *out) {
#define CALL(F)                                  \
  uint64_t acc[NLV], prod[NLV];                  \
  memset(acc, 0, sizeof acc);                    \
  for (size_t i = 0; i < n; i++) {               \
    F##_mul(prod, a + i * NLV, b + i * NLV);     \
    F##_add(acc, acc, prod);                     \
  }                                              \
  memcpy(out, acc, sizeof acc);
  DISPATCH(field, CALL)
#undef CALL
  return 0;
}
