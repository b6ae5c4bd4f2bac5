// lcpc_b200/csrc/host_matgen.cpp -- host-side setup of the encodings: dimension choosers, protocol
// parameters and the seeded Brakedown code generator.  None of this is on the data-parallel path;
// it runs once per encoding (the reference builds `enc` outside its timed loops too,
// lcpc-brakedown-pc/src/bench.rs:31-35) and its output crosses the C ABI as plain CSC arrays.
//
// The random streams follow the published behaviour of the crates the reference draws from
// (rand_core 0.6 seed_from_u64, rand_chacha 0.3 ChaCha20Rng, rand 0.8 Uniform, ff 0.12 Field::random);
// none of them is vendored in the reference tree, see DESIGN.md "parity status".
#include "host_matgen.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <thread>

#include "../../include/lcpc_b200.h"
#include "../../include/lcpc_b200_host.h"
#include "host_chacha.h"

namespace lcpc {
namespace host {

namespace {

constexpr size_t LAMBDA = 128;  // lcpc-ligero-pc/src/lib.rs:45, lcpc-brakedown-pc/src/lib.rs:54

struct FieldDesc {
  int limbs;             // u64 limbs
  unsigned bits, s;      // NUM_BITS, two-adicity
  uint64_t p[4];         // modulus, little-endian limbs
};

const FieldDesc *field_desc(int field) {
  // lcpc-test-fields/src/lib.rs:19,31,43,55
  static const FieldDesc k63{1, 63, 41, {0x46d0760000000001ull, 0, 0, 0}};
  static const FieldDesc k127{2, 127, 40, {0x7f2bd90000000001ull, 0x6e754097ba20e0bfull, 0, 0}};
  static const FieldDesc k191{3, 191, 41, {0xd246820000000001ull, 0x936888270ceecbcdull, 0x453708aa3fbc8ddaull, 0}};
  static const FieldDesc k255{4, 255, 41,
                              {0x02a4f20000000001ull, 0xef73c79086595f30ull, 0xfda9df04b9575969ull, 0x663c799b6e4d2900ull}};
  switch (field) {
    case LCPC_B200_FT63: return &k63;
    case LCPC_B200_FT127: return &k127;
    case LCPC_B200_FT191: return &k191;
    case LCPC_B200_FT255: return &k255;
  }
  return nullptr;
}

size_t log2_ceil(size_t v) {  // lcpc-2d/src/lib.rs:827-829
  size_t l = 0;
  while (((size_t)1 << l) < v) l++;
  return l;
}

size_t ceil_muldiv(size_t n, size_t num, size_t den) { return (n * num + den - 1) / den; }  // matgen.rs:23-25

double ent(double z) {  // codespec.rs:17-21
  double m = 1.0 - z;
  return -z * std::log2(z) - m * std::log2(m);
}

// ff's Field::random: limbs straight from next_u64, top limb masked to NUM_BITS, accept iff < p
void random_element(const FieldDesc &fd, ChaCha20Stream &rng, uint64_t *out) {
  const uint64_t top_mask = ~(uint64_t)0 >> (64 * fd.limbs - fd.bits);
  for (;;) {
    for (int i = 0; i < fd.limbs; i++) out[i] = rng.next_u64();
    out[fd.limbs - 1] &= top_mask;
    bool less = false;
    for (int i = fd.limbs - 1; i >= 0; i--) {
      if (out[i] != fd.p[i]) {
        less = out[i] < fd.p[i];
        break;
      }
    }
    if (less) return;
  }
}

// gen_code, matgen.rs:114-188
void generate_matrix(const FieldDesc &fd, const LevelDims &dim, ChaCha20Stream &rng, CscMatrix *M) {
  const size_t L = (size_t)fd.limbs;
  M->m = dim.m, M->n = dim.n;
  M->ptrs.assign(1, 0);
  M->ptrs.reserve(dim.n + 1);
  M->idxs.reserve(dim.n * dim.d);
  M->data.reserve(dim.n * dim.d * L);
  std::vector<uint64_t> picked;
  picked.reserve(dim.d);
  std::vector<uint64_t> val(L);
  for (size_t col = 0; col < dim.n; col++) {
    picked.clear();
    while (picked.size() < dim.d) {  // sample without replacement, quadratic membership test (:144-159)
      uint64_t x = rng.below(dim.m);
      if (std::find(picked.begin(), picked.end(), x) == picked.end()) picked.push_back(x);
    }
    std::sort(picked.begin(), picked.end());
    for (uint64_t row : picked) {  // one non-zero value per picked row (:166-183)
      do {
        random_element(fd, rng, val.data());
      } while (std::all_of(val.begin(), val.end(), [](uint64_t w) { return w == 0; }));
      M->idxs.push_back(row);
      M->data.insert(M->data.end(), val.begin(), val.end());
    }
    M->ptrs.push_back(M->idxs.size());
  }
}

}  // namespace

bool sdig_code_spec(int code, CodeSpec *out) {
  static const CodeSpec specs[6] = {
      {239, 2000, 71, 2500, 71, 50, 20},  {69, 500, 111, 2500, 147, 100, 20}, {89, 500, 61, 1000, 1521, 1000, 20},
      {1, 5, 41, 500, 41, 25, 20},        {211, 1000, 97, 1000, 202, 125, 20}, {119, 500, 241, 2000, 43, 25, 20},
  };
  if (code < 1 || code > 6) return false;
  *out = specs[code - 1];
  return true;
}

unsigned field_flog2(int field) {
  const FieldDesc *fd = field_desc(field);
  return fd ? fd->bits - 1 : 0;
}
unsigned field_two_adicity(int field) {
  const FieldDesc *fd = field_desc(field);
  return fd ? fd->s : 0;
}

size_t n_degree_tests(size_t lambda, size_t len, size_t flog2) {
  size_t den = flog2 - log2_ceil(len);
  return (lambda + den - 1) / den;
}

size_t ligero_n_col_opens(size_t rho_num, size_t rho_den) {
  double rho = (double)rho_num / (double)rho_den;
  return (size_t)std::ceil(-(double)LAMBDA / std::log2((1.0 + rho) / 2.0));
}

int ligero_get_dims(int field, size_t len, size_t rho_num, size_t rho_den, size_t *n_rows, size_t *n_per_row,
                    size_t *n_cols) {
  const FieldDesc *fd = field_desc(field);
  if (!fd || len == 0 || rho_num == 0 || rho_num >= rho_den) return LCPC_B200_ERR_BAD_ARG;
  const size_t flog2 = fd->bits - 1;
  const double rho = (double)rho_num / (double)rho_den;
  const size_t opens = ligero_n_col_opens(rho_num, rho_den);
  const double lncf = (double)(opens * len);
  const double ndt = (double)n_degree_tests(LAMBDA, (size_t)std::ceil(std::sqrt(lncf) / rho), flog2);
  const size_t want = (size_t)std::ceil(std::sqrt(lncf / ndt) / rho);
  const size_t nc1 = (size_t)1 << log2_ceil(want);
  if (log2_ceil(nc1) > fd->s) return LCPC_B200_ERR_TOO_BIG;  // `?` on None at :79-85
  const size_t np1 = nc1 * rho_num / rho_den;
  const size_t nc2 = nc1 / 2, np2 = np1 / 2;
  if (np1 == 0 || np2 == 0) return LCPC_B200_ERR_BAD_ARG;
  const size_t nr1 = (len + np1 - 1) / np1, nr2 = (len + np2 - 1) / np2;
  const size_t sz1 = opens * nr1 + (1 + n_degree_tests(LAMBDA, nc1, flog2)) * np1;
  const size_t sz2 = opens * nr2 + (1 + n_degree_tests(LAMBDA, nc2, flog2)) * np2;
  if (sz1 < sz2) *n_rows = nr1, *n_per_row = np1, *n_cols = nc1;
  else *n_rows = nr2, *n_per_row = np2, *n_cols = nc2;
  return LCPC_B200_OK;
}

size_t sdig_n_col_opens(const CodeSpec &s) {
  return (size_t)std::ceil(-(double)LAMBDA / std::log2(1.0 - s.dist() / 3.0));
}

// `ml`: SdigEncodingS::new_ml (lcpc-brakedown-pc/src/lib.rs:114-124) rounds the first guess up to a power of two
int sdig_choose_n_per_row(int field, const CodeSpec &s, size_t len, size_t *n_per_row, bool ml) {
  const FieldDesc *fd = field_desc(field);
  if (!fd || len == 0) return LCPC_B200_ERR_BAD_ARG;
  const size_t flog2 = fd->bits - 1;
  const size_t opens = sdig_n_col_opens(s);
  const double lncf = (double)(opens * len);
  const double ndt = (double)n_degree_tests(LAMBDA, (size_t)std::ceil(std::sqrt(lncf)) * 2, flog2);
  size_t np1 = (size_t)std::ceil(std::sqrt(lncf / ndt));
  if (ml) {  // checked_next_power_of_two
    size_t p = 1;
    while (p < np1) p <<= 1;
    np1 = p;
  }
  if (np1 > len) np1 = len;
  const size_t np2 = np1 / 2;
  if (np2 == 0) return LCPC_B200_ERR_BAD_ARG;
  const size_t nr1 = (len + np1 - 1) / np1, nr2 = (len + np2 - 1) / np2;
  const size_t sz1 = opens * nr1 + (1 + n_degree_tests(LAMBDA, np1 * 2, flog2)) * np1;
  const size_t sz2 = opens * nr2 + (1 + n_degree_tests(LAMBDA, np2 * 2, flog2)) * np2;
  *n_per_row = sz1 < sz2 ? np1 : np2;
  return LCPC_B200_OK;
}

int sdig_level_dims(int field, const CodeSpec &s, size_t n, std::vector<LevelDims> *pre, std::vector<LevelDims> *post) {
  const FieldDesc *fd = field_desc(field);
  if (!fd || n <= s.baselen) return LCPC_B200_ERR_BAD_ARG;  // assert!(n > baselen), matgen.rs:62
  const double log2p = (double)(fd->bits - 1);
  const double alpha = s.alpha(), beta = s.beta(), r = s.r();
  const double mu = r - 1.0 - r * alpha, nu = beta + alpha * beta + 0.03;
  const double cn1 = ent(beta) + alpha * ent(1.28 * beta / alpha), cn2 = beta * std::log2(alpha / (1.28 * beta));
  const double dn1 = r * alpha * ent(beta / r) + mu * ent(nu / mu), dn2 = alpha * beta * std::log2(mu / nu);
  std::vector<size_t> sizes;
  for (size_t ni = n; ni > s.baselen; ni = ceil_muldiv(ni, s.an, s.ad)) sizes.push_back(ni);
  sizes.push_back(ceil_muldiv(sizes.back(), s.an, s.ad));
  pre->clear(), post->clear();
  for (size_t i = 0; i + 1 < sizes.size(); i++) {
    const size_t ni = sizes[i], mi = sizes[i + 1];
    size_t cn = std::min(std::max(ceil_muldiv(ni, 32 * s.bn, 25 * s.bd), 4 + ceil_muldiv(ni, s.bn, s.bd)),
                         (size_t)std::ceil((110.0 / (double)ni + cn1) / cn2));
    pre->push_back({ni, mi, std::min(cn, mi)});
    const size_t nip = ceil_muldiv(mi, s.rn, s.rd);
    const size_t mip = ceil_muldiv(ni, s.rn, s.rd) - ni - nip;
    const size_t t1 = ceil_muldiv(ni, 2 * s.bn, s.bd), t2 = ceil_muldiv(ni, s.rn, s.rd) - ni + 110;
    size_t dn = std::min(t1 + (size_t)std::ceil((double)t2 / log2p), (size_t)std::ceil((110.0 / (double)ni + dn1) / dn2));
    post->push_back({nip, mip, std::min(dn, mip)});
  }
  return LCPC_B200_OK;
}

size_t SdigCode::codeword_length() const {
  size_t len = pre.front().n + post.back().n;
  for (size_t i = 0; i + 1 < pre.size(); i++) len += pre[i].m;
  for (const auto &q : post) len += q.m;
  return len;
}

int sdig_generate(int field, int code, size_t n_per_row, uint64_t seed, SdigCode *out) {
  const FieldDesc *fd = field_desc(field);
  CodeSpec spec;
  if (!fd || !sdig_code_spec(code, &spec)) return LCPC_B200_ERR_BAD_ARG;
  std::vector<LevelDims> pre_d, post_d;
  if (int rc = sdig_level_dims(field, spec, n_per_row, &pre_d, &post_d)) return rc;
  const size_t levels = pre_d.size();
  out->field = field, out->code = code;
  out->pre.assign(levels, {}), out->post.assign(levels, {});
  // one independent ChaCha20 stream per level: precode first, then postcode (matgen.rs:38-49)
  auto work = [&](size_t i) {
    ChaCha20Stream rng = ChaCha20Stream::seed_from_u64(seed);
    rng.set_stream(i);
    generate_matrix(*fd, pre_d[i], rng, &out->pre[i]);
    generate_matrix(*fd, post_d[i], rng, &out->post[i]);
  };
  std::vector<std::thread> pool;
  for (size_t i = 1; i < levels; i++) pool.emplace_back(work, i);
  work(0);
  for (auto &t : pool) t.join();
  return LCPC_B200_OK;
}

}  // namespace host
}  // namespace lcpc

// ---- extern "C" wrappers (include/lcpc_b200_host.h) ----
using namespace lcpc::host;

struct lcpc_b200_sdig_code {
  SdigCode code;
};

extern "C" {

size_t lcpc_b200_n_degree_tests(size_t lambda, size_t len, size_t flog2) { return n_degree_tests(lambda, len, flog2); }
unsigned lcpc_b200_field_flog2(int field) { return field_flog2(field); }
size_t lcpc_b200_ligero_n_col_opens(size_t rho_num, size_t rho_den) { return ligero_n_col_opens(rho_num, rho_den); }
int lcpc_b200_ligero_get_dims(int field, size_t len, size_t rho_num, size_t rho_den, size_t *n_rows, size_t *n_per_row,
                              size_t *n_cols) {
  if (!n_rows || !n_per_row || !n_cols) return LCPC_B200_ERR_BAD_ARG;
  return ligero_get_dims(field, len, rho_num, rho_den, n_rows, n_per_row, n_cols);
}
size_t lcpc_b200_sdig_n_col_opens(int code) {
  CodeSpec s;
  return sdig_code_spec(code, &s) ? sdig_n_col_opens(s) : 0;
}
int lcpc_b200_sdig_choose_n_per_row_ml(int field, int code, size_t n_vars, size_t *n_per_row) {
  CodeSpec s;
  if (!n_per_row || n_vars >= 8 * sizeof(size_t) - 1 || !sdig_code_spec(code, &s)) return LCPC_B200_ERR_BAD_ARG;
  return sdig_choose_n_per_row(field, s, (size_t)1 << n_vars, n_per_row, true);
}
int lcpc_b200_sdig_choose_n_per_row(int field, int code, size_t len, size_t *n_per_row) {
  CodeSpec s;
  if (!n_per_row || !sdig_code_spec(code, &s)) return LCPC_B200_ERR_BAD_ARG;
  return sdig_choose_n_per_row(field, s, len, n_per_row);
}
int lcpc_b200_sdig_code_generate(int field, int code, size_t n_per_row, uint64_t seed, lcpc_b200_sdig_code **out) {
  if (!out) return LCPC_B200_ERR_BAD_ARG;
  *out = nullptr;
  auto *c = new (std::nothrow) lcpc_b200_sdig_code;
  if (!c) return LCPC_B200_ERR_OOM;
  int rc = sdig_generate(field, code, n_per_row, seed, &c->code);
  if (rc != LCPC_B200_OK) {
    delete c;
    return rc;
  }
  *out = c;
  return LCPC_B200_OK;
}
void lcpc_b200_sdig_code_free(lcpc_b200_sdig_code *c) { delete c; }
size_t lcpc_b200_sdig_code_levels(const lcpc_b200_sdig_code *c) { return c ? c->code.pre.size() : 0; }
size_t lcpc_b200_sdig_code_n_per_row(const lcpc_b200_sdig_code *c) { return c ? c->code.pre.front().n : 0; }
size_t lcpc_b200_sdig_code_codeword_length(const lcpc_b200_sdig_code *c) { return c ? c->code.codeword_length() : 0; }
int lcpc_b200_sdig_code_matrix(const lcpc_b200_sdig_code *c, size_t level, int is_post, lcpc_b200_csc *out) {
  if (!c || !out || level >= c->code.pre.size()) return LCPC_B200_ERR_BAD_ARG;
  const CscMatrix &M = is_post ? c->code.post[level] : c->code.pre[level];
  out->m = M.m, out->n = M.n;
  out->ptrs = M.ptrs.data(), out->idxs = M.idxs.data(), out->data = M.data.data();
  return LCPC_B200_OK;
}
int lcpc_b200_sdig_new_from_code(lcpc_b200_ctx *ctx, const lcpc_b200_sdig_code *c, lcpc_b200_enc **out) {
  if (!ctx || !c || !out) return LCPC_B200_ERR_BAD_ARG;
  const size_t t = c->code.pre.size();
  std::vector<lcpc_b200_csc> pre(t), post(t);
  for (size_t i = 0; i < t; i++) {
    lcpc_b200_sdig_code_matrix(c, i, 0, &pre[i]);
    lcpc_b200_sdig_code_matrix(c, i, 1, &post[i]);
  }
  return lcpc_b200_sdig_new(ctx, c->code.field, t, pre.data(), post.data(), out);
}

}  // extern "C"
