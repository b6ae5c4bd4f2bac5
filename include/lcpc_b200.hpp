// include/lcpc_b200.hpp -- C++ host-side mirror of the reference's operator interface over the C ABI of
// include/lcpc_b200.h (+ lcpc_b200_host.h).  Header-only, C++17, no CUDA or torch types.
//
// The reference is compiled Rust; its toolchain is absent from the build image, so this is the compiled-language
// host layer a service would use directly (the Rust shim of INTEGRATION.md binds the same C symbols).  Names,
// argument meaning and error behaviour follow the reference (paths relative to the reference repository):
//
//   trait LcEncoding                      lcpc-2d/src/lib.rs:74-104       -> class LcEncoding
//   LigeroEncodingRho::new/new_ml/new_from_dims   lcpc-ligero-pc/src/lib.rs:121-148  -> LigeroEncoding::create*/...
//   SdigEncodingS::new/new_ml/new_from_dims       lcpc-brakedown-pc/src/lib.rs:103-137 -> SdigEncoding::create*/...
//   LcCommit::commit/get_root/prove/get_n_*       lcpc-2d/src/lib.rs:276-311      -> class LcCommit
//   LcRoot                                lcpc-2d/src/lib.rs:315-350      -> struct LcRoot
//   LcEvalProof::verify/get_n_cols/get_n_per_row  lcpc-2d/src/lib.rs:490-527      -> struct LcEvalProof
//   merlin::Transcript                    (merlin 2.0)                    -> class Transcript
//   ProverError / VerifierError           lcpc-2d/src/lib.rs:109-170      -> class Error (code() = the C status)
//
// Field elements are `uint64_t` limbs, L per element, Montgomery form, little-endian limb order: the in-memory image
// of the reference's `struct FtNNN([u64; L])`.  Nothing computes on the host: every method is one or two C-ABI calls,
// and without a CUDA device the constructors throw Error(LCPC_B200_ERR_CUDA).
#ifndef LCPC_B200_HPP
#define LCPC_B200_HPP

#include <array>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "lcpc_b200.h"
#include "lcpc_b200_host.h"

namespace lcpc_b200 {

enum class Field : int { Ft63 = LCPC_B200_FT63, Ft127 = LCPC_B200_FT127, Ft191 = LCPC_B200_FT191, Ft255 = LCPC_B200_FT255 };

/// A non-zero C status.  ProverError: TooBig = ERR_TOO_BIG, Encode = ERR_ENCODE, Commit = ERR_BAD_ARG, ColumnNumber =
/// ERR_COLUMN, OuterTensor = ERR_OUTER_TENSOR; VerifierError: the LCPC_B200_VERR_* codes.
class Error : public std::runtime_error {
 public:
  Error(int code, const std::string &what) : std::runtime_error("lcpc_b200: status " + std::to_string(code) + " " + what), code_(code) {}
  int code() const { return code_; }

 private:
  int code_;
};

namespace detail {
inline void check(int rc, const lcpc_b200_ctx *ctx = nullptr) {
  if (rc != LCPC_B200_OK) throw Error(rc, ctx ? lcpc_b200_last_error(ctx) : "");
}
inline size_t limbs(Field f) {
  int l = lcpc_b200_field_limbs(static_cast<int>(f));
  if (l < 0) throw Error(LCPC_B200_ERR_BAD_ARG, "unknown field");
  return static_cast<size_t>(l);
}
}  // namespace detail

/// One CUDA device + stream.  Shared ownership: encodings and commits keep their context alive.
class Context {
 public:
  explicit Context(int device = 0) {
    lcpc_b200_ctx *h = nullptr;
    detail::check(lcpc_b200_ctx_create(device, &h));
    h_.reset(h, lcpc_b200_ctx_destroy);
  }
  lcpc_b200_ctx *get() const { return h_.get(); }
  void synchronize() const { detail::check(lcpc_b200_ctx_synchronize(h_.get()), h_.get()); }
  uint64_t launch_count() const { return lcpc_b200_ctx_launch_count(h_.get()); }

 private:
  std::shared_ptr<lcpc_b200_ctx> h_;
};

/// merlin::Transcript (host side; sequential by construction).
class Transcript {
 public:
  explicit Transcript(const std::string &label) {
    lcpc_b200_transcript *h = nullptr;
    detail::check(lcpc_b200_transcript_new(reinterpret_cast<const uint8_t *>(label.data()), label.size(), &h));
    h_.reset(h, lcpc_b200_transcript_free);
  }
  void append_message(const std::string &label, const uint8_t *msg, size_t n) {
    detail::check(lcpc_b200_transcript_append_message(h_.get(), reinterpret_cast<const uint8_t *>(label.data()), label.size(), msg, n));
  }
  void append_message(const std::string &label, const std::string &msg) {
    append_message(label, reinterpret_cast<const uint8_t *>(msg.data()), msg.size());
  }
  void append_u64(const std::string &label, uint64_t x) {
    detail::check(lcpc_b200_transcript_append_u64(h_.get(), reinterpret_cast<const uint8_t *>(label.data()), label.size(), x));
  }
  std::vector<uint8_t> challenge_bytes(const std::string &label, size_t n) {
    std::vector<uint8_t> out(n);
    detail::check(lcpc_b200_transcript_challenge_bytes(h_.get(), reinterpret_cast<const uint8_t *>(label.data()), label.size(),
                                                       out.data(), n));
    return out;
  }
  lcpc_b200_transcript *get() const { return h_.get(); }

 private:
  std::shared_ptr<lcpc_b200_transcript> h_;
};

/// LcRoot<D, E>: the Merkle root (D = BLAKE3).
struct LcRoot {
  std::array<uint8_t, 32> root{};
  const std::array<uint8_t, 32> &as_ref() const { return root; }
  std::array<uint8_t, 32> into_raw() const { return root; }
  bool operator==(const LcRoot &o) const { return root == o.root; }
};

class LcCommit;
struct LcEvalProof;

/// trait LcEncoding over a device-side encoding object.
class LcEncoding {
 public:
  virtual ~LcEncoding() = default;
  Field field() const { return field_; }
  size_t limbs() const { return detail::limbs(field_); }
  size_t n_per_row() const { return n_per_row_; }
  size_t n_cols() const { return n_cols_; }
  /// LcEncoding::get_dims (:94)
  std::array<size_t, 3> get_dims(size_t len) const {
    size_t a = 0, b = 0, c = 0;
    detail::check(lcpc_b200_enc_get_dims(h_.get(), len, &a, &b, &c), ctx_.get());
    return {a, b, c};
  }
  /// LcEncoding::dims_ok (:97)
  bool dims_ok(size_t n_per_row, size_t n_cols) const { return lcpc_b200_enc_dims_ok(h_.get(), n_per_row, n_cols) == 1; }
  /// LcEncoding::get_n_col_opens (:100)
  virtual size_t get_n_col_opens() const = 0;
  /// LcEncoding::get_n_degree_tests (:103); LAMBDA = 128 for both reference encodings
  size_t get_n_degree_tests() const {
    return lcpc_b200_n_degree_tests(128, n_cols_, lcpc_b200_field_flog2(static_cast<int>(field_)));
  }
  /// LcEncoding::encode (:91), batched and in place: n_rows rows of n_cols elements each
  void encode(uint64_t *rows, size_t n_rows) const { detail::check(lcpc_b200_encode(h_.get(), rows, n_rows), ctx_.get()); }
  lcpc_b200_enc *get() const { return h_.get(); }
  const Context &context() const { return ctx_; }

 protected:
  LcEncoding(Context ctx, lcpc_b200_enc *h, Field f) : ctx_(std::move(ctx)), field_(f) {
    h_.reset(h, lcpc_b200_enc_free);
    auto d = get_dims(1);
    n_per_row_ = d[1], n_cols_ = d[2];
  }
  Context ctx_;
  std::shared_ptr<lcpc_b200_enc> h_;
  Field field_;
  size_t n_per_row_ = 0, n_cols_ = 0;
};

/// LigeroEncodingRho<Ft, Rn, Rd>; rho defaults to 1/2 like the alias LigeroEncoding<F> (lcpc-ligero-pc/src/lib.rs:189).
class LigeroEncoding : public LcEncoding {
 public:
  /// LigeroEncodingRho::new (:121-124)
  static LigeroEncoding create(const Context &ctx, Field f, size_t len, size_t rho_num = 1, size_t rho_den = 2) {
    size_t nr = 0, npr = 0, nc = 0;
    detail::check(lcpc_b200_ligero_get_dims(static_cast<int>(f), len, rho_num, rho_den, &nr, &npr, &nc));
    return new_from_dims(ctx, f, npr, nc, rho_num, rho_den);
  }
  /// LigeroEncodingRho::new_ml (:128-135)
  static LigeroEncoding new_ml(const Context &ctx, Field f, size_t n_vars, size_t rho_num = 1, size_t rho_den = 2) {
    const size_t n = static_cast<size_t>(1) << n_vars;
    size_t nr = 0, npr = 0, nc = 0;
    detail::check(lcpc_b200_ligero_get_dims(static_cast<int>(f), n, rho_num, rho_den, &nr, &npr, &nc));
    if ((nr & (nr - 1)) || (npr & (npr - 1)) || nr * npr != n) throw Error(LCPC_B200_ERR_BAD_ARG, "new_ml: dimensions are not powers of two");
    return new_from_dims(ctx, f, npr, nc, rho_num, rho_den);
  }
  /// LigeroEncodingRho::new_from_dims (:138-148)
  static LigeroEncoding new_from_dims(const Context &ctx, Field f, size_t n_per_row, size_t n_cols, size_t rho_num = 1,
                                      size_t rho_den = 2) {
    lcpc_b200_enc *h = nullptr;
    detail::check(lcpc_b200_ligero_new(ctx.get(), static_cast<int>(f), n_per_row, n_cols, &h), ctx.get());
    return LigeroEncoding(ctx, h, f, rho_num, rho_den);
  }
  size_t get_n_col_opens() const override { return lcpc_b200_ligero_n_col_opens(rho_num_, rho_den_); }

 private:
  LigeroEncoding(const Context &ctx, lcpc_b200_enc *h, Field f, size_t rn, size_t rd) : LcEncoding(ctx, h, f), rho_num_(rn), rho_den_(rd) {}
  size_t rho_num_, rho_den_;
};

/// SdigEncodingS<Ft, S>; code 3 = SdigCode3, the default alias SdigEncoding<F> (lcpc-brakedown-pc/src/lib.rs:19,179).
class SdigEncoding : public LcEncoding {
 public:
  /// SdigEncodingS::new (:103-110)
  static SdigEncoding create(const Context &ctx, Field f, size_t len, uint64_t seed, int code = 3) {
    size_t npr = 0;
    detail::check(lcpc_b200_sdig_choose_n_per_row(static_cast<int>(f), code, len, &npr));
    return new_from_dims(ctx, f, npr, 0, seed, code);
  }
  /// SdigEncodingS::new_ml (:114-124)
  static SdigEncoding new_ml(const Context &ctx, Field f, size_t n_vars, uint64_t seed, int code = 3) {
    size_t npr = 0;
    detail::check(lcpc_b200_sdig_choose_n_per_row_ml(static_cast<int>(f), code, n_vars, &npr));
    return new_from_dims(ctx, f, npr, 0, seed, code);
  }
  /// SdigEncodingS::new_from_dims (:126-137); n_cols = 0 skips the codeword-length assert
  static SdigEncoding new_from_dims(const Context &ctx, Field f, size_t n_per_row, size_t n_cols, uint64_t seed, int code = 3) {
    lcpc_b200_sdig_code *c = nullptr;
    detail::check(lcpc_b200_sdig_code_generate(static_cast<int>(f), code, n_per_row, seed, &c));
    std::unique_ptr<lcpc_b200_sdig_code, void (*)(lcpc_b200_sdig_code *)> guard(c, lcpc_b200_sdig_code_free);
    if (n_cols && lcpc_b200_sdig_code_codeword_length(c) != n_cols) throw Error(LCPC_B200_ERR_BAD_ARG, "codeword length != n_cols");
    lcpc_b200_enc *h = nullptr;
    detail::check(lcpc_b200_sdig_new_from_code(ctx.get(), c, &h), ctx.get());
    return SdigEncoding(ctx, h, f, code);
  }
  size_t get_n_col_opens() const override { return lcpc_b200_sdig_n_col_opens(code_); }

 private:
  SdigEncoding(const Context &ctx, lcpc_b200_enc *h, Field f, int code) : LcEncoding(ctx, h, f), code_(code) {}
  int code_;
};

/// LcEvalProof<D, E> (:490-500) as flat arrays: columns[i] = LcColumn{col: cols[i*n_rows..], path: paths[i*path_len*32..]}.
struct LcEvalProof {
  size_t n_cols = 0, n_per_row = 0, n_degree_tests = 0, n_columns = 0, n_rows = 0, path_len = 0;
  std::vector<uint64_t> p_eval, p_random_vec, cols;
  std::vector<uint8_t> paths;
  std::vector<uint64_t> col_idx;  // the prover's view of the opened column numbers (not part of the proof)

  size_t get_n_cols() const { return n_cols; }        // :507-509
  size_t get_n_per_row() const { return n_per_row; }  // :512-514

  /// LcEvalProof::verify (:518-527): returns the evaluation (L limbs) or throws Error carrying a VerifierError code.
  std::vector<uint64_t> verify(const LcRoot &root, const uint64_t *outer_tensor, size_t outer_len, const uint64_t *inner_tensor,
                               size_t inner_len, const LcEncoding &enc, Transcript &tr) const {
    lcpc_b200_proof pf{n_cols, n_per_row, n_degree_tests, n_columns, n_rows, path_len,
                       p_eval.data(), p_random_vec.data(), cols.data(), paths.data()};
    std::vector<uint64_t> eval(enc.limbs());
    detail::check(lcpc_b200_verify(enc.get(), tr.get(), nullptr, root.root.data(), outer_tensor, outer_len, inner_tensor, inner_len,
                                   enc.get_n_col_opens(), enc.get_n_degree_tests(), &pf, eval.data()),
                  enc.context().get());
    return eval;
  }
};

// ---- wire format: bincode 1.x (default options) over the reference's Wrapped* structs (:186-197, :353-357, :430-437,
// :551-560): usize and Vec lengths are little-endian u64, a field element is its L Montgomery limbs (a newtype around
// [u64; L], no length), a digest is WrappedOutput { #[serde(with = "serde_bytes")] bytes } = u64 length + raw bytes ----
namespace detail {
inline void put_u64(std::vector<uint8_t> &out, uint64_t v) {
  for (int i = 0; i < 8; i++) out.push_back(static_cast<uint8_t>(v >> (8 * i)));
}
inline void put_elems(std::vector<uint8_t> &out, const uint64_t *e, size_t n_elems, size_t L) {
  put_u64(out, n_elems);
  for (size_t i = 0; i < n_elems * L; i++) put_u64(out, e[i]);
}
struct Reader {
  const uint8_t *p;
  size_t n, o = 0;
  uint64_t u64() {
    if (o + 8 > n) throw Error(LCPC_B200_ERR_BAD_ARG, "wire: truncated");
    uint64_t v = 0;
    for (int i = 0; i < 8; i++) v |= static_cast<uint64_t>(p[o + i]) << (8 * i);
    o += 8;
    return v;
  }
  void elems(std::vector<uint64_t> &dst, size_t L, size_t *count) {
    const uint64_t k = u64();
    if (k > (n - o) / (8 * L)) throw Error(LCPC_B200_ERR_BAD_ARG, "wire: truncated");
    for (uint64_t i = 0; i < k * L; i++) dst.push_back(u64());
    *count = static_cast<size_t>(k);
  }
  void digest(std::vector<uint8_t> &dst) {
    if (u64() != 32 || o + 32 > n) throw Error(LCPC_B200_ERR_BAD_ARG, "wire: digest is not 32 bytes");
    dst.insert(dst.end(), p + o, p + o + 32);
    o += 32;
  }
};
}  // namespace detail

/// bincode::serialize(&LcRoot)
inline std::vector<uint8_t> serialize(const LcRoot &r) {
  std::vector<uint8_t> out;
  detail::put_u64(out, 32);
  out.insert(out.end(), r.root.begin(), r.root.end());
  return out;
}

/// bincode::serialize(&LcEvalProof): {n_cols, p_eval, p_random_vec, columns[{col, path}]}; L = limbs per element
inline std::vector<uint8_t> serialize(const LcEvalProof &p, size_t L) {
  std::vector<uint8_t> out;
  detail::put_u64(out, p.n_cols);
  detail::put_elems(out, p.p_eval.data(), p.n_per_row, L);
  detail::put_u64(out, p.n_degree_tests);
  for (size_t i = 0; i < p.n_degree_tests; i++) detail::put_elems(out, p.p_random_vec.data() + i * p.n_per_row * L, p.n_per_row, L);
  detail::put_u64(out, p.n_columns);
  for (size_t j = 0; j < p.n_columns; j++) {
    detail::put_elems(out, p.cols.data() + j * p.n_rows * L, p.n_rows, L);
    detail::put_u64(out, p.path_len);
    for (size_t l = 0; l < p.path_len; l++) {
      detail::put_u64(out, 32);
      const uint8_t *d = p.paths.data() + (j * p.path_len + l) * 32;
      out.insert(out.end(), d, d + 32);
    }
  }
  return out;
}

/// bincode::deserialize::<LcEvalProof>; throws Error(ERR_BAD_ARG) on truncated, ragged or trailing input
inline LcEvalProof deserialize_proof(const uint8_t *data, size_t n, size_t L) {
  detail::Reader r{data, n};
  LcEvalProof p;
  p.n_cols = static_cast<size_t>(r.u64());
  r.elems(p.p_eval, L, &p.n_per_row);
  p.n_degree_tests = static_cast<size_t>(r.u64());
  for (size_t i = 0; i < p.n_degree_tests; i++) {
    size_t k = 0;
    r.elems(p.p_random_vec, L, &k);
    if (k != p.n_per_row) throw Error(LCPC_B200_ERR_BAD_ARG, "wire: ragged proof");
  }
  p.n_columns = static_cast<size_t>(r.u64());
  for (size_t j = 0; j < p.n_columns; j++) {
    size_t k = 0;
    r.elems(p.cols, L, &k);
    const size_t pl = static_cast<size_t>(r.u64());
    if (j == 0) p.n_rows = k, p.path_len = pl;
    if (k != p.n_rows || pl != p.path_len) throw Error(LCPC_B200_ERR_BAD_ARG, "wire: ragged proof");
    for (size_t l = 0; l < pl; l++) r.digest(p.paths);
  }
  if (r.o != n) throw Error(LCPC_B200_ERR_BAD_ARG, "wire: trailing bytes");
  return p;
}

/// LcCommit<D, E> (:172-184), device resident; the fields are downloaded on request.
class LcCommit {
 public:
  /// LcCommit::commit (:299-301)
  static LcCommit commit(const uint64_t *coeffs, size_t len, const LcEncoding &enc) {
    lcpc_b200_commit *h = nullptr;
    detail::check(lcpc_b200_commit_new(enc.get(), coeffs, len, &h), enc.context().get());
    return LcCommit(enc, h);
  }
  /// Deserialize for LcCommit (:256-268): from host-side fields, sizes checked like check_comm (:672-688)
  static LcCommit from_fields(const LcEncoding &enc, const uint64_t *comm, size_t comm_len, const uint64_t *coeffs, size_t coeffs_len,
                              const uint8_t *hashes, size_t n_hashes, size_t n_rows) {
    lcpc_b200_commit *h = nullptr;
    detail::check(lcpc_b200_commit_from_host(enc.get(), comm, comm_len, coeffs, coeffs_len, hashes, n_hashes, n_rows, &h),
                  enc.context().get());
    return LcCommit(enc, h);
  }
  size_t get_n_rows() const { return n_rows_; }        // :294-296
  size_t get_n_cols() const { return n_cols_; }        // :289-291
  size_t get_n_per_row() const { return n_per_row_; }  // :284-286
  size_t n_hashes() const { return n_hashes_; }
  /// LcCommit::get_root (:276-281)
  LcRoot get_root() const {
    LcRoot r;
    detail::check(lcpc_b200_commit_root(h_.get(), r.root.data()), ctx_.get());
    return r;
  }
  /// the struct's fields (:178-183) copied to the host; any pointer may be null
  void download(uint64_t *comm, uint64_t *coeffs, uint8_t *hashes) const {
    detail::check(lcpc_b200_commit_download(h_.get(), comm, coeffs, hashes), ctx_.get());
  }
  /// collapse_columns (:1095-1123): poly (n_per_row elements) = sum_r tensor[r] * coeffs[r]
  std::vector<uint64_t> collapse(const uint64_t *tensor) const {
    std::vector<uint64_t> poly(n_per_row_ * limbs_);
    detail::check(lcpc_b200_commit_collapse(h_.get(), tensor, poly.data()), ctx_.get());
    return poly;
  }
  /// LcCommit::prove (:304-311)
  LcEvalProof prove(const uint64_t *outer_tensor, size_t outer_len, const LcEncoding &enc, Transcript &tr) const {
    LcEvalProof pf;
    pf.n_cols = n_cols_, pf.n_per_row = n_per_row_, pf.n_rows = n_rows_;
    pf.n_degree_tests = enc.get_n_degree_tests(), pf.n_columns = enc.get_n_col_opens();
    for (pf.path_len = 0; (static_cast<size_t>(1) << pf.path_len) < n_cols_;) pf.path_len++;
    pf.p_eval.resize(n_per_row_ * limbs_);
    pf.p_random_vec.resize(pf.n_degree_tests * n_per_row_ * limbs_);
    pf.col_idx.resize(pf.n_columns);
    pf.cols.resize(pf.n_columns * n_rows_ * limbs_);
    pf.paths.resize(pf.n_columns * pf.path_len * 32);
    detail::check(lcpc_b200_commit_prove(h_.get(), tr.get(), nullptr, outer_tensor, outer_len, pf.n_degree_tests, pf.n_columns,
                                         pf.p_eval.data(), pf.p_random_vec.data(), pf.col_idx.data(), pf.cols.data(),
                                         pf.paths.data()),
                  ctx_.get());
    return pf;
  }
  lcpc_b200_commit *get() const { return h_.get(); }

 private:
  LcCommit(const LcEncoding &enc, lcpc_b200_commit *h) : ctx_(enc.context()), limbs_(enc.limbs()) {
    h_.reset(h, lcpc_b200_commit_free);
    detail::check(lcpc_b200_commit_dims(h, &n_rows_, &n_per_row_, &n_cols_, &n_hashes_));
  }
  Context ctx_;
  std::shared_ptr<lcpc_b200_commit> h_;
  size_t limbs_ = 0, n_rows_ = 0, n_per_row_ = 0, n_cols_ = 0, n_hashes_ = 0;
};

}  // namespace lcpc_b200
#endif  // LCPC_B200_HPP
