// lcpc_b200/csrc/host_matgen.h -- host-side setup of the two encodings (no device work).
// Declared here for api.cu; the extern "C" wrappers are in include/lcpc_b200_host.h.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace lcpc {
namespace host {

struct CodeSpec {  // codespec.rs:24-129: alpha = an/ad, beta = bn/bd, r = rn/rd, baselen
  size_t an, ad, bn, bd, rn, rd, baselen;
  double alpha() const { return (double)an / (double)ad; }
  double beta() const { return (double)bn / (double)bd; }
  double r() const { return (double)rn / (double)rd; }
  double dist() const { return (double)(bn * rd) / (double)(bd * rn); }
};
bool sdig_code_spec(int code, CodeSpec *out);  // SdigCode1..6, codespec.rs:169-232

struct LevelDims { size_t n, m, d; };  // inputs, outputs, non-zeros per input column

struct CscMatrix {  // CsMat::new_csc((m, n), ptrs, idxs, data), matgen.rs:187
  size_t m = 0, n = 0;
  std::vector<uint64_t> ptrs, idxs, data;  // data: nnz * L Montgomery limbs
};

struct SdigCode {
  int field = 0, code = 0;
  std::vector<CscMatrix> pre, post;
  size_t codeword_length() const;  // encode.rs:18-33
};

unsigned field_flog2(int field);       // SizedField::FLOG2 = NUM_BITS - 1 (lcpc-2d/src/lib.rs:68-71)
unsigned field_two_adicity(int field); // PrimeField::S
size_t n_degree_tests(size_t lambda, size_t len, size_t flog2);  // lcpc-2d/src/lib.rs:613-616
size_t ligero_n_col_opens(size_t rho_num, size_t rho_den);       // lcpc-ligero-pc/src/lib.rs:61-64
int ligero_get_dims(int field, size_t len, size_t rho_num, size_t rho_den, size_t *n_rows, size_t *n_per_row,
                    size_t *n_cols);                             // lcpc-ligero-pc/src/lib.rs:70-112
size_t sdig_n_col_opens(const CodeSpec &s);                      // lcpc-brakedown-pc/src/lib.rs:57-61
int sdig_choose_n_per_row(int field, const CodeSpec &s, size_t len, size_t *n_per_row, bool ml = false);  // :69-124
int sdig_level_dims(int field, const CodeSpec &s, size_t n, std::vector<LevelDims> *pre,
                    std::vector<LevelDims> *post);               // matgen.rs:56-111
int sdig_generate(int field, int code, size_t n_per_row, uint64_t seed, SdigCode *out);  // matgen.rs:28-52

}  // namespace host
}  // namespace lcpc
