// tests/host/cpp_mirror_check.cpp -- exercises include/lcpc_b200.hpp (the C++ host-side mirror of the reference's
// operator interface) against the in-tree library.
//   cpp_mirror_check host        transcript conformance vector + "no device -> Error(ERR_CUDA)" (no GPU needed)
//   cpp_mirror_check wire FILE   writes the bincode image of a deterministic synthetic LcEvalProof (Ft255 shapes) to FILE
//                                after a serialize -> deserialize -> serialize round trip (no GPU needed)
//   cpp_mirror_check gpu KIND    KIND = ligero | sdig: commit, prove, verify, tamper, re-import of the commit's fields;
//                                prints `root <hex>`, `eval <hex>`, `cols <first opened column numbers>`, `ok`
// The coefficient / tensor values are small integers used directly as limb 0 (any value below p is a valid element),
// so tests/test_cpp_mirror.py can rebuild the same inputs and compare with the Python path.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/lcpc_b200.hpp"

using namespace lcpc_b200;

static std::string hex(const uint8_t *p, size_t n) {
  static const char *d = "0123456789abcdef";
  std::string s;
  for (size_t i = 0; i < n; i++) s += d[p[i] >> 4], s += d[p[i] & 15];
  return s;
}

static std::vector<uint64_t> small_elems(size_t n, size_t L, uint64_t mul, uint64_t add) {
  std::vector<uint64_t> v(n * L, 0);
  for (size_t i = 0; i < n; i++) v[i * L] = i * mul + add;
  return v;
}

static int run_host() {
  Transcript t("test protocol");
  t.append_message("some label", "some data");
  auto c = t.challenge_bytes("challenge", 32);
  printf("merlin %s\n", hex(c.data(), 32).c_str());
  try {
    Context ctx(0);
    printf("device present\n");
  } catch (const Error &e) {
    printf("no device: code %d\n", e.code());
    if (e.code() != LCPC_B200_ERR_CUDA) return 1;
  }
  return 0;
}

static int run_wire(const char *path) {
  const size_t L = 4;
  LcEvalProof p;
  p.n_cols = 1024, p.n_per_row = 512, p.n_degree_tests = 2, p.n_columns = 7, p.n_rows = 3, p.path_len = 10;
  for (size_t i = 0; i < p.n_per_row * L; i++) p.p_eval.push_back(0x0101010101010101ull * (i % 251) + i);
  for (size_t i = 0; i < p.n_degree_tests * p.n_per_row * L; i++) p.p_random_vec.push_back(i * 0x9e3779b97f4a7c15ull);
  for (size_t i = 0; i < p.n_columns * p.n_rows * L; i++) p.cols.push_back(~(uint64_t)i * 3);
  for (size_t i = 0; i < p.n_columns * p.path_len * 32; i++) p.paths.push_back((uint8_t)(i * 7 + 1));
  std::vector<uint8_t> w = serialize(p, L);
  LcEvalProof q = deserialize_proof(w.data(), w.size(), L);
  if (q.n_cols != p.n_cols || q.p_eval != p.p_eval || q.p_random_vec != p.p_random_vec || q.cols != p.cols || q.paths != p.paths ||
      q.n_rows != p.n_rows || q.path_len != p.path_len || serialize(q, L) != w)
    return 11;
  try {
    deserialize_proof(w.data(), w.size() - 1, L);
    return 12;
  } catch (const Error &) {
  }
  LcRoot r;
  for (int i = 0; i < 32; i++) r.root[i] = (uint8_t)(200 - i);
  std::vector<uint8_t> wr = serialize(r);
  FILE *f = fopen(path, "wb");
  if (!f) return 13;
  fwrite(w.data(), 1, w.size(), f);
  fwrite(wr.data(), 1, wr.size(), f);
  fclose(f);
  printf("wire %zu %zu\n", w.size(), wr.size());
  return 0;
}

static int run_gpu(const std::string &kind) {
  Context ctx(0);
  const Field f = Field::Ft127;
  const size_t L = 2, len = 3000;
  std::unique_ptr<LcEncoding> enc;
  if (kind == "ligero") enc.reset(new LigeroEncoding(LigeroEncoding::create(ctx, f, len)));
  else enc.reset(new SdigEncoding(SdigEncoding::create(ctx, f, len, /*seed=*/5)));
  auto coeffs = small_elems(len, L, 7, 1);
  LcCommit c = LcCommit::commit(coeffs.data(), len, *enc);
  auto dims = enc->get_dims(len);
  if (dims[0] != c.get_n_rows() || dims[1] != c.get_n_per_row() || dims[2] != c.get_n_cols()) return 2;
  LcRoot root = c.get_root();
  printf("root %s\n", hex(root.root.data(), 32).c_str());
  auto outer = small_elems(c.get_n_rows(), L, 11, 3), inner = small_elems(c.get_n_per_row(), L, 13, 2);
  Transcript tp("cpp mirror");
  LcEvalProof pf = c.prove(outer.data(), c.get_n_rows(), *enc, tp);
  Transcript tv("cpp mirror");
  auto ev = pf.verify(root, outer.data(), c.get_n_rows(), inner.data(), c.get_n_per_row(), *enc, tv);
  printf("eval %s\n", hex(reinterpret_cast<const uint8_t *>(ev.data()), 8 * L).c_str());
  printf("cols %llu %llu %llu\n", (unsigned long long)pf.col_idx[0], (unsigned long long)pf.col_idx[1], (unsigned long long)pf.col_idx[2]);
  // both transcripts end in the same state
  if (tp.challenge_bytes("after", 16) != tv.challenge_bytes("after", 16)) return 3;
  // a tampered opening must be rejected with a VerifierError code
  LcEvalProof bad = pf;
  bad.cols[0] ^= 1;
  Transcript tb("cpp mirror");
  try {
    bad.verify(root, outer.data(), c.get_n_rows(), inner.data(), c.get_n_per_row(), *enc, tb);
    return 4;
  } catch (const Error &e) {
    if (e.code() > LCPC_B200_VERR_NUM_COL_OPENS || e.code() < LCPC_B200_VERR_ENCODING_DIMS) return 5;
  }
  // a wrong outer tensor length is ProverError::OuterTensor
  Transcript tw("cpp mirror");
  try {
    c.prove(outer.data(), c.get_n_rows() - 1, *enc, tw);
    return 6;
  } catch (const Error &e) {
    if (e.code() != LCPC_B200_ERR_OUTER_TENSOR) return 7;
  }
  // Deserialize for LcCommit: fields out, back in, same root
  std::vector<uint64_t> comm(c.get_n_rows() * c.get_n_cols() * L), co(c.get_n_rows() * c.get_n_per_row() * L);
  std::vector<uint8_t> hashes(c.n_hashes() * 32);
  c.download(comm.data(), co.data(), hashes.data());
  LcCommit again = LcCommit::from_fields(*enc, comm.data(), comm.size() / L, co.data(), co.size() / L, hashes.data(), c.n_hashes(),
                                         c.get_n_rows());
  if (!(again.get_root() == root)) return 8;
  if (memcmp(hashes.data() + (c.n_hashes() - 1) * 32, root.root.data(), 32) != 0) return 9;
  printf("ok\n");
  return 0;
}

int main(int argc, char **argv) {
  try {
    if (argc >= 2 && std::string(argv[1]) == "host") return run_host();
    if (argc >= 3 && std::string(argv[1]) == "gpu") return run_gpu(argv[2]);
    if (argc >= 3 && std::string(argv[1]) == "wire") return run_wire(argv[2]);
  } catch (const std::exception &e) {
    fprintf(stderr, "exception: %s\n", e.what());
    return 20;
  }
  fprintf(stderr, "usage: cpp_mirror_check host | gpu ligero|sdig\n");
  return 64;
}
