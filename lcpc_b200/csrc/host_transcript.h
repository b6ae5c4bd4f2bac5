// lcpc_b200/csrc/host_transcript.h -- the Fiat-Shamir transcript prove()/verify() thread their challenges
// through (reference: `merlin::Transcript`, lcpc-2d/src/lib.rs:16, used at :46-49, :870-871, :891-898,
// :903-905, :1026-1027, :1043-1045, :1061-1063, :1073-1075).
//
// merlin 2.0 is not vendored in the reference tree; this is its published construction: STROBE-128 (the
// "lite" subset merlin uses: meta-AD, AD, PRF, KEY) over Keccak-f[1600], rate 166, protocol label
// "Merlin v1.0".  The transcript is inherently sequential host work (every absorb depends on the previous
// state), so it stays on the host exactly as in the reference; the device feeds it canonical bytes and
// consumes the 32-byte challenges (kernels_collapse.cu: expand_tensor_kernel).
#pragma once
#include <cstddef>
#include <cstdint>

namespace lcpc {
namespace host {

void keccak_f1600(uint64_t st[25]);
bool keccak_in_place();  // whether the selected build is fastest on the caller's own state (no copy)

class Strobe128 {
 public:
  explicit Strobe128(const uint8_t *protocol_label, size_t n);
  void meta_ad(const uint8_t *data, size_t n, bool more);
  void meta_ad_label_len(const uint8_t *label, size_t nl, const uint8_t len[4]);  // meta_ad(label) + meta_ad(len, more)
  void ad(const uint8_t *data, size_t n, bool more);
  void prf(uint8_t *data, size_t n, bool more);
  void key(const uint8_t *data, size_t n, bool more);

 private:
  static constexpr uint8_t R = 166;
  enum : uint8_t { FLAG_I = 1, FLAG_A = 2, FLAG_C = 4, FLAG_T = 8, FLAG_M = 16, FLAG_K = 32 };
  void run_f();
  void absorb(const uint8_t *data, size_t n);
  void overwrite(const uint8_t *data, size_t n);
  void squeeze(uint8_t *data, size_t n);
  void begin_op(uint8_t flags, bool more);
  uint64_t lanes_[25];  // the duplex state; STROBE addresses it as 200 little-endian bytes (st_), Keccak as 25 lanes
  uint8_t *st() { return reinterpret_cast<uint8_t *>(lanes_); }
  uint8_t pos_ = 0, pos_begin_ = 0, cur_flags_ = 0;
};

class Transcript {
 public:
  Transcript(const uint8_t *label, size_t n);                    // Transcript::new
  void append_message(const uint8_t *label, size_t nl, const uint8_t *msg, size_t n);
  void append_u64(const uint8_t *label, size_t nl, uint64_t x);
  void challenge_bytes(const uint8_t *label, size_t nl, uint8_t *out, size_t n);
  // FieldHash::transcript_update (lcpc-2d/src/lib.rs:46-49) for `count` elements in a row: element i is
  // append_message(label, repr[i*elem_bytes .. (i+1)*elem_bytes))
  void append_elems(const uint8_t *label, size_t nl, const uint8_t *repr, size_t elem_bytes, size_t count);

 private:
  Strobe128 strobe_;
};

}  // namespace host
}  // namespace lcpc

// the opaque handle of include/lcpc_b200_host.h
struct lcpc_b200_transcript {
  lcpc::host::Transcript tr;
  lcpc_b200_transcript(const uint8_t *label, size_t n) : tr(label, n) {}
};
