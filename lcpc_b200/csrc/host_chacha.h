// lcpc_b200/csrc/host_chacha.h -- rand_chacha 0.3's ChaCha20Rng on the host (setup and transcript-side
// sampling only: code generation, matgen.rs:43-44, and the column challenge, lcpc-2d/src/lib.rs:1073-1080).
#pragma once
#include <array>
#include <cstdint>
#include <cstring>

namespace lcpc {
namespace host {

// ChaCha20 keyed stream in the word order rand_chacha's BlockRng hands out
class ChaCha20Stream {
 public:
  // SeedableRng::seed_from_u64: a PCG32 walk fills the 32-byte key
  static ChaCha20Stream seed_from_u64(uint64_t state) {
    ChaCha20Stream r;
    for (auto &word : r.key_) {
      state = state * 6364136223846793005ull + 11634580027462260723ull;
      uint32_t xs = (uint32_t)(((state >> 18) ^ state) >> 27);
      unsigned rot = (unsigned)(state >> 59);
      word = (xs >> rot) | (xs << ((32 - rot) & 31));
    }
    return r;
  }
  // SeedableRng::from_seed: the 32 bytes are the key, read as little-endian words
  static ChaCha20Stream from_seed(const uint8_t key[32]) {
    ChaCha20Stream r;
    for (int i = 0; i < 8; i++)
      r.key_[i] = (uint32_t)key[4 * i] | ((uint32_t)key[4 * i + 1] << 8) | ((uint32_t)key[4 * i + 2] << 16) |
                  ((uint32_t)key[4 * i + 3] << 24);
    return r;
  }
  void set_stream(uint64_t s) { stream_ = s, pos_ = 16; }
  uint64_t next_u64() {
    uint64_t lo = next_u32();
    return lo | ((uint64_t)next_u32() << 32);
  }
  // rand 0.8 UniformInt<usize>: widening multiply with a rejection zone
  uint64_t below(uint64_t range) {
    const uint64_t reject = (0 - range) % range;
    const uint64_t zone = ~(uint64_t)0 - reject;
    for (;;) {
      unsigned __int128 wide = (unsigned __int128)next_u64() * range;
      if ((uint64_t)wide <= zone) return (uint64_t)(wide >> 64);
    }
  }

 private:
  static uint32_t rotl(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
  static void quarter(uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d) {
    a += b, d = rotl(d ^ a, 16);
    c += d, b = rotl(b ^ c, 12);
    a += b, d = rotl(d ^ a, 8);
    c += d, b = rotl(b ^ c, 7);
  }
  void refill() {
    std::array<uint32_t, 16> in = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
    for (int i = 0; i < 8; i++) in[4 + i] = key_[i];
    in[12] = (uint32_t)counter_, in[13] = (uint32_t)(counter_ >> 32);
    in[14] = (uint32_t)stream_, in[15] = (uint32_t)(stream_ >> 32);
    std::array<uint32_t, 16> x = in;
    for (int round = 0; round < 10; round++) {
      quarter(x[0], x[4], x[8], x[12]), quarter(x[1], x[5], x[9], x[13]);
      quarter(x[2], x[6], x[10], x[14]), quarter(x[3], x[7], x[11], x[15]);
      quarter(x[0], x[5], x[10], x[15]), quarter(x[1], x[6], x[11], x[12]);
      quarter(x[2], x[7], x[8], x[13]), quarter(x[3], x[4], x[9], x[14]);
    }
    for (int i = 0; i < 16; i++) block_[i] = x[i] + in[i];
    counter_++, pos_ = 0;
  }
  uint32_t next_u32() {
    if (pos_ >= 16) refill();
    return block_[pos_++];
  }
  std::array<uint32_t, 8> key_{};
  std::array<uint32_t, 16> block_{};
  uint64_t counter_ = 0, stream_ = 0;
  int pos_ = 16;
};

}  // namespace host
}  // namespace lcpc
