// lcpc_b200/csrc/blake3.cuh -- BLAKE3 compression function for the device, written from the
// published BLAKE3 specification (default hash mode, 32-byte output).
//
// The reference hashes with `D: digest::Digest` instantiated to blake3::Hasher
// (lcpc-2d/src/tests.rs:12, lcpc-ligero-pc/src/bench.rs:12) and only ever calls
// new / update / finalize (lcpc-2d/src/lib.rs:719-735, 770-775), so the device needs just the
// compression function plus the chunk / parent tree rules; those live in kernels_hash.cu.
#pragma once
#include <stdint.h>

namespace lcpc {
namespace b3 {

enum : uint32_t { CHUNK_START = 1u, CHUNK_END = 2u, PARENT = 4u, ROOT = 8u };
constexpr int BLOCK_LEN = 64;
constexpr int CHUNK_LEN = 1024;

__host__ __device__ constexpr uint32_t iv(int i) {
  constexpr uint32_t v[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                             0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
  return v[i];
}

// message word used at position i of round r (the spec's permutation applied r times)
__host__ __device__ constexpr int sched(int r, int i) {
  constexpr int perm[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};
  int idx = i;
  for (int k = 0; k < r; k++) idx = perm[idx];
  return idx;
}

__device__ __forceinline__ uint32_t rotr(uint32_t x, int n) { return __funnelshift_r(x, x, n); }

// BLAKE3 is all adds / xors / rotates, which ptxas places on the ALU pipe (64 lanes/clk/SM) while the multiplier
// pipe idles.  An add written as x * 1 + y with the 1 hidden in constant memory has to be an IMAD, which moves it
// to the other pipe: LCPC_B3_FMA_ADDS of the two three-input adds of every G are done that way (0, 1 or 2).
// Measured on the Ft255 leaf kernel at 2^24 (131072 columns x 129 compressions): 0.831 / 0.784 / 0.734 ms.
// Levels 3 / 4 (round 2: also the two-input adds c += d) measured 0.752 / 0.752 ms against 0.746 for level 2 in the
// same run (profiles/r02_ab_blake3_fma_adds.jsonl): the pipes are balanced at 2.
#ifndef LCPC_B3_FMA_ADDS
#define LCPC_B3_FMA_ADDS 2
#endif
__device__ __constant__ uint32_t kB3One = 1u;
__device__ __forceinline__ uint32_t add_on_fma(uint32_t x, uint32_t y) {
  uint32_t r;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(kB3One), "r"(y));
  return r;
}
__device__ __forceinline__ uint32_t add3_first(uint32_t a, uint32_t b, uint32_t m) {
  if (LCPC_B3_FMA_ADDS >= 1) return add_on_fma(add_on_fma(a, b), m);
  return a + b + m;
}
__device__ __forceinline__ uint32_t add3_second(uint32_t a, uint32_t b, uint32_t m) {
  if (LCPC_B3_FMA_ADDS >= 2) return add_on_fma(add_on_fma(a, b), m);
  return a + b + m;
}

// levels 3 and 4 also move the first / both two-input adds (c += d): one ALU op becomes one multiplier-pipe op
__device__ __forceinline__ uint32_t add2_first(uint32_t c, uint32_t d) {
  if (LCPC_B3_FMA_ADDS >= 3) return add_on_fma(c, d);
  return c + d;
}
__device__ __forceinline__ uint32_t add2_second(uint32_t c, uint32_t d) {
  if (LCPC_B3_FMA_ADDS >= 4) return add_on_fma(c, d);
  return c + d;
}

#define LCPC_B3_G(a, b, c, d, mx, my) \
  do {                                \
    a = add3_first(a, b, (mx));       \
    d = rotr(d ^ a, 16);              \
    c = add2_first(c, d);             \
    b = rotr(b ^ c, 12);              \
    a = add3_second(a, b, (my));      \
    d = rotr(d ^ a, 8);               \
    c = add2_second(c, d);            \
    b = rotr(b ^ c, 7);               \
  } while (0)

template <int R>
__device__ __forceinline__ void round_fn(uint32_t (&v)[16], const uint32_t (&m)[16]) {
  LCPC_B3_G(v[0], v[4], v[8], v[12], m[sched(R, 0)], m[sched(R, 1)]);
  LCPC_B3_G(v[1], v[5], v[9], v[13], m[sched(R, 2)], m[sched(R, 3)]);
  LCPC_B3_G(v[2], v[6], v[10], v[14], m[sched(R, 4)], m[sched(R, 5)]);
  LCPC_B3_G(v[3], v[7], v[11], v[15], m[sched(R, 6)], m[sched(R, 7)]);
  LCPC_B3_G(v[0], v[5], v[10], v[15], m[sched(R, 8)], m[sched(R, 9)]);
  LCPC_B3_G(v[1], v[6], v[11], v[12], m[sched(R, 10)], m[sched(R, 11)]);
  LCPC_B3_G(v[2], v[7], v[8], v[13], m[sched(R, 12)], m[sched(R, 13)]);
  LCPC_B3_G(v[3], v[4], v[9], v[14], m[sched(R, 14)], m[sched(R, 15)]);
}

// cv <- first 8 words of compress(cv, m, counter, block_len, flags)
__device__ __forceinline__ void compress(uint32_t (&cv)[8], const uint32_t (&m)[16], uint64_t counter,
                                         uint32_t block_len, uint32_t flags) {
  uint32_t v[16];
#pragma unroll
  for (int i = 0; i < 8; i++) v[i] = cv[i];
#pragma unroll
  for (int i = 0; i < 4; i++) v[8 + i] = iv(i);
  v[12] = (uint32_t)counter;
  v[13] = (uint32_t)(counter >> 32);
  v[14] = block_len;
  v[15] = flags;
  round_fn<0>(v, m);
  round_fn<1>(v, m);
  round_fn<2>(v, m);
  round_fn<3>(v, m);
  round_fn<4>(v, m);
  round_fn<5>(v, m);
  round_fn<6>(v, m);
#pragma unroll
  for (int i = 0; i < 8; i++) cv[i] = v[i] ^ v[i + 8];
}

__device__ __forceinline__ void set_iv(uint32_t (&cv)[8]) {
#pragma unroll
  for (int i = 0; i < 8; i++) cv[i] = iv(i);
}

}  // namespace b3
}  // namespace lcpc
