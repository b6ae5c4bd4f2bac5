// lcpc_b200/csrc/field.cuh -- prime-field arithmetic for the lcpc test fields on sm_100a.
//
// Replaces, for the device path, the arithmetic `#[derive(PrimeField)]` generates for
//   Ft63 / Ft127 / Ft191 / Ft255   (reference: lcpc-test-fields/src/lib.rs:18-22, 30-34, 42-46, 54-58).
// An element is the in-memory image of the reference's `struct FtNNN([u64; L])`: L little-endian
// u64 limbs holding the MONTGOMERY image x*R mod p, R = 2^(64 L).  On the device the same bytes are
// read as N = 2L little-endian 32-bit limbs, which keeps R unchanged, so results are bit-identical
// to the 64-bit-limb CPU arithmetic (all operations are exact mod p and results are canonical).
//
// Design notes (B200): the SM has no 64-bit integer multiplier; the unit of work is IMAD.WIDE.U32
// (32x32+64) on the FMA pipe, with adds/logic on the ALU pipe.  Montgomery multiplication is
// written as an interleaved (CIOS) product/reduction over two accumulator arrays that always hold
// 64-bit-aligned register pairs ("even"/"odd" alignment), so every multiply-add is one wide IMAD
// fed by a carry flag rather than a 3-instruction mul/add/add-carry group.  All four moduli are
// = 1 mod 2^32, hence -p^{-1} mod 2^32 = 0xffffffff: the Montgomery quotient digit is just the
// negated low limb and the p[0] column of every reduction step costs two adds instead of an IMAD.
#pragma once
#include <stdint.h>

namespace lcpc {

enum FieldId : int { FT63 = 1, FT127 = 2, FT191 = 3, FT255 = 4 };

// ---- moduli as 32-bit limbs (lcpc-test-fields/src/lib.rs:19,31,43,55 give them in decimal) ----
template <int FID> struct FieldP;
template <> struct FieldP<FT63> {
  static constexpr int N = 2;
  __host__ __device__ static constexpr uint32_t P(int i) {
    constexpr uint32_t p[N] = {0x00000001u, 0x46d07600u};
    return p[i];
  }
};
template <> struct FieldP<FT127> {
  static constexpr int N = 4;
  __host__ __device__ static constexpr uint32_t P(int i) {
    constexpr uint32_t p[N] = {0x00000001u, 0x7f2bd900u, 0xba20e0bfu, 0x6e754097u};
    return p[i];
  }
};
template <> struct FieldP<FT191> {
  static constexpr int N = 6;
  __host__ __device__ static constexpr uint32_t P(int i) {
    constexpr uint32_t p[N] = {0x00000001u, 0xd2468200u, 0x0ceecbcdu, 0x93688827u, 0x3fbc8ddau, 0x453708aau};
    return p[i];
  }
};
template <> struct FieldP<FT255> {
  static constexpr int N = 8;
  __host__ __device__ static constexpr uint32_t P(int i) {
    constexpr uint32_t p[N] = {0x00000001u, 0x02a4f200u, 0x86595f30u, 0xef73c790u,
                               0xb9575969u, 0xfda9df04u, 0x6e4d2900u, 0x663c799bu};
    return p[i];
  }
};

// The same limbs in the constant bank: ptxas fuses mad.lo.cc/madc.hi.cc into one IMAD.WIDE.U32.X only
// when the multiplier is a register or a c[bank][offset] operand, not a 32-bit immediate.
__device__ __constant__ uint32_t kModulusLimbs[5][8] = {
    {0, 0, 0, 0, 0, 0, 0, 0},
    {0x00000001u, 0x46d07600u, 0, 0, 0, 0, 0, 0},
    {0x00000001u, 0x7f2bd900u, 0xba20e0bfu, 0x6e754097u, 0, 0, 0, 0},
    {0x00000001u, 0xd2468200u, 0x0ceecbcdu, 0x93688827u, 0x3fbc8ddau, 0x453708aau, 0, 0},
    {0x00000001u, 0x02a4f200u, 0x86595f30u, 0xef73c790u, 0xb9575969u, 0xfda9df04u, 0x6e4d2900u, 0x663c799bu},
};

// -p^{-1} mod 2^32 (= 0xffffffff for every modulus here) kept opaque in the constant bank: when the
// quotient digit is produced by an ALU negate, ptxas 12.9 splits every dependent IMAD.WIDE.U32.X into
// IMAD.X + IMAD.HI.U32.X (twice the FMA-pipe work); produced by an IMAD it keeps them fused.
__device__ __constant__ uint32_t kMontInv32 = 0xffffffffu;
// a zero ptxas cannot see through (A/B knob LCPC_OPAQUE_ZERO): `x + 0 + carry` written against it has to stay
// a three-operand add on the ALU pipe instead of becoming IMAD.X on the multiplier pipe
__device__ __constant__ uint32_t kZero32 = 0u;

// ---- carry-flag primitives (PTX extended-precision integer arithmetic) ----
// Each statement is `asm volatile` so NVVM keeps them in program order; ptxas tracks CC itself.
#define LCPC_DEV __host__ __device__ __forceinline__

#ifdef __CUDA_ARCH__
LCPC_DEV void add_cc(uint32_t &d, uint32_t a, uint32_t b) { asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
LCPC_DEV void addc_cc(uint32_t &d, uint32_t a, uint32_t b) { asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
LCPC_DEV void addc(uint32_t &d, uint32_t a, uint32_t b) { asm volatile("addc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
// d += carry
LCPC_DEV void add_carry(uint32_t &d) {
#if defined(LCPC_OPAQUE_ZERO)
  asm volatile("addc.u32 %0, %0, %1;" : "+r"(d) : "r"(kZero32));
#else
  asm volatile("addc.u32 %0, %0, 0;" : "+r"(d));
#endif
}
LCPC_DEV void sub_cc(uint32_t &d, uint32_t a, uint32_t b) { asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
LCPC_DEV void subc_cc(uint32_t &d, uint32_t a, uint32_t b) { asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
LCPC_DEV void subc(uint32_t &d, uint32_t a, uint32_t b) { asm volatile("subc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
// (lo,hi) = a*b + (clo,chi) [+ carry], as a lo/hi pair that ptxas fuses into one IMAD.WIDE.U32
LCPC_DEV void mad_wide_cc(uint32_t &lo, uint32_t &hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi) {
  asm volatile("mad.lo.cc.u32 %0, %2, %3, %4;\n\tmadc.hi.cc.u32 %1, %2, %3, %5;"
               : "=r"(lo), "=r"(hi) : "r"(a), "r"(b), "r"(clo), "r"(chi));
}
LCPC_DEV void madc_wide_cc(uint32_t &lo, uint32_t &hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi) {
  asm volatile("madc.lo.cc.u32 %0, %2, %3, %4;\n\tmadc.hi.cc.u32 %1, %2, %3, %5;"
               : "=r"(lo), "=r"(hi) : "r"(a), "r"(b), "r"(clo), "r"(chi));
}
LCPC_DEV void madc_wide(uint32_t &lo, uint32_t &hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi) {
  asm volatile("madc.lo.cc.u32 %0, %2, %3, %4;\n\tmadc.hi.u32 %1, %2, %3, %5;"
               : "=r"(lo), "=r"(hi) : "r"(a), "r"(b), "r"(clo), "r"(chi));
}
LCPC_DEV void mul_wide(uint32_t &lo, uint32_t &hi, uint32_t a, uint32_t b) {
  asm volatile("mul.lo.u32 %0, %2, %3;\n\tmul.hi.u32 %1, %2, %3;" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
}
#else
// Host emulation of the same primitives with an explicit carry flag.  It exists so the carry-chain
// LOGIC of this header can be unit-tested on a machine without a GPU (tests/host_field_check.cu);
// no product entry point computes on the host.
namespace hostcc { inline thread_local uint32_t cf = 0; }
inline void add_cc(uint32_t &d, uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b; d = (uint32_t)s; hostcc::cf = (uint32_t)(s >> 32); }
inline void addc_cc(uint32_t &d, uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b + hostcc::cf; d = (uint32_t)s; hostcc::cf = (uint32_t)(s >> 32); }
inline void addc(uint32_t &d, uint32_t a, uint32_t b) { d = a + b + hostcc::cf; }
inline void add_carry(uint32_t &d) { d = d + hostcc::cf; }
inline void sub_cc(uint32_t &d, uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a - b; d = (uint32_t)s; hostcc::cf = (uint32_t)(s >> 63); }
inline void subc_cc(uint32_t &d, uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a - b - hostcc::cf; d = (uint32_t)s; hostcc::cf = (uint32_t)(s >> 63); }
inline void subc(uint32_t &d, uint32_t a, uint32_t b) { d = a - b - hostcc::cf; }
inline void mad_wide_cc(uint32_t &lo, uint32_t &hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi) {
  uint64_t pr = (uint64_t)a * b;
  uint64_t s = (pr & 0xffffffffu) + clo; lo = (uint32_t)s;
  uint64_t t = (pr >> 32) + chi + (s >> 32); hi = (uint32_t)t; hostcc::cf = (uint32_t)(t >> 32);
}
inline void madc_wide_cc(uint32_t &lo, uint32_t &hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi) {
  uint64_t pr = (uint64_t)a * b;
  uint64_t s = (pr & 0xffffffffu) + clo + hostcc::cf; lo = (uint32_t)s;
  uint64_t t = (pr >> 32) + chi + (s >> 32); hi = (uint32_t)t; hostcc::cf = (uint32_t)(t >> 32);
}
inline void madc_wide(uint32_t &lo, uint32_t &hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi) {
  madc_wide_cc(lo, hi, a, b, clo, chi);
}
inline void mul_wide(uint32_t &lo, uint32_t &hi, uint32_t a, uint32_t b) {
  uint64_t pr = (uint64_t)a * b; lo = (uint32_t)pr; hi = (uint32_t)(pr >> 32);
}
#endif

// the zero operand of carry-only adds: with LCPC_OPAQUE_ZERO a constant-bank load ptxas cannot fold, which keeps
// `x + 0 + carry` an IADD3.X on the ALU pipe instead of an IMAD.X on the multiplier pipe
LCPC_DEV uint32_t zero_operand() {
#if defined(__CUDA_ARCH__) && defined(LCPC_OPAQUE_ZERO)
  return kZero32;
#else
  return 0u;
#endif
}

template <int FID> struct Field {
  using FP = FieldP<FID>;
  static constexpr int N = FP::N;  // 32-bit limbs
  static constexpr int BYTES = 4 * N;

  struct Elem { uint32_t v[N]; };

  // modulus limb as a multiplier operand (constant bank on the device, see kModulusLimbs)
  LCPC_DEV static uint32_t PM(int i) {
#ifdef __CUDA_ARCH__
    return kModulusLimbs[FID][i];
#else
    return FP::P(i);
#endif
  }

  // R mod p = 2^(32N) mod p, the Montgomery image of 1, folded at compile time (32N doublings with a conditional
  // subtraction; a kernel that did this at run time spent a quarter of its 55 us on it)
  struct Limbs { uint32_t v[N]; };
  __host__ __device__ static constexpr Limbs compute_one() {
    Limbs r{};
    r.v[0] = 1;
    for (int it = 0; it < 32 * N; it++) {
      uint32_t carry = 0;
      for (int i = 0; i < N; i++) {  // r = 2r: r < p < 2^(32N-1), no carry out
        const uint32_t nv = (r.v[i] << 1) | carry;
        carry = r.v[i] >> 31;
        r.v[i] = nv;
      }
      bool ge = true;
      for (int i = N - 1; i >= 0; i--) {
        if (r.v[i] != FP::P(i)) {
          ge = r.v[i] > FP::P(i);
          break;
        }
      }
      if (ge) {
        uint64_t borrow = 0;
        for (int i = 0; i < N; i++) {
          const uint64_t d = (uint64_t)r.v[i] - FP::P(i) - borrow;
          r.v[i] = (uint32_t)d;
          borrow = (d >> 63) & 1u;
        }
      }
    }
    return r;
  }
  LCPC_DEV static Elem one() {
    constexpr Limbs o = compute_one();
    Elem e;
#pragma unroll
    for (int i = 0; i < N; i++) e.v[i] = o.v[i];
    return e;
  }

  LCPC_DEV static Elem zero() {
    Elem r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = 0;
    return r;
  }

  LCPC_DEV static bool is_zero(const Elem &a) {
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < N; i++) acc |= a.v[i];
    return acc == 0;
  }

  // r = t - p if t >= p else t   (t < 2p)
  LCPC_DEV static Elem cond_sub_p(const Elem &t) {
    Elem u;
    uint32_t borrow;
    sub_cc(u.v[0], t.v[0], FP::P(0));
#pragma unroll
    for (int i = 1; i < N; i++) subc_cc(u.v[i], t.v[i], FP::P(i));
    subc(borrow, 0u, 0u);  // 0xffffffff if t < p
    Elem r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = borrow ? t.v[i] : u.v[i];
    return r;
  }

  LCPC_DEV static Elem add(const Elem &a, const Elem &b) {
    Elem t;
    add_cc(t.v[0], a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) addc_cc(t.v[i], a.v[i], b.v[i]);
    addc(t.v[N - 1], a.v[N - 1], b.v[N - 1]);  // 2p < 2^(32N): no carry out
    return cond_sub_p(t);
  }

  LCPC_DEV static Elem sub(const Elem &a, const Elem &b) {
    Elem d;
    uint32_t mask;
    sub_cc(d.v[0], a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < N; i++) subc_cc(d.v[i], a.v[i], b.v[i]);
    subc(mask, 0u, 0u);  // all ones iff a < b
    add_cc(d.v[0], d.v[0], FP::P(0) & mask);
#pragma unroll
    for (int i = 1; i < N - 1; i++) addc_cc(d.v[i], d.v[i], FP::P(i) & mask);
    addc(d.v[N - 1], d.v[N - 1], FP::P(N - 1) & mask);
    return d;
  }

  // ---- Montgomery product, interleaved, two aligned accumulator arrays ----
  // Invariant at the top of step i: X holds limbs at positions 0..N-1, Y at positions 1..N of the
  // running value V = X + 2^32 Y.  A step adds a*b_i and m*p (m = -V mod 2^32), which clears
  // position 0, then positions shift down by one: the array that was Y becomes X, and the old X,
  // minus its two low limbs, becomes the new Y (the "rshift" below); old X[1] folds into new X[0]
  // and its carry feeds the chain that starts one position higher.
  template <bool FIRST, bool WITH_AB>
  LCPC_DEV static void step(uint32_t (&X)[N], uint32_t (&O)[N], const uint32_t (&a)[N], uint32_t bi) {
    // entry: X = array now at position 0, O = previous step's X (to be shifted into Y), unless FIRST
    if (FIRST) {
      if (WITH_AB) {
#pragma unroll
        for (int j = 0; j < N; j += 2) mul_wide(O[j], O[j + 1], a[j + 1], bi);   // Y = a_odd * b_i
#pragma unroll
        for (int j = 0; j < N; j += 2) mul_wide(X[j], X[j + 1], a[j], bi);       // X = a_even * b_i
      }
    } else {
      add_cc(X[0], X[0], O[1]);
      if (WITH_AB) {
#pragma unroll
        for (int j = 0; j < N - 2; j += 2) madc_wide_cc(O[j], O[j + 1], a[j + 1], bi, O[j + 2], O[j + 3]);
        madc_wide(O[N - 2], O[N - 1], a[N - 1], bi, 0u, 0u);
        mad_wide_cc(X[0], X[1], a[0], bi, X[0], X[1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) madc_wide_cc(X[j], X[j + 1], a[j], bi, X[j], X[j + 1]);
        add_carry(O[N - 1]);
      } else {
#pragma unroll
        for (int j = 0; j < N - 2; j++) addc_cc(O[j], O[j + 2], 0u);
        addc(O[N - 2], 0u, 0u);
        O[N - 1] = 0;
      }
    }
    // reduction digit; p = 1 mod 2^32 so m = -X[0] and the p[0] column is X[0] + m = 2^32 [X[0] != 0]
#if defined(__CUDA_ARCH__) && !defined(LCPC_M_ALU)
    uint32_t m = X[0] * kMontInv32;
#else
    uint32_t m = 0u - X[0];
#endif
    // Y += m * p_odd
    mad_wide_cc(O[0], O[1], m, PM(1), O[0], O[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) madc_wide_cc(O[j], O[j + 1], m, PM(j + 1), O[j], O[j + 1]);
    // X += m * p_even   (carry out of position N-1 lands in Y[N-1])
    add_cc(X[0], X[0], m);
    addc_cc(X[1], X[1], 0u);
#pragma unroll
    for (int j = 2; j < N; j += 2) madc_wide_cc(X[j], X[j + 1], m, PM(j), X[j], X[j + 1]);
    add_carry(O[N - 1]);
  }

  // merge the two arrays after the last step and bring the result into [0, p)
  LCPC_DEV static Elem finish(const uint32_t (&Y)[N], const uint32_t (&Xold)[N]) {
    // positions after the final shift: Y -> 0..N-1, Xold[1..N-1] -> 0..N-2
    Elem t;
    add_cc(t.v[0], Y[0], Xold[1]);
#pragma unroll
    for (int j = 1; j < N - 1; j++) addc_cc(t.v[j], Y[j], Xold[j + 1]);
    addc(t.v[N - 1], Y[N - 1], 0u);
    return cond_sub_p(t);
  }

  LCPC_DEV static Elem mul(const Elem &a, const Elem &b) {
#if defined(LCPC_KARATSUBA)
    if constexpr (N >= LCPC_KARATSUBA) return mul_karatsuba(a, b);
#endif
    uint32_t E[N], O[N];
    step<true, true>(E, O, a.v, b.v[0]);
#pragma unroll
    for (int i = 1; i < N; i += 2) {
      step<false, true>(O, E, a.v, b.v[i]);
      if (i + 1 < N) step<false, true>(E, O, a.v, b.v[i + 1]);
    }
    // N is even: the last step ran with X = O, so Y = E
    return finish(E, O);
  }

  // ---- double-width products, lazily reduced sums of products, stand-alone reduction ----
  // Used where many products feed ONE result (sparse rows of the expander code, the prover's
  // row combination): sum_k a_k b_k is accumulated as a 2N-limb integer and Montgomery-reduced once,
  // which removes the N(N-1) reduction multiplies from every term but the last.  REDC is linear
  // mod p, so the canonical result is the one the reference gets by reducing every product.
  struct Wide { uint32_t v[2 * N]; };

  // One row of a schoolbook product on two 64-bit-aligned accumulator arrays.  X receives
  // a[1],a[3],.. times bi on the pairs (X[0],X[1]), (X[2],X[3]), ..; Y receives a[0],a[2],.. times bi
  // on (Y[0],Y[1]), ..; X sits one limb above Y.  The top pair of X is first written by this row, so
  // the carry out of the Y chain can be absorbed there without rippling further.
  template <int M, bool X_FRESH, bool P0_IS_ONE>
  LCPC_DEV static void mad_row(uint32_t *X, uint32_t *Y, const uint32_t *a, uint32_t bi) {
    if (X_FRESH || M == 2) {
#pragma unroll
      for (int j = 1; j < M; j += 2) {
        if (X_FRESH || j == M - 1) mul_wide(X[j - 1], X[j], a[j], bi);
      }
    } else {
      mad_wide_cc(X[0], X[1], a[1], bi, X[0], X[1]);
#pragma unroll
      for (int j = 3; j < M - 1; j += 2) madc_wide_cc(X[j - 1], X[j], a[j], bi, X[j - 1], X[j]);
      madc_wide(X[M - 2], X[M - 1], a[M - 1], bi, 0u, 0u);
    }
    if (P0_IS_ONE) {  // a[0] == 1: the product is bi itself
      add_cc(Y[0], Y[0], bi);
      addc_cc(Y[1], Y[1], 0u);
    } else {
      mad_wide_cc(Y[0], Y[1], a[0], bi, Y[0], Y[1]);
    }
#pragma unroll
    for (int j = 2; j < M; j += 2) madc_wide_cc(Y[j], Y[j + 1], a[j], bi, Y[j], Y[j + 1]);
    addc(X[M - 1], X[M - 1], 0u);
  }

  // T[0..2M) = a[0..M) * b[0..M), M even: M^2 wide multiply-adds, M - 1 + 2M - 1 carry adds
  template <int M>
  LCPC_DEV static void mul_full_n(uint32_t *T, const uint32_t *a, const uint32_t *b) {
    uint32_t O[2 * M];  // O[k] sits at limb position k + 1
#pragma unroll
    for (int j = 0; j < M; j += 2) mul_wide(T[j], T[j + 1], a[j], b[0]);
#pragma unroll
    for (int j = 1; j < M; j += 2) mul_wide(O[j - 1], O[j], a[j], b[0]);
#pragma unroll
    for (int i = 1; i < M; i++) {
      if (i & 1) mad_row<M, false, false>(&T[i + 1], &O[i - 1], a, b[i]);
      else mad_row<M, false, false>(&O[i], &T[i], a, b[i]);
    }
    add_cc(T[1], T[1], O[0]);
#pragma unroll
    for (int k = 1; k < 2 * M - 2; k++) addc_cc(T[1 + k], T[1 + k], O[k]);
    addc(T[2 * M - 1], T[2 * M - 1], 0u);
  }

  LCPC_DEV static Wide mul_full(const Elem &a, const Elem &b) {
    Wide t;
    mul_full_n<N>(t.v, a.v, b.v);
    return t;
  }

  LCPC_DEV static Wide wide_zero() {
    Wide t;
#pragma unroll
    for (int i = 0; i < 2 * N; i++) t.v[i] = 0;
    return t;
  }

  // acc += t, then acc -= p * 2^(32N) if that leaves acc >= 2^(64N - 1).  Invariant: acc < 2^(64N-1)
  // before and after (every product is < p^2 < 0.19 * 2^(64N), and 0.27 <= p / 2^(32N) < 0.44 for the
  // four moduli); subtracting a multiple of p * R does not change REDC's result mod p.
  // acc -= p * 2^(32N) if acc >= 2^(64N - 1)
  LCPC_DEV static void wide_fold(Wide &acc) {
    const uint32_t mask = (uint32_t)((int32_t)acc.v[2 * N - 1] >> 31);
    sub_cc(acc.v[N], acc.v[N], FP::P(0) & mask);
#pragma unroll
    for (int i = 1; i < N - 1; i++) subc_cc(acc.v[N + i], acc.v[N + i], FP::P(i) & mask);
    subc(acc.v[2 * N - 1], acc.v[2 * N - 1], FP::P(N - 1) & mask);
  }
  LCPC_DEV static void wide_add(Wide &acc, const Wide &t) {
    add_cc(acc.v[0], acc.v[0], t.v[0]);
#pragma unroll
    for (int i = 1; i < 2 * N - 1; i++) addc_cc(acc.v[i], acc.v[i], t.v[i]);
    addc(acc.v[2 * N - 1], acc.v[2 * N - 1], t.v[2 * N - 1]);
  }
  LCPC_DEV static void wide_add_fold(Wide &acc, const Wide &t) {
    wide_add(acc, t);
    wide_fold(acc);
  }
  // merge two folded partial sums (both < 2^(64N-1)): the sum is < 2^(64N) and two folds bring it back under
  // 2^(64N-1) for every modulus here (0.27 <= p / 2^(32N) < 0.44)
  LCPC_DEV static void wide_merge(Wide &acc, const Wide &t) {
    wide_add(acc, t);
    wide_fold(acc);
    wide_fold(acc);
  }
  LCPC_DEV static void mac_wide(Wide &acc, const Elem &a, const Elem &b) { wide_add_fold(acc, mul_full(a, b)); }

  // ---- sums of MANY products with no carry propagation per term ----
  // mac_wide pays ~6N ALU operations per term on top of the N^2 wide multiply-adds (merging the product's two
  // accumulator arrays, adding it to the running sum, folding).  Here the two arrays of the schoolbook rows PERSIST
  // across terms -- E[k] at limb position k, O[k] at position k + 1, both 64-bit aligned for IMAD.WIDE -- every row
  // of every term multiply-adds straight into them, and the one carry that leaves each chain (2N per term) is
  // counted in C[q] (limb position N + q) instead of rippling upwards.  Nothing is merged or reduced until
  // sum_reduce: per term N^2 wide multiply-adds + 2N carry adds.  Exact for up to 2^30 terms (C and the top limb
  // count terms); the result is the canonical residue of sum_k a_k b_k / R, the same as reduce-every-product.
  struct Sum { uint32_t E[2 * N], O[2 * N], C[N + 1]; };
  LCPC_DEV static Sum sum_zero() {
    Sum s;
#pragma unroll
    for (int i = 0; i < 2 * N; i++) s.E[i] = 0, s.O[i] = 0;
#pragma unroll
    for (int i = 0; i <= N; i++) s.C[i] = 0;
    return s;
  }
  // d += carry, kept on the ALU pipe (the multiplier pipe is the one these loops saturate)
  LCPC_DEV static void count_carry(uint32_t &d) {
#ifdef __CUDA_ARCH__
    asm volatile("addc.u32 %0, %0, %1;" : "+r"(d) : "r"(kZero32));
#else
    add_carry(d);
#endif
  }
  LCPC_DEV static void sum_mac(Sum &s, const Elem &a, const Elem &b) {
#pragma unroll
    for (int i = 0; i < N; i++) {
      uint32_t *Y = (i & 1) ? &s.O[i - 1] : &s.E[i];  // limb position i
      uint32_t *X = (i & 1) ? &s.E[i + 1] : &s.O[i];  // limb position i + 1
      const uint32_t bi = b.v[i];
      mad_wide_cc(Y[0], Y[1], a.v[0], bi, Y[0], Y[1]);
#pragma unroll
      for (int j = 2; j < N; j += 2) madc_wide_cc(Y[j], Y[j + 1], a.v[j], bi, Y[j], Y[j + 1]);
      count_carry(s.C[i]);  // out of position i + N - 1 into position i + N
      mad_wide_cc(X[0], X[1], a.v[1], bi, X[0], X[1]);
#pragma unroll
      for (int j = 3; j < N; j += 2) madc_wide_cc(X[j - 1], X[j], a.v[j], bi, X[j - 1], X[j]);
      count_carry(s.C[i + 1]);  // into position i + N + 1
    }
  }
  // the sum as ONE integer of 2N+1 limbs (the top limb counts terms): what lanes that split a sum add up
  struct Collected { uint32_t v[2 * N + 1]; };
  LCPC_DEV static Collected sum_collect(const Sum &s) {
    Collected t;
    uint32_t *T = t.v;
    T[0] = s.E[0];
    add_cc(T[1], s.E[1], s.O[0]);
#pragma unroll
    for (int k = 2; k < 2 * N; k++) addc_cc(T[k], s.E[k], s.O[k - 1]);
    addc(T[2 * N], 0u, 0u);  // O[2N - 1] is never written
    add_cc(T[N], T[N], s.C[0]);
#pragma unroll
    for (int q = 1; q < N; q++) addc_cc(T[N + q], T[N + q], s.C[q]);
    addc(T[2 * N], T[2 * N], s.C[N]);
    return t;
  }
  LCPC_DEV static void collected_add(Collected &a, const Collected &b) {
    add_cc(a.v[0], a.v[0], b.v[0]);
#pragma unroll
    for (int k = 1; k < 2 * N; k++) addc_cc(a.v[k], a.v[k], b.v[k]);
    addc(a.v[2 * N], a.v[2 * N], b.v[2 * N]);
  }
  // (sum / R) mod p in [0, p)
  LCPC_DEV static Elem sum_reduce(const Sum &s) { return collected_reduce(sum_collect(s)); }
  LCPC_DEV static Elem collected_reduce(const Collected &ct) {
    const uint32_t *T = ct.v;
    // REDC: (T + q p) / 2^(32N) = redc_low(low half) + high half, an (N+1)-limb value V whose top limb counts terms
    Elem r = redc_low(T);
    uint32_t V[N + 1];
    add_cc(V[0], r.v[0], T[N]);
#pragma unroll
    for (int i = 1; i < N; i++) addc_cc(V[i], r.v[i], T[N + i]);
    addc(V[N], T[2 * N], 0u);
    // V mod p: quotient estimate from the top two limbs against the modulus' top limb + 1 (never too large, at most
    // one too small: the top limb of every modulus here is >= 2^30 and V / p < 2^24), one conditional subtraction
    const uint64_t top = ((uint64_t)V[N] << 32) | V[N - 1];
    const uint32_t qhat = (uint32_t)(top / ((uint64_t)FP::P(N - 1) + 1u));
    uint32_t Q[N + 1];
    uint64_t carry = 0;
#pragma unroll
    for (int k = 0; k < N; k++) {
      const uint64_t pr = (uint64_t)qhat * FP::P(k) + carry;
      Q[k] = (uint32_t)pr;
      carry = pr >> 32;
    }
    Q[N] = (uint32_t)carry;
    sub_cc(V[0], V[0], Q[0]);
#pragma unroll
    for (int k = 1; k < N; k++) subc_cc(V[k], V[k], Q[k]);
    subc(V[N], V[N], Q[N]);  // 0 now: V < 2p < 2^(32N)
    Elem t;
#pragma unroll
    for (int i = 0; i < N; i++) t.v[i] = V[i];
    return cond_sub_p(t);
  }

  // Montgomery reduction of the low half: (lo + q p) / 2^(32N) for the q that clears the low N limbs;
  // the result is <= p.  Rows of q*p are laid out as in mul_full_n (operand p, digits found on the fly);
  // `c` is the carry that the two arrays still owe to the position being cleared.
  LCPC_DEV static Elem redc_low(const uint32_t *lo) {
    uint32_t E[2 * N], O[2 * N];  // E[k] at limb position k, O[k] at position k + 1
    uint32_t pl[N];
#pragma unroll
    for (int j = 0; j < N; j++) { E[j] = lo[j]; pl[j] = PM(j); }
    uint32_t c = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {
      uint32_t *Y = (i & 1) ? &O[i - 1] : &E[i];
      uint32_t *X = (i & 1) ? &E[i + 1] : &O[i];
      const uint32_t w = i == 0 ? 0u : ((i & 1) ? E[i] : O[i - 1]);  // the other array's limb at position i
      const uint32_t v = Y[0] + w + c;
#ifdef __CUDA_ARCH__
      const uint32_t m = v * kMontInv32;
#else
      const uint32_t m = 0u - v;
#endif
      if (i == 0) mad_row<N, true, true>(X, Y, pl, m);
      else mad_row<N, false, true>(X, Y, pl, m);
      const uint32_t rest = Y[0] | w | c;  // position i now sums to 0 or 2^32
      c = rest != 0u ? 1u : 0u;
    }
    Elem t;
    uint32_t dummy;
    add_cc(dummy, c, 0xffffffffu);  // carry flag <- c
    (void)dummy;
#pragma unroll
    for (int k = 0; k < N - 1; k++) addc_cc(t.v[k], E[N + k], O[N + k - 1]);
    addc(t.v[N - 1], E[2 * N - 1], 0u);
    return t;
  }

  // REDC of a double-width value t < 2^(64N-1): t * R^{-1} mod p in [0, p).  SUBS = how many conditional
  // subtractions of p the bound on t needs (1 for a single product a*b with a, b < p; 2 for folded sums)
  template <int SUBS>
  LCPC_DEV static Elem redc(const Wide &t) {
    Elem r = redc_low(t.v);
    add_cc(r.v[0], r.v[0], t.v[N]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) addc_cc(r.v[i], r.v[i], t.v[N + i]);
    addc(r.v[N - 1], r.v[N - 1], t.v[2 * N - 1]);
#pragma unroll
    for (int s = 0; s < SUBS; s++) r = cond_sub_p(r);
    return r;
  }

  // ---- one-level subtractive Karatsuba product (N a multiple of 4) ----
  // a = a0 + a1 W, b = b0 + b1 W, W = 2^(16N):  a b = z0 + (z0 + z2 + (a0 - a1)(b1 - b0)) W + z2 W^2 with
  // z0 = a0 b0, z2 = a1 b1: three half-size schoolbook products (3 N^2/4 wide IMADs instead of N^2) paid for
  // with ~9N adds/logic ops on the ALU pipe, which the multiplier-bound kernels leave mostly idle.
  // The middle term is formed from |a0 - a1| |b1 - b0| and the product's sign; it is non-negative and
  // < 2^(32N+1), so the (N+1)-limb two's-complement sum below is exact.
  LCPC_DEV static Wide mul_full_karatsuba(const Elem &a, const Elem &b) {
    static_assert(N % 4 == 0, "Karatsuba split needs an even number of limbs per half");
    constexpr int H = N / 2;
    Wide t;
    const uint32_t Z = zero_operand();
    mul_full_n<H>(t.v, a.v, b.v);              // z0 at limbs 0..N-1
    mul_full_n<H>(t.v + N, a.v + H, b.v + H);  // z2 at limbs N..2N-1
    uint32_t da[H], db[H], z1[N], sa, sb;
    sub_cc(da[0], a.v[0], a.v[H]);
#pragma unroll
    for (int i = 1; i < H; i++) subc_cc(da[i], a.v[i], a.v[H + i]);
    subc(sa, Z, Z);  // all ones iff a0 < a1
    sub_cc(db[0], b.v[H], b.v[0]);
#pragma unroll
    for (int i = 1; i < H; i++) subc_cc(db[i], b.v[H + i], b.v[i]);
    subc(sb, Z, Z);  // all ones iff b1 < b0
    // |x| = (x ^ s) - s over H limbs (s = 0 or -1)
    sub_cc(da[0], da[0] ^ sa, sa);
#pragma unroll
    for (int i = 1; i < H - 1; i++) subc_cc(da[i], da[i] ^ sa, sa);
    subc(da[H - 1], da[H - 1] ^ sa, sa);
    sub_cc(db[0], db[0] ^ sb, sb);
#pragma unroll
    for (int i = 1; i < H - 1; i++) subc_cc(db[i], db[i] ^ sb, sb);
    subc(db[H - 1], db[H - 1] ^ sb, sb);
    mul_full_n<H>(z1, da, db);
    const uint32_t neg = sa ^ sb;  // all ones iff (a0 - a1)(b1 - b0) < 0
    uint32_t mid[N + 1];
    add_cc(mid[0], t.v[0], t.v[N]);
#pragma unroll
    for (int i = 1; i < N; i++) addc_cc(mid[i], t.v[i], t.v[N + i]);
    addc(mid[N], Z, Z);
    // mid += neg ? -z1 : z1   (two's complement over N+1 limbs: ~z1 + 1 with the sign limb -1)
    uint32_t dummy;
    add_cc(dummy, neg, neg);  // carry flag <- neg & 1
    (void)dummy;
#pragma unroll
    for (int i = 0; i < N; i++) addc_cc(mid[i], mid[i], z1[i] ^ neg);
    addc(mid[N], mid[N], neg);
    // t += mid * 2^(32H)
    add_cc(t.v[H], t.v[H], mid[0]);
#pragma unroll
    for (int i = 1; i <= N; i++) addc_cc(t.v[H + i], t.v[H + i], mid[i]);
#pragma unroll
    for (int i = H + N + 1; i < 2 * N - 1; i++) addc_cc(t.v[i], t.v[i], Z);
    if (H + N + 1 <= 2 * N - 1) addc(t.v[2 * N - 1], t.v[2 * N - 1], Z);
    return t;
  }
  LCPC_DEV static Elem mul_karatsuba(const Elem &a, const Elem &b) { return redc<1>(mul_full_karatsuba(a, b)); }

  // canonical integer of a Montgomery-form element: a * R^{-1} mod p  (what to_repr serialises,
  // reference: FieldHash::digest_update, lcpc-2d/src/lib.rs:42-57)
  LCPC_DEV static Elem from_mont(const Elem &a) { return cond_sub_p(redc_low(a.v)); }
};

}  // namespace lcpc
