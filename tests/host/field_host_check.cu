// tests/host/field_host_check.cu -- TEST ONLY.
// Runs lcpc_b200/csrc/field.cuh's carry-chain algorithms on the HOST (through the header's flag
// emulation) so their logic can be compared with the oracle in the CPU-only container.  The CUDA
// build of the very same header is compared with the oracle on the GPU in tests/test_gpu_parity.py.
#include <cstddef>
#include <cstdint>
#include <cstring>

#include "../../lcpc_b200/csrc/field.cuh"

using namespace lcpc;

template <int FID>
static int run(int op, uint64_t *r, const uint64_t *a, const uint64_t *b, size_t n) {
  using F = Field<FID>;
  typename F::Elem x, y, z;
  const size_t terms = (size_t)op >> 8;
  op &= 0xff;
  for (size_t i = 0; i < n; i++) {
    memcpy(x.v, a + i * (F::N / 2), F::BYTES);
    if (b) memcpy(y.v, b + i * (F::N / 2), F::BYTES);
    switch (op) {
      case 0: z = F::add(x, y); break;
      case 1: z = F::sub(x, y); break;
      case 2: z = F::mul(x, y); break;
      case 4: z = F::from_mont(x); break;
      case 5: z = F::template redc<1>(F::mul_full(x, y)); break;  // product and reduction done separately
      case 6: {  // lazily reduced sum of LAZY_TERMS products starting at element i (indices wrap)
        typename F::Wide acc = F::wide_zero();
        for (size_t k = 0; k < 37; k++) {
          typename F::Elem u, v;
          memcpy(u.v, a + ((i + k) % n) * (F::N / 2), F::BYTES);
          memcpy(v.v, b + ((i * 7 + k) % n) * (F::N / 2), F::BYTES);
          F::mac_wide(acc, u, v);
        }
        z = F::template redc<2>(acc);
        break;
      }
      case 8: {  // the same sum through the carry-counting accumulator (sum_mac / sum_reduce), n_terms in op >> 8
        typename F::Sum acc = F::sum_zero();
        for (size_t k = 0; k < terms; k++) {
          typename F::Elem u, v;
          memcpy(u.v, a + ((i + k) % n) * (F::N / 2), F::BYTES);
          memcpy(v.v, b + ((i * 7 + k) % n) * (F::N / 2), F::BYTES);
          F::sum_mac(acc, u, v);
        }
        z = F::sum_reduce(acc);
        break;
      }
      case 10: {  // the same sum split over 4 accumulators (what spmm_sum_split_kernel's lanes do), collected and added
        typename F::Sum part[4];
        for (int g = 0; g < 4; g++) part[g] = F::sum_zero();
        for (size_t k = 0; k < terms; k++) {
          typename F::Elem u, v;
          memcpy(u.v, a + ((i + k) % n) * (F::N / 2), F::BYTES);
          memcpy(v.v, b + ((i * 7 + k) % n) * (F::N / 2), F::BYTES);
          F::sum_mac(part[k % 4], u, v);
        }
        typename F::Collected t = F::sum_collect(part[0]);
        for (int g = 1; g < 4; g++) F::collected_add(t, F::sum_collect(part[g]));
        z = F::collected_reduce(t);
        break;
      }
      case 9: z = F::one(); break;  // R mod p folded at compile time
      case 7:  // Karatsuba product (fields with a multiple of 4 limbs; others fall back to mul)
        if constexpr (F::N % 4 == 0) z = F::mul_karatsuba(x, y);
        else z = F::mul(x, y);
        break;
      default: return -1;
    }
    memcpy(r + i * (F::N / 2), z.v, F::BYTES);
  }
  return 0;
}

extern "C" int hostcheck_field_op(int field, int op, uint64_t *r, const uint64_t *a, const uint64_t *b, size_t n) {
  switch (field) {
    case FT63: return run<FT63>(op, r, a, b, n);
    case FT127: return run<FT127>(op, r, a, b, n);
    case FT191: return run<FT191>(op, r, a, b, n);
    case FT255: return run<FT255>(op, r, a, b, n);
  }
  return -2;
}
