// lcpc_b200/csrc/kernels_ntt.cu -- Ligero row encoding: batched radix-2 NTT over shared memory.
//
// Replaces LigeroEncoding::encode (reference: lcpc-ligero-pc/src/lib.rs:162-164), which delegates to
// fffft 0.4 `fft_io_pc`: in-order input, bit-reversed output, decimation in frequency,
//   for gap = n/2 .. 1:  (a, b) <- (a + b, (a - b) * w^(idx * n/(2 gap))),  idx = position mod gap.
// Field arithmetic is exact, so any schedule of those butterflies gives bit-identical limbs.
//
// Schedule: the log2(n) stages are cut into passes of S <= 9 consecutive stages.  One CTA owns a tile
// of 2^S points along the pass's stride times C adjacent points (C*B contiguous bytes in HBM), stages
// it in shared memory, runs the S stages there and writes it back: one HBM read + one write per pass.
// The first pass reads the un-padded coefficient rows directly (implicit zeros beyond n_per_row), so
// the reference's separate pad/copy (lcpc-2d/src/lib.rs:640-651) costs no extra traffic.
#include <algorithm>

#include "field.cuh"
#include "kernels.h"

namespace lcpc {

int field_limbs32(int field) {
  switch (field) {
    case FT63: return 2;
    case FT127: return 4;
    case FT191: return 6;
    case FT255: return 8;
    default: return -1;
  }
}

// ---- global / shared element movement -------------------------------------------------------
// Global: N consecutive limbs.  Shared: split into planes of PW bytes so that consecutive elements
// are consecutive within a plane (conflict-free 64/128-bit accesses for a warp on adjacent elements).
template <int N> struct Planes {
  static constexpr int PW = (N % 4 == 0) ? 4 : 2;  // limbs per plane word (16 B or 8 B)
  static constexpr int NP = N / PW;
};

template <int N>
__device__ __forceinline__ void gload(uint32_t (&v)[N], const uint32_t *p) {
  if constexpr (N % 4 == 0) {
#pragma unroll
    for (int i = 0; i < N / 4; i++) {
      uint4 t = reinterpret_cast<const uint4 *>(p)[i];
      v[4 * i] = t.x, v[4 * i + 1] = t.y, v[4 * i + 2] = t.z, v[4 * i + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < N / 2; i++) {
      uint2 t = reinterpret_cast<const uint2 *>(p)[i];
      v[2 * i] = t.x, v[2 * i + 1] = t.y;
    }
  }
}

template <int N>
__device__ __forceinline__ void gload_ro(uint32_t (&v)[N], const uint32_t *p) {
  if constexpr (N % 4 == 0) {
#pragma unroll
    for (int i = 0; i < N / 4; i++) {
      uint4 t = __ldg(reinterpret_cast<const uint4 *>(p) + i);
      v[4 * i] = t.x, v[4 * i + 1] = t.y, v[4 * i + 2] = t.z, v[4 * i + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < N / 2; i++) {
      uint2 t = __ldg(reinterpret_cast<const uint2 *>(p) + i);
      v[2 * i] = t.x, v[2 * i + 1] = t.y;
    }
  }
}

template <int N>
__device__ __forceinline__ void gstore(uint32_t *p, const uint32_t (&v)[N]) {
  if constexpr (N % 4 == 0) {
#pragma unroll
    for (int i = 0; i < N / 4; i++)
      reinterpret_cast<uint4 *>(p)[i] = make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < N / 2; i++) reinterpret_cast<uint2 *>(p)[i] = make_uint2(v[2 * i], v[2 * i + 1]);
  }
}

template <int N>
__device__ __forceinline__ void sload(uint32_t (&v)[N], const uint32_t *smem, unsigned e, unsigned tile) {
  constexpr int PW = Planes<N>::PW, NP = Planes<N>::NP;
#pragma unroll
  for (int pl = 0; pl < NP; pl++) {
    const uint32_t *q = smem + ((size_t)pl * tile + e) * PW;
    if constexpr (PW == 4) {
      uint4 t = *reinterpret_cast<const uint4 *>(q);
      v[4 * pl] = t.x, v[4 * pl + 1] = t.y, v[4 * pl + 2] = t.z, v[4 * pl + 3] = t.w;
    } else {
      uint2 t = *reinterpret_cast<const uint2 *>(q);
      v[2 * pl] = t.x, v[2 * pl + 1] = t.y;
    }
  }
}

template <int N>
__device__ __forceinline__ void sstore(uint32_t *smem, unsigned e, unsigned tile, const uint32_t (&v)[N]) {
  constexpr int PW = Planes<N>::PW, NP = Planes<N>::NP;
#pragma unroll
  for (int pl = 0; pl < NP; pl++) {
    uint32_t *q = smem + ((size_t)pl * tile + e) * PW;
    if constexpr (PW == 4)
      *reinterpret_cast<uint4 *>(q) = make_uint4(v[4 * pl], v[4 * pl + 1], v[4 * pl + 2], v[4 * pl + 3]);
    else
      *reinterpret_cast<uint2 *>(q) = make_uint2(v[2 * pl], v[2 * pl + 1]);
  }
}

// ---- element-wise test hook -------------------------------------------------------------------
template <int FID>
__global__ void field_op_kernel(int op, uint32_t *r, const uint32_t *a, const uint32_t *b, size_t n) {
  using F = Field<FID>;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  typename F::Elem x, y, z;
  gload<F::N>(x.v, a + i * F::N);
  if (b) gload<F::N>(y.v, b + i * F::N);
  else y = F::zero();
  switch (op) {
    case 0: z = F::add(x, y); break;
    case 1: z = F::sub(x, y); break;
    case 2: z = F::mul(x, y); break;
    case 5: z = F::template redc<1>(F::mul_full(x, y)); break;
    case 6: {  // 37-term lazily reduced sum of products starting at element i (indices wrap)
      typename F::Wide acc = F::wide_zero();
      for (size_t k = 0; k < 37; k++) {
        typename F::Elem u, v;
        gload<F::N>(u.v, a + ((i + k) % n) * F::N);
        gload<F::N>(v.v, b + ((i * 7 + k) % n) * F::N);
        F::mac_wide(acc, u, v);
      }
      z = F::template redc<2>(acc);
      break;
    }
    default: z = F::from_mont(x); break;
  }
  gstore<F::N>(r + i * F::N, z.v);
}

cudaError_t launch_field_op(int field, int op, uint32_t *r, const uint32_t *a, const uint32_t *b, size_t n,
                            cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  unsigned grid = (unsigned)((n + 127) / 128);
  switch (field) {
    case FT63: field_op_kernel<FT63><<<grid, 128, 0, stream>>>(op, r, a, b, n); break;
    case FT127: field_op_kernel<FT127><<<grid, 128, 0, stream>>>(op, r, a, b, n); break;
    case FT191: field_op_kernel<FT191><<<grid, 128, 0, stream>>>(op, r, a, b, n); break;
    case FT255: field_op_kernel<FT255><<<grid, 128, 0, stream>>>(op, r, a, b, n); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

// ---- root table: roots[i] = w^i -----------------------------------------------------------------
// Each thread owns a run of consecutive exponents: w^(start) by square-and-multiply, then steps by w.
template <int FID>
__global__ void root_table_kernel(uint32_t *roots, const uint32_t *w_in, size_t half, unsigned run) {
  using F = Field<FID>;
  size_t start = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * run;
  if (start >= half) return;
  typename F::Elem w, cur, base;
  gload<F::N>(w.v, w_in);
  // Montgomery one = R mod p = from_mont^{-1}(1): obtain it as w^0 via x * R2 ... avoid constants:
  // one = (2^(32N) mod p); computed on the host and passed as w_in[N..2N)
  gload<F::N>(cur.v, w_in + F::N);
  base = w;
  for (size_t e = start; e; e >>= 1) {
    if (e & 1) cur = F::mul(cur, base);
    base = F::mul(base, base);
  }
  size_t end = start + run < half ? start + run : half;
  for (size_t i = start; i < end; i++) {
    gstore<F::N>(roots + i * F::N, cur.v);
    cur = F::mul(cur, w);
  }
}

cudaError_t launch_root_table(int field, uint32_t *roots, const uint32_t *w, size_t half, cudaStream_t stream) {
  if (half == 0) return cudaSuccess;
  const unsigned run = 16;
  size_t threads = (half + run - 1) / run;
  unsigned grid = (unsigned)((threads + 127) / 128);
  switch (field) {
    case FT63: root_table_kernel<FT63><<<grid, 128, 0, stream>>>(roots, w, half, run); break;
    case FT127: root_table_kernel<FT127><<<grid, 128, 0, stream>>>(roots, w, half, run); break;
    case FT191: root_table_kernel<FT191><<<grid, 128, 0, stream>>>(roots, w, half, run); break;
    case FT255: root_table_kernel<FT255><<<grid, 128, 0, stream>>>(roots, w, half, run); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

// ---- one NTT pass -----------------------------------------------------------------------------
struct NttPass {
  unsigned log_n;      // transform length 2^log_n
  unsigned hi;         // this pass runs the stages with gaps 2^(hi-1) .. 2^(hi-S)
  unsigned S;          // stages in this pass
  unsigned logC;       // log2 of adjacent points per tile
  unsigned tiles_per_row_log;  // log2(n / (2^S * C))
  size_t src_stride, src_valid, dst_stride;
};

constexpr int NTT_THREADS = 256;

template <int FID>
__global__ void __launch_bounds__(NTT_THREADS)
ntt_pass_kernel(const uint32_t *__restrict__ src, uint32_t *dst, const uint32_t *__restrict__ roots, NttPass p) {
  using F = Field<FID>;
  constexpr int N = F::N;
  extern __shared__ __align__(16) uint32_t smem[];
  const unsigned S = p.S, logC = p.logC;
  const unsigned tile = 1u << (S + logC);
  const size_t row = blockIdx.x >> p.tiles_per_row_log;
  const size_t tid_in_row = blockIdx.x & ((1u << p.tiles_per_row_log) - 1);
  const bool last = (p.hi == S);  // stride-1 pass: the tile is C contiguous sub-transforms
  // element (t, c) of the tile sits at  base + t * tstride + c * cstride  in the row
  size_t base, tstride;
  unsigned tshift, cshift;  // smem index e = (t << tshift) | (c << cshift)
  if (last) {
    base = tid_in_row << (S + logC);
    tstride = 1;
    tshift = 0, cshift = S;
  } else {
    const unsigned log_stride = p.hi - S;
    const size_t lowblocks_log = log_stride - logC;
    const size_t high = tid_in_row >> lowblocks_log;
    const size_t lowblock = tid_in_row & (((size_t)1 << lowblocks_log) - 1);
    base = (high << p.hi) + (lowblock << logC);
    tstride = (size_t)1 << log_stride;
    tshift = logC, cshift = 0;
  }
  const unsigned tmask = (1u << S) - 1, cmask = (1u << logC) - 1;
  const uint32_t *srow = src + row * p.src_stride * N;
  uint32_t *drow = dst + row * p.dst_stride * N;

  // load: smem index e enumerates the tile in its HBM-contiguous order
  for (unsigned e = threadIdx.x; e < tile; e += NTT_THREADS) {
    unsigned t = (e >> tshift) & tmask, c = (e >> cshift) & cmask;
    size_t j = last ? base + e : base + t * tstride + c;
    typename F::Elem x;
    if (j < p.src_valid) gload<N>(x.v, srow + j * N);
    else x = F::zero();
    sstore<N>(smem, e, tile, x.v);
  }
  __syncthreads();

  const unsigned half = tile >> 1;
  for (unsigned s = 0; s < S; s++) {
    const unsigned lg = S - 1 - s;          // log2 of the gap in t units
    const unsigned logG = p.hi - 1 - s;     // log2 of the gap in row positions
    const unsigned twshift = p.log_n - 1 - logG;
    for (unsigned w = threadIdx.x; w < half; w += NTT_THREADS) {
      unsigned q, c;
      if (last) q = w & ((1u << (S - 1)) - 1), c = w >> (S - 1);
      else c = w & cmask, q = w >> logC;
      unsigned t_lo = ((q >> lg) << (lg + 1)) | (q & ((1u << lg) - 1));
      unsigned e_lo = (t_lo << tshift) | (c << cshift);
      unsigned e_hi = e_lo + ((1u << lg) << tshift);
      typename F::Elem a, b;
      sload<N>(a.v, smem, e_lo, tile);
      sload<N>(b.v, smem, e_hi, tile);
      typename F::Elem sum = F::add(a, b);
      typename F::Elem dif = F::sub(a, b);
      if (logG != 0) {  // the gap-1 stage multiplies by w^0 = 1 only
        size_t j_lo = last ? base + ((size_t)c << S) + t_lo : base + t_lo * tstride + c;
        size_t tw = (j_lo & (((size_t)1 << logG) - 1)) << twshift;
        typename F::Elem wv;
        gload_ro<N>(wv.v, roots + tw * N);
        dif = F::mul(dif, wv);
      }
      sstore<N>(smem, e_lo, tile, sum.v);
      sstore<N>(smem, e_hi, tile, dif.v);
    }
    __syncthreads();
  }

  for (unsigned e = threadIdx.x; e < tile; e += NTT_THREADS) {
    unsigned t = (e >> tshift) & tmask, c = (e >> cshift) & cmask;
    size_t j = last ? base + e : base + t * tstride + c;
    typename F::Elem x;
    sload<N>(x.v, smem, e, tile);
    gstore<N>(drow + j * N, x.v);
  }
}

template <int FID>
static cudaError_t ntt_rows_impl(const uint32_t *src, size_t src_stride, size_t src_valid, uint32_t *dst,
                                 size_t dst_stride, const uint32_t *roots, unsigned log_n, size_t n_rows,
                                 cudaStream_t stream, int *n_launches) {
  using F = Field<FID>;
  static bool attr_set = false;
  constexpr unsigned LOG_TILE = 11;  // 2048 elements: 64 KiB for Ft255 -> 3 CTAs/SM
  constexpr unsigned MAX_S = 9;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(ntt_pass_kernel<FID>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)((1u << LOG_TILE) * F::BYTES));
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  if (n_launches) *n_launches = 0;
  if (n_rows == 0) return cudaSuccess;
  if (log_n == 0) {  // length-1 transform: identity
    if (src != dst)
      return cudaMemcpy2DAsync(dst, dst_stride * F::BYTES, src, src_stride * F::BYTES, F::BYTES, n_rows,
                               cudaMemcpyDeviceToDevice, stream);
    return cudaSuccess;
  }
  unsigned n_pass, S_first;
  if (log_n <= LOG_TILE) n_pass = 1;
  else n_pass = (log_n + MAX_S - 1) / MAX_S;
  unsigned hi = log_n;
  const uint32_t *cur_src = src;
  size_t cur_stride = src_stride, cur_valid = src_valid;
  for (unsigned ip = 0; ip < n_pass; ip++) {
    unsigned remaining = n_pass - ip;
    unsigned S = (hi + remaining - 1) / remaining;  // balanced split, larger passes first
    (void)S_first;
    NttPass p;
    p.log_n = log_n, p.hi = hi, p.S = S;
    unsigned logC = S >= LOG_TILE ? 0 : LOG_TILE - S;
    bool last = (hi == S);
    unsigned avail = last ? log_n - S : hi - S;  // log2 of adjacent points (or sub-transforms) available
    if (logC > avail) logC = avail;
    p.logC = logC;
    p.tiles_per_row_log = log_n - S - logC;
    p.src_stride = cur_stride, p.src_valid = cur_valid, p.dst_stride = dst_stride;
    size_t grid = n_rows << p.tiles_per_row_log;
    if (grid > 0x7fffffffu) return cudaErrorInvalidValue;
    size_t smem = ((size_t)1 << (S + logC)) * F::BYTES;
    ntt_pass_kernel<FID><<<(unsigned)grid, NTT_THREADS, smem, stream>>>(cur_src, dst, roots, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (n_launches) ++*n_launches;
    cur_src = dst, cur_stride = dst_stride, cur_valid = (size_t)1 << log_n;
    hi -= S;
  }
  return cudaSuccess;
}

cudaError_t launch_ntt_rows(int field, const uint32_t *src, size_t src_stride, size_t src_valid, uint32_t *dst,
                            size_t dst_stride, const uint32_t *roots, unsigned log_n, size_t n_rows,
                            cudaStream_t stream, int *n_launches) {
  switch (field) {
    case FT63: return ntt_rows_impl<FT63>(src, src_stride, src_valid, dst, dst_stride, roots, log_n, n_rows, stream, n_launches);
    case FT127: return ntt_rows_impl<FT127>(src, src_stride, src_valid, dst, dst_stride, roots, log_n, n_rows, stream, n_launches);
    case FT191: return ntt_rows_impl<FT191>(src, src_stride, src_valid, dst, dst_stride, roots, log_n, n_rows, stream, n_launches);
    case FT255: return ntt_rows_impl<FT255>(src, src_stride, src_valid, dst, dst_stride, roots, log_n, n_rows, stream, n_launches);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace lcpc
