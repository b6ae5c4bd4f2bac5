// lcpc_b200/csrc/kernels_verify.cu -- the verifier's per-column checks, batched over all opened columns, and the
// final inner product.
//
// Reference: verify() step 3 (lcpc-2d/src/lib.rs:926-949) runs, for each of the n_col_opens columns,
//   verify_column_value (:992-1000)  <tensor, column> == encoded-row[col]   once per degree test and once for p_eval
//   verify_column_path  (:954-989)   leaf = D(0^32 || repr(col[0]) || ...), then log2(n_cols) parent hashes up to root
// and step 4 (:944-951) returns <inner_tensor, p_eval>.  On the device the opened columns are turned into the
// row-major matrix [n_rows][n_open], so the leaf digests are one launch of the commit's own column-hash kernels and
// every tensor's dot products are one launch of collapse_kernel (kernels_hash.cu, kernels_collapse.cu); what is left
// for this file is the transpose, the per-column comparison + Merkle walk, and the dot product.
#include <algorithm>

#include "blake3.cuh"
#include "field.cuh"
#include "kernels.h"

namespace lcpc {

// out[r][j] = in[j][r]: opened columns (each contiguous, LcColumn::col) -> row-major matrix; W-limb pieces
template <typename V>
__global__ void transpose_columns_kernel(const V *__restrict__ in, V *__restrict__ out, size_t n_open, size_t n_rows,
                                         unsigned per_elem) {
  const size_t total = n_open * n_rows * per_elem;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const unsigned piece = (unsigned)(i % per_elem);
    const size_t e = i / per_elem, j = e % n_open, r = e / n_open;
    out[i] = in[(j * n_rows + r) * per_elem + piece];
  }
}

cudaError_t launch_transpose_columns(int field, const uint32_t *in, uint32_t *out, size_t n_open, size_t n_rows,
                                     cudaStream_t stream) {
  const int N = field_limbs32(field);
  if (N < 0) return cudaErrorInvalidValue;
  if (n_open == 0 || n_rows == 0) return cudaSuccess;
  const unsigned grid = 148 * 4;
  if (N % 4 == 0)
    transpose_columns_kernel<uint4><<<grid, 256, 0, stream>>>((const uint4 *)in, (uint4 *)out, n_open, n_rows, N / 4);
  else
    transpose_columns_kernel<uint2><<<grid, 256, 0, stream>>>((const uint2 *)in, (uint2 *)out, n_open, n_rows, N / 2);
  return cudaGetLastError();
}

// One thread per opened column j (column number cols[j]):
//   bit 0 of flags[j]: every degree-test value matches   evals[k][j] == rows[k][cols[j]],  k < n_tensors - 1
//   bit 1:             the evaluation value matches       (k = n_tensors - 1)
//   bit 2:             leaves[j] hashed up along paths[j][0..path_len) gives `root`
template <int N>
__global__ void check_columns_kernel(const uint32_t *__restrict__ evals, const uint32_t *__restrict__ rows,
                                     size_t row_stride, unsigned n_tensors, const uint64_t *__restrict__ cols,
                                     size_t n_open, const uint32_t *__restrict__ leaves,
                                     const uint32_t *__restrict__ paths, unsigned path_len,
                                     const uint32_t *__restrict__ root, uint32_t *__restrict__ flags) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_open) return;
  const size_t col = cols[j];
  uint32_t out = 0, degree_diff = 0, eval_diff = 0;
  for (unsigned k = 0; k < n_tensors; k++) {
    const uint32_t *a = evals + ((size_t)k * n_open + j) * N;
    const uint32_t *b = rows + ((size_t)k * row_stride + col) * N;
    uint32_t d = 0;
#pragma unroll
    for (int l = 0; l < N; l++) d |= a[l] ^ b[l];
    if (k + 1 == n_tensors) eval_diff |= d;
    else degree_diff |= d;
  }
  if (!degree_diff) out |= 1u;
  if (!eval_diff) out |= 2u;
  // verify_column_path (:971-986): even position -> D(hash || sibling), odd -> D(sibling || hash)
  uint32_t h[8];
#pragma unroll
  for (int q = 0; q < 8; q++) h[q] = leaves[j * 8 + q];
  size_t pos = col;
  for (unsigned l = 0; l < path_len; l++) {
    const uint32_t *sib = paths + ((size_t)j * path_len + l) * 8;
    uint32_t m[16];
    const bool right = (pos & 1) != 0;
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const uint32_t s = sib[q];
      m[q] = right ? s : h[q];
      m[8 + q] = right ? h[q] : s;
    }
    b3::set_iv(h);
    b3::compress(h, m, 0, 64, b3::CHUNK_START | b3::CHUNK_END | b3::ROOT);
    pos >>= 1;
  }
  uint32_t diff = 0;
#pragma unroll
  for (int q = 0; q < 8; q++) diff |= h[q] ^ root[q];
  if (!diff) out |= 4u;
  flags[j] = out;
}

cudaError_t launch_check_columns(int field, const uint32_t *evals, const uint32_t *rows, size_t row_stride,
                                 unsigned n_tensors, const uint64_t *cols, size_t n_open, const uint8_t *leaves,
                                 const uint8_t *paths, unsigned path_len, const uint8_t *root, uint32_t *flags,
                                 cudaStream_t stream) {
  if (n_open == 0) return cudaSuccess;
  const unsigned grid = (unsigned)((n_open + 63) / 64);
  const uint32_t *lv = (const uint32_t *)leaves, *pt = (const uint32_t *)paths, *rt = (const uint32_t *)root;
  switch (field_limbs32(field)) {
    case 2: check_columns_kernel<2><<<grid, 64, 0, stream>>>(evals, rows, row_stride, n_tensors, cols, n_open, lv, pt, path_len, rt, flags); break;
    case 4: check_columns_kernel<4><<<grid, 64, 0, stream>>>(evals, rows, row_stride, n_tensors, cols, n_open, lv, pt, path_len, rt, flags); break;
    case 6: check_columns_kernel<6><<<grid, 64, 0, stream>>>(evals, rows, row_stride, n_tensors, cols, n_open, lv, pt, path_len, rt, flags); break;
    case 8: check_columns_kernel<8><<<grid, 64, 0, stream>>>(evals, rows, row_stride, n_tensors, cols, n_open, lv, pt, path_len, rt, flags); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

// ---- <a, b> (verify step 4, :944-951); b == nullptr sums a --------------------------------------------------
// Each thread keeps a double-width lazily reduced sum (field.cuh mac_wide), reduces once, and the CTA adds its
// threads' results through shared memory; out[blockIdx.x] = the CTA's partial sum.
constexpr int DOT_THREADS = 128;
template <int FID>
__global__ void __launch_bounds__(DOT_THREADS)
dot_kernel(const uint32_t *__restrict__ a, const uint32_t *__restrict__ b, size_t n, uint32_t *__restrict__ out) {
  using F = Field<FID>;
  constexpr int N = F::N;
  __shared__ uint32_t part[DOT_THREADS][N + 1];
  typename F::Elem acc;
  if (b) {
    typename F::Wide wide = F::wide_zero();
    for (size_t i = (size_t)blockIdx.x * DOT_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * DOT_THREADS) {
      typename F::Elem x, y;
#pragma unroll
      for (int l = 0; l < N; l++) x.v[l] = a[i * N + l], y.v[l] = b[i * N + l];
      F::mac_wide(wide, x, y);
    }
    acc = F::template redc<2>(wide);
  } else {
    acc = F::zero();
    for (size_t i = (size_t)blockIdx.x * DOT_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * DOT_THREADS) {
      typename F::Elem x;
#pragma unroll
      for (int l = 0; l < N; l++) x.v[l] = a[i * N + l];
      acc = F::add(acc, x);
    }
  }
#pragma unroll
  for (int l = 0; l < N; l++) part[threadIdx.x][l] = acc.v[l];
  __syncthreads();
  for (int s = DOT_THREADS / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      typename F::Elem o;
#pragma unroll
      for (int l = 0; l < N; l++) o.v[l] = part[threadIdx.x + s][l];
      acc = F::add(acc, o);
#pragma unroll
      for (int l = 0; l < N; l++) part[threadIdx.x][l] = acc.v[l];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
#pragma unroll
    for (int l = 0; l < N; l++) out[(size_t)blockIdx.x * N + l] = acc.v[l];
  }
}

template <int FID>
static cudaError_t dot_impl(const uint32_t *a, const uint32_t *b, size_t n, uint32_t *partials, uint32_t *out,
                            cudaStream_t stream, int *n_launches) {
  unsigned ctas = (unsigned)std::min<size_t>(DOT_PARTIALS, (n + DOT_THREADS - 1) / DOT_THREADS);
  if (ctas == 0) ctas = 1;
  dot_kernel<FID><<<ctas, DOT_THREADS, 0, stream>>>(a, b, n, ctas == 1 ? out : partials);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (n_launches) ++*n_launches;
  if (ctas > 1) {
    dot_kernel<FID><<<1, DOT_THREADS, 0, stream>>>(partials, nullptr, ctas, out);
    e = cudaGetLastError();
    if (e == cudaSuccess && n_launches) ++*n_launches;
  }
  return e;
}

cudaError_t launch_dot(int field, const uint32_t *a, const uint32_t *b, size_t n, uint32_t *partials, uint32_t *out,
                       cudaStream_t stream, int *n_launches) {
  if (n_launches) *n_launches = 0;
  switch (field) {
    case FT63: return dot_impl<FT63>(a, b, n, partials, out, stream, n_launches);
    case FT127: return dot_impl<FT127>(a, b, n, partials, out, stream, n_launches);
    case FT191: return dot_impl<FT191>(a, b, n, partials, out, stream, n_launches);
    case FT255: return dot_impl<FT255>(a, b, n, partials, out, stream, n_launches);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace lcpc
