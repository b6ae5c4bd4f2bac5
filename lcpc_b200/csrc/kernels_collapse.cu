// lcpc_b200/csrc/kernels_collapse.cu -- the prover's random linear combination of rows and the
// column gather for openings.
//
// collapse_columns (reference: lcpc-2d/src/lib.rs:1095-1123, call sites :1034,:1055):
//     poly[c] = sum_r tensor[r] * coeffs[r * n_per_row + c]
// The reference reduces every product mod p and adds; here each thread keeps the UNREDUCED
// double-width sum of its products (a 2N+1-limb accumulator), and one Montgomery reduction is done
// per output element -- half the multiplier work of reduce-per-product.  The result is the same
// canonical residue: REDC is linear mod p and the final value is brought into [0, p).
// open_column's strided gather (lcpc-2d/src/lib.rs:802-808) is gather_columns_kernel.
#include <algorithm>

#include "field.cuh"
#include "kernels.h"

namespace lcpc {

constexpr int COL_TILE = 32;   // columns per CTA (one warp wide: coalesced 32*B-byte row reads)
constexpr int ROW_GROUPS = 8;  // row slices per CTA, reduced through shared memory

template <int N>
__device__ __forceinline__ void ld_elem(uint32_t (&v)[N], const uint32_t *p) {
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

template <int FID>
__global__ void __launch_bounds__(COL_TILE *ROW_GROUPS)
collapse_kernel(const uint32_t *__restrict__ coeffs, size_t row_stride, const uint32_t *__restrict__ tensor,
                uint32_t *__restrict__ poly, size_t n_rows, size_t n_per_row) {
  using F = Field<FID>;
  constexpr int N = F::N;
  __shared__ uint32_t part[ROW_GROUPS][N][COL_TILE];
  const unsigned lane = threadIdx.x % COL_TILE, grp = threadIdx.x / COL_TILE;
  const size_t col = (size_t)blockIdx.x * COL_TILE + lane;
  typename F::Wide wide = F::wide_zero();
  if (col < n_per_row) {
    for (size_t r = grp; r < n_rows; r += ROW_GROUPS) {
      typename F::Elem a, t;
      ld_elem<N>(a.v, coeffs + (r * row_stride + col) * N);
      ld_elem<N>(t.v, tensor + r * N);
      F::mac_wide(wide, a, t);
    }
  }
  typename F::Elem acc = F::template redc<2>(wide);
#pragma unroll
  for (int l = 0; l < N; l++) part[grp][l][lane] = acc.v[l];
  __syncthreads();
  if (grp == 0 && col < n_per_row) {
    for (int g = 1; g < ROW_GROUPS; g++) {
      typename F::Elem o;
#pragma unroll
      for (int l = 0; l < N; l++) o.v[l] = part[g][l][lane];
      acc = F::add(acc, o);
    }
    uint32_t *p = poly + col * N;
#pragma unroll
    for (int l = 0; l < N; l++) p[l] = acc.v[l];
  }
}

size_t collapse_scratch_bytes(int, size_t, size_t) { return 0; }

cudaError_t launch_collapse(int field, const uint32_t *coeffs, size_t row_stride, const uint32_t *tensor,
                            uint32_t *poly, size_t n_rows, size_t n_per_row, void *, cudaStream_t stream,
                            int *n_launches) {
  if (n_launches) *n_launches = 0;
  if (n_per_row == 0) return cudaSuccess;
  unsigned grid = (unsigned)((n_per_row + COL_TILE - 1) / COL_TILE);
  const int threads = COL_TILE * ROW_GROUPS;
  switch (field) {
    case FT63: collapse_kernel<FT63><<<grid, threads, 0, stream>>>(coeffs, row_stride, tensor, poly, n_rows, n_per_row); break;
    case FT127: collapse_kernel<FT127><<<grid, threads, 0, stream>>>(coeffs, row_stride, tensor, poly, n_rows, n_per_row); break;
    case FT191: collapse_kernel<FT191><<<grid, threads, 0, stream>>>(coeffs, row_stride, tensor, poly, n_rows, n_per_row); break;
    case FT255: collapse_kernel<FT255><<<grid, threads, 0, stream>>>(coeffs, row_stride, tensor, poly, n_rows, n_per_row); break;
    default: return cudaErrorInvalidValue;
  }
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && n_launches) *n_launches = 1;
  return e;
}

// out[i * n_rows + r] = comm[r * row_stride + cols[i]]
__global__ void gather_columns_kernel(const uint32_t *__restrict__ comm, size_t n_rows, size_t row_stride,
                                      const uint64_t *__restrict__ cols, size_t n_open, uint32_t *__restrict__ out,
                                      int n_limbs) {
  const size_t total = n_open * n_rows * (size_t)n_limbs;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t l = idx % n_limbs, e = idx / n_limbs;
    const size_t r = e % n_rows, i = e / n_rows;
    out[idx] = comm[(r * row_stride + cols[i]) * n_limbs + l];
  }
}

cudaError_t launch_gather_columns(int field, const uint32_t *comm, size_t n_rows, size_t row_stride,
                                  const uint64_t *cols, size_t n_open, uint32_t *out, cudaStream_t stream) {
  int nl = field_limbs32(field);
  if (nl < 0) return cudaErrorInvalidValue;
  size_t total = n_open * n_rows * (size_t)nl;
  if (total == 0) return cudaSuccess;
  unsigned grid = (unsigned)std::min<size_t>((total + 255) / 256, 148 * 16);
  gather_columns_kernel<<<grid, 256, 0, stream>>>(comm, n_rows, row_stride, cols, n_open, out, nl);
  return cudaGetLastError();
}

// pack_column_blocks: 16-byte (or 8-byte) granules, one thread each; reads are fully coalesced, writes are
// coalesced within a tile row
template <typename V>
__global__ void pack_column_blocks_kernel(const V *__restrict__ src, size_t n_rows, size_t n_cols, unsigned per_elem,
                                          unsigned n_blocks, const uint64_t *__restrict__ starts, V *__restrict__ dst) {
  __shared__ uint64_t s_starts[65];
  for (unsigned i = threadIdx.x; i <= n_blocks; i += blockDim.x) s_starts[i] = starts[i];
  __syncthreads();
  const size_t row_granules = n_cols * per_elem, total = n_rows * row_granules;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t r = idx / row_granules, g = idx % row_granules;
    const size_t c = g / per_elem, part = g % per_elem;
    unsigned h = 0;
    while (h + 1 < n_blocks && c >= s_starts[h + 1]) h++;
    const size_t start = s_starts[h], width = s_starts[h + 1] - start;
    dst[(n_rows * start + r * width + (c - start)) * per_elem + part] = src[idx];
  }
}

cudaError_t launch_pack_column_blocks(int field, const uint32_t *src, size_t n_rows, size_t n_cols, unsigned n_blocks,
                                      const uint64_t *d_starts, uint32_t *dst, cudaStream_t stream) {
  int nl = field_limbs32(field);
  if (nl < 0 || n_blocks == 0 || n_blocks > 64) return cudaErrorInvalidValue;
  if (n_rows * n_cols == 0) return cudaSuccess;
  if (nl % 4 == 0) {
    size_t total = n_rows * n_cols * (nl / 4);
    unsigned grid = (unsigned)std::min<size_t>((total + 255) / 256, 148 * 32);
    pack_column_blocks_kernel<uint4><<<grid, 256, 0, stream>>>((const uint4 *)src, n_rows, n_cols, nl / 4, n_blocks,
                                                               d_starts, (uint4 *)dst);
  } else {
    size_t total = n_rows * n_cols * (nl / 2);
    unsigned grid = (unsigned)std::min<size_t>((total + 255) / 256, 148 * 32);
    pack_column_blocks_kernel<uint2><<<grid, 256, 0, stream>>>((const uint2 *)src, n_rows, n_cols, nl / 2, n_blocks,
                                                               d_starts, (uint2 *)dst);
  }
  return cudaGetLastError();
}

// open_column's path walk (lcpc-2d/src/lib.rs:811-821): on layer l the sibling of node (col >> l) is
// (col >> l) ^ 1; layer l starts at offset sum_{q<l} np2 >> q of the flat hashes array.
__global__ void gather_paths_kernel(const uint8_t *__restrict__ hashes, size_t np2, const uint64_t *__restrict__ cols,
                                    size_t n_open, unsigned path_len, uint8_t *__restrict__ out) {
  const size_t total = n_open * path_len * 2;  // 16-byte halves
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t h = idx & 1, e = idx >> 1;
    const size_t l = e % path_len, i = e / path_len;
    size_t off = 0, len = np2;
    for (size_t q = 0; q < l; q++) off += len, len >>= 1;
    const size_t node = (cols[i] >> l) ^ 1;
    reinterpret_cast<uint4 *>(out)[idx] = reinterpret_cast<const uint4 *>(hashes)[(off + node) * 2 + h];
  }
}

cudaError_t launch_gather_paths(const uint8_t *hashes, size_t np2, const uint64_t *cols, size_t n_open,
                                unsigned path_len, uint8_t *out, cudaStream_t stream) {
  size_t total = n_open * path_len * 2;
  if (total == 0) return cudaSuccess;
  unsigned grid = (unsigned)std::min<size_t>((total + 255) / 256, 148 * 16);
  gather_paths_kernel<<<grid, 256, 0, stream>>>(hashes, np2, cols, n_open, path_len, out);
  return cudaGetLastError();
}

}  // namespace lcpc
