// lcpc_b200/csrc/kernels_hash.cu -- per-column BLAKE3 leaf digests and the Merkle layers.
//
// Replaces hash_columns / merkle_tree / merkle_layer (reference: lcpc-2d/src/lib.rs:706-785):
//   leaf[c] = D( 0x00*32 || to_repr(comm[0][c]) || ... || to_repr(comm[n_rows-1][c]) )   (:719-735)
//   node    = D( left || right )                                                          (:768-775)
// with D = BLAKE3 and to_repr = canonical little-endian bytes (FieldHash::digest_update, :42-57), so
// every element is taken out of Montgomery form on the fly, in registers, right after its load.
//
// Layout / parallelism: the leaf input of one column is Ltot = 32 + B*n_rows bytes = ceil(Ltot/1024)
// BLAKE3 chunks.  Chunks are independent until the tree merge, so the grid is (columns x chunks):
// adjacent lanes own adjacent columns, which makes every row read a fully coalesced run of
// 32*B bytes per warp; a second small kernel merges the chunk chaining values per column.
#include <algorithm>

#include "blake3.cuh"
#include "field.cuh"
#include "kernels.h"

namespace lcpc {

constexpr int HASH_THREADS = 128;
#ifndef LCPC_LEAF_PREFETCH
#define LCPC_LEAF_PREFETCH 0
#endif

template <int N>
__device__ __forceinline__ void gload_elem(uint32_t (&v)[N], const uint32_t *p) {
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

// Chunk `k` of column `col`.  Works in "slots" of one element (B bytes): slot s of the leaf input is
// zero for s < 32/B (the 32-byte zero prefix, lib.rs:722-723) and row s - 32/B after that.
// Requires B | 64 (Ft63, Ft127, Ft255).
template <int FID, bool SHORT_SOURCE = false>
__global__ void __launch_bounds__(HASH_THREADS)
leaf_chunk_kernel(const uint32_t *__restrict__ comm, size_t n_rows, size_t n_cols, size_t row_stride,
                  uint32_t *__restrict__ out, unsigned n_chunks, unsigned k_first, size_t out_cols, size_t col0,
                  size_t src_total, size_t col_stride) {
  using F = Field<FID>;
  constexpr int N = F::N;
  constexpr int B = F::BYTES;
  static_assert(64 % B == 0, "element must divide the BLAKE3 block");
  constexpr int SPB = 64 / B;        // slots per block
  constexpr int PRE = 32 / B;        // zero-prefix slots
  const size_t col = (size_t)blockIdx.x * HASH_THREADS + threadIdx.x;
  const unsigned k = k_first + blockIdx.y;
  if (col >= n_cols) return;
  const size_t total = 32 + (size_t)B * n_rows;
  const size_t chunk_off = (size_t)k * b3::CHUNK_LEN;
  const size_t chunk_len = (total - chunk_off < (size_t)b3::CHUNK_LEN) ? total - chunk_off : (size_t)b3::CHUNK_LEN;
  const unsigned n_blocks = (unsigned)((chunk_len + 63) / 64);
  uint32_t cv[8];
  b3::set_iv(cv);
  const uint32_t *cp = comm + col * col_stride * N;  // element (row r, column c) at (r * row_stride + c * col_stride)
  // rows of this column that exist in the source: all of them, unless the source ends inside the last row (element
  // (r, c) exists iff r * row_stride + c < src_total; beyond that the commit's zero padding is hashed)
  // (compiled in only for sources that can end early: the whole-commit path pays nothing for it)
  const size_t rows_here = !SHORT_SOURCE ? n_rows
                           : (src_total > col ? min(n_rows, (src_total - col + row_stride - 1) / row_stride) : 0);
#if LCPC_LEAF_PREFETCH
  // The elements of block b + 1 are requested before block b is compressed (from_mont(0) = 0, so slots outside the
  // column -- the zero prefix, the padding -- are just zero elements): without this every block starts by waiting for
  // its own loads, which is where 60 % of the kernel's stall samples sat (profiles/r02_ncu_hot_leaf_chunk_ft255.txt).
  // Measured (profiles/r02_ab_leaf_prefetch.jsonl): 0.748 against 0.734 ms (Ft255, 2^24), 0.285 against 0.277 ms
  // (Ft127): the second element buffer costs 16-19 registers, i.e. two of the ten resident CTAs per SM, and the
  // other warps were already covering those waits (ALU pipe 77-84 % busy).  Off.
  typename F::Elem nxt[SPB];
  auto request = [&](unsigned b) {
    const size_t slot0 = (chunk_off + (size_t)b * 64) / B;
#pragma unroll
    for (int i = 0; i < SPB; i++) {
      const size_t slot = slot0 + i;
      nxt[i] = F::zero();
      if (slot >= PRE && slot - PRE < rows_here) gload_elem<N>(nxt[i].v, cp + (slot - PRE) * row_stride * N);
    }
  };
  request(0);
  for (unsigned b = 0; b < n_blocks; b++) {
    uint32_t m[16];
#pragma unroll
    for (int i = 0; i < SPB; i++) {
      const typename F::Elem x = F::from_mont(nxt[i]);
#pragma unroll
      for (int l = 0; l < N; l++) m[i * N + l] = x.v[l];
    }
    if (b + 1 < n_blocks) request(b + 1);
    const bool lastb = (b + 1 == n_blocks);
    const uint32_t block_len = lastb ? (uint32_t)(chunk_len - (size_t)b * 64) : 64u;
    uint32_t flags = (b == 0 ? b3::CHUNK_START : 0u) | (lastb ? b3::CHUNK_END : 0u);
    if (lastb && n_chunks == 1) flags |= b3::ROOT;
    b3::compress(cv, m, k, block_len, flags);
  }
#else
  for (unsigned b = 0; b < n_blocks; b++) {
    uint32_t m[16];
    const size_t slot0 = (chunk_off + (size_t)b * 64) / B;
#pragma unroll
    for (int i = 0; i < SPB; i++) {
      const size_t slot = slot0 + i;
      typename F::Elem x = F::zero();
      if (slot >= PRE && slot - PRE < rows_here) {
        typename F::Elem raw;
        gload_elem<N>(raw.v, cp + (slot - PRE) * row_stride * N);
        x = F::from_mont(raw);  // (positions past src_total are the zero padding of a short last row: canonical zero)
      }
#pragma unroll
      for (int l = 0; l < N; l++) m[i * N + l] = x.v[l];
    }
    const bool lastb = (b + 1 == n_blocks);
    const uint32_t block_len = lastb ? (uint32_t)(chunk_len - (size_t)b * 64) : 64u;
    uint32_t flags = (b == 0 ? b3::CHUNK_START : 0u) | (lastb ? b3::CHUNK_END : 0u);
    if (lastb && n_chunks == 1) flags |= b3::ROOT;
    b3::compress(cv, m, k, block_len, flags);
  }
#endif
  // single chunk: this is the digest; else the chunk chaining value for the merge kernel
  uint32_t *o = out + ((size_t)k * out_cols + col0 + col) * 8;
  reinterpret_cast<uint4 *>(o)[0] = make_uint4(cv[0], cv[1], cv[2], cv[3]);
  reinterpret_cast<uint4 *>(o)[1] = make_uint4(cv[4], cv[5], cv[6], cv[7]);
}

// Ft191 (B = 24 does not divide 64): one generic word-granular walk; each block converts the <= 4
// elements it overlaps.  Kept simple: no reference test or bench uses this field with a commit.
template <int FID>
__global__ void __launch_bounds__(HASH_THREADS)
leaf_chunk_kernel_generic(const uint32_t *__restrict__ comm, size_t n_rows, size_t n_cols, size_t row_stride,
                          uint32_t *__restrict__ out, unsigned n_chunks, unsigned k_first, size_t out_cols, size_t col0,
                          size_t src_total, size_t col_stride) {
  using F = Field<FID>;
  constexpr int N = F::N;
  constexpr int B = F::BYTES;
  const size_t col = (size_t)blockIdx.x * HASH_THREADS + threadIdx.x;
  const unsigned k = k_first + blockIdx.y;
  if (col >= n_cols) return;
  const size_t total = 32 + (size_t)B * n_rows;
  const size_t chunk_off = (size_t)k * b3::CHUNK_LEN;
  const size_t chunk_len = (total - chunk_off < (size_t)b3::CHUNK_LEN) ? total - chunk_off : (size_t)b3::CHUNK_LEN;
  const unsigned n_blocks = (unsigned)((chunk_len + 63) / 64);
  uint32_t cv[8];
  b3::set_iv(cv);
  const uint32_t *cp = comm + col * col_stride * N;  // element (row r, column c) at (r * row_stride + c * col_stride)
  size_t cached_row = (size_t)-1;
  uint32_t canon[N];
  for (unsigned b = 0; b < n_blocks; b++) {
    uint32_t m[16];
    for (int w = 0; w < 16; w++) {
      const size_t off = chunk_off + (size_t)b * 64 + 4 * (size_t)w;
      uint32_t word = 0;
      if (off >= 32 && off < total) {
        const size_t row = (off - 32) / B;
        const unsigned limb = (unsigned)(((off - 32) % B) / 4);
        if (row != cached_row) {
          typename F::Elem raw = F::zero();
          if (row * row_stride + col < src_total) gload_elem<N>(raw.v, cp + row * row_stride * N);
          typename F::Elem c = F::from_mont(raw);
#pragma unroll
          for (int l = 0; l < N; l++) canon[l] = c.v[l];
          cached_row = row;
        }
#pragma unroll
        for (int l = 0; l < N; l++)
          if (l == (int)limb) word = canon[l];
      }
#pragma unroll
      for (int i = 0; i < 16; i++)
        if (i == w) m[i] = word;
    }
    const bool lastb = (b + 1 == n_blocks);
    const uint32_t block_len = lastb ? (uint32_t)(chunk_len - (size_t)b * 64) : 64u;
    uint32_t flags = (b == 0 ? b3::CHUNK_START : 0u) | (lastb ? b3::CHUNK_END : 0u);
    if (lastb && n_chunks == 1) flags |= b3::ROOT;
    b3::compress(cv, m, k, block_len, flags);
  }
  uint32_t *o = out + ((size_t)k * out_cols + col0 + col) * 8;
  reinterpret_cast<uint4 *>(o)[0] = make_uint4(cv[0], cv[1], cv[2], cv[3]);
  reinterpret_cast<uint4 *>(o)[1] = make_uint4(cv[4], cv[5], cv[6], cv[7]);
}

__device__ __forceinline__ void parent_cv(uint32_t (&out)[8], const uint32_t (&l)[8], const uint32_t (&r)[8],
                                          uint32_t extra_flags) {
  uint32_t m[16];
#pragma unroll
  for (int i = 0; i < 8; i++) m[i] = l[i], m[8 + i] = r[i];
  b3::set_iv(out);
  b3::compress(out, m, 0, 64, b3::PARENT | extra_flags);
}

// BLAKE3 tree over the n_chunks chaining values of each column (chunk-major scratch [k][col][8]).
constexpr int MAX_STACK = 24;
__global__ void __launch_bounds__(HASH_THREADS)
leaf_merge_kernel(const uint32_t *__restrict__ cvs, size_t n_cols, unsigned n_chunks, uint32_t *__restrict__ leaves) {
  const size_t col = (size_t)blockIdx.x * HASH_THREADS + threadIdx.x;
  if (col >= n_cols) return;
  uint32_t stack[MAX_STACK][8];
  int depth = 0;
  uint32_t cur[8];
  for (unsigned k = 0; k < n_chunks; k++) {
    const uint4 *p = reinterpret_cast<const uint4 *>(cvs + ((size_t)k * n_cols + col) * 8);
    uint4 a = __ldg(p), b = __ldg(p + 1);
    cur[0] = a.x, cur[1] = a.y, cur[2] = a.z, cur[3] = a.w, cur[4] = b.x, cur[5] = b.y, cur[6] = b.z, cur[7] = b.w;
    if (k + 1 == n_chunks) break;
    // add_chunk_chaining_value: merge completed subtrees (one per trailing zero bit of the count)
    unsigned total = k + 1;
    while ((total & 1u) == 0) {
      uint32_t merged[8];
      depth--;
      parent_cv(merged, stack[depth], cur, 0);
#pragma unroll
      for (int i = 0; i < 8; i++) cur[i] = merged[i];
      total >>= 1;
    }
#pragma unroll
    for (int i = 0; i < 8; i++) stack[depth][i] = cur[i];
    depth++;
  }
  // finalize: fold the stack from the top, ROOT on the last parent
  while (depth > 0) {
    uint32_t merged[8];
    depth--;
    parent_cv(merged, stack[depth], cur, depth == 0 ? b3::ROOT : 0u);
#pragma unroll
    for (int i = 0; i < 8; i++) cur[i] = merged[i];
  }
  uint32_t *o = leaves + col * 8;
  reinterpret_cast<uint4 *>(o)[0] = make_uint4(cur[0], cur[1], cur[2], cur[3]);
  reinterpret_cast<uint4 *>(o)[1] = make_uint4(cur[4], cur[5], cur[6], cur[7]);
}

static unsigned leaf_chunks(int field, size_t n_rows) {
  size_t total = 32 + field_bytes(field) * n_rows;
  return (unsigned)((total + b3::CHUNK_LEN - 1) / b3::CHUNK_LEN);
}

size_t hash_scratch_bytes(int field, size_t n_rows, size_t n_cols) {
  unsigned k = leaf_chunks(field, n_rows);
  return k > 1 ? (size_t)k * n_cols * 32 : 0;
}

// rows of comm that chunk k of a leaf input reads: [0, leaf_chunk_rows_end(field, n_rows, k)); a chunk can be
// hashed as soon as that many rows are encoded (commit() from host memory overlaps hashing with the copy)
size_t leaf_chunk_rows_end(int field, size_t n_rows, unsigned k) {
  const size_t B = field_bytes(field), total = 32 + B * n_rows;
  size_t end = std::min<size_t>((size_t)(k + 1) * b3::CHUNK_LEN, total);  // exclusive byte end of the chunk
  if (end <= 32) return 0;
  return std::min(n_rows, (end - 32 + B - 1) / B);
}
unsigned leaf_chunk_count(int field, size_t n_rows) { return leaf_chunks(field, n_rows); }

// chunk chaining values (or, for single-chunk leaves, the digests) of chunks [k_first, k_first + k_count) of the
// n_cols columns that start at `comm`; they are columns [col0, col0 + n_cols) of a commitment with total_cols
// columns (the slots of `leaves` / `scratch` they fill).  A commit can therefore hash column ranges from
// different matrices -- Brakedown's systematic columns straight from the coefficient rows, before the code's
// sparse products have produced the rest.
cudaError_t launch_leaf_chunks_range(int field, const uint32_t *comm, size_t n_rows, size_t n_cols, size_t row_stride,
                                     uint8_t *leaves, void *scratch, unsigned k_first, unsigned k_count, size_t total_cols,
                                     size_t col0, cudaStream_t stream, size_t src_total, size_t col_stride) {
  if (n_cols == 0 || k_count == 0) return cudaSuccess;
  const unsigned n_chunks = leaf_chunks(field, n_rows);
  if (n_chunks > 65535u || k_first + k_count > n_chunks || col0 + n_cols > total_cols) return cudaErrorInvalidValue;
  uint32_t *out = n_chunks > 1 ? (uint32_t *)scratch : (uint32_t *)leaves;
  if (n_chunks > 1 && !scratch) return cudaErrorInvalidValue;
  dim3 grid((unsigned)((n_cols + HASH_THREADS - 1) / HASH_THREADS), k_count);
  // A/B knob: unused dynamic shared memory caps the resident hash CTAs per SM, so that a column range hashed on a side
  // stream leaves registers for the kernels it runs beside (only applied to ranges, i.e. col0 or total_cols set)
  const size_t pad = (total_cols != n_cols) ? (size_t)std::min<long>(47, std::max<long>(0, tunable("LEAF_SMEM_PAD_KB", 0))) << 10 : 0;
  switch (field) {
    case FT63:
      if (src_total == ~(size_t)0) leaf_chunk_kernel<FT63, false><<<grid, HASH_THREADS, pad, stream>>>(comm, n_rows, n_cols, row_stride, out, n_chunks, k_first, total_cols, col0, src_total, col_stride);
      else leaf_chunk_kernel<FT63, true><<<grid, HASH_THREADS, pad, stream>>>(comm, n_rows, n_cols, row_stride, out, n_chunks, k_first, total_cols, col0, src_total, col_stride);
      break;
    case FT127:
      if (src_total == ~(size_t)0) leaf_chunk_kernel<FT127, false><<<grid, HASH_THREADS, pad, stream>>>(comm, n_rows, n_cols, row_stride, out, n_chunks, k_first, total_cols, col0, src_total, col_stride);
      else leaf_chunk_kernel<FT127, true><<<grid, HASH_THREADS, pad, stream>>>(comm, n_rows, n_cols, row_stride, out, n_chunks, k_first, total_cols, col0, src_total, col_stride);
      break;
    case FT191: leaf_chunk_kernel_generic<FT191><<<grid, HASH_THREADS, pad, stream>>>(comm, n_rows, n_cols, row_stride, out, n_chunks, k_first, total_cols, col0, src_total, col_stride); break;
    case FT255:
      if (src_total == ~(size_t)0) leaf_chunk_kernel<FT255, false><<<grid, HASH_THREADS, pad, stream>>>(comm, n_rows, n_cols, row_stride, out, n_chunks, k_first, total_cols, col0, src_total, col_stride);
      else leaf_chunk_kernel<FT255, true><<<grid, HASH_THREADS, pad, stream>>>(comm, n_rows, n_cols, row_stride, out, n_chunks, k_first, total_cols, col0, src_total, col_stride);
      break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t launch_leaf_chunks(int field, const uint32_t *comm, size_t n_rows, size_t n_cols, size_t row_stride,
                               uint8_t *leaves, void *scratch, unsigned k_first, unsigned k_count, cudaStream_t stream) {
  return launch_leaf_chunks_range(field, comm, n_rows, n_cols, row_stride, leaves, scratch, k_first, k_count, n_cols, 0, stream, ~(size_t)0, 1);
}

// BLAKE3 tree over the chunk chaining values of every column (no-op for single-chunk leaves)
cudaError_t launch_leaf_merge(int field, size_t n_rows, size_t n_cols, uint8_t *leaves, void *scratch, cudaStream_t stream,
                              int *n_launches) {
  if (n_launches) *n_launches = 0;
  const unsigned n_chunks = leaf_chunks(field, n_rows);
  if (n_chunks <= 1 || n_cols == 0) return cudaSuccess;
  const unsigned gx = (unsigned)((n_cols + HASH_THREADS - 1) / HASH_THREADS);
  leaf_merge_kernel<<<gx, HASH_THREADS, 0, stream>>>((const uint32_t *)scratch, n_cols, n_chunks, (uint32_t *)leaves);
  if (n_launches) *n_launches = 1;
  return cudaGetLastError();
}

cudaError_t launch_hash_columns(int field, const uint32_t *comm, size_t n_rows, size_t n_cols, size_t row_stride,
                                uint8_t *leaves, void *scratch, cudaStream_t stream, int *n_launches) {
  if (n_launches) *n_launches = 0;
  if (n_cols == 0) return cudaSuccess;
  const unsigned n_chunks = leaf_chunks(field, n_rows);
  cudaError_t e = launch_leaf_chunks(field, comm, n_rows, n_cols, row_stride, leaves, scratch, 0, n_chunks, stream);
  if (e != cudaSuccess) return e;
  if (n_launches) ++*n_launches;
  int nm = 0;
  e = launch_leaf_merge(field, n_rows, n_cols, leaves, scratch, stream, &nm);
  if (n_launches) *n_launches += nm;
  return e;
}

// ---- Merkle layers: node = BLAKE3(left || right), a single 64-byte block (lib.rs:768-775) ----
__global__ void __launch_bounds__(HASH_THREADS)
merkle_layer_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, size_t n_out) {
  const size_t i = (size_t)blockIdx.x * HASH_THREADS + threadIdx.x;
  if (i >= n_out) return;
  const uint4 *p = reinterpret_cast<const uint4 *>(in + i * 16);
  uint32_t m[16];
#pragma unroll
  for (int q = 0; q < 4; q++) {
    uint4 t = p[q];
    m[4 * q] = t.x, m[4 * q + 1] = t.y, m[4 * q + 2] = t.z, m[4 * q + 3] = t.w;
  }
  uint32_t cv[8];
  b3::set_iv(cv);
  b3::compress(cv, m, 0, 64, b3::CHUNK_START | b3::CHUNK_END | b3::ROOT);
  uint32_t *o = out + i * 8;
  reinterpret_cast<uint4 *>(o)[0] = make_uint4(cv[0], cv[1], cv[2], cv[3]);
  reinterpret_cast<uint4 *>(o)[1] = make_uint4(cv[4], cv[5], cv[6], cv[7]);
}

// Many layers in one launch: a CTA reduces a subtree of 2^LOG_SUB nodes in shared memory and writes
// every intermediate layer to its place in the flat `hashes` array.
constexpr int MERKLE_LOG_SUB = 8;  // 256 input nodes (8 KiB) per CTA -> 8 layers per launch
__global__ void __launch_bounds__(128)
merkle_subtree_kernel(uint32_t *hashes, size_t layer_off, size_t layer_len, unsigned n_layers) {
  __shared__ uint32_t buf[(1 << MERKLE_LOG_SUB) * 8];
  const unsigned sub = 1u << n_layers;  // input nodes per CTA
  const size_t first = (size_t)blockIdx.x * sub;
  for (unsigned i = threadIdx.x; i < sub * 2; i += blockDim.x)
    reinterpret_cast<uint4 *>(buf)[i] = reinterpret_cast<const uint4 *>(hashes + (layer_off + first) * 8)[i];
  __syncthreads();
  size_t in_off = layer_off, in_len = layer_len;
  unsigned width = sub;
  for (unsigned l = 0; l < n_layers; l++) {
    const size_t out_off = in_off + in_len, out_len = in_len >> 1;
    const unsigned n_out = width >> 1;
    uint32_t cv[8];
    const bool act = threadIdx.x < n_out;
    if (act) {
      uint32_t m[16];
#pragma unroll
      for (int q = 0; q < 16; q++) m[q] = buf[threadIdx.x * 16 + q];
      b3::set_iv(cv);
      b3::compress(cv, m, 0, 64, b3::CHUNK_START | b3::CHUNK_END | b3::ROOT);
    }
    __syncthreads();
    if (act) {
#pragma unroll
      for (int q = 0; q < 8; q++) buf[threadIdx.x * 8 + q] = cv[q];
      uint32_t *o = hashes + (out_off + ((size_t)blockIdx.x * n_out) + threadIdx.x) * 8;
      reinterpret_cast<uint4 *>(o)[0] = make_uint4(cv[0], cv[1], cv[2], cv[3]);
      reinterpret_cast<uint4 *>(o)[1] = make_uint4(cv[4], cv[5], cv[6], cv[7]);
    }
    __syncthreads();
    in_off = out_off, in_len = out_len, width = n_out;
  }
}

// n_layers layers above `n_leaves` nodes laid out [nodes | layer 1 | ... | layer n_layers]; n_leaves must be
// a multiple of 2^n_layers (a forest of equal subtrees side by side is fine)
cudaError_t launch_merkle_layers(uint8_t *hashes, size_t n_leaves, unsigned n_layers, cudaStream_t stream,
                                 int *n_launches) {
  if (n_launches) *n_launches = 0;
  if (n_layers && (n_leaves & (((size_t)1 << n_layers) - 1))) return cudaErrorInvalidValue;
  uint32_t *h = (uint32_t *)hashes;
  size_t off = 0, len = n_leaves;
  unsigned left = n_layers;
  while (left > 0) {
    // MERKLE_WIDE_MIN: layers of at least this many nodes run one thread per output node, one launch per layer;
    // below it a CTA reduces 256 nodes through 8 layers in shared memory (0 = subtrees only).  Measured
    // (profiles/r02_ab_merkle_wide.jsonl): subtrees only 0.0533 against 0.0503 ms at 2^19 leaves, 0.138 against
    // 0.108 ms at 2^21; wide from 300 000 nodes up 0.0471 ms -- within noise of the default, which stays.
    const size_t wide_min = (size_t)std::max<long>(0, tunable("MERKLE_WIDE_MIN", 2 * HASH_THREADS * 148));
    if (wide_min && len >= wide_min) {
      // wide layer: one thread per node keeps every SM busy
      size_t n_out = len >> 1;
      merkle_layer_kernel<<<(unsigned)((n_out + HASH_THREADS - 1) / HASH_THREADS), HASH_THREADS, 0, stream>>>(
          h + off * 8, h + (off + len) * 8, n_out);
      off += len, len = n_out, left -= 1;
    } else {
      unsigned nl = left < (unsigned)MERKLE_LOG_SUB ? left : (unsigned)MERKLE_LOG_SUB;
      merkle_subtree_kernel<<<(unsigned)(len >> nl), 128, 0, stream>>>(h, off, len, nl);
      for (unsigned l = 0; l < nl; l++) off += len, len >>= 1;
      left -= nl;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (n_launches) ++*n_launches;
  }
  return cudaSuccess;
}

cudaError_t launch_merkle_tree(uint8_t *hashes, size_t np2, cudaStream_t stream, int *n_launches) {
  unsigned log_len = 0;
  while (((size_t)1 << log_len) < np2) log_len++;
  return launch_merkle_layers(hashes, np2, log_len, stream, n_launches);
}

}  // namespace lcpc
