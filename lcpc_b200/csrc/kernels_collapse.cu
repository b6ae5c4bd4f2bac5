// lcpc_b200/csrc/kernels_collapse.cu -- the prover's random linear combination of rows and the
// column gather for openings.
//
// collapse_columns (reference: lcpc-2d/src/lib.rs:1095-1123, call sites :1034,:1055):
//     poly[c] = sum_r tensor[r] * coeffs[r * n_per_row + c]
// The reference reduces every product mod p and adds; here each thread keeps the UNREDUCED
// double-width sum of its products (a 2N-limb accumulator kept below 2^(64N-1) by subtracting p*R when it
// grows past that, field.cuh mac_wide), and one Montgomery reduction is done per output element -- half the
// multiplier work of reduce-per-product.  The result is the same canonical residue: REDC is linear mod p
// and the final value is brought into [0, p).
// open_column's strided gather (lcpc-2d/src/lib.rs:802-808) is gather_columns_kernel; the challenge tensors
// of the degree tests (:1026-1032) are expanded on the device by expand_tensor_kernel.
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
collapse_kernel(const uint32_t *__restrict__ coeffs, size_t row_stride, size_t col_stride, const uint32_t *__restrict__ tensor,
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
      ld_elem<N>(a.v, coeffs + (r * row_stride + col * col_stride) * N);
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
                            int *n_launches, size_t col_stride) {
  if (n_launches) *n_launches = 0;
  if (n_per_row == 0) return cudaSuccess;
  unsigned grid = (unsigned)((n_per_row + COL_TILE - 1) / COL_TILE);
  const int threads = COL_TILE * ROW_GROUPS;
  switch (field) {
    case FT63: collapse_kernel<FT63><<<grid, threads, 0, stream>>>(coeffs, row_stride, col_stride, tensor, poly, n_rows, n_per_row); break;
    case FT127: collapse_kernel<FT127><<<grid, threads, 0, stream>>>(coeffs, row_stride, col_stride, tensor, poly, n_rows, n_per_row); break;
    case FT191: collapse_kernel<FT191><<<grid, threads, 0, stream>>>(coeffs, row_stride, col_stride, tensor, poly, n_rows, n_per_row); break;
    case FT255: collapse_kernel<FT255><<<grid, threads, 0, stream>>>(coeffs, row_stride, col_stride, tensor, poly, n_rows, n_per_row); break;
    default: return cudaErrorInvalidValue;
  }
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && n_launches) *n_launches = 1;
  return e;
}

// ---- challenge tensor expansion (lcpc-2d/src/lib.rs:1026-1032: ChaCha20Rng::from_seed(key), then n_rows x
// F::random; the reference notes "could expand seed in parallel instead of in series") ----------------------
// F::random (ff_derive) draws L u64 words, masks the top limb to NUM_BITS and keeps the candidate iff it is < p,
// taking the accepted limbs as the Montgomery image.  A u64 is two consecutive u32 words of the ChaCha20
// stream, so candidate c is exactly words [2L c, 2L (c+1)) of the stream: every thread builds one candidate from
// the one or two ChaCha20 blocks holding its words, a block-wide scan ranks the accepted ones, and rounds of
// blockDim candidates repeat until n are out.  One CTA: n is the row count of the commitment (hundreds).
__device__ __forceinline__ uint32_t rotl32(uint32_t x, int n) { return __funnelshift_l(x, x, n); }
#define LCPC_CHACHA_QR(a, b, c, d)   \
  do {                              \
    a += b, d ^= a, d = rotl32(d, 16); \
    c += d, b ^= c, b = rotl32(b, 12); \
    a += b, d ^= a, d = rotl32(d, 8);  \
    c += d, b ^= c, b = rotl32(b, 7);  \
  } while (0)
__device__ void chacha20_block(const uint32_t (&key)[8], uint64_t counter, uint64_t stream, uint32_t (&out)[16]) {
  uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key[0], key[1], key[2], key[3],
                    key[4], key[5], key[6], key[7], (uint32_t)counter, (uint32_t)(counter >> 32),
                    (uint32_t)stream, (uint32_t)(stream >> 32)};
  uint32_t x[16];
#pragma unroll
  for (int i = 0; i < 16; i++) x[i] = s[i];
  for (int r = 0; r < 10; r++) {
    LCPC_CHACHA_QR(x[0], x[4], x[8], x[12]);
    LCPC_CHACHA_QR(x[1], x[5], x[9], x[13]);
    LCPC_CHACHA_QR(x[2], x[6], x[10], x[14]);
    LCPC_CHACHA_QR(x[3], x[7], x[11], x[15]);
    LCPC_CHACHA_QR(x[0], x[5], x[10], x[15]);
    LCPC_CHACHA_QR(x[1], x[6], x[11], x[12]);
    LCPC_CHACHA_QR(x[2], x[7], x[8], x[13]);
    LCPC_CHACHA_QR(x[3], x[4], x[9], x[14]);
  }
#pragma unroll
  for (int i = 0; i < 16; i++) out[i] = x[i] + s[i];
}

constexpr int EXPAND_THREADS = 256;
template <int FID>
__global__ void __launch_bounds__(EXPAND_THREADS)
expand_tensor_kernel(const uint32_t *__restrict__ key_in, uint64_t stream, size_t n, uint32_t num_bits,
                     uint32_t *__restrict__ out) {
  using F = Field<FID>;
  constexpr int N = F::N;
  __shared__ uint32_t warp_sums[EXPAND_THREADS / 32];
  __shared__ uint32_t round_total;
  uint32_t key[8];
#pragma unroll
  for (int i = 0; i < 8; i++) key[i] = key_in[i];
  const uint32_t top_mask = (num_bits % 32) ? ((1u << (num_bits % 32)) - 1u) : 0xffffffffu;
  size_t done = 0;
  for (uint64_t base = 0; done < n; base += EXPAND_THREADS) {
    const uint64_t c = base + threadIdx.x;          // candidate index
    const uint64_t w0 = c * N;                      // its first stream word
    uint32_t cand[N];
    {
      uint32_t blk[16];
      uint64_t b = w0 / 16;
      chacha20_block(key, b, stream, blk);
#pragma unroll
      for (int l = 0; l < N; l++) {
        const uint64_t w = w0 + l;
        if (w / 16 != b) {  // the candidate runs into the next block (Ft191: 6 words per candidate)
          b = w / 16;
          chacha20_block(key, b, stream, blk);
        }
        uint32_t v = 0;
#pragma unroll
        for (int q = 0; q < 16; q++)
          if (q == (int)(w % 16)) v = blk[q];
        cand[l] = v;
      }
    }
    cand[N - 1] &= top_mask;
    // accept iff cand < p
    bool lt = false, decided = false;
#pragma unroll
    for (int l = N - 1; l >= 0; l--) {
      if (!decided && cand[l] != FieldP<FID>::P(l)) lt = cand[l] < FieldP<FID>::P(l), decided = true;
    }
    const uint32_t ok = lt ? 1u : 0u;
    // rank among the accepted candidates of this round (warp ballot + warp totals)
    const unsigned lane = threadIdx.x % 32, wid = threadIdx.x / 32;
    const uint32_t ballot = __ballot_sync(0xffffffffu, ok);
    const uint32_t before = __popc(ballot & ((1u << lane) - 1u));
    if (lane == 0) warp_sums[wid] = __popc(ballot);
    __syncthreads();
    uint32_t off = 0, tot = 0;
#pragma unroll
    for (int q = 0; q < EXPAND_THREADS / 32; q++) {
      if (q < (int)wid) off += warp_sums[q];
      tot += warp_sums[q];
    }
    const size_t pos = done + off + before;
    if (ok && pos < n) {
#pragma unroll
      for (int l = 0; l < N; l++) out[pos * N + l] = cand[l];
    }
    if (threadIdx.x == 0) round_total = tot;
    __syncthreads();
    done += round_total;
    __syncthreads();
  }
}

cudaError_t launch_expand_tensor(int field, const uint32_t *d_key, uint64_t stream_id, size_t n, uint32_t *d_out,
                                 cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  switch (field) {  // NUM_BITS of the fields (lcpc-test-fields/src/lib.rs:13-59)
    case FT63: expand_tensor_kernel<FT63><<<1, EXPAND_THREADS, 0, stream>>>(d_key, stream_id, n, 63, d_out); break;
    case FT127: expand_tensor_kernel<FT127><<<1, EXPAND_THREADS, 0, stream>>>(d_key, stream_id, n, 127, d_out); break;
    case FT191: expand_tensor_kernel<FT191><<<1, EXPAND_THREADS, 0, stream>>>(d_key, stream_id, n, 191, d_out); break;
    case FT255: expand_tensor_kernel<FT255><<<1, EXPAND_THREADS, 0, stream>>>(d_key, stream_id, n, 255, d_out); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

// out[i * n_rows + r] = comm[r * row_stride + cols[i]]
__global__ void gather_columns_kernel(const uint32_t *__restrict__ comm, size_t n_rows, size_t row_stride,
                                      const uint64_t *__restrict__ cols, size_t n_open, uint32_t *__restrict__ out,
                                      int n_limbs, size_t col_stride) {
  const size_t total = n_open * n_rows * (size_t)n_limbs;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t l = idx % n_limbs, e = idx / n_limbs;
    const size_t r = e % n_rows, i = e / n_rows;
    out[idx] = comm[(r * row_stride + cols[i] * col_stride) * n_limbs + l];
  }
}

cudaError_t launch_gather_columns(int field, const uint32_t *comm, size_t n_rows, size_t row_stride,
                                  const uint64_t *cols, size_t n_open, uint32_t *out, cudaStream_t stream,
                                  size_t col_stride) {
  int nl = field_limbs32(field);
  if (nl < 0) return cudaErrorInvalidValue;
  size_t total = n_open * n_rows * (size_t)nl;
  if (total == 0) return cudaSuccess;
  unsigned grid = (unsigned)std::min<size_t>((total + 255) / 256, 148 * 16);
  gather_columns_kernel<<<grid, 256, 0, stream>>>(comm, n_rows, row_stride, cols, n_open, out, nl, col_stride);
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
