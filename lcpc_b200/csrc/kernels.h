// lcpc_b200/csrc/kernels.h -- host-side launchers of the device kernels (internal; the public
// boundary is include/lcpc_b200.h).  All pointers are DEVICE pointers; elements are N 32-bit limbs
// in Montgomery form (see field.cuh).  Every launcher enqueues on `stream` and returns the CUDA
// status of the launch; nothing here synchronises.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace lcpc {

// named integer knobs for A/B measurements (tunables.cpp): set value > LCPC_B200_<NAME> > dflt
long tunable(const char *name, long dflt);
void set_tunable(const char *name, long value);
void clear_tunable(const char *name);

int field_limbs32(int field);  // 2/4/6/8, or -1
inline size_t field_bytes(int field) { return 4 * (size_t)field_limbs32(field); }

// element-wise field ops (test hook): op 0 add, 1 sub, 2 mul, 4 from_mont, 5 mul_full+redc, 6 lazy 37-term sums,
// 7 Karatsuba product + redc
cudaError_t launch_field_op(int field, int op, uint32_t *r, const uint32_t *a, const uint32_t *b, size_t n,
                            cudaStream_t stream);

// Destination of a row-block encode that feeds column-sharded hashing (the multi-GPU commit): column block
// h = columns [starts[h], starts[h+1]) of every encoded row goes to the matrix dst[h][total rows][width_h],
// which may live on this GPU or be a peer-mapped buffer of another GPU reached over NVLink; this call's
// rows start at row0 of those matrices.  n_blocks == 0 means a plain row-major destination.
constexpr int MAX_SCATTER = 16;
struct Scatter {
  unsigned n_blocks;
  unsigned long long row0;
  unsigned long long starts[MAX_SCATTER + 1];
  uint32_t *dst[MAX_SCATTER];
};

// ---- Ligero: radix-2 DIF NTT, in-order in / bit-reversed out (fffft `fft_io_pc`) ----
// roots: w^0 .. w^(n_cols/2-1), w = root_of_unity()^(2^(S-log2 n_cols))  (FFTPrecomp)
// src rows hold `src_valid` leading elements (the rest of the n_cols-long row is implicit zeros) and
// are `src_stride` elements apart; dst rows are `dst_stride` apart.  src == dst (in place) is allowed
// when src_stride == dst_stride.  Returns the number of kernels launched through *n_launches.
// copy_dst (optional): the first pass also stores every source element it reads to copy_dst, rows
// copy_stride apart (commit()'s private copy of the coefficients without a second read).
cudaError_t launch_ntt_rows(int field, const uint32_t *src, size_t src_stride, size_t src_valid, uint32_t *dst,
                            size_t dst_stride, const uint32_t *roots, unsigned log_n, size_t n_rows,
                            cudaStream_t stream, int *n_launches, const Scatter *scatter = nullptr,
                            uint32_t *copy_dst = nullptr, size_t copy_stride = 0);
// powers of w into roots[0..half): roots[i] = w^i (Montgomery); `w` points at 2N limbs: [w | R mod p]
cudaError_t launch_root_table(int field, uint32_t *roots, const uint32_t *w, size_t half, cudaStream_t stream);

// ---- column hashing + Merkle (lcpc-2d/src/lib.rs:706-785) ----
// leaves[c] = BLAKE3(0^32 || repr(comm[0][c]) || ... || repr(comm[n_rows-1][c])) for c < n_cols;
// comm element (r, c) lives at comm + (r * row_stride + c) * N.  scratch must hold
// hash_scratch_bytes(field, n_rows, n_cols) bytes (0 if the leaf input fits one BLAKE3 chunk).
size_t hash_scratch_bytes(int field, size_t n_rows, size_t n_cols);
cudaError_t launch_hash_columns(int field, const uint32_t *comm, size_t n_rows, size_t n_cols, size_t row_stride,
                                uint8_t *leaves, void *scratch, cudaStream_t stream, int *n_launches);
// the same in pieces, for hashing that trails the row encode: chunk k of every leaf input needs the rows
// [0, leaf_chunk_rows_end(k)) only
unsigned leaf_chunk_count(int field, size_t n_rows);
size_t leaf_chunk_rows_end(int field, size_t n_rows, unsigned k);
cudaError_t launch_leaf_chunks(int field, const uint32_t *comm, size_t n_rows, size_t n_cols, size_t row_stride,
                               uint8_t *leaves, void *scratch, unsigned k_first, unsigned k_count, cudaStream_t stream);
// the same for a column range: the n_cols columns at `comm` are columns [col0, col0 + n_cols) of a commitment with
// total_cols columns; element (r, c) of the source exists only if r * row_stride + c < src_total (beyond that: zero,
// the padding of a short last row)
cudaError_t launch_leaf_chunks_range(int field, const uint32_t *comm, size_t n_rows, size_t n_cols, size_t row_stride,
                                     uint8_t *leaves, void *scratch, unsigned k_first, unsigned k_count, size_t total_cols,
                                     size_t col0, cudaStream_t stream, size_t src_total = ~(size_t)0, size_t col_stride = 1);
// (col_stride: elements between consecutive columns -- 1 for a row-major matrix; a column-major matrix, like the
// expander's work buffer W[position][row], is row_stride = 1, col_stride = n_rows: every column is contiguous)
cudaError_t launch_leaf_merge(int field, size_t n_rows, size_t n_cols, uint8_t *leaves, void *scratch, cudaStream_t stream,
                              int *n_launches);
// hashes = [leaves(np2) | layer 1 | ... | root]; leaves given, upper layers computed
cudaError_t launch_merkle_tree(uint8_t *hashes, size_t np2, cudaStream_t stream, int *n_launches);

// n_layers Merkle layers over n_leaves nodes (a forest of equal aligned subtrees side by side)
cudaError_t launch_merkle_layers(uint8_t *hashes, size_t n_leaves, unsigned n_layers, cudaStream_t stream,
                                 int *n_launches);

// ---- multi-GPU transpose step: split a row-block [n_rows][n_cols] into per-destination column-block
// tiles, tile h = columns [starts[h], starts[h+1]) stored contiguously as [n_rows][width_h] at element
// offset n_rows * starts[h] of dst (the send buffer of the all-to-all) ----
cudaError_t launch_pack_column_blocks(int field, const uint32_t *src, size_t n_rows, size_t n_cols,
                                      unsigned n_blocks, const uint64_t *d_starts, uint32_t *dst,
                                      cudaStream_t stream);

// ---- collapse_columns (lcpc-2d/src/lib.rs:1095-1123): poly[c] = sum_r tensor[r] * coeffs[r][c] ----
size_t collapse_scratch_bytes(int field, size_t n_rows, size_t n_per_row);
cudaError_t launch_collapse(int field, const uint32_t *coeffs, size_t row_stride, const uint32_t *tensor,
                            uint32_t *poly, size_t n_rows, size_t n_per_row, void *scratch, cudaStream_t stream,
                            int *n_launches, size_t col_stride = 1);  // element (r, c) at coeffs[r * row_stride + c * col_stride]

// ---- challenge tensor (lcpc-2d/src/lib.rs:1026-1032): out[0..n) = n x F::random from ChaCha20Rng::from_seed(key)
// with stream id `stream_id` (0 for from_seed); d_key: the 32-byte key as 8 little-endian words in device memory
cudaError_t launch_expand_tensor(int field, const uint32_t *d_key, uint64_t stream_id, size_t n, uint32_t *d_out,
                                 cudaStream_t stream);

// ---- open_column gather (lcpc-2d/src/lib.rs:802-808): out[i][r] = comm[r][cols[i]] ----
cudaError_t launch_gather_columns(int field, const uint32_t *comm, size_t n_rows, size_t row_stride,
                                  const uint64_t *cols, size_t n_open, uint32_t *out, cudaStream_t stream,
                                  size_t col_stride = 1);

// Merkle paths (lcpc-2d/src/lib.rs:811-821): out[i][l] = sibling of column cols[i]'s ancestor on layer l
cudaError_t launch_gather_paths(const uint8_t *hashes, size_t np2, const uint64_t *cols, size_t n_open,
                                unsigned path_len, uint8_t *out, cudaStream_t stream);

// ---- verifier side (lcpc-2d/src/lib.rs:926-951), kernels_verify.cu ----
// opened columns in[j][r] (each LcColumn::col contiguous) -> row-major out[r][j]
cudaError_t launch_transpose_columns(int field, const uint32_t *in, uint32_t *out, size_t n_open, size_t n_rows,
                                     cudaStream_t stream);
// flags[j]: bit 0 degree-test values match (evals[k][j] == rows[k][cols[j]] for k < n_tensors - 1), bit 1 the
// evaluation value matches (k = n_tensors - 1), bit 2 leaves[j] + paths[j][0..path_len) hash up to root
cudaError_t launch_check_columns(int field, const uint32_t *evals, const uint32_t *rows, size_t row_stride,
                                 unsigned n_tensors, const uint64_t *cols, size_t n_open, const uint8_t *leaves,
                                 const uint8_t *paths, unsigned path_len, const uint8_t *root, uint32_t *flags,
                                 cudaStream_t stream);
// out[0] = sum_i a[i] * b[i]; partials: DOT_PARTIALS elements of scratch
constexpr unsigned DOT_PARTIALS = 592;
cudaError_t launch_dot(int field, const uint32_t *a, const uint32_t *b, size_t n, uint32_t *partials, uint32_t *out,
                       cudaStream_t stream, int *n_launches);

}  // namespace lcpc
