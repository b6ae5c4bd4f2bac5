// lcpc_b200/csrc/expander.h -- device side of the Brakedown ("SDIG") expander code (internal).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <string>

#include "kernels.h"

namespace lcpc {

// borrowed view of one CsMat::new_csc((m, n), ptrs, idxs, data) (matgen.rs:187), host memory
struct CscView {
  size_t m, n;
  const uint64_t *ptrs, *idxs, *data;
};

struct ExpanderCode;

// a second stream + two events the caller lends for work that may run beside the chain (scatter mode only)
struct SideLane {
  cudaStream_t stream;
  cudaEvent_t fork, join;
};

// uploads the code in gather (row-compressed) form; returns an LCPC_B200_* status
int expander_build(int field, size_t n_levels, const CscView *pre, const CscView *post, cudaStream_t stream,
                   ExpanderCode **out, std::string *err);
void expander_free(ExpanderCode *code);   // drops one reference
void expander_retain(ExpanderCode *code);
size_t expander_n_in(const ExpanderCode *code);
size_t expander_codeword_length(const ExpanderCode *code);  // encode.rs:18-33
size_t expander_nnz(const ExpanderCode *code);
size_t expander_scratch_bytes(const ExpanderCode *code, size_t n_rows);
// encode n_rows rows: src row r holds `valid` >= n_in leading elements at src + r*src_stride; dst rows
// are dst_stride apart and receive the full codeword.  src may equal dst.  copy_dst (optional): the first
// n_in elements of every source row are also stored there, rows copy_stride apart.  src_total: elements that
// exist at src (row-major, src_stride apart); positions beyond it read as zero (the padded last row).
cudaError_t expander_encode_rows(const ExpanderCode *code, const uint32_t *src, size_t src_stride, size_t valid,
                                 uint32_t *dst, size_t dst_stride, size_t n_rows, void *scratch, cudaStream_t stream,
                                 int *n_launches, const Scatter *scatter = nullptr, uint32_t *copy_dst = nullptr,
                                 size_t copy_stride = 0, size_t src_total = ~(size_t)0, const SideLane *side = nullptr);

// dst == nullptr in expander_encode_rows (no scatter): the codewords stay in the work buffer, column-major
// (`scratch` = W[n_cols][n_rows], element (row r, position j) at (j * n_rows + r)); this produces the row-major matrix
// from it on demand
cudaError_t expander_untranspose(const ExpanderCode *code, const void *scratch, uint32_t *dst, size_t dst_stride, size_t n_rows,
                                 cudaStream_t stream, size_t n_pos = 0);

}  // namespace lcpc
