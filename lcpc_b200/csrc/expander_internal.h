// lcpc_b200/csrc/expander_internal.h -- the device-side representation of a Brakedown code, shared by the encoder
// (kernels_expander.cu) and the on-device code generator (device_matgen.cu).  Internal.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "expander.h"

namespace lcpc {

constexpr size_t FUSED_SMEM_BYTES = 40 << 10;  // window + temporary of the fused innermost levels
constexpr int MAX_FUSED_OPS = 16;

struct DeviceCsr {
  size_t m = 0, n = 0, nnz = 0;
  uint32_t *rowptr = nullptr;  // m + 1
  uint32_t *colidx = nullptr;  // nnz, ascending within a row
  uint32_t *vals = nullptr;    // nnz * N limbs
  // column-chunked schedule (see spmm_kernel): seg[q * m + i] = first non-zero of row i whose column is >= q * n / seg_q,
  // q = 0 .. seg_q; built on first use for a given chunk count and kept
  mutable uint32_t *seg = nullptr;
  mutable unsigned seg_q = 0;
  // matrices generated on the device (device_matgen.cu) also keep the reference's own column-compressed form:
  // exactly csc_d sorted distinct row indices per input column (matgen.rs:144-161), entry k of column c at c * csc_d + k
  uint32_t *csc_idx = nullptr;   // n * csc_d row indices
  uint32_t *csc_data = nullptr;  // n * csc_d elements
  size_t csc_d = 0;
};

struct ExpanderOp {
  int kind;  // 0 = sparse product, 1 = reed-solomon
  int mat;   // index into mats (kind 0)
  size_t in_off, in_len, out_off, out_len;
  bool out_tmp, in_tmp;  // x_t lives in the temporary
};

struct ExpanderCode {
  int refs = 1;  // encodings sharing this code (the per-context cache of seeded codes holds one as well); under the ctx mutex
  int field = 0;
  size_t n_levels = 0, n_in = 0, n_cols = 0, nnz = 0, tmp_len = 0;
  std::vector<DeviceCsr> mats;
  std::vector<ExpanderOp> ops;
  // ops [fuse_lo, fuse_hi] (the innermost levels around the Reed-Solomon base) touch only the codeword window
  // [win_lo, win_lo + win_len) and run as ONE kernel with that window in shared memory; fuse_lo > fuse_hi: none
  size_t fuse_lo = 1, fuse_hi = 0, win_lo = 0, win_len = 0;
};

// ops, offsets and the fused-window plan from the matrices' dimensions (mats[0..t) precodes, mats[t..2t) postcodes
// must already carry m and n); returns an LCPC_B200_* status
int expander_assemble(ExpanderCode *c, size_t t, std::string *err);

// matgen::generate (lcpc-brakedown-pc/src/matgen.rs:28-52) on the device: every level's precode and postcode drawn
// from ChaCha20Rng::seed_from_u64(seed) with stream id = level, straight into gather (row-compressed) form.
// dims: t (n, m, d) triples for the precodes, then t for the postcodes (matgen::get_dims, host arithmetic).
struct MatgenDims { size_t n, m, d; };
int device_matgen(int field, uint64_t seed, size_t t, const MatgenDims *pre, const MatgenDims *post, cudaStream_t stream,
                  ExpanderCode **out, std::string *err);

}  // namespace lcpc
