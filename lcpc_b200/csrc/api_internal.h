// lcpc_b200/csrc/api_internal.h -- objects behind the opaque handles of include/lcpc_b200.h, shared by the
// translation units that implement the C ABI (api.cu: single-GPU commit / prove / verify; shard.cu: the commit
// sharded over several GPUs).  Internal: nothing here is part of the boundary.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/lcpc_b200.h"
#include "../../include/lcpc_b200_host.h"
#include "expander.h"
#include "host_transcript.h"
#include "kernels.h"

// ------------------------------------------------------------------------------------------------
// Lifetimes: an encoding keeps its context alive and a commit keeps its encoding alive (reference counts), so the
// three `*_destroy` / `*_free` calls may come in any order -- a garbage-collected host (the Python layer here, a
// Drop order a Rust shim does not control) cannot turn the order of finalisers into a use-after-free.
struct lcpc_b200_ctx {
  std::atomic<int> refs{1};
  int device = 0;
  cudaStream_t stream = nullptr;
  // host->device staging stream + events: commit() from host memory copies the coefficient rows in
  // row-chunks on this stream while the engine stream encodes the chunks that have already landed
  cudaStream_t copy_stream = nullptr;
  static constexpr int MAX_CHUNKS = 16;
  cudaEvent_t chunk_ev[MAX_CHUNKS] = {};
  cudaEvent_t begin_ev = nullptr;
  // side stream: column hashing of already-encoded row chunks runs here next to the encode of later rows
  // (the transforms saturate the multiplier pipe, BLAKE3 the ALU pipe: they overlap on the same SMs)
  cudaStream_t side_stream = nullptr;
  cudaEvent_t side_ev[MAX_CHUNKS] = {};
  cudaEvent_t side_done = nullptr;
  cudaEvent_t lane_fork = nullptr, lane_join = nullptr;  // lent to the expander encode in scatter mode
  std::mutex mu;
  std::string err;
  uint64_t launches = 0;
  // grow-only device scratch shared by the stateless entry points
  void *scratch = nullptr;
  size_t scratch_bytes = 0;
  // Brakedown codes generated on this device from (field, code spec, n_per_row, seed): kept for the context's
  // lifetime, so building the same encoding again (another commit length with the same row length, a verifier next
  // to a prover) costs nothing (matgen_bench of the reference, lcpc-brakedown-pc/src/bench.rs:22-30, is this setup)
  struct SeededCode {
    int field, code;
    size_t n_per_row;
    uint64_t seed;
    lcpc::ExpanderCode *ptr;
  };
  std::vector<SeededCode> code_cache;
  // grow-only page-locked host staging (results the host has to read right away: canonical bytes for the transcript)
  void *h_stage = nullptr;
  size_t h_stage_bytes = 0;
};

inline int fail(lcpc_b200_ctx *ctx, int code, const char *fmt, ...) {
  if (ctx) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    ctx->err = buf;
  }
  return code;
}

inline int cuda_fail(lcpc_b200_ctx *ctx, cudaError_t e, const char *what) {
  int code = (e == cudaErrorMemoryAllocation) ? LCPC_B200_ERR_OOM : LCPC_B200_ERR_CUDA;
  return fail(ctx, code, "%s: %s", what, cudaGetErrorString(e));
}

#define CU(ctx, call)                                        \
  do {                                                       \
    cudaError_t e_ = (call);                                 \
    if (e_ != cudaSuccess) return cuda_fail(ctx, e_, #call); \
  } while (0)

inline int bind_device(lcpc_b200_ctx *ctx) {
  cudaError_t e = cudaSetDevice(ctx->device);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaSetDevice");
  return LCPC_B200_OK;
}

inline int ensure_scratch(lcpc_b200_ctx *ctx, size_t bytes) {
  if (bytes <= ctx->scratch_bytes) return LCPC_B200_OK;
  if (ctx->scratch) cudaFree(ctx->scratch);
  ctx->scratch = nullptr, ctx->scratch_bytes = 0;
  CU(ctx, cudaMalloc(&ctx->scratch, bytes));
  ctx->scratch_bytes = bytes;
  return LCPC_B200_OK;
}

// page-locked staging of at least `bytes`; nullptr if the allocation fails (callers fall back to pageable memory)
inline void *host_stage(lcpc_b200_ctx *ctx, size_t bytes) {
  if (bytes <= ctx->h_stage_bytes) return ctx->h_stage;
  if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
  ctx->h_stage = nullptr, ctx->h_stage_bytes = 0;
  if (cudaHostAlloc(&ctx->h_stage, bytes, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    ctx->h_stage = nullptr;
    return nullptr;
  }
  ctx->h_stage_bytes = bytes;
  return ctx->h_stage;
}

inline bool is_pow2(size_t v) { return v && !(v & (v - 1)); }
inline unsigned log2_ceil(size_t v) {  // lcpc-2d/src/lib.rs:827-829
  unsigned l = 0;
  while (((size_t)1 << l) < v) l++;
  return l;
}


// ------------------------------------------------------------------------------------------------
struct lcpc_b200_enc {
  std::atomic<int> refs{1};
  lcpc_b200_ctx *ctx = nullptr;
  int kind = 0, field = 0;
  size_t n_per_row = 0, n_cols = 0;
  // ligero
  unsigned log_n = 0;
  uint32_t *d_roots = nullptr;
  // sdig
  lcpc::ExpanderCode *code = nullptr;
};


// helpers of api.cu that shard.cu builds on (all expect the context's mutex held and its device bound); defined
// inside api.cu's extern "C" block, hence the linkage
struct Labels {
  const uint8_t *dt, *pr, *pe, *co;
  size_t dt_len, pr_len, pe_len, co_len;
};
struct HashTrail;
struct HostOut;
extern "C" {
void ctx_unref(lcpc_b200_ctx *ctx);
void enc_unref(lcpc_b200_enc *enc);
// encode n_rows rows: src (stride/valid) -> dst (stride n_cols), or per column block with a scatter; enqueues only
int encode_rows(lcpc_b200_enc *enc, const uint32_t *src, size_t src_stride, size_t valid, uint32_t *dst, size_t n_rows,
                void *enc_scratch, const lcpc::Scatter *scatter = nullptr);
size_t enc_scratch_bytes(const lcpc_b200_enc *enc, size_t n_rows);
// the same fed from host memory in row-chunks on the copy stream (see api.cu)
int encode_rows_from_host(lcpc_b200_enc *enc, const void *src, size_t len, uint32_t *d_coeffs, uint32_t *d_comm,
                          size_t n_rows, void *enc_scratch, cudaEvent_t first_ev, const lcpc::Scatter *scatter = nullptr,
                          HashTrail *trail = nullptr, const HostOut *host_out = nullptr,
                          cudaEvent_t coeffs_free_ev = nullptr);

// domain-separation labels of prove()/verify() with the reference's literal defaults (lcpc-2d/src/macros.rs:28-36)
Labels resolve_labels(const lcpc_b200_labels *in);
// ChaCha20Rng::from_seed(key) -> n x Uniform::new(0usize, n_cols) (lcpc-2d/src/lib.rs:1073-1080, :903-911)
void sample_columns(const uint8_t key[32], size_t n_cols, size_t n, uint64_t *out);
}  // extern "C"
