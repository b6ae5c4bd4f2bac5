// lcpc_b200/csrc/api.cu -- the extern "C" boundary declared in include/lcpc_b200.h.
// Owns contexts, encodings and device-resident commits; all compute is in the kernels_*.cu files.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "api_internal.h"
#include "expander_internal.h"
#include "field.cuh"
#include "host_chacha.h"
#include "host_matgen.h"

using namespace lcpc;

// ------------------------------------------------------------------------------------------------
// host-side setup scalars (one = R mod p, w = the n_cols-th root of unity).  These few values are
// computed once per encoding with field.cuh's host build of the same algorithms.
struct FieldHostInfo { unsigned two_adicity; uint32_t generator; };
static bool field_host_info(int field, FieldHostInfo *out) {
  switch (field) {  // lcpc-test-fields/src/lib.rs:20,32,44,56 (PrimeFieldGenerator) and 2-adicity of p-1
    case FT63: *out = {41, 10}; return true;
    case FT127: *out = {40, 3}; return true;
    case FT191: *out = {41, 5}; return true;
    case FT255: *out = {41, 5}; return true;
  }
  return false;
}

template <int FID>
static void host_root_of_unity(unsigned log_len, unsigned two_adicity, uint32_t gen, uint32_t *w_out, uint32_t *one_out) {
  using F = Field<FID>;
  constexpr int N = F::N;
  typename F::Elem one = F::zero();
  one.v[0] = 1;
  for (int i = 0; i < 32 * N; i++) one = F::add(one, one);  // R mod p
  typename F::Elem g = one;
  for (uint32_t i = 1; i < gen; i++) g = F::add(g, one);  // generator in Montgomery form
  // t = (p - 1) >> two_adicity
  uint32_t t[N];
  for (int i = 0; i < N; i++) t[i] = FieldP<FID>::P(i);
  t[0] -= 1;
  for (unsigned s = 0; s < two_adicity; s++) {
    for (int i = 0; i < N; i++) t[i] = (t[i] >> 1) | (i + 1 < N ? t[i + 1] << 31 : 0);
  }
  typename F::Elem acc = one, base = g;
  for (int i = 0; i < N; i++)
    for (int k = 0; k < 32; k++) {
      if ((t[i] >> k) & 1) acc = F::mul(acc, base);
      base = F::mul(base, base);
    }
  // PrimeField::root_of_unity() = acc; fffft squares it down to order 2^log_len
  for (unsigned i = log_len; i < two_adicity; i++) acc = F::mul(acc, acc);
  memcpy(w_out, acc.v, sizeof acc.v);
  memcpy(one_out, one.v, sizeof one.v);
}

static void host_root(int field, unsigned log_len, uint32_t *w, uint32_t *one) {
  FieldHostInfo fi;
  field_host_info(field, &fi);
  switch (field) {
    case FT63: host_root_of_unity<FT63>(log_len, fi.two_adicity, fi.generator, w, one); break;
    case FT127: host_root_of_unity<FT127>(log_len, fi.two_adicity, fi.generator, w, one); break;
    case FT191: host_root_of_unity<FT191>(log_len, fi.two_adicity, fi.generator, w, one); break;
    default: host_root_of_unity<FT255>(log_len, fi.two_adicity, fi.generator, w, one); break;
  }
}

// ------------------------------------------------------------------------------------------------
struct lcpc_b200_commit {
  lcpc_b200_enc *enc = nullptr;
  size_t n_rows = 0, n_per_row = 0, n_cols = 0, np2 = 0;
  uint32_t *d_coeffs = nullptr, *d_comm = nullptr;
  uint8_t *d_hashes = nullptr;
  void *d_hash_scratch = nullptr;
  void *d_enc_scratch = nullptr;
  // Brakedown, device-resident route: the codewords are left in the encoder's work buffer (d_enc_scratch,
  // column-major W[position][row]) -- column hashing and column openings read columns, which are contiguous there --
  // and the row-major d_comm is only made when somebody asks for it (ensure_comm)
  bool comm_in_w = false;    // Brakedown, device route: the codewords are in d_enc_scratch (column-major), d_comm is stale
  bool coeffs_in_w = false;  // ... and so are the padded coefficient rows (the systematic positions), d_coeffs is stale
  // prove-side staging
  uint32_t *d_tensor = nullptr, *d_poly = nullptr, *d_key = nullptr, *d_repr = nullptr;
  void *h_poly = nullptr;  // page-locked landing buffer for collapse results (pageable destinations copy from it)
  // phase boundaries of the last run: start | copy+pad | encode | leaf hash | merkle
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  int encode_launches = 0, hash_launches = 0, merkle_launches = 0;
};

extern "C" {

static int return_poly(lcpc_b200_commit *c, uint64_t *poly);

const char *lcpc_b200_version(void) { return "lcpc_b200 0.1 (sm_100a)"; }

int lcpc_b200_set_tunable(const char *name, long value) {
  if (!name || !*name) return LCPC_B200_ERR_BAD_ARG;
  set_tunable(name, value);
  return LCPC_B200_OK;
}
long lcpc_b200_get_tunable(const char *name, long dflt) { return name ? tunable(name, dflt) : dflt; }

int lcpc_b200_field_limbs(int field) {
  int n = field_limbs32(field);
  return n < 0 ? -1 : n / 2;
}

int lcpc_b200_field_one(int field, uint64_t *out) {
  FieldHostInfo fi;
  if (!out || !field_host_info(field, &fi)) return LCPC_B200_ERR_BAD_ARG;
  uint32_t w[8], one[8];
  host_root(field, 1, w, one);
  memcpy(out, one, field_bytes(field));
  return LCPC_B200_OK;
}

int lcpc_b200_ctx_create(int device, lcpc_b200_ctx **out) {
  if (!out) return LCPC_B200_ERR_BAD_ARG;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0 || device < 0 || device >= count) return LCPC_B200_ERR_CUDA;
  lcpc_b200_ctx *ctx = new (std::nothrow) lcpc_b200_ctx;
  if (!ctx) return LCPC_B200_ERR_OOM;
  ctx->device = device;
  bool ok = cudaSetDevice(device) == cudaSuccess;
  if (ok) {
    // A/B knob: L2 set-aside for persisting (evict_last) lines, in MiB; 0 leaves the device default
    const long mb = tunable("L2_PERSIST_MB", 0);
    if (mb > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)mb << 20) != cudaSuccess) cudaGetLastError();
  }
  ok = ok && cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreateWithFlags(&ctx->begin_ev, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&ctx->side_done, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&ctx->lane_fork, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&ctx->lane_join, cudaEventDisableTiming) == cudaSuccess;
  for (auto &e : ctx->chunk_ev) ok = ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
  for (auto &e : ctx->side_ev) ok = ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
  if (!ok) {
    lcpc_b200_ctx_destroy(ctx);
    return LCPC_B200_ERR_CUDA;
  }
  *out = ctx;
  return LCPC_B200_OK;
}

void ctx_unref(lcpc_b200_ctx *ctx) {
  if (ctx->refs.fetch_sub(1) != 1) return;  // encodings of this context are still alive
  cudaSetDevice(ctx->device);
  if (ctx->stream) {
    cudaStreamSynchronize(ctx->stream);
    cudaStreamDestroy(ctx->stream);
  }
  if (ctx->copy_stream) {
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamDestroy(ctx->copy_stream);
  }
  if (ctx->side_stream) {
    cudaStreamSynchronize(ctx->side_stream);
    cudaStreamDestroy(ctx->side_stream);
  }
  for (auto &e : ctx->chunk_ev)
    if (e) cudaEventDestroy(e);
  for (auto &e : ctx->side_ev)
    if (e) cudaEventDestroy(e);
  if (ctx->begin_ev) cudaEventDestroy(ctx->begin_ev);
  if (ctx->side_done) cudaEventDestroy(ctx->side_done);
  if (ctx->lane_fork) cudaEventDestroy(ctx->lane_fork);
  if (ctx->lane_join) cudaEventDestroy(ctx->lane_join);
  for (auto &e : ctx->code_cache) expander_free(e.ptr);
  if (ctx->scratch) cudaFree(ctx->scratch);
  if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
  delete ctx;
}

void lcpc_b200_ctx_destroy(lcpc_b200_ctx *ctx) {
  if (ctx) ctx_unref(ctx);
}

const char *lcpc_b200_last_error(const lcpc_b200_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int lcpc_b200_ctx_device(const lcpc_b200_ctx *ctx) { return ctx ? ctx->device : -1; }
void *lcpc_b200_ctx_stream(const lcpc_b200_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
uint64_t lcpc_b200_ctx_launch_count(const lcpc_b200_ctx *ctx) { return ctx ? ctx->launches : 0; }

int lcpc_b200_ctx_synchronize(lcpc_b200_ctx *ctx) {
  if (!ctx) return LCPC_B200_ERR_BAD_ARG;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (int rc = bind_device(ctx)) return rc;
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return LCPC_B200_OK;
}

// page-locked host memory: a caller that keeps its coefficient vector in such a buffer gets full-rate,
// truly asynchronous PCIe copies in commit() (pageable memory is staged by the driver at about half rate)
int lcpc_b200_host_alloc(size_t bytes, void **out) {
  if (!out) return LCPC_B200_ERR_BAD_ARG;
  *out = nullptr;
  cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable);
  if (e != cudaSuccess) return e == cudaErrorMemoryAllocation ? LCPC_B200_ERR_OOM : LCPC_B200_ERR_CUDA;
  return LCPC_B200_OK;
}
void lcpc_b200_host_free(void *p) {
  if (p) cudaFreeHost(p);
}
int lcpc_b200_host_register(void *p, size_t bytes) {
  if (!p || !bytes) return LCPC_B200_ERR_BAD_ARG;
  return cudaHostRegister(p, bytes, cudaHostRegisterPortable) == cudaSuccess ? LCPC_B200_OK : LCPC_B200_ERR_CUDA;
}
int lcpc_b200_host_unregister(void *p) {
  if (!p) return LCPC_B200_ERR_BAD_ARG;
  return cudaHostUnregister(p) == cudaSuccess ? LCPC_B200_OK : LCPC_B200_ERR_CUDA;
}

// ---------------------------------------------------------------------------------------- encodings
int lcpc_b200_ligero_new(lcpc_b200_ctx *ctx, int field, size_t n_per_row, size_t n_cols, lcpc_b200_enc **out) {
  if (!ctx || !out) return LCPC_B200_ERR_BAD_ARG;
  *out = nullptr;
  std::lock_guard<std::mutex> g(ctx->mu);
  FieldHostInfo fi;
  if (!field_host_info(field, &fi)) return fail(ctx, LCPC_B200_ERR_BAD_ARG, "unknown field %d", field);
  // dims_ok: lcpc-ligero-pc/src/lib.rs:114-118
  if (!(n_per_row < n_cols && is_pow2(n_cols)))
    return fail(ctx, LCPC_B200_ERR_BAD_ARG, "ligero dims not ok: n_per_row=%zu n_cols=%zu", n_per_row, n_cols);
  unsigned log_n = log2_ceil(n_cols);
  if (log_n > fi.two_adicity) return fail(ctx, LCPC_B200_ERR_TOO_BIG, "FFTError::TooBig: 2^%u points", log_n);
  if (int rc = bind_device(ctx)) return rc;
  lcpc_b200_enc *e = new (std::nothrow) lcpc_b200_enc;
  if (!e) return LCPC_B200_ERR_OOM;
  e->ctx = ctx, e->kind = LCPC_B200_ENC_LIGERO, e->field = field;
  e->n_per_row = n_per_row, e->n_cols = n_cols, e->log_n = log_n;
  const int N = field_limbs32(field);
  const size_t half = n_cols / 2;
  uint32_t seed[16];  // [w, one]
  host_root(field, log_n, seed, seed + N);
  uint32_t *d_seed = nullptr;
  cudaError_t ce = cudaMalloc(&e->d_roots, (half ? half : 1) * N * 4);
  if (ce == cudaSuccess) ce = cudaMalloc(&d_seed, 2 * N * 4);
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(d_seed, seed, 2 * N * 4, cudaMemcpyHostToDevice, ctx->stream);
  if (ce == cudaSuccess) ce = launch_root_table(field, e->d_roots, d_seed, half, ctx->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
  if (d_seed) cudaFree(d_seed);
  if (ce != cudaSuccess) {
    if (e->d_roots) cudaFree(e->d_roots);
    delete e;
    return cuda_fail(ctx, ce, "ligero_new");
  }
  ctx->launches += half ? 1 : 0;
  ctx->refs.fetch_add(1);
  *out = e;
  return LCPC_B200_OK;
}

int lcpc_b200_sdig_new(lcpc_b200_ctx *ctx, int field, size_t n_levels, const lcpc_b200_csc *pre,
                       const lcpc_b200_csc *post, lcpc_b200_enc **out) {
  if (!ctx || !out || !pre || !post) return LCPC_B200_ERR_BAD_ARG;
  *out = nullptr;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (field_limbs32(field) < 0) return fail(ctx, LCPC_B200_ERR_BAD_ARG, "unknown field %d", field);
  if (n_levels == 0 || n_levels > 64) return fail(ctx, LCPC_B200_ERR_BAD_ARG, "bad level count %zu", n_levels);
  if (int rc = bind_device(ctx)) return rc;
  std::vector<CscView> vpre(n_levels), vpost(n_levels);
  for (size_t i = 0; i < n_levels; i++) {
    vpre[i] = {pre[i].m, pre[i].n, pre[i].ptrs, pre[i].idxs, pre[i].data};
    vpost[i] = {post[i].m, post[i].n, post[i].ptrs, post[i].idxs, post[i].data};
  }
  std::string err;
  ExpanderCode *code = nullptr;
  int rc = expander_build(field, n_levels, vpre.data(), vpost.data(), ctx->stream, &code, &err);
  if (rc != LCPC_B200_OK) return fail(ctx, rc, "sdig_new: %s", err.c_str());
  lcpc_b200_enc *e = new (std::nothrow) lcpc_b200_enc;
  if (!e) {
    expander_free(code);
    return LCPC_B200_ERR_OOM;
  }
  e->ctx = ctx, e->kind = LCPC_B200_ENC_SDIG, e->field = field;
  e->n_per_row = expander_n_in(code), e->n_cols = expander_codeword_length(code);
  e->code = code;
  ctx->refs.fetch_add(1);
  *out = e;
  return LCPC_B200_OK;
}

// SdigEncodingS::new / _new_from_np1 (lcpc-brakedown-pc/src/lib.rs:69-110) with the code drawn ON THE DEVICE
// (device_matgen.cu) and cached per context
int lcpc_b200_sdig_new_seeded(lcpc_b200_ctx *ctx, int field, int code, size_t n_per_row, uint64_t seed, lcpc_b200_enc **out) {
  if (!ctx || !out) return LCPC_B200_ERR_BAD_ARG;
  *out = nullptr;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (field_limbs32(field) < 0) return fail(ctx, LCPC_B200_ERR_BAD_ARG, "unknown field %d", field);
  lcpc::host::CodeSpec spec;
  if (!lcpc::host::sdig_code_spec(code, &spec)) return fail(ctx, LCPC_B200_ERR_BAD_ARG, "unknown code spec %d", code);
  if (int rc = bind_device(ctx)) return rc;
  ExpanderCode *ec = nullptr;
  for (auto &e : ctx->code_cache)
    if (e.field == field && e.code == code && e.n_per_row == n_per_row && e.seed == seed) ec = e.ptr;
  if (!ec) {
    std::vector<lcpc::host::LevelDims> pre_d, post_d;
    if (int rc = lcpc::host::sdig_level_dims(field, spec, n_per_row, &pre_d, &post_d)) return fail(ctx, rc, "sdig: bad row length %zu", n_per_row);
    std::vector<MatgenDims> pre(pre_d.size()), post(post_d.size());
    for (size_t i = 0; i < pre_d.size(); i++) pre[i] = {pre_d[i].n, pre_d[i].m, pre_d[i].d}, post[i] = {post_d[i].n, post_d[i].m, post_d[i].d};
    std::string err;
    int rc = device_matgen(field, seed, pre.size(), pre.data(), post.data(), ctx->stream, &ec, &err);
    if (rc != LCPC_B200_OK) return fail(ctx, rc, "sdig_new_seeded: %s", err.c_str());
    ctx->launches += 14 * 2 * pre.size();
    ctx->code_cache.push_back({field, code, n_per_row, seed, ec});  // the cache owns the first reference
  }
  lcpc_b200_enc *e = new (std::nothrow) lcpc_b200_enc;
  if (!e) return LCPC_B200_ERR_OOM;
  expander_retain(ec);
  e->ctx = ctx, e->kind = LCPC_B200_ENC_SDIG, e->field = field;
  e->n_per_row = expander_n_in(ec), e->n_cols = expander_codeword_length(ec);
  e->code = ec;
  ctx->refs.fetch_add(1);
  *out = e;
  return LCPC_B200_OK;
}

// one matrix of a device-generated code in the reference's own form (CsMat::new_csc, matgen.rs:187): exactly *d sorted
// row indices per input column, so ptrs[c] = c * d; idxs: n*d row indices, data: n*d elements (either may be NULL)
int lcpc_b200_enc_sdig_matrix(lcpc_b200_enc *enc, size_t level, int is_post, size_t *m, size_t *n, size_t *d, uint64_t *idxs,
                              uint64_t *data) {
  if (!enc || enc->kind != LCPC_B200_ENC_SDIG || !enc->code) return LCPC_B200_ERR_BAD_ARG;
  lcpc_b200_ctx *ctx = enc->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  const ExpanderCode *c = enc->code;
  if (level >= c->n_levels) return fail(ctx, LCPC_B200_ERR_BAD_ARG, "level %zu of %zu", level, c->n_levels);
  const DeviceCsr &M = c->mats[is_post ? c->n_levels + level : level];
  if (m) *m = M.m;
  if (n) *n = M.n;
  if (d) *d = M.csc_d;
  if (!idxs && !data) return LCPC_B200_OK;
  if (!M.csc_idx) return fail(ctx, LCPC_B200_ERR_UNSUPPORTED, "this encoding was built from host matrices; the host holds them");
  if (int rc = bind_device(ctx)) return rc;
  if (idxs) {
    std::vector<uint32_t> tmp(M.nnz);
    CU(ctx, cudaMemcpy(tmp.data(), M.csc_idx, M.nnz * 4, cudaMemcpyDeviceToHost));
    for (size_t k = 0; k < M.nnz; k++) idxs[k] = tmp[k];
  }
  if (data) CU(ctx, cudaMemcpy(data, M.csc_data, M.nnz * field_bytes(enc->field), cudaMemcpyDeviceToHost));
  return LCPC_B200_OK;
}
size_t lcpc_b200_enc_sdig_levels(const lcpc_b200_enc *enc) { return (enc && enc->code) ? enc->code->n_levels : 0; }

void enc_unref(lcpc_b200_enc *enc) {
  if (enc->refs.fetch_sub(1) != 1) return;  // commits made with this encoding are still alive
  lcpc_b200_ctx *ctx = enc->ctx;
  {
    std::lock_guard<std::mutex> g(ctx->mu);
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (enc->d_roots) cudaFree(enc->d_roots);
    if (enc->code) expander_free(enc->code);
    delete enc;
  }
  ctx_unref(ctx);
}

void lcpc_b200_enc_free(lcpc_b200_enc *enc) {
  if (enc) enc_unref(enc);
}

int lcpc_b200_enc_kind(const lcpc_b200_enc *enc) { return enc ? enc->kind : LCPC_B200_ERR_BAD_ARG; }
int lcpc_b200_enc_field(const lcpc_b200_enc *enc) { return enc ? enc->field : LCPC_B200_ERR_BAD_ARG; }

int lcpc_b200_enc_get_dims(const lcpc_b200_enc *enc, size_t len, size_t *n_rows, size_t *n_per_row, size_t *n_cols) {
  if (!enc) return LCPC_B200_ERR_BAD_ARG;
  if (n_rows) *n_rows = (len + enc->n_per_row - 1) / enc->n_per_row;
  if (n_per_row) *n_per_row = enc->n_per_row;
  if (n_cols) *n_cols = enc->n_cols;
  return LCPC_B200_OK;
}

int lcpc_b200_enc_dims_ok(const lcpc_b200_enc *enc, size_t n_per_row, size_t n_cols) {
  if (!enc) return 0;
  bool ok = n_per_row < n_cols && n_per_row == enc->n_per_row && n_cols == enc->n_cols;
  if (enc->kind == LCPC_B200_ENC_LIGERO) ok = ok && is_pow2(n_cols);
  return ok ? 1 : 0;
}

// encode n_rows rows: src (stride/valid) -> dst (stride n_cols); enqueues only
int encode_rows(lcpc_b200_enc *enc, const uint32_t *src, size_t src_stride, size_t valid, uint32_t *dst,
                size_t n_rows, void *enc_scratch, const Scatter *scatter) {
  lcpc_b200_ctx *ctx = enc->ctx;
  int nl = 0;
  cudaError_t ce;
  if (enc->kind == LCPC_B200_ENC_LIGERO) {
    ce = launch_ntt_rows(enc->field, src, src_stride, valid, dst, enc->n_cols, enc->d_roots, enc->log_n, n_rows,
                         ctx->stream, &nl, scatter);
  } else {
    SideLane lane{ctx->side_stream, ctx->lane_fork, ctx->lane_join};
    ce = expander_encode_rows(enc->code, src, src_stride, valid, dst, enc->n_cols, n_rows, enc_scratch, ctx->stream, &nl,
                              scatter, nullptr, 0, ~(size_t)0, scatter ? &lane : nullptr);
  }
  ctx->launches += nl;
  if (ce != cudaSuccess) return cuda_fail(ctx, ce, "encode");
  return LCPC_B200_OK;
}

size_t enc_scratch_bytes(const lcpc_b200_enc *enc, size_t n_rows) {
  return enc->kind == LCPC_B200_ENC_SDIG ? expander_scratch_bytes(enc->code, n_rows) : 0;
}

// pad + copy (lcpc-2d/src/lib.rs:636-645) overlapped with the per-row encode (:648-653): rows are
// independent, so the `len` coefficients at host address `src` cross PCIe in row-chunks on the copy
// stream (into d_coeffs, zero-padded to n_rows * n_per_row) while the engine stream encodes the chunks
// that have landed into d_comm.  `first_ev` (optional) is recorded on the engine stream when the first
// chunk is in.
// leaf hashing that trails the row encode: chunk k of every leaf input is hashed as soon as the rows it reads
// are encoded (see leaf_chunk_rows_end)
struct HashTrail {
  uint8_t *leaves;
  void *scratch;
  unsigned next_chunk, n_chunks;
  int launches;
};

// host destinations of an eager commit: each row-chunk of comm (and of the padded coefficients) goes back over
// PCIe on the side stream as soon as it is final, while later chunks are still arriving and being encoded
// (PCIe is full duplex: the download hides behind the upload)
struct HostOut {
  uint64_t *comm, *coeffs;
};

int encode_rows_from_host(lcpc_b200_enc *enc, const void *src, size_t len, uint32_t *d_coeffs, uint32_t *d_comm,
                          size_t n_rows, void *enc_scratch, cudaEvent_t first_ev, const Scatter *scatter, HashTrail *trail,
                          const HostOut *host_out, cudaEvent_t coeffs_free_ev) {
  lcpc_b200_ctx *ctx = enc->ctx;
  cudaStream_t st = ctx->stream;
  const size_t B = field_bytes(enc->field), N = B / 4;
  const size_t n_per_row = enc->n_per_row, padded = n_rows * n_per_row;
  const size_t row_bytes = n_per_row * B;
  size_t n_chunks = std::min<size_t>(enc->kind == LCPC_B200_ENC_LIGERO ? (size_t)std::min<long>(std::max<long>(tunable("H2D_MAX_CHUNKS", 15), 1), 15) : 4,
                                     n_rows);  // + one short tail chunk, below
  while (n_chunks > 1 && (n_rows / n_chunks) * row_bytes < ((size_t)4 << 20)) n_chunks--;
  if (enc->kind == LCPC_B200_ENC_LIGERO) {
    // a row-chunk is one launch per transform pass, n_cols / 1024 CTAs per row: chunks of fewer rows than about two
    // waves of the 592 resident CTA slots leave the GPU half empty and make the encode, not the copy, the long pole
    // (a rank of the 8-GPU commit holds 32 rows: 16 chunks of 2 rows ran at 2.95 ms end to end)
    const size_t ctas_per_row = std::max<size_t>(1, enc->n_cols >> 10);
    const size_t min_rows = std::max<size_t>(1, (size_t)tunable("H2D_MIN_CHUNK_CTAS", 1184) / ctas_per_row);
    while (n_chunks > 1 && n_rows / n_chunks < min_rows) n_chunks--;
  } else {
    // the expander chain has a dozen launches per chunk, some of them one CTA per batch row: below ~16 rows a chunk
    // is launch- and occupancy-bound, and a rank of the 8-GPU commit holds only 9 rows at 2^24
    while (n_chunks > 1 && n_rows / n_chunks < (size_t)std::max<long>(1, tunable("H2D_MIN_CHUNK_ROWS_SDIG", 16))) n_chunks--;
  }
  if (coeffs_free_ev) {
    // the caller knows when the last reader of d_coeffs finished (an event recorded behind it): the copy may start
    // then, under whatever the engine stream still has queued behind that reader (hashing of the previous commit)
    CU(ctx, cudaStreamWaitEvent(ctx->copy_stream, coeffs_free_ev, 0));
  } else {
    CU(ctx, cudaEventRecord(ctx->begin_ev, st));  // earlier readers of d_coeffs on the engine stream
    CU(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->begin_ev, 0));
  }
  // What is left to do once the last byte has crossed PCIe is pure latency on top of the copy: the last chunk is
  // therefore a short one (a quarter of the others, H2D_TAIL_DIV), cut off the chunk before it -- its launches are
  // small, but nothing runs beside them anyway
  size_t cuts[lcpc_b200_ctx::MAX_CHUNKS + 1];  // row-chunk k = rows [cuts[k], cuts[k + 1])
  size_t tail_rows = 0;
  if (enc->kind == LCPC_B200_ENC_LIGERO && n_chunks >= 2 && n_chunks < (size_t)lcpc_b200_ctx::MAX_CHUNKS) {
    const long div = tunable("H2D_TAIL_DIV", 4);
    if (div > 1 && n_rows / n_chunks >= 2 * (size_t)div) tail_rows = (n_rows / n_chunks) / (size_t)div, n_chunks += 1;
  }
  {
    // ... and the chunk before it is twice the tail (H2D_PRE_TAIL): a full-size chunk's transform takes about half
    // of its own copy time, i.e. it would still be running when the short tail has landed
    size_t pre_rows = 0;
    if (tail_rows && tunable("H2D_PRE_TAIL", 1) != 0 && tail_rows * std::max<size_t>(1, enc->n_cols >> 10) >= (size_t)std::max<long>(1, tunable("H2D_PRE_TAIL_MIN_CTAS", 512)) &&
        n_chunks >= 4 && n_rows >= 8 * tail_rows) {
      pre_rows = 2 * tail_rows;
      if (n_chunks < (size_t)lcpc_b200_ctx::MAX_CHUNKS) n_chunks += 1;  // else one body chunk fewer
    }
    const size_t special = (tail_rows ? 1 : 0) + (pre_rows ? 1 : 0);
    const size_t body_rows = n_rows - tail_rows - pre_rows, body_chunks = n_chunks - special;
    for (size_t k = 0; k <= body_chunks; k++) cuts[k] = k * body_rows / body_chunks;
    if (pre_rows) cuts[body_chunks + 1] = body_rows + pre_rows;
    cuts[n_chunks] = n_rows;
  }
  // Expander code with trailing leaf hashing: all leaf chunks but the last one read rows [0, boundary) only (BLAKE3
  // chunks are 1024 bytes: 62 Ft127 rows after the 32-byte prefix), so the tail is cut exactly there -- those chunks
  // (16 of the 19 compressions per column at 72 rows) are hashed while the tail rows cross PCIe -- and the chunk
  // before the tail is as short as the tail, so that its chain and that hashing fit under the tail's copy
  // (H2D_TAIL_SDIG=0: the equal chunks of before).  Measured at 2^24/Ft127: see DESIGN.md section 5.
  if (enc->kind == LCPC_B200_ENC_SDIG && trail && trail->n_chunks >= 2 && n_chunks >= 2 && tunable("H2D_TAIL_SDIG", 1) != 0) {
    const size_t boundary = leaf_chunk_rows_end(enc->field, n_rows, trail->n_chunks - 2);
    const size_t tail = n_rows - boundary;
    if (boundary >= 16 && tail >= 1 && tail * 3 <= n_rows) {
      const size_t pre = std::min(tail, boundary / 4);
      const size_t big_rows = boundary - pre;
      size_t big = std::min<size_t>(3, std::max<size_t>(1, big_rows / (size_t)std::max<long>(1, tunable("H2D_MIN_CHUNK_ROWS_SDIG", 16))));
      n_chunks = 0;
      for (size_t k = 0; k < big; k++) cuts[n_chunks++] = k * big_rows / big;
      cuts[n_chunks++] = big_rows;
      cuts[n_chunks++] = boundary;
      cuts[n_chunks] = n_rows;
    }
  }
  for (size_t k = 0; k < n_chunks; k++) {
    const size_t r0 = cuts[k], r1 = cuts[k + 1];
    const size_t e0 = r0 * n_per_row, e1 = std::min(r1 * n_per_row, len);
    if (e1 > e0)
      CU(ctx, cudaMemcpyAsync((uint8_t *)d_coeffs + e0 * B, (const uint8_t *)src + e0 * B, (e1 - e0) * B,
                              cudaMemcpyHostToDevice, ctx->copy_stream));
    if (k + 1 == n_chunks && padded > len)
      CU(ctx, cudaMemsetAsync((uint8_t *)d_coeffs + len * B, 0, (padded - len) * B, ctx->copy_stream));
    CU(ctx, cudaEventRecord(ctx->chunk_ev[k], ctx->copy_stream));
    CU(ctx, cudaStreamWaitEvent(st, ctx->chunk_ev[k], 0));
    if (k == 0 && first_ev) CU(ctx, cudaEventRecord(first_ev, st));
    Scatter sc;
    if (scatter) {
      sc = *scatter;
      sc.row0 += r0;
    }
    if (int rc = encode_rows(enc, d_coeffs + r0 * n_per_row * N, n_per_row, n_per_row, d_comm + r0 * enc->n_cols * N,
                             r1 - r0, enc_scratch, scatter ? &sc : nullptr))
      return rc;
    if (host_out && (host_out->comm || host_out->coeffs)) {
      CU(ctx, cudaEventRecord(ctx->side_ev[k], st));
      CU(ctx, cudaStreamWaitEvent(ctx->side_stream, ctx->side_ev[k], 0));
      if (host_out->comm)
        CU(ctx, cudaMemcpyAsync((uint8_t *)host_out->comm + r0 * enc->n_cols * B, (uint8_t *)d_comm + r0 * enc->n_cols * B,
                                (r1 - r0) * enc->n_cols * B, cudaMemcpyDeviceToHost, ctx->side_stream));
      if (host_out->coeffs)
        CU(ctx, cudaMemcpyAsync((uint8_t *)host_out->coeffs + r0 * n_per_row * B, (uint8_t *)d_coeffs + r0 * n_per_row * B,
                                (r1 - r0) * n_per_row * B, cudaMemcpyDeviceToHost, ctx->side_stream));
    }
    if (trail && k + 1 < n_chunks) {  // the chunks still open after the last row-chunk are hashed by the caller
      unsigned ready = trail->next_chunk;
      while (ready < trail->n_chunks && leaf_chunk_rows_end(enc->field, n_rows, ready) <= r1) ready++;
      if (ready > trail->next_chunk) {
        cudaError_t ce = launch_leaf_chunks(enc->field, d_comm, n_rows, enc->n_cols, enc->n_cols, trail->leaves, trail->scratch,
                                            trail->next_chunk, ready - trail->next_chunk, st);
        if (ce != cudaSuccess) return cuda_fail(ctx, ce, "hash_columns");
        ctx->launches += 1, trail->launches += 1;
        trail->next_chunk = ready;
      }
    }
  }
  return LCPC_B200_OK;
}

// validate a scatter descriptor of the ABI and turn it into the kernels' by-value form
static int make_scatter(lcpc_b200_enc *enc, const lcpc_b200_scatter *in, Scatter *out) {
  lcpc_b200_ctx *ctx = enc->ctx;
  if (!in || !in->starts || !in->dst || in->n_blocks == 0 || in->n_blocks > (size_t)MAX_SCATTER)
    return fail(ctx, LCPC_B200_ERR_BAD_ARG, "scatter: 1..%d column blocks expected", MAX_SCATTER);
  if (in->starts[0] != 0 || in->starts[in->n_blocks] != enc->n_cols)
    return fail(ctx, LCPC_B200_ERR_BAD_ARG, "scatter: column blocks must cover [0, n_cols)");
  out->n_blocks = (unsigned)in->n_blocks, out->row0 = in->row0;
  for (size_t h = 0; h <= in->n_blocks; h++) {
    if (h && in->starts[h] < in->starts[h - 1]) return fail(ctx, LCPC_B200_ERR_BAD_ARG, "scatter: starts not monotone");
    out->starts[h] = in->starts[h];
  }
  for (size_t h = 0; h < in->n_blocks; h++) {
    if (!in->dst[h] && in->starts[h + 1] > in->starts[h]) return fail(ctx, LCPC_B200_ERR_BAD_ARG, "scatter: null block %zu", h);
    out->dst[h] = (uint32_t *)in->dst[h];
  }
  return LCPC_B200_OK;
}

int lcpc_b200_encode_rows_scatter_dev(lcpc_b200_enc *enc, const uint64_t *d_src, size_t src_stride, size_t valid,
                                      uint64_t *d_tmp, size_t n_rows, const lcpc_b200_scatter *scatter) {
  if (!enc || ((!d_src || !d_tmp) && n_rows)) return LCPC_B200_ERR_BAD_ARG;
  lcpc_b200_ctx *ctx = enc->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (valid > enc->n_cols || valid > src_stride)
    return fail(ctx, LCPC_B200_ERR_BAD_ARG, "valid %zu exceeds row (stride %zu, n_cols %zu)", valid, src_stride, enc->n_cols);
  Scatter sc;
  if (int rc = make_scatter(enc, scatter, &sc)) return rc;
  if (int rc = bind_device(ctx)) return rc;
  if (int rc = ensure_scratch(ctx, enc_scratch_bytes(enc, n_rows))) return rc;
  return encode_rows(enc, (const uint32_t *)d_src, src_stride, valid, (uint32_t *)d_tmp, n_rows, ctx->scratch, &sc);
}

int lcpc_b200_encode_rows_scatter_h2d(lcpc_b200_enc *enc, const uint64_t *rows, size_t len, uint64_t *d_coeffs,
                                      uint64_t *d_tmp, size_t n_rows, const lcpc_b200_scatter *scatter) {
  if (!enc || !rows || !d_coeffs || !d_tmp || n_rows == 0) return LCPC_B200_ERR_BAD_ARG;
  lcpc_b200_ctx *ctx = enc->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (len > n_rows * enc->n_per_row || len + enc->n_per_row <= n_rows * enc->n_per_row)
    return fail(ctx, LCPC_B200_ERR_BAD_ARG, "encode_rows_scatter_h2d: %zu elements do not fill %zu rows", len, n_rows);
  Scatter sc;
  if (int rc = make_scatter(enc, scatter, &sc)) return rc;
  if (int rc = bind_device(ctx)) return rc;
  if (int rc = ensure_scratch(ctx, enc_scratch_bytes(enc, n_rows))) return rc;
  return encode_rows_from_host(enc, rows, len, (uint32_t *)d_coeffs, (uint32_t *)d_tmp, n_rows, ctx->scratch, nullptr, &sc);
}

int lcpc_b200_encode_rows_h2d(lcpc_b200_enc *enc, const uint64_t *rows, size_t len, uint64_t *d_coeffs, uint64_t *d_dst,
                              size_t n_rows) {
  if (!enc || !rows || !d_coeffs || !d_dst || n_rows == 0) return LCPC_B200_ERR_BAD_ARG;
  lcpc_b200_ctx *ctx = enc->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (len > n_rows * enc->n_per_row || len + enc->n_per_row <= n_rows * enc->n_per_row)
    return fail(ctx, LCPC_B200_ERR_BAD_ARG, "encode_rows_h2d: %zu elements do not fill %zu rows", len, n_rows);
  if (int rc = bind_device(ctx)) return rc;
  if (int rc = ensure_scratch(ctx, enc_scratch_bytes(enc, n_rows))) return rc;
  return encode_rows_from_host(enc, rows, len, (uint32_t *)d_coeffs, (uint32_t *)d_dst, n_rows, ctx->scratch, nullptr);
}

int lcpc_b200_encode_dev(lcpc_b200_enc *enc, uint64_t *d_rows, size_t n_rows, size_t valid) {
  if (!enc || (!d_rows && n_rows)) return LCPC_B200_ERR_BAD_ARG;
  lcpc_b200_ctx *ctx = enc->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (valid > enc->n_cols) return fail(ctx, LCPC_B200_ERR_BAD_ARG, "valid %zu > n_cols %zu", valid, enc->n_cols);
  if (int rc = bind_device(ctx)) return rc;
  if (int rc = ensure_scratch(ctx, enc_scratch_bytes(enc, n_rows))) return rc;
  return encode_rows(enc, (const uint32_t *)d_rows, enc->n_cols, valid, (uint32_t *)d_rows, n_rows, ctx->scratch);
}

int lcpc_b200_encode_rows_dev(lcpc_b200_enc *enc, const uint64_t *d_src, size_t src_stride, size_t valid,
                              uint64_t *d_dst, size_t n_rows) {
  if (!enc || ((!d_src || !d_dst) && n_rows)) return LCPC_B200_ERR_BAD_ARG;
  lcpc_b200_ctx *ctx = enc->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (valid > enc->n_cols || valid > src_stride)
    return fail(ctx, LCPC_B200_ERR_BAD_ARG, "valid %zu exceeds row (stride %zu, n_cols %zu)", valid, src_stride, enc->n_cols);
  if (int rc = bind_device(ctx)) return rc;
  if (int rc = ensure_scratch(ctx, enc_scratch_bytes(enc, n_rows))) return rc;
  return encode_rows(enc, (const uint32_t *)d_src, src_stride, valid, (uint32_t *)d_dst, n_rows, ctx->scratch);
}

int lcpc_b200_encode(lcpc_b200_enc *enc, uint64_t *rows, size_t n_rows) {
  if (!enc || (!rows && n_rows)) return LCPC_B200_ERR_BAD_ARG;
  if (n_rows == 0) return LCPC_B200_OK;
  lcpc_b200_ctx *ctx = enc->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (int rc = bind_device(ctx)) return rc;
  // rows and the encoder's own scratch share the context's grow-only pool: the verifier calls this once per proof
  // with one or two rows (lcpc-2d/src/lib.rs:886, :918), so after the first call nothing is allocated or freed
  const size_t bytes = n_rows * enc->n_cols * field_bytes(enc->field);
  const size_t rows_al = (bytes + 255) & ~(size_t)255;
  if (int rc = ensure_scratch(ctx, rows_al + enc_scratch_bytes(enc, n_rows))) return rc;
  uint32_t *d = (uint32_t *)ctx->scratch;
  void *enc_scratch = (uint8_t *)ctx->scratch + rows_al;
  CU(ctx, cudaMemcpyAsync(d, rows, bytes, cudaMemcpyHostToDevice, ctx->stream));
  // the reference transforms the whole row: Ligero reads all n_cols entries, Brakedown the first n_per_row
  const size_t valid = enc->kind == LCPC_B200_ENC_LIGERO ? enc->n_cols : enc->n_per_row;
  if (int rc = encode_rows(enc, d, enc->n_cols, valid, d, n_rows, enc_scratch)) return rc;
  CU(ctx, cudaMemcpyAsync(rows, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return LCPC_B200_OK;
}

// ------------------------------------------------------------------------------------------ commit
static void commit_release(lcpc_b200_commit *c) {
  cudaFree(c->d_coeffs);
  cudaFree(c->d_comm);
  cudaFree(c->d_hashes);
  cudaFree(c->d_hash_scratch);
  cudaFree(c->d_enc_scratch);
  cudaFree(c->d_tensor);
  cudaFree(c->d_poly);
  cudaFree(c->d_key);
  cudaFree(c->d_repr);
  if (c->h_poly) cudaFreeHost(c->h_poly);
  for (auto &e : c->ev)
    if (e) cudaEventDestroy(e);
  delete c;
}

static int commit_alloc(lcpc_b200_enc *enc, size_t len, lcpc_b200_commit **out) {
  lcpc_b200_ctx *ctx = enc->ctx;
  if (len == 0) return fail(ctx, LCPC_B200_ERR_BAD_ARG, "commit: empty coefficient vector");
  size_t n_rows = (len + enc->n_per_row - 1) / enc->n_per_row, n_per_row = enc->n_per_row, n_cols = enc->n_cols;
  // asserts at lcpc-2d/src/lib.rs:630-632
  if (!(n_rows * n_per_row >= len && (n_rows - 1) * n_per_row < len && lcpc_b200_enc_dims_ok(enc, n_per_row, n_cols)))
    return fail(ctx, LCPC_B200_ERR_BAD_ARG, "commit: inconsistent dims");
  // n_cols.checked_next_power_of_two() (:656-658)
  unsigned lg = log2_ceil(n_cols);
  if (lg >= 8 * sizeof(size_t) - 2) return fail(ctx, LCPC_B200_ERR_TOO_BIG, "commit: n_cols too big");
  lcpc_b200_commit *c = new (std::nothrow) lcpc_b200_commit;
  if (!c) return LCPC_B200_ERR_OOM;
  c->enc = enc, c->n_rows = n_rows, c->n_per_row = n_per_row, c->n_cols = n_cols, c->np2 = (size_t)1 << lg;
  const size_t B = field_bytes(enc->field);
  size_t hs = hash_scratch_bytes(enc->field, n_rows, n_cols);
  size_t es = enc_scratch_bytes(enc, n_rows);
  cudaError_t ce = cudaMalloc(&c->d_coeffs, n_rows * n_per_row * B);
  if (ce == cudaSuccess) ce = cudaMalloc(&c->d_comm, n_rows * n_cols * B);
  if (ce == cudaSuccess) ce = cudaMalloc(&c->d_hashes, (2 * c->np2 - 1) * 32);
  if (ce == cudaSuccess && hs) ce = cudaMalloc(&c->d_hash_scratch, hs);
  if (ce == cudaSuccess && es) ce = cudaMalloc(&c->d_enc_scratch, es);
  if (ce == cudaSuccess) ce = cudaMalloc(&c->d_tensor, n_rows * B);
  if (ce == cudaSuccess) ce = cudaMalloc(&c->d_poly, n_per_row * B);
  for (auto &e : c->ev)
    if (ce == cudaSuccess) ce = cudaEventCreate(&e);
  if (ce != cudaSuccess) {
    commit_release(c);
    return cuda_fail(ctx, ce, "commit: cudaMalloc");
  }
  enc->refs.fetch_add(1);
  *out = c;
  return LCPC_B200_OK;
}

// make c->d_comm (row-major) valid; enqueue only
static int ensure_comm(lcpc_b200_commit *c) {
  if (!c->comm_in_w) return LCPC_B200_OK;
  lcpc_b200_ctx *ctx = c->enc->ctx;
  cudaError_t ce = expander_untranspose(c->enc->code, c->d_enc_scratch, c->d_comm, c->n_cols, c->n_rows, ctx->stream);
  ctx->launches += 1;
  if (ce != cudaSuccess) return cuda_fail(ctx, ce, "comm from the work buffer");
  c->comm_in_w = false;
  return LCPC_B200_OK;
}

// make c->d_coeffs (row-major) valid; enqueue only
static int ensure_coeffs(lcpc_b200_commit *c) {
  if (!c->coeffs_in_w) return LCPC_B200_OK;
  lcpc_b200_ctx *ctx = c->enc->ctx;
  cudaError_t ce = expander_untranspose(c->enc->code, c->d_enc_scratch, c->d_coeffs, c->n_per_row, c->n_rows, ctx->stream, c->n_per_row);
  ctx->launches += 1;
  if (ce != cudaSuccess) return cuda_fail(ctx, ce, "coefficients from the work buffer");
  c->coeffs_in_w = false;
  return LCPC_B200_OK;
}

// collapse_columns over the commit's coefficients wherever they are
static cudaError_t collapse_commit(lcpc_b200_commit *c, int *nl) {
  lcpc_b200_ctx *ctx = c->enc->ctx;
  if (c->coeffs_in_w)
    return launch_collapse(c->enc->field, (const uint32_t *)c->d_enc_scratch, 1, c->d_tensor, c->d_poly, c->n_rows, c->n_per_row, nullptr,
                           ctx->stream, nl, c->n_rows);
  return launch_collapse(c->enc->field, c->d_coeffs, c->n_per_row, c->d_tensor, c->d_poly, c->n_rows, c->n_per_row, nullptr, ctx->stream, nl);
}

// enqueue the whole commit pipeline; src is host or device memory holding `len` elements
static int commit_run(lcpc_b200_commit *c, const void *src, size_t len, cudaMemcpyKind kind,
                      const HostOut *host_out = nullptr) {
  lcpc_b200_enc *enc = c->enc;
  lcpc_b200_ctx *ctx = enc->ctx;
  const size_t B = field_bytes(enc->field);
  if ((len + c->n_per_row - 1) / c->n_per_row != c->n_rows)
    return fail(ctx, LCPC_B200_ERR_BAD_ARG, "commit: length %zu does not give %zu rows", len, c->n_rows);
  cudaStream_t st = ctx->stream;
  CU(ctx, cudaEventRecord(c->ev[0], st));
  const size_t padded = c->n_rows * c->n_per_row;
  uint64_t l0 = ctx->launches;
  HashTrail trail{c->d_hashes, c->d_hash_scratch, 0, leaf_chunk_count(enc->field, c->n_rows), 0};
  size_t early_cols = 0;  // leading columns whose leaf chunks were hashed on the side stream already
  if ((const void *)c->d_coeffs == src)  // re-commit of the commit's own coefficient rows: they have to be there
    if (int rc = ensure_coeffs(c)) return rc;
  c->comm_in_w = c->coeffs_in_w = false;
  if (kind == cudaMemcpyHostToDevice && c->n_rows > 1) {
    if (int rc = encode_rows_from_host(enc, src, len, c->d_coeffs, c->d_comm, c->n_rows, c->d_enc_scratch, c->ev[1], nullptr,
                                       &trail, host_out))
      return rc;
    if (host_out) host_out = nullptr;  // handled chunk by chunk
  } else if (kind == cudaMemcpyDeviceToDevice && enc->kind == LCPC_B200_ENC_LIGERO && padded == len && enc->log_n > 0 &&
             (const void *)c->d_coeffs != src) {
    // pad + copy (:636-645) folded into the first transform pass: it reads the caller's coefficient rows and
    // stores the commit's own copy on the way.  Rows can go in chunks (LCPC_B200_DEV_CHUNKS) so that leaf-input
    // chunks whose rows are done are hashed on the side stream while later rows are transformed; measured at
    // 2^24: 1/4/8/16 chunks = 5.63/5.72/5.90/6.14 ms -- the transform CTAs hold every register of the SM, so the
    // hash CTAs cannot co-reside and the extra launch tails cost more than the overlap gains.  Default: 1.
    CU(ctx, cudaEventRecord(c->ev[1], st));
    const size_t want_chunks = (size_t)std::min<long>(std::max<long>(tunable("DEV_CHUNKS", 1), 1), lcpc_b200_ctx::MAX_CHUNKS);
    const size_t N = B / 4;
    size_t n_rc = trail.n_chunks > 1 ? std::min(want_chunks, c->n_rows) : 1;
    while (n_rc > 1 && (c->n_rows / n_rc) * c->n_cols * B < ((size_t)32 << 20)) n_rc--;
    bool side_used = false;
    for (size_t k = 0; k < n_rc; k++) {
      const size_t r0 = k * c->n_rows / n_rc, r1 = (k + 1) * c->n_rows / n_rc;
      int nl = 0;
      cudaError_t ce = launch_ntt_rows(enc->field, (const uint32_t *)src + r0 * c->n_per_row * N, c->n_per_row, c->n_per_row,
                                       c->d_comm + r0 * c->n_cols * N, c->n_cols, enc->d_roots, enc->log_n, r1 - r0, st, &nl,
                                       nullptr, c->d_coeffs + r0 * c->n_per_row * N, c->n_per_row);
      ctx->launches += nl;
      if (ce != cudaSuccess) return cuda_fail(ctx, ce, "encode");
      if (k + 1 < n_rc) {
        unsigned ready = trail.next_chunk;
        while (ready < trail.n_chunks && leaf_chunk_rows_end(enc->field, c->n_rows, ready) <= r1) ready++;
        if (ready > trail.next_chunk) {
          CU(ctx, cudaEventRecord(ctx->side_ev[k], st));
          CU(ctx, cudaStreamWaitEvent(ctx->side_stream, ctx->side_ev[k], 0));
          ce = launch_leaf_chunks(enc->field, c->d_comm, c->n_rows, c->n_cols, c->n_cols, trail.leaves, trail.scratch,
                                  trail.next_chunk, ready - trail.next_chunk, ctx->side_stream);
          if (ce != cudaSuccess) return cuda_fail(ctx, ce, "hash_columns");
          ctx->launches += 1, trail.launches += 1;
          trail.next_chunk = ready;
          side_used = true;
        }
      }
    }
    if (side_used) {
      CU(ctx, cudaEventRecord(ctx->side_done, ctx->side_stream));
      CU(ctx, cudaStreamWaitEvent(st, ctx->side_done, 0));
    }
  } else if (kind == cudaMemcpyDeviceToDevice && enc->kind == LCPC_B200_ENC_SDIG && (const void *)c->d_coeffs != src) {
    // the same for the expander code: the transpose into the work buffer stores the copy and supplies the
    // zero padding of a short last row
    CU(ctx, cudaEventRecord(c->ev[1], st));
    // The code is systematic: columns [0, n_per_row) of comm ARE the coefficient rows (encode.rs: the codeword starts
    // with x_0), two thirds of all columns.  Their leaf digests do not depend on the sparse products at all, so they
    // are hashed straight from the caller's rows on the side stream while the chain runs: BLAKE3 lives on the ALU
    // pipe, the chain is bound by L2 gather bandwidth.  (A short last row reads as zeros past the caller's `len`
    // coefficients, exactly the padding of lcpc-2d/src/lib.rs:636-645.)
    // Measured at 2^24 (profiles/r02_ab_brakedown_early_hash.jsonl): the systematic columns' hash moves into the encode
    // phase (+0.17 ms) and out of the leaf phase (-0.15 ms) -- the sparse products hold every register of the SM, so
    // the two kernels take turns instead of sharing it; 1.675 ms per commit against 1.652.  Off by default.
    if (trail.n_chunks >= 1 && tunable("SDIG_EARLY_HASH", 0) != 0) {
      CU(ctx, cudaEventRecord(ctx->lane_fork, st));
      CU(ctx, cudaStreamWaitEvent(ctx->side_stream, ctx->lane_fork, 0));
      cudaError_t he = launch_leaf_chunks_range(enc->field, (const uint32_t *)src, c->n_rows, c->n_per_row, c->n_per_row, c->d_hashes,
                                                c->d_hash_scratch, 0, trail.n_chunks, c->n_cols, 0, ctx->side_stream, len);
      if (he != cudaSuccess) return cuda_fail(ctx, he, "hash_columns (systematic part)");
      CU(ctx, cudaEventRecord(ctx->side_done, ctx->side_stream));
      ctx->launches += 1, trail.launches += 1;
      early_cols = c->n_per_row;
    }
    // SDIG_LAZY_COMM (default on): no final transpose; the columns are hashed (and later opened) from the work buffer
    const bool lazy = !host_out && early_cols == 0 && tunable("SDIG_LAZY_COMM", 1) != 0;
    int nl = 0;
    cudaError_t ce = expander_encode_rows(enc->code, (const uint32_t *)src, c->n_per_row, c->n_per_row, lazy ? nullptr : c->d_comm,
                                          c->n_cols, c->n_rows, c->d_enc_scratch, st, &nl, nullptr, lazy ? nullptr : c->d_coeffs, c->n_per_row,
                                          len);
    ctx->launches += nl;
    if (ce != cudaSuccess) return cuda_fail(ctx, ce, "encode");
    c->comm_in_w = c->coeffs_in_w = lazy;
  } else {
    // pad + copy (lcpc-2d/src/lib.rs:636-645): coeffs = coeffs_in || zeros
    CU(ctx, cudaMemcpyAsync(c->d_coeffs, src, len * B, kind, st));
    if (padded > len) CU(ctx, cudaMemsetAsync((uint8_t *)c->d_coeffs + len * B, 0, (padded - len) * B, st));
    CU(ctx, cudaEventRecord(c->ev[1], st));
    // per-row encode (:648-653); reads the padded coefficient rows, writes comm
    if (int rc = encode_rows(enc, c->d_coeffs, c->n_per_row, c->n_per_row, c->d_comm, c->n_rows, c->d_enc_scratch)) return rc;
  }
  c->encode_launches = (int)(ctx->launches - l0);
  CU(ctx, cudaEventRecord(c->ev[2], st));
  // leaves beyond n_cols stay Output::default() = zeros (:665, :696)
  if (c->np2 > c->n_cols) CU(ctx, cudaMemsetAsync(c->d_hashes + c->n_cols * 32, 0, (c->np2 - c->n_cols) * 32, st));
  // column leaves (:706-745): whatever chunks did not already trail the encode, then the per-column chunk merge
  int nl = 0;
  cudaError_t ce =
      c->comm_in_w
          ? launch_leaf_chunks_range(enc->field, (const uint32_t *)c->d_enc_scratch, c->n_rows, c->n_cols, /*row_stride=*/1, c->d_hashes,
                                     c->d_hash_scratch, trail.next_chunk, trail.n_chunks - trail.next_chunk, c->n_cols, 0, st,
                                     ~(size_t)0, /*col_stride=*/c->n_rows)
          : launch_leaf_chunks_range(enc->field, c->d_comm + early_cols * (B / 4), c->n_rows, c->n_cols - early_cols, c->n_cols,
                                     c->d_hashes, c->d_hash_scratch, trail.next_chunk, trail.n_chunks - trail.next_chunk,
                                     c->n_cols, early_cols, st);
  if (early_cols) CU(ctx, cudaStreamWaitEvent(st, ctx->side_done, 0));
  if (ce == cudaSuccess) ce = launch_leaf_merge(enc->field, c->n_rows, c->n_cols, c->d_hashes, c->d_hash_scratch, st, &nl);
  nl += 1;
  ctx->launches += nl;
  c->hash_launches = nl + trail.launches;
  if (ce != cudaSuccess) return cuda_fail(ctx, ce, "hash_columns");
  CU(ctx, cudaEventRecord(c->ev[3], st));
  ce = launch_merkle_tree(c->d_hashes, c->np2, st, &nl);
  ctx->launches += nl;
  c->merkle_launches = nl;
  if (ce != cudaSuccess) return cuda_fail(ctx, ce, "merkle_tree");
  CU(ctx, cudaEventRecord(c->ev[4], st));
  if (host_out) {  // routes without row-chunks: whole arrays after the fact
    if (host_out->comm)
      CU(ctx, cudaMemcpyAsync(host_out->comm, c->d_comm, c->n_rows * c->n_cols * B, cudaMemcpyDeviceToHost, st));
    if (host_out->coeffs)
      CU(ctx, cudaMemcpyAsync(host_out->coeffs, c->d_coeffs, c->n_rows * c->n_per_row * B, cudaMemcpyDeviceToHost, st));
  }
  return LCPC_B200_OK;
}

static int commit_new_impl(lcpc_b200_enc *enc, const void *src, size_t len, cudaMemcpyKind kind, lcpc_b200_commit **out) {
  if (!enc || !out || (!src && len)) return LCPC_B200_ERR_BAD_ARG;
  *out = nullptr;
  lcpc_b200_ctx *ctx = enc->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (int rc = bind_device(ctx)) return rc;
  lcpc_b200_commit *c = nullptr;
  if (int rc = commit_alloc(enc, len, &c)) return rc;
  int rc = commit_run(c, src, len, kind);
  if (rc == LCPC_B200_OK) {
    cudaError_t ce = cudaStreamSynchronize(ctx->stream);
    if (ce != cudaSuccess) rc = cuda_fail(ctx, ce, "commit: synchronize");
  }
  if (rc != LCPC_B200_OK) {
    commit_release(c);
    enc->refs.fetch_sub(1);  // the caller's own reference keeps enc alive
    return rc;
  }
  *out = c;
  return LCPC_B200_OK;
}

int lcpc_b200_commit_new(lcpc_b200_enc *enc, const uint64_t *coeffs_in, size_t len, lcpc_b200_commit **out) {
  return commit_new_impl(enc, coeffs_in, len, cudaMemcpyHostToDevice, out);
}
int lcpc_b200_commit_new_dev(lcpc_b200_enc *enc, const uint64_t *d_coeffs_in, size_t len, lcpc_b200_commit **out) {
  return commit_new_impl(enc, d_coeffs_in, len, cudaMemcpyDeviceToDevice, out);
}

static int commit_rerun_impl(lcpc_b200_commit *c, const void *src, size_t len, cudaMemcpyKind kind, bool sync) {
  if (!c || !src) return LCPC_B200_ERR_BAD_ARG;
  lcpc_b200_ctx *ctx = c->enc->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (int rc = bind_device(ctx)) return rc;
  int rc = commit_run(c, src, len, kind);
  if (rc == LCPC_B200_OK && sync) CU(ctx, cudaStreamSynchronize(ctx->stream));
  return rc;
}
// device input: enqueue only (the caller times with events on lcpc_b200_ctx_stream and synchronises)
int lcpc_b200_commit_rerun_dev(lcpc_b200_commit *c, const uint64_t *d_coeffs_in, size_t len) {
  return commit_rerun_impl(c, d_coeffs_in, len, cudaMemcpyDeviceToDevice, false);
}
int lcpc_b200_commit_rerun(lcpc_b200_commit *c, const uint64_t *coeffs_in, size_t len) {
  return commit_rerun_impl(c, coeffs_in, len, cudaMemcpyHostToDevice, true);
}

// Deserialize for LcCommit (lcpc-2d/src/lib.rs:256-268) onto the device: the fields of a commitment that was made
// elsewhere (or earlier) become a device-resident commit that prove() can use.  Nothing is recomputed -- like the
// reference, which trusts the deserialized fields and only checks their consistency in prove() (check_comm, :672-688).
int lcpc_b200_commit_from_host(lcpc_b200_enc *enc, const uint64_t *comm, size_t comm_len, const uint64_t *coeffs,
                               size_t coeffs_len, const uint8_t *hashes, size_t n_hashes, size_t n_rows,
                               lcpc_b200_commit **out) {
  if (!enc || !out || !comm || !coeffs || !hashes) return LCPC_B200_ERR_BAD_ARG;
  *out = nullptr;
  lcpc_b200_ctx *ctx = enc->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  const size_t n_per_row = enc->n_per_row, n_cols = enc->n_cols;
  // check_comm: comm.len() == n_rows * n_cols, coeffs.len() == n_rows * n_per_row, hashes.len() == 2 * np2 - 1
  if (n_rows == 0 || comm_len != n_rows * n_cols || coeffs_len != n_rows * n_per_row ||
      n_hashes != 2 * ((size_t)1 << log2_ceil(n_cols)) - 1)
    return fail(ctx, LCPC_B200_ERR_BAD_ARG, "inconsistent commitment fields");  // ProverError::Commit
  if (int rc = bind_device(ctx)) return rc;
  lcpc_b200_commit *c = nullptr;
  if (int rc = commit_alloc(enc, (n_rows - 1) * n_per_row + 1, &c)) return rc;  // any length with this row count
  const size_t B = field_bytes(enc->field);
  cudaError_t ce = cudaMemcpyAsync(c->d_comm, comm, comm_len * B, cudaMemcpyHostToDevice, ctx->stream);
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(c->d_coeffs, coeffs, coeffs_len * B, cudaMemcpyHostToDevice, ctx->stream);
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(c->d_hashes, hashes, n_hashes * 32, cudaMemcpyHostToDevice, ctx->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
  if (ce != cudaSuccess) {
    commit_release(c);
    enc->refs.fetch_sub(1);
    return cuda_fail(ctx, ce, "commit_from_host");
  }
  *out = c;
  return LCPC_B200_OK;
}

void lcpc_b200_commit_free(lcpc_b200_commit *c) {
  if (!c) return;
  lcpc_b200_enc *enc = c->enc;
  {
    lcpc_b200_ctx *ctx = enc->ctx;
    std::lock_guard<std::mutex> g(ctx->mu);
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    commit_release(c);
  }
  enc_unref(enc);
}

int lcpc_b200_commit_dims(const lcpc_b200_commit *c, size_t *n_rows, size_t *n_per_row, size_t *n_cols, size_t *n_hashes) {
  if (!c) return LCPC_B200_ERR_BAD_ARG;
  if (n_rows) *n_rows = c->n_rows;
  if (n_per_row) *n_per_row = c->n_per_row;
  if (n_cols) *n_cols = c->n_cols;
  if (n_hashes) *n_hashes = 2 * c->np2 - 1;
  return LCPC_B200_OK;
}

int lcpc_b200_commit_root(lcpc_b200_commit *c, uint8_t root[32]) {
  if (!c || !root) return LCPC_B200_ERR_BAD_ARG;
  lcpc_b200_ctx *ctx = c->enc->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (int rc = bind_device(ctx)) return rc;
  CU(ctx, cudaMemcpyAsync(root, c->d_hashes + (2 * c->np2 - 2) * 32, 32, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return LCPC_B200_OK;
}

int lcpc_b200_commit_download(lcpc_b200_commit *c, uint64_t *comm, uint64_t *coeffs, uint8_t *hashes) {
  if (!c) return LCPC_B200_ERR_BAD_ARG;
  lcpc_b200_ctx *ctx = c->enc->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (int rc = bind_device(ctx)) return rc;
  const size_t B = field_bytes(c->enc->field);
  if (comm) {
    if (int rc = ensure_comm(c)) return rc;
    CU(ctx, cudaMemcpyAsync(comm, c->d_comm, c->n_rows * c->n_cols * B, cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (coeffs) {
    if (int rc = ensure_coeffs(c)) return rc;
    CU(ctx, cudaMemcpyAsync(coeffs, c->d_coeffs, c->n_rows * c->n_per_row * B, cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (hashes) CU(ctx, cudaMemcpyAsync(hashes, c->d_hashes, (2 * c->np2 - 1) * 32, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return LCPC_B200_OK;
}

int lcpc_b200_commit_phase_times(lcpc_b200_commit *c, float ms[4], int launches[3]) {
  if (!c || !ms) return LCPC_B200_ERR_BAD_ARG;
  lcpc_b200_ctx *ctx = c->enc->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (int rc = bind_device(ctx)) return rc;
  CU(ctx, cudaEventSynchronize(c->ev[4]));
  for (int i = 0; i < 4; i++) CU(ctx, cudaEventElapsedTime(&ms[i], c->ev[i], c->ev[i + 1]));
  if (launches) launches[0] = c->encode_launches, launches[1] = c->hash_launches, launches[2] = c->merkle_launches;
  return LCPC_B200_OK;
}

int lcpc_b200_commit_device_ptrs(lcpc_b200_commit *c, uint64_t **d_comm, uint64_t **d_coeffs, uint8_t **d_hashes) {
  if (!c) return LCPC_B200_ERR_BAD_ARG;
  if (d_comm || d_coeffs) {
    lcpc_b200_ctx *ctx = c->enc->ctx;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (int rc = bind_device(ctx)) return rc;
    if (d_comm)
      if (int rc = ensure_comm(c)) return rc;
    if (d_coeffs)
      if (int rc = ensure_coeffs(c)) return rc;
  }
  if (d_comm) *d_comm = (uint64_t *)c->d_comm;
  if (d_coeffs) *d_coeffs = (uint64_t *)c->d_coeffs;
  if (d_hashes) *d_hashes = c->d_hashes;
  return LCPC_B200_OK;
}

// commit + download with the download overlapped: see HostOut
int lcpc_b200_commit_rerun_to_host(lcpc_b200_commit *c, const uint64_t *coeffs_in, size_t len, uint64_t *comm,
                                   uint64_t *coeffs, uint8_t *hashes) {
  if (!c || !coeffs_in) return LCPC_B200_ERR_BAD_ARG;
  lcpc_b200_ctx *ctx = c->enc->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (int rc = bind_device(ctx)) return rc;
  HostOut ho{comm, coeffs};
  if (int rc = commit_run(c, coeffs_in, len, cudaMemcpyHostToDevice, &ho)) return rc;
  if (hashes)
    CU(ctx, cudaMemcpyAsync(hashes, c->d_hashes, (2 * c->np2 - 1) * 32, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->side_stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return LCPC_B200_OK;
}

int lcpc_b200_commit_to_host(lcpc_b200_enc *enc, const uint64_t *coeffs_in, size_t len, uint64_t *comm, uint64_t *coeffs,
                             uint8_t *hashes) {
  if (!enc || (!coeffs_in && len)) return LCPC_B200_ERR_BAD_ARG;
  lcpc_b200_ctx *ctx = enc->ctx;
  lcpc_b200_commit *c = nullptr;
  {
    std::lock_guard<std::mutex> g(ctx->mu);
    if (int rc = bind_device(ctx)) return rc;
    if (int rc = commit_alloc(enc, len, &c)) return rc;
  }
  int rc = lcpc_b200_commit_rerun_to_host(c, coeffs_in, len, comm, coeffs, hashes);
  lcpc_b200_commit_free(c);
  return rc;
}

// ------------------------------------------------------------------------------------------- prove
int lcpc_b200_commit_collapse(lcpc_b200_commit *c, const uint64_t *tensor, uint64_t *poly) {
  if (!c || !tensor || !poly) return LCPC_B200_ERR_BAD_ARG;
  lcpc_b200_ctx *ctx = c->enc->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (int rc = bind_device(ctx)) return rc;
  const int field = c->enc->field;
  const size_t B = field_bytes(field);
  CU(ctx, cudaMemcpyAsync(c->d_tensor, tensor, c->n_rows * B, cudaMemcpyHostToDevice, ctx->stream));
  int nl = 0;
  cudaError_t ce = collapse_commit(c, &nl);
  ctx->launches += nl;
  if (ce != cudaSuccess) return cuda_fail(ctx, ce, "collapse");
  return return_poly(c, poly);
}

// d_poly -> caller memory through the page-locked landing buffer (a direct copy into pageable memory is staged by
// the driver at a fraction of the PCIe rate); synchronises the engine stream
static int return_poly(lcpc_b200_commit *c, uint64_t *poly) {
  lcpc_b200_ctx *ctx = c->enc->ctx;
  const size_t bytes = c->n_per_row * field_bytes(c->enc->field);
  if (!c->h_poly && cudaHostAlloc(&c->h_poly, bytes, cudaHostAllocDefault) != cudaSuccess) {
    c->h_poly = nullptr;
    cudaGetLastError();
    CU(ctx, cudaMemcpyAsync(poly, c->d_poly, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return LCPC_B200_OK;
  }
  CU(ctx, cudaMemcpyAsync(c->h_poly, c->d_poly, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(poly, c->h_poly, bytes);
  return LCPC_B200_OK;
}

// challenge tensor on the device: key -> d_tensor (n_rows elements), enqueue only
static int expand_tensor_into(lcpc_b200_commit *c, const uint8_t key[32]) {
  lcpc_b200_ctx *ctx = c->enc->ctx;
  if (!c->d_key) CU(ctx, cudaMalloc(&c->d_key, 32));
  CU(ctx, cudaMemcpyAsync(c->d_key, key, 32, cudaMemcpyHostToDevice, ctx->stream));
  cudaError_t ce = launch_expand_tensor(c->enc->field, c->d_key, 0, c->n_rows, c->d_tensor, ctx->stream);
  ctx->launches += 1;
  if (ce != cudaSuccess) return cuda_fail(ctx, ce, "expand_tensor");
  return LCPC_B200_OK;
}

int lcpc_b200_commit_degree_test(lcpc_b200_commit *c, const uint8_t key[32], uint64_t *poly, uint64_t *tensor_out) {
  if (!c || !key || !poly) return LCPC_B200_ERR_BAD_ARG;
  lcpc_b200_ctx *ctx = c->enc->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (int rc = bind_device(ctx)) return rc;
  const int field = c->enc->field;
  const size_t B = field_bytes(field);
  if (int rc = expand_tensor_into(c, key)) return rc;
  int nl = 0;
  cudaError_t ce = collapse_commit(c, &nl);
  ctx->launches += nl;
  if (ce != cudaSuccess) return cuda_fail(ctx, ce, "collapse");
  if (tensor_out) CU(ctx, cudaMemcpyAsync(tensor_out, c->d_tensor, c->n_rows * B, cudaMemcpyDeviceToHost, ctx->stream));
  return return_poly(c, poly);
}

int lcpc_b200_expand_tensor(lcpc_b200_ctx *ctx, int field, const uint8_t key[32], size_t n, uint64_t *out) {
  if (!ctx || !key || (!out && n)) return LCPC_B200_ERR_BAD_ARG;
  if (field_limbs32(field) < 0) return fail(ctx, LCPC_B200_ERR_BAD_ARG, "unknown field %d", field);
  if (n == 0) return LCPC_B200_OK;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (int rc = bind_device(ctx)) return rc;
  const size_t bytes = n * field_bytes(field);
  if (int rc = ensure_scratch(ctx, 256 + bytes)) return rc;
  uint8_t *base = (uint8_t *)ctx->scratch;
  CU(ctx, cudaMemcpyAsync(base, key, 32, cudaMemcpyHostToDevice, ctx->stream));
  cudaError_t ce = launch_expand_tensor(field, (const uint32_t *)base, 0, n, (uint32_t *)(base + 256), ctx->stream);
  ctx->launches += 1;
  if (ce != cudaSuccess) return cuda_fail(ctx, ce, "expand_tensor");
  CU(ctx, cudaMemcpyAsync(out, base + 256, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return LCPC_B200_OK;
}

int lcpc_b200_collapse_dev(lcpc_b200_ctx *ctx, int field, const uint64_t *d_coeffs, size_t row_stride,
                           const uint64_t *d_tensor, uint64_t *d_poly, size_t n_rows, size_t n_per_row) {
  if (!ctx) return LCPC_B200_ERR_BAD_ARG;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (field_limbs32(field) < 0) return fail(ctx, LCPC_B200_ERR_BAD_ARG, "unknown field %d", field);
  if (int rc = bind_device(ctx)) return rc;
  int nl = 0;
  cudaError_t ce = launch_collapse(field, (const uint32_t *)d_coeffs, row_stride, (const uint32_t *)d_tensor,
                                   (uint32_t *)d_poly, n_rows, n_per_row, nullptr, ctx->stream, &nl);
  ctx->launches += nl;
  if (ce != cudaSuccess) return cuda_fail(ctx, ce, "collapse");
  return LCPC_B200_OK;
}

int lcpc_b200_collapse(lcpc_b200_ctx *ctx, int field, const uint64_t *coeffs, const uint64_t *tensor, uint64_t *poly,
                       size_t n_rows, size_t n_per_row) {
  if (!ctx || !coeffs || !tensor || !poly) return LCPC_B200_ERR_BAD_ARG;
  if (field_limbs32(field) < 0) return fail(ctx, LCPC_B200_ERR_BAD_ARG, "unknown field %d", field);
  const size_t B = field_bytes(field);
  {
    std::lock_guard<std::mutex> g(ctx->mu);
    if (int rc = bind_device(ctx)) return rc;
    size_t cb = n_rows * n_per_row * B, tb = n_rows * B, pb = n_per_row * B;
    size_t tb_al = (tb + 255) & ~(size_t)255, cb_al = (cb + 255) & ~(size_t)255;
    if (int rc = ensure_scratch(ctx, cb_al + tb_al + pb)) return rc;
    uint8_t *base = (uint8_t *)ctx->scratch;
    uint32_t *d_c = (uint32_t *)base, *d_t = (uint32_t *)(base + cb_al), *d_p = (uint32_t *)(base + cb_al + tb_al);
    CU(ctx, cudaMemcpyAsync(d_c, coeffs, cb, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(d_t, tensor, tb, cudaMemcpyHostToDevice, ctx->stream));
    int nl = 0;
    cudaError_t ce = launch_collapse(field, d_c, n_per_row, d_t, d_p, n_rows, n_per_row, nullptr, ctx->stream, &nl);
    ctx->launches += nl;
    if (ce != cudaSuccess) return cuda_fail(ctx, ce, "collapse");
    CU(ctx, cudaMemcpyAsync(poly, d_p, pb, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return LCPC_B200_OK;
}

// open_column for n columns; the context's mutex is held by the caller
static int open_columns_locked(lcpc_b200_commit *c, const uint64_t *cols, size_t n, uint64_t *cols_out,
                               uint8_t *paths_out) {
  lcpc_b200_ctx *ctx = c->enc->ctx;
  for (size_t i = 0; i < n; i++)
    if (cols[i] >= c->n_cols) return fail(ctx, LCPC_B200_ERR_COLUMN, "column %llu >= n_cols %zu", (unsigned long long)cols[i], c->n_cols);
  if (n == 0) return LCPC_B200_OK;
  if (int rc = bind_device(ctx)) return rc;
  const int field = c->enc->field;
  const size_t B = field_bytes(field);
  const unsigned path_len = log2_ceil(c->n_cols);
  // device staging: [cols (8n) | column values (n*n_rows*B) | paths (n*path_len*32)]
  size_t off_vals = (n * 8 + 255) & ~(size_t)255;
  size_t off_paths = (off_vals + n * c->n_rows * B + 255) & ~(size_t)255;
  size_t total = off_paths + n * (size_t)path_len * 32;
  if (int rc = ensure_scratch(ctx, total)) return rc;
  uint8_t *base = (uint8_t *)ctx->scratch;
  CU(ctx, cudaMemcpyAsync(base, cols, n * 8, cudaMemcpyHostToDevice, ctx->stream));
  // open_column reads whole columns: contiguous in the work buffer when the codewords were left there
  cudaError_t ce = c->comm_in_w ? launch_gather_columns(field, (const uint32_t *)c->d_enc_scratch, c->n_rows, 1, (const uint64_t *)base, n,
                                                        (uint32_t *)(base + off_vals), ctx->stream, c->n_rows)
                                : launch_gather_columns(field, c->d_comm, c->n_rows, c->n_cols, (const uint64_t *)base, n,
                                                        (uint32_t *)(base + off_vals), ctx->stream);
  if (ce == cudaSuccess && path_len)
    ce = launch_gather_paths(c->d_hashes, c->np2, (const uint64_t *)base, n, path_len, base + off_paths, ctx->stream);
  ctx->launches += path_len ? 2 : 1;
  if (ce != cudaSuccess) return cuda_fail(ctx, ce, "open_columns");
  CU(ctx, cudaMemcpyAsync(cols_out, base + off_vals, n * c->n_rows * B, cudaMemcpyDeviceToHost, ctx->stream));
  if (path_len) CU(ctx, cudaMemcpyAsync(paths_out, base + off_paths, n * (size_t)path_len * 32, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return LCPC_B200_OK;
}

int lcpc_b200_commit_open_columns(lcpc_b200_commit *c, const uint64_t *cols, size_t n, uint64_t *cols_out,
                                  uint8_t *paths_out) {
  if (!c || (n && (!cols || !cols_out || !paths_out))) return LCPC_B200_ERR_BAD_ARG;
  std::lock_guard<std::mutex> g(c->enc->ctx->mu);
  return open_columns_locked(c, cols, n, cols_out, paths_out);
}


// ------------------------------------------------------------------------------- prove() / verify()
// def_labels! (lcpc-2d/src/macros.rs:28-36) leaves `$l` unsubstituted inside the byte-string literals
static const uint8_t kLabelDT[] = "$l//DT", kLabelPR[] = "$l//PR", kLabelPE[] = "$l//PE", kLabelCO[] = "$l//CO";
Labels resolve_labels(const lcpc_b200_labels *in) {
  if (!in) return Labels{kLabelDT, kLabelPR, kLabelPE, kLabelCO, 6, 6, 6, 6};
  return Labels{in->dt, in->pr, in->pe, in->co, in->dt_len, in->pr_len, in->pe_len, in->co_len};
}

// layout of one call's device buffers inside the context's grow-only scratch: 256-byte aligned regions
struct DeviceBlock {
  size_t used = 0;
  size_t reserve(size_t bytes) {  // returns the offset of the region
    size_t off = (used + 255) & ~(size_t)255;
    used = off + bytes;
    return off;
  }
};

// ChaCha20Rng::from_seed(key) -> n x Uniform::new(0usize, n_cols) (lcpc-2d/src/lib.rs:1073-1080, :903-911)
void sample_columns(const uint8_t key[32], size_t n_cols, size_t n, uint64_t *out) {
  host::ChaCha20Stream rng = host::ChaCha20Stream::from_seed(key);
  for (size_t i = 0; i < n; i++) out[i] = rng.below(n_cols);
}

// index of the first element whose limbs are not < p, or (size_t)-1
static size_t first_noncanonical(int field, const uint64_t *v, size_t n, size_t L) {
  uint64_t p[4] = {0, 0, 0, 0};
  {
    uint32_t p32[8] = {0};
    const int N = field_limbs32(field);
    for (int i = 0; i < N; i++) {
      switch (field) {
        case FT63: p32[i] = FieldP<FT63>::P(i); break;
        case FT127: p32[i] = FieldP<FT127>::P(i); break;
        case FT191: p32[i] = FieldP<FT191>::P(i); break;
        default: p32[i] = FieldP<FT255>::P(i); break;
      }
    }
    for (size_t l = 0; l < L; l++) p[l] = (uint64_t)p32[2 * l] | ((uint64_t)p32[2 * l + 1] << 32);
  }
  for (size_t i = 0; i < n; i++) {
    const uint64_t *e = v + i * L;
    bool less = false;
    for (size_t l = L; l-- > 0;) {
      if (e[l] != p[l]) {
        less = e[l] < p[l];
        break;
      }
    }
    if (!less) return i;
  }
  return (size_t)-1;
}

int lcpc_b200_sample_columns(const uint8_t key[32], size_t n_cols, size_t n, uint64_t *out) {
  if (!key || n_cols == 0 || (!out && n)) return LCPC_B200_ERR_BAD_ARG;
  sample_columns(key, n_cols, n, out);
  return LCPC_B200_OK;
}

int lcpc_b200_commit_prove(lcpc_b200_commit *c, lcpc_b200_transcript *tr, const lcpc_b200_labels *labels,
                           const uint64_t *outer_tensor, size_t outer_len, size_t n_degree_tests, size_t n_col_opens,
                           uint64_t *p_eval, uint64_t *p_random, uint64_t *col_idx, uint64_t *cols_out,
                           uint8_t *paths_out) {
  if (!c || !tr || !outer_tensor || !p_eval || (n_degree_tests && !p_random) || (n_col_opens && (!cols_out || !paths_out)))
    return LCPC_B200_ERR_BAD_ARG;
  lcpc_b200_ctx *ctx = c->enc->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (outer_len != c->n_rows)  // ProverError::OuterTensor (:1016-1018)
    return fail(ctx, LCPC_B200_ERR_OUTER_TENSOR, "outer tensor has %zu entries, the commitment %zu rows", outer_len, c->n_rows);
  if (int rc = bind_device(ctx)) return rc;
  const Labels lb = resolve_labels(labels);
  const int field = c->enc->field;
  const size_t B = field_bytes(field), L = B / 8, pbytes = c->n_per_row * B;
  if (!c->d_repr) CU(ctx, cudaMalloc(&c->d_repr, pbytes));
  uint32_t *d_repr = c->d_repr;
  // results land in page-locked staging [poly | repr] (a pageable destination would be staged by the driver at a
  // fraction of the PCIe rate); pageable fallback if the allocation fails
  std::vector<uint8_t> pageable;
  uint8_t *stage = (uint8_t *)host_stage(ctx, 2 * pbytes);
  if (!stage) {
    pageable.resize(2 * pbytes);
    stage = pageable.data();
  }
  int rc = LCPC_B200_OK;
  // one collapse against the tensor in c->d_tensor: Montgomery limbs to `dst`, canonical bytes into the transcript
  auto collapse_and_absorb = [&](uint64_t *dst, const uint8_t *label, size_t label_len) -> int {
    int nl = 0;
    cudaError_t ce = collapse_commit(c, &nl);
    if (ce == cudaSuccess) ce = launch_field_op(field, 4, d_repr, c->d_poly, nullptr, c->n_per_row, ctx->stream);
    ctx->launches += nl + 1;
    if (ce != cudaSuccess) return cuda_fail(ctx, ce, "prove: collapse");
    CU(ctx, cudaMemcpyAsync(stage + pbytes, d_repr, pbytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaEventRecord(ctx->begin_ev, ctx->stream));
    CU(ctx, cudaMemcpyAsync(stage, c->d_poly, pbytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaEventSynchronize(ctx->begin_ev));  // the canonical bytes are in: absorb them while the limbs follow
    tr->tr.append_elems(label, label_len, stage + pbytes, B, c->n_per_row);  // transcript_update per coefficient (:1043-1045)
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(dst, stage, pbytes);
    return LCPC_B200_OK;
  };
  for (size_t i = 0; i < n_degree_tests && rc == LCPC_B200_OK; i++) {  // :1025-1048
    uint8_t key[32];
    tr->tr.challenge_bytes(lb.dt, lb.dt_len, key, 32);
    rc = expand_tensor_into(c, key);
    if (rc == LCPC_B200_OK) rc = collapse_and_absorb(p_random + i * c->n_per_row * L, lb.pr, lb.pr_len);
  }
  if (rc == LCPC_B200_OK) {  // :1051-1063
    cudaError_t ce = cudaMemcpyAsync(c->d_tensor, outer_tensor, c->n_rows * B, cudaMemcpyHostToDevice, ctx->stream);
    if (ce != cudaSuccess) rc = cuda_fail(ctx, ce, "prove: outer tensor");
    else rc = collapse_and_absorb(p_eval, lb.pe, lb.pe_len);
  }
  if (rc != LCPC_B200_OK) return rc;
  // :1066-1085
  uint8_t key[32];
  tr->tr.challenge_bytes(lb.co, lb.co_len, key, 32);
  std::vector<uint64_t> cols(n_col_opens);
  sample_columns(key, c->n_cols, n_col_opens, cols.data());
  if (col_idx) memcpy(col_idx, cols.data(), n_col_opens * 8);
  return open_columns_locked(c, cols.data(), n_col_opens, cols_out, paths_out);
}

int lcpc_b200_verify(lcpc_b200_enc *enc, lcpc_b200_transcript *tr, const lcpc_b200_labels *labels,
                     const uint8_t root[32], const uint64_t *outer_tensor, size_t outer_len,
                     const uint64_t *inner_tensor, size_t inner_len, size_t n_col_opens, size_t n_degree_tests,
                     const lcpc_b200_proof *proof, uint64_t *eval_out) {
  if (!enc || !tr || !root || !proof || !eval_out) return LCPC_B200_ERR_BAD_ARG;
  lcpc_b200_ctx *ctx = enc->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  // argument checks in the reference's order (:845-859)
  if (n_col_opens != proof->n_columns || n_col_opens == 0)
    return fail(ctx, LCPC_B200_VERR_NUM_COL_OPENS, "proof opens %zu columns, the encoding asks for %zu", proof->n_columns, n_col_opens);
  const size_t n_rows = proof->n_rows, n_cols = proof->n_cols, n_per_row = proof->n_per_row, n_open = proof->n_columns;
  if (inner_len != n_per_row) return fail(ctx, LCPC_B200_VERR_INNER_TENSOR, "inner tensor: %zu != n_per_row %zu", inner_len, n_per_row);
  if (outer_len != n_rows) return fail(ctx, LCPC_B200_VERR_OUTER_TENSOR, "outer tensor: %zu != n_rows %zu", outer_len, n_rows);
  if (!lcpc_b200_enc_dims_ok(enc, n_per_row, n_cols)) return fail(ctx, LCPC_B200_VERR_ENCODING_DIMS, "encoding dimension mismatch");
  // the reference indexes p_random_vec[i] for i < n_degree_tests (:883) and panics when the proof is short
  if (proof->n_degree_tests != n_degree_tests || n_rows == 0 || !outer_tensor || !inner_tensor || !proof->p_eval ||
      (n_degree_tests && !proof->p_random) || !proof->cols || (proof->path_len && !proof->paths) || proof->path_len > 64)
    return fail(ctx, LCPC_B200_ERR_BAD_ARG, "verify: malformed proof");
  // The proof comes from the other side of a trust boundary: every field element in it must be a canonical residue
  // (limbs < p).  The device arithmetic assumes it (field.cuh: `2p < 2^(32N)`, no carry out of add/sub), and
  // from_mont maps v and v + p to the same transcript bytes, so a non-canonical element would be both a malleable
  // proof and an encode input outside the kernels' precondition.  (The reference's derived Deserialize does not
  // check either; rejecting is the stricter, safe reading.)
  const int field = enc->field;
  {
    const size_t L = field_bytes(field) / 8;
    const size_t bad_pr = first_noncanonical(field, proof->p_random, n_degree_tests * n_per_row, L);
    const size_t bad_pe = first_noncanonical(field, proof->p_eval, n_per_row, L);
    const size_t bad_col = first_noncanonical(field, proof->cols, n_open * n_rows, L);
    if (bad_pr != (size_t)-1 || bad_pe != (size_t)-1 || bad_col != (size_t)-1)
      return fail(ctx, LCPC_B200_ERR_BAD_ARG, "verify: non-canonical field element in the proof (%s[%zu] >= p)",
                  bad_pr != (size_t)-1 ? "p_random" : bad_pe != (size_t)-1 ? "p_eval" : "columns",
                  bad_pr != (size_t)-1 ? bad_pr : bad_pe != (size_t)-1 ? bad_pe : bad_col);
  }
  // a Merkle path has exactly log2(n_cols.next_power_of_two()) siblings (open_column, :811-821); any other length
  // cannot lead from a leaf to the root of the committed tree
  if (proof->path_len != log2_ceil(n_cols))
    return fail(ctx, LCPC_B200_VERR_COLUMN_PATH, "merkle path length %zu, expected %u", (size_t)proof->path_len, log2_ceil(n_cols));
  if (int rc = bind_device(ctx)) return rc;
  const Labels lb = resolve_labels(labels);
  const size_t B = field_bytes(field), T = n_degree_tests + 1;
  const unsigned path_len = (unsigned)proof->path_len;

  DeviceBlock blk;
  const size_t o_in = blk.reserve(T * n_per_row * B), o_repr = blk.reserve(T * n_per_row * B);
  const size_t o_rows = blk.reserve(T * n_cols * B), o_tensors = blk.reserve(T * n_rows * B);
  const size_t o_keys = blk.reserve(T * 32), o_cols = blk.reserve(n_open * n_rows * B);
  const size_t o_colsT = blk.reserve(n_open * n_rows * B), o_idx = blk.reserve(n_open * 8);
  const size_t o_leaves = blk.reserve(n_open * 32), o_paths = blk.reserve(n_open * (size_t)path_len * 32 + 32);
  const size_t o_root = blk.reserve(32), o_evals = blk.reserve(T * n_open * B), o_flags = blk.reserve(n_open * 4);
  const size_t o_hs = blk.reserve(hash_scratch_bytes(field, n_rows, n_open) + 32);
  const size_t o_es = blk.reserve(enc_scratch_bytes(enc, T) + 32);
  const size_t o_inner = blk.reserve(n_per_row * B), o_part = blk.reserve((DOT_PARTIALS + 1) * B);
  if (int rc = ensure_scratch(ctx, blk.used)) return rc;
  uint8_t *d = (uint8_t *)ctx->scratch;
  auto w32 = [&](size_t off) { return (uint32_t *)(d + off); };
  cudaStream_t st = ctx->stream;

  // p_random rows and p_eval: to the device once; canonical bytes back for the transcript, rows on to the encode
  if (n_degree_tests)
    CU(ctx, cudaMemcpyAsync(d + o_in, proof->p_random, n_degree_tests * n_per_row * B, cudaMemcpyHostToDevice, st));
  CU(ctx, cudaMemcpyAsync(d + o_in + n_degree_tests * n_per_row * B, proof->p_eval, n_per_row * B, cudaMemcpyHostToDevice, st));
  cudaError_t ce = launch_field_op(field, 4, w32(o_repr), w32(o_in), nullptr, T * n_per_row, st);
  ctx->launches += 1;
  if (ce != cudaSuccess) return cuda_fail(ctx, ce, "verify: from_mont");
  const size_t repr_bytes = T * n_per_row * B;
  std::vector<uint8_t> pageable;
  uint8_t *repr = (uint8_t *)host_stage(ctx, repr_bytes);
  if (!repr) {
    pageable.resize(repr_bytes);
    repr = pageable.data();
  }
  CU(ctx, cudaMemcpyAsync(repr, d + o_repr, repr_bytes, cudaMemcpyDeviceToHost, st));
  CU(ctx, cudaEventRecord(ctx->begin_ev, st));
  // step 1b / step 2 (:883-888, :914-921): every row zero-extended to n_cols and encoded, in one batch
  if (int rc = encode_rows(enc, w32(o_in), n_per_row, n_per_row, w32(o_rows), T, d + o_es)) return rc;
  CU(ctx, cudaEventSynchronize(ctx->begin_ev));

  // host: replay the transcript while the device encodes (:866-911)
  std::vector<uint8_t> keys(T * 32);
  for (size_t i = 0; i < n_degree_tests; i++) {
    tr->tr.challenge_bytes(lb.dt, lb.dt_len, keys.data() + 32 * i, 32);
    tr->tr.append_elems(lb.pr, lb.pr_len, repr + i * n_per_row * B, B, n_per_row);
  }
  tr->tr.append_elems(lb.pe, lb.pe_len, repr + n_degree_tests * n_per_row * B, B, n_per_row);
  uint8_t key_co[32];
  tr->tr.challenge_bytes(lb.co, lb.co_len, key_co, 32);
  std::vector<uint64_t> cols(n_open);
  sample_columns(key_co, n_cols, n_open, cols.data());

  // tensors: the degree-test ones expanded from their keys on the device, then the outer tensor
  if (n_degree_tests) CU(ctx, cudaMemcpyAsync(d + o_keys, keys.data(), n_degree_tests * 32, cudaMemcpyHostToDevice, st));
  for (size_t i = 0; i < n_degree_tests; i++) {
    ce = launch_expand_tensor(field, w32(o_keys + 32 * i), 0, n_rows, w32(o_tensors + i * n_rows * B), st);
    ctx->launches += 1;
    if (ce != cudaSuccess) return cuda_fail(ctx, ce, "verify: expand_tensor");
  }
  CU(ctx, cudaMemcpyAsync(d + o_tensors + n_degree_tests * n_rows * B, outer_tensor, n_rows * B, cudaMemcpyHostToDevice, st));
  CU(ctx, cudaMemcpyAsync(d + o_cols, proof->cols, n_open * n_rows * B, cudaMemcpyHostToDevice, st));
  CU(ctx, cudaMemcpyAsync(d + o_idx, cols.data(), n_open * 8, cudaMemcpyHostToDevice, st));
  if (path_len) CU(ctx, cudaMemcpyAsync(d + o_paths, proof->paths, n_open * (size_t)path_len * 32, cudaMemcpyHostToDevice, st));
  CU(ctx, cudaMemcpyAsync(d + o_root, root, 32, cudaMemcpyHostToDevice, st));
  CU(ctx, cudaMemcpyAsync(d + o_inner, inner_tensor, n_per_row * B, cudaMemcpyHostToDevice, st));

  // step 3 (:926-942), batched: columns -> row-major [n_rows][n_open]; leaf digests with the commit's kernels;
  // <tensor_k, column_j> for all j with collapse_kernel; comparisons and Merkle walks in check_columns_kernel
  ce = launch_transpose_columns(field, w32(o_cols), w32(o_colsT), n_open, n_rows, st);
  ctx->launches += 1;
  int nl = 0;
  if (ce == cudaSuccess) ce = launch_hash_columns(field, w32(o_colsT), n_rows, n_open, n_open, d + o_leaves, d + o_hs, st, &nl);
  ctx->launches += nl;
  for (size_t k = 0; k < T && ce == cudaSuccess; k++) {
    ce = launch_collapse(field, w32(o_colsT), n_open, w32(o_tensors + k * n_rows * B), w32(o_evals + k * n_open * B), n_rows,
                         n_open, nullptr, st, &nl);
    ctx->launches += nl;
  }
  if (ce == cudaSuccess)
    ce = launch_check_columns(field, w32(o_evals), w32(o_rows), n_cols, (unsigned)T, (const uint64_t *)(d + o_idx), n_open,
                              d + o_leaves, d + o_paths, path_len, d + o_root, w32(o_flags), st);
  ctx->launches += 1;
  // step 4 (:944-951)
  if (ce == cudaSuccess)
    ce = launch_dot(field, w32(o_inner), w32(o_in + n_degree_tests * n_per_row * B), n_per_row, w32(o_part + B), w32(o_part), st, &nl);
  ctx->launches += nl;
  if (ce != cudaSuccess) return cuda_fail(ctx, ce, "verify: column checks");
  std::vector<uint32_t> flags(n_open);
  CU(ctx, cudaMemcpyAsync(flags.data(), d + o_flags, n_open * 4, cudaMemcpyDeviceToHost, st));
  CU(ctx, cudaMemcpyAsync(eval_out, d + o_part, B, cudaMemcpyDeviceToHost, st));
  CU(ctx, cudaStreamSynchronize(st));
  for (size_t j = 0; j < n_open; j++) {  // the match at :937-942
    if (!(flags[j] & 1u)) return fail(ctx, LCPC_B200_VERR_COLUMN_DEGREE, "column %zu (#%llu): degree test dot product failed", j, (unsigned long long)cols[j]);
    if (!(flags[j] & 2u)) return fail(ctx, LCPC_B200_VERR_COLUMN_EVAL, "column %zu (#%llu): eval dot product failed", j, (unsigned long long)cols[j]);
    if (!(flags[j] & 4u)) return fail(ctx, LCPC_B200_VERR_COLUMN_PATH, "column %zu (#%llu): merkle path failed", j, (unsigned long long)cols[j]);
  }
  return LCPC_B200_OK;
}

// --------------------------------------------------------------------------------- standalone pieces
int lcpc_b200_hash_columns_dev(lcpc_b200_ctx *ctx, int field, const uint64_t *d_comm, size_t n_rows, size_t n_cols,
                               size_t row_stride, uint8_t *d_leaves) {
  if (!ctx) return LCPC_B200_ERR_BAD_ARG;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (field_limbs32(field) < 0) return fail(ctx, LCPC_B200_ERR_BAD_ARG, "unknown field %d", field);
  if (int rc = bind_device(ctx)) return rc;
  if (int rc = ensure_scratch(ctx, hash_scratch_bytes(field, n_rows, n_cols))) return rc;
  int nl = 0;
  cudaError_t ce = launch_hash_columns(field, (const uint32_t *)d_comm, n_rows, n_cols, row_stride, d_leaves,
                                       ctx->scratch, ctx->stream, &nl);
  ctx->launches += nl;
  if (ce != cudaSuccess) return cuda_fail(ctx, ce, "hash_columns");
  return LCPC_B200_OK;
}

int lcpc_b200_merkle_tree_dev(lcpc_b200_ctx *ctx, uint8_t *d_hashes, size_t np2) {
  if (!ctx || !d_hashes || !is_pow2(np2)) return LCPC_B200_ERR_BAD_ARG;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (int rc = bind_device(ctx)) return rc;
  int nl = 0;
  cudaError_t ce = launch_merkle_tree(d_hashes, np2, ctx->stream, &nl);
  ctx->launches += nl;
  if (ce != cudaSuccess) return cuda_fail(ctx, ce, "merkle_tree");
  return LCPC_B200_OK;
}

int lcpc_b200_merkle_layers_dev(lcpc_b200_ctx *ctx, uint8_t *d_hashes, size_t n_leaves, unsigned n_layers) {
  if (!ctx || !d_hashes) return LCPC_B200_ERR_BAD_ARG;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (n_layers >= 48 || (n_leaves & (((size_t)1 << n_layers) - 1)))
    return fail(ctx, LCPC_B200_ERR_BAD_ARG, "merkle_layers: %zu nodes are not a multiple of 2^%u", n_leaves, n_layers);
  if (int rc = bind_device(ctx)) return rc;
  int nl = 0;
  cudaError_t ce = launch_merkle_layers(d_hashes, n_leaves, n_layers, ctx->stream, &nl);
  ctx->launches += nl;
  if (ce != cudaSuccess) return cuda_fail(ctx, ce, "merkle_layers");
  return LCPC_B200_OK;
}

int lcpc_b200_pack_column_blocks_dev(lcpc_b200_ctx *ctx, int field, const uint64_t *d_rows, size_t n_rows, size_t n_cols,
                                     size_t n_blocks, const uint64_t *d_starts, uint64_t *d_out) {
  if (!ctx || !d_rows || !d_starts || !d_out) return LCPC_B200_ERR_BAD_ARG;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (field_limbs32(field) < 0 || n_blocks == 0 || n_blocks > 64)
    return fail(ctx, LCPC_B200_ERR_BAD_ARG, "pack: bad field or block count");
  if (int rc = bind_device(ctx)) return rc;
  cudaError_t ce = launch_pack_column_blocks(field, (const uint32_t *)d_rows, n_rows, n_cols, (unsigned)n_blocks, d_starts,
                                             (uint32_t *)d_out, ctx->stream);
  ctx->launches += 1;
  if (ce != cudaSuccess) return cuda_fail(ctx, ce, "pack_column_blocks");
  return LCPC_B200_OK;
}

int lcpc_b200_merkleize(lcpc_b200_ctx *ctx, int field, const uint64_t *comm, size_t n_rows, size_t n_cols,
                        uint8_t *hashes) {
  if (!ctx || !comm || !hashes || n_cols == 0) return LCPC_B200_ERR_BAD_ARG;
  if (field_limbs32(field) < 0) return fail(ctx, LCPC_B200_ERR_BAD_ARG, "unknown field %d", field);
  std::lock_guard<std::mutex> g(ctx->mu);
  if (int rc = bind_device(ctx)) return rc;
  const size_t B = field_bytes(field);
  const size_t np2 = (size_t)1 << log2_ceil(n_cols);
  const size_t cb = (n_rows * n_cols * B + 255) & ~(size_t)255, hb = ((2 * np2 - 1) * 32 + 255) & ~(size_t)255;
  const size_t sb = hash_scratch_bytes(field, n_rows, n_cols);
  if (int rc = ensure_scratch(ctx, cb + hb + sb)) return rc;
  uint8_t *base = (uint8_t *)ctx->scratch;
  uint8_t *d_h = base + cb;
  CU(ctx, cudaMemcpyAsync(base, comm, n_rows * n_cols * B, cudaMemcpyHostToDevice, ctx->stream));
  if (np2 > n_cols) CU(ctx, cudaMemsetAsync(d_h + n_cols * 32, 0, (np2 - n_cols) * 32, ctx->stream));
  int nl = 0;
  cudaError_t ce = launch_hash_columns(field, (const uint32_t *)base, n_rows, n_cols, n_cols, d_h, base + cb + hb, ctx->stream, &nl);
  ctx->launches += nl;
  if (ce == cudaSuccess) {
    ce = launch_merkle_tree(d_h, np2, ctx->stream, &nl);
    ctx->launches += nl;
  }
  if (ce != cudaSuccess) return cuda_fail(ctx, ce, "merkleize");
  CU(ctx, cudaMemcpyAsync(hashes, d_h, (2 * np2 - 1) * 32, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return LCPC_B200_OK;
}

int lcpc_b200_field_op(lcpc_b200_ctx *ctx, int field, int op, uint64_t *r, const uint64_t *a, const uint64_t *b, size_t n) {
  if (!ctx || !r || !a) return LCPC_B200_ERR_BAD_ARG;
  if (field_limbs32(field) < 0 || !(op == 0 || op == 1 || op == 2 || op == 4 || op == 5 || op == 6 || op == 7))
    return fail(ctx, LCPC_B200_ERR_BAD_ARG, "bad field/op %d/%d", field, op);
  const bool two = op != 4;
  if (two && !b) return LCPC_B200_ERR_BAD_ARG;
  if (n == 0) return LCPC_B200_OK;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (int rc = bind_device(ctx)) return rc;
  const size_t bytes = n * field_bytes(field), al = (bytes + 255) & ~(size_t)255;
  if (int rc = ensure_scratch(ctx, 3 * al)) return rc;
  uint8_t *base = (uint8_t *)ctx->scratch;
  CU(ctx, cudaMemcpyAsync(base, a, bytes, cudaMemcpyHostToDevice, ctx->stream));
  if (two) CU(ctx, cudaMemcpyAsync(base + al, b, bytes, cudaMemcpyHostToDevice, ctx->stream));
  cudaError_t ce = launch_field_op(field, op, (uint32_t *)(base + 2 * al), (const uint32_t *)base,
                                   two ? (const uint32_t *)(base + al) : nullptr, n, ctx->stream);
  ctx->launches += 1;
  if (ce != cudaSuccess) return cuda_fail(ctx, ce, "field_op");
  CU(ctx, cudaMemcpyAsync(r, base + 2 * al, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return LCPC_B200_OK;
}

}  // extern "C"
