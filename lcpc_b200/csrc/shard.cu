// lcpc_b200/csrc/shard.cu -- LcCommit::commit / prove sharded over the GPUs of one box, behind the C ABI.
//
// The reference is one process whose two data-parallel loops are row-parallel (`enc.encode` per row,
// lcpc-2d/src/lib.rs:648-653) and column-parallel (`hash_columns`, :706-745); between them sits a transpose.
// Here GPU g encodes the row block [row_lo[g], row_lo[g+1]) and owns the column block [col_lo[g], col_lo[g+1])
// (a union of aligned Merkle subtrees, balanced over the real columns).  The transpose is FUSED INTO THE ENCODE:
// the last transform pass (or the expander's final transpose) stores column block h of every encoded row straight
// into GPU h's receive matrix through peer-mapped memory -- plain st.global over NVLink/NVSwitch, tile by tile as
// the transform finishes -- so there is no pack pass, no send buffer and no collective call on the data path.
//
// Every GPU exports one WINDOW (a single cudaMalloc allocation: flags, two receive matrices, the subtree roots of
// all ranks, the prover's exchange areas).  Peers map it either directly (same process: cudaDeviceEnablePeerAccess,
// the shape of a Rust host, which is one process like the reference) or through a CUDA IPC handle (one process per
// GPU: the shape bench.py is launched in).  Ordering between GPUs uses flags in those windows: after its stores a
// rank writes its epoch number into every peer's flag slot (st.release.sys behind __threadfence_system), and the
// consumer's stream spins on its own slots (ld.acquire.sys) in a one-warp kernel -- no host round trip, no NCCL.
// All exchange areas are double-buffered by epoch parity: a rank can only run two epochs ahead of a peer after the
// peer has signalled the epoch in between, which that peer does after consuming the older buffer (stream order).
#include <algorithm>
#include <cstring>
#include <new>
#include <vector>

#include "api_internal.h"
#include "field.cuh"

using namespace lcpc;

namespace {

constexpr unsigned MAX_WORLD = MAX_SCATTER;
constexpr unsigned N_CHANNELS = 4;  // A: encoded tiles landed, B: subtree roots landed, C: collapse partials, D: openings
enum { CH_TILES = 0, CH_ROOTS = 1, CH_PARTS = 2, CH_OPEN = 3 };

struct PeerPtrs {
  uint8_t *p[MAX_WORLD];
};

struct ShardPlan {
  unsigned world = 0;
  size_t n_rows = 0, n_per_row = 0, n_cols = 0, np2 = 0;
  size_t sub_leaves = 0, n_sub = 0, n_real_sub = 0;  // T leaves per aligned subtree, S = np2 / T subtrees
  size_t row_lo[MAX_WORLD + 1] = {}, sub_lo[MAX_WORLD + 1] = {}, col_lo[MAX_WORLD + 1] = {};
};

size_t next_pow2(size_t v) {
  size_t p = 1;
  while (p < v) p <<= 1;
  return p;
}

void make_plan(size_t n_rows, size_t n_per_row, size_t n_cols, unsigned world, ShardPlan *p) {
  constexpr size_t SUB_PER_RANK = 8;
  p->world = world, p->n_rows = n_rows, p->n_per_row = n_per_row, p->n_cols = n_cols;
  p->np2 = n_cols > 1 ? next_pow2(n_cols) : 1;
  p->n_sub = std::min(p->np2, next_pow2((size_t)world * SUB_PER_RANK));
  p->sub_leaves = p->np2 / p->n_sub;
  p->n_real_sub = (n_cols + p->sub_leaves - 1) / p->sub_leaves;
  for (unsigned g = 0; g <= world; g++) {
    p->row_lo[g] = (size_t)g * n_rows / world;
    p->sub_lo[g] = (size_t)g * p->n_real_sub / world;
    p->col_lo[g] = std::min(p->sub_lo[g] * p->sub_leaves, n_cols);
  }
  p->col_lo[world] = n_cols;
}

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// byte offsets inside a window; identical on every rank (same plan, same arguments)
struct WindowLayout {
  size_t flags = 0, status = 0, recv[2] = {}, top[2] = {}, parts[2] = {}, open_vals[2] = {}, open_paths[2] = {}, total = 0;
};

WindowLayout make_layout(const ShardPlan &p, size_t B, size_t max_open) {
  WindowLayout w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align256(off + bytes);
    return o;
  };
  w.flags = take(N_CHANNELS * MAX_WORLD * 4);
  w.status = take(256);
  size_t max_cols = 0;
  for (unsigned h = 0; h < p.world; h++) max_cols = std::max(max_cols, p.col_lo[h + 1] - p.col_lo[h]);
  const unsigned path_len = [&] {
    unsigned l = 0;
    while (((size_t)1 << l) < p.np2) l++;
    return l;
  }();
  for (int b = 0; b < 2; b++) w.recv[b] = take(p.n_rows * max_cols * B);
  for (int b = 0; b < 2; b++) w.top[b] = take(p.n_sub * 32);
  for (int b = 0; b < 2; b++) w.parts[b] = take((size_t)p.world * p.n_per_row * B);
  for (int b = 0; b < 2; b++) w.open_vals[b] = take(max_open * p.n_rows * B);
  for (int b = 0; b < 2; b++) w.open_paths[b] = take(max_open * (size_t)path_len * 32);
  w.total = off;
  return w;
}

// ---- device side of the flags -----------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// "everything this stream did before is visible; epoch `e` of channel `ch` has happened on rank `me`"
__global__ void shard_signal_kernel(PeerPtrs peers, size_t flag_off, unsigned world, unsigned me, uint32_t epoch) {
  if (threadIdx.x < world) {
    __threadfence_system();
    uint32_t *f = reinterpret_cast<uint32_t *>(peers.p[threadIdx.x] + flag_off) + me;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(epoch) : "memory");
  }
}

// spin until every rank's slot of this channel shows an epoch >= `epoch`; a peer that never arrives sets a bit in
// *status after timeout_ns instead of hanging the device
__global__ void shard_wait_kernel(const uint32_t *flags, unsigned world, uint32_t epoch, uint32_t *status,
                                  unsigned long long timeout_ns) {
  if (threadIdx.x < world) {
    const unsigned long long t0 = global_timer_ns();
    for (;;) {
      uint32_t v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + threadIdx.x) : "memory");
      if ((int32_t)(v - epoch) >= 0) break;
      if (global_timer_ns() - t0 > timeout_ns) {
        atomicOr(status, 1u << threadIdx.x);
        break;
      }
      __nanosleep(200);
    }
  }
}

// copy n16 16-byte granules from src to the same offset of every peer's window (many CTAs; signal separately)
__global__ void __launch_bounds__(256)
shard_push_kernel(const uint4 *__restrict__ src, size_t n16, PeerPtrs peers, size_t dst_off, unsigned world) {
  const size_t total = n16 * world;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t h = idx / n16, i = idx % n16;
    reinterpret_cast<uint4 *>(peers.p[h] + dst_off)[i] = src[i];
  }
}

// small payloads (subtree roots): one CTA copies to every peer and signals in the same launch
__global__ void __launch_bounds__(256)
shard_push_signal_kernel(const uint4 *__restrict__ src, size_t n16, PeerPtrs peers, size_t dst_off, size_t flag_off,
                         unsigned world, unsigned me, uint32_t epoch) {
  const size_t total = n16 * world;
  for (size_t idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const size_t h = idx / n16, i = idx % n16;
    reinterpret_cast<uint4 *>(peers.p[h] + dst_off)[i] = src[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < world) {
    __threadfence_system();
    uint32_t *f = reinterpret_cast<uint32_t *>(peers.p[threadIdx.x] + flag_off) + me;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(epoch) : "memory");
  }
}

// open_column (lcpc-2d/src/lib.rs:788-825) for the requested columns this rank owns: values from its receive
// matrix, siblings from its subtree forest and the replicated top tree; every opening goes to slot i of every
// peer's window, so that after one signal/wait every rank holds all of them.
struct OpenArgs {
  const uint32_t *recv;   // [n_rows][my_cols] elements
  const uint8_t *forest;  // [leaves | layer 1 | .. | roots] of my_subs aligned subtrees
  const uint8_t *top;     // [n_sub | n_sub/2 | .. | 1]
  const uint64_t *cols;
  size_t n_open, n_rows, my_cols, c0, c1, forest_leaves, sub_leaves, n_sub;
  unsigned n_limbs, sub_layers, path_len, world;
  size_t vals_off, paths_off;
};

__global__ void __launch_bounds__(256) shard_open_kernel(OpenArgs a, PeerPtrs peers) {
  // values: 16-byte (or 8-byte for Ft63 / Ft191) granules
  const unsigned gran = (a.n_limbs % 4 == 0) ? 4 : 2, gpe = a.n_limbs / gran;
  const size_t per_col = a.n_rows * gpe;
  const size_t total_v = a.n_open * per_col;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total_v; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t i = idx / per_col, rem = idx % per_col, r = rem / gpe, part = rem % gpe;
    const size_t col = a.cols[i];
    if (col < a.c0 || col >= a.c1) continue;
    const uint32_t *src = a.recv + ((r * a.my_cols + (col - a.c0)) * a.n_limbs + part * gran);
    const size_t dst_word = (i * a.n_rows + r) * a.n_limbs + part * gran;
    if (gran == 4) {
      const uint4 v = *reinterpret_cast<const uint4 *>(src);
      for (unsigned h = 0; h < a.world; h++) *reinterpret_cast<uint4 *>(peers.p[h] + a.vals_off + dst_word * 4) = v;
    } else {
      const uint2 v = *reinterpret_cast<const uint2 *>(src);
      for (unsigned h = 0; h < a.world; h++) *reinterpret_cast<uint2 *>(peers.p[h] + a.vals_off + dst_word * 4) = v;
    }
  }
  // paths: 16-byte halves of 32-byte digests
  const size_t total_p = a.n_open * a.path_len * 2;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total_p; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t half = idx & 1, e = idx >> 1, l = e % a.path_len, i = e / a.path_len;
    const size_t col = a.cols[i];
    if (col < a.c0 || col >= a.c1) continue;
    const uint4 *src;
    if (l < a.sub_layers) {  // sibling inside this rank's aligned subtrees
      size_t off = 0, len = a.forest_leaves;
      for (size_t q = 0; q < l; q++) off += len, len >>= 1;
      src = reinterpret_cast<const uint4 *>(a.forest) + (off + (((col - a.c0) >> l) ^ 1)) * 2 + half;
    } else {  // sibling in the tree over the subtree roots
      const size_t t = l - a.sub_layers, sub = col / a.sub_leaves;
      size_t off = 0, len = a.n_sub;
      for (size_t q = 0; q < t; q++) off += len, len >>= 1;
      src = reinterpret_cast<const uint4 *>(a.top) + (off + ((sub >> t) ^ 1)) * 2 + half;
    }
    const uint4 v = *src;
    for (unsigned h = 0; h < a.world; h++) reinterpret_cast<uint4 *>(peers.p[h] + a.paths_off)[idx] = v;
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
struct lcpc_b200_shard {
  lcpc_b200_enc *enc = nullptr;
  ShardPlan p;
  WindowLayout wl;
  unsigned rank = 0, sub_layers = 0, path_len = 0;
  size_t B = 0, N = 0, max_open = 0;
  size_t my_rows = 0, my_cols = 0, my_subs = 0, my_elems = 0;  // my_elems: coefficients this rank's rows really hold
  uint8_t *window = nullptr;
  PeerPtrs peers = {};
  bool peer_ipc[MAX_WORLD] = {};
  bool connected = false;
  uint32_t *d_coeffs = nullptr, *d_tmp = nullptr;
  uint8_t *d_forest = nullptr, *d_top = nullptr;
  size_t forest_leaves = 0, forest_nodes = 0, roots_off = 0;
  void *d_hash_scratch = nullptr, *d_enc_scratch = nullptr;
  uint32_t *d_tensor = nullptr, *d_part = nullptr, *d_poly = nullptr, *d_repr = nullptr, *d_key = nullptr, *d_ones = nullptr;
  uint64_t *d_cols = nullptr;
  uint32_t epoch = 0, collapse_seq = 0, collapse_done = 0, open_seq = 0, open_done = 0;
  size_t open_n = 0;
  bool collapse_want_repr = false;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t coeffs_free = nullptr;  // recorded behind the last enqueued reader of d_coeffs (encode pass, collapse)
  bool coeffs_free_valid = false;
  // SHARD_PIPELINE: steps 2 and 3 of a commit (exchange wait, column hashing, roots, top tree) run on their own stream,
  // so that the NEXT commit's encode starts right behind this commit's -- the hash phase of a small per-GPU share is
  // a string of short, latency-bound launches that would otherwise leave the GPU mostly idle
  bool pipelined = false;
  cudaStream_t hash_stream = nullptr;
  cudaEvent_t enc_done = nullptr, hash_done = nullptr;
  bool hash_pending = false;
  unsigned long long timeout_ns = (unsigned long long)std::max<long>(1, tunable("SHARD_TIMEOUT_MS", 20000)) * 1000000ull;
};

namespace {

uint32_t *flags_of(lcpc_b200_shard *s, unsigned ch) {
  return reinterpret_cast<uint32_t *>(s->window + s->wl.flags) + ch * MAX_WORLD;
}
size_t flag_off(const lcpc_b200_shard *s, unsigned ch) { return s->wl.flags + (size_t)ch * MAX_WORLD * 4; }

int signal_all(lcpc_b200_shard *s, unsigned ch, uint32_t epoch, cudaStream_t st = nullptr) {
  lcpc_b200_ctx *ctx = s->enc->ctx;
  shard_signal_kernel<<<1, 32, 0, st ? st : ctx->stream>>>(s->peers, flag_off(s, ch), s->p.world, s->rank, epoch);
  ctx->launches += 1;
  CU(ctx, cudaGetLastError());
  return LCPC_B200_OK;
}

int wait_all(lcpc_b200_shard *s, unsigned ch, uint32_t epoch, cudaStream_t st = nullptr) {
  lcpc_b200_ctx *ctx = s->enc->ctx;
  shard_wait_kernel<<<1, 32, 0, st ? st : ctx->stream>>>(flags_of(s, ch), s->p.world, epoch,
                                               reinterpret_cast<uint32_t *>(s->window + s->wl.status), s->timeout_ns);
  ctx->launches += 1;
  CU(ctx, cudaGetLastError());
  return LCPC_B200_OK;
}

// after a synchronisation: did a wait time out?
int check_status(lcpc_b200_shard *s) {
  lcpc_b200_ctx *ctx = s->enc->ctx;
  uint32_t st = 0;
  CU(ctx, cudaMemcpy(&st, s->window + s->wl.status, 4, cudaMemcpyDeviceToHost));
  if (st) return fail(ctx, LCPC_B200_ERR_CUDA, "sharded commit: ranks (mask 0x%x) did not arrive within the timeout", st);
  return LCPC_B200_OK;
}

void shard_release(lcpc_b200_shard *s) {
  for (unsigned h = 0; h < s->p.world && s->connected; h++)
    if (s->peer_ipc[h] && s->peers.p[h]) cudaIpcCloseMemHandle(s->peers.p[h]);
  cudaFree(s->window);
  cudaFree(s->d_coeffs);
  cudaFree(s->d_tmp);
  cudaFree(s->d_forest);
  cudaFree(s->d_top);
  cudaFree(s->d_hash_scratch);
  cudaFree(s->d_enc_scratch);
  cudaFree(s->d_tensor);
  cudaFree(s->d_part);
  cudaFree(s->d_poly);
  cudaFree(s->d_repr);
  cudaFree(s->d_key);
  cudaFree(s->d_ones);
  cudaFree(s->d_cols);
  for (auto &e : s->ev)
    if (e) cudaEventDestroy(e);
  if (s->coeffs_free) cudaEventDestroy(s->coeffs_free);
  if (s->enc_done) cudaEventDestroy(s->enc_done);
  if (s->hash_done) cudaEventDestroy(s->hash_done);
  if (s->hash_stream) {
    cudaStreamSynchronize(s->hash_stream);
    cudaStreamDestroy(s->hash_stream);
  }
  delete s;
}

// root of an all-padding subtree: T zero leaves hashed up (lcpc-2d/src/lib.rs:665,696 leave those leaves zero)
int zero_subtree_root(lcpc_b200_shard *s, uint8_t out[32]) {
  lcpc_b200_ctx *ctx = s->enc->ctx;
  uint8_t *buf = nullptr;
  CU(ctx, cudaMalloc(&buf, 96));
  cudaError_t ce = cudaMemsetAsync(buf, 0, 96, ctx->stream);
  for (unsigned l = 0; l < s->sub_layers && ce == cudaSuccess; l++) {
    int nl = 0;
    ce = launch_merkle_layers(buf, 2, 1, ctx->stream, &nl);
    ctx->launches += nl;
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(buf, buf + 64, 32, cudaMemcpyDeviceToDevice, ctx->stream);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(buf + 32, buf + 64, 32, cudaMemcpyDeviceToDevice, ctx->stream);
  }
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(out, buf, 32, cudaMemcpyDeviceToHost, ctx->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
  cudaFree(buf);
  if (ce != cudaSuccess) return cuda_fail(ctx, ce, "zero subtree root");
  return LCPC_B200_OK;
}

// the engine stream waits for whatever the hash stream still has in flight (pipelined mode); enqueue only
int join_streams(lcpc_b200_shard *s) {
  lcpc_b200_ctx *ctx = s->enc->ctx;
  if (s->pipelined && s->hash_pending) CU(ctx, cudaStreamWaitEvent(ctx->stream, s->hash_done, 0));
  return LCPC_B200_OK;
}

// One commit of this rank's row block, enqueued in three steps.  A one-process-per-GPU host runs them back to back;
// a single process driving several shards runs step 1 on every shard, then step 2 on every shard, then step 3, so
// that every wait kernel is enqueued behind all the signals it waits for -- whichever hardware queue the driver
// maps the streams to (shards that share a device share its queues), no stream can be stuck behind a wait.
//   step 1: encode the rows (host or device memory with my_elems elements, or nullptr to re-encode what d_coeffs
//           holds), storing column block h into rank h's receive matrix; signal "tiles landed"
//   step 2: wait for everybody's tiles; hash this rank's columns; reduce its subtrees; store the subtree roots into
//           every rank's copy of the top tree's leaves; signal "roots landed"
//   step 3: wait for everybody's roots; top tree -> LcRoot
int commit_step1(lcpc_b200_shard *s, const void *rows, size_t n_elems, bool rows_on_host) {
  lcpc_b200_enc *enc = s->enc;
  lcpc_b200_ctx *ctx = enc->ctx;
  const ShardPlan &p = s->p;
  if (!s->connected) return fail(ctx, LCPC_B200_ERR_BAD_ARG, "shard: not connected to its peers");
  if (rows && n_elems != s->my_elems)
    return fail(ctx, LCPC_B200_ERR_BAD_ARG, "shard %u holds %zu coefficients, %zu given", s->rank, s->my_elems, n_elems);
  cudaStream_t st = ctx->stream;
  const uint32_t e = ++s->epoch;
  const unsigned par = e & 1;
  // pipelined mode: the engine stream itself never waits for this commit's exchange, so before it overwrites the
  // receive matrices of two commits ago it makes sure every rank has finished hashing them ("roots landed" of e - 2
  // is sent behind a rank's leaf hashing and subtree reduction); normally long satisfied
  if (s->pipelined && e >= 3)
    if (int rc = wait_all(s, CH_ROOTS, e - 2, st)) return rc;
  CU(ctx, cudaEventRecord(s->ev[0], st));
  if (s->my_rows) {
    Scatter sc;
    sc.n_blocks = p.world, sc.row0 = p.row_lo[s->rank];
    for (unsigned h = 0; h <= p.world; h++) sc.starts[h] = p.col_lo[h];
    for (unsigned h = 0; h < p.world; h++) sc.dst[h] = reinterpret_cast<uint32_t *>(s->peers.p[h] + s->wl.recv[par]);
    if (rows && rows_on_host) {
      if (int rc = encode_rows_from_host(enc, rows, n_elems, s->d_coeffs, s->d_tmp, s->my_rows, s->d_enc_scratch, nullptr, &sc,
                                         nullptr, nullptr, s->coeffs_free_valid ? s->coeffs_free : nullptr))
        return rc;
    } else {
      if (rows && rows != s->d_coeffs) {
        CU(ctx, cudaMemcpyAsync(s->d_coeffs, rows, n_elems * s->B, cudaMemcpyDeviceToDevice, st));
        const size_t padded = s->my_rows * p.n_per_row;
        if (padded > n_elems) CU(ctx, cudaMemsetAsync((uint8_t *)s->d_coeffs + n_elems * s->B, 0, (padded - n_elems) * s->B, st));
      }
      // SHARD_ENC_CHUNKS (Ligero): the row block in chunks (of at least 8 rows) that alternate between the engine and
      // the side stream, so that one chunk's last pass (the one that stores over NVLink) runs beside the next chunk's first
      // (8 GPUs, 2^24: encode + peer stores 0.705 -> 0.672 ms with 2 or 4 chunks, profiles/r02_bench_g8_ligero_enc_chunks*.json)
      const size_t chunks = std::min<size_t>((size_t)std::max<long>(1, tunable("SHARD_ENC_CHUNKS", 4)), std::max<size_t>(1, s->my_rows / 8));
      if (chunks > 1 && enc->kind == LCPC_B200_ENC_LIGERO) {
        CU(ctx, cudaEventRecord(ctx->lane_fork, st));
        CU(ctx, cudaStreamWaitEvent(ctx->side_stream, ctx->lane_fork, 0));
        for (size_t k = 0; k < chunks; k++) {
          const size_t q0 = k * s->my_rows / chunks, q1 = (k + 1) * s->my_rows / chunks;
          Scatter sk = sc;
          sk.row0 += q0;
          int nl = 0;
          cudaError_t ce = launch_ntt_rows(enc->field, s->d_coeffs + q0 * p.n_per_row * s->N, p.n_per_row, p.n_per_row,
                                           s->d_tmp + q0 * p.n_cols * s->N, p.n_cols, enc->d_roots, enc->log_n, q1 - q0,
                                           (k & 1) ? ctx->side_stream : st, &nl, &sk);
          ctx->launches += nl;
          if (ce != cudaSuccess) return cuda_fail(ctx, ce, "shard: encode");
        }
        CU(ctx, cudaEventRecord(ctx->lane_join, ctx->side_stream));
        CU(ctx, cudaStreamWaitEvent(st, ctx->lane_join, 0));
      } else if (int rc = encode_rows(enc, s->d_coeffs, p.n_per_row, p.n_per_row, s->d_tmp, s->my_rows, s->d_enc_scratch, &sc)) {
        return rc;
      }
    }
  }
  CU(ctx, cudaEventRecord(s->ev[1], st));
  CU(ctx, cudaEventRecord(s->coeffs_free, st));
  s->coeffs_free_valid = true;
  if (int rc = signal_all(s, CH_TILES, e, st)) return rc;
  if (s->pipelined) CU(ctx, cudaEventRecord(s->enc_done, st));
  return LCPC_B200_OK;
}

int commit_step2(lcpc_b200_shard *s) {
  lcpc_b200_enc *enc = s->enc;
  lcpc_b200_ctx *ctx = enc->ctx;
  const ShardPlan &p = s->p;
  cudaStream_t st = s->pipelined ? s->hash_stream : ctx->stream;
  const uint32_t e = s->epoch;
  const unsigned par = e & 1;
  if (s->pipelined) CU(ctx, cudaStreamWaitEvent(st, s->enc_done, 0));
  if (int rc = wait_all(s, CH_TILES, e, st)) return rc;
  CU(ctx, cudaEventRecord(s->ev[2], st));
  if (s->my_cols) {
    int nl = 0;
    cudaError_t ce = launch_hash_columns(enc->field, reinterpret_cast<const uint32_t *>(s->window + s->wl.recv[par]), p.n_rows,
                                         s->my_cols, s->my_cols, s->d_forest, s->d_hash_scratch, st, &nl);
    ctx->launches += nl;
    if (ce == cudaSuccess && s->sub_layers) {
      ce = launch_merkle_layers(s->d_forest, s->forest_leaves, s->sub_layers, st, &nl);
      ctx->launches += nl;
    }
    if (ce != cudaSuccess) return cuda_fail(ctx, ce, "shard: hash_columns");
  }
  // subtree roots -> every rank's copy of the top tree's leaves, then "roots landed"
  shard_push_signal_kernel<<<1, 256, 0, st>>>(reinterpret_cast<const uint4 *>(s->d_forest + s->roots_off), s->my_subs * 2, s->peers,
                                              s->wl.top[par] + p.sub_lo[s->rank] * 32, flag_off(s, CH_ROOTS), p.world, s->rank, e);
  ctx->launches += 1;
  CU(ctx, cudaGetLastError());
  return LCPC_B200_OK;
}

int commit_step3(lcpc_b200_shard *s) {
  lcpc_b200_ctx *ctx = s->enc->ctx;
  const ShardPlan &p = s->p;
  cudaStream_t st = s->pipelined ? s->hash_stream : ctx->stream;
  const uint32_t e = s->epoch;
  const unsigned par = e & 1;
  if (int rc = wait_all(s, CH_ROOTS, e, st)) return rc;
  CU(ctx, cudaMemcpyAsync(s->d_top, s->window + s->wl.top[par], p.n_real_sub * 32, cudaMemcpyDeviceToDevice, st));
  if (p.n_sub > 1) {
    unsigned lg = 0;
    while (((size_t)1 << lg) < p.n_sub) lg++;
    int nl = 0;
    cudaError_t ce = launch_merkle_layers(s->d_top, p.n_sub, lg, st, &nl);
    ctx->launches += nl;
    if (ce != cudaSuccess) return cuda_fail(ctx, ce, "shard: top tree");
  }
  CU(ctx, cudaEventRecord(s->ev[3], st));
  if (s->pipelined) {
    CU(ctx, cudaEventRecord(s->hash_done, st));
    s->hash_pending = true;
  }
  return LCPC_B200_OK;
}

int commit_enqueue(lcpc_b200_shard *s, const void *rows, size_t n_elems, bool rows_on_host) {
  if (int rc = commit_step1(s, rows, n_elems, rows_on_host)) return rc;
  if (int rc = commit_step2(s)) return rc;
  return commit_step3(s);
}

}  // namespace

extern "C" {

int lcpc_b200_shard_plan(size_t n_rows, size_t n_per_row, size_t n_cols, unsigned world, size_t *row_lo, size_t *col_lo,
                         size_t *sub_lo, size_t *sub_leaves, size_t *n_sub) {
  if (world == 0 || world > MAX_WORLD || n_cols == 0) return LCPC_B200_ERR_BAD_ARG;
  ShardPlan p;
  make_plan(n_rows, n_per_row, n_cols, world, &p);
  for (unsigned g = 0; g <= world; g++) {
    if (row_lo) row_lo[g] = p.row_lo[g];
    if (col_lo) col_lo[g] = p.col_lo[g];
    if (sub_lo) sub_lo[g] = p.sub_lo[g];
  }
  if (sub_leaves) *sub_leaves = p.sub_leaves;
  if (n_sub) *n_sub = p.n_sub;
  return LCPC_B200_OK;
}

int lcpc_b200_shard_new(lcpc_b200_enc *enc, size_t len, unsigned world, unsigned rank, size_t max_open,
                        lcpc_b200_shard **out) {
  if (!enc || !out) return LCPC_B200_ERR_BAD_ARG;
  *out = nullptr;
  lcpc_b200_ctx *ctx = enc->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (world == 0 || world > MAX_WORLD || rank >= world) return fail(ctx, LCPC_B200_ERR_BAD_ARG, "shard: rank %u of %u", rank, world);
  if (len == 0) return fail(ctx, LCPC_B200_ERR_BAD_ARG, "commit: empty coefficient vector");
  if (int rc = bind_device(ctx)) return rc;
  lcpc_b200_shard *s = new (std::nothrow) lcpc_b200_shard;
  if (!s) return LCPC_B200_ERR_OOM;
  s->enc = enc, s->rank = rank, s->max_open = max_open;
  const size_t n_per_row = enc->n_per_row, n_cols = enc->n_cols, n_rows = (len + n_per_row - 1) / n_per_row;
  make_plan(n_rows, n_per_row, n_cols, world, &s->p);
  const ShardPlan &p = s->p;
  s->B = field_bytes(enc->field), s->N = s->B / 4;
  s->wl = make_layout(p, s->B, max_open);
  s->my_rows = p.row_lo[rank + 1] - p.row_lo[rank];
  s->my_cols = p.col_lo[rank + 1] - p.col_lo[rank];
  s->my_subs = p.sub_lo[rank + 1] - p.sub_lo[rank];
  {
    const size_t lo = p.row_lo[rank] * n_per_row, hi = std::min(p.row_lo[rank + 1] * n_per_row, len);
    s->my_elems = hi > lo ? hi - lo : 0;
  }
  while (((size_t)1 << s->sub_layers) < p.sub_leaves) s->sub_layers++;
  while (((size_t)1 << s->path_len) < p.np2) s->path_len++;
  s->forest_leaves = s->my_subs * p.sub_leaves;
  for (unsigned l = 0; l <= s->sub_layers; l++) {
    if (l == s->sub_layers) s->roots_off = s->forest_nodes * 32;
    s->forest_nodes += s->forest_leaves >> l;
  }
  const size_t B = s->B;
  cudaError_t ce = cudaMalloc(&s->window, s->wl.total);
  if (ce == cudaSuccess) ce = cudaMemset(s->window, 0, s->wl.recv[0]);  // flags + status
  auto alloc = [&](auto **ptr, size_t bytes) {
    if (ce == cudaSuccess) ce = cudaMalloc(ptr, std::max<size_t>(bytes, 256));
  };
  alloc(&s->d_coeffs, s->my_rows * n_per_row * B);
  alloc(&s->d_tmp, s->my_rows * n_cols * B);
  alloc(&s->d_forest, s->forest_nodes * 32);
  alloc(&s->d_top, (2 * p.n_sub - 1) * 32);
  alloc(&s->d_hash_scratch, hash_scratch_bytes(enc->field, n_rows, s->my_cols));
  alloc(&s->d_enc_scratch, enc_scratch_bytes(enc, s->my_rows));
  alloc(&s->d_tensor, n_rows * B);
  alloc(&s->d_part, n_per_row * B);
  alloc(&s->d_poly, n_per_row * B);
  alloc(&s->d_repr, n_per_row * B);
  alloc(&s->d_key, 32);
  alloc(&s->d_ones, world * B);
  alloc(&s->d_cols, max_open * 8);
  for (auto &e : s->ev)
    if (ce == cudaSuccess) ce = cudaEventCreate(&e);
  if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&s->coeffs_free, cudaEventDisableTiming);
  if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&s->enc_done, cudaEventDisableTiming);
  if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&s->hash_done, cudaEventDisableTiming);
  if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&s->hash_stream, cudaStreamNonBlocking);
  // 2 GPUs, 2^24 (profiles/r02_bench_g2_pipeline{0,1}.json): Brakedown 1.057 -> 0.968 ms per commit, Ligero unchanged
  // (its leaf hashing is a full-GPU ALU-bound kernel at that share); on by default
  s->pipelined = tunable("SHARD_PIPELINE", 1) != 0;
  if (ce == cudaSuccess) ce = cudaMemset(s->d_forest, 0, std::max<size_t>(s->forest_nodes * 32, 256));
  if (ce == cudaSuccess) ce = cudaMemset(s->d_top, 0, (2 * p.n_sub - 1) * 32);
  if (ce == cudaSuccess) ce = cudaMemset(s->d_coeffs, 0, std::max<size_t>(s->my_rows * n_per_row * B, 256));
  if (ce == cudaSuccess) {  // `world` copies of Field::one(): the tensor that sums the per-rank partial combinations
    std::vector<uint64_t> ones(world * (B / 8));
    for (unsigned h = 0; h < world; h++) lcpc_b200_field_one(enc->field, ones.data() + h * (B / 8));
    ce = cudaMemcpy(s->d_ones, ones.data(), world * B, cudaMemcpyHostToDevice);
  }
  if (ce != cudaSuccess) {
    shard_release(s);
    return cuda_fail(ctx, ce, "shard_new");
  }
  // all-padding subtrees (non-power-of-two n_cols): their constant root fills the tail of the top tree's leaves
  if (p.n_real_sub < p.n_sub) {
    uint8_t zr[32];
    int rc = zero_subtree_root(s, zr);
    if (rc == LCPC_B200_OK) {
      std::vector<uint8_t> pad((p.n_sub - p.n_real_sub) * 32);
      for (size_t i = 0; i < p.n_sub - p.n_real_sub; i++) memcpy(pad.data() + 32 * i, zr, 32);
      ce = cudaMemcpy(s->d_top + p.n_real_sub * 32, pad.data(), pad.size(), cudaMemcpyHostToDevice);
      if (ce != cudaSuccess) rc = cuda_fail(ctx, ce, "shard_new: top padding");
    }
    if (rc != LCPC_B200_OK) {
      shard_release(s);
      return rc;
    }
  }
  enc->refs.fetch_add(1);
  *out = s;
  return LCPC_B200_OK;
}

void lcpc_b200_shard_free(lcpc_b200_shard *s) {
  if (!s) return;
  lcpc_b200_enc *enc = s->enc;
  {
    lcpc_b200_ctx *ctx = enc->ctx;
    std::lock_guard<std::mutex> g(ctx->mu);
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    shard_release(s);
  }
  enc_unref(enc);
}

int lcpc_b200_shard_window(lcpc_b200_shard *s, void **d_ptr, size_t *bytes, uint8_t ipc_handle[64]) {
  if (!s) return LCPC_B200_ERR_BAD_ARG;
  lcpc_b200_ctx *ctx = s->enc->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (int rc = bind_device(ctx)) return rc;
  if (d_ptr) *d_ptr = s->window;
  if (bytes) *bytes = s->wl.total;
  if (ipc_handle) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    cudaIpcMemHandle_t h;
    CU(ctx, cudaIpcGetMemHandle(&h, s->window));
    memcpy(ipc_handle, &h, 64);
  }
  return LCPC_B200_OK;
}

int lcpc_b200_shard_connect(lcpc_b200_shard *s, void *const *peer_ptrs, const uint8_t *ipc_handles) {
  if (!s || (!peer_ptrs && !ipc_handles && s->p.world > 1)) return LCPC_B200_ERR_BAD_ARG;
  lcpc_b200_ctx *ctx = s->enc->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (s->connected) return fail(ctx, LCPC_B200_ERR_BAD_ARG, "shard: already connected");
  if (int rc = bind_device(ctx)) return rc;
  for (unsigned h = 0; h < s->p.world; h++) {
    if (h == s->rank) {
      s->peers.p[h] = s->window;
      continue;
    }
    if (peer_ptrs && peer_ptrs[h]) {  // same process: direct peer access
      cudaPointerAttributes at;
      CU(ctx, cudaPointerGetAttributes(&at, peer_ptrs[h]));
      if (at.type != cudaMemoryTypeDevice) return fail(ctx, LCPC_B200_ERR_BAD_ARG, "shard: peer %u window is not device memory", h);
      if (at.device != ctx->device) {
        int can = 0;
        CU(ctx, cudaDeviceCanAccessPeer(&can, ctx->device, at.device));
        if (!can) return fail(ctx, LCPC_B200_ERR_UNSUPPORTED, "shard: device %d cannot map device %d (no peer access)", ctx->device, at.device);
        cudaError_t ce = cudaDeviceEnablePeerAccess(at.device, 0);
        if (ce == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else if (ce != cudaSuccess) return cuda_fail(ctx, ce, "cudaDeviceEnablePeerAccess");
      }
      s->peers.p[h] = (uint8_t *)peer_ptrs[h];
    } else if (ipc_handles) {  // another process: map its window through the IPC handle
      cudaIpcMemHandle_t hd;
      memcpy(&hd, ipc_handles + 64 * (size_t)h, 64);
      void *ptr = nullptr;
      cudaError_t ce = cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess);
      if (ce != cudaSuccess) return cuda_fail(ctx, ce, "cudaIpcOpenMemHandle");
      s->peers.p[h] = (uint8_t *)ptr, s->peer_ipc[h] = true;
    } else {
      return fail(ctx, LCPC_B200_ERR_BAD_ARG, "shard: no window for peer %u", h);
    }
  }
  s->connected = true;
  return LCPC_B200_OK;
}

int lcpc_b200_shard_dims(const lcpc_b200_shard *s, size_t *n_rows, size_t *n_per_row, size_t *n_cols, size_t *row_lo,
                         size_t *row_hi, size_t *col_lo, size_t *col_hi, size_t *n_elems) {
  if (!s) return LCPC_B200_ERR_BAD_ARG;
  if (n_rows) *n_rows = s->p.n_rows;
  if (n_per_row) *n_per_row = s->p.n_per_row;
  if (n_cols) *n_cols = s->p.n_cols;
  if (row_lo) *row_lo = s->p.row_lo[s->rank];
  if (row_hi) *row_hi = s->p.row_lo[s->rank + 1];
  if (col_lo) *col_lo = s->p.col_lo[s->rank];
  if (col_hi) *col_hi = s->p.col_lo[s->rank + 1];
  if (n_elems) *n_elems = s->my_elems;
  return LCPC_B200_OK;
}

int lcpc_b200_shard_commit(lcpc_b200_shard *s, const uint64_t *rows, size_t n_elems) {
  if (!s || (!rows && n_elems)) return LCPC_B200_ERR_BAD_ARG;
  lcpc_b200_ctx *ctx = s->enc->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (int rc = bind_device(ctx)) return rc;
  return commit_enqueue(s, rows, n_elems, true);
}

int lcpc_b200_shard_commit_dev(lcpc_b200_shard *s, const uint64_t *d_rows, size_t n_elems) {
  if (!s) return LCPC_B200_ERR_BAD_ARG;
  lcpc_b200_ctx *ctx = s->enc->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (int rc = bind_device(ctx)) return rc;
  return commit_enqueue(s, d_rows, d_rows ? n_elems : 0, false);
}

int lcpc_b200_shard_commit_step(lcpc_b200_shard *s, int step) {
  if (!s || step < 1 || step > 3) return LCPC_B200_ERR_BAD_ARG;
  lcpc_b200_ctx *ctx = s->enc->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (int rc = bind_device(ctx)) return rc;
  if (step == 1) return commit_step1(s, nullptr, 0, false);
  if (s->epoch == 0) return fail(ctx, LCPC_B200_ERR_BAD_ARG, "shard: step %d before step 1", step);
  return step == 2 ? commit_step2(s) : commit_step3(s);
}

int lcpc_b200_shard_load_rows(lcpc_b200_shard *s, const uint64_t *rows, size_t n_elems) {
  if (!s || (!rows && n_elems)) return LCPC_B200_ERR_BAD_ARG;
  lcpc_b200_ctx *ctx = s->enc->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (n_elems != s->my_elems) return fail(ctx, LCPC_B200_ERR_BAD_ARG, "shard %u holds %zu coefficients, %zu given", s->rank, s->my_elems, n_elems);
  if (int rc = bind_device(ctx)) return rc;
  if (int rc = join_streams(s)) return rc;
  if (n_elems) CU(ctx, cudaMemcpyAsync(s->d_coeffs, rows, n_elems * s->B, cudaMemcpyHostToDevice, ctx->stream));
  const size_t padded = s->my_rows * s->p.n_per_row;
  if (padded > n_elems) CU(ctx, cudaMemsetAsync((uint8_t *)s->d_coeffs + n_elems * s->B, 0, (padded - n_elems) * s->B, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return LCPC_B200_OK;
}

int lcpc_b200_shard_root(lcpc_b200_shard *s, uint8_t root[32]) {
  if (!s || !root) return LCPC_B200_ERR_BAD_ARG;
  lcpc_b200_ctx *ctx = s->enc->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (int rc = bind_device(ctx)) return rc;
  if (s->epoch == 0) return fail(ctx, LCPC_B200_ERR_BAD_ARG, "shard: no commit has run");
  if (int rc = join_streams(s)) return rc;
  CU(ctx, cudaMemcpyAsync(root, s->d_top + (2 * s->p.n_sub - 2) * 32, 32, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return check_status(s);
}

int lcpc_b200_shard_join(lcpc_b200_shard *s) {
  if (!s) return LCPC_B200_ERR_BAD_ARG;
  lcpc_b200_ctx *ctx = s->enc->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (int rc = bind_device(ctx)) return rc;
  return join_streams(s);
}

int lcpc_b200_shard_root_enqueue(lcpc_b200_shard *s, uint8_t *root) {
  if (!s || !root) return LCPC_B200_ERR_BAD_ARG;
  lcpc_b200_ctx *ctx = s->enc->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (int rc = bind_device(ctx)) return rc;
  if (s->epoch == 0) return fail(ctx, LCPC_B200_ERR_BAD_ARG, "shard: no commit has run");
  // behind the top tree of the last enqueued commit, on whichever stream computes it
  CU(ctx, cudaMemcpyAsync(root, s->d_top + (2 * s->p.n_sub - 2) * 32, 32, cudaMemcpyDeviceToHost,
                          s->pipelined ? s->hash_stream : ctx->stream));
  if (s->pipelined) CU(ctx, cudaEventRecord(s->hash_done, s->hash_stream));
  return LCPC_B200_OK;
}

int lcpc_b200_shard_phase_times(lcpc_b200_shard *s, float ms[3]) {
  if (!s || !ms) return LCPC_B200_ERR_BAD_ARG;
  lcpc_b200_ctx *ctx = s->enc->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (int rc = bind_device(ctx)) return rc;
  CU(ctx, cudaEventSynchronize(s->ev[3]));
  for (int i = 0; i < 3; i++) CU(ctx, cudaEventElapsedTime(&ms[i], s->ev[i], s->ev[i + 1]));
  return LCPC_B200_OK;
}

int lcpc_b200_shard_device_ptrs(lcpc_b200_shard *s, uint64_t **d_recv, uint64_t **d_coeffs, uint8_t **d_leaves, uint8_t **d_top) {
  if (!s) return LCPC_B200_ERR_BAD_ARG;
  if (d_recv) *d_recv = (uint64_t *)(s->window + s->wl.recv[s->epoch & 1]);
  if (d_coeffs) *d_coeffs = (uint64_t *)s->d_coeffs;
  if (d_leaves) *d_leaves = s->d_forest;
  if (d_top) *d_top = s->d_top;
  return LCPC_B200_OK;
}

int lcpc_b200_shard_download(lcpc_b200_shard *s, uint64_t *cols_out, uint8_t *leaves_out) {
  if (!s) return LCPC_B200_ERR_BAD_ARG;
  lcpc_b200_ctx *ctx = s->enc->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (int rc = bind_device(ctx)) return rc;
  if (s->epoch == 0) return fail(ctx, LCPC_B200_ERR_BAD_ARG, "shard: no commit has run");
  if (int rc = join_streams(s)) return rc;
  if (cols_out && s->my_cols)
    CU(ctx, cudaMemcpyAsync(cols_out, s->window + s->wl.recv[s->epoch & 1], s->p.n_rows * s->my_cols * s->B, cudaMemcpyDeviceToHost, ctx->stream));
  if (leaves_out && s->my_cols) CU(ctx, cudaMemcpyAsync(leaves_out, s->d_forest, s->my_cols * 32, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return check_status(s);
}

// ------------------------------------------------------------------------------------------- prove, split-phase
// collapse_columns (lcpc-2d/src/lib.rs:1095-1123) with the coefficient rows sharded by row block: every rank combines its
// rows with its slice of the tensor, stores the partial vector into slot `rank` of every peer's window and signals;
// `finish` waits for all partials and sums them (a collapse with an all-ones tensor) -- no field arithmetic on the host.
int lcpc_b200_shard_collapse_begin(lcpc_b200_shard *s, const uint64_t *tensor, const uint8_t key[32]) {
  if (!s || (!tensor && !key)) return LCPC_B200_ERR_BAD_ARG;
  lcpc_b200_ctx *ctx = s->enc->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (!s->connected) return fail(ctx, LCPC_B200_ERR_BAD_ARG, "shard: not connected");
  if (s->collapse_seq != s->collapse_done) return fail(ctx, LCPC_B200_ERR_BAD_ARG, "shard: a collapse is already in flight");
  if (int rc = bind_device(ctx)) return rc;
  if (int rc = join_streams(s)) return rc;
  const ShardPlan &p = s->p;
  const int field = s->enc->field;
  cudaStream_t st = ctx->stream;
  const uint32_t k = ++s->collapse_seq;
  const unsigned par = k & 1;
  if (key) {  // challenge tensor expanded on the device (lcpc-2d/src/lib.rs:1026-1032)
    CU(ctx, cudaMemcpyAsync(s->d_key, key, 32, cudaMemcpyHostToDevice, st));
    cudaError_t ce = launch_expand_tensor(field, s->d_key, 0, p.n_rows, s->d_tensor, st);
    ctx->launches += 1;
    if (ce != cudaSuccess) return cuda_fail(ctx, ce, "shard: expand_tensor");
  } else {
    CU(ctx, cudaMemcpyAsync(s->d_tensor, tensor, p.n_rows * s->B, cudaMemcpyHostToDevice, st));
  }
  if (s->my_rows) {
    int nl = 0;
    cudaError_t ce = launch_collapse(field, s->d_coeffs, p.n_per_row, s->d_tensor + p.row_lo[s->rank] * s->N, s->d_part, s->my_rows,
                                     p.n_per_row, nullptr, st, &nl);
    ctx->launches += nl;
    if (ce != cudaSuccess) return cuda_fail(ctx, ce, "shard: collapse");
    CU(ctx, cudaEventRecord(s->coeffs_free, st));
  } else {
    CU(ctx, cudaMemsetAsync(s->d_part, 0, p.n_per_row * s->B, st));  // the zero element is all-zero limbs
  }
  const size_t n16 = p.n_per_row * s->B / 16, total = n16 * p.world;
  if (p.n_per_row * s->B % 16) return fail(ctx, LCPC_B200_ERR_UNSUPPORTED, "shard: odd row length for an 8-byte field");
  const unsigned grid = (unsigned)std::min<size_t>((total + 255) / 256, 148 * 8);
  shard_push_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const uint4 *>(s->d_part), n16, s->peers,
                                          s->wl.parts[par] + (size_t)s->rank * p.n_per_row * s->B, p.world);
  ctx->launches += 1;
  CU(ctx, cudaGetLastError());
  return signal_all(s, CH_PARTS, k);
}

int lcpc_b200_shard_collapse_finish(lcpc_b200_shard *s, uint64_t *poly, uint8_t *repr) {
  if (!s || !poly) return LCPC_B200_ERR_BAD_ARG;
  lcpc_b200_ctx *ctx = s->enc->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (s->collapse_seq == s->collapse_done) return fail(ctx, LCPC_B200_ERR_BAD_ARG, "shard: no collapse in flight");
  if (int rc = bind_device(ctx)) return rc;
  const ShardPlan &p = s->p;
  const int field = s->enc->field;
  cudaStream_t st = ctx->stream;
  const uint32_t k = s->collapse_seq;
  const unsigned par = k & 1;
  if (int rc = wait_all(s, CH_PARTS, k)) return rc;
  int nl = 0;
  cudaError_t ce = launch_collapse(field, reinterpret_cast<const uint32_t *>(s->window + s->wl.parts[par]), p.n_per_row, s->d_ones,
                                   s->d_poly, p.world, p.n_per_row, nullptr, st, &nl);
  ctx->launches += nl;
  if (ce == cudaSuccess && repr) {
    ce = launch_field_op(field, 4, s->d_repr, s->d_poly, nullptr, p.n_per_row, st);
    ctx->launches += 1;
  }
  if (ce != cudaSuccess) return cuda_fail(ctx, ce, "shard: collapse sum");
  const size_t bytes = p.n_per_row * s->B;
  if (repr) CU(ctx, cudaMemcpyAsync(repr, s->d_repr, bytes, cudaMemcpyDeviceToHost, st));
  CU(ctx, cudaMemcpyAsync(poly, s->d_poly, bytes, cudaMemcpyDeviceToHost, st));
  CU(ctx, cudaStreamSynchronize(st));
  s->collapse_done = k;
  return check_status(s);
}

// open_column (lcpc-2d/src/lib.rs:788-825) for `n` columns over the column-sharded commit (see shard_open_kernel)
int lcpc_b200_shard_open_begin(lcpc_b200_shard *s, const uint64_t *cols, size_t n) {
  if (!s || (n && !cols)) return LCPC_B200_ERR_BAD_ARG;
  lcpc_b200_ctx *ctx = s->enc->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (!s->connected || s->epoch == 0) return fail(ctx, LCPC_B200_ERR_BAD_ARG, "shard: nothing committed");
  if (s->open_seq != s->open_done) return fail(ctx, LCPC_B200_ERR_BAD_ARG, "shard: an opening is already in flight");
  if (n > s->max_open) return fail(ctx, LCPC_B200_ERR_BAD_ARG, "shard: %zu openings, capacity %zu", n, s->max_open);
  const ShardPlan &p = s->p;
  for (size_t i = 0; i < n; i++)
    if (cols[i] >= p.n_cols) return fail(ctx, LCPC_B200_ERR_COLUMN, "column %llu >= n_cols %zu", (unsigned long long)cols[i], p.n_cols);
  if (int rc = bind_device(ctx)) return rc;
  if (int rc = join_streams(s)) return rc;
  cudaStream_t st = ctx->stream;
  const uint32_t k = ++s->open_seq;
  const unsigned par = k & 1;
  s->open_n = n;
  if (n) {
    CU(ctx, cudaMemcpyAsync(s->d_cols, cols, n * 8, cudaMemcpyHostToDevice, st));
    OpenArgs a;
    a.recv = reinterpret_cast<const uint32_t *>(s->window + s->wl.recv[s->epoch & 1]);
    a.forest = s->d_forest, a.top = s->d_top, a.cols = s->d_cols;
    a.n_open = n, a.n_rows = p.n_rows, a.my_cols = s->my_cols, a.c0 = p.col_lo[s->rank], a.c1 = p.col_lo[s->rank + 1];
    a.forest_leaves = s->forest_leaves, a.sub_leaves = p.sub_leaves, a.n_sub = p.n_sub;
    a.n_limbs = (unsigned)s->N, a.sub_layers = s->sub_layers, a.path_len = s->path_len, a.world = p.world;
    a.vals_off = s->wl.open_vals[par], a.paths_off = s->wl.open_paths[par];
    const size_t work = n * p.n_rows * (s->N / ((s->N % 4 == 0) ? 4 : 2));
    const unsigned grid = (unsigned)std::min<size_t>((work + 255) / 256, 148 * 8);
    shard_open_kernel<<<std::max(grid, 1u), 256, 0, st>>>(a, s->peers);
    ctx->launches += 1;
    CU(ctx, cudaGetLastError());
  }
  return signal_all(s, CH_OPEN, k);
}

int lcpc_b200_shard_open_finish(lcpc_b200_shard *s, uint64_t *cols_out, uint8_t *paths_out) {
  if (!s) return LCPC_B200_ERR_BAD_ARG;
  lcpc_b200_ctx *ctx = s->enc->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (s->open_seq == s->open_done) return fail(ctx, LCPC_B200_ERR_BAD_ARG, "shard: no opening in flight");
  if (s->open_n && (!cols_out || (s->path_len && !paths_out))) return LCPC_B200_ERR_BAD_ARG;
  if (int rc = bind_device(ctx)) return rc;
  cudaStream_t st = ctx->stream;
  const uint32_t k = s->open_seq;
  const unsigned par = k & 1;
  if (int rc = wait_all(s, CH_OPEN, k)) return rc;
  if (s->open_n) {
    CU(ctx, cudaMemcpyAsync(cols_out, s->window + s->wl.open_vals[par], s->open_n * s->p.n_rows * s->B, cudaMemcpyDeviceToHost, st));
    if (s->path_len)
      CU(ctx, cudaMemcpyAsync(paths_out, s->window + s->wl.open_paths[par], s->open_n * (size_t)s->path_len * 32, cudaMemcpyDeviceToHost, st));
  }
  CU(ctx, cudaStreamSynchronize(st));
  s->open_done = k;
  return check_status(s);
}

// LcCommit::prove (lcpc-2d/src/lib.rs:1004-1093) on the sharded commit.  In the one-process-per-GPU shape every rank
// calls this with an identical transcript and gets the same proof; the outer tensor's combination (independent of
// the transcript) runs on the device while the host absorbs the last degree test's vector.
int lcpc_b200_shard_prove(lcpc_b200_shard *s, lcpc_b200_transcript *tr, const lcpc_b200_labels *labels,
                          const uint64_t *outer_tensor, size_t outer_len, size_t n_degree_tests, size_t n_col_opens,
                          uint64_t *p_eval, uint64_t *p_random, uint64_t *col_idx, uint64_t *cols_out, uint8_t *paths_out) {
  if (!s || !tr || !outer_tensor || !p_eval || (n_degree_tests && !p_random) || (n_col_opens && (!cols_out || !paths_out)))
    return LCPC_B200_ERR_BAD_ARG;
  lcpc_b200_ctx *ctx = s->enc->ctx;
  const ShardPlan &p = s->p;
  if (outer_len != p.n_rows) {  // ProverError::OuterTensor (:1016-1018)
    std::lock_guard<std::mutex> g(ctx->mu);
    return fail(ctx, LCPC_B200_ERR_OUTER_TENSOR, "outer tensor has %zu entries, the commitment %zu rows", outer_len, p.n_rows);
  }
  const Labels lb = resolve_labels(labels);
  const size_t B = s->B, L = B / 8, pbytes = p.n_per_row * B;
  // canonical bytes land in page-locked staging (two vectors: the one being absorbed, the one arriving)
  uint8_t *stage = nullptr;
  std::vector<uint8_t> pageable;
  {
    std::lock_guard<std::mutex> g(ctx->mu);
    stage = (uint8_t *)host_stage(ctx, 2 * pbytes);
  }
  if (!stage) {
    pageable.resize(2 * pbytes);
    stage = pageable.data();
  }
  uint8_t *repr_a = stage, *repr_b = stage + pbytes;
  int rc = LCPC_B200_OK;
  for (size_t i = 0; i < n_degree_tests && rc == LCPC_B200_OK; i++) {  // :1025-1048
    uint8_t key[32];
    tr->tr.challenge_bytes(lb.dt, lb.dt_len, key, 32);
    rc = lcpc_b200_shard_collapse_begin(s, nullptr, key);
    if (rc == LCPC_B200_OK) rc = lcpc_b200_shard_collapse_finish(s, p_random + i * p.n_per_row * L, repr_a);
    if (rc != LCPC_B200_OK) break;
    if (i + 1 == n_degree_tests) {  // p_eval's combination does not depend on the transcript: run it under the absorb
      rc = lcpc_b200_shard_collapse_begin(s, outer_tensor, nullptr);
      if (rc != LCPC_B200_OK) break;
    }
    tr->tr.append_elems(lb.pr, lb.pr_len, repr_a, B, p.n_per_row);
  }
  if (rc == LCPC_B200_OK && n_degree_tests == 0) rc = lcpc_b200_shard_collapse_begin(s, outer_tensor, nullptr);
  if (rc == LCPC_B200_OK) rc = lcpc_b200_shard_collapse_finish(s, p_eval, repr_b);  // :1051-1063
  if (rc != LCPC_B200_OK) return rc;
  tr->tr.append_elems(lb.pe, lb.pe_len, repr_b, B, p.n_per_row);
  uint8_t key[32];
  tr->tr.challenge_bytes(lb.co, lb.co_len, key, 32);  // :1066-1085
  std::vector<uint64_t> cols(n_col_opens);
  sample_columns(key, p.n_cols, n_col_opens, cols.data());
  if (col_idx) memcpy(col_idx, cols.data(), n_col_opens * 8);
  rc = lcpc_b200_shard_open_begin(s, cols.data(), n_col_opens);
  if (rc == LCPC_B200_OK) rc = lcpc_b200_shard_open_finish(s, cols_out, paths_out);
  return rc;
}

// ------------------------------------------------------------------------------------- one process, several GPUs
// The reference's shape: ONE process calls commit() / prove().  lcpc_b200_multi owns one shard per encoding (each
// encoding on its own context = its own GPU), connects them by direct peer access, and drives all of them from the
// calling thread; every call only enqueues on the GPUs' streams until a result has to come back.
}  // extern "C"

struct lcpc_b200_multi {
  std::vector<lcpc_b200_shard *> shards;
  size_t len = 0;
};

extern "C" {

void lcpc_b200_multi_free(lcpc_b200_multi *m) {
  if (!m) return;
  // a shard's stream may still be spinning on a peer's flag: drain every stream before any window disappears
  for (auto *s : m->shards) lcpc_b200_ctx_synchronize(s->enc->ctx);
  for (auto *s : m->shards) lcpc_b200_shard_free(s);
  delete m;
}

int lcpc_b200_multi_rerun(lcpc_b200_multi *m, const uint64_t *coeffs_in, size_t len) {
  if (!m || !coeffs_in || len != m->len) return LCPC_B200_ERR_BAD_ARG;
  // step by step over all shards (see commit_step1): every wait is enqueued behind the signals it waits for
  for (int step = 1; step <= 3; step++) {
    for (auto *s : m->shards) {
      lcpc_b200_ctx *ctx = s->enc->ctx;
      std::lock_guard<std::mutex> g(ctx->mu);
      if (int rc = bind_device(ctx)) return rc;
      const size_t lo = s->p.row_lo[s->rank] * s->p.n_per_row;
      int rc = step == 1 ? commit_step1(s, coeffs_in + lo * (s->B / 8), s->my_elems, true) : step == 2 ? commit_step2(s) : commit_step3(s);
      if (rc) return rc;
    }
  }
  return LCPC_B200_OK;
}

int lcpc_b200_commit_new_multi(lcpc_b200_enc *const *encs, size_t n_gpus, const uint64_t *coeffs_in, size_t len,
                               size_t max_open, lcpc_b200_multi **out) {
  if (!encs || !out || n_gpus == 0 || n_gpus > MAX_WORLD || (!coeffs_in && len)) return LCPC_B200_ERR_BAD_ARG;
  *out = nullptr;
  for (size_t g = 0; g < n_gpus; g++) {
    if (!encs[g]) return LCPC_B200_ERR_BAD_ARG;
    if (encs[g]->field != encs[0]->field || encs[g]->n_per_row != encs[0]->n_per_row || encs[g]->n_cols != encs[0]->n_cols ||
        encs[g]->kind != encs[0]->kind)
      return LCPC_B200_ERR_BAD_ARG;  // every GPU must hold the same encoding
  }
  lcpc_b200_multi *m = new (std::nothrow) lcpc_b200_multi;
  if (!m) return LCPC_B200_ERR_OOM;
  m->len = len;
  int rc = LCPC_B200_OK;
  for (size_t g = 0; g < n_gpus && rc == LCPC_B200_OK; g++) {
    lcpc_b200_shard *s = nullptr;
    rc = lcpc_b200_shard_new(encs[g], len, (unsigned)n_gpus, (unsigned)g, max_open, &s);
    if (rc == LCPC_B200_OK) m->shards.push_back(s);
  }
  std::vector<void *> wins(n_gpus, nullptr);
  for (size_t g = 0; g < m->shards.size() && rc == LCPC_B200_OK; g++) rc = lcpc_b200_shard_window(m->shards[g], &wins[g], nullptr, nullptr);
  for (size_t g = 0; g < m->shards.size() && rc == LCPC_B200_OK; g++) rc = lcpc_b200_shard_connect(m->shards[g], wins.data(), nullptr);
  if (rc == LCPC_B200_OK) rc = lcpc_b200_multi_rerun(m, coeffs_in, len);
  uint8_t root[32];
  for (size_t g = 0; g < m->shards.size() && rc == LCPC_B200_OK; g++) rc = lcpc_b200_shard_root(m->shards[g], root);  // completion + timeouts
  if (rc != LCPC_B200_OK) {
    lcpc_b200_multi_free(m);
    return rc;
  }
  *out = m;
  return LCPC_B200_OK;
}

int lcpc_b200_multi_root(lcpc_b200_multi *m, uint8_t root[32]) {
  if (!m || !root || m->shards.empty()) return LCPC_B200_ERR_BAD_ARG;
  // every shard holds the same root; draining all of them also makes the call a completion point for the commit
  int rc = LCPC_B200_OK;
  for (size_t g = m->shards.size(); g-- > 0 && rc == LCPC_B200_OK;) rc = lcpc_b200_shard_root(m->shards[g], root);
  return rc;
}

size_t lcpc_b200_multi_n_shards(const lcpc_b200_multi *m) { return m ? m->shards.size() : 0; }
lcpc_b200_shard *lcpc_b200_multi_shard(lcpc_b200_multi *m, size_t g) { return (m && g < m->shards.size()) ? m->shards[g] : nullptr; }

int lcpc_b200_multi_collapse(lcpc_b200_multi *m, const uint64_t *tensor, const uint8_t key[32], uint64_t *poly, uint8_t *repr) {
  if (!m || m->shards.empty() || !poly) return LCPC_B200_ERR_BAD_ARG;
  for (auto *s : m->shards)
    if (int rc = lcpc_b200_shard_collapse_begin(s, tensor, key)) return rc;
  int rc = lcpc_b200_shard_collapse_finish(m->shards[0], poly, repr);
  // the other shards' exchange areas were written too; mark their round as consumed (nothing to compute there)
  for (size_t g = 1; g < m->shards.size(); g++) m->shards[g]->collapse_done = m->shards[g]->collapse_seq;
  return rc;
}

int lcpc_b200_multi_open_columns(lcpc_b200_multi *m, const uint64_t *cols, size_t n, uint64_t *cols_out, uint8_t *paths_out) {
  if (!m || m->shards.empty()) return LCPC_B200_ERR_BAD_ARG;
  for (auto *s : m->shards)
    if (int rc = lcpc_b200_shard_open_begin(s, cols, n)) return rc;
  int rc = lcpc_b200_shard_open_finish(m->shards[0], cols_out, paths_out);
  for (size_t g = 1; g < m->shards.size(); g++) m->shards[g]->open_done = m->shards[g]->open_seq;
  return rc;
}

int lcpc_b200_multi_prove(lcpc_b200_multi *m, lcpc_b200_transcript *tr, const lcpc_b200_labels *labels,
                          const uint64_t *outer_tensor, size_t outer_len, size_t n_degree_tests, size_t n_col_opens,
                          uint64_t *p_eval, uint64_t *p_random, uint64_t *col_idx, uint64_t *cols_out, uint8_t *paths_out) {
  if (!m || m->shards.empty() || !tr || !outer_tensor || !p_eval || (n_degree_tests && !p_random) ||
      (n_col_opens && (!cols_out || !paths_out)))
    return LCPC_B200_ERR_BAD_ARG;
  lcpc_b200_shard *s0 = m->shards[0];
  lcpc_b200_ctx *ctx = s0->enc->ctx;
  const ShardPlan &p = s0->p;
  if (outer_len != p.n_rows) {
    std::lock_guard<std::mutex> g(ctx->mu);
    return fail(ctx, LCPC_B200_ERR_OUTER_TENSOR, "outer tensor has %zu entries, the commitment %zu rows", outer_len, p.n_rows);
  }
  const Labels lb = resolve_labels(labels);
  const size_t B = s0->B, L = B / 8, pbytes = p.n_per_row * B;
  uint8_t *stage = nullptr;
  std::vector<uint8_t> pageable;
  {
    std::lock_guard<std::mutex> g(ctx->mu);
    stage = (uint8_t *)host_stage(ctx, pbytes);
  }
  if (!stage) {
    pageable.resize(pbytes);
    stage = pageable.data();
  }
  for (size_t i = 0; i < n_degree_tests; i++) {
    uint8_t key[32];
    tr->tr.challenge_bytes(lb.dt, lb.dt_len, key, 32);
    if (int rc = lcpc_b200_multi_collapse(m, nullptr, key, p_random + i * p.n_per_row * L, stage)) return rc;
    tr->tr.append_elems(lb.pr, lb.pr_len, stage, B, p.n_per_row);
  }
  if (int rc = lcpc_b200_multi_collapse(m, outer_tensor, nullptr, p_eval, stage)) return rc;
  tr->tr.append_elems(lb.pe, lb.pe_len, stage, B, p.n_per_row);
  uint8_t key[32];
  tr->tr.challenge_bytes(lb.co, lb.co_len, key, 32);
  std::vector<uint64_t> cols(n_col_opens);
  sample_columns(key, p.n_cols, n_col_opens, cols.data());
  if (col_idx) memcpy(col_idx, cols.data(), n_col_opens * 8);
  return lcpc_b200_multi_open_columns(m, cols.data(), n_col_opens, cols_out, paths_out);
}

}  // extern "C"
