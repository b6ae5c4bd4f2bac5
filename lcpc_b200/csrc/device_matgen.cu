// lcpc_b200/csrc/device_matgen.cu -- the Brakedown code generator on the device.
//
// Replaces matgen::generate / gen_code (reference: lcpc-brakedown-pc/src/matgen.rs:28-52, :114-188) for encodings
// that are built from (n_per_row, seed): the matrices never exist on the host.  Per level i the reference draws from
// ONE sequential stream, ChaCha20Rng::seed_from_u64(seed) with set_stream(i) (:43-44): the precode first, then the
// postcode (:45-46).  For every input column gen_code draws
//     d distinct row indices    -- Uniform::new(0, m) samples, repeats thrown away (:144-159), then sorted (:160)
//     d non-zero field elements -- F::random, rejection-sampled (:171-177), one per sorted index
// so a column consumes a DATA-DEPENDENT number of 64-bit words (both kinds of draw are whole `next_u64`s) and column
// c+1 starts where column c ends.  That chain is resolved in parallel:
//   1. the keystream S[0..T) is generated up front (one ChaCha20 block per thread);
//   2. next[s] = where a column would end if it started at word s, for EVERY s (one thread per s, ~d + L d / acc reads);
//   3. the true column starts are the orbit of the matrix's first word under `next`: six pointer-doubling rounds give
//      next^64, one thread follows those long hops (n / 64 dependent loads), and a parallel pass fills in the 64
//      columns of every hop;
//   4. with its start known every column is drawn independently: indices sorted, values attached -- the reference's
//      CSC, kept as is (csc_idx / csc_data) -- and scattered into the row-compressed gather form the encoder reads
//      (row histogram, scan, scatter, per-row sort by column).
// Everything is integer work on u32/u64; results are bit-identical to the host generator (csrc/host_matgen.cpp),
// which the tests compare against matrix by matrix.
#include <algorithm>
#include <vector>

#include "../../include/lcpc_b200.h"
#include "expander_internal.h"
#include "field.cuh"
#include "kernels.h"

namespace lcpc {

namespace {

constexpr uint32_t SENT = 0xffffffffu;  // "runs past the generated keystream"
constexpr int MAX_D = 256;              // non-zeros per column the generator supports (reference presets stay below 64)
constexpr int LOG_HOP = 6;              // long hops of 2^6 columns

struct FieldParams {
  int limbs;            // u64 limbs
  uint64_t top_mask;    // NUM_BITS mask of the top limb
  uint64_t p[4];
};

FieldParams field_params(int field) {
  FieldParams f = {};
  const int n32 = field_limbs32(field);
  f.limbs = n32 / 2;
  uint32_t p32[8] = {0};
  unsigned bits = 0;
  switch (field) {
    case FT63: for (int i = 0; i < n32; i++) p32[i] = FieldP<FT63>::P(i); bits = 63; break;
    case FT127: for (int i = 0; i < n32; i++) p32[i] = FieldP<FT127>::P(i); bits = 127; break;
    case FT191: for (int i = 0; i < n32; i++) p32[i] = FieldP<FT191>::P(i); bits = 191; break;
    default: for (int i = 0; i < n32; i++) p32[i] = FieldP<FT255>::P(i); bits = 255; break;
  }
  for (int l = 0; l < f.limbs; l++) f.p[l] = (uint64_t)p32[2 * l] | ((uint64_t)p32[2 * l + 1] << 32);
  f.top_mask = ~(uint64_t)0 >> (64 * f.limbs - bits);
  return f;
}

// ---- 1. keystream: rand_chacha's ChaCha20Rng word order (block counter in words 12-13, stream id in 14-15) ----
__device__ __forceinline__ uint32_t rotl32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
#define LCPC_QR(a, b, c, d) \
  a += b, d = rotl32(d ^ a, 16), c += d, b = rotl32(b ^ c, 12), a += b, d = rotl32(d ^ a, 8), c += d, b = rotl32(b ^ c, 7)

struct Key8 { uint32_t w[8]; };

__global__ void __launch_bounds__(256)
matgen_keystream_kernel(Key8 key, uint64_t stream_id, size_t n_blocks, uint32_t *__restrict__ out) {
  const size_t blk = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (blk >= n_blocks) return;
  uint32_t in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
#pragma unroll
  for (int i = 0; i < 8; i++) in[4 + i] = key.w[i];
  in[12] = (uint32_t)blk, in[13] = (uint32_t)((uint64_t)blk >> 32);
  in[14] = (uint32_t)stream_id, in[15] = (uint32_t)(stream_id >> 32);
  uint32_t x[16];
#pragma unroll
  for (int i = 0; i < 16; i++) x[i] = in[i];
#pragma unroll 1
  for (int r = 0; r < 10; r++) {
    LCPC_QR(x[0], x[4], x[8], x[12]); LCPC_QR(x[1], x[5], x[9], x[13]);
    LCPC_QR(x[2], x[6], x[10], x[14]); LCPC_QR(x[3], x[7], x[11], x[15]);
    LCPC_QR(x[0], x[5], x[10], x[15]); LCPC_QR(x[1], x[6], x[11], x[12]);
    LCPC_QR(x[2], x[7], x[8], x[13]); LCPC_QR(x[3], x[4], x[9], x[14]);
  }
  uint4 *o = reinterpret_cast<uint4 *>(out + blk * 16);
#pragma unroll
  for (int q = 0; q < 4; q++)
    o[q] = make_uint4(x[4 * q] + in[4 * q], x[4 * q + 1] + in[4 * q + 1], x[4 * q + 2] + in[4 * q + 2], x[4 * q + 3] + in[4 * q + 3]);
}

// ---- one column of gen_code, replayed from word `s` of the keystream --------------------------------------------
// Returns the word after the column's last draw, or SENT if the keystream ends first.  EMIT: also hands out the
// picked row indices (in draw order) and, through `vals`, the d accepted elements (in draw order = sorted-index order).
template <bool EMIT>
__device__ __forceinline__ uint32_t replay_column(const uint64_t *__restrict__ S, uint32_t T, uint32_t s, uint32_t m, uint32_t d,
                                                  const FieldParams &f, uint32_t *picked, uint64_t *vals) {
  // Uniform::new(0usize, m): zone = u64::MAX - (2^64 - m) % m; v * m = (hi, lo); accept iff lo <= zone (rand 0.8)
  const uint64_t range = m;
  const uint64_t zone = ~(uint64_t)0 - ((0 - range) % range);
  uint32_t got = 0;
  uint32_t local[EMIT ? 1 : MAX_D];
  uint32_t *pk = EMIT ? picked : local;
  while (got < d) {
    if (s >= T) return SENT;
    const uint64_t v = S[s++];
    const uint64_t lo = v * range;
    if (lo > zone) continue;
    const uint32_t x = (uint32_t)__umul64hi(v, range);
    bool dup = false;
    for (uint32_t q = 0; q < got; q++) dup |= (pk[q] == x);  // `tmp.contains(&x)` (:150)
    if (!dup) pk[got++] = x;
  }
  // F::random (ff_derive): limbs from next_u64 in order, top limb masked, accept iff < p; zero is drawn again (:171-177)
  for (uint32_t k = 0; k < d; k++) {
    for (;;) {
      if (s + f.limbs > T) return SENT;
      uint64_t e[4];
      for (int l = 0; l < f.limbs; l++) e[l] = S[s + l];
      s += f.limbs;
      e[f.limbs - 1] &= f.top_mask;
      bool less = false, zero = true;
      for (int l = f.limbs - 1; l >= 0; l--) {
        if (e[l] != f.p[l]) {
          less = e[l] < f.p[l];
          break;
        }
      }
      for (int l = 0; l < f.limbs; l++) zero &= (e[l] == 0);
      if (less && !zero) {
        if (EMIT)
          for (int l = 0; l < f.limbs; l++) vals[(size_t)k * f.limbs + l] = e[l];
        break;
      }
    }
  }
  return s;
}

// ---- 2. next[s] for every word position ----
__global__ void __launch_bounds__(128)
matgen_next_kernel(const uint64_t *__restrict__ S, uint32_t T, uint32_t m, uint32_t d, FieldParams f, uint32_t *__restrict__ next) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= T) return;
  next[s] = replay_column<false>(S, T, s, m, d, f, nullptr, nullptr);
}

// ---- 3. orbit: pointer doubling, long hops, fill ----
__global__ void __launch_bounds__(256) matgen_double_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, uint32_t T) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= T) return;
  const uint32_t a = in[s];
  out[s] = (a == SENT || a >= T) ? SENT : in[a];
}

// one thread: hop[q] = start of column 64 q.  `first` is read from device memory (the precode's end, for a postcode)
__global__ void matgen_hops_kernel(const uint32_t *__restrict__ jump, const uint32_t *__restrict__ first, uint32_t n_hops,
                                   uint32_t T, uint32_t *__restrict__ hop, uint32_t *__restrict__ status) {
  if (threadIdx.x || blockIdx.x) return;
  uint32_t cur = *first;
  for (uint32_t q = 0; q < n_hops; q++) {
    hop[q] = cur;
    if (cur == SENT || cur >= T) {
      for (uint32_t r = q; r < n_hops; r++) hop[r] = SENT;
      atomicOr(status, 1u);
      return;
    }
    cur = jump[cur];
  }
}

// col_start[64 q + i] for i < 64; the entry for column n (one past the last) is the matrix's end word
__global__ void __launch_bounds__(128)
matgen_fill_kernel(const uint32_t *__restrict__ next, const uint32_t *__restrict__ hop, uint32_t n_hops, uint32_t n, uint32_t T,
                   uint32_t *__restrict__ col_start, uint32_t *__restrict__ status) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_hops) return;
  uint32_t cur = hop[q];
  for (uint32_t i = 0; i < (1u << LOG_HOP); i++) {
    const uint32_t c = (q << LOG_HOP) + i;
    if (c > n) break;
    col_start[c] = cur;
    if (c == n) break;
    if (cur == SENT || cur >= T) {
      atomicOr(status, 1u);
      cur = SENT;
    } else {
      cur = next[cur];
    }
  }
}

// ---- 4. draw every column; count the rows ----
__global__ void __launch_bounds__(64)
matgen_emit_kernel(const uint64_t *__restrict__ S, uint32_t T, const uint32_t *__restrict__ col_start, uint32_t n, uint32_t m,
                   uint32_t d, FieldParams f, uint32_t *__restrict__ csc_idx, uint64_t *__restrict__ csc_data,
                   uint32_t *__restrict__ row_count, uint32_t *__restrict__ status) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const uint32_t s = col_start[c];
  uint32_t *pk = csc_idx + (size_t)c * d;
  uint64_t *vals = csc_data + (size_t)c * d * f.limbs;
  if (s == SENT || replay_column<true>(S, T, s, m, d, f, pk, vals) == SENT) {
    atomicOr(status, 1u);
    return;
  }
  // tmp.sort_unstable() (:160): the values were drawn for the SORTED indices, so only the indices move
  for (uint32_t i = 1; i < d; i++) {
    const uint32_t v = pk[i];
    uint32_t j = i;
    while (j > 0 && pk[j - 1] > v) pk[j] = pk[j - 1], j--;
    pk[j] = v;
  }
  for (uint32_t k = 0; k < d; k++) atomicAdd(row_count + pk[k], 1u);
}

// exclusive scan of row_count[0..m) into rowptr[0..m], one CTA (m is at most a few hundred thousand)
__global__ void __launch_bounds__(1024) matgen_scan_kernel(const uint32_t *__restrict__ count, uint32_t m, uint32_t *__restrict__ rowptr) {
  __shared__ uint32_t part[1024];
  const uint32_t t = threadIdx.x, per = (m + 1023) / 1024;
  const uint32_t lo = min(t * per, m), hi = min(lo + per, m);
  uint32_t sum = 0;
  for (uint32_t i = lo; i < hi; i++) sum += count[i];
  part[t] = sum;
  __syncthreads();
  for (uint32_t off = 1; off < 1024; off <<= 1) {
    uint32_t v = t >= off ? part[t - off] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  uint32_t run = t ? part[t - 1] : 0;
  for (uint32_t i = lo; i < hi; i++) {
    rowptr[i] = run;
    run += count[i];
  }
  if (t == 1023) rowptr[m] = part[1023];
}

// scatter (column, source position) pairs into their rows, any order
__global__ void __launch_bounds__(256)
matgen_scatter_kernel(const uint32_t *__restrict__ csc_idx, size_t nnz, uint32_t d, const uint32_t *__restrict__ rowptr,
                      uint32_t *__restrict__ fill, uint32_t *__restrict__ colidx, uint32_t *__restrict__ src) {
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  const uint32_t row = csc_idx[k];
  const uint32_t pos = rowptr[row] + atomicAdd(fill + row, 1u);
  colidx[pos] = (uint32_t)(k / d);
  src[pos] = (uint32_t)k;
}

// ascending columns within every row (the encoder's column chunks cut rows by binary search), then move the values
__global__ void __launch_bounds__(128)
matgen_sort_rows_kernel(const uint32_t *__restrict__ rowptr, uint32_t m, uint32_t *__restrict__ colidx, uint32_t *__restrict__ src) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const uint32_t k0 = rowptr[i], k1 = rowptr[i + 1];
  for (uint32_t a = k0 + 1; a < k1; a++) {
    const uint32_t c = colidx[a], sidx = src[a];
    uint32_t b = a;
    while (b > k0 && colidx[b - 1] > c) colidx[b] = colidx[b - 1], src[b] = src[b - 1], b--;
    colidx[b] = c, src[b] = sidx;
  }
}

template <typename V>
__global__ void __launch_bounds__(256)
matgen_gather_vals_kernel(const V *__restrict__ csc_data, const uint32_t *__restrict__ src, size_t nnz, unsigned per_elem,
                          V *__restrict__ vals) {
  const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nnz * per_elem) return;
  const size_t k = g / per_elem, part = g % per_elem;
  vals[g] = csc_data[(size_t)src[k] * per_elem + part];
}

struct Scratch {
  std::vector<void *> ptrs;
  template <typename T> cudaError_t alloc(T **p, size_t count) {
    cudaError_t e = cudaMalloc(p, std::max<size_t>(count, 1) * sizeof(T));
    if (e == cudaSuccess) ptrs.push_back(*p);
    return e;
  }
  ~Scratch() {
    for (void *p : ptrs) cudaFree(p);
  }
};

unsigned grid_for(size_t items, unsigned block) { return (unsigned)((items + block - 1) / block); }

// expected 64-bit words per column, with headroom
size_t words_bound(const MatgenDims &dim, const FieldParams &f, double accept, double slack) {
  const double per_col = (double)dim.d * 1.02 + (double)dim.d * f.limbs / accept;
  return (size_t)((double)dim.n * per_col * slack) + 8192;
}

}  // namespace

int device_matgen(int field, uint64_t seed, size_t t, const MatgenDims *pre, const MatgenDims *post, cudaStream_t st,
                  ExpanderCode **out, std::string *err) {
  const FieldParams f = field_params(field);
  const size_t L = (size_t)f.limbs;
  // acceptance rate of F::random: p / 2^NUM_BITS
  const double accept = (double)f.p[L - 1] / ((double)f.top_mask + 1.0);
  Key8 key;
  {  // SeedableRng::seed_from_u64 (rand_core 0.6): PCG32 steps fill the 32-byte key
    uint64_t state = seed;
    for (auto &word : key.w) {
      state = state * 6364136223846793005ull + 11634580027462260723ull;
      uint32_t xs = (uint32_t)(((state >> 18) ^ state) >> 27);
      unsigned rot = (unsigned)(state >> 59);
      word = (xs >> rot) | (xs << ((32 - rot) & 31));
    }
  }
  ExpanderCode *c = new ExpanderCode;
  c->field = field;
  c->mats.resize(2 * t);
  auto bail = [&](int rc, const char *what, cudaError_t ce = cudaSuccess) {
    *err = what;
    if (ce != cudaSuccess) *err += std::string(": ") + cudaGetErrorString(ce);
    expander_free(c);
    return rc;
  };
  for (size_t lvl = 0; lvl < t; lvl++) {
    for (const MatgenDims &dim : {pre[lvl], post[lvl]}) {
      if (dim.d == 0 || dim.d > (size_t)MAX_D || dim.d > dim.m || dim.m >= 0x7fffffffull || dim.n >= 0x7fffffffull ||
          dim.n * dim.d >= 0xffffffffull)
        return bail(LCPC_B200_ERR_BAD_ARG, "matgen: level dimensions out of range");
    }
  }
  // scratch is sized once for the largest level (levels shrink geometrically) and reused: device allocations cost
  // more than the kernels here
  size_t max_n = 0, max_m = 0, max_nnz = 0;
  for (size_t lvl = 0; lvl < t; lvl++)
    for (const MatgenDims &dim : {pre[lvl], post[lvl]})
      max_n = std::max(max_n, dim.n), max_m = std::max(max_m, dim.m), max_nnz = std::max(max_nnz, dim.n * dim.d);
  double slack = 1.08;
  for (int attempt = 0;; attempt++, slack *= 1.5) {
    size_t max_T = 0;
    for (size_t lvl = 0; lvl < t; lvl++)
      max_T = std::max(max_T, words_bound(pre[lvl], f, accept, slack) + words_bound(post[lvl], f, accept, slack));
    if (max_T >= 0xfffffff0ull) return bail(LCPC_B200_ERR_TOO_BIG, "matgen: keystream too long for 32-bit word offsets");
    Scratch tmp;
    uint32_t *d_S = nullptr, *d_status = nullptr, *d_first = nullptr, *d_next = nullptr, *d_ja = nullptr, *d_jb = nullptr;
    uint32_t *d_hop = nullptr, *d_start = nullptr, *d_count = nullptr, *d_fill = nullptr, *d_src = nullptr;
    cudaError_t ce = tmp.alloc(&d_S, ((max_T + 7) / 8) * 16);
    if (ce == cudaSuccess) ce = tmp.alloc(&d_status, 1);
    if (ce == cudaSuccess) ce = tmp.alloc(&d_first, 1);
    if (ce == cudaSuccess) ce = tmp.alloc(&d_next, max_T);
    if (ce == cudaSuccess) ce = tmp.alloc(&d_ja, max_T);
    if (ce == cudaSuccess) ce = tmp.alloc(&d_jb, max_T);
    if (ce == cudaSuccess) ce = tmp.alloc(&d_hop, (max_n >> LOG_HOP) + 1);
    if (ce == cudaSuccess) ce = tmp.alloc(&d_start, max_n + 1);
    if (ce == cudaSuccess) ce = tmp.alloc(&d_count, max_m);
    if (ce == cudaSuccess) ce = tmp.alloc(&d_fill, max_m);
    if (ce == cudaSuccess) ce = tmp.alloc(&d_src, max_nnz);
    if (ce == cudaSuccess) ce = cudaMemsetAsync(d_status, 0, 4, st);
    if (ce != cudaSuccess) return bail(ce == cudaErrorMemoryAllocation ? LCPC_B200_ERR_OOM : LCPC_B200_ERR_CUDA, "matgen: scratch", ce);
    for (size_t lvl = 0; lvl < t; lvl++) {
      const size_t T = words_bound(pre[lvl], f, accept, slack) + words_bound(post[lvl], f, accept, slack);
      const size_t n_blocks = (T + 7) / 8;  // 16 u32 = 8 u64 per ChaCha20 block
      ce = cudaMemsetAsync(d_first, 0, 4, st);
      matgen_keystream_kernel<<<grid_for(n_blocks, 256), 256, 0, st>>>(key, (uint64_t)lvl, n_blocks, d_S);
      const uint64_t *S = reinterpret_cast<const uint64_t *>(d_S);
      const uint32_t T32 = (uint32_t)T;
      for (int which = 0; which < 2; which++) {
        const MatgenDims &dim = which ? post[lvl] : pre[lvl];
        DeviceCsr &M = c->mats[which ? t + lvl : lvl];
        const uint32_t n = (uint32_t)dim.n, m = (uint32_t)dim.m, d = (uint32_t)dim.d;
        const size_t nnz = (size_t)n * d;
        M.m = m, M.n = n, M.nnz = nnz, M.csc_d = d;
        // a retry regenerates every matrix from scratch
        cudaFree(M.rowptr), cudaFree(M.colidx), cudaFree(M.vals), cudaFree(M.csc_idx), cudaFree(M.csc_data);
        M.rowptr = M.colidx = M.vals = M.csc_idx = M.csc_data = nullptr;
        const uint32_t n_hops = (n >> LOG_HOP) + 1;
        if (ce == cudaSuccess) ce = cudaMalloc(&M.rowptr, ((size_t)m + 1) * 4);
        if (ce == cudaSuccess) ce = cudaMalloc(&M.colidx, std::max<size_t>(nnz, 1) * 4);
        if (ce == cudaSuccess) ce = cudaMalloc(&M.vals, std::max<size_t>(nnz, 1) * L * 8);
        if (ce == cudaSuccess) ce = cudaMalloc(&M.csc_idx, std::max<size_t>(nnz, 1) * 4);
        if (ce == cudaSuccess) ce = cudaMalloc(&M.csc_data, std::max<size_t>(nnz, 1) * L * 8);
        if (ce == cudaSuccess) ce = cudaMemsetAsync(d_count, 0, (size_t)m * 4, st);
        if (ce == cudaSuccess) ce = cudaMemsetAsync(d_fill, 0, (size_t)m * 4, st);
        if (ce != cudaSuccess) return bail(ce == cudaErrorMemoryAllocation ? LCPC_B200_ERR_OOM : LCPC_B200_ERR_CUDA, "matgen: buffers", ce);
        matgen_next_kernel<<<grid_for(T, 128), 128, 0, st>>>(S, T32, m, d, f, d_next);
        const uint32_t *jump = d_next;
        uint32_t *ping = d_ja, *pong = d_jb;
        for (int r = 0; r < LOG_HOP; r++) {
          matgen_double_kernel<<<grid_for(T, 256), 256, 0, st>>>(jump, ping, T32);
          jump = ping;
          std::swap(ping, pong);
        }
        matgen_hops_kernel<<<1, 1, 0, st>>>(jump, d_first, n_hops, T32, d_hop, d_status);
        matgen_fill_kernel<<<grid_for(n_hops, 128), 128, 0, st>>>(d_next, d_hop, n_hops, n, T32, d_start, d_status);
        // the postcode starts where the precode ended (the same rng is handed to both gen_code calls, :45-46)
        ce = cudaMemcpyAsync(d_first, d_start + n, 4, cudaMemcpyDeviceToDevice, st);
        matgen_emit_kernel<<<grid_for(n, 64), 64, 0, st>>>(S, T32, d_start, n, m, d, f, M.csc_idx, reinterpret_cast<uint64_t *>(M.csc_data),
                                                          d_count, d_status);
        matgen_scan_kernel<<<1, 1024, 0, st>>>(d_count, m, M.rowptr);
        matgen_scatter_kernel<<<grid_for(nnz, 256), 256, 0, st>>>(M.csc_idx, nnz, d, M.rowptr, d_fill, M.colidx, d_src);
        matgen_sort_rows_kernel<<<grid_for(m, 128), 128, 0, st>>>(M.rowptr, m, M.colidx, d_src);
        if (L % 2 == 0)
          matgen_gather_vals_kernel<uint4><<<grid_for(nnz * (L / 2), 256), 256, 0, st>>>(reinterpret_cast<const uint4 *>(M.csc_data), d_src,
                                                                                        nnz, (unsigned)(L / 2), reinterpret_cast<uint4 *>(M.vals));
        else
          matgen_gather_vals_kernel<uint2><<<grid_for(nnz * L, 256), 256, 0, st>>>(reinterpret_cast<const uint2 *>(M.csc_data), d_src, nnz,
                                                                                  (unsigned)L, reinterpret_cast<uint2 *>(M.vals));
        if (ce == cudaSuccess) ce = cudaGetLastError();
        if (ce != cudaSuccess) return bail(LCPC_B200_ERR_CUDA, "matgen: kernels", ce);
      }
    }
    // one look at the status word for the whole code: a keystream that was too short for some level's draw (an
    // unlucky run of rejections) makes everything be drawn again with more headroom
    uint32_t status = 0;
    ce = cudaMemcpyAsync(&status, d_status, 4, cudaMemcpyDeviceToHost, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    if (ce != cudaSuccess) return bail(LCPC_B200_ERR_CUDA, "matgen: kernels", ce);
    if (!status) break;
    if (attempt >= 4) return bail(LCPC_B200_ERR_CUDA, "matgen: keystream bound exceeded repeatedly");
  }
  int rc = expander_assemble(c, t, err);
  if (rc != LCPC_B200_OK) {
    expander_free(c);
    return rc;
  }
  *out = c;
  return LCPC_B200_OK;
}

}  // namespace lcpc
