// lcpc_b200/csrc/kernels_ntt.cu -- Ligero row encoding: batched NTT (radix-2 DIF semantics, radix-8 register rounds).
//
// Replaces LigeroEncoding::encode (reference: lcpc-ligero-pc/src/lib.rs:162-164), which delegates to
// fffft 0.4 `fft_io_pc`: in-order input, bit-reversed output, decimation in frequency,
//   for gap = n/2 .. 1:  (a, b) <- (a + b, (a - b) * w^(idx * n/(2 gap))),  idx = position mod gap.
// Field arithmetic is exact, so any schedule of those butterflies gives bit-identical limbs.
//
// Schedule: the log2(n) stages are cut into passes of S <= 10 consecutive stages.  One CTA owns a tile
// of 2^S points along the pass's stride times C adjacent points (C*B contiguous bytes in HBM), stages
// it in shared memory, runs the S stages there in radix-8 register rounds and writes it back: one HBM
// read + one write per pass.  The first pass reads the un-padded coefficient rows directly (implicit zeros
// beyond n_per_row) and can store the commit's copy of them on the way, so the reference's separate
// pad/copy (lcpc-2d/src/lib.rs:640-651) costs no extra read; the last pass can store per column block
// (multi-GPU exchange fused into the transform).
#include <algorithm>

#include <type_traits>

#include "field.cuh"
#include "kernels.h"

namespace lcpc {

int field_limbs32(int field) {
  switch (field) {
    case FT63: return 2;
    case FT127: return 4;
    case FT191: return 6;
    case FT255: return 8;
    default: return -1;
  }
}

// ---- global / shared element movement -------------------------------------------------------
// Global: N consecutive limbs.  Shared: split into planes of PW bytes so that consecutive elements
// are consecutive within a plane (conflict-free 64/128-bit accesses for a warp on adjacent elements).
template <int N> struct Planes {
  static constexpr int PW = (N % 4 == 0) ? 4 : 2;  // limbs per plane word (16 B or 8 B)
  static constexpr int NP = N / PW;
};

template <int N>
__device__ __forceinline__ void gload(uint32_t (&v)[N], const uint32_t *p) {
  if constexpr (N % 4 == 0) {
#pragma unroll
    for (int i = 0; i < N / 4; i++) {
      uint4 t = reinterpret_cast<const uint4 *>(p)[i];
      v[4 * i] = t.x, v[4 * i + 1] = t.y, v[4 * i + 2] = t.z, v[4 * i + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < N / 2; i++) {
      uint2 t = reinterpret_cast<const uint2 *>(p)[i];
      v[2 * i] = t.x, v[2 * i + 1] = t.y;
    }
  }
}

template <int N>
__device__ __forceinline__ void gload_ro(uint32_t (&v)[N], const uint32_t *p) {
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

template <int N>
__device__ __forceinline__ void gstore(uint32_t *p, const uint32_t (&v)[N]) {
  if constexpr (N % 4 == 0) {
#pragma unroll
    for (int i = 0; i < N / 4; i++)
      reinterpret_cast<uint4 *>(p)[i] = make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < N / 2; i++) reinterpret_cast<uint2 *>(p)[i] = make_uint2(v[2 * i], v[2 * i + 1]);
  }
}

// ---- element-wise test hook -------------------------------------------------------------------
template <int FID>
__global__ void field_op_kernel(int op, uint32_t *r, const uint32_t *a, const uint32_t *b, size_t n) {
  using F = Field<FID>;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  typename F::Elem x, y, z;
  gload<F::N>(x.v, a + i * F::N);
  if (b) gload<F::N>(y.v, b + i * F::N);
  else y = F::zero();
  switch (op) {
    case 0: z = F::add(x, y); break;
    case 1: z = F::sub(x, y); break;
    case 2: z = F::mul(x, y); break;
    case 5: z = F::template redc<1>(F::mul_full(x, y)); break;
    case 6: {  // 37-term lazily reduced sum of products starting at element i (indices wrap)
      typename F::Wide acc = F::wide_zero();
      for (size_t k = 0; k < 37; k++) {
        typename F::Elem u, v;
        gload<F::N>(u.v, a + ((i + k) % n) * F::N);
        gload<F::N>(v.v, b + ((i * 7 + k) % n) * F::N);
        F::mac_wide(acc, u, v);
      }
      z = F::template redc<2>(acc);
      break;
    }
    case 7:
      if constexpr (F::N % 4 == 0) z = F::mul_karatsuba(x, y);
      else z = F::mul(x, y);
      break;
    default: z = F::from_mont(x); break;
  }
  gstore<F::N>(r + i * F::N, z.v);
}

cudaError_t launch_field_op(int field, int op, uint32_t *r, const uint32_t *a, const uint32_t *b, size_t n,
                            cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  unsigned grid = (unsigned)((n + 127) / 128);
  switch (field) {
    case FT63: field_op_kernel<FT63><<<grid, 128, 0, stream>>>(op, r, a, b, n); break;
    case FT127: field_op_kernel<FT127><<<grid, 128, 0, stream>>>(op, r, a, b, n); break;
    case FT191: field_op_kernel<FT191><<<grid, 128, 0, stream>>>(op, r, a, b, n); break;
    case FT255: field_op_kernel<FT255><<<grid, 128, 0, stream>>>(op, r, a, b, n); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

// ---- root table: roots[i] = w^i -----------------------------------------------------------------
// Each thread owns a run of consecutive exponents: w^(start) by square-and-multiply, then steps by w.
template <int FID>
__global__ void root_table_kernel(uint32_t *roots, const uint32_t *w_in, size_t half, unsigned run) {
  using F = Field<FID>;
  size_t start = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * run;
  if (start >= half) return;
  typename F::Elem w, cur, base;
  gload<F::N>(w.v, w_in);
  // Montgomery one = R mod p = from_mont^{-1}(1): obtain it as w^0 via x * R2 ... avoid constants:
  // one = (2^(32N) mod p); computed on the host and passed as w_in[N..2N)
  gload<F::N>(cur.v, w_in + F::N);
  base = w;
  for (size_t e = start; e; e >>= 1) {
    if (e & 1) cur = F::mul(cur, base);
    base = F::mul(base, base);
  }
  size_t end = start + run < half ? start + run : half;
  for (size_t i = start; i < end; i++) {
    gstore<F::N>(roots + i * F::N, cur.v);
    cur = F::mul(cur, w);
  }
}

cudaError_t launch_root_table(int field, uint32_t *roots, const uint32_t *w, size_t half, cudaStream_t stream) {
  if (half == 0) return cudaSuccess;
  const unsigned run = 16;
  size_t threads = (half + run - 1) / run;
  unsigned grid = (unsigned)((threads + 127) / 128);
  switch (field) {
    case FT63: root_table_kernel<FT63><<<grid, 128, 0, stream>>>(roots, w, half, run); break;
    case FT127: root_table_kernel<FT127><<<grid, 128, 0, stream>>>(roots, w, half, run); break;
    case FT191: root_table_kernel<FT191><<<grid, 128, 0, stream>>>(roots, w, half, run); break;
    case FT255: root_table_kernel<FT255><<<grid, 128, 0, stream>>>(roots, w, half, run); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

// ---- one NTT pass -----------------------------------------------------------------------------
// Geometry of a pass (host side fills NttPass): the pass runs S consecutive stages, the ones whose
// butterfly gap in row positions is 2^(hi-1) .. 2^(hi-S).  A CTA stages a tile of 2^S "t" values (the
// S row-index bits the pass works on) times C = 2^logC other positions:
//   stride pass (hi > S):  element (t, c) is row position  base + t * 2^(hi-S) + c,  c = C adjacent points
//   last pass   (hi == S): element (t, c) is row position  base + c * 2^S + t,       C adjacent sub-transforms
// so a tile is always made of runs of >= C*B (resp. 2^S*B) contiguous bytes in HBM.
//
// Inside the tile the S stages run as ceil(S/3) ROUNDS: a thread pulls 8 (4, 2) elements that differ in
// the 3 (2, 1) t-bits of the round into registers, does those stages there, and puts them back -- one
// shared-memory round trip and one barrier per three stages, and up to four independent Montgomery
// products in flight per thread to keep the IMAD pipe fed.  In the final round of the last pass the
// twiddle exponent of lane r of every stage is r * n/(2 gap): the r = 0 products are by w^0 = 1 and are
// dropped at compile time (5 products per 8 points instead of 12).
struct NttPass {
  unsigned log_n;      // transform length 2^log_n
  unsigned hi;         // this pass runs the stages with gaps 2^(hi-1) .. 2^(hi-S)
  unsigned S;          // stages in this pass
  unsigned logC;       // log2 of adjacent points (or sub-transforms) per tile
  unsigned tiles_per_row_log;  // log2(n / (2^S * C))
  size_t src_stride, src_valid, dst_stride;
  Scatter sc;          // last pass only: store straight into per-column-block (possibly peer) matrices
  uint32_t *copy_dst;  // first pass only (optional): every source element read is also stored here, rows
  size_t copy_stride;  // copy_stride apart -- commit()'s own copy of the coefficients for free
};

// CTA shape, measured at 2^24 (encode ms): 256 threads x 2048-element tile x 2 CTAs/SM 4.97; 128 x 1024 x 4
// 4.75; 64 x 512 x 8 4.78; 512 x 4096 x 1 5.89 -- small CTAs hide each other's staging phases and barriers
#ifndef LCPC_NTT_THREADS
#define LCPC_NTT_THREADS 128
#endif
#ifndef LCPC_NTT_LOG_TILE
#define LCPC_NTT_LOG_TILE 10
#endif
constexpr int NTT_THREADS = LCPC_NTT_THREADS;
// tuning knobs (A/B builds): largest register round (3 = radix-8) and CTAs per SM the register budget allows
#ifndef LCPC_NTT_MAX_KL
#define LCPC_NTT_MAX_KL 3
#endif
#ifndef LCPC_NTT_MIN_BLOCKS
#define LCPC_NTT_MIN_BLOCKS 4
#endif

struct NttGeom {
  unsigned log_n, S, tile, tshift, cshift, tmask, cmask, log_stride;
  bool last;
  size_t base;
};

// shared-memory slot of tile element e: XOR the 16-byte-bank index with the next three index bits, so
// that the 8 lanes of a quarter warp hit 8 different bank groups in every round (their e differ either
// in bits 0-2 or in bits 3-5)
__device__ __forceinline__ unsigned swz(unsigned e) { return e ^ ((e >> 3) & 7u); }

template <int N> struct SmemLayout {
  static constexpr int PW = Planes<N>::PW, NP = Planes<N>::NP;
  // planes are PLANE_PAD elements apart beyond the tile so that the two 16-byte halves of an element
  // land in different bank groups (staging writes both halves of 16 elements per warp instruction)
  static constexpr unsigned PLANE_PAD = 4;
  __host__ __device__ static size_t bytes(unsigned tile) { return (size_t)NP * (tile + PLANE_PAD) * PW * 4; }
};

template <int N>
__device__ __forceinline__ void sload2(uint32_t (&v)[N], const uint32_t *smem, unsigned slot, unsigned tile) {
  constexpr int PW = SmemLayout<N>::PW, NP = SmemLayout<N>::NP;
#pragma unroll
  for (int pl = 0; pl < NP; pl++) {
    const uint32_t *q = smem + ((size_t)pl * (tile + SmemLayout<N>::PLANE_PAD) + slot) * PW;
    if constexpr (PW == 4) {
      uint4 t = *reinterpret_cast<const uint4 *>(q);
      v[4 * pl] = t.x, v[4 * pl + 1] = t.y, v[4 * pl + 2] = t.z, v[4 * pl + 3] = t.w;
    } else {
      uint2 t = *reinterpret_cast<const uint2 *>(q);
      v[2 * pl] = t.x, v[2 * pl + 1] = t.y;
    }
  }
}

template <int N>
__device__ __forceinline__ void sstore2(uint32_t *smem, unsigned slot, unsigned tile, const uint32_t (&v)[N]) {
  constexpr int PW = SmemLayout<N>::PW, NP = SmemLayout<N>::NP;
#pragma unroll
  for (int pl = 0; pl < NP; pl++) {
    uint32_t *q = smem + ((size_t)pl * (tile + SmemLayout<N>::PLANE_PAD) + slot) * PW;
    if constexpr (PW == 4)
      *reinterpret_cast<uint4 *>(q) = make_uint4(v[4 * pl], v[4 * pl + 1], v[4 * pl + 2], v[4 * pl + 3]);
    else
      *reinterpret_cast<uint2 *>(q) = make_uint2(v[2 * pl], v[2 * pl + 1]);
  }
}

// KL stages (t-bits lg_top .. lg_top-KL+1) on 2^KL elements per work item, in registers
// `tws` (last pass only): the pass's 2^(S-1) twiddles w^(k * n / 2^S) in shared memory, element k at tws + k*N
template <int FID, int KL, bool FINAL, bool SMEM_TW>
__device__ __forceinline__ void ntt_round(uint32_t *smem, const uint32_t *__restrict__ roots, const NttGeom &g,
                                          unsigned lg_top, const uint32_t *tws) {
  using F = Field<FID>;
  constexpr int N = F::N;
  constexpr int K = 1 << KL;
  const unsigned pos = lg_top + 1 - KL + g.tshift;  // tile-index bit of the round's lowest t-bit
  const unsigned items = g.tile >> KL;
  const unsigned log_h = g.log_stride + lg_top + 1 - KL;  // log2 row distance between the item's elements
  for (unsigned w = threadIdx.x; w < items; w += NTT_THREADS) {
    const unsigned e0 = ((w >> pos) << (pos + KL)) | (w & ((1u << pos) - 1u));
    typename F::Elem x[K];
#pragma unroll
    for (int q = 0; q < K; q++) sload2<N>(x[q].v, smem, swz(e0 | ((unsigned)q << pos)), g.tile);
    const unsigned t0 = (e0 >> g.tshift) & g.tmask, c = (e0 >> g.cshift) & g.cmask;
    const size_t j0 = g.last ? g.base + ((size_t)c << g.S) + t0 : g.base + ((size_t)t0 << g.log_stride) + c;
#pragma unroll
    for (int u = 0; u < KL; u++) {
      constexpr int dummy = 0;
      (void)dummy;
      const int gq = K >> (u + 1);
      const unsigned logG = g.log_stride + lg_top - u;  // log2 of this stage's gap in row positions
      const unsigned twshift = g.log_n - 1 - logG;
#pragma unroll
      for (int r = 0; r < gq; r++) {
        const bool unit = FINAL && r == 0;  // exponent r * n/(2 gap) = 0
        typename F::Elem wv;
        if (!unit) {
          const size_t k = (j0 + ((size_t)r << log_h)) & (((size_t)1 << logG) - 1);
          if (SMEM_TW) {  // exponent k * n / 2^(logG+1) = (k << (S-1-logG)) * n / 2^S
            const uint32_t *q = tws + ((k << (g.S - 1 - logG)) * N);
#pragma unroll
            for (int l = 0; l < N; l += Planes<N>::PW) {
              if constexpr (Planes<N>::PW == 4) {
                uint4 t = *reinterpret_cast<const uint4 *>(q + l);
                wv.v[l] = t.x, wv.v[l + 1] = t.y, wv.v[l + 2] = t.z, wv.v[l + 3] = t.w;
              } else {
                uint2 t = *reinterpret_cast<const uint2 *>(q + l);
                wv.v[l] = t.x, wv.v[l + 1] = t.y;
              }
            }
          } else {
            gload_ro<N>(wv.v, roots + (k << twshift) * N);
          }
        }
#pragma unroll
        for (int blk = 0; blk < K; blk += 2 * gq) {
          const int qa = blk + r, qb = qa + gq;
          typename F::Elem sum = F::add(x[qa], x[qb]);
          typename F::Elem dif = F::sub(x[qa], x[qb]);
          x[qa] = sum;
          x[qb] = unit ? dif : F::mul(dif, wv);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < K; q++) sstore2<N>(smem, swz(e0 | ((unsigned)q << pos)), g.tile, x[q].v);
  }
}

#ifndef LCPC_NTT_LOAD_BATCH
#define LCPC_NTT_LOAD_BATCH 4
#endif
#ifndef LCPC_NTT_STORE_BATCH
#define LCPC_NTT_STORE_BATCH 0
#endif

template <int FID>
__global__ void __launch_bounds__(NTT_THREADS, LCPC_NTT_MIN_BLOCKS)
ntt_pass_kernel(const uint32_t *__restrict__ src, uint32_t *dst, const uint32_t *__restrict__ roots, NttPass p) {
  using F = Field<FID>;
  constexpr int N = F::N;
  constexpr int PW = SmemLayout<N>::PW, NP = SmemLayout<N>::NP;
  extern __shared__ __align__(16) uint32_t smem[];
  NttGeom g;
  g.log_n = p.log_n, g.S = p.S;
  g.tile = 1u << (p.S + p.logC);
  g.last = (p.hi == p.S);
  const size_t row = blockIdx.x >> p.tiles_per_row_log;
  const size_t tid_in_row = blockIdx.x & ((1u << p.tiles_per_row_log) - 1);
  if (g.last) {
    g.base = tid_in_row << (p.S + p.logC);
    g.log_stride = 0;
    g.tshift = 0, g.cshift = p.S;
  } else {
    g.log_stride = p.hi - p.S;
    const size_t lowblocks_log = g.log_stride - p.logC;
    const size_t high = tid_in_row >> lowblocks_log;
    const size_t lowblock = tid_in_row & (((size_t)1 << lowblocks_log) - 1);
    g.base = (high << p.hi) + (lowblock << p.logC);
    g.tshift = p.logC, g.cshift = 0;
  }
  g.tmask = (1u << p.S) - 1, g.cmask = (1u << p.logC) - 1;
  const uint32_t *srow = src + row * p.src_stride * N;
  uint32_t *drow = dst + row * p.dst_stride * N;
  const unsigned granules = g.tile * NP;  // PW-limb pieces; consecutive lanes move consecutive pieces of HBM

  // Staging: LCPC_NTT_LOAD_BATCH granules per thread are requested from HBM before the first one is stored to shared
  // memory (a thread moves 16 granules of a 1024-element Ft255 tile; one at a time, or two as ptxas unrolled the plain
  // loop, each round trip to HBM was exposed: the per-instruction stall samples of the --set full capture put 9 % of
  // the kernel there).  0 = the plain loop.  Measured at 2^24 / Ft255 (profiles/r02_ab_ntt_load_batch.jsonl): encode 4.759 ms
  // plain, 4.710 / 4.657 / 4.724 / 4.700 / 4.740 ms with 2 / 4 / 6 / 8 / 16 in flight (more registers held across the phase cost what the
  // deeper queue gains); 4 is the default.
#if LCPC_NTT_LOAD_BATCH > 0
  for (unsigned g0 = threadIdx.x; g0 < granules; g0 += NTT_THREADS * LCPC_NTT_LOAD_BATCH) {
    using V = typename std::conditional<PW == 4, uint4, uint2>::type;
    V v[LCPC_NTT_LOAD_BATCH];
#pragma unroll
    for (int u = 0; u < LCPC_NTT_LOAD_BATCH; u++) {
      const unsigned gi = g0 + u * NTT_THREADS;
      const unsigned e = gi / NP, pl = gi % NP;
      const unsigned t = (e >> g.tshift) & g.tmask, c = (e >> g.cshift) & g.cmask;
      const size_t j = g.last ? g.base + e : g.base + ((size_t)t << g.log_stride) + c;
      if constexpr (PW == 4) v[u] = make_uint4(0, 0, 0, 0);
      else v[u] = make_uint2(0, 0);
      if (gi < granules && j < p.src_valid) v[u] = __ldg(reinterpret_cast<const V *>(srow + j * N) + pl);
    }
#pragma unroll
    for (int u = 0; u < LCPC_NTT_LOAD_BATCH; u++) {
      const unsigned gi = g0 + u * NTT_THREADS;
      if (gi >= granules) break;
      const unsigned e = gi / NP, pl = gi % NP;
      if (p.copy_dst) {
        const unsigned t = (e >> g.tshift) & g.tmask, c = (e >> g.cshift) & g.cmask;
        const size_t j = g.last ? g.base + e : g.base + ((size_t)t << g.log_stride) + c;
        if (j < p.src_valid) reinterpret_cast<V *>(p.copy_dst + (row * p.copy_stride + j) * N)[pl] = v[u];
      }
      *reinterpret_cast<V *>(smem + ((size_t)pl * (g.tile + SmemLayout<N>::PLANE_PAD) + swz(e)) * PW) = v[u];
    }
  }
#else
  for (unsigned gi = threadIdx.x; gi < granules; gi += NTT_THREADS) {
    const unsigned e = gi / NP, pl = gi % NP;
    const unsigned t = (e >> g.tshift) & g.tmask, c = (e >> g.cshift) & g.cmask;
    const size_t j = g.last ? g.base + e : g.base + ((size_t)t << g.log_stride) + c;
    uint32_t *q = smem + ((size_t)pl * (g.tile + SmemLayout<N>::PLANE_PAD) + swz(e)) * PW;
    if constexpr (PW == 4) {
      uint4 v = make_uint4(0, 0, 0, 0);
      if (j < p.src_valid) {
        v = __ldg(reinterpret_cast<const uint4 *>(srow + j * N) + pl);
        if (p.copy_dst) reinterpret_cast<uint4 *>(p.copy_dst + (row * p.copy_stride + j) * N)[pl] = v;
      }
      *reinterpret_cast<uint4 *>(q) = v;
    } else {
      uint2 v = make_uint2(0, 0);
      if (j < p.src_valid) {
        v = __ldg(reinterpret_cast<const uint2 *>(srow + j * N) + pl);
        if (p.copy_dst) reinterpret_cast<uint2 *>(p.copy_dst + (row * p.copy_stride + j) * N)[pl] = v;
      }
      *reinterpret_cast<uint2 *>(q) = v;
    }
  }
#endif
  // last pass: its stages only ever use the 2^(S-1) twiddles w^(k n / 2^S); keep them in shared memory behind the
  // tile instead of paying an L2 round trip per butterfly
  uint32_t *tws = smem + SmemLayout<N>::bytes(g.tile) / 4;
  if (g.last) {
    const unsigned n_tw = p.S ? 1u << (p.S - 1) : 0;
    const unsigned step_log = p.log_n - p.S;
    for (unsigned gi = threadIdx.x; gi < n_tw * NP; gi += NTT_THREADS) {
      const unsigned k = gi / NP, pl = gi % NP;
      const uint32_t *src_tw = roots + ((size_t)k << step_log) * N + pl * PW;
      if constexpr (PW == 4) *reinterpret_cast<uint4 *>(tws + k * N + pl * PW) = __ldg(reinterpret_cast<const uint4 *>(src_tw));
      else *reinterpret_cast<uint2 *>(tws + k * N + pl * PW) = __ldg(reinterpret_cast<const uint2 *>(src_tw));
    }
  }
  __syncthreads();

  unsigned rem = p.S, lg_top = p.S - 1;
  while (rem > 0) {
    if (LCPC_NTT_MAX_KL >= 3 && rem >= 3) {
      if (g.last && rem == 3) ntt_round<FID, 3, true, true>(smem, roots, g, lg_top, tws);
      else if (g.last) ntt_round<FID, 3, false, true>(smem, roots, g, lg_top, tws);
      else ntt_round<FID, 3, false, false>(smem, roots, g, lg_top, tws);
      rem -= 3, lg_top -= 3;
    } else if (rem >= 2) {
      if (g.last && rem == 2) ntt_round<FID, 2, true, true>(smem, roots, g, lg_top, tws);
      else if (g.last) ntt_round<FID, 2, false, true>(smem, roots, g, lg_top, tws);
      else ntt_round<FID, 2, false, false>(smem, roots, g, lg_top, tws);
      rem -= 2, lg_top -= 2;
    } else {
      if (g.last) ntt_round<FID, 1, true, true>(smem, roots, g, lg_top, tws);
      else ntt_round<FID, 1, false, false>(smem, roots, g, lg_top, tws);
      rem = 0;
    }
    __syncthreads();
  }

  // (the store side the same way -- LCPC_NTT_STORE_BATCH granules read from shared memory before the first store --
  // measured 4.680 / 4.682 ms for 4 / 8 against 4.657 ms for the plain loop: off)
#if LCPC_NTT_STORE_BATCH > 0
  for (unsigned g0 = threadIdx.x; g0 < granules; g0 += NTT_THREADS * LCPC_NTT_STORE_BATCH) {
    using V = typename std::conditional<PW == 4, uint4, uint2>::type;
    V v[LCPC_NTT_STORE_BATCH];
#pragma unroll
    for (int u = 0; u < LCPC_NTT_STORE_BATCH; u++) {
      const unsigned gi = min(g0 + u * NTT_THREADS, granules - 1);
      const unsigned e = gi / NP, pl = gi % NP;
      v[u] = *reinterpret_cast<const V *>(smem + ((size_t)pl * (g.tile + SmemLayout<N>::PLANE_PAD) + swz(e)) * PW);
    }
#pragma unroll
    for (int u = 0; u < LCPC_NTT_STORE_BATCH; u++) {
      const unsigned gi = g0 + u * NTT_THREADS;
      if (gi >= granules) break;
      const unsigned e = gi / NP, pl = gi % NP;
      const unsigned t = (e >> g.tshift) & g.tmask, c = (e >> g.cshift) & g.cmask;
      const size_t j = g.last ? g.base + e : g.base + ((size_t)t << g.log_stride) + c;
      uint32_t *out = drow + j * N;
      if (p.sc.n_blocks) {
        unsigned h = 0;
        while (h + 1 < p.sc.n_blocks && j >= p.sc.starts[h + 1]) h++;
        const size_t start = p.sc.starts[h], width = p.sc.starts[h + 1] - start;
        out = p.sc.dst[h] + ((p.sc.row0 + row) * width + (j - start)) * N;
      }
      reinterpret_cast<V *>(out)[pl] = v[u];
    }
  }
#else
  for (unsigned gi = threadIdx.x; gi < granules; gi += NTT_THREADS) {
    const unsigned e = gi / NP, pl = gi % NP;
    const unsigned t = (e >> g.tshift) & g.tmask, c = (e >> g.cshift) & g.cmask;
    const size_t j = g.last ? g.base + e : g.base + ((size_t)t << g.log_stride) + c;
    const uint32_t *q = smem + ((size_t)pl * (g.tile + SmemLayout<N>::PLANE_PAD) + swz(e)) * PW;
    uint32_t *out = drow + j * N;
    if (p.sc.n_blocks) {  // column block of position j, then (global row, local column) inside that block's matrix
      unsigned h = 0;
      while (h + 1 < p.sc.n_blocks && j >= p.sc.starts[h + 1]) h++;
      const size_t start = p.sc.starts[h], width = p.sc.starts[h + 1] - start;
      out = p.sc.dst[h] + ((p.sc.row0 + row) * width + (j - start)) * N;
    }
    if constexpr (PW == 4) reinterpret_cast<uint4 *>(out)[pl] = *reinterpret_cast<const uint4 *>(q);
    else reinterpret_cast<uint2 *>(out)[pl] = *reinterpret_cast<const uint2 *>(q);
  }
#endif
}

template <int FID>
static cudaError_t ntt_rows_impl(const uint32_t *src, size_t src_stride, size_t src_valid, uint32_t *dst,
                                 size_t dst_stride, const uint32_t *roots, unsigned log_n, size_t n_rows,
                                 cudaStream_t stream, int *n_launches, const Scatter *scatter, uint32_t *copy_dst,
                                 size_t copy_stride) {
  using F = Field<FID>;
  static bool attr_set_dev[64] = {};  // the attribute is per device: one process may hold contexts on several
  int dev_id = 0;
  cudaGetDevice(&dev_id);
  bool &attr_set = attr_set_dev[dev_id & 63];
  constexpr unsigned LOG_TILE = LCPC_NTT_LOG_TILE;  // 1024 elements: 32 KiB for Ft255
  constexpr unsigned MAX_S = 10;  // stages per pass: 2^19 points are 10 + 9 (two HBM round trips), 2^17 are 9 + 8
  if (!attr_set) {
    // tile + (last pass) up to 2^(LOG_TILE-1) twiddles
    cudaError_t e = cudaFuncSetAttribute(ntt_pass_kernel<FID>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(SmemLayout<F::N>::bytes(1u << LOG_TILE) + ((size_t)1 << (LOG_TILE - 1)) * F::BYTES +
                                               (64u << 10)));
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  if (n_launches) *n_launches = 0;
  if (n_rows == 0) return cudaSuccess;
  if (log_n == 0) {  // length-1 transform: identity
    if ((scatter && scatter->n_blocks) || copy_dst) return cudaErrorInvalidValue;
    if (src != dst)
      return cudaMemcpy2DAsync(dst, dst_stride * F::BYTES, src, src_stride * F::BYTES, F::BYTES, n_rows,
                               cudaMemcpyDeviceToDevice, stream);
    return cudaSuccess;
  }
  unsigned n_pass;
  if (log_n <= LOG_TILE) n_pass = 1;
  else n_pass = (log_n + MAX_S - 1) / MAX_S;
  unsigned hi = log_n;
  const uint32_t *cur_src = src;
  size_t cur_stride = src_stride, cur_valid = src_valid;
  for (unsigned ip = 0; ip < n_pass; ip++) {
    unsigned remaining = n_pass - ip;
    // balanced split, larger passes first (2^17: 9 + 8 measured 4.75 ms; 8 + 9, which would save 1.6 % of the
    // products in the final round, measured 4.78 ms)
    unsigned S = (hi + remaining - 1) / remaining;
    NttPass p;
    p.log_n = log_n, p.hi = hi, p.S = S;
    unsigned logC = S >= LOG_TILE ? 0 : LOG_TILE - S;
    bool last = (hi == S);
    unsigned avail = last ? log_n - S : hi - S;  // log2 of adjacent points (or sub-transforms) available
    if (logC > avail) logC = avail;
    p.logC = logC;
    p.tiles_per_row_log = log_n - S - logC;
    p.src_stride = cur_stride, p.src_valid = cur_valid, p.dst_stride = dst_stride;
    p.sc.n_blocks = 0;
    if (last && scatter && scatter->n_blocks) p.sc = *scatter;
    p.copy_dst = ip == 0 ? copy_dst : nullptr, p.copy_stride = copy_stride;
    size_t grid = n_rows << p.tiles_per_row_log;
    if (grid > 0x7fffffffu) return cudaErrorInvalidValue;
    size_t smem = SmemLayout<F::N>::bytes(1u << (S + logC)) + (last ? ((size_t)1 << (S - 1)) * F::BYTES : 0);
    // A/B knob: unused shared memory that lowers the CTAs per SM (e.g. to leave registers for a co-resident hash kernel)
    smem += (size_t)std::min<long>(64, std::max<long>(0, tunable("NTT_SMEM_PAD_KB", 0))) << 10;
    ntt_pass_kernel<FID><<<(unsigned)grid, NTT_THREADS, smem, stream>>>(cur_src, dst, roots, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (n_launches) ++*n_launches;
    cur_src = dst, cur_stride = dst_stride, cur_valid = (size_t)1 << log_n;
    hi -= S;
  }
  return cudaSuccess;
}

cudaError_t launch_ntt_rows(int field, const uint32_t *src, size_t src_stride, size_t src_valid, uint32_t *dst,
                            size_t dst_stride, const uint32_t *roots, unsigned log_n, size_t n_rows,
                            cudaStream_t stream, int *n_launches, const Scatter *scatter, uint32_t *copy_dst,
                            size_t copy_stride) {
  switch (field) {
    case FT63: return ntt_rows_impl<FT63>(src, src_stride, src_valid, dst, dst_stride, roots, log_n, n_rows, stream, n_launches, scatter, copy_dst, copy_stride);
    case FT127: return ntt_rows_impl<FT127>(src, src_stride, src_valid, dst, dst_stride, roots, log_n, n_rows, stream, n_launches, scatter, copy_dst, copy_stride);
    case FT191: return ntt_rows_impl<FT191>(src, src_stride, src_valid, dst, dst_stride, roots, log_n, n_rows, stream, n_launches, scatter, copy_dst, copy_stride);
    case FT255: return ntt_rows_impl<FT255>(src, src_stride, src_valid, dst, dst_stride, roots, log_n, n_rows, stream, n_launches, scatter, copy_dst, copy_stride);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace lcpc
