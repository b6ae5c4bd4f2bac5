// lcpc_b200/csrc/kernels_expander.cu -- Brakedown row encoding as a chain of batched sparse products.
//
// Replaces SdigEncodingS::encode -> encode() (reference: lcpc-brakedown-pc/src/lib.rs:150-153,
// lcpc-brakedown-pc/src/encode.rs:36-94) with reed_solomon (:97-110) as the base case.  The flat
// codeword of one row is
//     x_0 | x_1 | ... | x_{t-1} | RS(x_t) | v_{t-1} | ... | v_0
//   x_{i+1} = P_i x_i (precodes, :46-58; x_t goes to a temporary, :61-67)
//   RS(x_t)[k] = sum_j x_t[j] (k+1)^j  (Horner, :101-109)
//   v_i = Q_i (x_{i+1} | ... | v_{i+1})  (postcodes, reading one contiguous slice, :76-90)
// `M.dot(x)` is y[i] = sum_j M[i,j] x[j] (sprs CsMat::dot); arithmetic is exact so the summation
// order is free.
//
// All rows of the commit share the matrices, so each level is one SpMM.  The reference's CSC
// (scatter) form is turned into row-compressed (gather) form once at construction.  During the
// chain the batch lives TRANSPOSED in a work buffer W[position][row]: the gather of input position j
// then moves n_rows*B contiguous bytes for the whole batch, and the (index, value) pair of a non-zero
// is one broadcast load per warp instead of one load per row.  Two tiled transposes connect W to the
// row-major coefficient / commitment matrices of the C ABI.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/lcpc_b200.h"
#include "expander_internal.h"
#include "field.cuh"
#include "kernels.h"

namespace lcpc {

void expander_retain(ExpanderCode *c) { c->refs++; }
size_t expander_n_in(const ExpanderCode *c) { return c->n_in; }
size_t expander_codeword_length(const ExpanderCode *c) { return c->n_cols; }
size_t expander_nnz(const ExpanderCode *c) { return c->nnz; }

void expander_free(ExpanderCode *c) {
  if (!c || --c->refs > 0) return;
  for (auto &m : c->mats) {
    cudaFree(m.rowptr);
    cudaFree(m.colidx);
    cudaFree(m.vals);
    cudaFree(m.seg);
    cudaFree(m.csc_idx);
    cudaFree(m.csc_data);
  }
  delete c;
}

static int upload_csr(int field, const CscView &M, cudaStream_t stream, DeviceCsr *out, std::string *err) {
  const size_t L = field_bytes(field) / 8;  // u64 limbs per element
  if (!M.ptrs || (M.n && M.ptrs[0] != 0)) {
    *err = "bad column pointers";
    return LCPC_B200_ERR_BAD_ARG;
  }
  const size_t nnz = M.n ? M.ptrs[M.n] : 0;
  if (M.m >= 0xffffffffull || M.n >= 0xffffffffull || nnz >= 0xffffffffull) {
    *err = "matrix too large for 32-bit indices";
    return LCPC_B200_ERR_TOO_BIG;
  }
  std::vector<uint32_t> rowptr(M.m + 1, 0), colidx(nnz);
  std::vector<uint64_t> vals(nnz * L);
  for (size_t j = 0; j < M.n; j++) {
    if (M.ptrs[j + 1] < M.ptrs[j] || M.ptrs[j + 1] > nnz) {
      *err = "column pointers not monotone";
      return LCPC_B200_ERR_BAD_ARG;
    }
    for (uint64_t k = M.ptrs[j]; k < M.ptrs[j + 1]; k++) {
      if (M.idxs[k] >= M.m) {
        *err = "row index out of range";
        return LCPC_B200_ERR_BAD_ARG;
      }
      rowptr[M.idxs[k] + 1]++;
    }
  }
  for (size_t i = 0; i < M.m; i++) rowptr[i + 1] += rowptr[i];
  std::vector<uint32_t> fill(rowptr.begin(), rowptr.end() - 1);
  for (size_t j = 0; j < M.n; j++)
    for (uint64_t k = M.ptrs[j]; k < M.ptrs[j + 1]; k++) {
      uint32_t pos = fill[M.idxs[k]]++;
      colidx[pos] = (uint32_t)j;
      memcpy(&vals[(size_t)pos * L], M.data + k * L, L * 8);
    }
  out->m = M.m, out->n = M.n, out->nnz = nnz;
  cudaError_t e = cudaMalloc(&out->rowptr, (M.m + 1) * 4);
  if (e == cudaSuccess) e = cudaMalloc(&out->colidx, std::max<size_t>(nnz, 1) * 4);
  if (e == cudaSuccess) e = cudaMalloc(&out->vals, std::max<size_t>(nnz, 1) * L * 8);
  if (e == cudaSuccess) e = cudaMemcpyAsync(out->rowptr, rowptr.data(), (M.m + 1) * 4, cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess && nnz) e = cudaMemcpyAsync(out->colidx, colidx.data(), nnz * 4, cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess && nnz) e = cudaMemcpyAsync(out->vals, vals.data(), nnz * L * 8, cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);  // host vectors die at return
  if (e != cudaSuccess) {
    *err = cudaGetErrorString(e);
    return e == cudaErrorMemoryAllocation ? LCPC_B200_ERR_OOM : LCPC_B200_ERR_CUDA;
  }
  return LCPC_B200_OK;
}

int expander_assemble(ExpanderCode *c, size_t t, std::string *err) {
  // offsets exactly as encode.rs:36-94 walks them
  auto pre = [&](size_t i) -> const DeviceCsr & { return c->mats[i]; };
  auto post = [&](size_t i) -> const DeviceCsr & { return c->mats[t + i]; };
  c->n_levels = t;
  c->n_in = pre(0).n;
  size_t len = pre(0).n + post(t - 1).n;  // codeword_length, encode.rs:18-33
  for (size_t i = 0; i + 1 < t; i++) len += pre(i).m;
  for (size_t i = 0; i < t; i++) len += post(i).m;
  c->n_cols = len;
  c->ops.clear();
  size_t in_start = 0;
  for (size_t i = 0; i < t; i++) {
    if (i + 1 < t && pre(i + 1).n != pre(i).m) {
      *err = "precode dimensions do not chain";
      return LCPC_B200_ERR_BAD_ARG;
    }
    size_t in_end = in_start + pre(i).n;
    ExpanderOp op{0, (int)i, in_start, pre(i).n, in_end, pre(i).m, i + 1 == t, false};
    c->ops.push_back(op);
    in_start = in_end;
  }
  // base case: in_start is now the end of x_{t-1}; RS(x_t) fills [in_start, in_start + post[t-1].n)
  c->tmp_len = pre(t - 1).m;
  ExpanderOp rs{1, -1, 0, pre(t - 1).m, in_start, post(t - 1).n, false, true};
  c->ops.push_back(rs);
  size_t out_start = in_start + post(t - 1).n;
  in_start = in_start + pre(t - 1).m;
  for (size_t i = t; i-- > 0;) {
    in_start -= pre(i).m;
    if (out_start - in_start != post(i).n) {
      *err = "postcode dimensions do not match the codeword slice";
      return LCPC_B200_ERR_BAD_ARG;
    }
    ExpanderOp op{0, (int)(t + i), in_start, post(i).n, out_start, post(i).m, false, false};
    c->ops.push_back(op);
    out_start += post(i).m;
  }
  if (in_start != pre(0).n || out_start != len) {  // asserts at encode.rs:92-93
    *err = "codeword offsets do not close";
    return LCPC_B200_ERR_BAD_ARG;
  }
  c->nnz = 0;
  for (auto &m : c->mats) c->nnz += m.nnz;
  // innermost levels: grow the window outwards from the base code while it fits the shared-memory budget
  {
    // FUSED_SMEM_KB: the shared-memory budget of the fused window (A/B knob, read when the code is assembled)
    const size_t fused_bytes = (size_t)std::min<long>(200, std::max<long>(8, tunable("FUSED_SMEM_KB", (long)(FUSED_SMEM_BYTES >> 10)))) << 10;
    const size_t cap_elems = fused_bytes / field_bytes(c->field);
    const size_t q = t;  // index of the Reed-Solomon op: ops = pre_0..pre_{t-1}, RS, post_{t-1}..post_0
    size_t lo = q, hi = q;
    size_t wlo = c->ops[q].out_off, wend = c->ops[q].out_off + c->ops[q].out_len;
    while (lo > 0 && hi + 1 < c->ops.size()) {
      const ExpanderOp &pre_op = c->ops[lo - 1], &post_op = c->ops[hi + 1];
      const size_t nlo = pre_op.in_off, nend = post_op.out_off + post_op.out_len;
      if (nend - nlo + c->tmp_len > cap_elems || (hi - lo + 1) + 2 > (size_t)MAX_FUSED_OPS) break;
      lo--, hi++, wlo = nlo, wend = nend;
    }
    c->fuse_lo = 1, c->fuse_hi = 0;
    if (hi > lo) c->fuse_lo = lo, c->fuse_hi = hi, c->win_lo = wlo, c->win_len = wend - wlo;
  }
  return LCPC_B200_OK;
}

int expander_build(int field, size_t t, const CscView *pre, const CscView *post, cudaStream_t stream,
                   ExpanderCode **out, std::string *err) {
  ExpanderCode *c = new ExpanderCode;
  c->field = field;
  c->mats.resize(2 * t);
  int rc = LCPC_B200_OK;
  for (size_t i = 0; i < t && rc == LCPC_B200_OK; i++) rc = upload_csr(field, pre[i], stream, &c->mats[i], err);
  for (size_t i = 0; i < t && rc == LCPC_B200_OK; i++) rc = upload_csr(field, post[i], stream, &c->mats[t + i], err);
  if (rc == LCPC_B200_OK) rc = expander_assemble(c, t, err);
  if (rc != LCPC_B200_OK) {
    expander_free(c);
    return rc;
  }
  *out = c;
  return LCPC_B200_OK;
}

size_t expander_scratch_bytes(const ExpanderCode *c, size_t n_rows) {
  // W[n_cols][n_rows] | T[tmp_len][n_rows] | one
  return ((c->n_cols + c->tmp_len) * n_rows + 1) * field_bytes(c->field);
}

// ---- kernels ------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void ldv(uint32_t (&v)[N], const uint32_t *p) {
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
// the same load as a volatile asm statement: it keeps its place in front of the (volatile asm) multiply chains,
// so a batch of gathers is really in flight before the first product instead of being sunk next to its use
template <int N>
__device__ __forceinline__ void ldv_early(uint32_t (&v)[N], const uint32_t *p) {
  if constexpr (N % 4 == 0) {
#pragma unroll
    for (int i = 0; i < N / 4; i++)
      asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(v[4 * i]), "=r"(v[4 * i + 1]), "=r"(v[4 * i + 2]), "=r"(v[4 * i + 3])
                   : "l"(p + 4 * i));
  } else {
#pragma unroll
    for (int i = 0; i < N / 2; i++)
      asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(v[2 * i]), "=r"(v[2 * i + 1]) : "l"(p + 2 * i));
  }
}
// the same with an L2 eviction policy (createpolicy): the gathered work-buffer window is asked to stay (evict_last),
// the matrix stream (values, column indices) to leave first
template <int N>
__device__ __forceinline__ void ldv_early_pol(uint32_t (&v)[N], const uint32_t *p, uint64_t pol) {
  if constexpr (N % 4 == 0) {
#pragma unroll
    for (int i = 0; i < N / 4; i++)
      asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
                   : "=r"(v[4 * i]), "=r"(v[4 * i + 1]), "=r"(v[4 * i + 2]), "=r"(v[4 * i + 3])
                   : "l"(p + 4 * i), "l"(pol));
  } else {
#pragma unroll
    for (int i = 0; i < N / 2; i++)
      asm volatile("ld.global.nc.L2::cache_hint.v2.u32 {%0, %1}, [%2], %3;"
                   : "=r"(v[2 * i]), "=r"(v[2 * i + 1])
                   : "l"(p + 2 * i), "l"(pol));
  }
}
__device__ __forceinline__ uint32_t ld_u32_pol(const uint32_t *p, uint64_t pol) {
  uint32_t v;
  asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}
template <int N>
__device__ __forceinline__ void stv(uint32_t *p, const uint32_t (&v)[N]) {
  if constexpr (N % 4 == 0) {
#pragma unroll
    for (int i = 0; i < N / 4; i++)
      reinterpret_cast<uint4 *>(p)[i] = make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < N / 2; i++) reinterpret_cast<uint2 *>(p)[i] = make_uint2(v[2 * i], v[2 * i + 1]);
  }
}

// tiled transpose of elements: dst[c * dst_ld + r] = src[r * src_ld + c], r < n_r, c < n_c.  With a scatter
// descriptor (the final transpose of a multi-GPU row block) destination row c, position r goes to column
// block h(r)'s matrix instead: sc.dst[h][(row0 + c) * width_h + (r - start_h)].
constexpr int TT = 32;
template <int N>
__global__ void __launch_bounds__(256)
transpose_kernel(const uint32_t *__restrict__ src, size_t src_ld, uint32_t *__restrict__ dst, size_t dst_ld,
                 size_t n_r, size_t n_c, Scatter sc, uint32_t *__restrict__ copy_dst, size_t copy_ld, size_t src_total,
                 size_t pos0) {
  __shared__ uint32_t tile[N][TT][TT + 1];
  const size_t c0 = (size_t)blockIdx.x * TT, r0 = (size_t)blockIdx.y * TT;
  const unsigned tx = threadIdx.x % TT, ty = threadIdx.x / TT;  // 32 x 8
  for (unsigned rr = ty; rr < TT; rr += 8) {
    size_t r = r0 + rr, cidx = c0 + tx;
    if (r < n_r && cidx < n_c) {
      uint32_t v[N];
      if (r * src_ld + cidx < src_total) {
        ldv<N>(v, src + (r * src_ld + cidx) * N);
      } else {  // beyond the caller's coefficients: the zero padding of the last row (lcpc-2d/src/lib.rs:636-645)
#pragma unroll
        for (int l = 0; l < N; l++) v[l] = 0;
      }
      if (copy_dst) stv<N>(copy_dst + (r * copy_ld + cidx) * N, v);  // commit()'s own copy of the coefficients
#pragma unroll
      for (int l = 0; l < N; l++) tile[l][rr][tx] = v[l];
    }
  }
  __syncthreads();
  for (unsigned cc = ty; cc < TT; cc += 8) {
    size_t cidx = c0 + cc, r = r0 + tx;
    if (r < n_r && cidx < n_c) {
      uint32_t v[N];
#pragma unroll
      for (int l = 0; l < N; l++) v[l] = tile[l][tx][cc];
      const size_t pos = pos0 + r;  // destination position (pos0 > 0: the source starts at work-buffer position pos0)
      uint32_t *out = dst + (cidx * dst_ld + pos) * N;
      if (sc.n_blocks) {
        unsigned h = 0;
        while (h + 1 < sc.n_blocks && pos >= sc.starts[h + 1]) h++;
        const size_t start = sc.starts[h], width = sc.starts[h + 1] - start;
        out = sc.dst[h] + ((sc.row0 + cidx) * width + (pos - start)) * N;
      }
      stv<N>(out, v);
    }
  }
}

// The systematic part of the codeword (positions [0, n_in) = the coefficients themselves, two thirds of it) is
// final before the chain starts: in the multi-GPU commit it leaves for its column owners on a side stream while
// the sparse products run, straight from the row-major source rows (no transpose involved).
template <int N>
__global__ void __launch_bounds__(256)
scatter_rows_kernel(const uint32_t *__restrict__ src, size_t src_ld, size_t src_total, size_t n_rows, size_t n_in,
                    Scatter sc) {
  const size_t total = n_rows * n_in;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t r = idx / n_in, pos = idx % n_in;
    uint32_t v[N];
    if (r * src_ld + pos < src_total) {
      ldv<N>(v, src + (r * src_ld + pos) * N);
    } else {
#pragma unroll
      for (int l = 0; l < N; l++) v[l] = 0;
    }
    unsigned h = 0;
    while (h + 1 < sc.n_blocks && pos >= sc.starts[h + 1]) h++;
    const size_t start = sc.starts[h], width = sc.starts[h + 1] - start;
    stv<N>(sc.dst[h] + ((sc.row0 + r) * width + (pos - start)) * N, v);
  }
}

// y[i][r] = sum_k vals[k] * x[colidx[k]][r]; one thread per (output i, batch row r), r fastest.
// The sum is kept double-width and Montgomery-reduced once per output (field.cuh, mac_wide / redc):
// a sparse row has 8..45 terms, so this halves the multiplier work against reduce-every-product.
// The gathers are what matters: every input position is read once per non-zero of its column, at random, so the
// chain pulls nnz * n_rows * B bytes (6x the algorithmic bytes of the encode) through L2 -- and from DRAM as well
// unless the gathered window of the work buffer stays L2-resident (round 1: 1.62 GB of DRAM reads per big level,
// 19 % L2 hits).  Two measures keep it resident:
//  * HINT: gathers carry an evict_last L2 policy, the matrix stream (values, indices) evict_first, so that the
//    window is not pushed out by data that is used once;
//  * column chunks: a level whose window n * n_rows * B exceeds the L2 budget runs as Q launches, launch q
//    covering the non-zeros with columns in [q n/Q, (q+1) n/Q) (the rows are column-sorted, `seg` holds the
//    cut points) and ACCUMulating onto y: y = y + REDC(partial sum), exact mod p, so the result is the same
//    canonical element.  Each window chunk then comes from DRAM once.
// The loop issues SPMM_UNROLL index loads, then that many gathers, before the first multiply.
// __launch_bounds__(256, 4) is what makes ptxas keep all eight 16-byte loads of a batch in front of the first
// product (64 registers); left to itself it settles on 48 registers and sinks each load next to its use, which
// measured 6 % slower on the 2^24 chain.  Wider elements (Ft191, Ft255) batch two deep under a 128-register budget.
template <int FID, bool HINT, bool ACCUM>
__global__ void __launch_bounds__(256, (Field<FID>::N <= 4 ? 4 : 2))
spmm_kernel(const uint32_t *__restrict__ seg_lo, const uint32_t *__restrict__ seg_hi, const uint32_t *__restrict__ colidx,
            const uint32_t *__restrict__ vals, const uint32_t *__restrict__ x, uint32_t *__restrict__ y, size_t m,
            size_t n_rows, unsigned r0, unsigned rg, float keep_frac) {
  using F = Field<FID>;
  constexpr int N = F::N;
  constexpr int SPMM_UNROLL = N <= 4 ? 4 : 2;
  const size_t item = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= m * rg) return;
  const size_t i = item / rg, r = r0 + item % rg;
  const uint32_t k0 = __ldg(seg_lo + i), k1 = __ldg(seg_hi + i);
  uint64_t pol_keep = 0, pol_stream = 0;
  if (HINT) {
    asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, %1;" : "=l"(pol_keep) : "f"(keep_frac));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
  }
  typename F::Wide acc = F::wide_zero();
  // address arithmetic kept off the multiplier pipe: one 32x32->64 product per gather, pointer bumps elsewhere
  const uint32_t *xr = x + r * N;
  const uint32_t pos_stride = (uint32_t)(n_rows * N);  // limbs between consecutive positions (host checks < 2^32)
  const uint32_t *vp = vals + (size_t)k0 * N;
  const uint32_t *cp = colidx + k0;
  uint32_t left = k1 - k0;
  for (; left >= SPMM_UNROLL; left -= SPMM_UNROLL, vp += SPMM_UNROLL * N, cp += SPMM_UNROLL) {
    uint32_t j[SPMM_UNROLL];
    typename F::Elem a[SPMM_UNROLL], xv[SPMM_UNROLL];
#pragma unroll
    for (int u = 0; u < SPMM_UNROLL; u++) j[u] = HINT ? ld_u32_pol(cp + u, pol_stream) : __ldg(cp + u);
#pragma unroll
    for (int u = 0; u < SPMM_UNROLL; u++) {
      if (HINT) {
        ldv_early_pol<N>(xv[u].v, xr + (size_t)j[u] * pos_stride, pol_keep);
        ldv_early_pol<N>(a[u].v, vp + u * N, pol_stream);
      } else {
        ldv_early<N>(xv[u].v, xr + (size_t)j[u] * pos_stride);
        ldv_early<N>(a[u].v, vp + u * N);
      }
    }
#pragma unroll
    for (int u = 0; u < SPMM_UNROLL; u++) F::mac_wide(acc, a[u], xv[u]);
  }
  for (; left; left--, vp += N, cp++) {
    const uint32_t j = __ldg(cp);
    typename F::Elem a, xv;
    ldv<N>(a.v, vp);
    if (HINT) ldv_early_pol<N>(xv.v, xr + (size_t)j * pos_stride, pol_keep);
    else ldv<N>(xv.v, xr + (size_t)j * pos_stride);
    F::mac_wide(acc, a, xv);
  }
  typename F::Elem out = F::template redc<2>(acc);
  uint32_t *yp = y + (i * n_rows + r) * N;
  if (ACCUM) {
    typename F::Elem prev;
    ldv<N>(prev.v, yp);
    out = F::add(out, prev);
  }
  stv<N>(yp, out.v);
}

// The same product with the carry-counting accumulator of field.cuh (Field::Sum): the rows of every term
// multiply-add straight into two persistent accumulator arrays and only the 2N chain carries per term are counted,
// instead of forming each product, merging its two arrays, adding it to the running sum and folding
// (mac_wide: ~6N ALU operations per term, and ptxas turns a quarter of them into IMAD.X / IMAD.MOV on the multiplier
// pipe -- the pipe that bounds this kernel: ncu sm__throughput 75-77 % with issue 52 %, ALU 50 %).
// Same canonical result (sum_reduce brings the exact sum to [0, p)).  SPMM_MAC=0 selects spmm_kernel.
// DEEP: twice as many gathers in flight per thread (8 for the 2- and 4-limb fields), the matrix values loaded just in
// time (they are shared by all batch rows of an output: L1 hits), three CTAs per SM instead of four -- an A/B knob
// (SPMM_SUM_DEEP) for the levels whose gathers miss in L2.
template <int FID, bool DEEP>
__global__ void __launch_bounds__(256, (Field<FID>::N <= 4 ? (DEEP ? 3 : 4) : 2))
spmm_sum_kernel(const uint32_t *__restrict__ rowptr, const uint32_t *__restrict__ colidx, const uint32_t *__restrict__ vals,
                const uint32_t *__restrict__ x, uint32_t *__restrict__ y, size_t m, size_t n_rows) {
  using F = Field<FID>;
  constexpr int N = F::N;
  constexpr int U = (N <= 4 ? 4 : 2) * (DEEP ? 2 : 1);
  const size_t item = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= m * n_rows) return;
  const size_t i = item / n_rows, r = item % n_rows;
  const uint32_t k0 = __ldg(rowptr + i), k1 = __ldg(rowptr + i + 1);
  typename F::Sum acc = F::sum_zero();
  const uint32_t *xr = x + r * N;
  const uint32_t pos_stride = (uint32_t)(n_rows * N);
  const uint32_t *vp = vals + (size_t)k0 * N;
  const uint32_t *cp = colidx + k0;
  uint32_t left = k1 - k0;
  for (; left >= U; left -= U, vp += U * N, cp += U) {
    uint32_t j[U];
    typename F::Elem xv[U];
#pragma unroll
    for (int u = 0; u < U; u++) j[u] = __ldg(cp + u);
    if constexpr (DEEP) {
#pragma unroll
      for (int u = 0; u < U; u++) ldv_early<N>(xv[u].v, xr + (size_t)j[u] * pos_stride);
#pragma unroll
      for (int u = 0; u < U; u++) {
        typename F::Elem a;
        ldv<N>(a.v, vp + u * N);
        F::sum_mac(acc, xv[u], a);
      }
    } else {
      typename F::Elem a[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        ldv_early<N>(xv[u].v, xr + (size_t)j[u] * pos_stride);
        ldv_early<N>(a[u].v, vp + u * N);
      }
#pragma unroll
      for (int u = 0; u < U; u++) F::sum_mac(acc, xv[u], a[u]);
    }
  }
  for (; left; left--, vp += N, cp++) {
    const uint32_t j = __ldg(cp);
    typename F::Elem a, xv;
    ldv<N>(a.v, vp);
    ldv<N>(xv.v, xr + (size_t)j * pos_stride);
    F::sum_mac(acc, xv, a);
  }
  typename F::Elem out = F::sum_reduce(acc);
  stv<N>(y + (i * n_rows + r) * N, out.v);
}

// Small levels (a few thousand outputs x the batch rows: fewer threads than the GPU has slots) are pure latency in
// spmm_sum_kernel -- one thread walks its output's ~45 non-zeros in rounds of four dependent gathers.  Here G lanes
// share an output: lane g takes non-zeros g, g + G, ..., the partial sums are collected (Field::Collected, one
// 2N+1-limb integer each) and added across the G lanes with shuffles, lane 0 reduces and stores.  Lanes with the same
// g hold consecutive batch rows, so a gather is still a run of 32 / G elements.  Same exact sum, same result.
template <int FID, int G>
__global__ void __launch_bounds__(256)
spmm_sum_split_kernel(const uint32_t *__restrict__ rowptr, const uint32_t *__restrict__ colidx, const uint32_t *__restrict__ vals,
                      const uint32_t *__restrict__ x, uint32_t *__restrict__ y, size_t m, size_t n_rows) {
  using F = Field<FID>;
  constexpr int N = F::N;
  constexpr int U = N <= 4 ? 4 : 2;
  constexpr unsigned PER_WARP = 32 / G;  // (output, batch row) items per warp
  const unsigned lane = threadIdx.x & 31u, g = lane / PER_WARP;
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t item = warp * PER_WARP + lane % PER_WARP, total = m * n_rows;
  const bool live = item < total;
  const size_t i = live ? item / n_rows : 0, r = live ? item % n_rows : 0;
  typename F::Sum acc = F::sum_zero();
  if (live) {
    const uint32_t k0 = __ldg(rowptr + i), k1 = __ldg(rowptr + i + 1);
    const uint32_t *xr = x + r * N;
    const uint32_t pos_stride = (uint32_t)(n_rows * N);
    for (uint32_t kb = k0 + g; kb < k1; kb += U * G) {
      uint32_t j[U];
      typename F::Elem a[U], xv[U];
#pragma unroll
      for (int u = 0; u < U; u++) j[u] = __ldg(colidx + min(kb + u * G, k1 - 1));
#pragma unroll
      for (int u = 0; u < U; u++) {
        ldv_early<N>(xv[u].v, xr + (size_t)j[u] * pos_stride);
        ldv_early<N>(a[u].v, vals + (size_t)min(kb + u * G, k1 - 1) * N);
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        if (kb + u * G >= k1) {  // past the row's end: a zero factor adds nothing
#pragma unroll
          for (int l = 0; l < N; l++) a[u].v[l] = 0;
        }
        F::sum_mac(acc, xv[u], a[u]);
      }
    }
  }
  typename F::Collected t = F::sum_collect(acc);
  __syncwarp();
#pragma unroll
  for (unsigned off = PER_WARP; off < 32; off <<= 1) {
    typename F::Collected o;
#pragma unroll
    for (int l = 0; l < 2 * N + 1; l++) o.v[l] = __shfl_xor_sync(0xffffffffu, t.v[l], off);
    F::collected_add(t, o);
  }
  if (live && g == 0) {
    typename F::Elem out = F::collected_reduce(t);
    stv<N>(y + (i * n_rows + r) * N, out.v);
  }
}

// The same product with the gathers SOFTWARE-PIPELINED one batch ahead: while a thread multiplies batch b, the
// gathers of batch b+1 are already in flight and the column indices of batch b+2 are on their way, so a warp never
// sits in front of an empty load queue between its multiply phases (spmm_kernel alternates "wait for 4 gathers" and
// "4 products", and ran at 3.1 KB/clk of L2 traffic, half of what the L2 delivers).  Costs registers for a second
// batch of gathered elements (launch bounds allow 3 CTAs per SM instead of 4 for 2- and 4-limb fields); the matrix
// values are loaded just in time (they are shared by all batch rows of an output: L1 hits).  A ragged tail is one
// more full batch with clamped indices and masked products, not a serial remainder loop -- which is what makes short
// rows (column chunks) affordable.
template <int FID, bool ACCUM>
__global__ void __launch_bounds__(256, (Field<FID>::N <= 4 ? 3 : 2))
spmm_pipe_kernel(const uint32_t *__restrict__ seg_lo, const uint32_t *__restrict__ seg_hi, const uint32_t *__restrict__ colidx,
                 const uint32_t *__restrict__ vals, const uint32_t *__restrict__ x, uint32_t *__restrict__ y, size_t m,
                 size_t n_rows, unsigned r0, unsigned rg) {
  using F = Field<FID>;
  constexpr int N = F::N;
  constexpr int U = N <= 4 ? 4 : 2;
  const size_t item = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= m * rg) return;
  const size_t i = item / rg, r = r0 + item % rg;
  const uint32_t k0 = __ldg(seg_lo + i), k1 = __ldg(seg_hi + i);
  typename F::Wide acc = F::wide_zero();
  const uint32_t *xr = x + r * N;
  const uint32_t pos_stride = (uint32_t)(n_rows * N);
  uint32_t *yp = y + (i * n_rows + r) * N;
  if (k1 > k0) {
    const uint32_t last = k1 - 1;
    const uint32_t nb = (k1 - k0 + U - 1) / U;
    auto load_idx = [&](uint32_t (&j)[U], uint32_t b) {
#pragma unroll
      for (int u = 0; u < U; u++) {
        const uint32_t k = min(k0 + b * U + u, last);
        asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(j[u]) : "l"(colidx + k));
      }
    };
    auto gather = [&](typename F::Elem (&xv)[U], const uint32_t (&j)[U]) {
#pragma unroll
      for (int u = 0; u < U; u++) ldv_early<N>(xv[u].v, xr + (size_t)j[u] * pos_stride);
    };
    auto compute = [&](const typename F::Elem (&xv)[U], uint32_t b) {
#pragma unroll
      for (int u = 0; u < U; u++) {
        const uint32_t k = k0 + b * U + u;
        if (k <= last) {
          typename F::Elem a;
          ldv_early<N>(a.v, vals + (size_t)k * N);
          F::mac_wide(acc, a, xv[u]);
        }
      }
    };
    uint32_t jA[U], jB[U];
    typename F::Elem xA[U], xB[U];
    load_idx(jA, 0);
    if (nb > 1) load_idx(jB, 1);
    gather(xA, jA);
    uint32_t b = 0;
    for (;;) {
      if (b + 1 < nb) gather(xB, jB);
      if (b + 2 < nb) load_idx(jA, b + 2);
      compute(xA, b);
      if (++b >= nb) break;
      if (b + 1 < nb) gather(xA, jA);
      if (b + 2 < nb) load_idx(jB, b + 2);
      compute(xB, b);
      if (++b >= nb) break;
    }
  }
  typename F::Elem out = F::template redc<2>(acc);
  if (ACCUM) {
    typename F::Elem prev;
    ldv<N>(prev.v, yp);
    out = F::add(out, prev);
  }
  stv<N>(yp, out.v);
}

// ---- the same product with the gathers done by the SM's bulk-copy engine (TMA, cp.async.bulk) -----------------------
// A gather of input position j for the whole batch is ONE contiguous run of n_rows * B bytes of the work buffer
// (1152 B at 2^24): exactly what a 1-D bulk copy moves.  A producer warp walks the non-zero stream of each group of
// threads (the outputs [i0, i1) of a group are consecutive, so their (column, value) pairs are consecutive in the
// row-compressed arrays) and keeps a ring of STAGES x U gathers per group in flight -- global -> shared memory,
// completion counted in bytes on an mbarrier -- while the consumer threads (one per output-in-progress and batch
// row) multiply out of shared memory.  No registers are held by loads in flight, the depth does not depend on
// occupancy, and a thread never waits for its own gather.  Fields whose element is a multiple of 16 bytes only
// (Ft127, Ft255: what the reference benchmarks); the others keep spmm_kernel.
namespace bulk {
constexpr int U = 4;        // non-zeros per stage
constexpr int STAGES = 4;   // ring depth per group
constexpr int MAX_GROUPS = 32;
constexpr int CONSUMERS = 256;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
}  // namespace bulk

// grid.x: CTAs over groups of outputs; grid.y: batch-row tiles of RT rows.  Block = 32 producer lanes + CONSUMERS.
// Dynamic shared memory: [full barriers G*STAGES | empty barriers G*STAGES | per group, per stage: U*RT elements, U values]
template <int FID>
__global__ void __launch_bounds__(32 + bulk::CONSUMERS)
spmm_bulk_kernel(const uint32_t *__restrict__ rowptr, const uint32_t *__restrict__ colidx, const uint32_t *__restrict__ vals,
                 const uint32_t *__restrict__ x, uint32_t *__restrict__ y, uint32_t m, uint32_t n_rows, uint32_t RT, uint32_t G,
                 uint32_t outs_per_group) {
  using F = Field<FID>;
  constexpr int N = F::N;
  constexpr int B = F::BYTES;
  using namespace bulk;
  extern __shared__ __align__(128) uint8_t bsm[];
  uint64_t *full = reinterpret_cast<uint64_t *>(bsm);
  uint64_t *empty = full + G * STAGES;
  const uint32_t stage_bytes = (U * RT * B + U * B + 127) & ~127u;  // x[U][RT] then vals[U]
  uint8_t *ring = bsm + ((2 * G * STAGES * 8 + 127) & ~127u);
  const uint32_t r0 = blockIdx.y * RT;
  const uint32_t rt = min(RT, n_rows - r0);  // rows of this tile
  const uint32_t tid = threadIdx.x;
  if (tid == 0) {
    for (uint32_t q = 0; q < G * STAGES; q++) {
      mbar_init(full + q, 1);
      mbar_init(empty + q, rt);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  // group g of this CTA works on outputs [i0, i1) = non-zeros [K0, K1)
  auto group_range = [&](uint32_t g, uint32_t &i0, uint32_t &i1) {
    const uint64_t first = ((uint64_t)blockIdx.x * G + g) * outs_per_group;
    i0 = (uint32_t)(first < m ? first : m);
    i1 = (uint32_t)(first + outs_per_group < m ? first + outs_per_group : m);
  };
  if (tid < 32) {
    // ---- producer: lane = (group within a batch of 8, non-zero within the stage) ----
    const uint32_t lane = tid, u = lane & 3;
    const uint32_t n_batches = (G + 7) / 8;
    uint32_t K0[MAX_GROUPS / 8], K1[MAX_GROUPS / 8], s_idx[MAX_GROUPS / 8], jn[MAX_GROUPS / 8];
#pragma unroll
    for (uint32_t b = 0; b < MAX_GROUPS / 8; b++) {
      K0[b] = K1[b] = 0, s_idx[b] = 0, jn[b] = 0;
      const uint32_t g = b * 8 + (lane >> 2);
      if (b < n_batches && g < G) {
        uint32_t i0, i1;
        group_range(g, i0, i1);
        K0[b] = __ldg(rowptr + i0), K1[b] = __ldg(rowptr + i1);
        const uint32_t k = K0[b] + u;
        if (k < K1[b]) jn[b] = __ldg(colidx + k);
      }
    }
    bool any = true;
    while (any) {
      any = false;
#pragma unroll
      for (uint32_t b = 0; b < MAX_GROUPS / 8; b++) {
        if (b >= n_batches) break;
        const uint32_t g = b * 8 + (lane >> 2);
        const uint32_t kbase = K0[b] + s_idx[b] * U;
        const bool live = g < G && kbase < K1[b];
        if (live) {
          const uint32_t slot = s_idx[b] % STAGES, round = s_idx[b] / STAGES;
          uint64_t *fb = full + g * STAGES + slot, *eb = empty + g * STAGES + slot;
          if (round > 0) mbar_wait(eb, (round - 1) & 1);  // the consumers are done with what this slot held
          const uint32_t n_valid = min((uint32_t)U, K1[b] - kbase);
          uint8_t *st = ring + (size_t)(g * STAGES + slot) * stage_bytes;
          if (u == 0) mbar_expect_tx(fb, n_valid * rt * B + n_valid * B);
          __syncwarp(__activemask());
          if (u < n_valid) bulk_g2s(st + (size_t)u * RT * B, x + ((size_t)jn[b] * n_rows + r0) * N, rt * B, fb);
          if (u == 0) bulk_g2s(st + (size_t)U * RT * B, vals + (size_t)kbase * N, n_valid * B, fb);
          s_idx[b]++;
          const uint32_t kn = kbase + U + u;
          if (kn < K1[b]) jn[b] = __ldg(colidx + kn);  // next stage's column, in flight until the next visit
          if (kbase + U < K1[b]) any = true;
        }
      }
      any = __any_sync(0xffffffffu, any);
    }
  } else {
    // ---- consumers: thread = (group, batch row) ----
    const uint32_t ct = tid - 32;
    const uint32_t g = ct / rt, r = ct % rt;
    if (g < G) {
      uint32_t i0, i1;
      group_range(g, i0, i1);
      if (i0 < i1) {
        const uint32_t K0 = __ldg(rowptr + i0), K1 = __ldg(rowptr + i1);
        uint32_t i = i0, row_end = __ldg(rowptr + i0 + 1);
        typename F::Wide acc = F::wide_zero();
        auto finish_row = [&]() {
          typename F::Elem out = F::template redc<2>(acc);
          stv<N>(y + ((size_t)i * n_rows + r0 + r) * N, out.v);
          acc = F::wide_zero();
          i++;
          if (i < i1) row_end = __ldg(rowptr + i + 1);
        };
        uint32_t s_idx = 0;
        for (uint32_t kbase = K0; kbase < K1; kbase += U, s_idx++) {
          const uint32_t slot = s_idx % STAGES, round = s_idx / STAGES;
          uint64_t *fb = full + g * STAGES + slot, *eb = empty + g * STAGES + slot;
          mbar_wait(fb, round & 1);
          const uint8_t *st = ring + (size_t)(g * STAGES + slot) * stage_bytes;
          const uint32_t n_valid = min((uint32_t)U, K1 - kbase);
#pragma unroll
          for (int uu = 0; uu < U; uu++) {
            if ((uint32_t)uu < n_valid) {
              while (kbase + uu == row_end && i < i1) finish_row();
              typename F::Elem a, xv;
              ldv<N>(a.v, reinterpret_cast<const uint32_t *>(st + (size_t)U * RT * B) + uu * N);
              ldv<N>(xv.v, reinterpret_cast<const uint32_t *>(st) + ((size_t)uu * RT + r) * N);
              F::mac_wide(acc, a, xv);
            }
          }
          mbar_arrive(eb);
        }
        while (i < i1) finish_row();  // the last output, and any empty rows behind it
      }
    }
  }
}

// seg[q * m + i] = first k in [rowptr[i], rowptr[i+1]) with colidx[k] >= q * n / Q  (q = 0 .. Q)
__global__ void __launch_bounds__(256)
spmm_segments_kernel(const uint32_t *__restrict__ rowptr, const uint32_t *__restrict__ colidx, size_t m, size_t n,
                     unsigned Q, uint32_t *__restrict__ seg) {
  const size_t item = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= m * (Q + 1)) return;
  const size_t q = item / m, i = item % m;
  uint32_t lo = rowptr[i], hi = rowptr[i + 1];
  const uint32_t bound = (uint32_t)(q * n / Q);
  if (q == Q) lo = hi;
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (colidx[mid] < bound) lo = mid + 1;
    else hi = mid;
  }
  seg[q * m + i] = lo;
}

// reed_solomon (encode.rs:97-110): out[k][r] = Horner of in[.][r] at the point k+1
template <int FID>
__global__ void __launch_bounds__(128)
reed_solomon_kernel(const uint32_t *__restrict__ xin, size_t n_in, uint32_t *__restrict__ out, size_t n_out,
                    size_t n_rows, const uint32_t *__restrict__ one_mont) {
  using F = Field<FID>;
  constexpr int N = F::N;
  const size_t item = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= n_out * n_rows) return;
  const size_t k = item / n_rows, r = item % n_rows;
  typename F::Elem one, pt;
  ldv<N>(one.v, one_mont);
  pt = one;
  for (size_t q = 0; q < k; q++) pt = F::add(pt, one);  // the field element k+1
  typename F::Elem acc = F::zero();
  for (size_t j = n_in; j-- > 0;) {
    typename F::Elem c;
    ldv<N>(c.v, xin + (j * n_rows + r) * N);
    acc = F::add(F::mul(acc, pt), c);
  }
  stv<N>(out + (k * n_rows + r) * N, acc.v);
}

// The innermost levels of the chain for one batch row per CTA, the touched codeword window staged in shared
// memory: sizes there are a few hundred outputs per level, so separate launches are pure latency.
struct FusedOp {
  int kind;  // 0 sparse product, 1 reed-solomon
  const uint32_t *rowptr, *colidx, *vals;
  uint32_t in_off, in_len, out_off, out_len;  // offsets relative to the window start
  int in_tmp, out_tmp;
  uint32_t group;  // sparse product: lanes sharing one output (power of two <= 32), so that short levels use the CTA
};
struct FusedOps {
  int n;
  FusedOp op[MAX_FUSED_OPS];
};

constexpr int FUSED_THREADS = 1024;
template <int FID>
__global__ void __launch_bounds__(FUSED_THREADS)
fused_levels_kernel(FusedOps ops, uint32_t *__restrict__ W, size_t n_rows, uint32_t win_lo, uint32_t win_len,
                    uint32_t first_in_len, uint32_t tmp_len) {
  using F = Field<FID>;
  constexpr int N = F::N;
  extern __shared__ __align__(16) uint32_t fsm[];
  uint32_t *win = fsm, *tmp = fsm + (size_t)win_len * N;
  const size_t r = blockIdx.x;
  // the first op's input (x_i) is the only part of the window computed before this kernel
  for (uint32_t e = threadIdx.x; e < first_in_len; e += blockDim.x) {
    uint32_t v[N];
    ldv<N>(v, W + ((size_t)(win_lo + e) * n_rows + r) * N);
    stv<N>(win + (size_t)e * N, v);
  }
  __syncthreads();
  const typename F::Elem one = F::one();  // R mod p
  for (int o = 0; o < ops.n; o++) {
    const FusedOp &op = ops.op[o];
    const uint32_t *in = (op.in_tmp ? tmp : win + (size_t)op.in_off * N);
    uint32_t *out = (op.out_tmp ? tmp : win + (size_t)op.out_off * N);
    if (op.kind == 0) {
      // `group` lanes split the non-zeros of one output and merge their double-width partial sums with shuffles:
      // these levels have tens to hundreds of outputs, far fewer than the CTA has threads
      const unsigned G = op.group, gl = threadIdx.x & (G - 1), per_pass = blockDim.x / G;
      for (uint32_t base = 0; base < op.out_len; base += per_pass) {  // same trip count for every thread
        const uint32_t i = base + threadIdx.x / G;
        const bool act = i < op.out_len;
        typename F::Wide acc = F::wide_zero();
        if (act) {
          const uint32_t k0 = __ldg(op.rowptr + i), k1 = __ldg(op.rowptr + i + 1);
          // column indices and values of up to FB non-zeros are requested before the first product: the matrices of
          // these levels sit in L2, and a lane walks ~3..15 non-zeros, so one L2 round trip per non-zero was most
          // of this kernel's time (it runs on n_rows CTAs only: pure latency)
          constexpr int FB = N <= 4 ? 4 : 2;
          for (uint32_t kb = k0 + gl; kb < k1; kb += FB * G) {
            uint32_t j[FB];
            typename F::Elem a[FB];
#pragma unroll
            for (int u = 0; u < FB; u++) {
              const uint32_t k = min(kb + u * G, k1 - 1);
              asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(j[u]) : "l"(op.colidx + k));
              ldv_early<N>(a[u].v, op.vals + (size_t)k * N);
            }
#pragma unroll
            for (int u = 0; u < FB; u++) {
              if (kb + u * G < k1) {
                typename F::Elem xv;
                ldv<N>(xv.v, in + (size_t)j[u] * N);
                F::mac_wide(acc, a[u], xv);
              }
            }
          }
        }
        for (unsigned d = G >> 1; d >= 1; d >>= 1) {
          typename F::Wide other;
#pragma unroll
          for (int l = 0; l < 2 * N; l++) other.v[l] = __shfl_down_sync(0xffffffffu, acc.v[l], d, G);
          F::wide_merge(acc, other);
        }
        if (act && gl == 0) {
          typename F::Elem res = F::template redc<2>(acc);
          stv<N>(out + (size_t)i * N, res.v);
        }
      }
    } else {  // reed_solomon (encode.rs:97-110): Horner at the point i + 1
      for (uint32_t i = threadIdx.x; i < op.out_len; i += blockDim.x) {
        typename F::Elem pt = one;
        for (uint32_t q = 0; q < i; q++) pt = F::add(pt, one);
        typename F::Elem res = F::zero();
        for (uint32_t j = op.in_len; j-- > 0;) {
          typename F::Elem cf;
          ldv<N>(cf.v, in + (size_t)j * N);
          res = F::add(F::mul(res, pt), cf);
        }
        stv<N>(out + (size_t)i * N, res.v);
      }
    }
    __syncthreads();
  }
  for (uint32_t e = first_in_len + threadIdx.x; e < win_len; e += blockDim.x) {
    uint32_t v[N];
    ldv<N>(v, win + (size_t)e * N);
    stv<N>(W + ((size_t)(win_lo + e) * n_rows + r) * N, v);
  }
}

template <int FID> __global__ void one_mont_kernel(uint32_t *out) {
  using F = Field<FID>;
  const typename F::Elem one = F::one();  // 2^(32N) mod p
  for (int l = 0; l < F::N; l++) out[l] = one.v[l];
}

template <int FID>
static cudaError_t encode_impl(const ExpanderCode *c, const uint32_t *src, size_t src_stride, size_t valid,
                               uint32_t *dst, size_t dst_stride, size_t n_rows, void *scratch, cudaStream_t st,
                               int *n_launches, const Scatter *scatter, uint32_t *copy_dst, size_t copy_stride,
                               size_t src_total, const SideLane *side) {
  using F = Field<FID>;
  constexpr int N = F::N;
  if (n_launches) *n_launches = 0;
  if (n_rows == 0) return cudaSuccess;
  if (valid < c->n_in || !scratch) return cudaErrorInvalidValue;
  if (n_rows * N >= 0xffffffffull) return cudaErrorInvalidValue;  // spmm_kernel's 32-bit position stride
  uint32_t *W = (uint32_t *)scratch;                       // [n_cols][n_rows]
  uint32_t *T = W + c->n_cols * n_rows * N;                // [tmp_len][n_rows]
  int launches = 0;
  const bool early = scatter && scatter->n_blocks && side && side->stream;
  if (early) {  // systematic positions leave now, on the side stream, overlapped with the chain below
    cudaError_t e = cudaEventRecord(side->fork, st);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(side->stream, side->fork, 0);
    if (e != cudaSuccess) return e;
    const size_t total = n_rows * c->n_in;
    const unsigned grid = (unsigned)std::min<size_t>((total + 255) / 256, 148 * 16);
    scatter_rows_kernel<N><<<grid, 256, 0, side->stream>>>(src, src_stride, src_total, n_rows, c->n_in, *scatter);
    launches++;
    e = cudaEventRecord(side->join, side->stream);
    if (e != cudaSuccess) return e;
  }
  // coefficients -> W[0 .. n_in)
  {
    dim3 grid((unsigned)((c->n_in + TT - 1) / TT), (unsigned)((n_rows + TT - 1) / TT));
    Scatter none;
    none.n_blocks = 0;
    transpose_kernel<N><<<grid, 256, 0, st>>>(src, src_stride, W, n_rows, n_rows, c->n_in, none, copy_dst, copy_stride, src_total, 0);
    launches++;
  }
  for (size_t oi = 0; oi < c->ops.size(); oi++) {
    const ExpanderOp &op = c->ops[oi];
    if (oi == c->fuse_lo && c->fuse_hi > c->fuse_lo) {
      FusedOps fo;
      fo.n = 0;
      for (size_t q = c->fuse_lo; q <= c->fuse_hi; q++) {
        const ExpanderOp &g = c->ops[q];
        FusedOp &f = fo.op[fo.n++];
        f.kind = g.kind, f.in_tmp = g.in_tmp, f.out_tmp = g.out_tmp;
        f.in_len = (uint32_t)g.in_len, f.out_len = (uint32_t)g.out_len;
        f.in_off = g.in_tmp ? 0 : (uint32_t)(g.in_off - c->win_lo);
        f.out_off = g.out_tmp ? 0 : (uint32_t)(g.out_off - c->win_lo);
        f.group = 1;
        while (f.group < 32 && (size_t)f.out_len * (f.group * 2) <= (size_t)FUSED_THREADS) f.group *= 2;
        if (g.kind == 0) {
          const DeviceCsr &M = c->mats[g.mat];
          f.rowptr = M.rowptr, f.colidx = M.colidx, f.vals = M.vals;
        } else {
          f.rowptr = f.colidx = f.vals = nullptr;
        }
      }
      const size_t smem = (c->win_len + c->tmp_len) * F::BYTES;
      static bool attr_set_dev[64] = {};  // per device
      int dev_id = 0;
      cudaGetDevice(&dev_id);
      bool &attr_set = attr_set_dev[dev_id & 63];
      if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(fused_levels_kernel<FID>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             200 << 10);
        if (e != cudaSuccess) return e;
        attr_set = true;
      }
      fused_levels_kernel<FID><<<(unsigned)n_rows, FUSED_THREADS, smem, st>>>(fo, W, n_rows, (uint32_t)c->win_lo, (uint32_t)c->win_len,
                                                                    (uint32_t)c->ops[c->fuse_lo].in_len, (uint32_t)c->tmp_len);
      launches++;
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) return e;
      oi = c->fuse_hi;
      continue;
    }
    if (op.kind == 0) {
      const DeviceCsr &M = c->mats[op.mat];
      const uint32_t *x = W + op.in_off * n_rows * N;
      uint32_t *y = op.out_tmp ? T : W + op.out_off * n_rows * N;
      // Schedule of this level (all knobs are A/B tunables, see tunables.cpp; none changes the result):
      //   SPMM_HINTS      1: L2 policies on the gathers / the matrix stream (default), 0: plain loads
      //   SPMM_WINDOW_KB  L2 budget for the gathered window; a larger window runs as column chunks (0: never)
      //   SPMM_SLICE_KB   round-1 alternative: split the BATCH rows so that the slice fits (0: off)
      // Measured at 2^24 (profiles/r02_ab_brakedown_schedules.jsonl, r02_ncu_spmm_*.csv): column chunks cut the first
      // level's DRAM reads from 1.62 GB to 0.65 GB but the level takes 719 us instead of 444 us -- this kernel is
      // bound by dependent gather rounds per thread, not by DRAM bytes, and shorter per-launch rows mean more
      // rounds; the hints change nothing measurable.  Both therefore default to off here.
      const bool hints = tunable("SPMM_HINTS", 0) != 0;
      const bool pipe = tunable("SPMM_PIPE", 0) != 0;  // software-pipelined gathers (spmm_pipe_kernel)
      // A/B knob: unused dynamic shared memory that lowers the resident CTAs per SM (room for a co-resident hash kernel)
      const size_t spmm_pad = (size_t)std::min<long>(100, std::max<long>(0, tunable("SPMM_SMEM_PAD_KB", 0))) << 10;
      if (spmm_pad > ((size_t)47 << 10)) {
        cudaError_t e = cudaFuncSetAttribute(spmm_kernel<FID, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 << 10);
        if (e != cudaSuccess) return e;
      }
      const size_t cap = (size_t)std::max<long>(0, tunable("SPMM_WINDOW_KB", 0)) << 10;
      const size_t slice_cap = (size_t)std::max<long>(0, tunable("SPMM_SLICE_KB", 0)) << 10;
      const size_t window = M.n * n_rows * F::BYTES;
      size_t rg = n_rows;
      unsigned Q = 1;
      if (slice_cap && window > slice_cap) rg = std::min(n_rows, std::max<size_t>(8, slice_cap / (M.n * F::BYTES)));
      else if (cap && window > cap) Q = (unsigned)std::min<size_t>((window + cap - 1) / cap, 16);
      if (Q > 1 && (M.seg_q != Q || !M.seg)) {  // cut points, once per (matrix, chunk count)
        if (M.seg) cudaFree(M.seg);
        M.seg = nullptr, M.seg_q = 0;
        cudaError_t e = cudaMalloc(&M.seg, (size_t)(Q + 1) * M.m * 4);
        if (e != cudaSuccess) return e;
        const size_t items = M.m * (Q + 1);
        spmm_segments_kernel<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(M.rowptr, M.colidx, M.m, M.n, Q, M.seg);
        M.seg_q = Q;
        launches++;
      }
      const size_t eff_window = Q > 1 ? window / Q : (rg < n_rows ? M.n * rg * F::BYTES : window);
      const float keep = (!cap || eff_window <= cap) ? 1.0f : (float)((double)cap / (double)eff_window);
      // SPMM_BULK 1: gathers by the bulk-copy engine (spmm_bulk_kernel), 16-byte-multiple elements and whole levels only
      if constexpr (N % 4 == 0) if (tunable("SPMM_BULK", 0) != 0 && Q == 1 && rg == n_rows && M.m && M.nnz) {
        const unsigned rt_count = (unsigned)((n_rows + bulk::CONSUMERS - 1) / bulk::CONSUMERS);
        const unsigned RT = (unsigned)((n_rows + rt_count - 1) / rt_count);
        unsigned G = std::min<unsigned>(bulk::MAX_GROUPS, std::max<unsigned>(1, bulk::CONSUMERS / RT));
        const size_t stage_bytes = ((size_t)bulk::U * RT * F::BYTES + bulk::U * F::BYTES + 127) & ~(size_t)127;
        auto smem_of = [&](unsigned g) { return (((size_t)2 * g * bulk::STAGES * 8 + 127) & ~(size_t)127) + (size_t)g * bulk::STAGES * stage_bytes; };
        while (G > 1 && smem_of(G) > (size_t)100 << 10) G--;
        const size_t smem = smem_of(G);
        if (smem <= (size_t)200 << 10) {
          static bool attr_set_dev[64] = {};
          int dev_id = 0;
          cudaGetDevice(&dev_id);
          if (!attr_set_dev[dev_id & 63]) {
            cudaError_t e = cudaFuncSetAttribute(spmm_bulk_kernel<FID>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 << 10);
            if (e != cudaSuccess) return e;
            attr_set_dev[dev_id & 63] = true;
          }
          const unsigned ctas_per_sm = (unsigned)std::max<size_t>(1, std::min<size_t>(7, ((size_t)220 << 10) / (smem + 1024)));
          const size_t groups_target = (size_t)148 * ctas_per_sm * G * 4;  // about four waves of groups
          const unsigned opg = (unsigned)std::max<size_t>(1, (M.m + groups_target - 1) / groups_target);
          dim3 grid((unsigned)((M.m + (size_t)G * opg - 1) / ((size_t)G * opg)), rt_count);
          spmm_bulk_kernel<FID><<<grid, 32 + bulk::CONSUMERS, smem, st>>>(M.rowptr, M.colidx, M.vals, x, y, (uint32_t)M.m, (uint32_t)n_rows,
                                                                          RT, G, opg);
          launches++;
          cudaError_t e = cudaGetLastError();
          if (e != cudaSuccess) return e;
          continue;
        }
      }
      if (tunable("SPMM_MAC", 1) != 0 && Q == 1 && rg == n_rows && !hints && !pipe && M.m) {
        const size_t items = M.m * n_rows;
        // SPMM_SPLIT (default 1): levels with fewer (output, row) items than a quarter of the GPU's thread slots split
        // every output's non-zeros over 4 lanes, those under an eighth over 8.  Measured (profiles/r02_ab_brakedown_split.jsonl):
        // 2^20 (18 rows) encode 0.161 -> 0.145 ms, a 9-row share of 2^24 0.275 -> 0.256 ms; at 72 rows the 95 K- and
        // 134 K-item levels gain nothing from it (one wave either way), hence the quarter.
        const size_t slots = (size_t)148 * 2048;
        const long split = tunable("SPMM_SPLIT", 1);
        const unsigned G = split <= 0 ? 1 : (split > 1 ? (unsigned)split : (items * 8 <= slots ? 8 : (items * 4 <= slots ? 4 : 1)));
        if (G == 4 || G == 8) {
          const size_t threads = ((items + 32 / G - 1) / (32 / G)) * 32;
          const unsigned grid = (unsigned)((threads + 255) / 256);
          if (G == 4) spmm_sum_split_kernel<FID, 4><<<grid, 256, 0, st>>>(M.rowptr, M.colidx, M.vals, x, y, M.m, n_rows);
          else spmm_sum_split_kernel<FID, 8><<<grid, 256, 0, st>>>(M.rowptr, M.colidx, M.vals, x, y, M.m, n_rows);
        } else if (tunable("SPMM_SUM_DEEP", 0) != 0)
          spmm_sum_kernel<FID, true><<<(unsigned)((items + 255) / 256), 256, spmm_pad, st>>>(M.rowptr, M.colidx, M.vals, x, y, M.m, n_rows);
        else
          spmm_sum_kernel<FID, false><<<(unsigned)((items + 255) / 256), 256, spmm_pad, st>>>(M.rowptr, M.colidx, M.vals, x, y, M.m, n_rows);
        launches++;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        continue;
      }
      for (unsigned q = 0; q < Q; q++) {
        const uint32_t *lo = Q > 1 ? M.seg + (size_t)q * M.m : M.rowptr;
        const uint32_t *hi = Q > 1 ? M.seg + (size_t)(q + 1) * M.m : M.rowptr + 1;
        for (size_t r0 = 0; r0 < n_rows && M.m; r0 += rg) {
          const size_t cnt = std::min(rg, n_rows - r0), items = M.m * cnt;
          const unsigned grid = (unsigned)((items + 255) / 256);
#define LCPC_SPMM_LAUNCH(H, A) \
  spmm_kernel<FID, H, A><<<grid, 256, spmm_pad, st>>>(lo, hi, M.colidx, M.vals, x, y, M.m, n_rows, (unsigned)r0, (unsigned)cnt, keep)
          if (pipe && q) spmm_pipe_kernel<FID, true><<<grid, 256, 0, st>>>(lo, hi, M.colidx, M.vals, x, y, M.m, n_rows, (unsigned)r0, (unsigned)cnt);
          else if (pipe) spmm_pipe_kernel<FID, false><<<grid, 256, 0, st>>>(lo, hi, M.colidx, M.vals, x, y, M.m, n_rows, (unsigned)r0, (unsigned)cnt);
          else if (hints && q) LCPC_SPMM_LAUNCH(true, true);
          else if (hints) LCPC_SPMM_LAUNCH(true, false);
          else if (q) LCPC_SPMM_LAUNCH(false, true);
          else LCPC_SPMM_LAUNCH(false, false);
#undef LCPC_SPMM_LAUNCH
          launches++;
        }
      }
    } else {
      size_t items = op.out_len * n_rows;
      // R mod p (the field's `one`) sits in a dedicated slot right after T (see expander_scratch_bytes)
      uint32_t *one_slot = T + c->tmp_len * n_rows * N;
      one_mont_kernel<FID><<<1, 1, 0, st>>>(one_slot);
      reed_solomon_kernel<FID><<<(unsigned)((items + 127) / 128), 128, 0, st>>>(T, op.in_len, W + op.out_off * n_rows * N,
                                                                               op.out_len, n_rows, one_slot);
      launches += 2;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  // W -> row-major codewords (or per-column-block matrices); positions [0, n_in) are skipped if they already left.
  // dst == nullptr: the caller keeps the codewords in the work buffer (column-major: every column contiguous, which
  // is what column hashing and column openings want) and calls expander_untranspose only if somebody asks for the
  // row-major matrix
  if (dst) {
    const size_t pos0 = early ? c->n_in : 0, n_pos = c->n_cols - pos0;
    dim3 grid((unsigned)((n_rows + TT - 1) / TT), (unsigned)((n_pos + TT - 1) / TT));
    Scatter sc;
    sc.n_blocks = 0;
    if (scatter && scatter->n_blocks) sc = *scatter;
    transpose_kernel<N><<<grid, 256, 0, st>>>(W + pos0 * n_rows * N, n_rows, dst, dst_stride, n_pos, n_rows, sc, nullptr, 0,
                                              ~(size_t)0, pos0);
    launches++;
    if (early) {
      cudaError_t e = cudaStreamWaitEvent(st, side->join, 0);
      if (e != cudaSuccess) return e;
    }
  }
  if (n_launches) *n_launches = launches;
  return cudaGetLastError();
}

template <int N>
static cudaError_t untranspose_impl(const void *scratch, uint32_t *dst, size_t dst_stride, size_t n_rows, size_t n_pos,
                                    cudaStream_t st) {
  const uint32_t *W = (const uint32_t *)scratch;
  dim3 grid((unsigned)((n_rows + TT - 1) / TT), (unsigned)((n_pos + TT - 1) / TT));
  Scatter none;
  none.n_blocks = 0;
  transpose_kernel<N><<<grid, 256, 0, st>>>(W, n_rows, dst, dst_stride, n_pos, n_rows, none, nullptr, 0, ~(size_t)0, 0);
  return cudaGetLastError();
}

// the work buffer of the last encode (W[position][row]) -> row-major dst[row][position] for positions [0, n_pos):
// n_pos = n_cols gives the codewords, n_pos = n_in their systematic part, i.e. the (padded) coefficient rows
cudaError_t expander_untranspose(const ExpanderCode *c, const void *scratch, uint32_t *dst, size_t dst_stride, size_t n_rows,
                                 cudaStream_t st, size_t n_pos) {
  if (!scratch || !dst || n_rows == 0 || n_pos > c->n_cols) return cudaErrorInvalidValue;
  if (n_pos == 0) n_pos = c->n_cols;
  switch (c->field) {
    case FT63: return untranspose_impl<2>(scratch, dst, dst_stride, n_rows, n_pos, st);
    case FT127: return untranspose_impl<4>(scratch, dst, dst_stride, n_rows, n_pos, st);
    case FT191: return untranspose_impl<6>(scratch, dst, dst_stride, n_rows, n_pos, st);
    case FT255: return untranspose_impl<8>(scratch, dst, dst_stride, n_rows, n_pos, st);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t expander_encode_rows(const ExpanderCode *c, const uint32_t *src, size_t src_stride, size_t valid,
                                 uint32_t *dst, size_t dst_stride, size_t n_rows, void *scratch, cudaStream_t st,
                                 int *n_launches, const Scatter *scatter, uint32_t *copy_dst, size_t copy_stride,
                                 size_t src_total, const SideLane *side) {
  switch (c->field) {
    case FT63: return encode_impl<FT63>(c, src, src_stride, valid, dst, dst_stride, n_rows, scratch, st, n_launches, scatter, copy_dst, copy_stride, src_total, side);
    case FT127: return encode_impl<FT127>(c, src, src_stride, valid, dst, dst_stride, n_rows, scratch, st, n_launches, scatter, copy_dst, copy_stride, src_total, side);
    case FT191: return encode_impl<FT191>(c, src, src_stride, valid, dst, dst_stride, n_rows, scratch, st, n_launches, scatter, copy_dst, copy_stride, src_total, side);
    case FT255: return encode_impl<FT255>(c, src, src_stride, valid, dst, dst_stride, n_rows, scratch, st, n_launches, scatter, copy_dst, copy_stride, src_total, side);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace lcpc
