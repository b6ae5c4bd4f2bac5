/*
 * oracle/algo_tmpl.h -- TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
 *
 * Per-field hot loops of the reference's commit/prove path. Included once per field after
 * field_tmpl.h with the same FT / NL macros. Each function cites the reference lines it restates.
 */
#define FN(name) CAT(FT, name)

/* ---------- Ligero encode: lcpc-ligero-pc/src/lib.rs:162-164 -> fffft 0.4 `fft_io_pc` ----------
 * fffft is NOT in /root/reference (Cargo dependency `fffft = "0.4"`, lcpc-ligero-pc/Cargo.toml:13).
 * Published algorithm restated: in-order input, bit-reversed output,
 * Gentleman-Sande decimation in frequency with w = root_of_unity()^(2^(S-log_len)):
 *   for gap = N/2, N/4, .., 1:  (a, b) <- (a + b, (a - b) * w^(idx * N/(2 gap)))
 * so position i ends up holding X[bitrev(i)], X[k] = sum_j x_j w^(jk).
 * PARITY UNPINNED by the reference (its tests only check linearity and ifft_oi(fft_io(x)) == x,
 * lcpc-2d/src/tests.rs:193-236); tests/ pin this function against a direct big-int DFT.
 */
static void FN(roots_of_unity)(uint64_t *roots /* (len/2) * NL */, uint32_t log_len, int inverse) {
  uint64_t w[NL];
  memcpy(w, FN(ROU), sizeof w);
  for (uint32_t i = 0; i < FN(S) - log_len; i++) FN(mul)(w, w, w);
  if (inverse) FN(inv)(w, w);
  size_t half = log_len ? ((size_t)1 << (log_len - 1)) : 0;
  uint64_t cur[NL];
  memcpy(cur, FN(R), sizeof cur);
  for (size_t i = 0; i < half; i++) {
    memcpy(roots + i * NL, cur, sizeof cur);
    FN(mul)(cur, cur, w);
  }
}

static void FN(fft_io)(uint64_t *x, uint32_t log_len, const uint64_t *roots) {
  size_t len = (size_t)1 << log_len;
  for (size_t gap = len / 2; gap > 0; gap /= 2) {
    size_t nchunks = len / (2 * gap);
    for (size_t cidx = 0; cidx < nchunks; cidx++) {
      size_t offset = 2 * cidx * gap;
      for (size_t idx = 0; idx < gap; idx++) {
        uint64_t *a = x + (offset + idx) * NL;
        uint64_t *b = x + (offset + idx + gap) * NL;
        uint64_t neg[NL];
        FN(sub)(neg, a, b);
        FN(add)(a, a, b);
        FN(mul)(b, neg, roots + nchunks * idx * NL);
      }
    }
  }
}

/* inverse of fft_io (fffft `ifft_oi`): bit-reversed input, in-order output, scaled by 1/N.
 * Only used by tests (round trip, lcpc-2d/src/tests.rs:226-233). roots must be the INVERSE roots. */
static void FN(ifft_oi)(uint64_t *x, uint32_t log_len, const uint64_t *inv_roots) {
  size_t len = (size_t)1 << log_len;
  for (size_t gap = 1; gap < len; gap *= 2) {
    size_t nchunks = len / (2 * gap);
    for (size_t cidx = 0; cidx < nchunks; cidx++) {
      size_t offset = 2 * cidx * gap;
      for (size_t idx = 0; idx < gap; idx++) {
        uint64_t *a = x + (offset + idx) * NL;
        uint64_t *b = x + (offset + idx + gap) * NL;
        uint64_t t[NL];
        FN(mul)(t, b, inv_roots + nchunks * idx * NL);
        FN(sub)(b, a, t);
        FN(add)(a, a, t);
      }
    }
  }
  uint64_t n[NL], ninv[NL];
  memset(n, 0, sizeof n);
  n[0] = (uint64_t)len;
  FN(to_mont)(n, n);
  FN(inv)(ninv, n);
  for (size_t i = 0; i < len; i++) FN(mul)(x + i * NL, x + i * NL, ninv);
}

/* ---------- hash_columns base case: lcpc-2d/src/lib.rs:716-735 ----------
 * One digest per column, started with 32 zero bytes (:722-723), fed row by row (:727-731) with
 * to_repr of the element (FieldHash::digest_update :42-44), finalized into hashes[col] (:733-735).
 */
static void FN(hash_column_block)(const uint64_t *comm, uint8_t *hashes /* 32 B each */,
                                  size_t n_rows, size_t n_cols, size_t offset, size_t count) {
  b3_hasher dig[1 << LCPC_LOG_MIN_NCOLS];
  static const uint8_t zeros[32] = {0};
  for (size_t c = 0; c < count; c++) {
    b3_init(&dig[c]);
    b3_update(&dig[c], zeros, 32);
  }
  for (size_t row = 0; row < n_rows; row++) {
    for (size_t c = 0; c < count; c++) {
      uint8_t repr[8 * NL];
      FN(to_repr)(repr, comm + (row * n_cols + offset + c) * NL);
      b3_update(&dig[c], repr, 8 * NL);
    }
  }
  for (size_t c = 0; c < count; c++) b3_finalize(&dig[c], hashes + 32 * c);
}

/* hash_columns recursion (:736-744) flattened: the rayon::join tree bottoms out in blocks of
 * <= 32 columns obtained by repeated halving; results do not depend on the split. */
static void FN(hash_columns)(const uint64_t *comm, uint8_t *hashes, size_t n_rows, size_t n_cols,
                             size_t offset, size_t count) {
  if (count <= ((size_t)1 << LCPC_LOG_MIN_NCOLS)) {
    FN(hash_column_block)(comm, hashes, n_rows, n_cols, offset, count);
  } else {
    size_t half = count / 2;
#pragma omp task default(shared) if (count > 64)
    FN(hash_columns)(comm, hashes, n_rows, n_cols, offset, half);
#pragma omp task default(shared) if (count > 64)
    FN(hash_columns)(comm, hashes + 32 * half, n_rows, n_cols, offset + half, count - half);
#pragma omp taskwait
  }
}

/* ---------- collapse_columns: lcpc-2d/src/lib.rs:1095-1123 ----------
 * poly[col] += coeffs[row*n_per_row + offset + col] * tensor[row]; poly is NOT cleared here
 * (callers pre-zero it, :1033,1054). Same halving recursion as hash_columns.
 */
static void FN(collapse_columns)(const uint64_t *coeffs, const uint64_t *tensor, uint64_t *poly,
                                 size_t n_rows, size_t n_per_row, size_t offset, size_t count) {
  if (count <= ((size_t)1 << LCPC_LOG_MIN_NCOLS)) {
    for (size_t row = 0; row < n_rows; row++) {
      const uint64_t *tv = tensor + row * NL;
      for (size_t col = 0; col < count; col++) {
        uint64_t prod[NL];
        FN(mul)(prod, coeffs + (row * n_per_row + offset + col) * NL, tv);
        FN(add)(poly + col * NL, poly + col * NL, prod);
      }
    }
  } else {
    size_t half = count / 2;
#pragma omp task default(shared) if (count > 64)
    FN(collapse_columns)(coeffs, tensor, poly, n_rows, n_per_row, offset, half);
#pragma omp task default(shared) if (count > 64)
    FN(collapse_columns)(coeffs, tensor, poly + half * NL, n_rows, n_per_row, offset + half,
                         count - half);
#pragma omp taskwait
  }
}

/* serial twin: eval_outer_ser, lcpc-2d/src/lib.rs:1205-1226 */
static void FN(collapse_columns_ser)(const uint64_t *coeffs, const uint64_t *tensor, uint64_t *poly,
                                     size_t n_rows, size_t n_per_row) {
  for (size_t row = 0; row < n_rows; row++)
    for (size_t col = 0; col < n_per_row; col++) {
      uint64_t prod[NL];
      FN(mul)(prod, coeffs + (row * n_per_row + col) * NL, tensor + row * NL);
      FN(add)(poly + col * NL, poly + col * NL, prod);
    }
}

/* ---------- Brakedown: sparse code matrices and encode ----------
 * CsMat::new_csc((m, n), ptrs, idxs, data) (matgen.rs:187): m output rows, n input columns.
 * `mat.dot(x)` (encode.rs:51-52,65-66,83-84; un-vendored sprs 0.10):
 *   y[i] = sum_j M[i,j] x[j], scatter over the CSC columns. Exact arithmetic: order-independent.
 */
static void FN(csc_dot)(const lcpc_csc *M, const uint64_t *x, uint64_t *y) {
  memset(y, 0, M->m * NL * sizeof(uint64_t));
  for (size_t j = 0; j < M->n; j++) {
    const uint64_t *xj = x + j * NL;
    for (uint64_t k = M->ptrs[j]; k < M->ptrs[j + 1]; k++) {
      uint64_t prod[NL];
      uint64_t *yi = y + M->idxs[k] * NL;
      FN(mul)(prod, M->data + k * NL, xj);
      FN(add)(yi, yi, prod);
    }
  }
}

/* reed_solomon: lcpc-brakedown-pc/src/encode.rs:97-110 (Horner at the points 1, 2, 3, ...) */
static void FN(reed_solomon)(const uint64_t *xi, size_t n_in, uint64_t *xo, size_t n_out) {
  uint64_t x[NL];
  memcpy(x, FN(R), sizeof x);
  for (size_t r = 0; r < n_out; r++) {
    uint64_t *acc = xo + r * NL;
    memset(acc, 0, NL * sizeof(uint64_t));
    for (size_t j = n_in; j-- > 0;) {
      FN(mul)(acc, acc, x);
      FN(add)(acc, acc, xi + j * NL);
    }
    FN(add)(x, x, FN(R));
  }
}

/* encode: lcpc-brakedown-pc/src/encode.rs:36-94. xi has codeword_length entries, in place. */
static int FN(sdig_encode)(uint64_t *xi, size_t xi_len, const lcpc_csc *pre, const lcpc_csc *post,
                           size_t n_levels) {
  if (xi_len != lcpc_codeword_length(pre, post, n_levels)) return -1; /* assert at :42 */
  size_t in_start = 0;
  size_t max_rows = 0;
  for (size_t i = 0; i < n_levels; i++) {
    if (pre[i].m > max_rows) max_rows = pre[i].m;
    if (post[i].m > max_rows) max_rows = post[i].m;
  }
  uint64_t *tmp = (uint64_t *)malloc((max_rows + 1) * NL * sizeof(uint64_t));
  if (!tmp) return -2;
  /* precodes all the way down (:45-58) */
  for (size_t i = 0; i + 1 < n_levels; i++) {
    size_t in_end = in_start + pre[i].n;
    FN(csc_dot)(&pre[i], xi + in_start * NL, tmp);
    memcpy(xi + in_end * NL, tmp, pre[i].m * NL * sizeof(uint64_t));
    in_start = in_end;
  }
  /* base case (:61-74): last precode into temporary storage, then Reed-Solomon */
  size_t out_start;
  {
    const lcpc_csc *pc = &pre[n_levels - 1];
    size_t in_end = in_start + pc->n;
    FN(csc_dot)(pc, xi + in_start * NL, tmp);
    size_t out_end = in_end + post[n_levels - 1].n;
    FN(reed_solomon)(tmp, pc->m, xi + in_end * NL, out_end - in_end);
    in_start = in_end + pc->m;
    out_start = out_end;
  }
  /* postcodes back up (:76-90) */
  for (size_t i = n_levels; i-- > 0;) {
    in_start -= pre[i].m;
    if (out_start - in_start != post[i].n) {
      free(tmp);
      return -3; /* sprs would panic on the dimension mismatch */
    }
    FN(csc_dot)(&post[i], xi + in_start * NL, tmp);
    memcpy(xi + out_start * NL, tmp, post[i].m * NL * sizeof(uint64_t));
    out_start += post[i].m;
  }
  free(tmp);
  if (in_start != pre[0].n || out_start != xi_len) return -4; /* asserts at :92-93 */
  return 0;
}

/* gen_code: lcpc-brakedown-pc/src/matgen.rs:114-188. n columns, each with d distinct sorted row
 * indices in [0,m) and non-zero random values, drawn from one RNG stream in this exact order. */
static int FN(gen_code)(lcpc_csc *M, size_t n, size_t m, size_t d, chacha_rng *rng) {
  M->m = m;
  M->n = n;
  M->ptrs = (uint64_t *)malloc((n + 1) * sizeof(uint64_t));
  M->idxs = (uint64_t *)malloc((d * n + 1) * sizeof(uint64_t));
  M->data = (uint64_t *)malloc((d * n + 1) * NL * sizeof(uint64_t));
  if (!M->ptrs || !M->idxs || !M->data) return -1;
  uint64_t *cols = (uint64_t *)malloc((d + 1) * sizeof(uint64_t));
  size_t nnz = 0;
  M->ptrs[0] = 0;
  for (size_t c = 0; c < n; c++) {
    size_t have = 0;
    while (have < d) { /* rejection of repeats, quadratic scan (:144-159) */
      uint64_t x = chacha_uniform(rng, m);
      int seen = 0;
      for (size_t k = 0; k < have; k++) seen |= (cols[k] == x);
      if (!seen) cols[have++] = x;
    }
    /* sort_unstable (:160); values are distinct so any sort gives the same order */
    for (size_t a = 1; a < d; a++) {
      uint64_t v = cols[a];
      size_t b = a;
      while (b > 0 && cols[b - 1] > v) {
        cols[b] = cols[b - 1];
        b--;
      }
      cols[b] = v;
    }
    for (size_t k = 0; k < d; k++) { /* (:166-183); the repeat check never fires on distinct cols */
      uint64_t *val = M->data + nnz * NL;
      do {
        FN(random)(val, chacha_next_u64, rng);
      } while (FN(is_zero)(val));
      M->idxs[nnz] = cols[k];
      nnz++;
    }
    M->ptrs[c + 1] = nnz;
  }
  free(cols);
  return 0;
}

#undef FN
