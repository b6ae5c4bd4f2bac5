/* tests/host/multi_gpu_check.c -- ONE process drives N GPUs through include/lcpc_b200.h alone (plain C, no CUDA, no
 * Python): lcpc_b200_commit_new_multi + lcpc_b200_multi_prove against the single-GPU lcpc_b200_commit_new +
 * lcpc_b200_commit_prove on the same coefficients.  This is the call sequence a Rust host makes.
 *
 *   multi_gpu_check <n_gpus> <ligero|sdig> <log2 length> [devices: d0,d1,...]
 * prints "root <hex>", "proof <fnv64 of all proof fields>" for both paths and "ok" when they agree. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/lcpc_b200.h"
#include "../../include/lcpc_b200_host.h"

#define CHECK(call)                                                                 \
  do {                                                                              \
    int rc_ = (call);                                                               \
    if (rc_ != LCPC_B200_OK) {                                                      \
      fprintf(stderr, "%s -> %d (%s)\n", #call, rc_, lcpc_b200_last_error(ctx[0])); \
      return 1;                                                                     \
    }                                                                               \
  } while (0)

static uint64_t fnv(uint64_t h, const void *p, size_t n) {
  const unsigned char *b = (const unsigned char *)p;
  for (size_t i = 0; i < n; i++) h = (h ^ b[i]) * 0x100000001b3ull;
  return h;
}

int main(int argc, char **argv) {
  if (argc < 4) return 2;
  const size_t n_gpus = (size_t)atoi(argv[1]);
  const int sdig = strcmp(argv[2], "sdig") == 0;
  const size_t len = (size_t)1 << atoi(argv[3]);
  int dev[16] = {0};
  if (argc > 4) {
    char *tok = strtok(argv[4], ",");
    for (size_t g = 0; g < n_gpus && tok; g++, tok = strtok(NULL, ",")) dev[g] = atoi(tok);
  }
  const int field = sdig ? LCPC_B200_FT127 : LCPC_B200_FT255;
  const size_t L = (size_t)lcpc_b200_field_limbs(field);
  lcpc_b200_ctx *ctx[17] = {0};
  lcpc_b200_enc *enc[17] = {0};
  /* encodings: one per GPU (+ one more on the first device for the single-GPU twin) */
  size_t n_rows, n_per_row, n_cols;
  lcpc_b200_sdig_code *code = NULL;
  if (sdig) {
    if (lcpc_b200_sdig_choose_n_per_row(field, 3, len, &n_per_row) || lcpc_b200_sdig_code_generate(field, 3, n_per_row, 0, &code)) return 3;
  } else if (lcpc_b200_ligero_get_dims(field, len, 1, 2, &n_rows, &n_per_row, &n_cols)) {
    return 3;
  }
  for (size_t g = 0; g <= n_gpus; g++) {
    if (lcpc_b200_ctx_create(g < n_gpus ? dev[g] : dev[0], &ctx[g])) {
      fprintf(stderr, "no device %d\n", g < n_gpus ? dev[g] : dev[0]);
      return 4;
    }
    if (sdig) CHECK(lcpc_b200_sdig_new_from_code(ctx[g], code, &enc[g]));
    else CHECK(lcpc_b200_ligero_new(ctx[g], field, n_per_row, n_cols, &enc[g]));
  }
  CHECK(lcpc_b200_enc_get_dims(enc[0], len, &n_rows, &n_per_row, &n_cols));
  /* coefficients: small canonical values pushed through the device's own field_op would need a device; any limbs
   * below p are valid Montgomery images, so take a simple pattern with a zero top limb */
  uint64_t *x = NULL;
  CHECK(lcpc_b200_host_alloc(len * L * 8, (void **)&x));
  for (size_t i = 0; i < len; i++)
    for (size_t l = 0; l < L; l++) x[i * L + l] = l + 1 < L ? (i * 0x9E3779B97F4A7C15ull + l) : (i % 1000);
  const size_t ndt = lcpc_b200_n_degree_tests(128, n_cols, lcpc_b200_field_flog2(field));
  const size_t nco = sdig ? lcpc_b200_sdig_n_col_opens(3) : lcpc_b200_ligero_n_col_opens(1, 2);
  size_t path_len = 0;
  while (((size_t)1 << path_len) < n_cols) path_len++;
  uint64_t *outer = (uint64_t *)calloc(n_rows * L, 8);
  for (size_t i = 0; i < n_rows; i++) outer[i * L] = 3 * i + 1;
  uint64_t h[2];
  uint8_t root[2][32];
  for (int which = 0; which < 2; which++) {
    uint64_t *p_eval = (uint64_t *)malloc(n_per_row * L * 8), *p_rand = (uint64_t *)malloc((ndt ? ndt : 1) * n_per_row * L * 8);
    uint64_t *idx = (uint64_t *)malloc(nco * 8), *cols = (uint64_t *)malloc(nco * n_rows * L * 8);
    uint8_t *paths = (uint8_t *)malloc(nco * path_len * 32 + 1);
    lcpc_b200_transcript *tr = NULL;
    CHECK(lcpc_b200_transcript_new((const uint8_t *)"multi gpu check", 15, &tr));
    if (which == 0) {
      lcpc_b200_multi *m = NULL;
      CHECK(lcpc_b200_commit_new_multi(enc, n_gpus, x, len, nco, &m));
      CHECK(lcpc_b200_multi_root(m, root[0]));
      CHECK(lcpc_b200_multi_rerun(m, x, len)); /* a second commit into the same object */
      CHECK(lcpc_b200_multi_root(m, root[0]));
      CHECK(lcpc_b200_multi_prove(m, tr, NULL, outer, n_rows, ndt, nco, p_eval, p_rand, idx, cols, paths));
      lcpc_b200_multi_free(m);
    } else {
      lcpc_b200_commit *c = NULL;
      CHECK(lcpc_b200_commit_new(enc[n_gpus], x, len, &c));
      CHECK(lcpc_b200_commit_root(c, root[1]));
      CHECK(lcpc_b200_commit_prove(c, tr, NULL, outer, n_rows, ndt, nco, p_eval, p_rand, idx, cols, paths));
      lcpc_b200_commit_free(c);
    }
    uint64_t f = 0xcbf29ce484222325ull;
    f = fnv(f, p_eval, n_per_row * L * 8);
    f = fnv(f, p_rand, ndt * n_per_row * L * 8);
    f = fnv(f, idx, nco * 8);
    f = fnv(f, cols, nco * n_rows * L * 8);
    f = fnv(f, paths, nco * path_len * 32);
    h[which] = f;
    printf("%s root ", which == 0 ? "multi " : "single");
    for (int i = 0; i < 32; i++) printf("%02x", root[which][i]);
    printf(" proof %016llx\n", (unsigned long long)f);
    lcpc_b200_transcript_free(tr);
    free(p_eval), free(p_rand), free(idx), free(cols), free(paths);
  }
  free(outer);
  lcpc_b200_host_free(x);
  for (size_t g = 0; g <= n_gpus; g++) lcpc_b200_enc_free(enc[g]), lcpc_b200_ctx_destroy(ctx[g]);
  if (code) lcpc_b200_sdig_code_free(code);
  if (memcmp(root[0], root[1], 32) || h[0] != h[1]) {
    printf("MISMATCH\n");
    return 5;
  }
  printf("ok\n");
  return 0;
}
