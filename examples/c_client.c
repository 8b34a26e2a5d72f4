/* A plain-C client of liblmm.so: the drop-in boundary exercised without Python or torch.
 * BASELINE config 1 (OILMM p=3, m=2, N=50, SEKernel + Matern32): logpdf, posterior, marginals at 7 points,
 * save / load of the posterior, and the identities the reference's tests rely on:
 *   logpdf from the posterior call == logpdf call;  OILMM == ILMM with the dense H = U sqrt(S) (test/oilmm.jl:10-14);
 *   a reloaded posterior predicts bit-identically.
 * Build: gcc -std=c99 -O2 -Iinclude examples/c_client.c -Llinearmixingmodels.jl_b200 -llmm -lm -o examples/c_client
 * Exit code 0 = all checks passed; 77 = no CUDA device (the library has no CPU fallback). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "lmm.h"

#define N 50
#define P 3
#define M 2
#define NS 7

static double frand(unsigned* s) { *s = *s * 1664525u + 1013904223u; return (double)(*s >> 8) / 16777216.0; }

int main(int argc, char** argv) {
  lmm_ctx* ctx = NULL;
  int rc = lmm_ctx_create(0, &ctx);
  if (rc == LMM_E_CUDA) { printf("no CUDA device: liblmm has no CPU fallback\n"); return 77; }
  if (rc != LMM_OK) { printf("lmm_ctx_create failed: %d\n", rc); return 1; }
  printf("%s\n", lmm_version());
  unsigned seed = 12345u;
  double x[N], xs[NS], y[P * N];
  for (int i = 0; i < N; ++i) x[i] = 5.0 * (i + frand(&seed)) / N;
  for (int i = 0; i < NS; ++i) xs[i] = 5.0 * frand(&seed);
  for (int i = 0; i < P * N; ++i) y[i] = sin(3.0 * x[i % N] + i / N) + 0.3 * (frand(&seed) - 0.5);
  /* U (p x m, column-major): two orthonormal columns by Gram-Schmidt; S positive */
  double U[P * M] = {1.0, 2.0, 2.0, 2.0, -1.0, 0.5}, S[M] = {1.7, 0.6};
  double n0 = sqrt(U[0] * U[0] + U[1] * U[1] + U[2] * U[2]);
  for (int j = 0; j < P; ++j) U[j] /= n0;
  double d = U[0] * U[3] + U[1] * U[4] + U[2] * U[5];
  for (int j = 0; j < P; ++j) U[3 + j] -= d * U[j];
  double n1 = sqrt(U[3] * U[3] + U[4] * U[4] + U[5] * U[5]);
  for (int j = 0; j < P; ++j) U[3 + j] /= n1;
  if (lmm_orthogonal_validate(U, P, M) != LMM_OK) { printf("U not orthogonal\n"); return 1; }
  lmm_gp_desc lat[M] = {{LMM_KERNEL_SE, 0, 1.0, 1.0, 0.0, NULL, 0.0}, {LMM_KERNEL_MATERN32, 0, 0.8, 1.3, 0.2, NULL, 0.0}};
  const double sigma2 = 0.1;
  int il = -1, fails = 0;

  double lp = 0.0, terms[M + 1];
  rc = lmm_oilmm_logpdf(ctx, lat, M, x, N, 1, U, S, P, sigma2, y, P, &lp, terms, &il);
  if (rc != LMM_OK) { printf("lmm_oilmm_logpdf: %d (%s)\n", rc, lmm_last_error(ctx)); return 1; }
  printf("logpdf = %.12f  (latent terms %.6f %.6f, regulariser %.6f)\n", lp, terms[0], terms[1], terms[2]);

  lmm_post* post = NULL;
  double lp2 = 0.0;
  rc = lmm_oilmm_posterior(ctx, lat, M, x, N, 1, U, S, P, sigma2, y, P, &post, &lp2, NULL, &il);
  if (rc != LMM_OK) { printf("lmm_oilmm_posterior: %d (%s)\n", rc, lmm_last_error(ctx)); return 1; }
  if (lp2 != lp) { printf("FAIL: logpdf from the shared factorisation differs (%.17g vs %.17g)\n", lp2, lp); ++fails; }

  /* OILMM == ILMM with the dense mixing matrix H = U sqrt(S)  (test/oilmm.jl:10-14) */
  double H[P * M], lpi = 0.0;
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < P; ++j) H[i * P + j] = U[i * P + j] * sqrt(S[i]);
  rc = lmm_ilmm_logpdf(ctx, lat, M, x, N, 1, H, P, sigma2, y, P, LMM_ILMM_FORM_PROJECTED, &lpi, &il);
  if (rc != LMM_OK) { printf("lmm_ilmm_logpdf: %d (%s)\n", rc, lmm_last_error(ctx)); return 1; }
  printf("ILMM logpdf with H = U sqrt(S): %.12f (relative difference %.2e)\n", lpi, fabs(lpi - lp) / fabs(lp));
  if (!(fabs(lpi - lp) <= 1.5e-8 * fabs(lp))) { printf("FAIL: OILMM != ILMM\n"); ++fails; }

  double mean[P * NS], var[P * NS], mean2[P * NS], var2[P * NS];
  rc = lmm_post_mean_and_var(post, xs, NS, sigma2, mean, var);
  if (rc != LMM_OK) { printf("lmm_post_mean_and_var: %d (%s)\n", rc, lmm_last_error(ctx)); return 1; }
  for (int k = 0; k < P * NS; ++k)
    if (!(var[k] > sigma2) || !isfinite(mean[k])) { printf("FAIL: marginal %d is not sane (%g, %g)\n", k, mean[k], var[k]); ++fails; }
  printf("posterior marginal of output 1 at x*=%.3f: mean %.6f, std %.6f\n", xs[0], mean[0], sqrt(var[0]));

  /* serialise, reload, predict again: bit-identical */
  const char* path = argc > 1 ? argv[1] : "/tmp/lmm_c_client_post.bin";
  lmm_post* back = NULL;
  if ((rc = lmm_post_save(post, path)) != LMM_OK || (rc = lmm_post_load(ctx, path, &back)) != LMM_OK) {
    printf("save/load: %d (%s)\n", rc, lmm_last_error(ctx));
    return 1;
  }
  rc = lmm_post_mean_and_var(back, xs, NS, sigma2, mean2, var2);
  if (rc != LMM_OK || memcmp(mean, mean2, sizeof mean) != 0 || memcmp(var, var2, sizeof var) != 0) {
    printf("FAIL: reloaded posterior predicts differently\n");
    ++fails;
  }
  /* errors map to the reference's: wrong out_dim -> "out dim of x != out dim of f." (src/ilmm.jl:52) */
  rc = lmm_oilmm_logpdf(ctx, lat, M, x, N, 1, U, S, P, sigma2, y, P + 1, &lp2, NULL, &il);
  if (rc != LMM_E_OUT_DIM) { printf("FAIL: out_dim mismatch returned %d\n", rc); ++fails; }
  long long launches = 0, h2d = 0, d2h = 0;
  lmm_ctx_counters(ctx, (int64_t*)&launches, (int64_t*)&h2d, (int64_t*)&d2h);
  printf("%lld kernel launches, %lld bytes H2D, %lld bytes D2H\n", launches, h2d, d2h);
  lmm_post_free(back);
  lmm_post_free(post);
  lmm_ctx_destroy(ctx);
  remove(path);
  printf(fails ? "FAILED (%d)\n" : "c_client ok\n", fails);
  return fails ? 1 : 0;
}
