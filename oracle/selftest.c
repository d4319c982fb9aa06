/* selftest.c -- drives every oracle entry point on small inputs; built with
 * -fsanitize=address,undefined (make -C oracle selftest_asan) so that the checker itself is
 * checked for out-of-bounds accesses, leaks and undefined behaviour.  Test infrastructure only.
 * Also pins the SURVEY.md 8c known-answer vector (n = 6, k = 3) from C.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int oracle_knn_brute(const double *, int64_t, const double *, int64_t, int, int, int32_t *, double *, int);
int oracle_knn_kdtree(const double *, int64_t, const double *, int64_t, int, int, int32_t *, double *, int, double *,
                      double *);
int oracle_nn_graph(const int32_t *, int64_t, int, int64_t, int64_t *, int32_t *, int32_t *, double *);
int oracle_compute_snn(const int32_t *, int64_t, int, double, int, void **, int64_t *);
int oracle_compute_snn_rows(const int32_t *, int64_t, int, double, int64_t, int64_t, int, void **, int64_t *, double *);
void oracle_snn_fetch(void *, int64_t *, int32_t *, double *);
void oracle_l2norm(double *, int64_t, int);

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static double rnd(void) {
  rng_state ^= rng_state << 13;
  rng_state ^= rng_state >> 7;
  rng_state ^= rng_state << 17;
  return (double)(rng_state >> 11) / 9007199254740992.0 - 0.5;
}

#define CHECK(c)                                                      \
  do {                                                                \
    if (!(c)) {                                                       \
      fprintf(stderr, "selftest: %s failed (line %d)\n", #c, __LINE__); \
      return 1;                                                       \
    }                                                                 \
  } while (0)

int main(void) {
  /* 1. known-answer SNN */
  const int32_t nn6[18] = {1, 2, 3, 2, 1, 3, 3, 2, 4, 4, 5, 3, 5, 4, 3, 6, 5, 4};
  const int64_t p_want[7] = {0, 5, 10, 16, 22, 28, 32};
  void *h = NULL;
  int64_t nnz = 0;
  CHECK(oracle_compute_snn(nn6, 6, 3, 1.0 / 15, 1, &h, &nnz) == 0);
  CHECK(nnz == 32);
  int64_t p[7];
  int32_t ii[32];
  double xx[32];
  oracle_snn_fetch(h, p, ii, xx);
  CHECK(memcmp(p, p_want, sizeof(p)) == 0);
  CHECK(ii[0] == 0 && ii[4] == 4 && xx[0] == 1.0 && xx[2] == 0.5 && xx[3] == 0.2 && ii[31] == 5 && xx[31] == 1.0);

  /* 2. kNN: brute force vs kd-tree on random data, self and query != reference, ragged sizes */
  const int sizes[][3] = {{1, 1, 1}, {7, 3, 7}, {129, 5, 10}, {300, 12, 20}};
  for (unsigned s = 0; s < sizeof(sizes) / sizeof(sizes[0]); ++s) {
    const int n = sizes[s][0], d = sizes[s][1], k = sizes[s][2], nq = n / 2 + 1;
    double *X = malloc(sizeof(double) * (size_t)n * d), *Q = malloc(sizeof(double) * (size_t)nq * d);
    for (int t = 0; t < n * d; ++t) X[t] = rnd();
    for (int t = 0; t < nq * d; ++t) Q[t] = rnd();
    int32_t *i1 = malloc(sizeof(int32_t) * (size_t)n * k), *i2 = malloc(sizeof(int32_t) * (size_t)n * k);
    double *d1 = malloc(sizeof(double) * (size_t)n * k), *d2 = malloc(sizeof(double) * (size_t)n * k);
    CHECK(oracle_knn_brute(X, n, X, n, d, k, i1, d1, 2) == 0);
    CHECK(oracle_knn_kdtree(X, n, X, n, d, k, i2, d2, 2, NULL, NULL) == 0);
    CHECK(memcmp(d1, d2, sizeof(double) * (size_t)n * k) == 0);
    CHECK(memcmp(i1, i2, sizeof(int32_t) * (size_t)n * k) == 0);
    CHECK(oracle_knn_brute(X, n, Q, nq, d, k, i1, d1, 1) == 0);
    CHECK(oracle_knn_kdtree(X, n, Q, nq, d, k, i2, d2, 1, NULL, NULL) == 0);
    CHECK(memcmp(d1, d2, sizeof(double) * (size_t)nq * k) == 0);
    CHECK(oracle_knn_brute(X, n, X, n, d, n + 1, i1, d1, 1) == 1); /* k > nr is refused */
    /* 3. graphs on that index: nn Graph two-call protocol, SNN in full and by row range */
    CHECK(oracle_knn_brute(X, n, X, n, d, k, i1, d1, 1) == 0);
    int64_t nnz_nn = 0;
    CHECK(oracle_nn_graph(i1, n, k, n, &nnz_nn, NULL, NULL, NULL) == 0);
    CHECK(nnz_nn == (int64_t)n * k);
    int32_t *gp = malloc(sizeof(int32_t) * (size_t)(n + 1)), *gi = malloc(sizeof(int32_t) * (size_t)nnz_nn);
    double *gx = malloc(sizeof(double) * (size_t)nnz_nn);
    CHECK(oracle_nn_graph(i1, n, k, n, &nnz_nn, gp, gi, gx) == 0);
    CHECK(gp[n] == nnz_nn);
    const double prunes[3] = {0.0, 1.0 / 15, 1.0};
    for (int pi = 0; pi < 3; ++pi) {
      int64_t nz_full = 0, nz_a = 0, nz_b = 0;
      CHECK(oracle_compute_snn(i1, n, k, prunes[pi], 2, &h, &nz_full) == 0);
      oracle_snn_fetch(h, NULL, NULL, NULL);
      CHECK(oracle_compute_snn_rows(i1, n, k, prunes[pi], 0, n / 2, 1, &h, &nz_a, NULL) == 0);
      oracle_snn_fetch(h, NULL, NULL, NULL);
      CHECK(oracle_compute_snn_rows(i1, n, k, prunes[pi], n / 2, n, 1, &h, &nz_b, NULL) == 0);
      oracle_snn_fetch(h, NULL, NULL, NULL);
      CHECK(nz_a + nz_b == nz_full);
    }
    oracle_l2norm(X, n, d);
    double s2 = 0;
    for (int j = 0; j < d; ++j) s2 += X[j] * X[j];
    CHECK(fabs(s2 - 1.0) < 1e-12);
    free(X); free(Q); free(i1); free(i2); free(d1); free(d2); free(gp); free(gi); free(gx);
  }
  /* 4. an index outside [1, n] is reported, not dereferenced */
  int32_t bad[6] = {1, 2, 3, 1, 2, 9};
  CHECK(oracle_compute_snn(bad, 2, 3, 0.0, 1, &h, &nnz) != 0);
  printf("oracle selftest ok\n");
  return 0;
}
