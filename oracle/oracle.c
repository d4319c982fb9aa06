/*
 * oracle.c -- CPU restatement of Seurat's FindNeighbors hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (seurat_b200/, the
 * C-ABI library libb200nn.so) may import, link or execute this file.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs use it, and there only as the checker / the timed CPU arm.
 *
 * PARITY STATUS: "parity unpinned" at the third-party boundaries.  The
 * reference's own tests hold no golden vector for nn.method="rann", nn2 or
 * ComputeSNN, and neither R, RANN/ANN nor Eigen
 * exist in the build container, so the reference itself cannot be run.  The
 * pins used instead are (i) the known-answer vectors of SURVEY.md 8c, frozen
 * under tests/golden/, and (ii) independent cross-checks against
 * scipy.sparse (A @ A.T) and sklearn brute-force kNN in tests/.
 *
 * What each function follows:
 *   oracle_compute_snn     /root/reference/src/snn.cpp:14-38 (ComputeSNN)
 *   oracle_nn_graph        /root/reference/R/clustering.R:664-670
 *   oracle_knn_brute       RANN::nn2 semantics (call site
 *                          /root/reference/R/clustering.R:1664-1667): exact
 *                          Euclidean kNN in double, sequential sum over
 *                          dimensions, ascending distance, 1-based indices,
 *                          sqrt at the end.  Ties ordered by lower index.
 *   oracle_knn_kdtree      the published ANN 1.1.x algorithm RANN wraps
 *                          (Mount & Arya): kd-tree, bucket size 1, sliding
 *                          midpoint split, standard (depth-first) k-search
 *                          with incremental box distances and partial-distance
 *                          early exit, eps = 0.  Restated from the published
 *                          description; RANN is an un-vendored, un-pinned CRAN
 *                          dependency (/root/reference/DESCRIPTION:72).
 *   oracle_l2norm          /root/reference/R/utilities.R:2107-2122 (L2Norm)
 *   oracle_write_edge_file / oracle_fast_dist / oracle_snn_smallest_nonzero_dist
 *                          /root/reference/src/snn.cpp:40-59, 82-126; src/fast_NN_dist.cpp:8-69
 *
 * Build:  make -C oracle        (gcc -O2 -ffp-contract=off -fopenmp)
 * All matrices are row-major (one cell per row) unless stated otherwise.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ */
/* kNN, brute force                                                    */
/* ------------------------------------------------------------------ */

/* squared distance, summed in dimension order exactly as ANN's leaf loop
 * does (dist += t*t, one rounding for the product and one for the sum;
 * compile with -ffp-contract=off so no FMA is formed). */
static inline double sqdist(const double *q, const double *p, int d) {
  double dist = 0.0;
  for (int j = 0; j < d; ++j) {
    double t = q[j] - p[j];
    dist = dist + t * t;
  }
  return dist;
}

/* keep the k smallest (key, idx) pairs, ascending by key then by idx.
 * Points arrive in ascending idx order, so "insert after equal keys"
 * realises the (d2, index) order. */
static inline void topk_insert(double *key, int32_t *val, int k, int *n,
                               double kv, int32_t iv) {
  int i = *n;
  if (i == k) {
    if (!(kv < key[k - 1])) return;
    i = k - 1;
  } else {
    (*n)++;
  }
  for (; i > 0; --i) {
    if (key[i - 1] > kv) {
      key[i] = key[i - 1];
      val[i] = val[i - 1];
    } else
      break;
  }
  key[i] = kv;
  val[i] = iv;
}

/*
 * Exact kNN by exhaustive search.
 *   ref [nr x d], qry [nq x d] row-major doubles; qry may equal ref.
 *   idx_out [nq x k] row-major, 1-based (RANN returns index+1)
 *   dist_out[nq x k] row-major, sqrt of the squared distance
 * Returns 0, or 1 when k > nr (RANN: "Cannot find more nearest neighbours
 * than there are points").
 */
ORACLE_API int oracle_knn_brute(const double *ref, int64_t nr, const double *qry,
                                int64_t nq, int d, int k, int32_t *idx_out,
                                double *dist_out, int nthreads) {
  if (k > nr || k <= 0) return 1;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
  {
    double *key = (double *)malloc(sizeof(double) * (size_t)(k + 1));
    int32_t *val = (int32_t *)malloc(sizeof(int32_t) * (size_t)(k + 1));
#pragma omp for schedule(dynamic, 16)
    for (int64_t qi = 0; qi < nq; ++qi) {
      const double *q = qry + qi * d;
      int n = 0;
      for (int64_t ri = 0; ri < nr; ++ri) {
        double d2 = sqdist(q, ref + ri * d, d);
        topk_insert(key, val, k, &n, d2, (int32_t)ri);
      }
      for (int j = 0; j < k; ++j) {
        idx_out[qi * k + j] = val[j] + 1;
        dist_out[qi * k + j] = sqrt(key[j]);
      }
    }
    free(key);
    free(val);
  }
  return 0;
}

/* ------------------------------------------------------------------ */
/* kNN, ANN-style kd-tree (the timed CPU baseline)                     */
/* ------------------------------------------------------------------ */

typedef struct {
  int32_t cut_dim;   /* -1 => leaf */
  int32_t left, right; /* child node ids, or for a leaf: left = point index (-1 if empty) */
  double cut_val;
  double lo_bnd, hi_bnd; /* bounds of the cell along cut_dim */
} kd_node;

typedef struct {
  const double *pts;
  int64_t n;
  int d;
  kd_node *nodes;
  int64_t n_nodes, cap_nodes;
  int32_t *pidx;
  int32_t root;
  double *bnd_lo, *bnd_hi; /* bounding box of all points */
} kd_tree;

static int32_t kd_new_node(kd_tree *t) {
  if (t->n_nodes == t->cap_nodes) {
    t->cap_nodes = t->cap_nodes ? t->cap_nodes * 2 : 1024;
    t->nodes = (kd_node *)realloc(t->nodes, sizeof(kd_node) * (size_t)t->cap_nodes);
  }
  return (int32_t)(t->n_nodes++);
}

#define KD_ERR 0.001 /* ANN's sl_midpt_split tolerance on the longest side */

/* sliding-midpoint split of pidx[lo,hi) inside box [blo,bhi]; returns the
 * number of points placed on the low side and fills cut_dim/cut_val. */
static int64_t kd_sl_midpt_split(kd_tree *t, int64_t lo, int64_t hi, double *blo,
                                 double *bhi, int *cut_dim, double *cut_val) {
  const int d = t->d;
  const double *pts = t->pts;
  int32_t *pidx = t->pidx;
  double max_length = bhi[0] - blo[0];
  for (int j = 1; j < d; ++j) {
    double len = bhi[j] - blo[j];
    if (len > max_length) max_length = len;
  }
  /* among the (nearly) longest sides pick the one with the widest point spread */
  double max_spread = -1.0;
  int cd = 0;
  for (int j = 0; j < d; ++j) {
    if ((bhi[j] - blo[j]) >= (1.0 - KD_ERR) * max_length) {
      double mn = pts[(int64_t)pidx[lo] * d + j], mx = mn;
      for (int64_t i = lo + 1; i < hi; ++i) {
        double c = pts[(int64_t)pidx[i] * d + j];
        if (c < mn) mn = c; else if (c > mx) mx = c;
      }
      double spr = mx - mn;
      if (spr > max_spread) { max_spread = spr; cd = j; }
    }
  }
  double ideal = (blo[cd] + bhi[cd]) / 2.0;
  double mn = pts[(int64_t)pidx[lo] * d + cd], mx = mn;
  for (int64_t i = lo + 1; i < hi; ++i) {
    double c = pts[(int64_t)pidx[i] * d + cd];
    if (c < mn) mn = c; else if (c > mx) mx = c;
  }
  double cv = ideal;
  if (ideal < mn) cv = mn; else if (ideal > mx) cv = mx; /* slide to the points */
  /* three-way partition: < cv | == cv | > cv */
  int64_t l = lo, r = hi - 1;
  for (;;) {
    while (l < hi && pts[(int64_t)pidx[l] * d + cd] < cv) l++;
    while (r >= lo && pts[(int64_t)pidx[r] * d + cd] >= cv) r--;
    if (l > r) break;
    int32_t tmp = pidx[l]; pidx[l] = pidx[r]; pidx[r] = tmp; l++; r--;
  }
  int64_t br1 = l; /* [lo,br1) < cv */
  r = hi - 1;
  for (;;) {
    while (l < hi && pts[(int64_t)pidx[l] * d + cd] <= cv) l++;
    while (r >= br1 && pts[(int64_t)pidx[r] * d + cd] > cv) r--;
    if (l > r) break;
    int32_t tmp = pidx[l]; pidx[l] = pidx[r]; pidx[r] = tmp; l++; r--;
  }
  int64_t br2 = l; /* [br1,br2) == cv */
  int64_t n = hi - lo, n_lo;
  if (ideal < mn) n_lo = 1;
  else if (ideal > mx) n_lo = n - 1;
  else if (br1 - lo > n / 2) n_lo = br1 - lo;
  else if (br2 - lo < n / 2) n_lo = br2 - lo;
  else n_lo = n / 2;
  *cut_dim = cd;
  *cut_val = cv;
  return n_lo;
}

static int32_t kd_build_rec(kd_tree *t, int64_t lo, int64_t hi, double *blo, double *bhi) {
  int64_t n = hi - lo;
  int32_t id = kd_new_node(t);
  if (n <= 1) { /* bucket size 1 */
    t->nodes[id].cut_dim = -1;
    t->nodes[id].left = n == 1 ? t->pidx[lo] : -1;
    t->nodes[id].right = -1;
    return id;
  }
  int cd; double cv;
  int64_t n_lo = kd_sl_midpt_split(t, lo, hi, blo, bhi, &cd, &cv);
  double lv = blo[cd], hv = bhi[cd];
  bhi[cd] = cv;
  int32_t L = kd_build_rec(t, lo, lo + n_lo, blo, bhi);
  bhi[cd] = hv;
  blo[cd] = cv;
  int32_t R = kd_build_rec(t, lo + n_lo, hi, blo, bhi);
  blo[cd] = lv;
  kd_node *nd = &t->nodes[id];
  nd->cut_dim = cd; nd->cut_val = cv; nd->lo_bnd = lv; nd->hi_bnd = hv;
  nd->left = L; nd->right = R;
  return id;
}

static kd_tree *kd_build(const double *pts, int64_t n, int d) {
  kd_tree *t = (kd_tree *)calloc(1, sizeof(kd_tree));
  t->pts = pts; t->n = n; t->d = d;
  t->pidx = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
  for (int64_t i = 0; i < n; ++i) t->pidx[i] = (int32_t)i;
  t->bnd_lo = (double *)malloc(sizeof(double) * (size_t)d);
  t->bnd_hi = (double *)malloc(sizeof(double) * (size_t)d);
  for (int j = 0; j < d; ++j) {
    double mn = pts[j], mx = pts[j];
    for (int64_t i = 1; i < n; ++i) {
      double c = pts[i * d + j];
      if (c < mn) mn = c; else if (c > mx) mx = c;
    }
    t->bnd_lo[j] = mn; t->bnd_hi[j] = mx;
  }
  double *blo = (double *)malloc(sizeof(double) * (size_t)d);
  double *bhi = (double *)malloc(sizeof(double) * (size_t)d);
  memcpy(blo, t->bnd_lo, sizeof(double) * (size_t)d);
  memcpy(bhi, t->bnd_hi, sizeof(double) * (size_t)d);
  t->root = kd_build_rec(t, 0, n, blo, bhi);
  free(blo); free(bhi);
  return t;
}

static void kd_free(kd_tree *t) {
  free(t->nodes); free(t->pidx); free(t->bnd_lo); free(t->bnd_hi); free(t);
}

typedef struct {
  const kd_tree *t;
  const double *q;
  int k, n;
  double *key;
  int32_t *val;
} kd_query;

static void kd_search_rec(kd_query *s, int32_t id, double box_dist) {
  const kd_tree *t = s->t;
  const kd_node *nd = &t->nodes[id];
  if (nd->cut_dim < 0) {
    int32_t p = nd->left;
    if (p < 0) return;
    double min_dist = s->n == s->k ? s->key[s->k - 1] : INFINITY;
    const double *pp = t->pts + (int64_t)p * t->d;
    double dist = 0.0;
    int j;
    for (j = 0; j < t->d; ++j) { /* partial-distance early exit */
      double tt = s->q[j] - pp[j];
      dist = dist + tt * tt;
      if (dist > min_dist) break;
    }
    if (j >= t->d) { /* self matches allowed */
      /* insert keeping ascending keys; equal keys keep visit order */
      int i = s->n;
      if (i == s->k) {
        if (!(dist < s->key[s->k - 1])) return;
        i = s->k - 1;
      } else s->n++;
      for (; i > 0; --i) {
        if (s->key[i - 1] > dist) { s->key[i] = s->key[i - 1]; s->val[i] = s->val[i - 1]; }
        else break;
      }
      s->key[i] = dist; s->val[i] = p;
    }
    return;
  }
  double cut_diff = s->q[nd->cut_dim] - nd->cut_val;
  if (cut_diff < 0) {
    kd_search_rec(s, nd->left, box_dist);
    double box_diff = nd->lo_bnd - s->q[nd->cut_dim];
    if (box_diff < 0) box_diff = 0;
    box_dist = box_dist + (cut_diff * cut_diff - box_diff * box_diff);
    double kth = s->n == s->k ? s->key[s->k - 1] : INFINITY;
    if (box_dist < kth) kd_search_rec(s, nd->right, box_dist); /* eps = 0 */
  } else {
    kd_search_rec(s, nd->right, box_dist);
    double box_diff = s->q[nd->cut_dim] - nd->hi_bnd;
    if (box_diff < 0) box_diff = 0;
    box_dist = box_dist + (cut_diff * cut_diff - box_diff * box_diff);
    double kth = s->n == s->k ? s->key[s->k - 1] : INFINITY;
    if (box_dist < kth) kd_search_rec(s, nd->left, box_dist);
  }
}

/*
 * ANN-style exact kNN.  Same outputs as oracle_knn_brute except that equal
 * distances come out in tree-visit order (RANN's tie order is
 * implementation-defined).  build_seconds/search_seconds (may be NULL) report
 * the two phases so that a sub-sampled search can be extrapolated.
 */
ORACLE_API int oracle_knn_kdtree(const double *ref, int64_t nr, const double *qry,
                                 int64_t nq, int d, int k, int32_t *idx_out,
                                 double *dist_out, int nthreads,
                                 double *build_seconds, double *search_seconds) {
  if (k > nr || k <= 0) return 1;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
  double t0 = omp_get_wtime();
#endif
  kd_tree *t = kd_build(ref, nr, d);
#ifdef _OPENMP
  double t1 = omp_get_wtime();
#endif
#pragma omp parallel
  {
    kd_query s;
    s.t = t; s.k = k;
    s.key = (double *)malloc(sizeof(double) * (size_t)(k + 1));
    s.val = (int32_t *)malloc(sizeof(int32_t) * (size_t)(k + 1));
#pragma omp for schedule(dynamic, 16)
    for (int64_t qi = 0; qi < nq; ++qi) {
      s.q = qry + qi * d; s.n = 0;
      double box_dist = 0.0; /* distance from q to the root bounding box */
      for (int j = 0; j < d; ++j) {
        double df = 0;
        if (s.q[j] < t->bnd_lo[j]) df = t->bnd_lo[j] - s.q[j];
        else if (s.q[j] > t->bnd_hi[j]) df = s.q[j] - t->bnd_hi[j];
        box_dist = box_dist + df * df;
      }
      kd_search_rec(&s, t->root, box_dist);
      for (int j = 0; j < k; ++j) {
        idx_out[qi * k + j] = s.val[j] + 1;
        dist_out[qi * k + j] = sqrt(s.key[j]);
      }
    }
    free(s.key); free(s.val);
  }
#ifdef _OPENMP
  double t2 = omp_get_wtime();
  if (build_seconds) *build_seconds = t1 - t0;
  if (search_seconds) *search_seconds = t2 - t1;
#else
  if (build_seconds) *build_seconds = 0;
  if (search_seconds) *search_seconds = 0;
#endif
  kd_free(t);
  return 0;
}

/* ------------------------------------------------------------------ */
/* nn Graph (R/clustering.R:664-670) and ComputeSNN (src/snn.cpp:14-38) */
/* ------------------------------------------------------------------ */

/* Build the column-compressed adjacency A (n_rows x n_cols) of the kNN lists:
 * A[i, nn(i,j)-1] += 1, duplicates summed (sparseMatrix / setFromTriplets).
 * nn is row-major [n_rows x k], 1-based.  Outputs: colptr[n_cols+1] (int64),
 * then *rows_out / *vals_out (malloc'd, length colptr[n_cols]) with row
 * indices ascending inside each column. */
static int build_csc(const int32_t *nn, int64_t n_rows, int k, int64_t n_cols,
                     int64_t *colptr, int32_t **rows_out, double **vals_out) {
  int64_t *cnt = (int64_t *)calloc((size_t)n_cols + 1, sizeof(int64_t));
  for (int64_t t = 0; t < n_rows * k; ++t) {
    int64_t c = (int64_t)nn[t] - 1;
    if (c < 0 || c >= n_cols) { free(cnt); return 2; }
    cnt[c + 1]++;
  }
  for (int64_t c = 0; c < n_cols; ++c) cnt[c + 1] += cnt[c];
  int64_t total = cnt[n_cols];
  int32_t *rows = (int32_t *)malloc(sizeof(int32_t) * (size_t)(total ? total : 1));
  int64_t *cur = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n_cols + 1));
  memcpy(cur, cnt, sizeof(int64_t) * (size_t)(n_cols + 1));
  /* rows visited in ascending order => each column comes out sorted */
  for (int64_t i = 0; i < n_rows; ++i)
    for (int j = 0; j < k; ++j) {
      int64_t c = (int64_t)nn[i * k + j] - 1;
      rows[cur[c]++] = (int32_t)i;
    }
  free(cur);
  /* merge duplicates (same row listed twice in a column) */
  double *vals = (double *)malloc(sizeof(double) * (size_t)(total ? total : 1));
  int64_t w = 0;
  colptr[0] = 0;
  for (int64_t c = 0; c < n_cols; ++c) {
    int64_t b = cnt[c], e = cnt[c + 1];
    for (int64_t p = b; p < e; ++p) {
      if (p > b && rows[p] == rows[p - 1]) vals[w - 1] += 1.0;
      else { rows[w] = rows[p]; vals[w] = 1.0; w++; }
    }
    colptr[c + 1] = w;
  }
  free(cnt);
  *rows_out = rows; *vals_out = vals;
  return 0;
}

/* nn Graph as a dgCMatrix (n x n): p[n+1], i[nnz], x[nnz].  Two-call
 * protocol: call with i_out == NULL to get nnz, then again to fill. */
ORACLE_API int oracle_nn_graph(const int32_t *nn, int64_t n_rows, int k, int64_t n,
                               int64_t *nnz_out, int32_t *p_out, int32_t *i_out,
                               double *x_out) {
  int64_t *colptr = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 1));
  int32_t *rows; double *vals;
  int rc = build_csc(nn, n_rows, k, n, colptr, &rows, &vals);
  if (rc) { free(colptr); return rc; }
  *nnz_out = colptr[n];
  if (i_out) {
    for (int64_t c = 0; c <= n; ++c) p_out[c] = (int32_t)colptr[c];
    memcpy(i_out, rows, sizeof(int32_t) * (size_t)colptr[n]);
    memcpy(x_out, vals, sizeof(double) * (size_t)colptr[n]);
  }
  free(colptr); free(rows); free(vals);
  return 0;
}

static int oracle_cmp_i32(const void *a, const void *b) {
  int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
  return (x > y) - (x < y);
}

typedef struct {
  int64_t n, nnz;
  int64_t *p;
  int32_t *i;
  double *x;
} snn_result;

/*
 * ComputeSNN.  nn_ranked row-major [n x k], 1-based.  Row-streaming Gustavson
 * product of A with its transpose, never materialising the pre-prune matrix
 * (the reference does, see SURVEY.md section 6 for its 2^31 limit); per stored
 * entry v <- v / (k + (k - v)) in double, strict `v < prune` -> 0, exact zeros
 * dropped (snn.cpp:30-36).  Result rows == columns (S is symmetric), indices
 * ascending.  Returns an opaque handle; fetch with oracle_snn_fetch.
 */
/* rows [row_begin, row_end) only (the transpose is always built in full);
 * seconds[0] = transpose, seconds[1] = rows, for sub-sampled baselines. */
ORACLE_API int oracle_compute_snn_rows(const int32_t *nn, int64_t n, int k, double prune,
                                       int64_t row_begin, int64_t row_end, int nthreads,
                                       void **handle, int64_t *nnz_out, double *seconds) {
#ifdef _OPENMP
  double t0 = omp_get_wtime();
#endif
  int64_t *colptr = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 1));
  int32_t *crow; double *cval;
  int rc = build_csc(nn, n, k, n, colptr, &crow, &cval);
  if (rc) { free(colptr); return rc; }
#ifdef _OPENMP
  double t1 = omp_get_wtime();
  if (nthreads > 0) omp_set_num_threads(nthreads);
  int nt = omp_get_max_threads();
#else
  int nt = 1;
#endif
  const int64_t nrows = row_end - row_begin;
  snn_result *res = (snn_result *)calloc(1, sizeof(snn_result));
  res->n = nrows;
  res->p = (int64_t *)calloc((size_t)nrows + 1, sizeof(int64_t));
  /* per-thread output chunks over contiguous row blocks */
  int64_t nblk = nt * 8; if (nblk > nrows) nblk = nrows > 0 ? nrows : 1;
  int32_t **bi = (int32_t **)calloc((size_t)nblk, sizeof(int32_t *));
  double **bx = (double **)calloc((size_t)nblk, sizeof(double *));
  int64_t *bn = (int64_t *)calloc((size_t)nblk, sizeof(int64_t));
  const double kd = (double)k;
#pragma omp parallel
  {
    double *acc = (double *)calloc((size_t)(n ? n : 1), sizeof(double));
    int32_t *touched = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n ? n : 1));
#pragma omp for schedule(dynamic, 1)
    for (int64_t b = 0; b < nblk; ++b) {
      int64_t r0 = row_begin + nrows * b / nblk, r1 = row_begin + nrows * (b + 1) / nblk;
      int64_t cap = (r1 - r0) * 64 + 64, cnt = 0;
      int32_t *oi = (int32_t *)malloc(sizeof(int32_t) * (size_t)cap);
      double *ox = (double *)malloc(sizeof(double) * (size_t)cap);
      for (int64_t i = r0; i < r1; ++i) {
        int64_t nt_ = 0;
        /* S[i, :] = sum_c A[i,c] * A[:,c]  (A[i,c] counted once per listing) */
        for (int j = 0; j < k; ++j) {
          int64_t c = (int64_t)nn[i * k + j] - 1;
          for (int64_t p = colptr[c]; p < colptr[c + 1]; ++p) {
            int32_t r = crow[p];
            if (acc[r] == 0.0) touched[nt_++] = r;
            acc[r] += cval[p];
          }
        }
        /* ascending column order */
        if (nt_ > 1) {
          /* insertion sort is fine for short lists, qsort otherwise */
          if (nt_ < 32) {
            for (int64_t a = 1; a < nt_; ++a) {
              int32_t v = touched[a]; int64_t bb = a;
              while (bb > 0 && touched[bb - 1] > v) { touched[bb] = touched[bb - 1]; bb--; }
              touched[bb] = v;
            }
          } else {
            qsort(touched, (size_t)nt_, sizeof(int32_t), oracle_cmp_i32);
          }
        }
        int64_t row_n = 0;
        for (int64_t a = 0; a < nt_; ++a) {
          int32_t r = touched[a];
          double v = acc[r];
          acc[r] = 0.0;
          v = v / (kd + (kd - v));
          if (v < prune) v = 0;
          if (v != 0) { /* SNN.prune(0.0) keeps everything that is not exactly zero */
            if (cnt == cap) {
              cap *= 2;
              oi = (int32_t *)realloc(oi, sizeof(int32_t) * (size_t)cap);
              ox = (double *)realloc(ox, sizeof(double) * (size_t)cap);
            }
            oi[cnt] = r; ox[cnt] = v; cnt++; row_n++;
          }
        }
        res->p[i - row_begin + 1] = row_n;
      }
      bi[b] = oi; bx[b] = ox; bn[b] = cnt;
    }
    free(acc); free(touched);
  }
  for (int64_t i = 0; i < nrows; ++i) res->p[i + 1] += res->p[i];
  res->nnz = res->p[nrows];
  res->i = (int32_t *)malloc(sizeof(int32_t) * (size_t)(res->nnz ? res->nnz : 1));
  res->x = (double *)malloc(sizeof(double) * (size_t)(res->nnz ? res->nnz : 1));
  int64_t off = 0;
  for (int64_t b = 0; b < nblk; ++b) {
    if (bn[b]) {
      memcpy(res->i + off, bi[b], sizeof(int32_t) * (size_t)bn[b]);
      memcpy(res->x + off, bx[b], sizeof(double) * (size_t)bn[b]);
    }
    off += bn[b];
    free(bi[b]); free(bx[b]);
  }
  free(bi); free(bx); free(bn);
  free(colptr); free(crow); free(cval);
  *handle = res;
  *nnz_out = res->nnz;
#ifdef _OPENMP
  if (seconds) { seconds[0] = t1 - t0; seconds[1] = omp_get_wtime() - t1; }
#else
  if (seconds) { seconds[0] = 0; seconds[1] = 0; }
#endif
  return 0;
}

ORACLE_API int oracle_compute_snn(const int32_t *nn, int64_t n, int k, double prune,
                                  int nthreads, void **handle, int64_t *nnz_out) {
  return oracle_compute_snn_rows(nn, n, k, prune, 0, n, nthreads, handle, nnz_out, NULL);
}

/* copy out (p as int64 so that callers can detect dgCMatrix overflow) and free */
ORACLE_API void oracle_snn_fetch(void *handle, int64_t *p_out, int32_t *i_out, double *x_out) {
  snn_result *res = (snn_result *)handle;
  if (p_out) memcpy(p_out, res->p, sizeof(int64_t) * (size_t)(res->n + 1));
  if (i_out) memcpy(i_out, res->i, sizeof(int32_t) * (size_t)res->nnz);
  if (x_out) memcpy(x_out, res->x, sizeof(double) * (size_t)res->nnz);
  free(res->p); free(res->i); free(res->x); free(res);
}

/* ------------------------------------------------------------------ */
/* L2Norm (R/utilities.R:2107-2122): rows scaled to unit length,        */
/* non-finite results set to 0.  In place, row-major.                   */
/* ------------------------------------------------------------------ */
ORACLE_API void oracle_l2norm(double *m, int64_t n, int d) {
  for (int64_t i = 0; i < n; ++i) {
    long double acc = 0.0L; /* R's sum() accumulates doubles in long double */
    for (int j = 0; j < d; ++j) acc += (long double)(m[i * d + j] * m[i * d + j]);
    double s = sqrt((double)acc);
    for (int j = 0; j < d; ++j) {
      double v = m[i * d + j] / s;
      m[i * d + j] = isfinite(v) ? v : 0.0;
    }
  }
}

/* ------------------------------------------------------------------ */
/* SNN side consumers                                     */
/* ------------------------------------------------------------------ */

/* WriteEdgeFile, /root/reference/src/snn.cpp:40-59: strict lower triangle, "col\trow\tvalue" with
 * std::setprecision(15) on the default float format, i.e. "%.15g". */
#include <stdio.h>
ORACLE_API int64_t oracle_write_edge_file(const int32_t *p, const int32_t *i, const double *x, int64_t n,
                                          const char *filename) {
  FILE *f = fopen(filename, "w");
  if (!f) return -1;
  int64_t lines = 0;
  for (int64_t k = 0; k < n; ++k)
    for (int64_t e = p[k]; e < p[k + 1]; ++e) {
      if (k >= i[e]) continue; /* it.col() >= it.row() */
      fprintf(f, "%lld\t%d\t%.15g\n", (long long)k, i[e], x[e]);
      ++lines;
    }
  fclose(f);
  return lines;
}

/* fast_dist, /root/reference/src/fast_NN_dist.cpp:8-69: rval += pow(d1 - d2, 2) over the columns,
 * sqrt(rval).  x [nx x d], y [ny x d] row-major; lists flattened as ptr / nbr (1-based). */
ORACLE_API int oracle_fast_dist(const double *x, int64_t nx, const double *y, int64_t ny, int d, const int64_t *ptr,
                                const int32_t *nbr, double *out) {
  for (int64_t r = 0; r < nx; ++r)
    for (int64_t q = ptr[r]; q < ptr[r + 1]; ++q) {
      const int64_t j = (int64_t)nbr[q] - 1;
      if (j < 0 || j >= ny) return -1;
      double rval = 0;
      for (int c = 0; c < d; ++c) rval += pow(x[r * d + c] - y[j * d + c], 2);
      out[q] = sqrt(rval);
    }
  return 0;
}

/* SNN_SmallestNonzero_Dist, /root/reference/src/snn.cpp:82-126.  The row distance is Eigen's
 * (a - b).norm() there; restated as the square root of the squares summed in column order. */
static int cmp_desc(const void *a, const void *b) {
  const double u = *(const double *)a, v = *(const double *)b;
  return u < v ? 1 : (u > v ? -1 : 0);
}
ORACLE_API void oracle_snn_smallest_nonzero_dist(const int32_t *p, const int32_t *ri, const double *x, int64_t n,
                                                 const double *mat, int d, int n_small, const double *nearest,
                                                 double *out) {
  for (int64_t i = 0; i < n; ++i) {
    const int64_t b = p[i], cnt = p[i + 1] - b;
    int64_t *order = (int64_t *)malloc(sizeof(int64_t) * (size_t)(cnt > 0 ? cnt : 1));
    double *dists = (double *)malloc(sizeof(double) * (size_t)(cnt > 0 ? cnt : 1));
    /* sort_indexes (:70-79): std::stable_sort of 0..cnt-1 by value */
    for (int64_t a = 0; a < cnt; ++a) order[a] = a;
    for (int64_t a = 1; a < cnt; ++a) { /* stable insertion sort */
      const int64_t o = order[a];
      int64_t q = a;
      while (q > 0 && x[b + o] < x[b + order[q - 1]]) {
        order[q] = order[q - 1];
        --q;
      }
      order[q] = o;
    }
    int64_t n_i = n_small;
    if (n_i > cnt) n_i = cnt;
    int64_t nd = 0;
    for (int64_t j = 0; j < cnt; ++j) {
      const int64_t cell = ri[b + order[j]];
      if (nd < n_i || x[b + order[j]] == x[b + order[n_i - 1]]) {
        double acc = 0;
        for (int c = 0; c < d; ++c) {
          const double t = mat[cell * d + c] - mat[i * d + c];
          acc += t * t;
        }
        double res = sqrt(acc);
        if (nearest[i] > 0) {
          res = res - nearest[i];
          if (res < 0) res = 0;
        }
        dists[nd++] = res;
      } else {
        break;
      }
    }
    double avg = 0;
    if (nd > n_i) {
      qsort(dists, (size_t)nd, sizeof(double), cmp_desc);
      for (int64_t a = 0; a < n_i; ++a) avg += dists[a];
      avg /= (double)n_i;
    } else {
      for (int64_t a = 0; a < nd; ++a) avg += dists[a];
      avg /= (double)nd;
    }
    out[i] = avg;
    free(order);
    free(dists);
  }
}

ORACLE_API int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
