/*
 * b200nn_r.c -- the `.Call` shim between R and libb200nn.so.
 *
 * NOT compiled in the build container (no R headers there); it is the binding a
 * Seurat maintainer adds under src/ next to RcppExports.cpp.  It only marshals:
 * every computation happens behind the C ABI of include/b200nn.h, which is what
 * the test-suite exercises (through ctypes).
 *
 * Entry points and what they replace in the reference:
 *   b200_nn2(data, query, k)        RANN::nn2(data, query, k, searchtype = "standard",
 *                                   eps = 0) as reached from NNHelper,
 *                                   R/clustering.R:1664-1667  -> list(nn.idx, nn.dists)
 *   b200_ComputeSNN(nn_ranked, prune)  .Call('_Seurat_ComputeSNN', ...),
 *                                   R/RcppExports.R:96-98, src/RcppExports.cpp:313-323
 *                                   -> dgCMatrix (no dimnames; the caller sets them)
 *   b200_FindNeighbors(data, k, prune, compute_snn)  the body of FindNeighbors.default,
 *                                   R/clustering.R:633-682, with the index kept on the GPU
 *                                   -> list(nn = dgCMatrix, snn = dgCMatrix)
 *   With more than one GPU visible b200_nn2 and b200_FindNeighbors go through b200nn_multi_*:
 *   the call is sharded over the devices inside the library (B200NN_NUM_DEVICES limits it,
 *   B200NN_DEVICE pins it to one device).
 *   b200_index(data) / b200_index_nn2(index, query, k)  the cache.index = TRUE / index =
 *                                   arguments of NNHelper / FindNeighbors, R/clustering.R:596-597,
 *                                   :1686-1688: the reference embedding stays on the GPU behind an
 *                                   external pointer (finaliser pattern of src/valid_pointer.c:4-6)
 *   The step after FindNeighbors and the SNN side consumers, same arguments as the
 *   Rcpp wrappers they stand in for:
 *   b200_RunModularityClusteringCpp(SNN, modularityFunction, resolution, algorithm, nRandomStarts,
 *                                   nIterations, randomSeed, printOutput, edgefilename)
 *                                   _Seurat_RunModularityClusteringCpp, src/RcppExports.cpp:14-32
 *   b200_WriteEdgeFile(snn, filename, display_progress)        src/RcppExports.cpp:324-333
 *   b200_DirectSNNToFile(nn_ranked, prune, display_progress, filename)      :335-346
 *   b200_fast_dist(x, y, n)                                    :253-264
 *   b200_SNN_SmallestNonzero_Dist(snn, mat, n, nearest_dist)   :348-360
 * Registration follows src/RcppExports.cpp:407-445 (R_CallMethodDef table,
 * R_registerRoutines, R_useDynamicSymbols(dll, FALSE)).
 *
 * Rules honoured: R owns every input and output vector (allocated here with
 * Rf_allocVector and PROTECTed); the library never keeps an R pointer after a call;
 * the C ABI never long-jumps, so the return code is checked after all transient state
 * is released and only then turned into Rf_error().
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <stdint.h>

#include "b200nn.h"

static b200nn_ctx *g_ctx = NULL;      /* first device: stages that start from an existing index */
static b200nn_multi *g_multi = NULL;  /* every visible device: kNN and the fused FindNeighbors    */
static int g_live_indexes = 0;        /* resident reference sets that still point into g_ctx      */

static b200nn_ctx *get_ctx(void) {
  if (g_ctx == NULL) {
    const char *dev = getenv("B200NN_DEVICE");
    if (b200nn_ctx_create(dev ? atoi(dev) : 0, &g_ctx) != B200NN_OK) Rf_error("%s", b200nn_last_error());
  }
  return g_ctx;
}

/* All visible GPUs (or the first B200NN_NUM_DEVICES of them) behind one handle; NULL when only
 * one device is to be used, in which case the single-device entry points are called. */
static b200nn_multi *get_multi(void) {
  if (g_multi == NULL) {
    const char *nd = getenv("B200NN_NUM_DEVICES");
    int want = nd ? atoi(nd) : b200nn_device_count();
    if (getenv("B200NN_DEVICE") != NULL) want = 1; /* an explicit device pins the call to it */
    if (want <= 1) return NULL;
    if (b200nn_multi_create(NULL, want, &g_multi) != B200NN_OK) Rf_error("%s", b200nn_last_error());
  }
  return g_multi;
}

/* nn2 accepts integer matrices as well (R coerces for .C): hand back a REALSXP, PROTECTed by the caller */
static SEXP as_real_matrix(SEXP m, const char *what) {
  if (!Rf_isMatrix(m) || !(Rf_isReal(m) || Rf_isInteger(m) || Rf_isLogical(m)))
    Rf_error("%s must be a numeric matrix", what);
  return Rf_isReal(m) ? m : Rf_coerceVector(m, REALSXP);
}

static SEXP make_dgCMatrix(int n, int64_t nnz, SEXP p, SEXP i, SEXP x) {
  SEXP cls = PROTECT(R_do_MAKE_CLASS("dgCMatrix"));
  SEXP m = PROTECT(R_do_new_object(cls));
  SEXP dim = PROTECT(Rf_allocVector(INTSXP, 2));
  INTEGER(dim)[0] = n;
  INTEGER(dim)[1] = n;
  R_do_slot_assign(m, Rf_install("Dim"), dim);
  R_do_slot_assign(m, Rf_install("p"), p);
  R_do_slot_assign(m, Rf_install("i"), i);
  R_do_slot_assign(m, Rf_install("x"), x);
  UNPROTECT(3);
  (void)nnz;
  return m;
}

/* RANN::nn2 drop-in: data, query are numeric matrices (column-major); query may be R_NilValue */
SEXP b200_nn2(SEXP data_, SEXP query_, SEXP k_) {
  SEXP data = PROTECT(as_real_matrix(data_, "data"));
  const int self = Rf_isNull(query_);
  SEXP query = PROTECT(self ? R_NilValue : as_real_matrix(query_, "query"));
  const int nr = Rf_nrows(data), d = Rf_ncols(data), k = Rf_asInteger(k_);
  const int nq = self ? nr : Rf_nrows(query);
  if (!self && Rf_ncols(query) != d) Rf_error("Query and data tables must have same dimensions");
  if (k > nr) Rf_error("Cannot find more nearest neighbours than there are points");
  SEXP idx = PROTECT(Rf_allocMatrix(INTSXP, nq, k));
  SEXP dist = PROTECT(Rf_allocMatrix(REALSXP, nq, k));
  b200nn_opts o = {B200NN_COL_MAJOR, B200NN_KNN_AUTO, 0, 0};
  b200nn_multi *mu = get_multi();
  /* non-finite input is refused inside the library (nn2's .C(NAOK = FALSE) behaviour) */
  int rc = mu ? b200nn_multi_knn(mu, REAL(data), nr, self ? NULL : REAL(query), nq, d, k, INTEGER(idx),
                                 REAL(dist), &o, NULL)
              : b200nn_knn(get_ctx(), REAL(data), nr, self ? NULL : REAL(query), nq, d, k, INTEGER(idx),
                           REAL(dist), &o, NULL);
  if (rc != B200NN_OK) {
    UNPROTECT(4);
    Rf_error("%s", b200nn_last_error());
  }
  SEXP res = PROTECT(Rf_allocVector(VECSXP, 2));
  SEXP nms = PROTECT(Rf_allocVector(STRSXP, 2));
  SET_VECTOR_ELT(res, 0, idx);
  SET_VECTOR_ELT(res, 1, dist);
  SET_STRING_ELT(nms, 0, Rf_mkChar("nn.idx"));
  SET_STRING_ELT(nms, 1, Rf_mkChar("nn.dists"));
  Rf_setAttrib(res, R_NamesSymbol, nms);
  UNPROTECT(6);
  return res;
}

/* ComputeSNN drop-in: nn_ranked integer or double matrix (1-based), prune double */
SEXP b200_ComputeSNN(SEXP nn_ranked, SEXP prune_) {
  SEXP nn = PROTECT(Rf_coerceVector(nn_ranked, INTSXP)); /* Rcpp coerces to double; values are whole */
  SEXP dims = Rf_getAttrib(nn_ranked, R_DimSymbol);
  if (Rf_isNull(dims)) {
    UNPROTECT(1);
    Rf_error("nn_ranked must be a matrix");
  }
  const int n = INTEGER(dims)[0], k = INTEGER(dims)[1];
  int64_t nnz = 0;
  b200nn_opts o = {B200NN_COL_MAJOR, 0, 0, 0};
  int rc = b200nn_snn_count(get_ctx(), INTEGER(nn), n, k, Rf_asReal(prune_), &o, &nnz, NULL);
  if (rc != B200NN_OK) {
    UNPROTECT(1);
    Rf_error("%s", b200nn_last_error());
  }
  SEXP p = PROTECT(Rf_allocVector(INTSXP, n + 1));
  SEXP i = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)nnz));
  SEXP x = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)nnz));
  rc = b200nn_snn_fill(get_ctx(), INTEGER(p), INTEGER(i), REAL(x));
  if (rc != B200NN_OK) {
    UNPROTECT(4);
    Rf_error("%s", b200nn_last_error());
  }
  SEXP m = make_dgCMatrix(n, nnz, p, i, x);
  UNPROTECT(4);
  return m;
}

/* fused FindNeighbors.default body for query == object; all visible GPUs take a share */
SEXP b200_FindNeighbors(SEXP data_, SEXP k_, SEXP prune_, SEXP compute_snn_) {
  SEXP data = PROTECT(as_real_matrix(data_, "object"));
  const int n = Rf_nrows(data), d = Rf_ncols(data), k = Rf_asInteger(k_);
  const int snn = Rf_asLogical(compute_snn_);
  int64_t nnz_nn = 0, nnz_snn = 0;
  b200nn_opts o = {B200NN_COL_MAJOR, B200NN_KNN_AUTO, 0, 0};
  b200nn_multi *mu = get_multi();
  int rc = mu ? b200nn_multi_find_neighbors_count(mu, REAL(data), n, d, k, Rf_asReal(prune_), snn, NULL, NULL,
                                                  NULL, NULL, NULL, &o, &nnz_nn, &nnz_snn, NULL)
              : b200nn_find_neighbors_count(get_ctx(), REAL(data), n, d, k, Rf_asReal(prune_), snn, NULL, NULL,
                                            NULL, NULL, NULL, &o, &nnz_nn, &nnz_snn, NULL);
  if (rc != B200NN_OK) {
    UNPROTECT(1);
    Rf_error("%s", b200nn_last_error());
  }
  SEXP np = PROTECT(Rf_allocVector(INTSXP, n + 1));
  SEXP ni = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)nnz_nn));
  SEXP nx = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)nnz_nn));
  SEXP sp = PROTECT(Rf_allocVector(INTSXP, snn ? n + 1 : 0));
  SEXP si = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)nnz_snn));
  SEXP sx = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)nnz_snn));
  rc = mu ? b200nn_multi_find_neighbors_fill(mu, INTEGER(np), INTEGER(ni), REAL(nx), snn ? INTEGER(sp) : NULL,
                                             snn ? INTEGER(si) : NULL, snn ? REAL(sx) : NULL)
          : b200nn_find_neighbors_fill(get_ctx(), INTEGER(np), INTEGER(ni), REAL(nx), snn ? INTEGER(sp) : NULL,
                                       snn ? INTEGER(si) : NULL, snn ? REAL(sx) : NULL);
  if (rc != B200NN_OK) {
    UNPROTECT(7);
    Rf_error("%s", b200nn_last_error());
  }
  /* nothing of the library is pending any more: a user interrupt may unwind from here */
  R_CheckUserInterrupt();
  SEXP res = PROTECT(Rf_allocVector(VECSXP, snn ? 2 : 1));
  SEXP nms = PROTECT(Rf_allocVector(STRSXP, snn ? 2 : 1));
  SET_VECTOR_ELT(res, 0, make_dgCMatrix(n, nnz_nn, np, ni, nx));
  SET_STRING_ELT(nms, 0, Rf_mkChar("nn"));
  if (snn) {
    SET_VECTOR_ELT(res, 1, make_dgCMatrix(n, nnz_snn, sp, si, sx));
    SET_STRING_ELT(nms, 1, Rf_mkChar("snn"));
  }
  Rf_setAttrib(res, R_NamesSymbol, nms);
  UNPROTECT(9);
  return res;
}

/* resident reference set: external pointer with a finaliser */
static void index_finalizer(SEXP ptr) {
  b200nn_index *ix = (b200nn_index *)R_ExternalPtrAddr(ptr);
  if (ix != NULL) {
    b200nn_index_destroy(g_ctx, ix); /* g_ctx outlives every index: b200_release refuses otherwise */
    R_ClearExternalPtr(ptr);
    if (g_live_indexes > 0) --g_live_indexes;
  }
}

SEXP b200_index(SEXP data_) {
  SEXP data = PROTECT(as_real_matrix(data_, "data"));
  b200nn_opts o = {B200NN_COL_MAJOR, B200NN_KNN_AUTO, 0, 0};
  b200nn_index *ix = NULL;
  int rc = b200nn_index_create(get_ctx(), REAL(data), Rf_nrows(data), Rf_ncols(data), &o, &ix);
  UNPROTECT(1);
  if (rc != B200NN_OK) Rf_error("%s", b200nn_last_error());
  SEXP ptr = PROTECT(R_MakeExternalPtr(ix, Rf_install("b200nn_index"), R_NilValue));
  R_RegisterCFinalizerEx(ptr, index_finalizer, TRUE);
  ++g_live_indexes;
  UNPROTECT(1);
  return ptr;
}

/* nn2 against a resident reference set; query may be R_NilValue (self) */
SEXP b200_index_nn2(SEXP index, SEXP query_, SEXP k_) {
  b200nn_index *ix = (b200nn_index *)R_ExternalPtrAddr(index);
  if (ix == NULL) Rf_error("the cached index is no longer valid");
  const int nr = (int)b200nn_index_rows(ix), d = b200nn_index_dims(ix), k = Rf_asInteger(k_);
  const int self = Rf_isNull(query_);
  SEXP query = PROTECT(self ? R_NilValue : as_real_matrix(query_, "query"));
  const int nq = self ? nr : Rf_nrows(query);
  if (!self && Rf_ncols(query) != d) Rf_error("Query and data tables must have same dimensions");
  SEXP idx = PROTECT(Rf_allocMatrix(INTSXP, nq, k));
  SEXP dist = PROTECT(Rf_allocMatrix(REALSXP, nq, k));
  b200nn_opts o = {B200NN_COL_MAJOR, B200NN_KNN_AUTO, 0, 0};
  int rc = b200nn_index_knn(get_ctx(), ix, self ? NULL : REAL(query), nq, k, INTEGER(idx), REAL(dist), &o, NULL);
  if (rc != B200NN_OK) {
    UNPROTECT(3);
    Rf_error("%s", b200nn_last_error());
  }
  SEXP res = PROTECT(Rf_allocVector(VECSXP, 2));
  SEXP nms = PROTECT(Rf_allocVector(STRSXP, 2));
  SET_VECTOR_ELT(res, 0, idx);
  SET_VECTOR_ELT(res, 1, dist);
  SET_STRING_ELT(nms, 0, Rf_mkChar("nn.idx"));
  SET_STRING_ELT(nms, 1, Rf_mkChar("nn.dists"));
  Rf_setAttrib(res, R_NamesSymbol, nms);
  UNPROTECT(5);
  return res;
}

/* ---- the step after FindNeighbors and the SNN side consumers ------------------------------------- */
static void dgc_slots(SEXP m, int *n, const int **p, const int **i, const double **x) {
  SEXP dim = R_do_slot(m, Rf_install("Dim"));
  *n = INTEGER(dim)[1];
  *p = INTEGER(R_do_slot(m, Rf_install("p")));
  *i = INTEGER(R_do_slot(m, Rf_install("i")));
  *x = REAL(R_do_slot(m, Rf_install("x")));
}

/* RunModularityClusteringCpp drop-in (src/RModularityOptimizer.cpp:21-173): SNN dgCMatrix (only its strict
 * lower triangle is read) or, with a non-empty edgefilename, the edge file; returns the integer vector of
 * 0-based cluster ids, clusters ordered by decreasing size.  The wrapper's console lines are kept. */
SEXP b200_RunModularityClusteringCpp(SEXP SNN, SEXP modularityFunction, SEXP resolution, SEXP algorithm,
                                     SEXP nRandomStarts, SEXP nIterations, SEXP randomSeed, SEXP printOutput,
                                     SEXP edgefilename) {
  const int mf = Rf_asInteger(modularityFunction), alg = Rf_asInteger(algorithm);
  const int ns = Rf_asInteger(nRandomStarts), ni = Rf_asInteger(nIterations), print = Rf_asLogical(printOutput);
  const double res = Rf_asReal(resolution);
  const uint64_t seed = (uint64_t)(int64_t)Rf_asInteger(randomSeed); /* JavaRandom(int) sign-extends */
  const char *file = Rf_isNull(edgefilename) ? "" : CHAR(STRING_ELT(edgefilename, 0));
  b200nn_modopt_stats st;
  SEXP out;
  int rc;
  if (print) Rprintf("Modularity Optimizer version 1.3.0 by Ludo Waltman and Nees Jan van Eck\n\n");
  if (file[0] != '\0') {
    int64_t n = 0;
    if (print) Rprintf("Reading input file...\n\n");
    rc = b200nn_modularity_cluster_edge_file(get_ctx(), file, mf, res, alg, ns, ni, seed, NULL, 0, &n, NULL);
    if (rc != B200NN_OK) Rf_error("%s", b200nn_last_error());
    out = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)n));
    rc = b200nn_modularity_cluster_edge_file(get_ctx(), file, mf, res, alg, ns, ni, seed, INTEGER(out), n, &n, &st);
  } else {
    int n;
    const int *p, *i;
    const double *x;
    dgc_slots(SNN, &n, &p, &i, &x);
    out = PROTECT(Rf_allocVector(INTSXP, n));
    rc = b200nn_modularity_cluster(get_ctx(), n, p, i, x, mf, res, alg, ns, ni, seed, INTEGER(out), &st);
  }
  if (rc != B200NN_OK) {
    UNPROTECT(1);
    Rf_error("%s", b200nn_last_error()); /* the wrapper's stop() texts */
  }
  if (print) {
    Rprintf("Number of nodes: %d\nNumber of edges: %d\n\n", (int)st.n_nodes, (int)st.n_edges);
    Rprintf("Running %s...\n", alg == 1 ? "Louvain algorithm"
                               : alg == 2 ? "Louvain algorithm with multilevel refinement"
                                          : "smart local moving algorithm");
    if (ns == 1) Rprintf("Modularity: %.4f\n", st.modularity);
    else Rprintf("Maximum modularity in %d random starts: %.4f\n", ns, st.modularity);
    Rprintf("Number of communities: %d\nElapsed time: %d seconds\n", (int)st.n_clusters, (int)(st.ms_total / 1000.0));
  }
  UNPROTECT(1);
  return out;
}

/* WriteEdgeFile drop-in (src/snn.cpp:40-59) */
SEXP b200_WriteEdgeFile(SEXP snn, SEXP filename, SEXP display_progress) {
  int n;
  const int *p, *i;
  const double *x;
  dgc_slots(snn, &n, &p, &i, &x);
  if (Rf_asLogical(display_progress) == TRUE) REprintf("Writing SNN as edge file\n");
  if (b200nn_write_edge_file(p, i, x, n, CHAR(STRING_ELT(filename, 0)), NULL) != B200NN_OK)
    Rf_error("%s", b200nn_last_error());
  return R_NilValue;
}

SEXP b200_ComputeSNN(SEXP nn_ranked, SEXP prune_);
/* DirectSNNToFile drop-in (src/snn.cpp:63-69): ComputeSNN, then the edge file, the graph is returned */
SEXP b200_DirectSNNToFile(SEXP nn_ranked, SEXP prune, SEXP display_progress, SEXP filename) {
  SEXP g = PROTECT(b200_ComputeSNN(nn_ranked, prune));
  b200_WriteEdgeFile(g, filename, display_progress);
  UNPROTECT(1);
  return g;
}

/* fast_dist drop-in (src/fast_NN_dist.cpp:34-69): n is a list of (1-based) row indices of y, one vector per
 * row of x; returns the list of distance vectors (an empty list when the shapes do not fit, as the reference) */
SEXP b200_fast_dist(SEXP x_, SEXP y_, SEXP n_) {
  SEXP x = PROTECT(as_real_matrix(x_, "x"));
  SEXP y = PROTECT(as_real_matrix(y_, "y"));
  const R_xlen_t nx = Rf_nrows(x), m = Rf_xlength(n_);
  if (nx != m || Rf_ncols(x) != Rf_ncols(y)) {
    UNPROTECT(2);
    return Rf_allocVector(VECSXP, 0);
  }
  int64_t *ptr = (int64_t *)R_alloc((size_t)m + 1, sizeof(int64_t));
  ptr[0] = 0;
  for (R_xlen_t r = 0; r < m; ++r) ptr[r + 1] = ptr[r] + Rf_xlength(VECTOR_ELT(n_, r));
  int32_t *nbr = (int32_t *)R_alloc((size_t)(ptr[m] > 0 ? ptr[m] : 1), sizeof(int32_t));
  for (R_xlen_t r = 0; r < m; ++r) {
    SEXP v = VECTOR_ELT(n_, r);
    const R_xlen_t len = Rf_xlength(v);
    for (R_xlen_t j = 0; j < len; ++j)
      nbr[ptr[r] + j] = Rf_isInteger(v) ? INTEGER(v)[j] : (int32_t)REAL(v)[j];
  }
  double *flat = (double *)R_alloc((size_t)(ptr[m] > 0 ? ptr[m] : 1), sizeof(double));
  int rc = b200nn_fast_dist(get_ctx(), REAL(x), nx, REAL(y), Rf_nrows(y), Rf_ncols(x), B200NN_COL_MAJOR, ptr, nbr,
                            flat);
  if (rc != B200NN_OK) {
    UNPROTECT(2);
    Rf_error("%s", b200nn_last_error());
  }
  SEXP res = PROTECT(Rf_duplicate(n_)); /* clone(n): names are kept */
  for (R_xlen_t r = 0; r < m; ++r) {
    const R_xlen_t len = (R_xlen_t)(ptr[r + 1] - ptr[r]);
    SEXP d = PROTECT(Rf_allocVector(REALSXP, len));
    for (R_xlen_t j = 0; j < len; ++j) REAL(d)[j] = flat[ptr[r] + j];
    SET_VECTOR_ELT(res, r, d);
    UNPROTECT(1);
  }
  UNPROTECT(3);
  return res;
}

/* SNN_SmallestNonzero_Dist drop-in (src/snn.cpp:82-126) */
SEXP b200_SNN_SmallestNonzero_Dist(SEXP snn, SEXP mat_, SEXP n_, SEXP nearest_dist) {
  int n;
  const int *p, *i;
  const double *x;
  dgc_slots(snn, &n, &p, &i, &x);
  SEXP mat = PROTECT(as_real_matrix(mat_, "mat"));
  SEXP nd = PROTECT(Rf_coerceVector(nearest_dist, REALSXP));
  SEXP out = PROTECT(Rf_allocVector(REALSXP, n));
  int rc = b200nn_snn_smallest_nonzero_dist(get_ctx(), p, i, x, n, REAL(mat), Rf_ncols(mat), B200NN_COL_MAJOR,
                                            Rf_asInteger(n_), REAL(nd), REAL(out));
  UNPROTECT(3);
  if (rc != B200NN_OK) Rf_error("%s", b200nn_last_error());
  return out;
}

/* Frees the device state.  Resident indexes live inside the single-device context: while any
 * external pointer to one is still alive the context stays (their finaliser needs it). */
SEXP b200_release(void) {
  if (g_multi) {
    b200nn_multi_destroy(g_multi);
    g_multi = NULL;
  }
  if (g_ctx && g_live_indexes == 0) {
    b200nn_ctx_destroy(g_ctx);
    g_ctx = NULL;
  }
  return Rf_ScalarLogical(g_ctx == NULL);
}

static const R_CallMethodDef CallEntries[] = {
    {"b200_nn2", (DL_FUNC)&b200_nn2, 3},
    {"b200_ComputeSNN", (DL_FUNC)&b200_ComputeSNN, 2},
    {"b200_FindNeighbors", (DL_FUNC)&b200_FindNeighbors, 4},
    {"b200_index", (DL_FUNC)&b200_index, 1},
    {"b200_index_nn2", (DL_FUNC)&b200_index_nn2, 3},
    {"b200_RunModularityClusteringCpp", (DL_FUNC)&b200_RunModularityClusteringCpp, 9},
    {"b200_WriteEdgeFile", (DL_FUNC)&b200_WriteEdgeFile, 3},
    {"b200_DirectSNNToFile", (DL_FUNC)&b200_DirectSNNToFile, 4},
    {"b200_fast_dist", (DL_FUNC)&b200_fast_dist, 3},
    {"b200_SNN_SmallestNonzero_Dist", (DL_FUNC)&b200_SNN_SmallestNonzero_Dist, 4},
    {"b200_release", (DL_FUNC)&b200_release, 0},
    {NULL, NULL, 0}};

void R_init_b200nn(DllInfo *dll) {
  R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
