/*
 * b200nn.h -- C ABI of libb200nn.so, the B200 (sm_100a) implementation of
 * Seurat's FindNeighbors hot path: exact Euclidean kNN + nn Graph + SNN graph.
 *
 * Plain pointers and sizes only: this is what the R package's `.Call` shim
 * (shim/b200nn_r.c), Python's ctypes (seurat_b200/_lib.py) or any other FFI
 * binds.  Every entry point returns 0 on success or a B200NN_ERR_* code;
 * b200nn_last_error() gives the message.  Nothing here throws or long-jumps.
 * There is NO CPU fallback: without a CUDA device every compute call fails
 * with B200NN_ERR_CUDA.
 *
 * Reference interfaces replaced (paths relative to the Seurat tree):
 *   b200nn_knn              RANN::nn2(data, query, k, searchtype="standard",
 *                           eps=0) as called by NNHelper, R/clustering.R:1664-1667
 *   b200nn_snn_*            ComputeSNN(nn_ranked, prune): src/snn.cpp:14-38,
 *                           glue R/RcppExports.R:96-98, src/RcppExports.cpp:313-323
 *   b200nn_nn_graph_*       the sparseMatrix() nn Graph, R/clustering.R:664-670
 *   b200nn_find_neighbors_* the body of FindNeighbors.default,
 *                           R/clustering.R:633-682, fused so that the kNN index
 *                           never leaves the device between the stages
 *   b200nn_dev_*            the same stages on buffers already resident in
 *                           HBM (used by the multi-GPU host code, which shards
 *                           query rows over one process per GPU and all-gathers
 *                           the index between the kNN and the SNN stage)
 */
#ifndef B200NN_H
#define B200NN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200NN_VERSION 2

#if defined(__GNUC__)
#define B200NN_API __attribute__((visibility("default")))
#else
#define B200NN_API
#endif

enum {
  B200NN_OK = 0,
  B200NN_ERR_INVALID = 1,      /* bad argument (NULL, d<=0, k<=0, ...)               */
  B200NN_ERR_K_TOO_LARGE = 2,  /* k > number of reference points (nn2's stop())     */
  B200NN_ERR_CUDA = 3,         /* CUDA runtime / driver failure, or no device       */
  B200NN_ERR_INDEX_RANGE = 4,  /* neighbour index outside [1, n]                    */
  B200NN_ERR_NNZ_OVERFLOW = 5, /* result does not fit a dgCMatrix (nnz >= 2^31)     */
  B200NN_ERR_NOMEM = 6,        /* device allocation failed                          */
  B200NN_ERR_STATE = 7         /* *_fill called without a matching *_count          */
};

/* memory layout of embedding inputs and of idx/dist/nn_ranked matrices */
enum {
  B200NN_COL_MAJOR = 0, /* R matrices: element (row, col) at [col * nrow + row]    */
  B200NN_ROW_MAJOR = 1  /* C / numpy:  element (row, col) at [row * ncol + col]    */
};

/* kNN algorithm selection */
enum {
  B200NN_KNN_AUTO = 0,   /* tensor-core path when applicable, else exact scan       */
  B200NN_KNN_EXACT = 1,  /* FP64 exhaustive scan (reference arithmetic everywhere)  */
  B200NN_KNN_TENSOR = 2  /* tcgen05 FP16 candidate generation + FP64 re-rank with a
                            per-row exactness certificate and FP64 fallback rows     */
};

typedef struct b200nn_ctx b200nn_ctx;

typedef struct b200nn_opts {
  int32_t layout;      /* B200NN_COL_MAJOR (default, R) or B200NN_ROW_MAJOR          */
  int32_t knn_algo;    /* B200NN_KNN_*                                               */
  int32_t n_cand;      /* widened candidate count k' of the tensor path, 0 = auto    */
  int32_t scan_mode;   /* tensor path: 0 = auto, 1 = exhaustive (every reference tile),
                          2 = grouped (tile-pruned); same result, for benchmarking   */
  /* b200nn_dev_knn only: this call computes share `shard_rank` of `shard_world` of a SELF
     search (d_qry = NULL).  d_idx0 / d_dist are then FULL [nr x k] buffers of which only the
     rows of this share are written.  Which rows a share owns is the library's choice (whole
     groups of the tile-pruned search, so that every share keeps full work items); every share
     of a call with the same data makes the same choice, and the shares partition the rows.
     Callers zero-fill d_idx0 and combine the shares with a sum (all-reduce).  0 / 0 = off. */
  int32_t shard_rank;
  int32_t shard_world;
} b200nn_opts;

typedef struct b200nn_stats {
  double ms_h2d;          /* host->device copies                                     */
  double ms_prep;         /* layout / precision conversion of the embedding          */
  double ms_knn_cand;     /* tensor-core candidate generation (K1)                   */
  double ms_knn_rerank;   /* FP64 re-rank + certificate (K2)                         */
  double ms_knn_fallback; /* FP64 exhaustive scan (all rows, or uncertified rows)    */
  double ms_nn_graph;     /* transpose / nn Graph (K3)                               */
  double ms_snn;          /* SNN rows (K4)                                           */
  double ms_d2h;          /* device->host copies                                     */
  int64_t n_uncertified;  /* query rows that needed the FP64 fallback scan           */
  int64_t nnz_nn;         /* entries of the nn Graph                                 */
  int64_t nnz_snn;        /* entries of the SNN graph                                */
  int64_t snn_expanded;   /* sum over rows of expanded candidates (sum_c indeg^2)    */
  int64_t max_indegree;   /* longest reverse-neighbour list                          */
  int64_t gpu_launches;   /* kernels launched by this call                           */
  int64_t knn_tiles_visited; /* (work item, reference tile) pairs the tensor kernel scanned  */
  int64_t knn_tiles_total;   /* ... out of this many (equal for the exhaustive scan)         */
  double ms_knn_pass2;       /* the dominant launch of K1 alone (pass 2 of the grouped search, or the
                                exhaustive scan); part of ms_knn_cand                         */
} b200nn_stats;

/* ---- context ---------------------------------------------------------- */

/* Creates a context on CUDA device `device` (owns one stream and a growable
 * workspace; buffers are reused across calls). */
B200NN_API int b200nn_ctx_create(int device, b200nn_ctx **ctx_out);
B200NN_API void b200nn_ctx_destroy(b200nn_ctx *ctx);
/* Blocks until all work queued on the context's stream has finished. */
B200NN_API int b200nn_ctx_sync(b200nn_ctx *ctx);
/* The context's cudaStream_t (as void*), e.g. to record timing events on. */
B200NN_API void *b200nn_ctx_stream(b200nn_ctx *ctx);
/* Message of the last error raised on the calling thread. */
B200NN_API const char *b200nn_last_error(void);
B200NN_API int b200nn_version(void);
/* Number of CUDA devices visible (0 when there is none / no driver). */
B200NN_API int b200nn_device_count(void);

/* ---- host-buffer entry points (what the R shim binds) ------------------ */

/*
 * Exact k nearest reference rows of every query row (Euclidean).
 *   ref  [nr x d], qry [nq x d] doubles in opts->layout; qry == NULL means
 *   qry = ref (self search: column 1 is the point itself unless duplicates).
 *   idx_out  [nq x k] int32, 1-based reference row numbers, opts->layout
 *   dist_out [nq x k] double, ascending Euclidean distances,  opts->layout
 * Ties in distance are ordered by lower index.
 */
B200NN_API int b200nn_knn(b200nn_ctx *ctx, const double *ref, int64_t nr, const double *qry,
               int64_t nq, int32_t d, int32_t k, int32_t *idx_out, double *dist_out,
               const b200nn_opts *opts, b200nn_stats *stats);

/*
 * ComputeSNN in two phases so that the caller (R) can allocate the result.
 *   nn_ranked [n x k] int32, 1-based, opts->layout.  count: computes the graph
 *   on the device and reports nnz.  fill: copies p[n+1], i[nnz] (0-based,
 *   ascending within each column), x[nnz] -- the slots of a dgCMatrix.
 */
B200NN_API int b200nn_snn_count(b200nn_ctx *ctx, const int32_t *nn_ranked, int64_t n, int32_t k,
                     double prune, const b200nn_opts *opts, int64_t *nnz_out,
                     b200nn_stats *stats);
B200NN_API int b200nn_snn_fill(b200nn_ctx *ctx, int32_t *p, int32_t *i, double *x);

/*
 * nn Graph: sparseMatrix(i = row, j = nn_ranked, x = 1, dims = c(n, n)) with
 * duplicates summed.  nn_ranked is [n_rows x k]; n_rows may differ from n
 * (the reference's query != object quirk); indices must lie in [1, n].
 */
B200NN_API int b200nn_nn_graph_count(b200nn_ctx *ctx, const int32_t *nn_ranked, int64_t n_rows,
                          int32_t k, int64_t n, const b200nn_opts *opts,
                          int64_t *nnz_out, b200nn_stats *stats);
B200NN_API int b200nn_nn_graph_fill(b200nn_ctx *ctx, int32_t *p, int32_t *i, double *x);

/*
 * FindNeighbors.default for query == object: kNN, nn Graph and (optionally)
 * SNN with the index kept on the device between stages.  idx_out / dist_out
 * may be NULL when the caller only wants the graphs.  nn_p_out [n + 1],
 * nn_i_out [n * k], nn_x_out [n * k] are optional: the nn Graph never has more
 * than n * k entries, so a caller may hand its buffers in already here; they are
 * then filled while the SNN stage runs (page-locked buffers make that a true
 * overlap) and _fill can be called with NULL for them.  The first *nnz_nn_out
 * entries are valid.
 */
B200NN_API int b200nn_find_neighbors_count(b200nn_ctx *ctx, const double *data, int64_t n, int32_t d,
                                int32_t k, double prune, int32_t compute_snn,
                                int32_t *idx_out, double *dist_out,
                                int32_t *nn_p_out, int32_t *nn_i_out, double *nn_x_out,
                                const b200nn_opts *opts, int64_t *nnz_nn_out,
                                int64_t *nnz_snn_out, b200nn_stats *stats);
B200NN_API int b200nn_find_neighbors_fill(b200nn_ctx *ctx, int32_t *nn_p, int32_t *nn_i, double *nn_x,
                               int32_t *snn_p, int32_t *snn_i, double *snn_x);

/* ---- resident reference set ("index") ----------------------------------- */
/*
 * The reference's other kNN callers (FindNN / FindWeights / MappingScore /
 * ProjectUMAP, R/integration.R:4411-4531, :2469-2664, R/dimensional_reduction.R:366-377)
 * query one reference embedding many times through NNHelper(..., cache.index =
 * TRUE) / FindNeighbors(index = ), R/clustering.R:596-597, :1658-1690.  For the
 * exact search the "index" is the reference embedding kept on the device: it is
 * uploaded once and every later query skips that transfer.  The object belongs
 * to the context it was created on and must be destroyed before it.
 */
typedef struct b200nn_index b200nn_index;
B200NN_API int b200nn_index_create(b200nn_ctx *ctx, const double *ref, int64_t nr, int32_t d,
                                   const b200nn_opts *opts, b200nn_index **index_out);
B200NN_API void b200nn_index_destroy(b200nn_ctx *ctx, b200nn_index *index);
B200NN_API int64_t b200nn_index_rows(const b200nn_index *index);
B200NN_API int32_t b200nn_index_dims(const b200nn_index *index);
/* same contract as b200nn_knn with the reference taken from the index; qry == NULL: self */
B200NN_API int b200nn_index_knn(b200nn_ctx *ctx, const b200nn_index *index, const double *qry, int64_t nq,
                                int32_t k, int32_t *idx_out, double *dist_out, const b200nn_opts *opts,
                                b200nn_stats *stats);

/* ---- several devices from one process ------------------------------------ */
/*
 * FindNeighbors.default is called from a single-threaded R process
 * (R/clustering.R:633-682), so the sharding over 1/2/4/8 GPUs lives inside the
 * library: one worker thread, context and stream per device.  Query cells are
 * split contiguously; every device uploads its slice of the embedding over its
 * own PCIe link, the slices and later the int32 kNN index are exchanged between
 * the devices over NVLink (peer copies), every device emits the SNN rows of its
 * share, and all results land in the caller's host buffers as ONE Neighbor /
 * dgCMatrix -- same contract as the single-device entry points above.
 * devices == NULL: devices 0 .. n_devices-1; n_devices <= 0: every visible device.
 * A device may be listed more than once (the shards then share it).
 */
typedef struct b200nn_multi b200nn_multi;
B200NN_API int b200nn_multi_create(const int32_t *devices, int32_t n_devices, b200nn_multi **multi_out);
B200NN_API void b200nn_multi_destroy(b200nn_multi *multi);
B200NN_API int32_t b200nn_multi_size(const b200nn_multi *multi);
/* same contract as b200nn_knn */
B200NN_API int b200nn_multi_knn(b200nn_multi *multi, const double *ref, int64_t nr, const double *qry, int64_t nq,
                                int32_t d, int32_t k, int32_t *idx_out, double *dist_out,
                                const b200nn_opts *opts, b200nn_stats *stats);
/* same contract as b200nn_find_neighbors_count / _fill (stats: maximum over the devices of the
 * stage times, sums of the counters; ms_h2d includes the NVLink exchange of the embedding) */
B200NN_API int b200nn_multi_find_neighbors_count(b200nn_multi *multi, const double *data, int64_t n, int32_t d,
                                                 int32_t k, double prune, int32_t compute_snn,
                                                 int32_t *idx_out, double *dist_out, int32_t *nn_p_out,
                                                 int32_t *nn_i_out, double *nn_x_out, const b200nn_opts *opts,
                                                 int64_t *nnz_nn_out, int64_t *nnz_snn_out, b200nn_stats *stats);
B200NN_API int b200nn_multi_find_neighbors_fill(b200nn_multi *multi, int32_t *nn_p, int32_t *nn_i, double *nn_x,
                                                int32_t *snn_p, int32_t *snn_i, double *snn_x);

/* ---- device-buffer entry points ---------------------------------------- */
/* All pointers are device pointers on the context's device.  Work is queued
 * on the context's stream; calls that return sizes synchronise internally. */

/*
 * kNN on resident data.  d_ref [nr x d], d_qry [nq x d] (NULL = self) doubles,
 * ROW-MAJOR.  Outputs, both ROW-MAJOR [nq x k]: d_idx0 int32 **0-based**
 * reference rows (the form the SNN stage consumes), d_dist double.
 */
B200NN_API int b200nn_dev_knn(b200nn_ctx *ctx, const double *d_ref, int64_t nr, const double *d_qry,
                   int64_t nq, int32_t d, int32_t k, int32_t *d_idx0, double *d_dist,
                   const b200nn_opts *opts, b200nn_stats *stats);

/*
 * nn Graph + SNN rows [row_begin, row_end) from the FULL 0-based index
 * d_idx0 [n x k] (row-major).  Results stay on the device inside the context:
 * fetch pointers with b200nn_dev_result.  With compute_snn == 0 only the nn
 * Graph is built.  The nn Graph is always the full n x n matrix.
 */
B200NN_API int b200nn_dev_graphs(b200nn_ctx *ctx, const int32_t *d_idx0, int64_t n, int32_t k,
                      double prune, int32_t compute_snn, int64_t row_begin,
                      int64_t row_end, int64_t *nnz_nn_out, int64_t *nnz_snn_out,
                      b200nn_stats *stats);

/*
 * Candidate stage of the tensor-core kNN path alone (diagnostics / profiling):
 * prep + tcgen05 candidate kernel, no re-rank.  d_ref/d_qry as in b200nn_dev_knn
 * (d <= 61).  Outputs (device, any may be NULL): d_cand [nq x n_cand] int32
 * 0-based candidate rows (unordered), d_tau [nq] float selection thresholds,
 * d_keys [256 x 128] float raw accumulator keys of the first 256 queries against
 * the first 128 reference rows.  h_scale_mu (host, 65 doubles, may be NULL)
 * receives the power-of-two scale and the centring vector of the prep stage.
 * scan_mode: 0 = auto, 1 = exhaustive scan, 2 = grouped (tile-pruned) search;
 * d_keys forces the exhaustive scan.
 */
B200NN_API int b200nn_dev_knn_candidates(b200nn_ctx *ctx, const double *d_ref, int64_t nr,
                                         const double *d_qry, int64_t nq, int32_t d,
                                         int32_t n_cand, int32_t scan_mode, int32_t *d_cand,
                                         float *d_tau, float *d_keys, double *h_scale_mu);

typedef struct b200nn_dev_csr {
  int64_t n_rows;      /* rows (== columns for the symmetric SNN block)             */
  int64_t nnz;
  const int64_t *p;    /* [n_rows + 1] offsets (64-bit on the device)               */
  const int32_t *i;    /* [nnz] 0-based, ascending within a row/column              */
  const double *x;     /* [nnz]                                                     */
} b200nn_dev_csr;

/* which: 0 = nn Graph (CSC of the n x n matrix), 1 = SNN rows of the last
 * b200nn_dev_graphs call (row-compressed block; column indices are global). */
B200NN_API int b200nn_dev_result(b200nn_ctx *ctx, int32_t which, b200nn_dev_csr *out);

/* Device-to-device copy of a result into caller-owned device buffers
 * (d_p: int64[n_rows + 1], d_i: int32[nnz], d_x: double[nnz]; NULL = skip).
 * Queued on the context's stream and synchronised before returning. */
B200NN_API int b200nn_dev_copy_result(b200nn_ctx *ctx, int32_t which, int64_t *d_p, int32_t *d_i,
                                      double *d_x);

/* ---- device-resident embedding hand-off ------------------- */
/*
 * What FindNeighbors does to the embedding on the host before the search, for an embedding that already
 * lives in HBM (a PCA computed on the device): the column subset Embeddings(object[[reduction]])[, dims]
 * (/root/reference/R/clustering.R:843-846) and, with l2_norm != 0, L2Norm (R/clustering.R:619-622,
 * R/utilities.R:2107-2122: rows divided by sqrt(sum(x^2)), R's long-double sum reproduced bit for bit,
 * non-finite results set to 0).  d_src: [n x n_cols] doubles on the context's device in src_layout;
 * dims: n_dims 0-based column numbers in HOST memory (NULL = all columns in order); d_out: [n x n_dims]
 * row-major on the device -- what b200nn_dev_knn / b200nn_dev_graphs take, so no H2D of the embedding
 * remains.  Must not alias d_src.
 */
B200NN_API int b200nn_dev_embedding_prepare(b200nn_ctx *ctx, const double *d_src, int64_t n, int32_t n_cols,
                                            int32_t src_layout, const int32_t *dims, int32_t n_dims,
                                            int32_t l2_norm, double *d_out);

/* ---- modularity optimiser (the step after FindNeighbors) --- */
/*
 * Replaces RunModularityClusteringCpp, /root/reference/src/RModularityOptimizer.cpp:21-173
 * (registered as _Seurat_RunModularityClusteringCpp, /root/reference/src/RcppExports.cpp; called from
 * RunModularityClustering, /root/reference/R/clustering.R:1881-1906, i.e. FindClusters with
 * algorithm = 1 / 2).  Same arguments, same result for the same random_seed: the cluster of every
 * node, 0-based, clusters ordered by decreasing size.  Only the strict lower triangle of the matrix
 * is read (RModularityOptimizer.cpp:72-81), weights must be positive.
 *   algorithm 1 = Louvain, 2 = Louvain with multilevel refinement, 3 = smart local moving; 4
 *   (Leiden: FindClusters routes it to an R-side package, never here) returns B200NN_ERR_INVALID.
 * Error texts are the wrapper's stop() messages.
 */
typedef struct b200nn_modopt_stats {
  int64_t n_nodes, n_edges;
  int32_t n_clusters, pad_;
  double modularity;           /* quality function of the returned clustering (best start)      */
  int64_t windows;             /* windows of the node permutation executed in parallel           */
  int64_t rounds;              /* fixed-point rounds over all windows                            */
  int64_t sequential_steps;    /* nodes handled by the single-node (reference) step              */
  int64_t window_retries;      /* windows repeated with a smaller size                           */
  int64_t local_moving_calls, levels, max_window;
  double ms_total, ms_build, ms_local_moving, ms_reduce, ms_quality;
  double ms_subnetworks;       /* SLM: local moving inside the clusters                          */
  int64_t gpu_launches;
} b200nn_modopt_stats;

/* p [n + 1], i [nnz], x [nnz]: dgCMatrix slots of the n x n graph in HOST memory (p[0] = 0).
 * cluster_out [n] int32 in host memory. */
B200NN_API int b200nn_modularity_cluster(b200nn_ctx *ctx, int64_t n, const int32_t *p, const int32_t *i,
                                         const double *x, int32_t modularity_function, double resolution,
                                         int32_t algorithm, int32_t n_random_starts, int32_t n_iterations,
                                         uint64_t random_seed, int32_t *cluster_out, b200nn_modopt_stats *stats);
/* The same on an edge list in host memory: edge e joins the 0-based nodes node1[e] and node2[e] with
 * weight[e] (NULL: 1).  As in the reference (matrixToNetwork, ModularityOptimizer.cpp:759-803) only
 * edges with node1 < node2 count, and every node's neighbours keep the order of the list. */
B200NN_API int b200nn_modularity_cluster_edges(b200nn_ctx *ctx, int64_t n_nodes, const int32_t *node1,
                                               const int32_t *node2, const double *weight, int64_t n_edges,
                                               int32_t modularity_function, double resolution, int32_t algorithm,
                                               int32_t n_random_starts, int32_t n_iterations, uint64_t random_seed,
                                               int32_t *cluster_out, b200nn_modopt_stats *stats);
/* The edge.file.name input of RunModularityClusteringCpp (RModularityOptimizer.cpp:55-63; file format of
 * readInputFile, ModularityOptimizer.cpp:806-838 = what b200nn_write_edge_file writes): lines
 * "node1<TAB>node2[<TAB>weight]", the number of nodes is the largest id + 1 (*n_nodes_out).  Call with
 * cluster_out = NULL to learn n_nodes, then with cluster_capacity >= n_nodes.  A file that cannot be opened
 * or parsed gives B200NN_ERR_INVALID with the wrapper's "Could not parse edge file.". */
B200NN_API int b200nn_modularity_cluster_edge_file(b200nn_ctx *ctx, const char *filename,
                                                   int32_t modularity_function, double resolution,
                                                   int32_t algorithm, int32_t n_random_starts,
                                                   int32_t n_iterations, uint64_t random_seed,
                                                   int32_t *cluster_out, int64_t cluster_capacity,
                                                   int64_t *n_nodes_out, b200nn_modopt_stats *stats);
/* The same on the SNN graph the context still holds in HBM after b200nn_find_neighbors_count /
 * b200nn_dev_graphs (all rows): no D2H + H2D of the graph between FindNeighbors and FindClusters.
 * n_expected > 0 is checked against the resident graph's size. */
B200NN_API int b200nn_modularity_cluster_resident(b200nn_ctx *ctx, int32_t modularity_function, double resolution,
                                                  int32_t algorithm, int32_t n_random_starts,
                                                  int32_t n_iterations, uint64_t random_seed,
                                                  int32_t *cluster_out, int64_t n_expected,
                                                  b200nn_modopt_stats *stats);

/* ---- SNN side consumers ----------------------------------- */
/*
 * WriteEdgeFile, /root/reference/src/snn.cpp:40-59 (and the file half of DirectSNNToFile, :63-69):
 * the strict lower triangle of the dgCMatrix (p, i, x) as lines "col<TAB>row<TAB>value", columns
 * in order, value with 15 significant digits ("%.15g" = std::setprecision(15)).  Host I/O only.
 * n_written (may be NULL) receives the number of lines.
 */
B200NN_API int b200nn_write_edge_file(const int32_t *p, const int32_t *i, const double *x, int64_t n,
                                      const char *filename, int64_t *n_written);
/*
 * fast_dist, /root/reference/src/fast_NN_dist.cpp:34-69: Euclidean distance between row r of x
 * [nx x d] and every row of y [ny x d] listed for it; the lists are given as ptr [nx + 1] /
 * nbr [ptr[nx]] (1-based rows of y, R's list of index vectors flattened); out [ptr[nx]].
 * Arithmetic as the reference's: squares summed in dimension order, sqrt at the end (bit-exact).
 */
B200NN_API int b200nn_fast_dist(b200nn_ctx *ctx, const double *x, int64_t nx, const double *y, int64_t ny,
                                int32_t d, int32_t layout, const int64_t *ptr, const int32_t *nbr, double *out);
/*
 * SNN_SmallestNonzero_Dist, /root/reference/src/snn.cpp:82-126: for every column of the SNN graph
 * the mean distance (minus nearest_dist, floored at 0) to the n_small cells with the smallest
 * non-zero weights (all cells tied with the n_small-th are included; if that makes more than
 * n_small, the n_small largest distances are averaged).  mat [n x d] in `layout`; out [n].
 * Distances agree with the reference's Eigen norm to rounding (1e-12 relative), see extras_core.inl.
 */
B200NN_API int b200nn_snn_smallest_nonzero_dist(b200nn_ctx *ctx, const int32_t *p, const int32_t *i,
                                                const double *x, int64_t n, const double *mat, int32_t d,
                                                int32_t layout, int32_t n_small, const double *nearest_dist,
                                                double *out);

#ifdef __cplusplus
}
#endif
#endif /* B200NN_H */
