// internal.cuh -- shared declarations of libb200nn (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/b200nn.h"

// ---------------------------------------------------------------------------
// error plumbing: every internal function returns a B200NN_* code
// ---------------------------------------------------------------------------
void b200nn_set_error(const char *fmt, ...);

#define CU_TRY(expr)                                                              \
  do {                                                                            \
    cudaError_t _e = (expr);                                                      \
    if (_e != cudaSuccess) {                                                      \
      b200nn_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e),        \
                       __FILE__, __LINE__, cudaGetErrorString(_e));               \
      (void)cudaGetLastError();                                                   \
      return _e == cudaErrorMemoryAllocation ? B200NN_ERR_NOMEM : B200NN_ERR_CUDA; \
    }                                                                             \
  } while (0)

#define RC_TRY(expr)           \
  do {                         \
    int _rc = (expr);          \
    if (_rc != B200NN_OK) return _rc; \
  } while (0)

#define KERNEL_CHECK(ctx)            \
  do {                               \
    (ctx)->launches++;               \
    CU_TRY(cudaGetLastError());      \
  } while (0)

// ---------------------------------------------------------------------------
// context: one stream + a slot-addressed workspace that only ever grows
// ---------------------------------------------------------------------------
enum WsSlot {
  WS_STAGE_A = 0,  // raw host copies (embedding / nn_ranked) before re-layout
  WS_STAGE_B,
  WS_REF64,        // reference embedding, row-major fp64
  WS_QRY64,        // query embedding, row-major fp64
  WS_IDX0,         // kNN index, row-major 0-based int32
  WS_DIST,         // kNN distances, row-major fp64
  WS_OUT_IDX,      // idx in caller layout (1-based)
  WS_OUT_DIST,     // dist in caller layout
  WS_INDEG,        // in-degree per column (int32, n+1)
  WS_COLPTR,       // exclusive scan of in-degree (int64, n+1)
  WS_CURSOR,       // scatter cursors (int32, n)
  WS_REV,          // reverse lists = rows of each column (int32, n_rows*k)
  WS_UNIQ,         // unique entries per column (int32, n+1)
  WS_NN_P,         // nn Graph CSC: p (int64, n+1)
  WS_NN_I,         //               i (int32)
  WS_NN_X,         //               x (fp64)
  WS_ROW_BOUND,    // survivor bound per row (int32)
  WS_TMP_OFF,      // exclusive scan of bounds (int64)
  WS_TMP,          // packed survivors (uint32 or uint64 words)
  WS_ROW_NNZ,      // survivors per row (int32)
  WS_SNN_P,        // SNN CSR: p (int64, rows+1)
  WS_SNN_I,
  WS_SNN_X,
  WS_CLASS_LIST,   // row ids grouped by work class (int32, 4 * rows)
  WS_SCAN_TMP,     // block sums of the scan
  WS_INFO,         // small device struct with counters / flags
  WS_LUT_KEEP,     // keep[s] bytes
  WS_LUT_JAC,      // jaccard[s] doubles
  WS_HUGE_TABLE,   // global-memory hash tables of the "huge row" path
  WS_OUT_P32,      // int32 copies of p for dgCMatrix
  WS_LONG_LIST,    // columns handled by the block-level sort
  WS_OVF_LIST,     // rows that overflowed the filtered SNN kernel's exact table
  // tensor-core kNN path
  WS_TC_REF16,     // fp16 operand matrix of the reference set (K-padded, augmented)
  WS_TC_QRY16,     // fp16 operand matrix of the queries
  WS_TC_QRYERR,
  WS_TC_CAND,      // candidate indices (int32, nq * k')
  WS_TC_CANDKEY,   // selection threshold tau per query (fp32, nq)
  WS_TC_FLAGS,     // per-query certificate flags / fallback list
  WS_TC_FAILTAU,   // upper bound of the k-th squared distance of every fallback row
  WS_TC_MISC,
  WS_TC_Y16R,      // FP16 rows of the reference set in original order (grouped search)
  WS_TC_Y16Q,
  WS_TC_SPLIT,     // FP16 triple split of |yh|^2 per reference row
  WS_TC_IBUF,      // group ids, permutations, offsets, list counts
  WS_TC_FBUF,      // centroids, radii, item balls
  WS_TC_TILES,     // per-item tile lists
  WS_FLAG,         // one int32 flag (finite check)
  WS_LOW_OFF,      // symmetric SNN: offsets of the mirrored (lower-triangle) segments
  WS_LOW_WORDS,    //                packed words of the lower triangle, unsorted / sorted per row
  WS_LOW_SORTED,
  WS_LONG_LIST2,
  WS_TC_CENT16,    // fp16 operand rows of the group centroids (tensor-core assignment)
  WS_REV_SORTED,   // reverse lists sorted per column (nn Graph side stream)
  WS_SCAN_TMP2,    // block sums of scans queued on the side stream
  WS_DIST_FULL,    // multi-device search: full-size distance scratch of a group-aligned share
  WS_NUM
};

struct WsBuf {
  void *p = nullptr;
  size_t cap = 0;
};

// host mirror of the device-side info block (one D2H per phase)
struct DevInfo {
  int32_t err_index_range;   // some neighbour index was outside [0, n)
  int32_t has_dups;          // some row lists the same neighbour twice
  int32_t max_indegree;
  int32_t max_row_e;         // largest expanded-candidate count of a row
  int64_t sum_row_e;
  int64_t sum_bound;
  int64_t nnz_nn;
  int64_t nnz_snn;
  int32_t class_count[8];    // rows per work class (5 used)
  int32_t n_uncertified;
  int32_t n_long;            // columns too long for the warp-level sort
  int32_t n_ovf;             // rows the filtered SNN kernel handed to the big-table kernel
  int32_t pad_;
  long long dbg_hot_keys;    // experimental single-walk kernel: distinct hot keys verified (all rows)
  int32_t n_long2;           // long lower-triangle segments of the symmetric SNN assembly
  int32_t pad2_;
};

struct CsrResult {
  bool valid = false;
  int64_t n_rows = 0;
  int64_t nnz = 0;
  const int64_t *p = nullptr;
  const int32_t *i = nullptr;
  const double *x = nullptr;
};

struct b200nn_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // device->host copies that overlap later stages
  cudaStream_t side_stream = nullptr;  // nn Graph columns (sort / emit) while the SNN stage runs
  cudaEvent_t ev[16] = {};
  cudaEvent_t ev_copy = nullptr;
  cudaEvent_t ev_side = nullptr;   // nn Graph side stream: start marker
  cudaEvent_t ev_sorted = nullptr; // ... its sorted columns are complete
  WsBuf ws[WS_NUM];
  std::vector<uint8_t> h_keep;     // SNN lookup tables of the last call (host copies must outlive the H2D)
  std::vector<double> h_jac;
  DevInfo *h_info = nullptr;  // pinned (device-visible: small read-backs are stored here by a kernel)
  long long *h_scalars = nullptr;  // pinned, 8 values
  int sm_count = 148;
  int64_t launches = 0;
  CsrResult nn_graph, snn;
  // pending host-buffer results (between *_count and *_fill)
  bool pending_nn = false, pending_snn = false;
  // rows written by the last sharded kNN call (opts.shard_world > 1): either the `owned_slots`
  // entries of owned_rows (device; -1 = padding) or, when owned_rows is null, rows [owned_b, owned_e)
  const int32_t *owned_rows = nullptr;
  int64_t owned_slots = 0, owned_b = 0, owned_e = 0;
};

// grow-only workspace allocation; returns nullptr + error code via rc
int ws_reserve(b200nn_ctx *ctx, WsSlot slot, size_t bytes, void **out);

template <typename T>
static inline int ws_get(b200nn_ctx *ctx, WsSlot slot, size_t count, T **out) {
  void *p = nullptr;
  int rc = ws_reserve(ctx, slot, count * sizeof(T), &p);
  *out = static_cast<T *>(p);
  return rc;
}

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------
// stage entry points implemented in the other translation units
// ---------------------------------------------------------------------------

// util.cu
int scan_exclusive_i32_to_i64(b200nn_ctx *ctx, const int32_t *d_in, int64_t *d_out, int64_t n,
                              cudaStream_t stream = nullptr, int tmp_slot = -1);
// out[r * ncol + c] = in[c * nrow + r] (+ add)   (column-major -> row-major)
int transpose_f64(b200nn_ctx *ctx, const double *d_in, double *d_out, int64_t nrow, int64_t ncol);
int transpose_i32_add(b200nn_ctx *ctx, const int32_t *d_in, int32_t *d_out, int64_t nrow,
                      int64_t ncol, int32_t add);
int add_i32(b200nn_ctx *ctx, const int32_t *d_in, int32_t *d_out, int64_t n, int32_t add);
int narrow_i64_to_i32(b200nn_ctx *ctx, const int64_t *d_in, int32_t *d_out, int64_t n, cudaStream_t stream = nullptr);
// B200NN_ERR_INVALID ("NA/NaN/Inf in foreign function call") when x holds a non-finite value
int check_finite_f64(b200nn_ctx *ctx, const double *d_x, int64_t count, const char *what);

// knn_exact.cu : FP64 exhaustive scan.  d_qsel == nullptr: queries 0..nq-1;
// otherwise the kernel processes the n_sel query rows listed in d_qsel
// (count read from the device at launch-time upper bound n_sel).
int knn_exact_launch(b200nn_ctx *ctx, const double *d_ref, int64_t nr, const double *d_qry,
                     int64_t nq, int d, int k, const int32_t *d_qsel, int64_t n_sel,
                     int32_t *d_idx0, double *d_dist,
                     const double *d_tau0 = nullptr);

// Reference rows clustered by the tensor-core search (all device pointers): groups of
// scaled, FP16-rounded points yh with centre and radius, stored group by group in a padded
// order (`rperm`: position -> original row, -1 for padding; group g owns positions
// [goff[g], goff[g+1]), a multiple of 128).
struct GroupedRefs {
  int G;                 // groups
  int ld;                // leading dimension of cent / yq16 (elements)
  const float *cent;     // [G x ld] group centres
  const float *grad;     // [G] radius (upper bound) of the group's points around its centre
  const int32_t *goff;   // [G + 1]
  const int32_t *rperm;
  const uint16_t *yq16;  // [nq x ld] FP16 bits of the scaled query rows, original order
  const double *q_err;   // [nq] |y_q - yh_q|
  const double *scale;   // s: y = s * (x - mu)
  const double *e_max;   // max over reference rows of |y_r - yh_r|
};
// knn_exact.cu : same result as knn_exact_launch for the listed rows, but only the groups
// that can hold a point within the row's bound d_tau0 (squared, reference arithmetic) are scanned
int knn_exact_grouped_launch(b200nn_ctx *ctx, const double *d_ref, const double *d_qry, int d, int k,
                             const int32_t *d_qsel, int64_t n_sel, const double *d_tau0,
                             const GroupedRefs &R, int32_t *d_idx0, double *d_dist);

// knn_tc.cu : tcgen05 candidate generation + FP64 re-rank + certificate
// scan_mode: 0 = auto, 1 = exhaustive scan of every key tile, 2 = grouped (tile-pruned) search
bool knn_tc_applicable(int64_t nr, int64_t nq, int d, int k, int scan_mode);
struct TcRefCache;  // reference-side state of the grouped search, kept by a resident index
TcRefCache *tc_ref_cache_create();
void tc_ref_cache_destroy(TcRefCache *c);
int knn_tc_launch(b200nn_ctx *ctx, const double *d_ref, int64_t nr, const double *d_qry,
                  int64_t nq, int d, int k, int n_cand, int scan_mode, int32_t *d_idx0, double *d_dist,
                  b200nn_stats *stats, int shard_rank = 0, int shard_world = 0, TcRefCache *cache = nullptr);
// true when knn_tc_launch would run the grouped (tile-pruned) search for this shape
bool knn_tc_grouped(int64_t nr, int scan_mode);

int knn_tc_debug(b200nn_ctx *ctx, const double *d_ref, int64_t nr, const double *d_qry, int64_t nq,
                 int d, int kp, int scan_mode, float *d_keys_out, int32_t *d_cand_out, float *d_tau_out,
                 double *h_scale_mu);

// api.cu : algorithm selection + argument checks of the kNN stage (device pointers, row-major)
b200nn_opts b200nn_default_opts();
int knn_dispatch(b200nn_ctx *ctx, const double *d_ref, int64_t nr, const double *d_qry, int64_t nq,
                 int d, int k, const b200nn_opts &o, int32_t *d_idx0, double *d_dist, b200nn_stats *stats,
                 TcRefCache *cache = nullptr);

// snn.cu
// optional host destination of the nn Graph (capacity n_rows * k entries): copied on the
// context's copy stream as soon as the graph exists, i.e. while the SNN stage runs
struct NnGraphHostOut {
  int32_t *p = nullptr, *i = nullptr;
  double *x = nullptr;
};
int graphs_launch(b200nn_ctx *ctx, const int32_t *d_idx0, int64_t n_rows, int k, int64_t n,
                  double prune, bool want_nn_graph, bool compute_snn, int64_t row_begin,
                  int64_t row_end, b200nn_stats *stats, const NnGraphHostOut *early_nn = nullptr);
