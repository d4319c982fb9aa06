// api.cu -- the C ABI of libb200nn.so (see include/b200nn.h).
#include <string.h>

#include <algorithm>
#include <new>

#include "internal.cuh"

namespace {

struct Timer {
  b200nn_ctx *ctx;
  cudaEvent_t a, b;
  Timer(b200nn_ctx *c, int slot) : ctx(c), a(c->ev[slot]), b(c->ev[slot + 1]) {}
  int start() {
    CU_TRY(cudaEventRecord(a, ctx->stream));
    return B200NN_OK;
  }
  int stop(double *ms_out) {
    CU_TRY(cudaEventRecord(b, ctx->stream));
    CU_TRY(cudaEventSynchronize(b));
    float ms = 0;
    CU_TRY(cudaEventElapsedTime(&ms, a, b));
    if (ms_out) *ms_out += ms;
    return B200NN_OK;
  }
};

b200nn_opts default_opts() {
  b200nn_opts o;
  memset(&o, 0, sizeof(o));
  o.layout = B200NN_COL_MAJOR;
  o.knn_algo = B200NN_KNN_AUTO;
  return o;
}

int check_ctx(b200nn_ctx *ctx) {
  if (!ctx) {
    b200nn_set_error("NULL context");
    return B200NN_ERR_INVALID;
  }
  CU_TRY(cudaSetDevice(ctx->device));
  return B200NN_OK;
}

// host matrix [rows x cols] doubles in `layout` -> device row-major
int upload_matrix_f64(b200nn_ctx *ctx, const double *h, int64_t rows, int64_t cols, int layout,
                      WsSlot stage, WsSlot dst_slot, double **d_out) {
  size_t cnt = (size_t)rows * (size_t)cols;
  double *dst;
  RC_TRY(ws_get(ctx, dst_slot, cnt, &dst));
  if (layout == B200NN_ROW_MAJOR) {
    CU_TRY(cudaMemcpyAsync(dst, h, cnt * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  } else {
    double *raw;
    RC_TRY(ws_get(ctx, stage, cnt, &raw));
    CU_TRY(cudaMemcpyAsync(raw, h, cnt * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    RC_TRY(transpose_f64(ctx, raw, dst, rows, cols));
  }
  *d_out = dst;
  return B200NN_OK;
}

// host nn_ranked [rows x k] int32 1-based in `layout` -> device row-major 0-based
int upload_nn_ranked(b200nn_ctx *ctx, const int32_t *h, int64_t rows, int k, int layout,
                     int32_t **d_out) {
  size_t cnt = (size_t)rows * (size_t)k;
  int32_t *raw, *dst;
  RC_TRY(ws_get(ctx, WS_STAGE_A, cnt, &raw));
  RC_TRY(ws_get(ctx, WS_IDX0, cnt, &dst));
  CU_TRY(cudaMemcpyAsync(raw, h, cnt * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  if (layout == B200NN_ROW_MAJOR)
    RC_TRY(add_i32(ctx, raw, dst, (int64_t)cnt, -1));
  else
    RC_TRY(transpose_i32_add(ctx, raw, dst, rows, k, -1));
  *d_out = dst;
  return B200NN_OK;
}

// device row-major idx0 / dist [nq x k] -> host buffers in `layout`, idx 1-based
// Layout conversion on the compute stream, then the device->host copies on the copy stream:
// they run while later kernels of the compute stream execute.  download_knn_end waits for them.
int download_knn_begin(b200nn_ctx *ctx, const int32_t *d_idx0, const double *d_dist, int64_t nq, int k,
                       int layout, int32_t *h_idx, double *h_dist) {
  size_t cnt = (size_t)nq * (size_t)k;
  int32_t *oi = nullptr;
  const double *od = d_dist;
  if (h_idx) {
    RC_TRY(ws_get(ctx, WS_OUT_IDX, cnt, &oi));
    if (layout == B200NN_ROW_MAJOR)
      RC_TRY(add_i32(ctx, d_idx0, oi, (int64_t)cnt, 1));
    else  // row-major [nq x k] == column-major [k x nq]; transposing that gives col-major [nq x k]
      RC_TRY(transpose_i32_add(ctx, d_idx0, oi, k, nq, 1));
  }
  if (h_dist && layout != B200NN_ROW_MAJOR) {
    double *o;
    RC_TRY(ws_get(ctx, WS_OUT_DIST, cnt, &o));
    RC_TRY(transpose_f64(ctx, d_dist, o, k, nq));
    od = o;
  }
  CU_TRY(cudaEventRecord(ctx->ev_copy, ctx->stream));
  CU_TRY(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_copy, 0));
  if (h_idx) CU_TRY(cudaMemcpyAsync(h_idx, oi, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->copy_stream));
  if (h_dist) CU_TRY(cudaMemcpyAsync(h_dist, od, cnt * sizeof(double), cudaMemcpyDeviceToHost, ctx->copy_stream));
  return B200NN_OK;
}

int download_knn_end(b200nn_ctx *ctx) {
  CU_TRY(cudaStreamSynchronize(ctx->copy_stream));
  return B200NN_OK;
}

int download_knn(b200nn_ctx *ctx, const int32_t *d_idx0, const double *d_dist, int64_t nq, int k,
                 int layout, int32_t *h_idx, double *h_dist) {
  RC_TRY(download_knn_begin(ctx, d_idx0, d_dist, nq, k, layout, h_idx, h_dist));
  return download_knn_end(ctx);
}

int download_csr(b200nn_ctx *ctx, const CsrResult &r, int32_t *p, int32_t *i, double *x) {
  if (!r.valid) {
    b200nn_set_error("no result pending: call the matching *_count entry point first");
    return B200NN_ERR_STATE;
  }
  if (r.nnz >= 0x7fffffffLL) {
    b200nn_set_error("graph has %lld entries, more than a dgCMatrix can hold", (long long)r.nnz);
    return B200NN_ERR_NNZ_OVERFLOW;
  }
  if (p) {
    int32_t *p32;
    RC_TRY(ws_get(ctx, WS_OUT_P32, (size_t)r.n_rows + 1, &p32));
    RC_TRY(narrow_i64_to_i32(ctx, r.p, p32, r.n_rows + 1));
    CU_TRY(cudaMemcpyAsync(p, p32, sizeof(int32_t) * ((size_t)r.n_rows + 1), cudaMemcpyDeviceToHost,
                           ctx->stream));
  }
  if (i && r.nnz)
    CU_TRY(cudaMemcpyAsync(i, r.i, sizeof(int32_t) * (size_t)r.nnz, cudaMemcpyDeviceToHost,
                           ctx->stream));
  if (x && r.nnz)
    CU_TRY(cudaMemcpyAsync(x, r.x, sizeof(double) * (size_t)r.nnz, cudaMemcpyDeviceToHost,
                           ctx->stream));
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  return B200NN_OK;
}

}  // namespace

b200nn_opts b200nn_default_opts() { return default_opts(); }

int knn_dispatch(b200nn_ctx *ctx, const double *d_ref, int64_t nr, const double *d_qry, int64_t nq,
                 int d, int k, const b200nn_opts &o, int32_t *d_idx0, double *d_dist,
                 b200nn_stats *stats, TcRefCache *cache) {
  if (!d_ref || nr <= 0 || nq < 0 || d <= 0 || k <= 0 || !d_idx0 || !d_dist) {
    b200nn_set_error("knn: invalid argument (nr=%lld nq=%lld d=%d k=%d)", (long long)nr,
                     (long long)nq, d, k);
    return B200NN_ERR_INVALID;
  }
  if ((int64_t)k > nr) {
    b200nn_set_error("Cannot find more nearest neighbours than there are points");
    return B200NN_ERR_K_TOO_LARGE;
  }
  if (nr >= 0x7fffffffLL || nq >= 0x7fffffffLL) {
    b200nn_set_error("knn: more than 2^31-1 points are not supported");
    return B200NN_ERR_INVALID;
  }
  if (nq == 0) return B200NN_OK;
  const double *q = d_qry ? d_qry : d_ref;
  int algo = o.knn_algo;
  ctx->owned_rows = nullptr;
  ctx->owned_slots = 0;
  ctx->owned_b = 0;
  ctx->owned_e = nq;
  const bool sharded = o.shard_world > 1;
  if (sharded) {
    if (q != d_ref || nq != nr || o.shard_rank < 0 || o.shard_rank >= o.shard_world || o.shard_world > 64) {
      b200nn_set_error("knn: a sharded call must be a self search (d_qry = NULL) with 0 <= rank < world <= 64");
      return B200NN_ERR_INVALID;
    }
    const bool tc = algo != B200NN_KNN_EXACT && knn_tc_applicable(nr, nq, d, k, o.scan_mode & 3);
    if (!(tc && knn_tc_grouped(nr, o.scan_mode & 3))) {
      // no grouping to align the shares with: contiguous rows [b, e) of the full-size outputs
      const int64_t per = ceil_div64(nr, o.shard_world);
      const int64_t b = std::min<int64_t>((int64_t)o.shard_rank * per, nr), e = std::min<int64_t>(b + per, nr);
      ctx->owned_b = b;
      ctx->owned_e = e;
      if (e <= b) return B200NN_OK;
      b200nn_opts o1 = o;
      o1.shard_rank = o1.shard_world = 0;
      int rc = knn_dispatch(ctx, d_ref, nr, d_ref + b * d, e - b, d, k, o1, d_idx0 + b * k, d_dist + b * k, stats);
      ctx->owned_rows = nullptr;
      ctx->owned_b = b;
      ctx->owned_e = e;
      return rc;
    }
  }
  if (algo == B200NN_KNN_AUTO)
    algo = knn_tc_applicable(nr, nq, d, k, o.scan_mode & 3) ? B200NN_KNN_TENSOR : B200NN_KNN_EXACT;
  if (algo == B200NN_KNN_TENSOR) {
    if (!knn_tc_applicable(nr, nq, d, k, o.scan_mode & 3)) {
      b200nn_set_error("tensor-core kNN path does not cover nr=%lld d=%d k=%d", (long long)nr, d, k);
      return B200NN_ERR_INVALID;
    }
    return knn_tc_launch(ctx, d_ref, nr, q, nq, d, k, o.n_cand, o.scan_mode & 3, d_idx0, d_dist, stats,
                         sharded ? o.shard_rank : 0, sharded ? o.shard_world : 0, cache);
  }
  RC_TRY(check_finite_f64(ctx, d_ref, nr * (int64_t)d, "data"));
  if (q != d_ref) RC_TRY(check_finite_f64(ctx, q, nq * (int64_t)d, "query"));
  Timer t(ctx, 4);
  RC_TRY(t.start());
  RC_TRY(knn_exact_launch(ctx, d_ref, nr, q, nq, d, k, nullptr, 0, d_idx0, d_dist));
  RC_TRY(t.stop(stats ? &stats->ms_knn_fallback : nullptr));
  return B200NN_OK;
}


extern "C" {

int b200nn_version(void) { return B200NN_VERSION; }

int b200nn_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return n;
}

int b200nn_ctx_create(int device, b200nn_ctx **ctx_out) {
  if (!ctx_out) {
    b200nn_set_error("ctx_out is NULL");
    return B200NN_ERR_INVALID;
  }
  *ctx_out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    (void)cudaGetLastError();
    b200nn_set_error("no CUDA device available (%s); libb200nn has no CPU fallback",
                     e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    return B200NN_ERR_CUDA;
  }
  if (device < 0 || device >= n) {
    b200nn_set_error("device %d out of range (%d visible)", device, n);
    return B200NN_ERR_INVALID;
  }
  CU_TRY(cudaSetDevice(device));
  b200nn_ctx *ctx = new (std::nothrow) b200nn_ctx();
  if (!ctx) return B200NN_ERR_NOMEM;
  ctx->device = device;
  // from here on a failure releases what was created so far
  cudaError_t ce = cudaSuccess;
  cudaDeviceProp prop;
  ce = cudaGetDeviceProperties(&prop, device);
  if (ce == cudaSuccess && prop.major < 10) {
    b200nn_set_error("device %d is sm_%d%d; libb200nn is built for sm_100a (B200) only", device,
                     prop.major, prop.minor);
    delete ctx;
    return B200NN_ERR_CUDA;
  }
  if (ce == cudaSuccess) {
    ctx->sm_count = prop.multiProcessorCount;
    ce = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  }
  if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
  if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking);
  if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&ctx->ev_copy, cudaEventDisableTiming);
  if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&ctx->ev_side, cudaEventDisableTiming);
  if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&ctx->ev_sorted, cudaEventDisableTiming);
  for (auto &ev : ctx->ev)
    if (ce == cudaSuccess) ce = cudaEventCreate(&ev);
  if (ce == cudaSuccess) ce = cudaMallocHost(reinterpret_cast<void **>(&ctx->h_info), sizeof(DevInfo));
  if (ce == cudaSuccess) ce = cudaMallocHost(reinterpret_cast<void **>(&ctx->h_scalars), 8 * sizeof(long long));
  if (ce != cudaSuccess) {
    b200nn_set_error("ctx_create: %s", cudaGetErrorString(ce));
    b200nn_ctx_destroy(ctx);
    return B200NN_ERR_CUDA;
  }
  *ctx_out = ctx;
  return B200NN_OK;
}

void b200nn_ctx_destroy(b200nn_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (auto &b : ctx->ws)
    if (b.p) cudaFree(b.p);
  for (auto &ev : ctx->ev)
    if (ev) cudaEventDestroy(ev);
  if (ctx->h_info) cudaFreeHost(ctx->h_info);
  if (ctx->h_scalars) cudaFreeHost(ctx->h_scalars);
  if (ctx->side_stream) {
    cudaStreamSynchronize(ctx->side_stream);
    cudaStreamDestroy(ctx->side_stream);
  }
  if (ctx->copy_stream) {
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamDestroy(ctx->copy_stream);
  }
  if (ctx->ev_copy) cudaEventDestroy(ctx->ev_copy);
  if (ctx->ev_side) cudaEventDestroy(ctx->ev_side);
  if (ctx->ev_sorted) cudaEventDestroy(ctx->ev_sorted);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  (void)cudaGetLastError();
  delete ctx;
}

int b200nn_ctx_sync(b200nn_ctx *ctx) {
  RC_TRY(check_ctx(ctx));
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  return B200NN_OK;
}

void *b200nn_ctx_stream(b200nn_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

// ---- device-buffer entry points -------------------------------------------

int b200nn_dev_knn(b200nn_ctx *ctx, const double *d_ref, int64_t nr, const double *d_qry,
                   int64_t nq, int32_t d, int32_t k, int32_t *d_idx0, double *d_dist,
                   const b200nn_opts *opts, b200nn_stats *stats) {
  RC_TRY(check_ctx(ctx));
  b200nn_opts o = opts ? *opts : default_opts();
  if (stats) memset(stats, 0, sizeof(*stats));
  int64_t l0 = ctx->launches;
  int rc = knn_dispatch(ctx, d_ref, nr, d_qry, nq, d, k, o, d_idx0, d_dist, stats);
  if (stats) stats->gpu_launches = ctx->launches - l0;
  return rc;
}

int b200nn_dev_knn_candidates(b200nn_ctx *ctx, const double *d_ref, int64_t nr, const double *d_qry,
                              int64_t nq, int32_t d, int32_t n_cand, int32_t scan_mode,
                              int32_t *d_cand, float *d_tau, float *d_keys, double *h_scale_mu) {
  RC_TRY(check_ctx(ctx));
  if (!d_ref || nr <= 0 || nq <= 0 || d <= 0 || d > 61 || n_cand < 8 || n_cand > 48 || (n_cand & 7)) {
    b200nn_set_error("knn_candidates: invalid argument (d <= 61, n_cand a multiple of 8 in [8, 48])");
    return B200NN_ERR_INVALID;
  }
  return knn_tc_debug(ctx, d_ref, nr, d_qry, nq, d, n_cand, scan_mode & 3, d_keys, d_cand, d_tau, h_scale_mu);
}

int b200nn_dev_graphs(b200nn_ctx *ctx, const int32_t *d_idx0, int64_t n, int32_t k, double prune,
                      int32_t compute_snn, int64_t row_begin, int64_t row_end,
                      int64_t *nnz_nn_out, int64_t *nnz_snn_out, b200nn_stats *stats) {
  RC_TRY(check_ctx(ctx));
  if (!d_idx0) {
    b200nn_set_error("d_idx0 is NULL");
    return B200NN_ERR_INVALID;
  }
  if (stats) memset(stats, 0, sizeof(*stats));
  int64_t l0 = ctx->launches;
  int rc = graphs_launch(ctx, d_idx0, n, k, n, prune, true, compute_snn != 0, row_begin, row_end,
                         stats);
  if (stats) stats->gpu_launches = ctx->launches - l0;
  if (rc != B200NN_OK) return rc;
  if (nnz_nn_out) *nnz_nn_out = ctx->nn_graph.nnz;
  if (nnz_snn_out) *nnz_snn_out = compute_snn ? ctx->snn.nnz : 0;
  return B200NN_OK;
}

int b200nn_dev_result(b200nn_ctx *ctx, int32_t which, b200nn_dev_csr *out) {
  if (!ctx || !out) {
    b200nn_set_error("NULL argument");
    return B200NN_ERR_INVALID;
  }
  const CsrResult &r = which == 0 ? ctx->nn_graph : ctx->snn;
  if (!r.valid) {
    b200nn_set_error("no %s result in this context", which == 0 ? "nn Graph" : "SNN");
    return B200NN_ERR_STATE;
  }
  out->n_rows = r.n_rows;
  out->nnz = r.nnz;
  out->p = r.p;
  out->i = r.i;
  out->x = r.x;
  return B200NN_OK;
}

int b200nn_dev_copy_result(b200nn_ctx *ctx, int32_t which, int64_t *d_p, int32_t *d_i,
                           double *d_x) {
  RC_TRY(check_ctx(ctx));
  const CsrResult &r = which == 0 ? ctx->nn_graph : ctx->snn;
  if (!r.valid) {
    b200nn_set_error("no %s result in this context", which == 0 ? "nn Graph" : "SNN");
    return B200NN_ERR_STATE;
  }
  if (d_p)
    CU_TRY(cudaMemcpyAsync(d_p, r.p, sizeof(int64_t) * ((size_t)r.n_rows + 1),
                           cudaMemcpyDeviceToDevice, ctx->stream));
  if (d_i && r.nnz)
    CU_TRY(cudaMemcpyAsync(d_i, r.i, sizeof(int32_t) * (size_t)r.nnz, cudaMemcpyDeviceToDevice,
                           ctx->stream));
  if (d_x && r.nnz)
    CU_TRY(cudaMemcpyAsync(d_x, r.x, sizeof(double) * (size_t)r.nnz, cudaMemcpyDeviceToDevice,
                           ctx->stream));
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  return B200NN_OK;
}

// ---- host-buffer entry points ----------------------------------------------

int b200nn_knn(b200nn_ctx *ctx, const double *ref, int64_t nr, const double *qry, int64_t nq,
               int32_t d, int32_t k, int32_t *idx_out, double *dist_out, const b200nn_opts *opts,
               b200nn_stats *stats) {
  RC_TRY(check_ctx(ctx));
  b200nn_opts o = opts ? *opts : default_opts();
  if (stats) memset(stats, 0, sizeof(*stats));
  if (!ref || nr <= 0 || d <= 0 || k <= 0 || (!idx_out && !dist_out)) {
    b200nn_set_error("knn: invalid argument");
    return B200NN_ERR_INVALID;
  }
  if ((int64_t)k > nr) {
    b200nn_set_error("Cannot find more nearest neighbours than there are points");
    return B200NN_ERR_K_TOO_LARGE;
  }
  if (!qry) nq = nr;
  int64_t l0 = ctx->launches;
  Timer th(ctx, 6);
  RC_TRY(th.start());
  double *d_ref, *d_qry = nullptr;
  RC_TRY(upload_matrix_f64(ctx, ref, nr, d, o.layout, WS_STAGE_A, WS_REF64, &d_ref));
  if (qry) RC_TRY(upload_matrix_f64(ctx, qry, nq, d, o.layout, WS_STAGE_B, WS_QRY64, &d_qry));
  RC_TRY(th.stop(stats ? &stats->ms_h2d : nullptr));
  int32_t *d_idx0;
  double *d_dist;
  RC_TRY(ws_get(ctx, WS_IDX0, (size_t)nq * k, &d_idx0));
  RC_TRY(ws_get(ctx, WS_DIST, (size_t)nq * k, &d_dist));
  RC_TRY(knn_dispatch(ctx, d_ref, nr, d_qry, nq, d, k, o, d_idx0, d_dist, stats));
  RC_TRY(th.start());
  RC_TRY(download_knn(ctx, d_idx0, d_dist, nq, k, o.layout, idx_out, dist_out));
  RC_TRY(th.stop(stats ? &stats->ms_d2h : nullptr));
  if (stats) stats->gpu_launches = ctx->launches - l0;
  return B200NN_OK;
}

// ---------------------------------------------------------------------------
// resident reference set
// ---------------------------------------------------------------------------
struct b200nn_index {
  b200nn_ctx *owner = nullptr;
  double *d_ref = nullptr;  // [nr x d] row-major, own allocation (not a workspace slot)
  int64_t nr = 0;
  int32_t d = 0;
  // prepared operands of the tensor search (centring / scale, group-sorted FP16 rows, grouping, balls):
  // filled by the first query, reused by the later ones
  TcRefCache *cache = nullptr;
};

int b200nn_index_create(b200nn_ctx *ctx, const double *ref, int64_t nr, int32_t d, const b200nn_opts *opts,
                        b200nn_index **index_out) {
  RC_TRY(check_ctx(ctx));
  b200nn_opts o = opts ? *opts : default_opts();
  if (!ref || nr <= 0 || d <= 0 || !index_out) {
    b200nn_set_error("index_create: invalid argument");
    return B200NN_ERR_INVALID;
  }
  *index_out = nullptr;
  b200nn_index *ix = new (std::nothrow) b200nn_index;
  if (!ix) return B200NN_ERR_NOMEM;
  const size_t bytes = (size_t)nr * (size_t)d * sizeof(double);
  if (cudaMalloc(&ix->d_ref, bytes) != cudaSuccess) {
    cudaGetLastError();
    delete ix;
    b200nn_set_error("index_create: cannot allocate %zu bytes of device memory", bytes);
    return B200NN_ERR_NOMEM;
  }
  int rc = B200NN_OK;
  if (o.layout == B200NN_ROW_MAJOR) {
    if (cudaMemcpyAsync(ix->d_ref, ref, bytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) rc = B200NN_ERR_CUDA;
  } else {
    double *raw = nullptr;
    rc = ws_get(ctx, WS_STAGE_A, (size_t)nr * d, &raw);
    if (rc == B200NN_OK && cudaMemcpyAsync(raw, ref, bytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
      rc = B200NN_ERR_CUDA;
    if (rc == B200NN_OK) rc = transpose_f64(ctx, raw, ix->d_ref, nr, d);
  }
  if (rc == B200NN_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = B200NN_ERR_CUDA;
  if (rc != B200NN_OK) {
    if (rc == B200NN_ERR_CUDA) b200nn_set_error("index_create: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(ix->d_ref);
    delete ix;
    return rc;
  }
  ix->owner = ctx;
  ix->nr = nr;
  ix->d = d;
  ix->cache = tc_ref_cache_create();
  *index_out = ix;
  return B200NN_OK;
}

void b200nn_index_destroy(b200nn_ctx *ctx, b200nn_index *index) {
  if (!index) return;
  if (ctx && ctx == index->owner) {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
  }
  if (index->d_ref) cudaFree(index->d_ref);
  tc_ref_cache_destroy(index->cache);
  delete index;
}

int64_t b200nn_index_rows(const b200nn_index *index) { return index ? index->nr : 0; }
int32_t b200nn_index_dims(const b200nn_index *index) { return index ? index->d : 0; }

int b200nn_index_knn(b200nn_ctx *ctx, const b200nn_index *index, const double *qry, int64_t nq, int32_t k,
                     int32_t *idx_out, double *dist_out, const b200nn_opts *opts, b200nn_stats *stats) {
  RC_TRY(check_ctx(ctx));
  b200nn_opts o = opts ? *opts : default_opts();
  if (stats) memset(stats, 0, sizeof(*stats));
  if (!index || index->owner != ctx) {
    b200nn_set_error("index_knn: the index does not belong to this context");
    return B200NN_ERR_STATE;
  }
  if (k <= 0 || (!idx_out && !dist_out) || (qry && nq <= 0)) {
    b200nn_set_error("index_knn: invalid argument");
    return B200NN_ERR_INVALID;
  }
  const int64_t nr = index->nr;
  const int32_t d = index->d;
  if ((int64_t)k > nr) {
    b200nn_set_error("Cannot find more nearest neighbours than there are points");
    return B200NN_ERR_K_TOO_LARGE;
  }
  if (!qry) nq = nr;
  int64_t l0 = ctx->launches;
  Timer th(ctx, 6);
  RC_TRY(th.start());
  double *d_qry = nullptr;
  if (qry) RC_TRY(upload_matrix_f64(ctx, qry, nq, d, o.layout, WS_STAGE_B, WS_QRY64, &d_qry));
  RC_TRY(th.stop(stats ? &stats->ms_h2d : nullptr));
  int32_t *d_idx0;
  double *d_dist;
  RC_TRY(ws_get(ctx, WS_IDX0, (size_t)nq * k, &d_idx0));
  RC_TRY(ws_get(ctx, WS_DIST, (size_t)nq * k, &d_dist));
  RC_TRY(knn_dispatch(ctx, index->d_ref, nr, d_qry, nq, d, k, o, d_idx0, d_dist, stats, index->cache));
  RC_TRY(th.start());
  RC_TRY(download_knn(ctx, d_idx0, d_dist, nq, k, o.layout, idx_out, dist_out));
  RC_TRY(th.stop(stats ? &stats->ms_d2h : nullptr));
  if (stats) stats->gpu_launches = ctx->launches - l0;
  return B200NN_OK;
}

int b200nn_snn_count(b200nn_ctx *ctx, const int32_t *nn_ranked, int64_t n, int32_t k, double prune,
                     const b200nn_opts *opts, int64_t *nnz_out, b200nn_stats *stats) {
  RC_TRY(check_ctx(ctx));
  b200nn_opts o = opts ? *opts : default_opts();
  if (stats) memset(stats, 0, sizeof(*stats));
  ctx->pending_snn = false;
  if (!nn_ranked || n <= 0 || k <= 0 || !nnz_out) {
    b200nn_set_error("snn: invalid argument");
    return B200NN_ERR_INVALID;
  }
  int64_t l0 = ctx->launches;
  Timer th(ctx, 6);
  RC_TRY(th.start());
  int32_t *d_idx0;
  RC_TRY(upload_nn_ranked(ctx, nn_ranked, n, k, o.layout, &d_idx0));
  RC_TRY(th.stop(stats ? &stats->ms_h2d : nullptr));
  RC_TRY(graphs_launch(ctx, d_idx0, n, k, n, prune, false, true, 0, n, stats));
  if (ctx->snn.nnz >= 0x7fffffffLL) {
    b200nn_set_error("SNN graph has %lld entries, more than a dgCMatrix can hold",
                     (long long)ctx->snn.nnz);
    return B200NN_ERR_NNZ_OVERFLOW;
  }
  *nnz_out = ctx->snn.nnz;
  ctx->pending_snn = true;
  if (stats) stats->gpu_launches = ctx->launches - l0;
  return B200NN_OK;
}

int b200nn_snn_fill(b200nn_ctx *ctx, int32_t *p, int32_t *i, double *x) {
  RC_TRY(check_ctx(ctx));
  if (!ctx->pending_snn) {
    b200nn_set_error("b200nn_snn_fill without a successful b200nn_snn_count");
    return B200NN_ERR_STATE;
  }
  return download_csr(ctx, ctx->snn, p, i, x);
}

int b200nn_nn_graph_count(b200nn_ctx *ctx, const int32_t *nn_ranked, int64_t n_rows, int32_t k,
                          int64_t n, const b200nn_opts *opts, int64_t *nnz_out,
                          b200nn_stats *stats) {
  RC_TRY(check_ctx(ctx));
  b200nn_opts o = opts ? *opts : default_opts();
  if (stats) memset(stats, 0, sizeof(*stats));
  ctx->pending_nn = false;
  if (!nn_ranked || n_rows <= 0 || n <= 0 || k <= 0 || !nnz_out) {
    b200nn_set_error("nn_graph: invalid argument");
    return B200NN_ERR_INVALID;
  }
  int64_t l0 = ctx->launches;
  int32_t *d_idx0;
  RC_TRY(upload_nn_ranked(ctx, nn_ranked, n_rows, k, o.layout, &d_idx0));
  RC_TRY(graphs_launch(ctx, d_idx0, n_rows, k, n, 0.0, true, false, 0, n_rows, stats));
  *nnz_out = ctx->nn_graph.nnz;
  ctx->pending_nn = true;
  if (stats) stats->gpu_launches = ctx->launches - l0;
  return B200NN_OK;
}

int b200nn_nn_graph_fill(b200nn_ctx *ctx, int32_t *p, int32_t *i, double *x) {
  RC_TRY(check_ctx(ctx));
  if (!ctx->pending_nn) {
    b200nn_set_error("b200nn_nn_graph_fill without a successful b200nn_nn_graph_count");
    return B200NN_ERR_STATE;
  }
  return download_csr(ctx, ctx->nn_graph, p, i, x);
}

int b200nn_find_neighbors_count(b200nn_ctx *ctx, const double *data, int64_t n, int32_t d,
                                int32_t k, double prune, int32_t compute_snn, int32_t *idx_out,
                                double *dist_out, int32_t *nn_p_out, int32_t *nn_i_out, double *nn_x_out,
                                const b200nn_opts *opts, int64_t *nnz_nn_out, int64_t *nnz_snn_out,
                                b200nn_stats *stats) {
  RC_TRY(check_ctx(ctx));
  b200nn_opts o = opts ? *opts : default_opts();
  if (stats) memset(stats, 0, sizeof(*stats));
  ctx->pending_nn = ctx->pending_snn = false;
  if (!data || n <= 0 || d <= 0 || k <= 0) {
    b200nn_set_error("find_neighbors: invalid argument");
    return B200NN_ERR_INVALID;
  }
  if ((int64_t)k > n) {
    b200nn_set_error("Cannot find more nearest neighbours than there are points");
    return B200NN_ERR_K_TOO_LARGE;
  }
  int64_t l0 = ctx->launches;
  Timer th(ctx, 6);
  RC_TRY(th.start());
  double *d_ref;
  RC_TRY(upload_matrix_f64(ctx, data, n, d, o.layout, WS_STAGE_A, WS_REF64, &d_ref));
  RC_TRY(th.stop(stats ? &stats->ms_h2d : nullptr));
  int32_t *d_idx0;
  double *d_dist;
  RC_TRY(ws_get(ctx, WS_IDX0, (size_t)n * k, &d_idx0));
  RC_TRY(ws_get(ctx, WS_DIST, (size_t)n * k, &d_dist));
  RC_TRY(knn_dispatch(ctx, d_ref, n, nullptr, n, d, k, o, d_idx0, d_dist, stats));
  // the kNN result travels to the host while the graphs are computed
  const bool want_knn = idx_out || dist_out;
  if (want_knn) RC_TRY(download_knn_begin(ctx, d_idx0, d_dist, n, k, o.layout, idx_out, dist_out));
  b200nn_stats gs;
  memset(&gs, 0, sizeof(gs));
  NnGraphHostOut early;
  early.p = nn_p_out;
  early.i = nn_i_out;
  early.x = nn_x_out;
  const bool want_early = nn_p_out || nn_i_out || nn_x_out;
  int grc = graphs_launch(ctx, d_idx0, n, k, n, prune, true, compute_snn != 0, 0, n, &gs, &early);
  if (grc != B200NN_OK) {
    if (want_knn || want_early) cudaStreamSynchronize(ctx->copy_stream);  // no copy into caller memory stays in flight
    return grc;
  }
  if (stats) {
    stats->ms_nn_graph = gs.ms_nn_graph;
    stats->ms_snn = gs.ms_snn;
    stats->nnz_nn = gs.nnz_nn;
    stats->nnz_snn = gs.nnz_snn;
    stats->snn_expanded = gs.snn_expanded;
    stats->max_indegree = gs.max_indegree;
  }
  if (want_knn || want_early) {
    RC_TRY(th.start());
    RC_TRY(download_knn_end(ctx));
    RC_TRY(th.stop(stats ? &stats->ms_d2h : nullptr));  // what is left of the copy after the graphs
  }
  if (compute_snn && ctx->snn.nnz >= 0x7fffffffLL) {
    b200nn_set_error("SNN graph has %lld entries, more than a dgCMatrix can hold",
                     (long long)ctx->snn.nnz);
    return B200NN_ERR_NNZ_OVERFLOW;
  }
  if (nnz_nn_out) *nnz_nn_out = ctx->nn_graph.nnz;
  if (nnz_snn_out) *nnz_snn_out = compute_snn ? ctx->snn.nnz : 0;
  ctx->pending_nn = true;
  ctx->pending_snn = compute_snn != 0;
  if (stats) stats->gpu_launches = ctx->launches - l0;
  return B200NN_OK;
}

int b200nn_find_neighbors_fill(b200nn_ctx *ctx, int32_t *nn_p, int32_t *nn_i, double *nn_x,
                               int32_t *snn_p, int32_t *snn_i, double *snn_x) {
  RC_TRY(check_ctx(ctx));
  if (!ctx->pending_nn) {
    b200nn_set_error("b200nn_find_neighbors_fill without a successful _count");
    return B200NN_ERR_STATE;
  }
  Timer th(ctx, 6);
  RC_TRY(th.start());
  if (nn_p || nn_i || nn_x) RC_TRY(download_csr(ctx, ctx->nn_graph, nn_p, nn_i, nn_x));
  if (snn_p || snn_i || snn_x) {
    if (!ctx->pending_snn) {
      b200nn_set_error("SNN was not computed by the last b200nn_find_neighbors_count");
      return B200NN_ERR_STATE;
    }
    RC_TRY(download_csr(ctx, ctx->snn, snn_p, snn_i, snn_x));
  }
  return B200NN_OK;
}

}  // extern "C"
