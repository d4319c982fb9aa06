// multi.cu -- several devices driven from ONE process (b200nn_multi_*, see include/b200nn.h).
//
// FindNeighbors shards naturally with one exchange step.  The R call site
// (/root/reference/R/clustering.R:633-682) is a single-threaded process, so the sharding lives
// inside the library: one b200nn_ctx, one host worker thread and one stream per device.
//
//   1. every device uploads ITS slice of the host embedding over its own PCIe link and the
//      slices are exchanged between the devices over NVLink (peer copies), so the embedding
//      crosses PCIe once in total instead of once per device;
//   2. device g computes the kNN rows of its share of the query cells against the full reference
//      set.  With peer access between all devices the shares are whole GROUPS of the tile-pruned
//      search (opts.shard_rank / shard_world: every device derives the same deterministic grouping
//      and takes the groups dealt to it), so every share keeps full 256-query work items;
//   3. k_bcast_rows -- the exchange step, one kernel: every finished index row is stored straight
//      into the index buffer of EVERY peer (4*n*k bytes per device over NVLink, peer stores), and
//      its distance row into the buffer of the one device that delivers that row to the host.
//      Without peer access: contiguous row shares and cudaMemcpyPeer of the index slices;
//   4. every device transposes the full index and emits the SNN rows of its share; device 0 also
//      emits the nn Graph;
//   5. every device copies its rows of idx / dist / SNN straight into the caller's host
//      buffers (row offsets and the nnz prefix are known), so the result is ONE dgCMatrix.
//
// A device id may be listed more than once: the shards then share a device (used by the tests
// on a single-GPU box to exercise the sharding, exchange and gathering logic).
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "internal.cuh"

namespace {

// barrier that also tells every participant whether all of them are still healthy
class Gate {
 public:
  explicit Gate(int n) : n_(n) {}
  bool wait(bool my_ok) {
    std::unique_lock<std::mutex> lk(mu_);
    if (!my_ok) ok_ = false;
    const int gen = gen_;
    if (++count_ == n_) {
      count_ = 0;
      ++gen_;
      cv_.notify_all();
    } else {
      cv_.wait(lk, [&] { return gen != gen_; });
    }
    return ok_;
  }

 private:
  std::mutex mu_;
  std::condition_variable cv_;
  int n_, count_ = 0, gen_ = 0;
  bool ok_ = true;
};

struct Shard {
  b200nn_ctx *ctx = nullptr;
  int dev = 0;
  int64_t b = 0, e = 0;        // rows [b, e) of the query set
  double *d_ref = nullptr;     // full reference set on this device
  double *d_qry = nullptr;     // full query set (query != reference calls)
  int32_t *d_idx = nullptr;    // full index [nq x k] (own rows computed, the others received)
  double *d_dist = nullptr;    // distances of rows [b, e): what this device delivers to the host
  double *d_dist_full = nullptr;  // group-aligned shares: scratch [nq x k], rows of the own share valid
  int rc = B200NN_OK;
  std::string err;
  b200nn_stats st;
  int64_t nnz_snn = 0, nnz_nn = 0;
  double ms_upload = 0, ms_exchange = 0, ms_d2h = 0;
};

double now_ms() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

#define SH_TRY(expr)                         \
  do {                                       \
    int _rc = (expr);                        \
    if (_rc != B200NN_OK) {                  \
      s.rc = _rc;                            \
      s.err = b200nn_last_error();           \
      return _rc;                            \
    }                                        \
  } while (0)

#define SH_CU(expr)                                                                     \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      (void)cudaGetLastError();                                                         \
      s.rc = _e == cudaErrorMemoryAllocation ? B200NN_ERR_NOMEM : B200NN_ERR_CUDA;      \
      char buf[256];                                                                    \
      snprintf(buf, sizeof(buf), "CUDA error %s at %s:%d (device %d): %s", cudaGetErrorName(_e), __FILE__, \
               __LINE__, s.dev, cudaGetErrorString(_e));                                \
      s.err = buf;                                                                      \
      return s.rc;                                                                      \
    }                                                                                   \
  } while (0)

// rows [b, e) of a host matrix [n x d] (layout) -> rows [b, e) of the device row-major matrix
int upload_rows(Shard &s, const double *h, int64_t n, int d, int layout, double *d_full, int64_t b, int64_t e) {
  const int64_t rows = e - b;
  if (rows <= 0) return B200NN_OK;
  cudaStream_t st = s.ctx->stream;
  if (layout == B200NN_ROW_MAJOR) {
    SH_CU(cudaMemcpyAsync(d_full + b * d, h + b * d, sizeof(double) * (size_t)rows * d, cudaMemcpyHostToDevice, st));
  } else {
    double *stage;
    SH_TRY(ws_get(s.ctx, WS_STAGE_A, (size_t)rows * d, &stage));
    // column c of the slice: rows contiguous doubles at h + c * n + b
    SH_CU(cudaMemcpy2DAsync(stage, sizeof(double) * (size_t)rows, h + b, sizeof(double) * (size_t)n,
                            sizeof(double) * (size_t)rows, (size_t)d, cudaMemcpyHostToDevice, st));
    SH_TRY(transpose_f64(s.ctx, stage, d_full + b * d, rows, d));
  }
  return B200NN_OK;
}

// copy `count` elements at element offset `off` of the peer's buffer into the same place of mine
template <typename T>
int pull_from(Shard &s, const Shard &peer, T *mine, const T *theirs, int64_t off, int64_t count) {
  if (count <= 0) return B200NN_OK;
  if (peer.dev == s.dev)
    SH_CU(cudaMemcpyAsync(mine + off, theirs + off, sizeof(T) * (size_t)count, cudaMemcpyDeviceToDevice, s.ctx->stream));
  else
    SH_CU(cudaMemcpyPeerAsync(mine + off, s.dev, theirs + off, peer.dev, sizeof(T) * (size_t)count, s.ctx->stream));
  return B200NN_OK;
}

// own kNN rows -> the caller's host matrices [nq x k] (layout), 1-based indices; copy stream
int download_rows(Shard &s, int64_t nq, int k, int layout, int32_t *h_idx, double *h_dist) {
  const int64_t rows = s.e - s.b;
  if (rows <= 0 || (!h_idx && !h_dist)) return B200NN_OK;
  b200nn_ctx *ctx = s.ctx;
  const size_t cnt = (size_t)rows * k;
  const int32_t *src_i = s.d_idx + s.b * k;
  int32_t *oi = nullptr;
  const double *od = s.d_dist;
  if (h_idx) {
    SH_TRY(ws_get(ctx, WS_OUT_IDX, cnt, &oi));
    if (layout == B200NN_ROW_MAJOR)
      SH_TRY(add_i32(ctx, src_i, oi, (int64_t)cnt, 1));
    else
      SH_TRY(transpose_i32_add(ctx, src_i, oi, k, rows, 1));  // -> column-major [rows x k]
  }
  if (h_dist && layout != B200NN_ROW_MAJOR) {
    double *o;
    SH_TRY(ws_get(ctx, WS_OUT_DIST, cnt, &o));
    SH_TRY(transpose_f64(ctx, s.d_dist, o, k, rows));
    od = o;
  }
  SH_CU(cudaEventRecord(ctx->ev_copy, ctx->stream));
  SH_CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_copy, 0));
  if (layout == B200NN_ROW_MAJOR) {
    if (h_idx)
      SH_CU(cudaMemcpyAsync(h_idx + s.b * k, oi, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->copy_stream));
    if (h_dist)
      SH_CU(cudaMemcpyAsync(h_dist + s.b * k, od, cnt * sizeof(double), cudaMemcpyDeviceToHost, ctx->copy_stream));
  } else {
    if (h_idx)
      SH_CU(cudaMemcpy2DAsync(h_idx + s.b, sizeof(int32_t) * (size_t)nq, oi, sizeof(int32_t) * (size_t)rows,
                              sizeof(int32_t) * (size_t)rows, (size_t)k, cudaMemcpyDeviceToHost, ctx->copy_stream));
    if (h_dist)
      SH_CU(cudaMemcpy2DAsync(h_dist + s.b, sizeof(double) * (size_t)nq, od, sizeof(double) * (size_t)rows,
                              sizeof(double) * (size_t)rows, (size_t)k, cudaMemcpyDeviceToHost, ctx->copy_stream));
  }
  return B200NN_OK;
}

// The exchange step of the sharded search as one kernel over peer memory.  One warp per row slot:
// the index row goes to every other shard's full index buffer, the distance row to the shard that
// owns the contiguous output slice the row falls into (possibly this one).
constexpr int MAX_DIRECT = 16;
struct PeerPtrs {
  int32_t *idx[MAX_DIRECT];
  double *dist[MAX_DIRECT];
};
__global__ void __launch_bounds__(256)
k_bcast_rows(const int32_t *__restrict__ rows, int64_t n_slots, int64_t row_b, int k, int64_t per, int G, int self_g,
             const int32_t *__restrict__ idx_local, const double *__restrict__ dist_local, PeerPtrs P) {
  const int lane = threadIdx.x & 31;
  const int64_t slot = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (slot >= n_slots) return;
  const int64_t q = rows ? (int64_t)rows[slot] : row_b + slot;
  if (q < 0) return;
  const int ho = (int)(q / per);
  double *dd = P.dist[ho] + (q - (int64_t)ho * per) * k;
  for (int t = lane; t < k; t += 32) {
    const int32_t v = idx_local[q * k + t];
    for (int h = 0; h < G; ++h)
      if (h != self_g) P.idx[h][q * k + t] = v;
    dd[t] = dist_local[q * k + t];
  }
}

__global__ void k_narrow_add(const int64_t *__restrict__ in, int32_t *__restrict__ out, int64_t n, int64_t add) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int32_t)(in[i] + add);
}

struct Job {
  // inputs
  const double *ref = nullptr, *qry = nullptr;
  int64_t nr = 0, nq = 0;
  int d = 0, k = 0;
  double prune = 0;
  bool graphs = false, compute_snn = false;
  int32_t *idx_out = nullptr;
  double *dist_out = nullptr;
  NnGraphHostOut early;
  b200nn_opts o;
};

}  // namespace

struct b200nn_multi {
  std::vector<Shard> sh;
  bool direct = false;  // every pair of shards can store into each other's memory
  // what the last *_count call left pending for *_fill
  bool pending = false, pending_snn = false;
  int64_t n = 0;
  int k = 0;
};

namespace {

int shard_count_phase(b200nn_multi *m, int g, const Job &J, Gate &gate) {
  Shard &s = m->sh[g];
  const int G = (int)m->sh.size();
  b200nn_ctx *ctx = s.ctx;
  s.rc = B200NN_OK;
  s.err.clear();
  memset(&s.st, 0, sizeof(s.st));
  s.nnz_snn = s.nnz_nn = 0;
  bool ok = true;
  static const bool trace = getenv("B200NN_TRACE") != nullptr;  // phase progress of every shard on stderr
  int phase_no = 0;
  const double t_call = now_ms();
  auto phase = [&](int rc) {  // all shards leave a phase together; a failure anywhere ends the call
    if (trace) fprintf(stderr, "[b200nn multi] shard %d/%d (device %d) finished phase %d rc=%d at %.2f ms, waits\n", g, G, s.dev, phase_no, rc, now_ms() - t_call);
    ok = gate.wait(rc == B200NN_OK) && rc == B200NN_OK;
    if (trace) fprintf(stderr, "[b200nn multi] shard %d/%d leaves phase %d (%s) at %.2f ms\n", g, G, phase_no, ok ? "ok" : "failed", now_ms() - t_call);
    ++phase_no;
    return ok;
  };
  const bool self = J.qry == nullptr;
  const int64_t nq = self ? J.nr : J.nq;

  // ---- 1. upload own slices ------------------------------------------------------------------
  auto p1 = [&]() -> int {
    SH_CU(cudaSetDevice(s.dev));
    const double t0 = now_ms();
    SH_TRY(ws_get(ctx, WS_REF64, (size_t)J.nr * J.d, &s.d_ref));
    const int64_t per_r = ceil_div64(J.nr, G);
    const int64_t rb = std::min<int64_t>(g * per_r, J.nr), re = std::min<int64_t>(rb + per_r, J.nr);
    SH_TRY(upload_rows(s, J.ref, J.nr, J.d, J.o.layout, s.d_ref, rb, re));
    // NA / NaN / Inf are refused before any search kernel sees them (nn2's .C(NAOK = FALSE)): every
    // shard checks the slice it uploaded, and a failure anywhere stops all shards at this phase
    SH_TRY(check_finite_f64(ctx, s.d_ref + rb * J.d, (re - rb) * (int64_t)J.d, "data"));
    s.d_qry = nullptr;
    if (!self) {  // every device needs only its own query rows
      SH_TRY(ws_get(ctx, WS_QRY64, (size_t)nq * J.d, &s.d_qry));
      SH_TRY(upload_rows(s, J.qry, nq, J.d, J.o.layout, s.d_qry, s.b, s.e));
      SH_TRY(check_finite_f64(ctx, s.d_qry + s.b * J.d, (s.e - s.b) * (int64_t)J.d, "query"));
    }
    SH_TRY(ws_get(ctx, WS_IDX0, (size_t)nq * J.k, &s.d_idx));
    SH_TRY(ws_get(ctx, WS_DIST, (size_t)std::max<int64_t>(s.e - s.b, 1) * J.k, &s.d_dist));
    SH_CU(cudaStreamSynchronize(ctx->stream));
    s.ms_upload = now_ms() - t0;
    return B200NN_OK;
  };
  if (!phase(p1())) return s.rc;

  // ---- 2. the other slices of the reference set arrive over NVLink ----------------------------
  auto p2 = [&]() -> int {
    const double t0 = now_ms();
    const int64_t per_r = ceil_div64(J.nr, G);
    for (int h = 1; h < G; ++h) {
      const int pg = (g + h) % G;  // staggered so that the peers are not all read in the same order
      const int64_t rb = std::min<int64_t>(pg * per_r, J.nr), re = std::min<int64_t>(rb + per_r, J.nr);
      SH_TRY(pull_from(s, m->sh[pg], s.d_ref, m->sh[pg].d_ref, rb * J.d, (re - rb) * J.d));
    }
    SH_CU(cudaStreamSynchronize(ctx->stream));
    s.ms_exchange = now_ms() - t0;
    return B200NN_OK;
  };
  if (!phase(p2())) return s.rc;

  // ---- 3. kNN of the own share of the query rows (+ the exchange kernel) -----------------------
  const bool group_shares = self && G > 1 && m->direct && G <= MAX_DIRECT;
  auto p3a = [&]() -> int {  // full-size scratch for the distances of a group-aligned share
    if (group_shares) SH_TRY(ws_get(ctx, WS_DIST_FULL, (size_t)nq * J.k, &s.d_dist_full));
    return B200NN_OK;
  };
  if (!phase(p3a())) return s.rc;  // every shard's buffers exist before anybody stores into them
  auto p3 = [&]() -> int {
    const int64_t rows = s.e - s.b;
    if (group_shares) {
      b200nn_opts o1 = J.o;
      o1.shard_rank = g;
      o1.shard_world = G;
      SH_TRY(knn_dispatch(ctx, s.d_ref, J.nr, nullptr, J.nr, J.d, J.k, o1, s.d_idx, s.d_dist_full, &s.st));
      PeerPtrs P;
      memset(&P, 0, sizeof(P));
      for (int h = 0; h < G; ++h) {
        P.idx[h] = m->sh[h].d_idx;
        P.dist[h] = m->sh[h].d_dist;
      }
      const int64_t per = ceil_div64(nq, G);
      const int64_t slots = ctx->owned_rows ? ctx->owned_slots : ctx->owned_e - ctx->owned_b;
      if (slots > 0) {
        k_bcast_rows<<<(unsigned)ceil_div64(slots * 32, 256), 256, 0, ctx->stream>>>(
            ctx->owned_rows, slots, ctx->owned_b, J.k, per, G, g, s.d_idx, s.d_dist_full, P);
        ctx->launches++;
        SH_CU(cudaGetLastError());
      }
    } else if (rows > 0) {
      const double *q = self ? s.d_ref + s.b * J.d : s.d_qry + s.b * J.d;
      if (self && G == 1) q = nullptr;  // one shard: the plain self search
      SH_TRY(knn_dispatch(ctx, s.d_ref, J.nr, q, rows, J.d, J.k, J.o, s.d_idx + s.b * J.k, s.d_dist, &s.st));
    }
    SH_CU(cudaStreamSynchronize(ctx->stream));
    return B200NN_OK;
  };
  if (!phase(p3())) return s.rc;

  // ---- 4. own rows to the host (copy stream); index exchange where it is still due; graphs -------
  auto p4 = [&]() -> int {
    SH_TRY(download_rows(s, nq, J.k, J.o.layout, J.idx_out, J.dist_out));
    if (!J.graphs) return B200NN_OK;
    if (!group_shares) {
      const double t0 = now_ms();
      for (int h = 1; h < G; ++h) {
        const Shard &p = m->sh[(g + h) % G];
        SH_TRY(pull_from(s, p, s.d_idx, p.d_idx, p.b * J.k, (p.e - p.b) * J.k));
      }
      SH_CU(cudaStreamSynchronize(ctx->stream));
      s.ms_exchange += now_ms() - t0;
    }
    b200nn_stats gs;
    memset(&gs, 0, sizeof(gs));
    SH_TRY(graphs_launch(ctx, s.d_idx, J.nr, J.k, J.nr, J.prune, g == 0, J.compute_snn, s.b, s.e, &gs,
                         g == 0 ? &J.early : nullptr));
    s.st.ms_nn_graph = gs.ms_nn_graph;
    s.st.ms_snn = gs.ms_snn;
    s.st.snn_expanded = gs.snn_expanded;
    s.st.max_indegree = gs.max_indegree;
    s.nnz_nn = g == 0 ? ctx->nn_graph.nnz : 0;
    s.nnz_snn = J.compute_snn ? ctx->snn.nnz : 0;
    return B200NN_OK;
  };
  if (!phase(p4())) {
    cudaStreamSynchronize(ctx->copy_stream);
    return s.rc;
  }
  // ---- 5. the copies into the caller's buffers have landed --------------------------------------
  auto p5 = [&]() -> int {
    const double t0 = now_ms();
    SH_CU(cudaStreamSynchronize(ctx->copy_stream));
    s.ms_d2h = now_ms() - t0;
    return B200NN_OK;
  };
  phase(p5());
  return s.rc;
}

int shard_fill_phase(b200nn_multi *m, int g, int32_t *nn_p, int32_t *nn_i, double *nn_x, int32_t *snn_p,
                     int32_t *snn_i, double *snn_x) {
  Shard &s = m->sh[g];
  b200nn_ctx *ctx = s.ctx;
  s.rc = B200NN_OK;
  s.err.clear();
  SH_CU(cudaSetDevice(s.dev));
  cudaStream_t st = ctx->stream;
  if (g == 0 && (nn_p || nn_i || nn_x)) {
    const CsrResult &r = ctx->nn_graph;
    if (nn_p) {
      int32_t *p32;
      SH_TRY(ws_get(ctx, WS_OUT_P32, (size_t)r.n_rows + 1, &p32));
      SH_TRY(narrow_i64_to_i32(ctx, r.p, p32, r.n_rows + 1));
      SH_CU(cudaMemcpyAsync(nn_p, p32, sizeof(int32_t) * ((size_t)r.n_rows + 1), cudaMemcpyDeviceToHost, st));
      SH_CU(cudaStreamSynchronize(st));  // p32 is reused below
    }
    if (nn_i && r.nnz) SH_CU(cudaMemcpyAsync(nn_i, r.i, sizeof(int32_t) * (size_t)r.nnz, cudaMemcpyDeviceToHost, st));
    if (nn_x && r.nnz) SH_CU(cudaMemcpyAsync(nn_x, r.x, sizeof(double) * (size_t)r.nnz, cudaMemcpyDeviceToHost, st));
  }
  if (m->pending_snn && (snn_p || snn_i || snn_x)) {
    const CsrResult &r = ctx->snn;
    int64_t off = 0;
    for (int h = 0; h < g; ++h) off += m->sh[h].nnz_snn;
    const int64_t rows = s.e - s.b;
    if (snn_p) {
      // rows [b, e) of the global p: local offsets + the entries of the shards before this one;
      // the last shard also writes p[n]
      const int64_t cnt = rows + (g + 1 == (int)m->sh.size() ? 1 : 0);
      if (cnt > 0) {
        int32_t *p32;
        SH_TRY(ws_get(ctx, WS_OUT_P32, (size_t)cnt, &p32));
        k_narrow_add<<<(unsigned)ceil_div64(cnt, 256), 256, 0, st>>>(r.p, p32, cnt, off);
        ctx->launches++;
        SH_CU(cudaGetLastError());
        SH_CU(cudaMemcpyAsync(snn_p + s.b, p32, sizeof(int32_t) * (size_t)cnt, cudaMemcpyDeviceToHost, st));
      }
    }
    if (snn_i && r.nnz)
      SH_CU(cudaMemcpyAsync(snn_i + off, r.i, sizeof(int32_t) * (size_t)r.nnz, cudaMemcpyDeviceToHost, st));
    if (snn_x && r.nnz)
      SH_CU(cudaMemcpyAsync(snn_x + off, r.x, sizeof(double) * (size_t)r.nnz, cudaMemcpyDeviceToHost, st));
  }
  SH_CU(cudaStreamSynchronize(st));
  return B200NN_OK;
}

template <typename F>
int run_shards(b200nn_multi *m, F fn) {
  const int G = (int)m->sh.size();
  std::vector<std::thread> th;
  th.reserve(G);
  for (int g = 1; g < G; ++g) th.emplace_back([&, g] { fn(g); });
  fn(0);
  for (auto &t : th) t.join();
  for (int g = 0; g < G; ++g)
    if (m->sh[g].rc != B200NN_OK) {
      b200nn_set_error("%s", m->sh[g].err.c_str());
      return m->sh[g].rc;
    }
  return B200NN_OK;
}

void merge_stats(const b200nn_multi *m, b200nn_stats *out) {
  if (!out) return;
  memset(out, 0, sizeof(*out));
  for (const Shard &s : m->sh) {
    const b200nn_stats &a = s.st;
    out->ms_h2d = std::max(out->ms_h2d, s.ms_upload + s.ms_exchange);
    out->ms_prep = std::max(out->ms_prep, a.ms_prep);
    out->ms_knn_cand = std::max(out->ms_knn_cand, a.ms_knn_cand);
    out->ms_knn_pass2 = std::max(out->ms_knn_pass2, a.ms_knn_pass2);
    out->ms_knn_rerank = std::max(out->ms_knn_rerank, a.ms_knn_rerank);
    out->ms_knn_fallback = std::max(out->ms_knn_fallback, a.ms_knn_fallback);
    out->ms_nn_graph = std::max(out->ms_nn_graph, a.ms_nn_graph);
    out->ms_snn = std::max(out->ms_snn, a.ms_snn);
    out->ms_d2h = std::max(out->ms_d2h, s.ms_d2h);
    out->n_uncertified += a.n_uncertified;
    out->nnz_nn += s.nnz_nn;
    out->nnz_snn += s.nnz_snn;
    out->snn_expanded += a.snn_expanded;
    out->max_indegree = std::max(out->max_indegree, a.max_indegree);
    out->gpu_launches += a.gpu_launches;
    out->knn_tiles_visited += a.knn_tiles_visited;
    out->knn_tiles_total += a.knn_tiles_total;
  }
}

void set_rows(b200nn_multi *m, int64_t nq) {
  const int G = (int)m->sh.size();
  const int64_t per = ceil_div64(nq > 0 ? nq : 1, G);
  for (int g = 0; g < G; ++g) {
    m->sh[g].b = std::min<int64_t>(g * per, nq);
    m->sh[g].e = std::min<int64_t>(m->sh[g].b + per, nq);
  }
}

int run_count(b200nn_multi *m, const Job &J, b200nn_stats *stats) {
  const int G = (int)m->sh.size();
  std::vector<int64_t> l0(G);
  for (int g = 0; g < G; ++g) l0[g] = m->sh[g].ctx->launches;
  Gate gate(G);
  int rc = run_shards(m, [&](int g) { shard_count_phase(m, g, J, gate); });
  for (int g = 0; g < G; ++g) m->sh[g].st.gpu_launches = m->sh[g].ctx->launches - l0[g];
  merge_stats(m, stats);
  return rc;
}

}  // namespace

extern "C" {

int b200nn_multi_create(const int32_t *devices, int32_t n_devices, b200nn_multi **out) {
  if (!out) {
    b200nn_set_error("multi_create: out is NULL");
    return B200NN_ERR_INVALID;
  }
  *out = nullptr;
  int visible = b200nn_device_count();
  if (visible <= 0) {
    b200nn_set_error("no CUDA device available; libb200nn has no CPU fallback");
    return B200NN_ERR_CUDA;
  }
  if (n_devices <= 0) n_devices = visible;  // all visible devices
  if (n_devices > 64) {
    b200nn_set_error("multi_create: at most 64 shards");
    return B200NN_ERR_INVALID;
  }
  b200nn_multi *m = new (std::nothrow) b200nn_multi();
  if (!m) return B200NN_ERR_NOMEM;
  m->sh.resize((size_t)n_devices);
  for (int g = 0; g < n_devices; ++g) {
    const int dev = devices ? devices[g] : g;
    m->sh[g].dev = dev;
    int rc = b200nn_ctx_create(dev, &m->sh[g].ctx);
    if (rc != B200NN_OK) {
      b200nn_multi_destroy(m);
      return rc;
    }
  }
  // direct peer access where the hardware offers it (NVLink / NVSwitch); without it the peer
  // copies are staged by the driver and still correct
  m->direct = getenv("B200NN_MULTI_NO_DIRECT") == nullptr;
  for (int a = 0; a < n_devices; ++a)
    for (int b = 0; b < n_devices; ++b) {
      const int da = m->sh[a].dev, db = m->sh[b].dev;
      if (da == db) continue;
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, da, db) == cudaSuccess && can) {
        cudaSetDevice(da);
        cudaError_t e = cudaDeviceEnablePeerAccess(db, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) m->direct = false;
        (void)cudaGetLastError();
      } else {
        (void)cudaGetLastError();
        m->direct = false;
      }
    }
  *out = m;
  return B200NN_OK;
}

void b200nn_multi_destroy(b200nn_multi *m) {
  if (!m) return;
  for (Shard &s : m->sh)
    if (s.ctx) b200nn_ctx_destroy(s.ctx);
  delete m;
}

int32_t b200nn_multi_size(const b200nn_multi *m) { return m ? (int32_t)m->sh.size() : 0; }

int b200nn_multi_knn(b200nn_multi *m, const double *ref, int64_t nr, const double *qry, int64_t nq, int32_t d,
                     int32_t k, int32_t *idx_out, double *dist_out, const b200nn_opts *opts,
                     b200nn_stats *stats) {
  if (!m) {
    b200nn_set_error("NULL multi-device context");
    return B200NN_ERR_INVALID;
  }
  if (stats) memset(stats, 0, sizeof(*stats));
  m->pending = m->pending_snn = false;
  if (!ref || nr <= 0 || d <= 0 || k <= 0 || (!idx_out && !dist_out) || (qry && nq <= 0)) {
    b200nn_set_error("knn: invalid argument");
    return B200NN_ERR_INVALID;
  }
  if ((int64_t)k > nr) {
    b200nn_set_error("Cannot find more nearest neighbours than there are points");
    return B200NN_ERR_K_TOO_LARGE;
  }
  Job J;
  J.ref = ref;
  J.qry = qry;
  J.nr = nr;
  J.nq = qry ? nq : nr;
  J.d = d;
  J.k = k;
  J.idx_out = idx_out;
  J.dist_out = dist_out;
  J.o = opts ? *opts : b200nn_default_opts();
  set_rows(m, J.nq);
  return run_count(m, J, stats);
}

int b200nn_multi_find_neighbors_count(b200nn_multi *m, const double *data, int64_t n, int32_t d, int32_t k,
                                      double prune, int32_t compute_snn, int32_t *idx_out, double *dist_out,
                                      int32_t *nn_p_out, int32_t *nn_i_out, double *nn_x_out,
                                      const b200nn_opts *opts, int64_t *nnz_nn_out, int64_t *nnz_snn_out,
                                      b200nn_stats *stats) {
  if (!m) {
    b200nn_set_error("NULL multi-device context");
    return B200NN_ERR_INVALID;
  }
  if (stats) memset(stats, 0, sizeof(*stats));
  m->pending = m->pending_snn = false;
  if (!data || n <= 0 || d <= 0 || k <= 0) {
    b200nn_set_error("find_neighbors: invalid argument");
    return B200NN_ERR_INVALID;
  }
  if ((int64_t)k > n) {
    b200nn_set_error("Cannot find more nearest neighbours than there are points");
    return B200NN_ERR_K_TOO_LARGE;
  }
  Job J;
  J.ref = data;
  J.nr = J.nq = n;
  J.d = d;
  J.k = k;
  J.prune = prune;
  J.graphs = true;
  J.compute_snn = compute_snn != 0;
  J.idx_out = idx_out;
  J.dist_out = dist_out;
  J.early.p = nn_p_out;
  J.early.i = nn_i_out;
  J.early.x = nn_x_out;
  J.o = opts ? *opts : b200nn_default_opts();
  set_rows(m, n);
  RC_TRY(run_count(m, J, stats));
  int64_t nnz_snn = 0;
  for (const Shard &s : m->sh) nnz_snn += s.nnz_snn;
  if (nnz_snn >= 0x7fffffffLL) {
    b200nn_set_error("SNN graph has %lld entries, more than a dgCMatrix can hold", (long long)nnz_snn);
    return B200NN_ERR_NNZ_OVERFLOW;
  }
  if (nnz_nn_out) *nnz_nn_out = m->sh[0].nnz_nn;
  if (nnz_snn_out) *nnz_snn_out = nnz_snn;
  m->pending = true;
  m->pending_snn = compute_snn != 0;
  m->n = n;
  m->k = k;
  return B200NN_OK;
}

int b200nn_multi_find_neighbors_fill(b200nn_multi *m, int32_t *nn_p, int32_t *nn_i, double *nn_x,
                                     int32_t *snn_p, int32_t *snn_i, double *snn_x) {
  if (!m || !m->pending) {
    b200nn_set_error("b200nn_multi_find_neighbors_fill without a successful _count");
    return B200NN_ERR_STATE;
  }
  if ((snn_p || snn_i || snn_x) && !m->pending_snn) {
    b200nn_set_error("SNN was not computed by the last b200nn_multi_find_neighbors_count");
    return B200NN_ERR_STATE;
  }
  return run_shards(m, [&](int g) { shard_fill_phase(m, g, nn_p, nn_i, nn_x, snn_p, snn_i, snn_x); });
}

}  // extern "C"
