// modopt.cu -- modularity optimiser on the device: CUDA backend of modopt_core.inl and
// the C entry points b200nn_modularity_cluster*.
//
// Replaces RunModularityClusteringCpp (/root/reference/src/RModularityOptimizer.cpp:21-173) with
// results identical to the reference for the same seed; see modopt_core.inl for how the sequential
// algorithm is executed in parallel.  Here: the kernels behind the backend's parallel steps
// (k_pfor: one logical thread per element, grid-stride; CUB radix sort / scan for the grouping
// steps; k_seq_sum: the order-fixed running sums), a stream-ordered memory pool so that the many
// short-lived buffers cost no cudaMalloc, and argument checks with the wrapper's messages.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <errno.h>
#include <stdio.h>
#include <string.h>
#include <time.h>

#include <string>
#include <utility>
#include <vector>

#include "internal.cuh"

#define MO_HD __device__
template <typename T>
__device__ __forceinline__ void be_atomic_min(T *a, T v) { atomicMin(a, v); }
template <typename T>
__device__ __forceinline__ void be_atomic_max(T *a, T v) { atomicMax(a, v); }
__device__ __forceinline__ int be_atomic_add(int *a, int v) { return atomicAdd(a, v); }
__device__ __forceinline__ void be_atomic_max64(int64_t *a, int64_t v) {
  atomicMax(reinterpret_cast<long long *>(a), (long long)v);
}

#include "modopt_core.inl"

namespace {

template <class F>
__global__ void __launch_bounds__(256) k_pfor(int64_t n, F f) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) f(i);
}

template <class F>
__global__ void k_single(F f) {
  f();
}

// running sum in the order of the array (the reference's std::accumulate / += loops): one warp,
// coalesced loads, the additions themselves strictly one after the other
__global__ void __launch_bounds__(32) k_seq_sum(const double *__restrict__ x, int64_t n, double init, double *out) {
  const int lane = threadIdx.x;
  double s = init;
  double nxt = lane < n ? x[lane] : 0.0;
  for (int64_t base = 0; base < n; base += 32) {
    const double v = nxt;
    const int64_t nb = base + 32 + lane;
    nxt = nb < n ? x[nb] : 0.0;
    const int cnt = (int)(n - base < 32 ? n - base : 32);
    if (cnt == 32) {
#pragma unroll
      for (int l = 0; l < 32; ++l) s = __dadd_rn(s, __shfl_sync(0xffffffffu, v, l));
    } else {
      for (int l = 0; l < cnt; ++l) s = __dadd_rn(s, __shfl_sync(0xffffffffu, v, l));
    }
  }
  if (lane == 0) *out = s;
}

// The long chains of a local-moving round, one warp per chain (modopt_core.inl chain_walk is the
// definition): 32 events are loaded coalesced, the running weight takes them one after the other
// (every lane keeps the same running value; a leave is the addition of the negated weight, which is
// the same IEEE operation as the subtraction), each lane keeps the value after its own event.
__global__ void __launch_bounds__(256)
k_long_chains(const uint64_t *__restrict__ ks, double *__restrict__ ev_cw, int32_t *__restrict__ ev_nn,
              const int32_t *__restrict__ hend, const double *__restrict__ cw, const int32_t *__restrict__ nn,
              const int32_t *__restrict__ list, const int32_t *__restrict__ count) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int n_long = *count;
  for (int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < n_long; q += warps) {
    const int s = list[q], e = hend[s];
    const int32_t c = (int32_t)(ks[s] >> 32);
    double w = cw[c];
    int cnt = nn[c];
    for (int base = s; base < e; base += 32) {
      const int p = base + lane;
      const bool in = p < e;
      const bool join = in && ((uint32_t)ks[p] & 1u);
      const double x = in ? (join ? ev_cw[p] : -ev_cw[p]) : 0.0;
      // node count after my event: inclusive prefix of +1 / -1
      int d = in ? (join ? 1 : -1) : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, d, o);
        if (lane >= o) d += v;
      }
      const int m = e - base < 32 ? e - base : 32;
      double mine = 0.0;
      if (m == 32) {
#pragma unroll
        for (int l = 0; l < 32; ++l) {
          w = __dadd_rn(w, __shfl_sync(0xffffffffu, x, l));
          if (l == lane) mine = w;
        }
      } else {
        for (int l = 0; l < m; ++l) {
          w = __dadd_rn(w, __shfl_sync(0xffffffffu, x, l));
          if (l == lane) mine = w;
        }
      }
      if (in) {
        ev_cw[p] = mine;
        ev_nn[p] = cnt + d;
      }
      cnt += __shfl_sync(0xffffffffu, d, 31);
    }
  }
}

struct CudaBackend {
  b200nn_ctx *ctx;
  cudaStream_t st;
  int rc = B200NN_OK;  // first failure; later steps become no-ops
  int64_t win0 = 8192, winmax = (int64_t)1 << 40;
  double *d_scalar = nullptr;

  explicit CudaBackend(b200nn_ctx *c) : ctx(c), st(c->stream) {
    if (const char *e = getenv("B200NN_MODOPT_WIN0")) win0 = atoll(e);
    if (const char *e = getenv("B200NN_MODOPT_WINMAX")) winmax = atoll(e);
    trace = getenv("B200NN_MODOPT_TRACE") != nullptr;
    // keep the stream-ordered pool's memory between the many short-lived buffers of a call
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, c->device) == cudaSuccess) {
      uint64_t keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    (void)cudaGetLastError();
  }
  void fail(cudaError_t e, const char *what) {
    if (rc != B200NN_OK) return;
    b200nn_set_error("modularity optimiser: CUDA error %s in %s: %s", cudaGetErrorName(e), what, cudaGetErrorString(e));
    (void)cudaGetLastError();
    rc = e == cudaErrorMemoryAllocation ? B200NN_ERR_NOMEM : B200NN_ERR_CUDA;
  }
  template <typename T>
  T *alloc(int64_t n) {
    void *p = nullptr;
    if (rc != B200NN_OK) return nullptr;
    cudaError_t e = cudaMallocAsync(&p, sizeof(T) * (size_t)(n > 0 ? n : 1), st);
    if (e != cudaSuccess) {
      fail(e, "cudaMallocAsync");
      return nullptr;
    }
    return static_cast<T *>(p);
  }
  void free(void *p) {
    if (p) cudaFreeAsync(p, st);
  }
  template <class F>
  void pfor(int64_t n, F f) {
    if (rc != B200NN_OK || n <= 0) return;
    const int64_t blocks = std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 16);
    k_pfor<<<(unsigned)blocks, 256, 0, st>>>(n, f);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) fail(e, "k_pfor launch");
  }
  template <class F>
  void single(F f) {
    if (rc != B200NN_OK) return;
    k_single<<<1, 1, 0, st>>>(f);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) fail(e, "k_single launch");
  }
  void sort_keys(const uint64_t *in, uint64_t *out, int64_t n, int bits) {
    if (rc != B200NN_OK || n <= 0) return;
    size_t tb = 0;
    cudaError_t e = cub::DeviceRadixSort::SortKeys(nullptr, tb, in, out, n, 0, bits, st);
    if (e != cudaSuccess) return fail(e, "SortKeys (size)");
    void *tmp = alloc<uint8_t>((int64_t)tb);
    if (!tmp) return;
    e = cub::DeviceRadixSort::SortKeys(tmp, tb, in, out, n, 0, bits, st);
    ctx->launches += 3;
    if (e != cudaSuccess) fail(e, "SortKeys");
    free(tmp);
  }
  void sort_pairs(const uint64_t *kin, uint64_t *kout, const uint32_t *vin, uint32_t *vout, int64_t n, int bits) {
    if (rc != B200NN_OK || n <= 0) return;
    size_t tb = 0;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, tb, kin, kout, vin, vout, n, 0, bits, st);
    if (e != cudaSuccess) return fail(e, "SortPairs (size)");
    void *tmp = alloc<uint8_t>((int64_t)tb);
    if (!tmp) return;
    e = cub::DeviceRadixSort::SortPairs(tmp, tb, kin, kout, vin, vout, n, 0, bits, st);
    ctx->launches += 3;
    if (e != cudaSuccess) fail(e, "SortPairs");
    free(tmp);
  }
  // out[0..n]: exclusive prefix sums of in[0..n-1]; out[n] = the total (in[] has n + 1 slots)
  void excl_scan(const int64_t *in, int64_t *out, int64_t n) {
    if (rc != B200NN_OK) return;
    size_t tb = 0;
    cudaError_t e = cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, n + 1, st);
    if (e != cudaSuccess) return fail(e, "ExclusiveSum (size)");
    void *tmp = alloc<uint8_t>((int64_t)tb);
    if (!tmp) return;
    e = cub::DeviceScan::ExclusiveSum(tmp, tb, in, out, n + 1, st);
    ctx->launches += 2;
    if (e != cudaSuccess) fail(e, "ExclusiveSum");
    free(tmp);
  }
  void h2d(void *d, const void *s, size_t b) {
    if (rc != B200NN_OK || !d) return;
    cudaError_t e = cudaMemcpyAsync(d, s, b, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // the host buffer may be a temporary
    if (e != cudaSuccess) fail(e, "h2d");
  }
  void d2h(void *d, const void *s, size_t b) {
    if (rc != B200NN_OK || !s) {
      memset(d, 0, b);
      return;
    }
    cudaError_t e = cudaMemcpyAsync(d, s, b, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      fail(e, "d2h");
      memset(d, 0, b);
    }
  }
  void d2d(void *d, const void *s, size_t b) {
    if (rc != B200NN_OK || !d || !s) return;
    cudaError_t e = cudaMemcpyAsync(d, s, b, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) fail(e, "d2d");
  }
  double seq_sum_plain(const double *x, int64_t n, double init) {
    if (rc != B200NN_OK) return 0.0;
    if (!d_scalar) d_scalar = alloc<double>(1);
    if (!d_scalar) return 0.0;
    k_seq_sum<<<1, 32, 0, st>>>(x, n, init, d_scalar);
    ctx->launches++;
    double h = 0.0;
    d2h(&h, d_scalar, sizeof(double));
    return h;
  }
  bool failed() const { return rc != B200NN_OK; }
  void long_chains(const uint64_t *ks, double *ev_cw, int32_t *ev_nn, const int32_t *hend, const double *cw,
                   const int32_t *nn, const int32_t *list, const int32_t *count) {
    if (rc != B200NN_OK) return;
    k_long_chains<<<ctx->sm_count * 2, 256, 0, st>>>(ks, ev_cw, ev_nn, hend, cw, nn, list, count);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) fail(e, "k_long_chains launch");
  }
  // B200NN_MODOPT_TRACE=1: time between marks (stream synchronised), summed per name, printed at the end
  bool trace = false;
  double t_mark = 0;
  std::vector<std::pair<std::string, double>> acc;
  void mark(const char *name) {
    if (!trace) return;
    const double t = now_ms();
    for (auto &a : acc)
      if (a.first == name) {
        a.second += t - t_mark;
        t_mark = now_ms();
        return;
      }
    acc.push_back({name, t - t_mark});
    t_mark = now_ms();
  }
  void note_window(int64_t n, int64_t w, int rounds, bool converged, int64_t max_chain) {
    if (trace) fprintf(stderr, "[b200nn modopt] network %lld nodes: window %lld, %d rounds, longest chain %lld%s\n",
                       (long long)n, (long long)w, rounds, (long long)max_chain,
                       converged ? "" : " (not converged, retried smaller)");
  }
  void report() {
    if (!trace) return;
    for (auto &a : acc) fprintf(stderr, "[b200nn modopt] %-22s %10.2f ms\n", a.first.c_str(), a.second);
  }
  int64_t initial_window(int64_t) { return win0; }
  int64_t max_window(int64_t) { return winmax; }
  double now_ms() {
    cudaStreamSynchronize(st);
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
  }
};

int check_args(int modularity_function, double resolution, int algorithm, int n_random_starts, int n_iterations) {
  // the wrapper's stop() messages, RModularityOptimizer.cpp:32-41
  if (modularity_function != 1 && modularity_function != 2) {
    b200nn_set_error("Modularity parameter must be equal to 1 or 2.");
    return B200NN_ERR_INVALID;
  }
  if (algorithm != 1 && algorithm != 2 && algorithm != 3 && algorithm != 4) {
    b200nn_set_error("Algorithm for modularity optimization must be 1, 2, 3, or 4");
    return B200NN_ERR_INVALID;
  }
  if (n_random_starts < 1) {
    b200nn_set_error("Have to have at least one start");
    return B200NN_ERR_INVALID;
  }
  if (n_iterations < 1) {
    b200nn_set_error("Need at least one interation");
    return B200NN_ERR_INVALID;
  }
  if (modularity_function == 2 && resolution > 1.0) {
    b200nn_set_error("error: resolution<1 for alternative modularity");
    return B200NN_ERR_INVALID;
  }
  if (algorithm == 4) {
    // FindClusters sends algorithm 4 to RunLeiden (R/clustering.R:364-375, an R-side package), never
    // to this entry point; the wrapper itself would run nothing for it
    b200nn_set_error("modularity optimiser: algorithm 4 (Leiden) is not part of RunModularityClusteringCpp; "
                     "1 = Louvain, 2 = Louvain with multilevel refinement, 3 = smart local moving");
    return B200NN_ERR_INVALID;
  }
  return B200NN_OK;
}

void fill_stats(const modopt::Stats &s, int64_t n, int64_t edges, int32_t n_clusters, double q, double ms_total,
                int64_t launches, b200nn_modopt_stats *out) {
  if (!out) return;
  memset(out, 0, sizeof(*out));
  out->n_nodes = n;
  out->n_edges = edges;
  out->n_clusters = n_clusters;
  out->modularity = q;
  out->windows = s.windows;
  out->rounds = s.rounds;
  out->sequential_steps = s.seq_steps;
  out->window_retries = s.shrinks;
  out->local_moving_calls = s.local_moving_calls;
  out->levels = s.levels;
  out->max_window = s.max_window;
  out->ms_total = ms_total;
  out->ms_build = s.ms_build;
  out->ms_local_moving = s.ms_local_moving;
  out->ms_reduce = s.ms_reduce;
  out->ms_quality = s.ms_quality;
  out->ms_subnetworks = s.ms_subnetworks;
  out->gpu_launches = launches;
}

// stream-ordered device buffer released on every exit path of an entry point
struct PoolBuf {
  void *p = nullptr;
  cudaStream_t st;
  explicit PoolBuf(cudaStream_t s) : st(s) {}
  ~PoolBuf() {
    if (p) cudaFreeAsync(p, st);
  }
  PoolBuf(const PoolBuf &) = delete;
  PoolBuf &operator=(const PoolBuf &) = delete;
  template <typename T>
  T *as() const { return static_cast<T *>(p); }
};

template <typename PT>
int cluster_device_csc(b200nn_ctx *ctx, int64_t n, const PT *d_p, const int32_t *d_i, const double *d_x,
                       int modularity_function, double resolution, int algorithm, int n_random_starts,
                       int n_iterations, uint64_t seed, int32_t *cluster_out, b200nn_modopt_stats *stats) {
  CudaBackend be(ctx);
  const int64_t l0 = ctx->launches;
  const double t0 = be.now_ms();
  modopt::Net net;
  modopt::Stats st;
  const int64_t E = modopt::build_network(be, (int32_t)n, d_p, d_i, d_x, modularity_function, net);
  st.ms_build = be.now_ms() - t0;
  if (be.rc != B200NN_OK) return be.rc;
  if (E == 0) {
    b200nn_set_error("Matrix contained no network data.  Check format.");
    return B200NN_ERR_INVALID;
  }
  double q = 0;
  const int32_t k = modopt::run_clustering(be, net, algorithm, resolution, modularity_function, n_random_starts,
                                           n_iterations, seed, cluster_out, &q, &st);
  modopt::free_net(be, net);
  be.free(be.d_scalar);
  be.report();
  const double t1 = be.now_ms();
  if (be.rc != B200NN_OK) return be.rc;
  if (k < 0) {
    b200nn_set_error("Clustering step failed.");
    return B200NN_ERR_STATE;
  }
  fill_stats(st, n, E, k, q, t1 - t0, ctx->launches - l0, stats);
  return B200NN_OK;
}

// the same on an edge list in device memory (the edge.file.name input)
int cluster_device_edges(b200nn_ctx *ctx, int64_t n, const int32_t *d_n1, const int32_t *d_n2, const double *d_w,
                         int64_t m, int modularity_function, double resolution, int algorithm, int n_random_starts,
                         int n_iterations, uint64_t seed, int32_t *cluster_out, b200nn_modopt_stats *stats) {
  CudaBackend be(ctx);
  const int64_t l0 = ctx->launches;
  const double t0 = be.now_ms();
  modopt::Net net;
  modopt::Stats st;
  const int64_t E = modopt::build_network_edges(be, (int32_t)n, d_n1, d_n2, d_w, m, modularity_function, net);
  st.ms_build = be.now_ms() - t0;
  if (be.rc != B200NN_OK) return be.rc;
  if (E == 0) {
    b200nn_set_error("Matrix contained no network data.  Check format.");
    return B200NN_ERR_INVALID;
  }
  double q = 0;
  const int32_t k = modopt::run_clustering(be, net, algorithm, resolution, modularity_function, n_random_starts,
                                           n_iterations, seed, cluster_out, &q, &st);
  modopt::free_net(be, net);
  be.free(be.d_scalar);
  be.report();
  const double t1 = be.now_ms();
  if (be.rc != B200NN_OK) return be.rc;
  if (k < 0) {
    b200nn_set_error("Clustering step failed.");
    return B200NN_ERR_STATE;
  }
  fill_stats(st, n, E, k, q, t1 - t0, ctx->launches - l0, stats);
  return B200NN_OK;
}

// readInputFile (/root/reference/src/ModularityOptimizer.cpp:806-838): one edge per line,
// "node1<TAB>node2[<TAB>weight]" (weight 1 when absent), node ids 0-based; the number of nodes is the
// largest id + 1.  std::stoi / std::stod semantics: leading white space skipped, trailing text ignored,
// no digits -> failure (the wrapper turns every failure into "Could not parse edge file.").
bool parse_edge_file(const char *fname, std::vector<int32_t> &n1, std::vector<int32_t> &n2, std::vector<double> &w) {
  FILE *f = fopen(fname, "r");
  if (!f) return false;
  std::string line;
  char buf[4096];
  bool ok = true;
  auto take_line = [&]() -> bool {  // std::getline: a last line without '\n' counts, an empty tail does not
    line.clear();
    bool any = false;
    while (fgets(buf, sizeof(buf), f)) {
      any = true;
      const size_t len = strlen(buf);
      if (len && buf[len - 1] == '\n') {
        line.append(buf, len - 1);
        return true;
      }
      line.append(buf, len);
    }
    return any;
  };
  while (ok && take_line()) {
    const char *p = line.c_str();
    const char *t1 = strchr(p, '\t');
    if (!t1) {  // splittedLine[1] does not exist: the reference reads out of bounds; refuse
      ok = false;
      break;
    }
    const char *t2 = strchr(t1 + 1, '\t');
    char *end = nullptr;
    errno = 0;
    const long a = strtol(p, &end, 10);
    if (end == p || end > t1 || errno == ERANGE || a < 0 || a > 0x7ffffffeL) ok = false;
    const long b = strtol(t1 + 1, &end, 10);
    if (end == t1 + 1 || (t2 && end > t2) || errno == ERANGE || b < 0 || b > 0x7ffffffeL) ok = false;
    double wt = 1.0;
    if (t2 && t2[1] != '\0') {  // split() drops an empty last token
      wt = strtod(t2 + 1, &end);
      if (end == t2 + 1) ok = false;
    }
    n1.push_back((int32_t)a);
    n2.push_back((int32_t)b);
    w.push_back(wt);
  }
  fclose(f);
  return ok;
}

}  // namespace

extern "C" {

int b200nn_modularity_cluster(b200nn_ctx *ctx, int64_t n, const int32_t *p, const int32_t *i, const double *x,
                              int32_t modularity_function, double resolution, int32_t algorithm,
                              int32_t n_random_starts, int32_t n_iterations, uint64_t random_seed,
                              int32_t *cluster_out, b200nn_modopt_stats *stats) {
  if (!ctx) {
    b200nn_set_error("NULL context");
    return B200NN_ERR_INVALID;
  }
  if (!p || !i || !x || !cluster_out || n <= 0 || n >= 0x7fffffffLL) {
    b200nn_set_error("modularity_cluster: invalid argument");
    return B200NN_ERR_INVALID;
  }
  RC_TRY(check_args(modularity_function, resolution, algorithm, n_random_starts, n_iterations));
  const int64_t nnz = p[n];
  if (p[0] != 0 || nnz < 0) {
    b200nn_set_error("modularity_cluster: p is not a dgCMatrix column pointer");
    return B200NN_ERR_INVALID;
  }
  CU_TRY(cudaSetDevice(ctx->device));
  int rc;
  {
    PoolBuf bp(ctx->stream), bi(ctx->stream), bx(ctx->stream);  // freed on every path, error returns included
    CU_TRY(cudaMallocAsync(&bp.p, sizeof(int32_t) * (size_t)(n + 1), ctx->stream));
    CU_TRY(cudaMallocAsync(&bi.p, sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1), ctx->stream));
    CU_TRY(cudaMallocAsync(&bx.p, sizeof(double) * (size_t)(nnz > 0 ? nnz : 1), ctx->stream));
    CU_TRY(cudaMemcpyAsync(bp.p, p, sizeof(int32_t) * (size_t)(n + 1), cudaMemcpyHostToDevice, ctx->stream));
    if (nnz > 0) {
      CU_TRY(cudaMemcpyAsync(bi.p, i, sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice, ctx->stream));
      CU_TRY(cudaMemcpyAsync(bx.p, x, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, ctx->stream));
    }
    rc = cluster_device_csc(ctx, n, bp.as<int32_t>(), bi.as<int32_t>(), bx.as<double>(), modularity_function,
                            resolution, algorithm, n_random_starts, n_iterations, random_seed, cluster_out, stats);
  }
  cudaStreamSynchronize(ctx->stream);
  return rc;
}

int b200nn_modularity_cluster_edges(b200nn_ctx *ctx, int64_t n_nodes, const int32_t *node1, const int32_t *node2,
                                    const double *weight, int64_t n_edges, int32_t modularity_function,
                                    double resolution, int32_t algorithm, int32_t n_random_starts,
                                    int32_t n_iterations, uint64_t random_seed, int32_t *cluster_out,
                                    b200nn_modopt_stats *stats) {
  if (!ctx) {
    b200nn_set_error("NULL context");
    return B200NN_ERR_INVALID;
  }
  if (!node1 || !node2 || !cluster_out || n_nodes <= 0 || n_nodes >= 0x7fffffffLL || n_edges < 0 ||
      n_edges >= 0xffffffffLL) {
    b200nn_set_error("modularity_cluster_edges: invalid argument");
    return B200NN_ERR_INVALID;
  }
  RC_TRY(check_args(modularity_function, resolution, algorithm, n_random_starts, n_iterations));
  for (int64_t e = 0; e < n_edges; ++e)
    if (node1[e] < 0 || node2[e] < 0 || node1[e] >= n_nodes || node2[e] >= n_nodes) {
      b200nn_set_error("modularity_cluster_edges: edge %lld names a node outside [0, %lld)", (long long)e,
                       (long long)n_nodes);
      return B200NN_ERR_INVALID;
    }
  if (n_edges == 0) {
    b200nn_set_error("Matrix contained no network data.  Check format.");
    return B200NN_ERR_INVALID;
  }
  CU_TRY(cudaSetDevice(ctx->device));
  int rc;
  {
    PoolBuf b1(ctx->stream), b2(ctx->stream), bw(ctx->stream);
    CU_TRY(cudaMallocAsync(&b1.p, sizeof(int32_t) * (size_t)n_edges, ctx->stream));
    CU_TRY(cudaMallocAsync(&b2.p, sizeof(int32_t) * (size_t)n_edges, ctx->stream));
    CU_TRY(cudaMallocAsync(&bw.p, sizeof(double) * (size_t)n_edges, ctx->stream));
    CU_TRY(cudaMemcpyAsync(b1.p, node1, sizeof(int32_t) * (size_t)n_edges, cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(cudaMemcpyAsync(b2.p, node2, sizeof(int32_t) * (size_t)n_edges, cudaMemcpyHostToDevice, ctx->stream));
    if (weight) {
      CU_TRY(cudaMemcpyAsync(bw.p, weight, sizeof(double) * (size_t)n_edges, cudaMemcpyHostToDevice, ctx->stream));
    } else {  // unweighted list: every edge weighs 1 (readInputFile's default)
      std::vector<double> ones((size_t)n_edges, 1.0);
      CU_TRY(cudaMemcpyAsync(bw.p, ones.data(), sizeof(double) * (size_t)n_edges, cudaMemcpyHostToDevice, ctx->stream));
      CU_TRY(cudaStreamSynchronize(ctx->stream));
    }
    rc = cluster_device_edges(ctx, n_nodes, b1.as<int32_t>(), b2.as<int32_t>(), bw.as<double>(), n_edges,
                              modularity_function, resolution, algorithm, n_random_starts, n_iterations, random_seed,
                              cluster_out, stats);
  }
  cudaStreamSynchronize(ctx->stream);
  return rc;
}

int b200nn_modularity_cluster_edge_file(b200nn_ctx *ctx, const char *filename, int32_t modularity_function,
                                        double resolution, int32_t algorithm, int32_t n_random_starts,
                                        int32_t n_iterations, uint64_t random_seed, int32_t *cluster_out,
                                        int64_t cluster_capacity, int64_t *n_nodes_out, b200nn_modopt_stats *stats) {
  if (!ctx) {
    b200nn_set_error("NULL context");
    return B200NN_ERR_INVALID;
  }
  if (!filename || !n_nodes_out) {
    b200nn_set_error("modularity_cluster_edge_file: invalid argument");
    return B200NN_ERR_INVALID;
  }
  *n_nodes_out = 0;
  RC_TRY(check_args(modularity_function, resolution, algorithm, n_random_starts, n_iterations));
  std::vector<int32_t> n1, n2;
  std::vector<double> w;
  if (!parse_edge_file(filename, n1, n2, w) || n1.empty()) {
    b200nn_set_error("Could not parse edge file.");
    return B200NN_ERR_INVALID;
  }
  int32_t mx = 0;
  for (size_t e = 0; e < n1.size(); ++e) mx = std::max(mx, std::max(n1[e], n2[e]));
  const int64_t n = (int64_t)mx + 1;
  *n_nodes_out = n;
  if (!cluster_out || cluster_capacity < n) {  // size query: the caller allocates n ids and calls again
    if (cluster_out) {
      b200nn_set_error("modularity_cluster_edge_file: the file names %lld nodes, cluster_out holds %lld",
                       (long long)n, (long long)cluster_capacity);
      return B200NN_ERR_INVALID;
    }
    return B200NN_OK;
  }
  return b200nn_modularity_cluster_edges(ctx, n, n1.data(), n2.data(), w.data(), (int64_t)n1.size(),
                                         modularity_function, resolution, algorithm, n_random_starts, n_iterations,
                                         random_seed, cluster_out, stats);
}

int b200nn_modularity_cluster_resident(b200nn_ctx *ctx, int32_t modularity_function, double resolution,
                                       int32_t algorithm, int32_t n_random_starts, int32_t n_iterations,
                                       uint64_t random_seed, int32_t *cluster_out, int64_t n_expected,
                                       b200nn_modopt_stats *stats) {
  if (!ctx) {
    b200nn_set_error("NULL context");
    return B200NN_ERR_INVALID;
  }
  RC_TRY(check_args(modularity_function, resolution, algorithm, n_random_starts, n_iterations));
  const CsrResult &r = ctx->snn;
  if (!r.valid || !r.p || !cluster_out) {
    b200nn_set_error("modularity_cluster_resident: no SNN graph is resident (run b200nn_find_neighbors_count / "
                     "b200nn_dev_graphs with compute_snn first)");
    return B200NN_ERR_STATE;
  }
  if (n_expected > 0 && r.n_rows != n_expected) {
    b200nn_set_error("modularity_cluster_resident: the resident SNN holds %lld rows (a shard?), expected %lld",
                     (long long)r.n_rows, (long long)n_expected);
    return B200NN_ERR_STATE;
  }
  CU_TRY(cudaSetDevice(ctx->device));
  // S is symmetric, so its CSR arrays are its CSC arrays (semantics item 5)
  return cluster_device_csc(ctx, r.n_rows, r.p, r.i, r.x, modularity_function, resolution, algorithm,
                            n_random_starts, n_iterations, random_seed, cluster_out, stats);
}

}  // extern "C"
