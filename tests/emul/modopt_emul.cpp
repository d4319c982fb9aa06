// modopt_emul.cpp -- serial emulation backend for seurat_b200/csrc/modopt_core.inl.
//
// TEST INFRASTRUCTURE: runs the very step functions the CUDA library launches as kernels, one
// "thread" after the other, so that the windowed fixed-point execution of the modularity optimiser
// can be compared with the real reference (oracle/_ref/libmodopt_ref.so) on a box without a GPU.
// Never linked into libb200nn.so; the product path has no CPU fallback.
//
//   g++ -O2 -std=c++17 -ffp-contract=off -shared -fPIC tests/emul/modopt_emul.cpp -o tests/emul/libmodopt_emul.so
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <vector>

#define MO_HD
template <typename T> static inline void be_atomic_min(T *a, T v) { if (v < *a) *a = v; }
template <typename T> static inline void be_atomic_max(T *a, T v) { if (v > *a) *a = v; }
static inline void be_atomic_max64(int64_t *a, int64_t v) { if (v > *a) *a = v; }
static inline int be_atomic_add(int *a, int v) { const int o = *a; *a += v; return o; }

#include "../../seurat_b200/csrc/modopt_core.inl"

namespace {
struct EmulBackend {
  int64_t win0 = 1024, winmax = 1 << 20;
  template <typename T> T *alloc(int64_t n) { return (T *)malloc(sizeof(T) * (size_t)(n > 0 ? n : 1)); }
  void free(void *p) { ::free(p); }
  template <class F> void pfor(int64_t n, F f) { for (int64_t i = 0; i < n; ++i) f(i); }
  template <class F> void single(F f) { f(); }
  void sort_keys(const uint64_t *in, uint64_t *out, int64_t n, int) {
    memcpy(out, in, sizeof(uint64_t) * (size_t)n);
    std::sort(out, out + n);
  }
  void sort_pairs(const uint64_t *kin, uint64_t *kout, const uint32_t *vin, uint32_t *vout, int64_t n, int) {
    std::vector<int64_t> ix((size_t)n);
    for (int64_t i = 0; i < n; ++i) ix[i] = i;
    std::stable_sort(ix.begin(), ix.end(), [&](int64_t a, int64_t b) { return kin[a] < kin[b]; });
    for (int64_t i = 0; i < n; ++i) { kout[i] = kin[ix[i]]; vout[i] = vin[ix[i]]; }
  }
  void excl_scan(const int64_t *in, int64_t *out, int64_t n) {
    int64_t acc = 0;
    for (int64_t i = 0; i < n; ++i) { const int64_t v = in[i]; out[i] = acc; acc += v; }
    out[n] = acc;
  }
  void h2d(void *d, const void *s, size_t b) { memcpy(d, s, b); }
  void d2h(void *d, const void *s, size_t b) { memcpy(d, s, b); }
  void d2d(void *d, const void *s, size_t b) { memcpy(d, s, b); }
  double seq_sum_plain(const double *x, int64_t n, double init) {
    double s = init;
    for (int64_t i = 0; i < n; ++i) s = s + x[i];
    return s;
  }
  bool failed() const { return false; }
  void mark(const char *) {}
  void long_chains(const uint64_t *ks, double *ev_cw, int32_t *ev_nn, const int32_t *hend, const double *cw,
                   const int32_t *nn, const int32_t *list, const int32_t *count) {
    for (int32_t q = 0; q < *count; ++q) {
      const int32_t s = list[q];
      const int32_t c = (int32_t)(ks[s] >> 32);
      modopt::chain_walk(ks, ev_cw, ev_nn, s, hend[s], cw[c], nn[c]);
    }
  }
  void note_window(int64_t n, int64_t w, int r, bool, int64_t mc) { if (getenv("EMUL_TRACE")) fprintf(stderr, "net %lld window %lld rounds %d max_chain %lld\n", (long long)n, (long long)w, r, (long long)mc); }

  int64_t initial_window(int64_t) { return win0; }
  int64_t max_window(int64_t) { return winmax; }
  double now_ms() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
  }
};
}  // namespace

extern "C" int emul_modularity_cluster(int n, const int *p, const int *i, const double *x, int modularityFunction,
                                       double resolution, int algorithm, int nRandomStarts, int nIterations,
                                       uint64_t seed, int *cluster_out, double *modularity_out, int64_t win0,
                                       int64_t winmax, int64_t *stats_out) {
  using namespace modopt;
  EmulBackend be;
  if (win0 > 0) be.win0 = win0;
  if (winmax > 0) be.winmax = winmax;
  Net net;
  const i64 E = build_network(be, n, p, i, x, modularityFunction, net);
  if (E == 0) return -2;
  Stats st;
  const i32 k = run_clustering(be, net, algorithm, resolution, modularityFunction, nRandomStarts, nIterations, seed,
                               cluster_out, modularity_out, &st);
  if (stats_out) {
    stats_out[0] = st.windows; stats_out[1] = st.rounds; stats_out[2] = st.seq_steps; stats_out[3] = st.shrinks;
    stats_out[4] = st.local_moving_calls; stats_out[5] = st.levels; stats_out[6] = st.max_window;
    stats_out[7] = (int64_t)st.ms_local_moving; stats_out[8] = (int64_t)st.ms_reduce; stats_out[9] = (int64_t)st.ms_quality; stats_out[10] = (int64_t)st.ms_subnetworks;
  }
  free_net(be, net);
  return k;
}

extern "C" int emul_modularity_cluster_edges(int n, const int *node1, const int *node2, const double *w, int64_t m,
                                             int modularityFunction, double resolution, int algorithm,
                                             int nRandomStarts, int nIterations, uint64_t seed, int *cluster_out,
                                             double *modularity_out) {
  using namespace modopt;
  EmulBackend be;
  Net net;
  const i64 E = build_network_edges(be, n, node1, node2, w, m, modularityFunction, net);
  if (E == 0) return -2;
  Stats st;
  const i32 k = run_clustering(be, net, algorithm, resolution, modularityFunction, nRandomStarts, nIterations, seed,
                               cluster_out, modularity_out, &st);
  free_net(be, net);
  return k;
}

extern "C" double emul_seq_sum(const double *x, int64_t n, double init) {
  EmulBackend be;
  return modopt::seq_sum(be, x, n, init);
}
