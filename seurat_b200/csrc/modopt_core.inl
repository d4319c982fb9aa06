// modopt_core.inl -- the modularity optimiser as backend-neutral parallel steps.
//
// Replaces RunModularityClusteringCpp, /root/reference/src/RModularityOptimizer.cpp:21-173, and the
// classes it drives, /root/reference/src/ModularityOptimizer.cpp (Network, Clustering,
// VOSClusteringTechnique: local moving :474-581, Louvain :583-602, multilevel refinement :616-636,
// reduced network :315-367, quality function :451-472).  Results are IDENTICAL to the reference's
// for the same seed: the reference is a sequential randomised algorithm whose floating-point state
// (cluster weights updated by -= / += in visiting order) feeds every later decision, so this is
// not "a parallel Louvain" but a parallel EXECUTION of that very sequence:
//
//   * local moving runs over WINDOWS of consecutive positions of the node permutation.  All nodes
//     of a window decide at once from speculative inputs (the decisions of the window's earlier
//     nodes as of the previous round), every cluster's weight is replayed as one sequential chain
//     of its leave / join events in visiting order (the reference's exact additions), and the
//     round repeats until no decision changes.  Position t is final after at most t + 1 rounds
//     (induction over the visiting order), so the fixed point IS the sequential execution; in
//     practice 2-4 rounds.  Window sizes adapt; W = 1 is the reference step itself;
//   * sums whose order the reference fixes (node weights, cluster weights, reduced edges, the
//     quality function) are evaluated in that order -- in parallel ACROSS chains, sequentially
//     along each one;
//   * the Java LCG and the swap chain of the permutation (ModularityOptimizer.cpp:33-77) are host
//     control data (n draws), like the Jaccard table of the SNN stage.
//
// This file is included twice: by modopt.cu (nvcc; Backend = CUDA kernels + CUB sort/scan) and by
// tests/emul/modopt_emul.cpp (g++; Backend = plain loops).  The emulation exists so that the
// step logic can be checked against the real reference (oracle/_ref) on a box without a GPU; it
// is test infrastructure and is never linked into libb200nn.so.
//
// Every parallel step is a pfor whose iterations write disjoint locations (or use the integer
// atomics atomic_min / atomic_max / atomic_or), so the serial emulation is faithful.
#pragma once
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <vector>

namespace modopt {

typedef uint64_t u64;
typedef int64_t i64;
typedef int32_t i32;
typedef uint32_t u32;

// ---- reference arithmetic: no contraction into FMA, round to nearest -------------------------
MO_HD inline double dadd(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}
MO_HD inline double dsub(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dsub_rn(a, b);
#else
  return a - b;
#endif
}
MO_HD inline double dmul(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}

// ---- java.util.Random as the reference restates it (ModularityOptimizer.cpp:33-61) -----------
struct JavaRandom {
  u64 seed;
  explicit JavaRandom(u64 s) { seed = (s ^ 0x5DEECE66DULL) & ((1ULL << 48) - 1); }
  int next(int bits) {
    seed = (seed * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1);
    return (int)(seed >> (48 - bits));
  }
  int nextInt(int n) {
    if ((n & -n) == n) return (int)(((u64)n * (u64)next(31)) >> 31);
    int bits, val;
    do {
      bits = next(31);
      val = bits % n;
    } while (bits - val + (n - 1) < 0);
    return val;
  }
};

// ModularityOptimizer.cpp:64-77: n swaps with a uniformly drawn partner (not Fisher-Yates)
inline void random_permutation(JavaRandom &rng, int n, std::vector<i32> &perm) {
  perm.resize((size_t)n);
  for (int i = 0; i < n; ++i) perm[i] = i;
  for (int i = 0; i < n; ++i) {
    const int j = rng.nextInt(n);
    const i32 k = perm[i];
    perm[i] = perm[j];
    perm[j] = k;
  }
}

// ---- a network level (CSR adjacency, both directions stored) ----------------------------------
struct Net {
  i32 n = 0;
  i64 m2 = 0;              // stored adjacency entries (2 x edges)
  i64 maxdeg = 0;
  i64 *first = nullptr;    // [n + 1]
  i32 *nbr = nullptr;      // [m2]
  double *ew = nullptr;    // [m2]
  double *nw = nullptr;    // [n] node weights
  bool owned = false;
};

struct Stats {
  i64 windows = 0, rounds = 0, seq_steps = 0, shrinks = 0, local_moving_calls = 0, levels = 0;
  i64 max_window = 0;
  double ms_local_moving = 0, ms_reduce = 0, ms_quality = 0, ms_build = 0, ms_subnetworks = 0;
};

MO_HD inline u32 hash_cluster(i32 c) { return (u32)c * 0x9E3779B1u; }

// chains of at least this many events go to the backend's cooperative walker
constexpr i64 LONG_CHAIN = 96;

// Replays one cluster's events [s, e) of the sorted event list: ev_cw[] holds the weight of each
// event's node on entry and the cluster's weight right after the event on exit (a leave subtracts,
// a join adds: the reference's -= / += in visiting order); ev_nn[] the node count after the event.
MO_HD inline void chain_walk(const u64 *ks, double *ev_cw, i32 *ev_nn, i64 s, i64 e, double w, i32 cnt) {
  for (i64 q = s; q < e; ++q) {
    const double wj = ev_cw[q];
    if ((u32)ks[q] & 1u) {
      w = dadd(w, wj);
      ++cnt;
    } else {
      w = dsub(w, wj);
      --cnt;
    }
    ev_cw[q] = w;
    ev_nn[q] = cnt;
  }
}

// =================================================================================================
// Local moving (VOSClusteringTechnique::runLocalMovingAlgorithm, ModularityOptimizer.cpp:474-581)
// =================================================================================================
template <class B>
struct LocalMoving {
  B &be;
  const Net &net;
  double resolution;
  Stats *stats;

  // committed state
  i32 *cl;          // [n] cluster of every node (caller's array, updated in place)
  double *cw;       // [n] cluster weights
  i32 *nn;          // [n] nodes per cluster
  i32 *unused;      // [n] stack of unused cluster ids
  i32 *perm, *pos;  // [n] visiting order and its inverse
  // window scratch (capacity Wcap positions, Tcap table slots)
  i64 Wcap = 0, Tcap = 0;
  i32 *D = nullptr, *D2 = nullptr;   // decisions of the previous / this round
  i32 *c0 = nullptr;                 // cluster of node(t) when the window starts
  i32 *leave_nn = nullptr;           // nodes left in c0 right after node(t) left it
  i64 *toff = nullptr;               // [W + 1] table offsets
  i64 *tcap_in = nullptr;            // [W] table capacities (scan input)
  i32 *tcl = nullptr;                // [Tcap] table: cluster ids (-1 = empty)
  double *tsum = nullptr;            // [Tcap] table: edge weight towards the cluster
  u64 *keys = nullptr, *keys_s = nullptr;  // [2W] events (cluster << 32 | t << 1 | join)
  double *ev_cw = nullptr;           // [2W] cluster weight after the event
  i32 *ev_nn = nullptr;              // [2W] node count after the event
  i32 *hend = nullptr;               // [2W] at a segment head: end of the segment
  i32 *long_list = nullptr;          // [2W / LONG_CHAIN + 1] heads of the long chains of a round
  i32 *segfirst = nullptr;           // [n] first event of a cluster in keys_s, -1 = none this round
  i64 *pushflag = nullptr, *pushoff = nullptr;  // [W + 1]
  i32 *flags = nullptr;              // device: [0] changed, [1] first popper, [2] first unstable, [3] last unstable
  i32 nUnused = 0;

  LocalMoving(B &b, const Net &nt, double res, Stats *st) : be(b), net(nt), resolution(res), stats(st) {}

  void alloc_window(i64 W, i64 T) {
    if (W > Wcap) {
      free_window_w();
      Wcap = W;
      D = be.template alloc<i32>(W);
      D2 = be.template alloc<i32>(W);
      c0 = be.template alloc<i32>(W);
      leave_nn = be.template alloc<i32>(W);
      toff = be.template alloc<i64>(W + 1);
      tcap_in = be.template alloc<i64>(W + 1);
      keys = be.template alloc<u64>(2 * W);
      keys_s = be.template alloc<u64>(2 * W);
      ev_cw = be.template alloc<double>(2 * W);
      ev_nn = be.template alloc<i32>(2 * W);
      hend = be.template alloc<i32>(2 * W);
      long_list = be.template alloc<i32>(2 * W / LONG_CHAIN + 2);
      pushflag = be.template alloc<i64>(W + 1);
      pushoff = be.template alloc<i64>(W + 1);
    }
    if (T > Tcap) {
      be.free(tcl);
      be.free(tsum);
      Tcap = T + T / 4;
      tcl = be.template alloc<i32>(Tcap);
      tsum = be.template alloc<double>(Tcap);
    }
  }
  void free_window_w() {
    be.free(D); be.free(D2); be.free(c0); be.free(leave_nn); be.free(toff); be.free(tcap_in);
    be.free(keys); be.free(keys_s); be.free(ev_cw); be.free(ev_nn); be.free(hend); be.free(long_list);
    be.free(pushflag); be.free(pushoff);
    D = D2 = c0 = leave_nn = nullptr; toff = tcap_in = nullptr; keys = keys_s = nullptr;
    ev_cw = nullptr; ev_nn = hend = nullptr; long_list = nullptr; pushflag = pushoff = nullptr;
  }
  void free_all() {
    free_window_w();
    be.free(tcl); be.free(tsum); be.free(segfirst); be.free(flags);
    be.free(cw); be.free(nn); be.free(unused); be.free(perm); be.free(pos);
    tcl = nullptr; tsum = nullptr; segfirst = nullptr; flags = nullptr;
    cw = nullptr; nn = nullptr; unused = nullptr; perm = pos = nullptr;
    Wcap = Tcap = 0;
  }

  // ---- one exact step of the reference for the node at permutation position p (single thread):
  // the W = 1 window; also handles the "popper" (a node that leaves a non-empty cluster for an
  // unused cluster id, ModularityOptimizer.cpp:541-545), whose id comes off the unused stack.
  // out[0] = 1 when the node stayed where it was.
  void step_sequential(i64 p, i32 *d_out) {
    const Net nt = net;
    i32 *cl_ = cl, *nn_ = nn, *unused_ = unused, *perm_ = perm, *tcl_ = tcl;
    double *cw_ = cw, *tsum_ = tsum;
    const double res = resolution;
    i32 *nun = flags + 4;  // device copy of nUnused
    be.single([=] MO_HD() {
      const i32 j = perm_[p];
      // neighbouring clusters in order of first appearance; tcl_/tsum_ as a plain list
      i32 cnt = 0;
      for (i64 e = nt.first[j]; e < nt.first[j + 1]; ++e) {
        const i32 l = cl_[nt.nbr[e]];
        i32 s = 0;
        while (s < cnt && tcl_[s] != l) ++s;
        if (s == cnt) {
          tcl_[cnt] = l;
          tsum_[cnt] = 0.0;
          ++cnt;
        }
        tsum_[s] = dadd(tsum_[s], nt.ew[e]);
      }
      const i32 own = cl_[j];
      const double w = nt.nw[j];
      cw_[own] = dsub(cw_[own], w);
      nn_[own]--;
      i32 nu = *nun;
      if (nn_[own] == 0) unused_[nu++] = own;
      i32 best = -1;
      double mx = 0.0;
      for (i32 s = 0; s < cnt; ++s) {
        const i32 l = tcl_[s];
        const double q = dsub(tsum_[s], dmul(dmul(w, cw_[l]), res));
        if (q > mx || (q == mx && l < best)) {
          best = l;
          mx = q;
        }
      }
      if (mx == 0.0) best = unused_[--nu];
      cw_[best] = dadd(cw_[best], w);
      nn_[best]++;
      *nun = nu;
      d_out[0] = best == own ? 1 : 0;
      cl_[j] = best;
    });
  }

  // returns true when any node changed its cluster (the reference's `update`)
  // host_perm: the visiting order when the caller drew it already (subnetworks of the SLM step)
  bool run(i32 *cl_io, i32 *n_clusters_io, JavaRandom &rng, const i32 *host_perm = nullptr);
};

template <class B>
bool LocalMoving<B>::run(i32 *cl_io, i32 *n_clusters_io, JavaRandom &rng, const i32 *host_perm) {
  const i32 n = net.n;
  if (n == 1) return false;
  if (stats) stats->local_moving_calls++;
  cl = cl_io;
  cw = be.template alloc<double>(n);
  nn = be.template alloc<i32>(n);
  unused = be.template alloc<i32>(n);
  perm = be.template alloc<i32>(n);
  pos = be.template alloc<i32>(n);
  segfirst = be.template alloc<i32>(n);
  flags = be.template alloc<i32>(8);
  const Net nt = net;
  const double res = resolution;

  // ---- cluster weights and sizes: clusterWeight[cluster[i]] += nodeWeight[i] for ascending i
  //      (:484-488) = one sequential chain per cluster over its nodes in index order
  {
    u64 *k1 = be.template alloc<u64>(n), *k2 = be.template alloc<u64>(n);
    i32 *cl_ = cl, *nn_ = nn, *seg_ = segfirst;
    double *cw_ = cw;
    be.pfor(n, [=] MO_HD(i64 v) {
      k1[v] = ((u64)(u32)cl_[v] << 32) | (u64)(u32)v;
      cw_[v] = 0.0;
      nn_[v] = 0;
      seg_[v] = -1;
    });
    be.sort_keys(k1, k2, n, 64);
    double *nws = be.template alloc<double>(n);
    be.pfor(n, [=] MO_HD(i64 s) { nws[s] = nt.nw[(i32)(u32)k2[s]]; });
    be.pfor(n, [=] MO_HD(i64 s) {
      const i32 c = (i32)(k2[s] >> 32);
      if (s > 0 && (i32)(k2[s - 1] >> 32) == c) return;
      double acc = 0.0;
      i32 cnt = 0;
      for (i64 e = s; e < n && (i32)(k2[e] >> 32) == c; ++e) {
        acc = dadd(acc, nws[e]);
        ++cnt;
      }
      cw_[c] = acc;
      nn_[c] = cnt;
    });
    be.free(nws);
    be.free(k1);
    be.free(k2);
  }
  // ---- unused cluster ids, ascending (:490-498)
  {
    i64 *f = be.template alloc<i64>(n + 1), *o = be.template alloc<i64>(n + 1);
    i32 *nn_ = nn, *unused_ = unused;
    be.pfor(n, [=] MO_HD(i64 c) { f[c] = nn_[c] == 0 ? 1 : 0; });
    be.excl_scan(f, o, n);
    be.pfor(n, [=] MO_HD(i64 c) {
      if (nn_[c] == 0) unused_[o[c]] = (i32)c;
    });
    i64 tot = 0;
    be.d2h(&tot, o + n, sizeof(i64));
    nUnused = (i32)tot;
    be.free(f);
    be.free(o);
  }
  // ---- the visiting order (host control data) and its inverse
  {
    std::vector<i32> hp;
    if (!host_perm) random_permutation(rng, n, hp);
    be.h2d(perm, host_perm ? host_perm : hp.data(), sizeof(i32) * (size_t)n);
    i32 *perm_ = perm, *pos_ = pos;
    be.pfor(n, [=] MO_HD(i64 t) { pos_[perm_[t]] = (i32)t; });
  }

  bool update = false;
  i64 nStable = 0;
  i64 i = 0;                 // position in the permutation
  i64 W = std::min<i64>(n, be.initial_window(n));
  const i64 Wmax = std::min<i64>(n, be.max_window(n));
  const int round_cap = 24;
  while (nStable < n && !be.failed()) {
    i64 Wn = std::min<i64>(W, n - i);
    // ---------------------------------------------------------------- W = 1: the reference step
    if (Wn <= 1) {
      alloc_window(1, 2 * net.maxdeg + 4);
      i32 h_nu = nUnused;
      be.h2d(flags + 4, &h_nu, sizeof(i32));
      step_sequential(i, flags + 5);
      i32 out[2];
      be.d2h(out, flags + 4, 2 * sizeof(i32));
      nUnused = out[0];
      if (out[1]) nStable++; else { nStable = 1; update = true; }
      if (stats) stats->seq_steps++;
      i = (i < n - 1) ? i + 1 : 0;
      if (W < 2 && Wmax >= 2) W = 2;
      continue;
    }
    // ---------------------------------------------------------------- a window of Wn positions
    alloc_window(Wn, 0);
    {  // slots of a node's table: a power of two >= 2 * degree (>= 2)
      i32 *perm_ = perm;
      i64 *tcap_ = tcap_in;
      const i64 t0 = i;
      be.pfor(Wn, [=] MO_HD(i64 t) {
        const i32 j = perm_[t0 + t];
        const i64 deg = nt.first[j + 1] - nt.first[j];
        i64 c = 2;
        while (c < 2 * deg) c <<= 1;
        tcap_[t] = c;
      });
    }
    be.excl_scan(tcap_in, toff, Wn);
    i64 T = 0;
    be.d2h(&T, toff + Wn, sizeof(i64));
    alloc_window(Wn, T);
    {
      i32 *perm_ = perm, *cl_ = cl, *D_ = D, *c0_ = c0;
      const i64 t0 = i;
      be.pfor(Wn, [=] MO_HD(i64 t) {
        const i32 c = cl_[perm_[t0 + t]];
        c0_[t] = c;
        D_[t] = c;  // round 0: everybody stays
      });
    }
    int rounds = 0;
    bool converged = false;
    i64 max_chain = 0;
    while (rounds < round_cap && !be.failed()) {
      ++rounds;
      const i64 t0 = i;
      i32 *perm_ = perm, *pos_ = pos, *cl_ = cl, *D_ = D, *D2_ = D2, *c0_ = c0, *tcl_ = tcl, *nn_ = nn;
      i32 *leave_ = leave_nn, *hend_ = hend, *seg_ = segfirst, *flags_ = flags, *evnn_ = ev_nn;
      double *tsum_ = tsum, *cw_ = cw, *evcw_ = ev_cw;
      i64 *toff_ = toff;
      u64 *keys_ = keys, *ks_ = keys_s;
      be.mark("lm.other");
      // (1) clear the tables, reset the flags
      be.pfor(T, [=] MO_HD(i64 s) { tcl_[s] = -1; });
      be.single([=] MO_HD() {
        flags_[0] = 0;
        flags_[1] = 0x7fffffff;
        flags_[2] = 0x7fffffff;
        flags_[3] = -1;
        flags_[6] = 0;  // longest chain of the round (diagnostics)
        flags_[7] = 0;  // number of long chains
      });
      be.mark("lm.clear");
      // (2) edge weight towards every neighbouring cluster, summed in edge order (:507-516);
      //     neighbours visited earlier in this window count with last round's decision
      be.pfor(Wn, [=] MO_HD(i64 t) {
        const i32 j = perm_[t0 + t];
        const i64 base = toff_[t];
        const u32 mask = (u32)(toff_[t + 1] - base) - 1u;
        for (i64 e = nt.first[j]; e < nt.first[j + 1]; ++e) {
          const i32 v = nt.nbr[e];
          const i64 pv = (i64)pos_[v] - t0;
          const i32 l = (pv >= 0 && pv < t) ? D_[pv] : cl_[v];
          u32 h = hash_cluster(l) & mask;
          while (tcl_[base + h] != l && tcl_[base + h] != -1) h = (h + 1) & mask;
          if (tcl_[base + h] == -1) {
            tcl_[base + h] = l;
            tsum_[base + h] = 0.0;
          }
          tsum_[base + h] = dadd(tsum_[base + h], nt.ew[e]);
        }
        // (3a) the node's two events: it leaves c0 at time t and joins D[t] right after
        keys_[2 * t] = ((u64)(u32)c0_[t] << 32) | ((u64)t << 1);
        keys_[2 * t + 1] = ((u64)(u32)D_[t] << 32) | ((u64)t << 1) | 1ull;
      });
      be.mark("lm.gather");
      // (3b) events grouped by cluster, in visiting order
      be.sort_keys(keys, keys_s, 2 * Wn, 64);
      be.mark("lm.sort");
      // (4) every cluster's weight and size replayed along its chain (:518-519, :547-548);
      //     ev_cw first holds the weight of the event's node so that the chain streams
      be.pfor(2 * Wn, [=] MO_HD(i64 s) { evcw_[s] = nt.nw[perm_[t0 + ((u32)ks_[s] >> 1)]]; });
      // short chains: the thread that finds the head walks it; long chains (a large cluster owns a
      // large share of the window) are listed and replayed by the backend's cooperative walker
      // (one warp per chain on the device: coalesced loads, the additions still strictly in order)
      i32 *longl_ = long_list;
      be.pfor(2 * Wn, [=] MO_HD(i64 s) {
        const i32 c = (i32)(ks_[s] >> 32);
        if (s > 0 && (i32)(ks_[s - 1] >> 32) == c) return;
        i64 lo = s + 1, hi = 2 * Wn;  // end of the segment: first key of a later cluster
        while (lo < hi) {
          const i64 mid = (lo + hi) >> 1;
          if ((i32)(ks_[mid] >> 32) == c) lo = mid + 1; else hi = mid;
        }
        const i64 e = lo;
        seg_[c] = (i32)s;
        hend_[s] = (i32)e;
        be_atomic_max(&flags_[6], (i32)(e - s));  // longest chain (diagnostics)
        if (e - s >= LONG_CHAIN) {
          longl_[be_atomic_add(&flags_[7], 1)] = (i32)s;
          return;
        }
        chain_walk(ks_, evcw_, evnn_, s, e, cw_[c], nn_[c]);
      });
      be.long_chains(ks_, evcw_, evnn_, hend_, cw_, nn_, long_list, flags + 7);
      be.mark("lm.chain");
      // (5) every node decides (:521-545)
      be.pfor(Wn, [=] MO_HD(i64 t) {
        const i32 j = perm_[t0 + t];
        const double w = nt.nw[j];
        const i32 own = c0_[t];
        const u32 mine = (u32)t << 1;
        // own cluster right after leaving it: the leave event itself
        i32 own_left = 0;
        {
          i32 lo = seg_[own], hi = hend_[lo];
          while (hi - lo > 1) {  // last event with low word <= mine
            const i32 mid = (lo + hi) >> 1;
            if ((u32)ks_[mid] <= mine) lo = mid; else hi = mid;
          }
          own_left = evnn_[lo];
        }
        leave_[t] = own_left;
        const i64 base = toff_[t];
        const i64 cap = toff_[t + 1] - base;
        i32 best = -1;
        double mx = 0.0;
        for (i64 s = 0; s < cap; ++s) {
          const i32 l = tcl_[base + s];
          if (l < 0) continue;
          // weight of cluster l as node(t) sees it: after every event before time t, and after
          // its own departure when l is its own cluster
          double cwl = cw_[l];
          const i32 sf = seg_[l];
          if (sf >= 0) {
            i32 lo = sf, hi = hend_[sf];
            const u32 lim = (l == own) ? mine + 1u : mine;  // events with low word < lim
            if ((u32)ks_[lo] < lim) {
              while (hi - lo > 1) {
                const i32 mid = (lo + hi) >> 1;
                if ((u32)ks_[mid] < lim) lo = mid; else hi = mid;
              }
              cwl = evcw_[lo];
            }
          }
          const double q = dsub(tsum_[base + s], dmul(dmul(w, cwl), res));
          if (q > mx || (q == mx && l < best)) {
            best = l;
            mx = q;
          }
        }
        if (mx == 0.0) {
          if (own_left == 0) {
            best = own;           // pushes its own id and pops it again
          } else {
            best = own;           // placeholder; the window is cut in front of this node
            be_atomic_min(&flags_[1], (i32)t);
          }
        }
        D2_[t] = best;
        if (best != D_[t]) flags_[0] = 1;
      });
      be.mark("lm.decide");
      // (6) forget this round's segments
      be.pfor(2 * Wn, [=] MO_HD(i64 s) {
        const i32 c = (i32)(ks_[s] >> 32);
        if (s > 0 && (i32)(ks_[s - 1] >> 32) == c) return;
        seg_[c] = -1;
      });
      i32 hf[7];
      be.d2h(hf, flags, 7 * sizeof(i32));
      max_chain = std::max<i64>(max_chain, hf[6]);
      be.mark("lm.unseg+readback");
      std::swap(D, D2);
      if (!hf[0]) {
        converged = true;
        break;
      }
    }
    be.note_window(n, Wn, rounds, converged, max_chain);
    if (stats) {
      stats->rounds += rounds;
      stats->windows++;
      stats->max_window = std::max<i64>(stats->max_window, Wn);
    }
    if (!converged) {  // long dependency chains: retry with a smaller window
      W = std::max<i64>(1, Wn / 4);
      if (stats) stats->shrinks++;
      continue;
    }
    // ---- the fixed point is the sequential execution up to the first popper ---------------------
    // D (after the swap) holds the converged decisions; the chains / tables of the last round
    // belong to exactly these decisions.
    i32 hf[4];
    be.d2h(hf, flags, 2 * sizeof(i32));
    i64 tc = std::min<i64>(Wn, (i64)hf[1]);  // commit positions [0, tc)
    {
      i32 *D_ = D, *c0_ = c0, *flags_ = flags;
      be.pfor(tc, [=] MO_HD(i64 t) {
        if (D_[t] != c0_[t]) {
          be_atomic_min(&flags_[2], (i32)t);
          be_atomic_max(&flags_[3], (i32)t);
        }
      });
      be.d2h(hf, flags, 4 * sizeof(i32));
    }
    const i64 u1 = std::min<i64>(tc, (i64)hf[2]);
    bool done = false;
    if (nStable + u1 >= n) {  // n consecutive stable nodes are reached inside the stable prefix
      tc = n - nStable;       // the node that completes them is the last one processed (:565)
      done = true;
    }
    const bool any_unstable = hf[3] >= 0 && (i64)hf[3] < tc;
    // ---- commit [0, tc)
    if (tc > 0) {
      const i64 t0 = i;
      i32 *perm_ = perm, *cl_ = cl, *D_ = D, *c0_ = c0, *leave_ = leave_nn, *unused_ = unused, *nn_ = nn;
      i32 *hend_ = hend, *evnn_ = ev_nn;
      double *cw_ = cw, *evcw_ = ev_cw;
      i64 *pf = pushflag, *po = pushoff;
      u64 *ks_ = keys_s;
      const i32 nu0 = nUnused;
      be.pfor(tc, [=] MO_HD(i64 t) {
        cl_[perm_[t0 + t]] = D_[t];
        pf[t] = (leave_[t] == 0 && D_[t] != c0_[t]) ? 1 : 0;
      });
      be.excl_scan(pushflag, pushoff, tc);
      be.pfor(tc, [=] MO_HD(i64 t) {
        if (pf[t]) unused_[nu0 + po[t]] = c0_[t];
      });
      const u32 lim = (u32)tc << 1;
      be.pfor(2 * Wn, [=] MO_HD(i64 s) {
        const i32 c = (i32)(ks_[s] >> 32);
        if (s > 0 && (i32)(ks_[s - 1] >> 32) == c) return;
        // segment bounds were published in hend at the head
        i32 lo = (i32)s, hi = hend_[s];
        if ((u32)ks_[lo] >= lim) return;  // no event of this cluster before the cut
        while (hi - lo > 1) {
          const i32 mid = (lo + hi) >> 1;
          if ((u32)ks_[mid] < lim) lo = mid; else hi = mid;
        }
        cw_[c] = evcw_[lo];
        nn_[c] = evnn_[lo];
      });
      i64 pushed = 0;
      be.d2h(&pushed, pushoff + tc, sizeof(i64));
      nUnused += (i32)pushed;
      if (any_unstable) {
        nStable = tc - (i64)hf[3];
        update = true;
      } else {
        nStable += tc;
      }
      i += tc;
      if (i >= n) i = 0;
    }
    if (done || nStable >= n) break;
    if (tc < Wn && (i64)hf[1] == tc) {
      // the node at the cut takes an id off the unused stack: the reference step itself
      alloc_window(1, 2 * net.maxdeg + 4);
      i32 h_nu = nUnused;
      be.h2d(flags + 4, &h_nu, sizeof(i32));
      step_sequential(i, flags + 5);
      i32 out[2];
      be.d2h(out, flags + 4, 2 * sizeof(i32));
      nUnused = out[0];
      if (out[1]) nStable++; else { nStable = 1; update = true; }
      if (stats) stats->seq_steps++;
      i = (i < n - 1) ? i + 1 : 0;
    }
    // adapt the window: few rounds -> larger, many -> smaller
    if (rounds <= 4) W = std::min<i64>(Wmax, std::max<i64>(W, Wn) * 2);
    else if (rounds > 10) W = std::max<i64>(std::min<i64>(2, Wmax), Wn / 2);
  }

  // ---- renumber the clusters in ascending order of their ids (:567-578)
  {
    i64 *f = be.template alloc<i64>(n + 1), *o = be.template alloc<i64>(n + 1);
    i32 *nn_ = nn, *cl_ = cl;
    be.pfor(n, [=] MO_HD(i64 c) { f[c] = nn_[c] > 0 ? 1 : 0; });
    be.excl_scan(f, o, n);
    be.pfor(n, [=] MO_HD(i64 v) { cl_[v] = (i32)o[cl_[v]]; });
    i64 tot = 0;
    be.d2h(&tot, o + n, sizeof(i64));
    *n_clusters_io = (i32)tot;
    be.free(f);
    be.free(o);
  }
  free_all();
  return update;
}

// =================================================================================================
// Order-fixed running sums (std::accumulate / `+=` loops of the reference) without walking them
// element by element.  For s in [2^e, 2^(e+1)) and x >= 0 with s + x < 2^(e+1), fl(s + x) is
// s + rn(x / u) * u with u = ulp(s) = 2^(e-52): while the running sum stays inside one binade the
// sequence of rounded additions is an INTEGER prefix sum of r_t = rn(x_t / u), exact and
// associative -- except for exact ties (frac(x_t / u) = 1/2), whose rounding depends on the parity
// of the running sum.  So: block sums of r_t in parallel, one pass over the block sums to find the
// first block that leaves the binade or holds a tie (or a negative / non-finite term), that block
// alone with real additions, then on with the next binade.  About 25 binades cover 15 M terms.
// =================================================================================================
constexpr i64 SUM_BLOCK = 2048;
constexpr u64 TWO53 = 1ull << 53;

template <class B>
double seq_sum(B &be, const double *x, i64 n, double init) {
  double s = init;
  i64 pos = 0;
  if (n < 4 * SUM_BLOCK) return be.seq_sum_plain(x, n, s);
  const i64 nb_all = (n + SUM_BLOCK - 1) / SUM_BLOCK;
  u64 *bsum = be.template alloc<u64>(nb_all);
  i32 *bflag = be.template alloc<i32>(nb_all);
  u64 *res = be.template alloc<u64>(2);
  u64 *gsum = be.template alloc<u64>(nb_all / 64 + 1);
  i32 *gflag = be.template alloc<i32>(nb_all / 64 + 1);
  while (pos < n && !be.failed()) {
    int e = 0;
    const bool ok = s > 0.0 && s < 1e300 && s >= 2.3e-308 && (frexp(s, &e), true);
    if (!ok) {  // zero / subnormal / huge / negative running sum: this block with real additions
      const i64 len = std::min<i64>(SUM_BLOCK, n - pos);
      s = be.seq_sum_plain(x + pos, len, s);
      pos += len;
      continue;
    }
    const int ex = e - 1;             // s in [2^ex, 2^(ex+1))
    const int shift = 52 - ex;        // x / u = ldexp(x, shift)
    const u64 S = (u64)ldexp(s, shift);
    const i64 nb = (n - pos + SUM_BLOCK - 1) / SUM_BLOCK;
    const double *xp = x + pos;
    const i64 rem = n - pos;
    be.pfor(nb, [=] MO_HD(i64 b) {
      u64 acc = 0;
      i32 flag = 0;
      const i64 hi = (b + 1) * SUM_BLOCK < rem ? (b + 1) * SUM_BLOCK : rem;
      for (i64 t = b * SUM_BLOCK; t < hi; ++t) {
        const double v = xp[t];
        if (!(v >= 0.0) || v > 1e300) {  // negative, NaN or huge: real additions
          flag = 1;
          break;
        }
        const double y = ldexp(v, shift);
        if (y >= 9007199254740992.0) {   // >= 2^53: certainly leaves the binade
          flag = 1;
          break;
        }
        const double fl = floor(y);
        const double fr = y - fl;        // exact
        if (fr == 0.5) {                 // tie: depends on the parity of the running sum
          flag = 1;
          break;
        }
        acc += (u64)fl + (fr > 0.5 ? 1ull : 0ull);
        if (acc >= TWO53) {
          acc = TWO53;
          break;
        }
      }
      bsum[b] = acc;
      bflag[b] = flag;
    });
    // first block that cannot be taken wholesale: groups of 64 blocks first, then inside the group
    const i64 ng = (nb + 63) / 64;
    be.pfor(ng, [=] MO_HD(i64 g) {
      u64 acc = 0;
      i32 flag = 0;
      const i64 hi = (g + 1) * 64 < nb ? (g + 1) * 64 : nb;
      for (i64 b = g * 64; b < hi; ++b) {
        flag |= bflag[b];
        acc += bsum[b];
        if (acc >= TWO53) acc = TWO53;
      }
      gsum[g] = acc;
      gflag[g] = flag;
    });
    be.single([=] MO_HD() {
      u64 acc = S;
      i64 g = 0;
      for (; g < ng; ++g) {
        if (gflag[g] || acc + gsum[g] >= TWO53) break;
        acc += gsum[g];
      }
      i64 b = g * 64;
      for (; b < nb; ++b) {
        if (bflag[b] || acc + bsum[b] >= TWO53) break;
        acc += bsum[b];
      }
      res[0] = acc;
      res[1] = (u64)b;
    });
    u64 h[2];
    be.d2h(h, res, sizeof(h));
    s = ldexp((double)h[0], -shift);   // exact: an integer below 2^53 times a power of two
    pos += (i64)h[1] * SUM_BLOCK;
    if (pos >= n) break;
    if ((i64)h[1] < nb) {               // the block that could not be taken wholesale
      const i64 len = std::min<i64>(SUM_BLOCK, n - pos);
      s = be.seq_sum_plain(x + pos, len, s);
      pos += len;
    }
  }
  be.free(bsum);
  be.free(bflag);
  be.free(res);
  be.free(gsum);
  be.free(gflag);
  return s;
}

// =================================================================================================
// Networks
// =================================================================================================
template <class B>
void free_net(B &be, Net &nt) {
  if (!nt.owned) return;
  be.free(nt.first); be.free(nt.nbr); be.free(nt.ew); be.free(nt.nw);
  nt = Net();
}

template <class B>
void finish_net(B &be, Net &nt) {  // maximum degree (sizes the scratch of the single-node step)
  i64 *mx = be.template alloc<i64>(1);
  const Net c = nt;
  be.single([=] MO_HD() { mx[0] = 0; });
  be.pfor(nt.n, [=] MO_HD(i64 v) { be_atomic_max64(mx, c.first[v + 1] - c.first[v]); });
  be.d2h(&nt.maxdeg, mx, sizeof(i64));
  be.free(mx);
}

// first[v] = first sorted record whose node (key >> shift) is >= v
template <class B>
void first_from_sorted(B &be, const u64 *ks, i64 m, int shift, i64 *first, i32 n) {
  be.pfor((i64)n + 1, [=] MO_HD(i64 v) {
    i64 lo = 0, hi = m;  // first index with (key >> shift) >= v
    while (lo < hi) {
      const i64 mid = (lo + hi) >> 1;
      if ((i64)(ks[mid] >> shift) < v) lo = mid + 1; else hi = mid;
    }
    first[v] = lo;
  });
}

// The network of the strict lower triangle of a column-compressed matrix, as
// RModularityOptimizer.cpp:72-86 + matrixToNetwork (ModularityOptimizer.cpp:729-770) build it:
// edge list in column-major order (node1 = column < node2 = row); the adjacency of v lists, in that
// order, first the edges in which v is the row, then those in which it is the column.
// Returns the number of edges (0: "Matrix contained no network data").
template <class B, typename PT>
i64 build_network(B &be, i32 n, const PT *p, const i32 *ri, const double *x, int modularityFunction, Net &out) {
  i64 *cnt = be.template alloc<i64>((i64)n + 1), *off = be.template alloc<i64>((i64)n + 1);
  be.pfor(n, [=] MO_HD(i64 c) {
    i64 k = 0;
    for (i64 e = (i64)p[c]; e < (i64)p[c + 1]; ++e) k += ri[e] > c ? 1 : 0;
    cnt[c] = k;
  });
  be.excl_scan(cnt, off, n);
  i64 E = 0;
  be.d2h(&E, off + n, sizeof(i64));
  if (E == 0) {
    be.free(cnt); be.free(off);
    return 0;
  }
  u64 *k1 = be.template alloc<u64>(2 * E), *k2 = be.template alloc<u64>(2 * E);
  i32 *e_row = be.template alloc<i32>(E), *e_col = be.template alloc<i32>(E);
  double *e_w = be.template alloc<double>(E);
  be.pfor(n, [=] MO_HD(i64 c) {
    i64 o = off[c];
    for (i64 e = (i64)p[c]; e < (i64)p[c + 1]; ++e) {
      const i32 r = ri[e];
      if (r <= c) continue;
      e_row[o] = r;
      e_col[o] = (i32)c;
      e_w[o] = x[e];
      k1[2 * o] = ((u64)(u32)r << 33) | (u64)o;                       // v is the row: first part
      k1[2 * o + 1] = ((u64)(u32)c << 33) | (1ull << 32) | (u64)o;    // v is the column: second part
      ++o;
    }
  });
  be.sort_keys(k1, k2, 2 * E, 64);
  out = Net();
  out.n = n;
  out.m2 = 2 * E;
  out.owned = true;
  out.first = be.template alloc<i64>((i64)n + 1);
  out.nbr = be.template alloc<i32>(2 * E);
  out.ew = be.template alloc<double>(2 * E);
  out.nw = be.template alloc<double>(n);
  first_from_sorted(be, k2, 2 * E, 33, out.first, n);
  {
    i32 *nbr = out.nbr;
    double *ew = out.ew;
    be.pfor(2 * E, [=] MO_HD(i64 s) {
      const u64 k = k2[s];
      const i64 e = (i64)(u32)k;
      nbr[s] = (k >> 32) & 1ull ? e_row[e] : e_col[e];
      ew[s] = e_w[e];
    });
  }
  {
    const Net c = out;
    double *nw = out.nw;
    if (modularityFunction == 1)  // getTotalEdgeWeightPerNode (:253-259): adjacency order
      be.pfor(n, [=] MO_HD(i64 v) {
        double acc = 0.0;
        for (i64 e = c.first[v]; e < c.first[v + 1]; ++e) acc = dadd(acc, c.ew[e]);
        nw[v] = acc;
      });
    else
      be.pfor(n, [=] MO_HD(i64 v) { nw[v] = 1.0; });
  }
  be.free(cnt); be.free(off); be.free(k1); be.free(k2); be.free(e_row); be.free(e_col); be.free(e_w);
  finish_net(be, out);
  return E;
}

// The network of an arbitrary edge list (the edge.file.name input of RunModularityClusteringCpp,
// RModularityOptimizer.cpp:55-63: readInputFile, ModularityOptimizer.cpp:806-838, then the same
// matrixToNetwork :759-803).  Only edges with node1 < node2 count (:765, :784); the adjacency of v lists
// its edges in the order of the list, whichever end v is.  Returns the number of edges kept.
template <class B>
i64 build_network_edges(B &be, i32 n, const i32 *node1, const i32 *node2, const double *w, i64 M,
                        int modularityFunction, Net &out) {
  if (M <= 0) return 0;
  i64 *cnt = be.template alloc<i64>(M + 1), *off = be.template alloc<i64>(M + 1);
  be.pfor(M, [=] MO_HD(i64 e) { cnt[e] = node1[e] < node2[e] ? 1 : 0; });
  be.excl_scan(cnt, off, M);
  i64 E = 0;
  be.d2h(&E, off + M, sizeof(i64));
  if (E == 0) {
    be.free(cnt); be.free(off);
    return 0;
  }
  u64 *k1 = be.template alloc<u64>(2 * E), *k2 = be.template alloc<u64>(2 * E);
  be.pfor(M, [=] MO_HD(i64 e) {
    if (!(node1[e] < node2[e])) return;
    const i64 o = off[e];  // rank among the kept edges (o < 2^32)
    k1[2 * o] = ((u64)(u32)node1[e] << 33) | ((u64)o << 1) | 1ull;   // v = node1: neighbour node2
    k1[2 * o + 1] = ((u64)(u32)node2[e] << 33) | ((u64)o << 1);      // v = node2: neighbour node1
  });
  // kept rank -> position in the list
  i64 *src = be.template alloc<i64>(E);
  be.pfor(M, [=] MO_HD(i64 e) { if (node1[e] < node2[e]) src[off[e]] = e; });
  be.sort_keys(k1, k2, 2 * E, 64);
  out = Net();
  out.n = n;
  out.m2 = 2 * E;
  out.owned = true;
  out.first = be.template alloc<i64>((i64)n + 1);
  out.nbr = be.template alloc<i32>(2 * E);
  out.ew = be.template alloc<double>(2 * E);
  out.nw = be.template alloc<double>(n);
  first_from_sorted(be, k2, 2 * E, 33, out.first, n);
  {
    i32 *nbr = out.nbr;
    double *ew = out.ew;
    be.pfor(2 * E, [=] MO_HD(i64 s) {
      const u64 k = k2[s];
      const i64 e = src[(i64)((k >> 1) & 0xffffffffull)];
      nbr[s] = (k & 1ull) ? node2[e] : node1[e];
      ew[s] = w[e];
    });
  }
  {
    const Net c = out;
    double *nw = out.nw;
    if (modularityFunction == 1)
      be.pfor(n, [=] MO_HD(i64 v) {
        double acc = 0.0;
        for (i64 e = c.first[v]; e < c.first[v + 1]; ++e) acc = dadd(acc, c.ew[e]);
        nw[v] = acc;
      });
    else
      be.pfor(n, [=] MO_HD(i64 v) { nw[v] = 1.0; });
  }
  be.free(cnt); be.free(off); be.free(k1); be.free(k2); be.free(src);
  finish_net(be, out);
  return E;
}

// Network::createReducedNetwork (ModularityOptimizer.cpp:315-367).  The reference scans cluster by
// cluster, node by node (ascending), edge by edge, adds every edge weight to a per-target running
// sum and lists the targets in order of first appearance.  Here: every adjacency entry gets its
// rank `seq` in that scan, a STABLE sort by (cluster, target) keeps each pair's entries in scan
// order, one sequential chain per pair gives the reference's sum, and the pairs of a cluster are
// ordered by the rank of their first entry.  (totalEdgeWeightSelfLinks of a reduced network is
// never read by the optimiser and is not computed.)
template <class B>
void reduce_network(B &be, const Net &nt, const i32 *cl, i32 nC, Net &out) {
  const i32 n = nt.n;
  const Net c = nt;
  u64 *k1 = be.template alloc<u64>(n), *k2 = be.template alloc<u64>(n);
  be.pfor(n, [=] MO_HD(i64 v) { k1[v] = ((u64)(u32)cl[v] << 32) | (u64)(u32)v; });
  be.sort_keys(k1, k2, n, 64);
  // scan rank of every node's first entry
  i64 *dg = be.template alloc<i64>((i64)n + 1), *sb = be.template alloc<i64>((i64)n + 1);
  double *nws = be.template alloc<double>(n);
  be.pfor(n, [=] MO_HD(i64 s) {
    const i32 v = (i32)(u32)k2[s];
    dg[s] = c.first[v + 1] - c.first[v];
    nws[s] = c.nw[v];
  });
  be.excl_scan(dg, sb, n);
  out = Net();
  out.n = nC;
  out.owned = true;
  out.nw = be.template alloc<double>(nC);
  out.first = be.template alloc<i64>((i64)nC + 1);
  {
    double *nw = out.nw;
    be.pfor(n, [=] MO_HD(i64 s) {  // node weights: cluster members in ascending index order (:335)
      const i32 cc = (i32)(k2[s] >> 32);
      if (s > 0 && (i32)(k2[s - 1] >> 32) == cc) return;
      double acc = 0.0;
      for (i64 e = s; e < n && (i32)(k2[e] >> 32) == cc; ++e) acc = dadd(acc, nws[e]);
      nw[cc] = acc;
    });
  }
  const i64 m2 = nt.m2;
  u64 *pk1 = be.template alloc<u64>(m2 > 0 ? m2 : 1), *pk2 = be.template alloc<u64>(m2 > 0 ? m2 : 1);
  u32 *pv1 = be.template alloc<u32>(m2 > 0 ? m2 : 1), *pv2 = be.template alloc<u32>(m2 > 0 ? m2 : 1);
  double *wseq = be.template alloc<double>(m2 > 0 ? m2 : 1);
  be.pfor(n, [=] MO_HD(i64 s) {
    const i32 v = (i32)(u32)k2[s];
    const i32 cv = (i32)(k2[s] >> 32);
    i64 q = sb[s];
    for (i64 e = c.first[v]; e < c.first[v + 1]; ++e, ++q) {
      const i32 ct = cl[c.nbr[e]];
      pk1[q] = ct == cv ? ~0ull : (((u64)(u32)cv << 32) | (u64)(u32)ct);
      pv1[q] = (u32)q;
      wseq[q] = c.ew[e];
    }
  });
  be.sort_pairs(pk1, pk2, pv1, pv2, m2, 64);
  // pairs (cluster, target): heads of equal-key runs (the all-ones key = inside a cluster, dropped)
  i64 *hf = be.template alloc<i64>(m2 + 1), *ho = be.template alloc<i64>(m2 + 1);
  be.pfor(m2, [=] MO_HD(i64 s) { hf[s] = (pk2[s] != ~0ull && (s == 0 || pk2[s - 1] != pk2[s])) ? 1 : 0; });
  be.excl_scan(hf, ho, m2);
  i64 G = 0;
  be.d2h(&G, ho + m2, sizeof(i64));
  u64 *gk1 = be.template alloc<u64>(G > 0 ? G : 1), *gk2 = be.template alloc<u64>(G > 0 ? G : 1);
  u32 *gv1 = be.template alloc<u32>(G > 0 ? G : 1), *gv2 = be.template alloc<u32>(G > 0 ? G : 1);
  double *gsum = be.template alloc<double>(G > 0 ? G : 1);
  i32 *gtgt = be.template alloc<i32>(G > 0 ? G : 1);
  be.pfor(m2, [=] MO_HD(i64 s) {
    if (!hf[s]) return;
    const u64 key = pk2[s];
    double acc = 0.0;
    for (i64 e = s; e < m2 && pk2[e] == key; ++e) acc = dadd(acc, wseq[pv2[e]]);
    const i64 g = ho[s];
    gsum[g] = acc;
    gtgt[g] = (i32)(u32)key;
    gk1[g] = (key & 0xffffffff00000000ull) | (u64)pv2[s];  // (cluster, rank of the first entry)
    gv1[g] = (u32)g;
  });
  be.sort_pairs(gk1, gk2, gv1, gv2, G, 64);
  out.m2 = G;
  out.nbr = be.template alloc<i32>(G > 0 ? G : 1);
  out.ew = be.template alloc<double>(G > 0 ? G : 1);
  {
    i32 *nbr = out.nbr;
    double *ew = out.ew;
    be.pfor(G, [=] MO_HD(i64 s) {
      nbr[s] = gtgt[gv2[s]];
      ew[s] = gsum[gv2[s]];
    });
  }
  first_from_sorted(be, gk2, G, 32, out.first, nC);
  be.free(k1); be.free(k2); be.free(dg); be.free(sb); be.free(nws); be.free(pk1); be.free(pk2); be.free(pv1);
  be.free(pv2); be.free(wseq); be.free(hf); be.free(ho); be.free(gk1); be.free(gk2); be.free(gv1); be.free(gv2);
  be.free(gsum); be.free(gtgt);
  finish_net(be, out);
}

// VOSClusteringTechnique::calcQualityFunction (ModularityOptimizer.cpp:451-472): ONE running sum over
// all intra-cluster adjacency entries in node / edge order, then minus cw^2 * resolution cluster by
// cluster, divided by (2 * total edge weight + self links).
template <class B>
double quality_function(B &be, const Net &nt, const i32 *cl, i32 nC, double resolution, double total_edge_weight,
                        double self_links) {
  const Net c = nt;
  const i32 n = nt.n;
  const i64 m2 = nt.m2;
  // the intra-cluster entries, compacted in order
  i64 *f = be.template alloc<i64>(m2 + 1), *o = be.template alloc<i64>(m2 + 1);
  be.pfor(n, [=] MO_HD(i64 v) {
    const i32 cv = cl[v];
    for (i64 e = c.first[v]; e < c.first[v + 1]; ++e) f[e] = cl[c.nbr[e]] == cv ? 1 : 0;
  });
  be.excl_scan(f, o, m2);
  i64 K = 0;
  be.d2h(&K, o + m2, sizeof(i64));
  double *xs = be.template alloc<double>(K > 0 ? K : 1);
  be.pfor(m2, [=] MO_HD(i64 e) {
    if (f[e]) xs[o[e]] = c.ew[e];
  });
  double q = seq_sum(be, xs, K, 0.0);
  q = q + self_links;
  // cluster weights in node order, then the penalty cluster by cluster
  u64 *k1 = be.template alloc<u64>(n), *k2 = be.template alloc<u64>(n);
  double *nws = be.template alloc<double>(n), *cw = be.template alloc<double>(nC);
  be.pfor(n, [=] MO_HD(i64 v) { k1[v] = ((u64)(u32)cl[v] << 32) | (u64)(u32)v; });
  be.sort_keys(k1, k2, n, 64);
  be.pfor(n, [=] MO_HD(i64 s) { nws[s] = c.nw[(i32)(u32)k2[s]]; });
  be.pfor(nC, [=] MO_HD(i64 cc) { cw[cc] = 0.0; });
  be.pfor(n, [=] MO_HD(i64 s) {
    const i32 cc = (i32)(k2[s] >> 32);
    if (s > 0 && (i32)(k2[s - 1] >> 32) == cc) return;
    double acc = 0.0;
    for (i64 e = s; e < n && (i32)(k2[e] >> 32) == cc; ++e) acc = dadd(acc, nws[e]);
    cw[cc] = acc;
  });
  double *pen = be.template alloc<double>(nC);
  be.pfor(nC, [=] MO_HD(i64 cc) { pen[cc] = -dmul(dmul(cw[cc], cw[cc]), resolution); });
  q = seq_sum(be, pen, nC, q);
  q /= 2 * total_edge_weight + self_links;
  be.free(f); be.free(o); be.free(xs); be.free(k1); be.free(k2); be.free(nws); be.free(cw); be.free(pen);
  return q;
}

// =================================================================================================
// Louvain / Louvain with multilevel refinement (ModularityOptimizer.cpp:583-602, :616-636)
// =================================================================================================
template <class B>
struct Optimizer {
  B &be;
  double resolution;
  Stats *stats;
  JavaRandom &rng;

  bool local_moving(const Net &nt, i32 *cl, i32 *nC) {
    const double t0 = be.now_ms();
    LocalMoving<B> lm(be, nt, resolution, stats);
    const bool u = lm.run(cl, nC, rng);
    if (stats) stats->ms_local_moving += be.now_ms() - t0;
    return u;
  }

  bool slm(const Net &nt, i32 *cl, i32 *nC, int depth = 0);  // algorithm 3

  // algorithm 1 (refine = false) and 2 (refine = true)
  bool louvain(const Net &nt, i32 *cl, i32 *nC, bool refine, int depth = 0) {
    if (nt.n <= 1 || be.failed()) return false;
    if (stats) stats->levels = std::max<i64>(stats->levels, depth + 1);
    bool update = local_moving(nt, cl, nC);
    if (*nC < nt.n) {
      Net red;
      const double t0 = be.now_ms();
      reduce_network(be, nt, cl, *nC, red);
      if (stats) stats->ms_reduce += be.now_ms() - t0;
      i32 *cl2 = be.template alloc<i32>(red.n);
      i32 nC2 = red.n;
      be.pfor(red.n, [=] MO_HD(i64 v) { cl2[v] = (i32)v; });
      const bool update2 = louvain(red, cl2, &nC2, refine, depth + 1);
      if (update2) {
        update = true;
        be.pfor(nt.n, [=] MO_HD(i64 v) { cl[v] = cl2[cl[v]]; });  // mergeClusters (:159-163)
        *nC = nC2;
        if (refine) local_moving(nt, cl, nC);
      }
      be.free(cl2);
      free_net(be, red);
    }
    return update;
  }
};

// =================================================================================================
// Smart local moving (ModularityOptimizer.cpp:649-692): local moving, then local moving INSIDE every
// cluster (its subnetwork, from singletons), then the same on the reduced network of the refined
// clusters, started from the unrefined partition.
// =================================================================================================
constexpr i32 SLM_SMALL_MAX = 1024;  // subnetworks up to this size: one thread each runs the reference loop

// Network::createSubnetworks (:296-306, :397-440) for all clusters at once: nodes in (cluster, index)
// order = position g; cluster c owns positions [off[c], off[c+1]); edges that stay inside the
// cluster, in adjacency order, with LOCAL neighbour ids; node weights copied (not recomputed, :421).
struct SubNets {
  i64 *off = nullptr;   // [nC + 1]
  i32 *node = nullptr;  // [n] position -> node
  i64 *first = nullptr; // [n + 1]
  i32 *nbr = nullptr;
  double *ew = nullptr, *nw = nullptr;
  i64 m = 0;
};

template <class B>
void build_subnets(B &be, const Net &nt, const i32 *cl, i32 nC, SubNets &S) {
  const i32 n = nt.n;
  const Net c = nt;
  u64 *k1 = be.template alloc<u64>(n), *k2 = be.template alloc<u64>(n);
  i32 *rank = be.template alloc<i32>(n);
  i64 *dg = be.template alloc<i64>((i64)n + 1);
  S.off = be.template alloc<i64>((i64)nC + 1);
  S.node = be.template alloc<i32>(n);
  S.first = be.template alloc<i64>((i64)n + 1);
  S.nw = be.template alloc<double>(n);
  be.pfor(n, [=] MO_HD(i64 v) { k1[v] = ((u64)(u32)cl[v] << 32) | (u64)(u32)v; });
  be.sort_keys(k1, k2, n, 64);
  first_from_sorted(be, k2, n, 32, S.off, nC);
  {
    i32 *node = S.node;
    double *nw = S.nw;
    be.pfor(n, [=] MO_HD(i64 g) {
      const i32 v = (i32)(u32)k2[g];
      node[g] = v;
      rank[v] = (i32)g;
      nw[g] = c.nw[v];
      const i32 cv = cl[v];
      i64 k = 0;
      for (i64 e = c.first[v]; e < c.first[v + 1]; ++e) k += cl[c.nbr[e]] == cv ? 1 : 0;
      dg[g] = k;
    });
  }
  be.excl_scan(dg, S.first, n);
  be.d2h(&S.m, S.first + n, sizeof(i64));
  S.nbr = be.template alloc<i32>(S.m > 0 ? S.m : 1);
  S.ew = be.template alloc<double>(S.m > 0 ? S.m : 1);
  {
    const SubNets T = S;
    be.pfor(n, [=] MO_HD(i64 g) {
      const i32 v = T.node[g];
      const i32 cv = cl[v];
      const i64 o = T.off[cv];
      i64 q = T.first[g];
      for (i64 e = c.first[v]; e < c.first[v + 1]; ++e) {
        const i32 u = c.nbr[e];
        if (cl[u] != cv) continue;
        T.nbr[q] = (i32)(rank[u] - o);
        T.ew[q] = c.ew[e];
        ++q;
      }
    });
  }
  be.free(k1); be.free(k2); be.free(rank); be.free(dg);
}

template <class B>
void free_subnets(B &be, SubNets &S) {
  be.free(S.off); be.free(S.node); be.free(S.first); be.free(S.nbr); be.free(S.ew); be.free(S.nw);
  S = SubNets();
}

// scratch of the one-thread-per-subnetwork local moving, all [n], addressed at the subnetwork's offset
struct SmallScratch {
  i32 *clus, *nn, *unused, *nbc, *newc, *perm, *sub_nc;
  double *cw, *ewpc;
};

// runLocalMovingAlgorithm (:474-581) of ONE subnetwork, from singletons, exactly as the reference
// loops (these networks are small; many of them run side by side, one thread each)
MO_HD inline void lm_small(const SubNets &S, const SmallScratch &W, double res, i32 c) {
  const i64 o = S.off[c];
  const i32 ns = (i32)(S.off[c + 1] - o);
  if (ns == 1) {  // "if (network->getNNodes() == 1) return false" -- no draw, one cluster
    W.clus[o] = 0;
    W.sub_nc[c] = 1;
    return;
  }
  i32 *clus = W.clus + o, *nn = W.nn + o, *unused = W.unused + o, *nbc = W.nbc + o, *newc = W.newc + o;
  const i32 *perm = W.perm + o;
  double *cw = W.cw + o, *ewpc = W.ewpc + o;
  const double *nw = S.nw + o;
  const i64 *first = S.first + o;
  for (i32 j = 0; j < ns; ++j) {
    clus[j] = j;
    cw[j] = dadd(0.0, nw[j]);
    nn[j] = 1;
    ewpc[j] = 0.0;
  }
  i32 nUnused = 0, nStable = 0, i = 0;
  do {
    const i32 j = perm[i];
    i32 k = 0;
    for (i64 e = first[j]; e < first[j + 1]; ++e) {
      const i32 l = clus[S.nbr[e]];
      if (ewpc[l] == 0.0) nbc[k++] = l;
      ewpc[l] = dadd(ewpc[l], S.ew[e]);
    }
    const i32 own = clus[j];
    cw[own] = dsub(cw[own], nw[j]);
    if (--nn[own] == 0) unused[nUnused++] = own;
    i32 best = -1;
    double mx = 0.0;
    for (i32 q = 0; q < k; ++q) {
      const i32 l = nbc[q];
      const double qf = dsub(ewpc[l], dmul(dmul(nw[j], cw[l]), res));
      if (qf > mx || (qf == mx && l < best)) {
        best = l;
        mx = qf;
      }
      ewpc[l] = 0.0;
    }
    if (mx == 0.0) best = unused[--nUnused];
    cw[best] = dadd(cw[best], nw[j]);
    nn[best]++;
    if (best == own) {
      nStable++;
    } else {
      clus[j] = best;
      nStable = 1;
    }
    i = (i < ns - 1) ? i + 1 : 0;
  } while (nStable < ns);
  i32 cnt = 0;
  for (i32 q = 0; q < ns; ++q)
    if (nn[q] > 0) newc[q] = cnt++;
  for (i32 j = 0; j < ns; ++j) clus[j] = newc[clus[j]];
  W.sub_nc[c] = cnt;
}

template <class B>
bool Optimizer<B>::slm(const Net &nt, i32 *cl, i32 *nC, int depth) {
  if (nt.n <= 1 || be.failed()) return false;
  if (stats) stats->levels = std::max<i64>(stats->levels, depth + 1);
  bool update = local_moving(nt, cl, nC);
  if (*nC >= nt.n) return update;
  const i32 n = nt.n, nC0 = *nC;
  const double t0 = be.now_ms();
  SubNets S;
  build_subnets(be, nt, cl, nC0, S);
  std::vector<i64> hoff((size_t)nC0 + 1);
  be.d2h(hoff.data(), S.off, sizeof(i64) * ((size_t)nC0 + 1));
  // the visiting orders, drawn cluster by cluster as the reference does (:660-663 -> :500)
  std::vector<i32> hperm((size_t)n), one;
  for (i32 c = 0; c < nC0; ++c) {
    const i32 ns = (i32)(hoff[c + 1] - hoff[c]);
    if (ns <= 1) {
      if (ns == 1) hperm[hoff[c]] = 0;
      continue;
    }
    random_permutation(rng, ns, one);
    std::copy(one.begin(), one.end(), hperm.begin() + hoff[c]);
  }
  SmallScratch W;
  W.clus = be.template alloc<i32>(n); W.nn = be.template alloc<i32>(n); W.unused = be.template alloc<i32>(n);
  W.nbc = be.template alloc<i32>(n); W.newc = be.template alloc<i32>(n); W.perm = be.template alloc<i32>(n);
  W.sub_nc = be.template alloc<i32>(nC0); W.cw = be.template alloc<double>(n); W.ewpc = be.template alloc<double>(n);
  be.h2d(W.perm, hperm.data(), sizeof(i32) * (size_t)n);
  {
    const SubNets T = S;
    const SmallScratch V = W;
    const double res = resolution;
    be.pfor(nC0, [=] MO_HD(i64 c) {
      if (T.off[c + 1] - T.off[c] <= SLM_SMALL_MAX) lm_small(T, V, res, (i32)c);
    });
  }
  // the large subnetworks through the windowed local moving, one after the other
  for (i32 c = 0; c < nC0 && !be.failed(); ++c) {
    const i64 o = hoff[c];
    const i32 ns = (i32)(hoff[c + 1] - o);
    if (ns <= SLM_SMALL_MAX) continue;
    i64 fe[2];
    be.d2h(&fe[0], S.first + o, sizeof(i64));
    be.d2h(&fe[1], S.first + o + ns, sizeof(i64));
    Net sub;
    sub.n = ns;
    sub.m2 = fe[1] - fe[0];
    sub.first = be.template alloc<i64>((i64)ns + 1);
    sub.nbr = S.nbr + fe[0];
    sub.ew = S.ew + fe[0];
    sub.nw = S.nw + o;
    {
      i64 *f = sub.first;
      const i64 *sf = S.first + o;
      const i64 b0 = fe[0];
      be.pfor((i64)ns + 1, [=] MO_HD(i64 j) { f[j] = sf[j] - b0; });
    }
    finish_net(be, sub);
    i32 *cls = W.clus + o;
    be.pfor(ns, [=] MO_HD(i64 j) { cls[j] = (i32)j; });
    i32 snc = ns;
    LocalMoving<B> lm(be, sub, resolution, stats);
    lm.run(cls, &snc, rng, hperm.data() + o);
    be.h2d(W.sub_nc + c, &snc, sizeof(i32));
    be.free(sub.first);
  }
  // refined clustering: the sub-clusters of cluster c follow those of the clusters before it (:664-667)
  i64 *cin = be.template alloc<i64>((i64)nC0 + 1), *base = be.template alloc<i64>((i64)nC0 + 1);
  {
    const i32 *snc = W.sub_nc;
    be.pfor(nC0, [=] MO_HD(i64 c) { cin[c] = snc[c]; });
  }
  be.excl_scan(cin, base, nC0);
  i64 nRef = 0;
  be.d2h(&nRef, base + nC0, sizeof(i64));
  {
    const SubNets T = S;
    const i32 *clus = W.clus;
    be.pfor(n, [=] MO_HD(i64 g) {
      const i32 v = T.node[g];
      cl[v] = (i32)(base[cl[v]] + clus[g]);
    });
  }
  if (stats) stats->ms_subnetworks += be.now_ms() - t0;
  // the reduced network of the refined clusters, started from the unrefined partition (:670-684)
  Net red;
  const double t1 = be.now_ms();
  reduce_network(be, nt, cl, (i32)nRef, red);
  if (stats) stats->ms_reduce += be.now_ms() - t1;
  i32 *cl2 = be.template alloc<i32>(red.n);
  be.pfor(nRef, [=] MO_HD(i64 r) {
    i64 lo = 0, hi = nC0;  // the cluster c with base[c] <= r < base[c + 1]
    while (hi - lo > 1) {
      const i64 mid = (lo + hi) >> 1;
      if (base[mid] <= r) lo = mid; else hi = mid;
    }
    cl2[r] = (i32)lo;
  });
  i32 nC2 = nC0;
  update |= slm(red, cl2, &nC2, depth + 1);
  be.pfor(n, [=] MO_HD(i64 v) { cl[v] = cl2[cl[v]]; });  // mergeClusters (:159-163)
  *nC = nC2;
  be.free(cl2); be.free(cin); be.free(base);
  be.free(W.clus); be.free(W.nn); be.free(W.unused); be.free(W.nbc); be.free(W.newc); be.free(W.perm);
  be.free(W.sub_nc); be.free(W.cw); be.free(W.ewpc);
  free_subnets(be, S);
  free_net(be, red);
  return update;
}

// RunModularityClusteringCpp (RModularityOptimizer.cpp:21-173) on a built network.
// cluster_out: host array [n].  Returns the number of clusters; *modularity_out = the best start's value.
template <class B>
i32 run_clustering(B &be, const Net &net, int algorithm, double resolution, int modularityFunction, int nRandomStarts,
                   int nIterations, u64 seed, i32 *cluster_out, double *modularity_out, Stats *stats) {
  const i32 n = net.n;
  const double total_edge_weight = seq_sum(be, net.ew, net.m2, 0.0) / 2.0;  // getTotalEdgeWeight (:242-244)
  const double self_links = 0.0;  // matrixToNetwork drops nothing onto the diagonal
  const double resolution2 =
      modularityFunction == 1 ? resolution / (2 * total_edge_weight + self_links) : resolution;
  JavaRandom rng(seed);
  Optimizer<B> opt{be, resolution2, stats, rng};
  i32 *cl = be.template alloc<i32>(n), *best = be.template alloc<i32>(n);
  i32 bestC = 0;
  bool have = false;
  double maxQ = -INFINITY, q = 0;
  for (int s = 0; s < nRandomStarts && !be.failed(); ++s) {
    i32 nC = n;
    be.pfor(n, [=] MO_HD(i64 v) { cl[v] = (i32)v; });
    int j = 0;
    bool update = true;
    do {
      if (algorithm == 1)
        update = opt.louvain(net, cl, &nC, false);
      else if (algorithm == 2)
        update = opt.louvain(net, cl, &nC, true);
      else if (algorithm == 3)
        opt.slm(net, cl, &nC);  // the wrapper does not look at SLM's return value (:122-123)
      ++j;
    } while (j < nIterations && update && !be.failed());
    // only the value after a start's last iteration takes part in the comparison (:134-138)
    const double t0 = be.now_ms();
    q = quality_function(be, net, cl, nC, resolution2, total_edge_weight, self_links);
    if (stats) stats->ms_quality += be.now_ms() - t0;
    if (q > maxQ) {
      be.d2d(best, cl, sizeof(i32) * (size_t)n);
      bestC = nC;
      maxQ = q;
      have = true;
    }
  }
  if (!have) {  // "Clustering step failed." (RModularityOptimizer.cpp:149-151)
    be.free(cl);
    be.free(best);
    return -3;
  }
  // Clustering::orderClustersByNNodes (:130-157): stable, descending size; host work on nC counts
  std::vector<i32> h((size_t)n);
  be.d2h(h.data(), best, sizeof(i32) * (size_t)n);
  std::vector<i32> cnt((size_t)bestC, 0);
  for (i32 v = 0; v < n; ++v) cnt[h[v]]++;
  std::vector<std::pair<i32, i32>> order;
  order.reserve((size_t)bestC);
  for (i32 c = 0; c < bestC; ++c) order.push_back({cnt[c], c});
  std::stable_sort(order.begin(), order.end(),
                   [](const std::pair<i32, i32> &a, const std::pair<i32, i32> &b) { return b.first < a.first; });
  std::vector<i32> newc((size_t)bestC, 0);
  i32 k = 0;
  do {
    newc[order[k].second] = k;
    ++k;
  } while (k < bestC && order[k].first > 0);
  for (i32 v = 0; v < n; ++v) cluster_out[v] = newc[h[v]];
  if (modularity_out) *modularity_out = maxQ;
  be.free(cl);
  be.free(best);
  return k;
}

}  // namespace modopt
