// extras_core.inl -- per-row bodies of the SNN side consumers, shared by extras.cu
// (device) and tests/emul/extras_emul.cpp (host emulation, test infrastructure).
//
//   fast_dist                  /root/reference/src/fast_NN_dist.cpp:34-69 (euclidean_distance :8-31)
//   SNN_SmallestNonzero_Dist   /root/reference/src/snn.cpp:82-126 (sort_indexes :70-79)
#pragma once
#include <math.h>
#include <stdint.h>

namespace extras {

// fast_NN_dist.cpp:8-31: rval += pow(d1 - d2, 2) in dimension order, sqrt at the end.
// pow(t, 2) is the correctly rounded t * t; no contraction into FMA.
EX_HD inline double euclid_seq(const double *a, const double *b, int d) {
  double r = 0.0;
  for (int j = 0; j < d; ++j) {
#ifdef __CUDA_ARCH__
    const double t = __dsub_rn(a[j], b[j]);
    r = __dadd_rn(r, __dmul_rn(t, t));
#else
    const double t = a[j] - b[j];
    r = r + t * t;
#endif
  }
#ifdef __CUDA_ARCH__
  return __dsqrt_rn(r);
#else
  return sqrt(r);
#endif
}

// One column i of SNN_SmallestNonzero_Dist.  ord / dist: scratch of the column's nnz entries.
//   p, ri, x : dgCMatrix slots; mat [n_cells x d] row-major; n_small: the `n` argument.
// The distance of two rows is (mat.row(cell) - mat.row(i)).norm() in the reference: Eigen's
// reduction order is unspecified (vectorised), here the squares are summed in dimension order --
// results agree to rounding (documented tolerance 1e-12 relative), not necessarily bit for bit.
EX_HD inline double smallest_nonzero_dist_col(const int64_t *p, const int32_t *ri, const double *x, const double *mat,
                                              int d, int n_small, const double *nearest, int64_t i, int32_t *ord,
                                              double *dist) {
  const int64_t b = p[i], cnt = p[i + 1] - b;
  // sort_indexes: stable ascending by value (insertion sort keeps equal values in storage order)
  for (int64_t a = 0; a < cnt; ++a) {
    const double v = x[b + a];
    int64_t q = a;
    while (q > 0 && x[b + ord[q - 1]] > v) {
      ord[q] = ord[q - 1];
      --q;
    }
    ord[q] = (int32_t)a;
  }
  int64_t n_i = n_small;
  if (n_i > cnt) n_i = cnt;
  int64_t nd = 0;
  for (int64_t j = 0; j < cnt; ++j) {
    if (nd < n_i || x[b + ord[j]] == x[b + ord[n_i - 1]]) {
      const int64_t cell = ri[b + ord[j]];
      double res = euclid_seq(mat + cell * d, mat + i * d, d);
      if (nearest[i] > 0) {
        res = res - nearest[i];
        if (res < 0) res = 0;
      }
      dist[nd++] = res;
    } else {
      break;
    }
  }
  double acc = 0.0;
  if (nd > n_i) {
    // std::sort(rbegin, rend): descending; the n_i largest are averaged
    for (int64_t a = 1; a < nd; ++a) {
      const double v = dist[a];
      int64_t q = a;
      while (q > 0 && dist[q - 1] < v) {
        dist[q] = dist[q - 1];
        --q;
      }
      dist[q] = v;
    }
    for (int64_t a = 0; a < n_i; ++a) acc = acc + dist[a];
    return acc / (double)n_i;
  }
  for (int64_t a = 0; a < nd; ++a) acc = acc + dist[a];
  return acc / (double)nd;  // 0 / 0 = NaN for an empty column, as in the reference
}

}  // namespace extras
