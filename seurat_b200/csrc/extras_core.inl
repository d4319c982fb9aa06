// extras_core.inl -- per-row bodies of the SNN side consumers, shared by extras.cu
// (device) and tests/emul/extras_emul.cpp (host emulation, test infrastructure).
//
//   fast_dist                  /root/reference/src/fast_NN_dist.cpp:34-69 (euclidean_distance :8-31)
//   SNN_SmallestNonzero_Dist   /root/reference/src/snn.cpp:82-126 (sort_indexes :70-79)
//   L2Norm                     /root/reference/R/utilities.R:2107-2122 (the embedding prepared on
//                              the device -- dims slicing, l2.norm -- instead of on the host)
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

// ---- L2Norm on the device, bit for bit ---------------------------------------------------------
// L2Norm divides every row by sqrt(sum(x^2)).  R's sum() accumulates the doubles x^2 in LONG DOUBLE
// (x87 extended precision on x86: 64-bit significand, round to nearest even, one rounding per
// addition) and rounds the total to double once.  The device has no such type, so the accumulation
// is restated with integer arithmetic: X87Sum adds NON-NEGATIVE doubles exactly as the x87 does.
struct X87Sum {
  uint64_t m;    // significand, top bit set unless the sum is zero; value = m * 2^e
  int e;
  bool bad;      // a term was +inf or NaN: the total is not finite
};

EX_HD inline int x87_clz64(uint64_t v) {
#ifdef __CUDA_ARCH__
  return __clzll((long long)v);
#else
  return __builtin_clzll(v);
#endif
}

EX_HD inline uint64_t x87_bits(double v) {
#ifdef __CUDA_ARCH__
  return (uint64_t)__double_as_longlong(v);
#else
  uint64_t b;
  __builtin_memcpy(&b, &v, sizeof(b));
  return b;
#endif
}

EX_HD inline void x87_init(X87Sum &s) {
  s.m = 0;
  s.e = 0;
  s.bad = false;
}

// s += t for t >= 0 (a square).  One rounding to 64 significant bits, ties to even.
EX_HD inline void x87_add(X87Sum &s, double t) {
  const uint64_t bits = x87_bits(t);
  const int ex = (int)((bits >> 52) & 0x7ff);
  uint64_t mt = bits & 0xfffffffffffffull;
  if (ex == 0x7ff || (bits >> 63)) {  // inf / NaN (or a negative term, which a square never is)
    if (ex == 0x7ff || mt != 0 || ex != 0) s.bad = true;
    return;
  }
  int et;
  if (ex == 0) {
    if (mt == 0) return;  // + 0
    et = -1074;
  } else {
    mt |= 1ull << 52;
    et = ex - 1075;
  }
  const int sh0 = x87_clz64(mt);
  mt <<= sh0;
  et -= sh0;
  if (s.m == 0) {
    s.m = mt;
    s.e = et;
    return;
  }
  uint64_t ma = s.m, mb = mt;
  int ea = s.e, eb = et;
  if (eb > ea) {
    ma = mt; ea = et;
    mb = s.m; eb = s.e;
  }
  const int sh = ea - eb;
  // 128-bit window: a = ma:0, b = (mb:0) >> sh with a sticky bit for what falls below the window
  uint64_t bhi, blo;
  bool sticky = false;
  if (sh == 0) {
    bhi = mb; blo = 0;
  } else if (sh < 64) {
    bhi = mb >> sh; blo = mb << (64 - sh);
  } else if (sh == 64) {
    bhi = 0; blo = mb;
  } else if (sh < 128) {
    bhi = 0; blo = mb >> (sh - 64);
    sticky = (mb << (128 - sh)) != 0;
  } else {
    bhi = 0; blo = 0;
    sticky = true;
  }
  uint64_t hi = ma + bhi, lo = blo;
  if (hi < ma) {  // carry out of the window: one more significant bit
    sticky = sticky || (lo & 1ull);
    lo = (lo >> 1) | (hi << 63);
    hi = (hi >> 1) | (1ull << 63);
    ea += 1;
  }
  const uint64_t half = 1ull << 63;
  if (lo > half || (lo == half && (sticky || (hi & 1ull)))) {
    hi += 1;
    if (hi == 0) {
      hi = half;
      ea += 1;
    }
  }
  s.m = hi;
  s.e = ea;
}

// the total as a double: one more rounding (ties to even), like the (double) of a long double
EX_HD inline double x87_to_double(const X87Sum &s) {
  if (s.bad) return INFINITY;
  if (s.m == 0) return 0.0;
  const int E = s.e + 63;  // exponent of the leading bit
  int drop = 11;
  if (E < -1022) drop += -1022 - E;  // subnormal result: coarser quantum
  uint64_t r;
  if (drop >= 65) {
    r = 0;
  } else if (drop == 64) {
    r = s.m > (1ull << 63) ? 1 : 0;
  } else {
    const uint64_t keep = s.m >> drop, rem = s.m & ((1ull << drop) - 1), half = 1ull << (drop - 1);
    r = keep + ((rem > half || (rem == half && (keep & 1ull))) ? 1 : 0);
  }
  return ldexp((double)r, s.e + drop);  // exact: r <= 2^53; overflows to +inf like the x87 conversion
}

// sqrt(sum(x^2)) of `nd` values at x[0], x[stride], ... -- the STATS of L2Norm's Sweep
EX_HD inline double l2norm_of_row(const double *x, int64_t stride, int nd) {
  X87Sum acc;
  x87_init(acc);
  for (int j = 0; j < nd; ++j) {
    const double v = x[(int64_t)j * stride];
#ifdef __CUDA_ARCH__
    x87_add(acc, __dmul_rn(v, v));
#else
    x87_add(acc, v * v);
#endif
  }
#ifdef __CUDA_ARCH__
  return __dsqrt_rn(x87_to_double(acc));
#else
  return sqrt(x87_to_double(acc));
#endif
}

// one entry of the normalised matrix: x / norm, non-finite results become 0 (R/utilities.R:2120)
EX_HD inline double l2norm_entry(double v, double norm) {
#ifdef __CUDA_ARCH__
  const double q = __ddiv_rn(v, norm);
#else
  const double q = v / norm;
#endif
  return isfinite(q) ? q : 0.0;
}

}  // namespace extras
