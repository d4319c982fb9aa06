// extras_emul.cpp -- host emulation of the per-row bodies in seurat_b200/csrc/extras_core.inl (the code
// the CUDA kernels of extras.cu run).  TEST INFRASTRUCTURE; never linked into libb200nn.so.
//   g++ -O2 -std=c++17 -ffp-contract=off -shared -fPIC tests/emul/extras_emul.cpp -o tests/emul/libextras_emul.so
#define EX_HD
#include "../../seurat_b200/csrc/extras_core.inl"

#include <vector>

extern "C" void emul_fast_dist(const double *x, int64_t nx, const double *y, int d, const int64_t *ptr,
                               const int32_t *nbr, double *out) {
  for (int64_t r = 0; r < nx; ++r)
    for (int64_t q = ptr[r]; q < ptr[r + 1]; ++q) out[q] = extras::euclid_seq(x + r * d, y + (int64_t)(nbr[q] - 1) * d, d);
}

extern "C" void emul_smallest_nonzero(const int64_t *p, const int32_t *ri, const double *x, int64_t n, const double *mat,
                                      int d, int n_small, const double *nearest, double *out) {
  std::vector<int32_t> ord((size_t)p[n] + 1);
  std::vector<double> dist((size_t)p[n] + 1);
  for (int64_t i = 0; i < n; ++i)
    out[i] = extras::smallest_nonzero_dist_col(p, ri, x, mat, d, n_small, nearest, i, ord.data() + p[i], dist.data() + p[i]);
}

// L2Norm of a row-major matrix with the integer restatement of the long-double sum (extras_core.inl) ...
extern "C" void emul_l2norm(const double *m, int64_t n, int d, double *out) {
  for (int64_t i = 0; i < n; ++i) {
    const double s = extras::l2norm_of_row(m + i * d, 1, d);
    for (int j = 0; j < d; ++j) out[i * d + j] = extras::l2norm_entry(m[i * d + j], s);
  }
}
// ... the total alone, and the same total from the compiler's own long double (x87 on x86-64): the pin
extern "C" double emul_x87_sum(const double *t, int64_t n) {
  extras::X87Sum s;
  extras::x87_init(s);
  for (int64_t i = 0; i < n; ++i) extras::x87_add(s, t[i]);
  return extras::x87_to_double(s);
}
extern "C" double native_long_double_sum(const double *t, int64_t n) {
  volatile long double acc = 0.0L;
  for (int64_t i = 0; i < n; ++i) acc = acc + (long double)t[i];
  return (double)acc;
}
extern "C" int native_long_double_mantissa_bits() { return __LDBL_MANT_DIG__; }
