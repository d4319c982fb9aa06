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
