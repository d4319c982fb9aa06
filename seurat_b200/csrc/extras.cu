// extras.cu -- SNN side consumers: the edge-file writer, fast_dist and
// SNN_SmallestNonzero_Dist behind the C ABI.
//
//   b200nn_write_edge_file           WriteEdgeFile, /root/reference/src/snn.cpp:40-59 (host I/O: formatting
//                                    and fwrite; the format is std::setprecision(15) = "%.15g")
//   b200nn_fast_dist                 fast_dist, /root/reference/src/fast_NN_dist.cpp:34-69
//   b200nn_snn_smallest_nonzero_dist SNN_SmallestNonzero_Dist, /root/reference/src/snn.cpp:82-126
//   b200nn_dev_embedding_prepare      what FindNeighbors.Seurat / .default do to the embedding on
//                                    the host before the search -- Embeddings(...)[, dims]
//                                    (/root/reference/R/clustering.R:843-846) and L2Norm
//                                    (R/clustering.R:619-622, R/utilities.R:2107-2122) -- on an embedding that
//                                    already lives in HBM
#include <string.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "internal.cuh"

#define EX_HD __host__ __device__
#include "extras_core.inl"

namespace {

// one thread per (row, listed neighbour) pair
__global__ void __launch_bounds__(256)
k_fast_dist(const double *__restrict__ x, const double *__restrict__ y, int d, const int64_t *__restrict__ ptr,
            const int32_t *__restrict__ nbr, int64_t nx, int64_t ny, int64_t total, double *__restrict__ out,
            int32_t *__restrict__ bad) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= total) return;
  int64_t lo = 0, hi = nx;  // row i with ptr[i] <= q < ptr[i + 1]
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (ptr[mid] <= q) lo = mid; else hi = mid;
  }
  const int64_t j = (int64_t)nbr[q] - 1;
  if (j < 0 || j >= ny) {
    *bad = 1;
    out[q] = nan("");
    return;
  }
  out[q] = extras::euclid_seq(x + lo * d, y + j * d, d);
}

__global__ void __launch_bounds__(128)
k_smallest_nonzero(const int64_t *__restrict__ p, const int32_t *__restrict__ ri, const double *__restrict__ x,
                   const double *__restrict__ mat, int d, int n_small, const double *__restrict__ nearest, int64_t n,
                   int32_t *__restrict__ ord, double *__restrict__ dist, double *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = extras::smallest_nonzero_dist_col(p, ri, x, mat, d, n_small, nearest, i, ord + p[i], dist + p[i]);
}

// One thread per cell: gather the selected columns of a row-major or column-major device matrix, optionally
// divide by the row's L2 norm (R's arithmetic, extras_core.inl), stage 128 rows in shared memory and write them
// out as one contiguous block of the row-major result.  STAGED = false: rows too wide for shared memory are
// written directly.
template <bool STAGED>
__global__ void __launch_bounds__(128)
k_embedding_prepare(const double *__restrict__ src, int64_t n, int n_cols, int col_major, const int32_t *__restrict__ dims,
                    int nd, int l2, double *__restrict__ out) {
  extern __shared__ double tile[];  // [128][nd | 1]
  const int64_t r0 = (int64_t)blockIdx.x * 128, row = r0 + threadIdx.x;
  const int ts = nd | 1;
  double *mine = STAGED ? tile + (size_t)threadIdx.x * ts : out + row * nd;
  if (row < n) {
    const int64_t rs = col_major ? 1 : n_cols, cs = col_major ? n : 1;
    for (int j = 0; j < nd; ++j) mine[j] = src[row * rs + (int64_t)(dims ? dims[j] : j) * cs];
    if (l2) {
      const double norm = extras::l2norm_of_row(mine, 1, nd);
      for (int j = 0; j < nd; ++j) mine[j] = extras::l2norm_entry(mine[j], norm);
    }
  }
  if (STAGED) {
    __syncthreads();
    const int64_t rows = n - r0 < 128 ? n - r0 : 128, total = rows * nd;
    for (int64_t t = threadIdx.x; t < total; t += 128) out[r0 * nd + t] = tile[(t / nd) * ts + (t % nd)];
  }
}

__global__ void k_widen_p(const int32_t *__restrict__ in, int64_t *__restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}

struct DevBufs {  // cudaMallocAsync'd scratch released on every exit path
  cudaStream_t st;
  std::vector<void *> ptrs;
  explicit DevBufs(cudaStream_t s) : st(s) {}
  ~DevBufs() {
    for (void *p : ptrs) cudaFreeAsync(p, st);
    cudaStreamSynchronize(st);
  }
  template <typename T>
  int get(size_t count, T **out) {
    void *p = nullptr;
    CU_TRY(cudaMallocAsync(&p, sizeof(T) * (count ? count : 1), st));
    ptrs.push_back(p);
    *out = static_cast<T *>(p);
    return B200NN_OK;
  }
};

// host matrix [rows x d] in `layout` -> device row-major
int upload_rowmajor(b200nn_ctx *ctx, DevBufs &B, const double *h, int64_t rows, int d, int layout, double **out) {
  double *dst;
  RC_TRY(B.get((size_t)rows * d, &dst));
  if (layout == B200NN_ROW_MAJOR) {
    CU_TRY(cudaMemcpyAsync(dst, h, sizeof(double) * (size_t)rows * d, cudaMemcpyHostToDevice, ctx->stream));
  } else {
    double *raw;
    RC_TRY(B.get((size_t)rows * d, &raw));
    CU_TRY(cudaMemcpyAsync(raw, h, sizeof(double) * (size_t)rows * d, cudaMemcpyHostToDevice, ctx->stream));
    RC_TRY(transpose_f64(ctx, raw, dst, rows, d));
  }
  *out = dst;
  return B200NN_OK;
}

}  // namespace

extern "C" {

int b200nn_write_edge_file(const int32_t *p, const int32_t *i, const double *x, int64_t n, const char *filename,
                           int64_t *n_written) {
  if (!p || !i || !x || !filename || n < 0) {
    b200nn_set_error("write_edge_file: invalid argument");
    return B200NN_ERR_INVALID;
  }
  FILE *f = fopen(filename, "w");
  if (!f) {
    b200nn_set_error("write_edge_file: cannot open %s", filename);
    return B200NN_ERR_INVALID;
  }
  // "%.15g" of every distinct value once (an SNN graph holds at most k + 1 of them)
  std::unordered_map<uint64_t, std::string> text;
  std::vector<char> buf;
  buf.reserve(1 << 22);
  int64_t lines = 0;
  char tmp[64];
  for (int64_t c = 0; c < n; ++c) {
    for (int64_t e = p[c]; e < p[c + 1]; ++e) {
      const int32_t r = i[e];
      if (c >= r) continue;  // it.col() >= it.row(): only the strict lower triangle is written
      uint64_t bits;
      memcpy(&bits, &x[e], sizeof(bits));
      auto it = text.find(bits);
      if (it == text.end()) {
        if (text.size() > (1u << 20)) text.clear();
        snprintf(tmp, sizeof(tmp), "%.15g", x[e]);
        it = text.emplace(bits, std::string(tmp)).first;
      }
      const int len = snprintf(tmp, sizeof(tmp), "%lld\t%d\t", (long long)c, r);
      buf.insert(buf.end(), tmp, tmp + len);
      buf.insert(buf.end(), it->second.begin(), it->second.end());
      buf.push_back('\n');
      ++lines;
      if (buf.size() > (1u << 22) - 128) {
        if (fwrite(buf.data(), 1, buf.size(), f) != buf.size()) {
          fclose(f);
          b200nn_set_error("write_edge_file: write to %s failed", filename);
          return B200NN_ERR_INVALID;
        }
        buf.clear();
      }
    }
  }
  const bool ok = fwrite(buf.data(), 1, buf.size(), f) == buf.size();
  if (fclose(f) != 0 || !ok) {
    b200nn_set_error("write_edge_file: write to %s failed", filename);
    return B200NN_ERR_INVALID;
  }
  if (n_written) *n_written = lines;
  return B200NN_OK;
}

int b200nn_fast_dist(b200nn_ctx *ctx, const double *x, int64_t nx, const double *y, int64_t ny, int32_t d,
                     int32_t layout, const int64_t *ptr, const int32_t *nbr, double *out) {
  if (!ctx) {
    b200nn_set_error("NULL context");
    return B200NN_ERR_INVALID;
  }
  if (!x || !y || !ptr || nx <= 0 || ny <= 0 || d <= 0 || ptr[0] != 0 || ptr[nx] < 0 || (ptr[nx] > 0 && (!nbr || !out))) {
    b200nn_set_error("fast_dist: invalid argument");
    return B200NN_ERR_INVALID;
  }
  const int64_t total = ptr[nx];
  if (total == 0) return B200NN_OK;
  CU_TRY(cudaSetDevice(ctx->device));
  DevBufs B(ctx->stream);
  double *dx, *dy, *dout;
  int64_t *dptr;
  int32_t *dnbr, *dbad;
  RC_TRY(upload_rowmajor(ctx, B, x, nx, d, layout, &dx));
  if (y == x && ny == nx) dy = dx; else RC_TRY(upload_rowmajor(ctx, B, y, ny, d, layout, &dy));
  RC_TRY(B.get((size_t)nx + 1, &dptr));
  RC_TRY(B.get((size_t)total, &dnbr));
  RC_TRY(B.get((size_t)total, &dout));
  RC_TRY(B.get(1, &dbad));
  CU_TRY(cudaMemcpyAsync(dptr, ptr, sizeof(int64_t) * ((size_t)nx + 1), cudaMemcpyHostToDevice, ctx->stream));
  CU_TRY(cudaMemcpyAsync(dnbr, nbr, sizeof(int32_t) * (size_t)total, cudaMemcpyHostToDevice, ctx->stream));
  CU_TRY(cudaMemsetAsync(dbad, 0, sizeof(int32_t), ctx->stream));
  k_fast_dist<<<(unsigned)ceil_div64(total, 256), 256, 0, ctx->stream>>>(dx, dy, d, dptr, dnbr, nx, ny, total, dout, dbad);
  KERNEL_CHECK(ctx);
  int32_t bad = 0;
  CU_TRY(cudaMemcpyAsync(out, dout, sizeof(double) * (size_t)total, cudaMemcpyDeviceToHost, ctx->stream));
  CU_TRY(cudaMemcpyAsync(&bad, dbad, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  if (bad) {
    b200nn_set_error("fast_dist: a neighbour index is outside [1, %lld]", (long long)ny);
    return B200NN_ERR_INDEX_RANGE;
  }
  return B200NN_OK;
}

int b200nn_snn_smallest_nonzero_dist(b200nn_ctx *ctx, const int32_t *p, const int32_t *i, const double *x, int64_t n,
                                     const double *mat, int32_t d, int32_t layout, int32_t n_small,
                                     const double *nearest_dist, double *out) {
  if (!ctx) {
    b200nn_set_error("NULL context");
    return B200NN_ERR_INVALID;
  }
  if (!p || !mat || !nearest_dist || !out || n <= 0 || d <= 0 || n_small < 0 || p[0] != 0 || p[n] < 0 ||
      (p[n] > 0 && (!i || !x))) {
    b200nn_set_error("snn_smallest_nonzero_dist: invalid argument");
    return B200NN_ERR_INVALID;
  }
  const int64_t nnz = p[n];
  for (int64_t e = 0; e < nnz; ++e)
    if (i[e] < 0 || i[e] >= n) {
      b200nn_set_error("snn_smallest_nonzero_dist: row index out of range");
      return B200NN_ERR_INDEX_RANGE;
    }
  CU_TRY(cudaSetDevice(ctx->device));
  DevBufs B(ctx->stream);
  double *dmat, *dx, *dnear, *dout, *ddist;
  int32_t *dp32, *di, *dord;
  int64_t *dp;
  RC_TRY(upload_rowmajor(ctx, B, mat, n, d, layout, &dmat));
  RC_TRY(B.get((size_t)n + 1, &dp32));
  RC_TRY(B.get((size_t)n + 1, &dp));
  RC_TRY(B.get((size_t)nnz, &di));
  RC_TRY(B.get((size_t)nnz, &dx));
  RC_TRY(B.get((size_t)nnz, &dord));
  RC_TRY(B.get((size_t)nnz, &ddist));
  RC_TRY(B.get((size_t)n, &dnear));
  RC_TRY(B.get((size_t)n, &dout));
  CU_TRY(cudaMemcpyAsync(dp32, p, sizeof(int32_t) * ((size_t)n + 1), cudaMemcpyHostToDevice, ctx->stream));
  if (nnz) {
    CU_TRY(cudaMemcpyAsync(di, i, sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(cudaMemcpyAsync(dx, x, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, ctx->stream));
  }
  CU_TRY(cudaMemcpyAsync(dnear, nearest_dist, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
  k_widen_p<<<(unsigned)ceil_div64(n + 1, 256), 256, 0, ctx->stream>>>(dp32, dp, n + 1);
  KERNEL_CHECK(ctx);
  k_smallest_nonzero<<<(unsigned)ceil_div64(n, 128), 128, 0, ctx->stream>>>(dp, di, dx, dmat, d, n_small, dnear, n, dord,
                                                                          ddist, dout);
  KERNEL_CHECK(ctx);
  CU_TRY(cudaMemcpyAsync(out, dout, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  return B200NN_OK;
}

int b200nn_dev_embedding_prepare(b200nn_ctx *ctx, const double *d_src, int64_t n, int32_t n_cols, int32_t src_layout,
                                 const int32_t *dims, int32_t n_dims, int32_t l2_norm, double *d_out) {
  if (!ctx) {
    b200nn_set_error("NULL context");
    return B200NN_ERR_INVALID;
  }
  if (!d_src || !d_out || n <= 0 || n_cols <= 0 || (dims && n_dims <= 0) || (!dims && n_dims != 0 && n_dims != n_cols) ||
      (src_layout != B200NN_ROW_MAJOR && src_layout != B200NN_COL_MAJOR)) {
    b200nn_set_error("dev_embedding_prepare: invalid argument");
    return B200NN_ERR_INVALID;
  }
  const int nd = dims ? n_dims : n_cols;
  for (int j = 0; dims && j < nd; ++j)
    if (dims[j] < 0 || dims[j] >= n_cols) {
      b200nn_set_error("dev_embedding_prepare: dimension %d is outside the embedding's %d columns (subscript out of bounds)",
                       dims[j] + 1, n_cols);
      return B200NN_ERR_INDEX_RANGE;
    }
  CU_TRY(cudaSetDevice(ctx->device));
  DevBufs B(ctx->stream);
  int32_t *d_dims = nullptr;
  if (dims) {
    RC_TRY(B.get((size_t)nd, &d_dims));
    CU_TRY(cudaMemcpyAsync(d_dims, dims, sizeof(int32_t) * (size_t)nd, cudaMemcpyHostToDevice, ctx->stream));
  }
  const size_t smem = sizeof(double) * 128 * (size_t)(nd | 1);
  const unsigned grid = (unsigned)ceil_div64(n, 128);
  if (smem <= 200 * 1024) {
    CU_TRY(cudaFuncSetAttribute(k_embedding_prepare<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_embedding_prepare<true><<<grid, 128, smem, ctx->stream>>>(d_src, n, n_cols, src_layout == B200NN_COL_MAJOR, d_dims,
                                                               nd, l2_norm != 0, d_out);
  } else {
    k_embedding_prepare<false><<<grid, 128, 0, ctx->stream>>>(d_src, n, n_cols, src_layout == B200NN_COL_MAJOR, d_dims, nd,
                                                             l2_norm != 0, d_out);
  }
  KERNEL_CHECK(ctx);
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  return B200NN_OK;
}

}  // extern "C"
