// util.cu -- workspace, error strings, scan and re-layout kernels of libb200nn.
#include <stdarg.h>

#include "internal.cuh"

static thread_local char g_err[512] = "";

void b200nn_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char *b200nn_last_error(void) { return g_err; }

int ws_reserve(b200nn_ctx *ctx, WsSlot slot, size_t bytes, void **out) {
  WsBuf &b = ctx->ws[slot];
  if (bytes == 0) bytes = 16;
  if (b.cap < bytes) {
    if (b.p) {
      // earlier kernels on the stream may still read the old buffer
      CU_TRY(cudaStreamSynchronize(ctx->stream));
      CU_TRY(cudaFree(b.p));
      b.p = nullptr;
      b.cap = 0;
    }
    size_t want = (bytes + 255) & ~size_t(255);
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      b.p = nullptr;
      b200nn_set_error("device allocation of %zu bytes failed (workspace slot %d): %s", want,
                       (int)slot, cudaGetErrorString(e));
      return B200NN_ERR_NOMEM;
    }
    b.cap = want;
  }
  *out = b.p;
  return B200NN_OK;
}

// ---------------------------------------------------------------------------
// exclusive scan int32 -> int64 (reduce / scan block sums / down-sweep).
// out has n + 1 entries; out[n] is the total.
// ---------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;  // per thread
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int64_t warp_incl_scan(int64_t v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int64_t t = __shfl_up_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) >= o) v += t;
  }
  return v;
}

// block-wide exclusive scan of one value per thread; returns the exclusive
// prefix and writes the block total to *total (same for every thread)
__device__ __forceinline__ int64_t block_excl_scan(int64_t v, int64_t *total) {
  __shared__ int64_t wsum[33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  int64_t inc = warp_incl_scan(v);
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  if (w == 0) {
    int64_t s = lane < nw ? wsum[lane] : 0;
    int64_t si = warp_incl_scan(s);
    wsum[lane] = si - s;              // exclusive offset of warp `lane`
    if (lane == 31) wsum[32] = si;    // block total
  }
  __syncthreads();
  int64_t r = wsum[w] + inc - v;
  *total = wsum[32];
  __syncthreads();  // wsum is reused by the next call
  return r;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const int32_t *__restrict__ in,
                                                                int64_t n,
                                                                int64_t *__restrict__ bsum) {
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
  int64_t s = 0;
  for (int it = 0; it < SCAN_ITEMS; ++it) {
    int64_t idx = base + (int64_t)it * SCAN_THREADS + threadIdx.x;
    if (idx < n) s += in[idx];
  }
  int64_t total;
  block_excl_scan(s, &total);
  if (threadIdx.x == 0) bsum[blockIdx.x] = total;
}

// single block: in-place exclusive scan of bsum[0..nb), total -> *grand
__global__ void __launch_bounds__(1024) k_scan_bsums(int64_t *__restrict__ bsum, int64_t nb,
                                                       int64_t *__restrict__ grand) {
  int64_t carry = 0;
  for (int64_t base = 0; base < nb; base += blockDim.x) {
    int64_t idx = base + threadIdx.x;
    int64_t v = idx < nb ? bsum[idx] : 0;
    int64_t total;
    int64_t ex = block_excl_scan(v, &total);
    if (idx < nb) bsum[idx] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) *grand = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_down(const int32_t *__restrict__ in,
                                                              int64_t n,
                                                              const int64_t *__restrict__ bsum,
                                                              int64_t *__restrict__ out) {
  // thread t owns SCAN_ITEMS consecutive items so that a plain running sum works
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int32_t v[SCAN_ITEMS];
  int64_t s = 0;
#pragma unroll
  for (int it = 0; it < SCAN_ITEMS; ++it) {
    int64_t idx = base + it;
    v[it] = idx < n ? in[idx] : 0;
    s += v[it];
  }
  int64_t total;
  int64_t ex = block_excl_scan(s, &total) + bsum[blockIdx.x];
#pragma unroll
  for (int it = 0; it < SCAN_ITEMS; ++it) {
    int64_t idx = base + it;
    if (idx < n) out[idx] = ex;
    ex += v[it];
  }
}

int scan_exclusive_i32_to_i64(b200nn_ctx *ctx, const int32_t *d_in, int64_t *d_out, int64_t n,
                              cudaStream_t stream, int tmp_slot) {
  // a scan queued on another stream than the context's needs its own block-sum scratch
  cudaStream_t st = stream ? stream : ctx->stream;
  int64_t nb = ceil_div64(n > 0 ? n : 1, SCAN_TILE);
  int64_t *bsum;
  RC_TRY(ws_get(ctx, tmp_slot >= 0 ? (WsSlot)tmp_slot : WS_SCAN_TMP, (size_t)nb + 1, &bsum));
  k_scan_reduce<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(d_in, n, bsum);
  KERNEL_CHECK(ctx);
  k_scan_bsums<<<1, 1024, 0, st>>>(bsum, nb, d_out + n);
  KERNEL_CHECK(ctx);
  k_scan_down<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(d_in, n, bsum, d_out);
  KERNEL_CHECK(ctx);
  return B200NN_OK;
}

// ---------------------------------------------------------------------------
// re-layout kernels
// ---------------------------------------------------------------------------
// in: column-major [nrow x ncol] (element (r,c) at c*nrow + r)
// out: row-major   [nrow x ncol] (element (r,c) at r*ncol + c)
template <typename T>
__global__ void __launch_bounds__(256) k_transpose(const T *__restrict__ in, T *__restrict__ out,
                                                   int64_t nrow, int64_t ncol, T add,
                                                   int64_t gx) {
  __shared__ T tile[32][33];
  int64_t r0 = ((int64_t)blockIdx.x % gx) * 32, c0 = ((int64_t)blockIdx.x / gx) * 32;
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    int64_t r = r0 + tx, c = c0 + j;
    if (r < nrow && c < ncol) tile[j][tx] = in[c * nrow + r];
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    int64_t r = r0 + j, c = c0 + tx;
    if (r < nrow && c < ncol) out[r * ncol + c] = tile[tx][j] + add;
  }
}

template <typename T>
static int transpose_any(b200nn_ctx *ctx, const T *d_in, T *d_out, int64_t nrow, int64_t ncol,
                         T add) {
  if (nrow <= 0 || ncol <= 0) return B200NN_OK;
  int64_t gx = ceil_div64(nrow, 32), gy = ceil_div64(ncol, 32);
  if (gx * gy > 0x7fffffffLL) {
    b200nn_set_error("matrix too large for the re-layout kernel (%lld x %lld)", (long long)nrow,
                     (long long)ncol);
    return B200NN_ERR_INVALID;
  }
  k_transpose<T><<<(unsigned)(gx * gy), 256, 0, ctx->stream>>>(d_in, d_out, nrow, ncol, add, gx);
  KERNEL_CHECK(ctx);
  return B200NN_OK;
}

int transpose_f64(b200nn_ctx *ctx, const double *d_in, double *d_out, int64_t nrow, int64_t ncol) {
  return transpose_any<double>(ctx, d_in, d_out, nrow, ncol, 0.0);
}

int transpose_i32_add(b200nn_ctx *ctx, const int32_t *d_in, int32_t *d_out, int64_t nrow,
                      int64_t ncol, int32_t add) {
  return transpose_any<int32_t>(ctx, d_in, d_out, nrow, ncol, add);
}

__global__ void k_add_i32(const int32_t *__restrict__ in, int32_t *__restrict__ out, int64_t n,
                          int32_t add) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] + add;
}

int add_i32(b200nn_ctx *ctx, const int32_t *d_in, int32_t *d_out, int64_t n, int32_t add) {
  if (n <= 0) return B200NN_OK;
  k_add_i32<<<(unsigned)ceil_div64(n, 256), 256, 0, ctx->stream>>>(d_in, d_out, n, add);
  KERNEL_CHECK(ctx);
  return B200NN_OK;
}

__global__ void k_narrow(const int64_t *__restrict__ in, int32_t *__restrict__ out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int32_t)in[i];
}

int narrow_i64_to_i32(b200nn_ctx *ctx, const int64_t *d_in, int32_t *d_out, int64_t n, cudaStream_t stream) {
  if (n <= 0) return B200NN_OK;
  k_narrow<<<(unsigned)ceil_div64(n, 256), 256, 0, stream ? stream : ctx->stream>>>(d_in, d_out, n);
  KERNEL_CHECK(ctx);
  return B200NN_OK;
}

// ---------------------------------------------------------------------------
// RANN::nn2 goes through .C(..., NAOK = FALSE): R refuses NA / NaN / Inf before the search
// starts.  The tensor path folds this test into its norm pass; the FP64 scan calls this.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_check_finite(const double *__restrict__ x, int64_t n, int32_t *flag) {
  int bad = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    bad |= !isfinite(x[i]);
  if (bad) *flag = 1;
}

int check_finite_f64(b200nn_ctx *ctx, const double *d_x, int64_t count, const char *what) {
  if (count <= 0) return B200NN_OK;
  int32_t *flag;
  RC_TRY(ws_get(ctx, WS_FLAG, 4, &flag));
  CU_TRY(cudaMemsetAsync(flag, 0, sizeof(int32_t), ctx->stream));
  int64_t blocks = ceil_div64(count, 256 * 8);
  if (blocks > (int64_t)ctx->sm_count * 16) blocks = (int64_t)ctx->sm_count * 16;
  k_check_finite<<<(unsigned)blocks, 256, 0, ctx->stream>>>(d_x, count, flag);
  KERNEL_CHECK(ctx);
  int32_t h = 0;
  CU_TRY(cudaMemcpyAsync(&h, flag, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  if (h) {
    b200nn_set_error("NA/NaN/Inf in foreign function call (arg %s)", what);
    return B200NN_ERR_INVALID;
  }
  return B200NN_OK;
}
