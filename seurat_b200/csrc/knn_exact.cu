// knn_exact.cu -- FP64 exhaustive-scan kNN (reference arithmetic).
//
// Replaces RANN::nn2 (call site /root/reference/R/clustering.R:1664-1667) with
// the arithmetic the oracle pins: d2 = sum_j (q_j - p_j)^2 accumulated in
// dimension order with one rounding per product and per sum (no FMA), ordered
// by (d2, index), sqrt at the end.  It is the parity anchor, the path for
// shapes the tensor-core kernel does not cover, and the fallback for query
// rows whose tensor-core candidate set could not be certified exact.
//
// Work layout: a CTA owns QB query rows; the reference set streams through
// shared memory in tiles of 128 rows; thread t computes the distances of tile
// row t to all QB queries (QB independent FP64 accumulators).  Rows that beat
// the running k-th best of a query are appended to that query's candidate
// buffer and merged into its sorted top-k list by one warp after every tile.
#include <math_constants.h>

#include "internal.cuh"

namespace {

constexpr int KX_THREADS = 128;    // threads per CTA == reference rows per tile
constexpr int KX_DC = 64;          // dimensions staged per chunk
constexpr int KX_RS = KX_DC + 1;   // padded row stride of the reference tile (doubles)

__host__ __device__ inline size_t kx_smem_bytes(int QB, int k) {
  size_t b = 0;
  b += sizeof(double) * (size_t)QB * KX_DC;        // q_s
  b += sizeof(double) * (size_t)KX_THREADS * KX_RS; // r_s
  b += sizeof(double) * (size_t)QB * k;            // top_d
  b += sizeof(double) * (size_t)QB * KX_THREADS;   // buf_d
  b += sizeof(double) * (size_t)QB;                // tau
  b += sizeof(int64_t) * (size_t)QB;               // qid
  b += sizeof(int32_t) * (size_t)QB * k;           // top_i
  b += sizeof(int32_t) * (size_t)QB * KX_THREADS;  // buf_i
  b += sizeof(int32_t) * (size_t)QB * 2;           // buf_n, top_n
  return b + 16;
}

// (d, i) < (e, j) in the order the oracle sorts by
__device__ __forceinline__ bool lex_less(double d, int32_t i, double e, int32_t j) {
  return d < e || (d == e && i < j);
}

template <int QB>
__global__ void __launch_bounds__(KX_THREADS)
k_knn_exact(const double *__restrict__ ref, int64_t nr, const double *__restrict__ qry, int d, int k,
            const int32_t *__restrict__ qsel, const int32_t *__restrict__ n_sel_ptr,
            int64_t n_items, int32_t *__restrict__ idx0, double *__restrict__ dist) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *q_s = reinterpret_cast<double *>(smem_raw);
  double *r_s = q_s + QB * KX_DC;
  double *top_d = r_s + KX_THREADS * KX_RS;
  double *buf_d = top_d + QB * k;
  double *tau = buf_d + QB * KX_THREADS;
  int64_t *qid = reinterpret_cast<int64_t *>(tau + QB);
  int32_t *top_i = reinterpret_cast<int32_t *>(qid + QB);
  int32_t *buf_i = top_i + QB * k;
  int32_t *buf_n = buf_i + QB * KX_THREADS;
  int32_t *top_n = buf_n + QB;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int64_t n_it = n_items;
  if (n_sel_ptr != nullptr) {
    int64_t c = *n_sel_ptr;
    if (c < n_it) n_it = c;
  }
  const int64_t item0 = (int64_t)blockIdx.x * QB;
  if (item0 >= n_it) return;

  if (tid < QB) {
    int64_t item = item0 + tid;
    qid[tid] = item < n_it ? (qsel != nullptr ? (int64_t)qsel[item] : item) : -1;
    top_n[tid] = 0;
    buf_n[tid] = 0;
    tau[tid] = CUDART_INF;
  }
  __syncthreads();

  const bool single_chunk = d <= KX_DC;
  if (single_chunk) {
    for (int e = tid; e < QB * d; e += KX_THREADS) {
      int qi = e / d, j = e - qi * d;
      q_s[qi * KX_DC + j] = qid[qi] >= 0 ? qry[qid[qi] * d + j] : 0.0;
    }
  }

  for (int64_t tile = 0; tile < nr; tile += KX_THREADS) {
    const int rows_here = (int)((nr - tile) < KX_THREADS ? (nr - tile) : KX_THREADS);
    double acc[QB];
#pragma unroll
    for (int qi = 0; qi < QB; ++qi) acc[qi] = 0.0;

    for (int c0 = 0; c0 < d; c0 += KX_DC) {
      const int dc = (d - c0) < KX_DC ? (d - c0) : KX_DC;
      __syncthreads();  // everyone is done with the previous contents of r_s / q_s
      if (!single_chunk) {
        for (int e = tid; e < QB * dc; e += KX_THREADS) {
          int qi = e / dc, j = e - qi * dc;
          q_s[qi * KX_DC + j] = qid[qi] >= 0 ? qry[qid[qi] * d + c0 + j] : 0.0;
        }
      }
      for (int e = tid; e < rows_here * dc; e += KX_THREADS) {
        int rl = e / dc, j = e - rl * dc;
        r_s[rl * KX_RS + j] = ref[(tile + rl) * d + c0 + j];
      }
      __syncthreads();
      if (tid < rows_here) {
        const double *rr = r_s + tid * KX_RS;
        for (int j = 0; j < dc; ++j) {
          const double rj = rr[j];
#pragma unroll
          for (int qi = 0; qi < QB; ++qi) {
            // t = q - p; dist = dist + t*t  -- two roundings, never an FMA
            double t = __dsub_rn(q_s[qi * KX_DC + j], rj);
            acc[qi] = __dadd_rn(acc[qi], __dmul_rn(t, t));
          }
        }
      }
    }

    if (tid < rows_here) {
      const int32_t row = (int32_t)(tile + tid);
#pragma unroll
      for (int qi = 0; qi < QB; ++qi) {
        if (qid[qi] >= 0 && acc[qi] <= tau[qi]) {
          int pos = atomicAdd(&buf_n[qi], 1);
          buf_d[qi * KX_THREADS + pos] = acc[qi];
          buf_i[qi * KX_THREADS + pos] = row;
        }
      }
    }
    __syncthreads();

    // one warp merges the buffer of a query into its sorted top-k list
    for (int qi = warp; qi < QB; qi += KX_THREADS / 32) {
      const int m = buf_n[qi];
      if (m == 0) continue;
      double *td = top_d + qi * k;
      int32_t *ti = top_i + qi * k;
      int tn = top_n[qi];
      for (int b = 0; b < m; ++b) {
        const double dv = buf_d[qi * KX_THREADS + b];
        const int32_t iv = buf_i[qi * KX_THREADS + b];
        if (tn == k && !lex_less(dv, iv, td[k - 1], ti[k - 1])) continue;
        int cnt = 0;
        for (int t = lane; t < tn; t += 32) cnt += lex_less(td[t], ti[t], dv, iv) ? 1 : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        const int pos = cnt;
        const int newn = tn < k ? tn + 1 : k;
        // shift [pos, newn-1) one slot to the right, highest chunk first
        for (int base = newn - 1; base > pos; base -= 32) {
          const int t = base - lane;
          const bool act = t > pos;
          double sd = 0.0;
          int32_t si = 0;
          if (act) {
            sd = td[t - 1];
            si = ti[t - 1];
          }
          __syncwarp();
          if (act) {
            td[t] = sd;
            ti[t] = si;
          }
          __syncwarp();
        }
        if (lane == 0) {
          td[pos] = dv;
          ti[pos] = iv;
        }
        __syncwarp();
        tn = newn;
      }
      if (lane == 0) {
        top_n[qi] = tn;
        buf_n[qi] = 0;
        tau[qi] = tn == k ? td[k - 1] : CUDART_INF;
      }
    }
    // the __syncthreads at the top of the next tile orders these writes
  }
  __syncthreads();

  for (int qi = 0; qi < QB; ++qi) {
    const int64_t q = qid[qi];
    if (q < 0) continue;
    const int tn = top_n[qi];
    for (int t = tid; t < k; t += KX_THREADS) {
      if (t < tn) {
        idx0[q * k + t] = top_i[qi * k + t];
        dist[q * k + t] = sqrt(top_d[qi * k + t]);
      } else {  // only reachable with non-finite input
        idx0[q * k + t] = -1;
        dist[q * k + t] = CUDART_NAN;
      }
    }
  }
}

template <int QB>
int launch_qb(b200nn_ctx *ctx, const double *d_ref, int64_t nr, const double *d_qry, int d, int k,
              const int32_t *d_qsel, const int32_t *d_nsel, int64_t n_items, int32_t *d_idx0,
              double *d_dist) {
  size_t smem = kx_smem_bytes(QB, k);
  if (smem > 227 * 1024) {
    b200nn_set_error("k = %d is too large for the exact kNN kernel (shared memory %zu B)", k, smem);
    return B200NN_ERR_INVALID;
  }
  CU_TRY(cudaFuncSetAttribute(k_knn_exact<QB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)smem));
  int64_t blocks = ceil_div64(n_items, QB);
  if (blocks <= 0) return B200NN_OK;
  k_knn_exact<QB><<<(unsigned)blocks, KX_THREADS, smem, ctx->stream>>>(
      d_ref, nr, d_qry, d, k, d_qsel, d_nsel, n_items, d_idx0, d_dist);
  KERNEL_CHECK(ctx);
  return B200NN_OK;
}

}  // namespace

int knn_exact_launch(b200nn_ctx *ctx, const double *d_ref, int64_t nr, const double *d_qry,
                     int64_t nq, int d, int k, const int32_t *d_qsel, int64_t n_sel,
                     int32_t *d_idx0, double *d_dist) {
  // d_qsel != nullptr: process the rows listed in d_qsel; their count lives on
  // the device right behind the list's capacity is unknown to the host, so the
  // caller passes the device counter through d_qsel[-1] convention instead:
  // the counter is the int32 just before the list.
  const int32_t *d_nsel = d_qsel ? d_qsel - 1 : nullptr;
  int64_t n_items = d_qsel ? n_sel : nq;
  if (n_items <= 0) return B200NN_OK;
  if (kx_smem_bytes(16, k) <= 110 * 1024 && d_qsel == nullptr)
    return launch_qb<16>(ctx, d_ref, nr, d_qry, d, k, d_qsel, d_nsel, n_items, d_idx0, d_dist);
  if (kx_smem_bytes(4, k) <= 227 * 1024)
    return launch_qb<4>(ctx, d_ref, nr, d_qry, d, k, d_qsel, d_nsel, n_items, d_idx0, d_dist);
  return launch_qb<1>(ctx, d_ref, nr, d_qry, d, k, d_qsel, d_nsel, n_items, d_idx0, d_dist);
}
