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
#include <cuda_fp16.h>
#include <math_constants.h>

#include <algorithm>

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
  b += sizeof(double) * (size_t)QB * 2;            // tau, tau0
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

// one warp merges m buffered (d2, row) pairs into a sorted list of at most k; returns its length
__device__ __forceinline__ int topk_merge(double *td, int32_t *ti, int tn, int k, const double *bd,
                                          const int32_t *bi, int m, int lane) {
  for (int b = 0; b < m; ++b) {
    const double dv = bd[b];
    const int32_t iv = bi[b];
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
  return tn;
}

template <int QB>
__global__ void __launch_bounds__(KX_THREADS)
k_knn_exact(const double *__restrict__ ref, int64_t nr, const double *__restrict__ qry, int d, int k,
            const int32_t *__restrict__ qsel, const int32_t *__restrict__ n_sel_ptr,
            int64_t n_items, int32_t *__restrict__ idx0, double *__restrict__ dist,
            int64_t slice_len, int n_slices, int32_t *__restrict__ part_idx,
            double *__restrict__ part_d2, const double *__restrict__ tau0_in) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *q_s = reinterpret_cast<double *>(smem_raw);
  double *r_s = q_s + QB * KX_DC;
  double *top_d = r_s + KX_THREADS * KX_RS;
  double *buf_d = top_d + QB * k;
  double *tau = buf_d + QB * KX_THREADS;
  double *tau0 = tau + QB;  // caller's upper bound of the k-th squared distance (+inf: none)
  int64_t *qid = reinterpret_cast<int64_t *>(tau0 + QB);
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
    tau0[tid] = (tau0_in != nullptr && item < n_it) ? tau0_in[item] : CUDART_INF;
    tau[tid] = tau0[tid];
  }
  __syncthreads();

  const bool single_chunk = d <= KX_DC;
  if (single_chunk) {
    for (int e = tid; e < QB * d; e += KX_THREADS) {
      int qi = e / d, j = e - qi * d;
      q_s[qi * KX_DC + j] = qid[qi] >= 0 ? qry[qid[qi] * d + j] : 0.0;
    }
  }

  // blockIdx.y selects a slice of the reference set (split scan of a few fallback rows)
  const int64_t r_lo = (int64_t)blockIdx.y * slice_len;
  const int64_t r_hi = (r_lo + slice_len) < nr ? (r_lo + slice_len) : nr;
  for (int64_t tile = r_lo; tile < r_hi; tile += KX_THREADS) {
    const int rows_here = (int)((r_hi - tile) < KX_THREADS ? (r_hi - tile) : KX_THREADS);
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
      tn = topk_merge(td, ti, tn, k, buf_d + qi * KX_THREADS, buf_i + qi * KX_THREADS, m, lane);
      if (lane == 0) {
        top_n[qi] = tn;
        buf_n[qi] = 0;
        tau[qi] = tn == k ? td[k - 1] : tau0[qi];
      }
    }
    // the __syncthreads at the top of the next tile orders these writes
  }
  __syncthreads();

  if (part_idx != nullptr) {
    // split scan: this slice's sorted partial result (squared distances) for the merge kernel
    for (int qi = 0; qi < QB; ++qi) {
      if (qid[qi] < 0) continue;
      const int64_t item = item0 + qi;
      const int tn = top_n[qi];
      const int64_t o = (item * n_slices + blockIdx.y) * k;
      for (int t = tid; t < k; t += KX_THREADS) {
        part_idx[o + t] = t < tn ? top_i[qi * k + t] : 0x7fffffff;
        part_d2[o + t] = t < tn ? top_d[qi * k + t] : CUDART_INF;
      }
    }
    return;
  }
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

// ---------------------------------------------------------------------------
// Fallback rows of the tensor-core path: the same scan restricted to the groups of
// reference rows that can hold a point within the row's bound.  One CTA per (row, slice).
//   |x_q - x_r| = |y_q - y_r| / s >= (|yh_q - c_g| - rho_g - e_q - e_max) / s   for r in group g
// so group g is skipped when that exceeds sqrt(tau0) (tau0 >= the true k-th squared distance,
// obtained from a real candidate in reference arithmetic).
// ---------------------------------------------------------------------------
constexpr int KG_MAXG = 1024;

__device__ __forceinline__ void kx_cp_async8(void *s, const void *g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(s)), "l"(g)
               : "memory");
}

__host__ __device__ inline size_t kg_smem_bytes(int k) {
  size_t b = 0;
  b += sizeof(double) * KX_DC;                      // q_s
  b += sizeof(double) * KX_DC;                      // yq
  b += sizeof(double) * (size_t)KX_THREADS * KX_RS; // r_s
  b += sizeof(double) * (size_t)k;                  // top_d
  b += sizeof(double) * KX_THREADS;                 // buf_d
  b += sizeof(int32_t) * (size_t)k;                 // top_i
  b += sizeof(int32_t) * KX_THREADS;                // buf_i
  b += sizeof(int32_t) * KX_THREADS;                // row_s
  b += sizeof(int32_t) * KG_MAXG;                   // sel
  return b + 64;
}

__global__ void __launch_bounds__(KX_THREADS)
k_knn_exact_grouped(const double *__restrict__ ref, const double *__restrict__ qry, int d, int k,
                    const int32_t *__restrict__ qsel, const int32_t *__restrict__ n_sel_ptr,
                    const double *__restrict__ tau0_in, GroupedRefs R, int n_slices,
                    int32_t *__restrict__ part_idx, double *__restrict__ part_d2,
                    int32_t *__restrict__ idx0, double *__restrict__ dist) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *q_s = reinterpret_cast<double *>(smem_raw);
  double *yq = q_s + KX_DC;
  double *r_s = yq + KX_DC;
  double *top_d = r_s + KX_THREADS * KX_RS;
  double *buf_d = top_d + k;
  int32_t *top_i = reinterpret_cast<int32_t *>(buf_d + KX_THREADS);
  int32_t *buf_i = top_i + k;
  int32_t *row_s = buf_i + KX_THREADS;
  int32_t *sel = row_s + KX_THREADS;
  __shared__ int s_bufn, s_topn;
  __shared__ double s_tau;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t item = blockIdx.x;
  if (item >= (int64_t)*n_sel_ptr) return;
  const int64_t q = qsel[item];
  const double tau0 = tau0_in[item];
  if (tid < KX_DC) {
    q_s[tid] = tid < d ? qry[q * d + tid] : 0.0;
    yq[tid] = tid < R.ld ? (double)__half2float(__ushort_as_half(R.yq16[q * R.ld + tid])) : 0.0;
  }
  if (tid == 0) {
    s_bufn = 0;
    s_topn = 0;
    s_tau = tau0;
  }
  __syncthreads();
  // groups that may hold a row within sqrt(tau0)
  const double reach = (*R.scale) * sqrt(tau0) * (1.0 + 1e-9) + R.q_err[q] + (*R.e_max) + 1e-6;
  for (int g = tid; g < R.G; g += KX_THREADS) {
    if (R.goff[g + 1] <= R.goff[g]) {
      sel[g] = 0;
      continue;
    }
    const float *c = R.cent + (size_t)g * R.ld;
    double s2 = 0.0;
    for (int j = 0; j < d; ++j) {  // columns >= d of the operand rows are not coordinates
      const double t = yq[j] - (double)c[j];
      s2 = fma(t, t, s2);
    }
    const double lb = sqrt(s2) * (1.0 - 1e-6) - (double)R.grad[g];
    sel[g] = !(lb > reach);  // also keeps everything when tau0 is +inf / NaN
  }
  __syncthreads();
  // every slice of this row enumerates the selected tiles in the same (group) order
  int n = 0;
  for (int g = 0; g < R.G; ++g) {
    if (!sel[g]) continue;
    const int t0 = R.goff[g] / KX_THREADS, nt = (R.goff[g + 1] - R.goff[g]) / KX_THREADS;
    for (int t = 0; t < nt; ++t, ++n) {
      if (n % n_slices != (int)blockIdx.y) continue;
      __syncthreads();  // the previous tile's r_s / row_s / buffer are no longer in use
      row_s[tid] = R.rperm[(int64_t)(t0 + t) * KX_THREADS + tid];
      __syncthreads();
      // the rows of a tile lie all over the reference set: asynchronous copies put every row in flight at
      // once (one memory round trip per tile; a load-then-store loop paid one per row, ~22 us per tile)
      for (int rl = warp; rl < KX_THREADS; rl += KX_THREADS / 32) {
        const int r = row_s[rl];
        if (r >= 0)
          for (int j = lane; j < d; j += 32) kx_cp_async8(r_s + rl * KX_RS + j, ref + (int64_t)r * d + j);
      }
      asm volatile("cp.async.wait_all;" ::: "memory");
      __syncthreads();
      const int row = row_s[tid];
      if (row >= 0) {
        const double *rr = r_s + tid * KX_RS;
        double acc = 0.0;
        for (int j = 0; j < d; ++j) {
          const double tt = __dsub_rn(q_s[j], rr[j]);
          acc = __dadd_rn(acc, __dmul_rn(tt, tt));
        }
        if (acc <= s_tau) {
          const int pos = atomicAdd(&s_bufn, 1);
          buf_d[pos] = acc;
          buf_i[pos] = row;
        }
      }
      __syncthreads();
      if (warp == 0 && s_bufn > 0) {
        const int tn = topk_merge(top_d, top_i, s_topn, k, buf_d, buf_i, s_bufn, lane);
        if (lane == 0) {
          s_topn = tn;
          s_bufn = 0;
          s_tau = tn == k ? top_d[k - 1] : tau0;
        }
      }
    }
  }
  __syncthreads();
  const int tn = s_topn;
  if (part_idx != nullptr) {
    const int64_t o = (item * n_slices + blockIdx.y) * k;
    for (int t = tid; t < k; t += KX_THREADS) {
      part_idx[o + t] = t < tn ? top_i[t] : 0x7fffffff;
      part_d2[o + t] = t < tn ? top_d[t] : CUDART_INF;
    }
    return;
  }
  for (int t = tid; t < k; t += KX_THREADS) {
    idx0[q * k + t] = t < tn ? top_i[t] : -1;
    dist[q * k + t] = t < tn ? sqrt(top_d[t]) : CUDART_NAN;
  }
}

// merge the per-slice partial results of one fallback row: sort n_slices * k (d2, idx)
// pairs ascending (ascending-comparator bitonic network, virtual +inf padding) and keep k
__global__ void __launch_bounds__(256)
k_merge_parts(const int32_t *__restrict__ qsel, int k, int n_slices,
              const int32_t *__restrict__ part_idx, const double *__restrict__ part_d2,
              int32_t *__restrict__ idx0, double *__restrict__ dist) {
  extern __shared__ __align__(16) unsigned char ms_raw[];
  const int m = n_slices * k;
  unsigned long long *key = reinterpret_cast<unsigned long long *>(ms_raw);
  int32_t *val = reinterpret_cast<int32_t *>(key + m);
  const int64_t item = blockIdx.x;
  const int tid = threadIdx.x, nthr = blockDim.x;
  for (int t = tid; t < m; t += nthr) {
    key[t] = (unsigned long long)__double_as_longlong(part_d2[item * m + t]);  // d2 >= 0
    val[t] = part_idx[item * m + t];
  }
  __syncthreads();
  int P = 1;
  while (P < m) P <<= 1;
  const int half = P >> 1;
  for (int kk = 2; kk <= P; kk <<= 1) {
    for (int jj = kk >> 1, first = 1; jj > 0; jj >>= 1, first = 0) {
      for (int t = tid; t < half; t += nthr) {
        int lo, hi;
        if (first) {
          const int hk = kk >> 1, blk = t / hk, off = t - blk * hk;
          lo = blk * kk + off;
          hi = blk * kk + kk - 1 - off;
        } else {
          const int blk = t / jj, off = t - blk * jj;
          lo = blk * 2 * jj + off;
          hi = lo + jj;
        }
        if (hi < m) {
          unsigned long long ka = key[lo], kb = key[hi];
          int32_t ia = val[lo], ib = val[hi];
          if (ka > kb || (ka == kb && ia > ib)) {
            key[lo] = kb; key[hi] = ka;
            val[lo] = ib; val[hi] = ia;
          }
        }
      }
      __syncthreads();
    }
  }
  const int64_t q = qsel[item];
  for (int t = tid; t < k; t += nthr) {
    idx0[q * k + t] = val[t];
    dist[q * k + t] = sqrt(__longlong_as_double((long long)key[t]));
  }
}

template <int QB>
int launch_qb(b200nn_ctx *ctx, const double *d_ref, int64_t nr, const double *d_qry, int d, int k,
              const int32_t *d_qsel, const int32_t *d_nsel, int64_t n_items, int32_t *d_idx0,
              double *d_dist, int n_slices = 1, int32_t *part_idx = nullptr,
              double *part_d2 = nullptr, const double *d_tau0 = nullptr) {
  size_t smem = kx_smem_bytes(QB, k);
  if (smem > 227 * 1024) {
    b200nn_set_error("k = %d is too large for the exact kNN kernel (shared memory %zu B)", k, smem);
    return B200NN_ERR_INVALID;
  }
  CU_TRY(cudaFuncSetAttribute(k_knn_exact<QB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)smem));
  int64_t blocks = ceil_div64(n_items, QB);
  if (blocks <= 0) return B200NN_OK;
  const int64_t slice_len = n_slices > 1 ? ceil_div64(ceil_div64(nr, n_slices), KX_THREADS) * KX_THREADS : nr;
  dim3 grid((unsigned)blocks, (unsigned)n_slices);
  k_knn_exact<QB><<<grid, KX_THREADS, smem, ctx->stream>>>(d_ref, nr, d_qry, d, k, d_qsel, d_nsel,
                                                          n_items, d_idx0, d_dist, slice_len,
                                                          n_slices, part_idx, part_d2, d_tau0);
  KERNEL_CHECK(ctx);
  return B200NN_OK;
}

}  // namespace

int knn_exact_launch(b200nn_ctx *ctx, const double *d_ref, int64_t nr, const double *d_qry,
                     int64_t nq, int d, int k, const int32_t *d_qsel, int64_t n_sel,
                     int32_t *d_idx0, double *d_dist, const double *d_tau0) {
  // d_tau0 (optional, with d_qsel): per listed row an upper bound of its k-th squared distance
  // in the reference arithmetic; rows beyond it are never buffered.
  // d_qsel != nullptr: scan only the n_sel query rows listed there; the int32 just before
  // the list (d_qsel[-1]) holds the same count on the device.
  const int32_t *d_nsel = d_qsel ? d_qsel - 1 : nullptr;
  int64_t n_items = d_qsel ? n_sel : nq;
  if (n_items <= 0) return B200NN_OK;
  if (d_qsel == nullptr) {
    if (kx_smem_bytes(16, k) <= 110 * 1024)
      return launch_qb<16>(ctx, d_ref, nr, d_qry, d, k, nullptr, nullptr, n_items, d_idx0, d_dist);
    if (kx_smem_bytes(4, k) <= 227 * 1024)
      return launch_qb<4>(ctx, d_ref, nr, d_qry, d, k, nullptr, nullptr, n_items, d_idx0, d_dist);
    return launch_qb<1>(ctx, d_ref, nr, d_qry, d, k, nullptr, nullptr, n_items, d_idx0, d_dist);
  }
  // A handful of rows against the whole reference set: split the scan over slices of the
  // reference set so that the grid fills the machine, then merge the partial top-k lists.
  const int64_t blocks4 = ceil_div64(n_items, 4);
  int64_t want = ceil_div64((int64_t)ctx->sm_count * 16, blocks4);
  int64_t cap = 8192 / k;
  int64_t by_rows = ceil_div64(nr, 4 * KX_THREADS);
  int n_slices = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(want, cap), std::min<int64_t>(by_rows, 512)));
  if (n_slices <= 1 || kx_smem_bytes(4, k) > 227 * 1024) {
    if (kx_smem_bytes(4, k) <= 227 * 1024)
      return launch_qb<4>(ctx, d_ref, nr, d_qry, d, k, d_qsel, d_nsel, n_items, d_idx0, d_dist, 1, nullptr,
                          nullptr, d_tau0);
    return launch_qb<1>(ctx, d_ref, nr, d_qry, d, k, d_qsel, d_nsel, n_items, d_idx0, d_dist, 1, nullptr,
                        nullptr, d_tau0);
  }
  int32_t *part_idx;
  double *part_d2;
  const size_t cnt = (size_t)n_items * n_slices * k;
  RC_TRY(ws_get(ctx, WS_STAGE_B, cnt * 2, &part_d2));  // doubles, then the int32 block
  part_idx = reinterpret_cast<int32_t *>(part_d2 + cnt);
  RC_TRY((launch_qb<4>(ctx, d_ref, nr, d_qry, d, k, d_qsel, d_nsel, n_items, d_idx0, d_dist,
                       n_slices, part_idx, part_d2, d_tau0)));
  const size_t msmem = (size_t)n_slices * k * 12 + 16;
  CU_TRY(cudaFuncSetAttribute(k_merge_parts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem));
  k_merge_parts<<<(unsigned)n_items, 256, msmem, ctx->stream>>>(d_qsel, k, n_slices, part_idx, part_d2,
                                                               d_idx0, d_dist);
  KERNEL_CHECK(ctx);
  return B200NN_OK;
}

int knn_exact_grouped_launch(b200nn_ctx *ctx, const double *d_ref, const double *d_qry, int d, int k,
                             const int32_t *d_qsel, int64_t n_sel, const double *d_tau0,
                             const GroupedRefs &R, int32_t *d_idx0, double *d_dist) {
  if (n_sel <= 0) return B200NN_OK;
  if (d > KX_DC || R.ld > KX_DC || R.G > KG_MAXG || kg_smem_bytes(k) > 227 * 1024) {
    b200nn_set_error("grouped exact scan: unsupported shape (d = %d, k = %d, groups = %d)", d, k, R.G);
    return B200NN_ERR_INVALID;
  }
  // few rows: split every row's tile sequence over slices so that the grid fills the machine
  const int64_t want = ceil_div64((int64_t)ctx->sm_count * 4, n_sel);
  const int n_slices = (int)std::max<int64_t>(1, std::min<int64_t>(want, std::min<int64_t>(8192 / k, 64)));
  const size_t smem = kg_smem_bytes(k);
  CU_TRY(cudaFuncSetAttribute(k_knn_exact_grouped, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int32_t *part_idx = nullptr;
  double *part_d2 = nullptr;
  if (n_slices > 1) {
    const size_t cnt = (size_t)n_sel * n_slices * k;
    RC_TRY(ws_get(ctx, WS_STAGE_B, cnt * 2, &part_d2));  // doubles, then the int32 block
    part_idx = reinterpret_cast<int32_t *>(part_d2 + cnt);
  }
  dim3 grid((unsigned)n_sel, (unsigned)n_slices);
  k_knn_exact_grouped<<<grid, KX_THREADS, smem, ctx->stream>>>(d_ref, d_qry, d, k, d_qsel, d_qsel - 1, d_tau0, R,
                                                              n_slices, part_idx, part_d2, d_idx0, d_dist);
  KERNEL_CHECK(ctx);
  if (n_slices > 1) {
    const size_t msmem = (size_t)n_slices * k * 12 + 16;
    CU_TRY(cudaFuncSetAttribute(k_merge_parts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem));
    k_merge_parts<<<(unsigned)n_sel, 256, msmem, ctx->stream>>>(d_qsel, k, n_slices, part_idx, part_d2, d_idx0,
                                                               d_dist);
    KERNEL_CHECK(ctx);
  }
  return B200NN_OK;
}
