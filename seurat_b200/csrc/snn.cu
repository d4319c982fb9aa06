// snn.cu -- nn Graph (K3) and shared-nearest-neighbour graph (K4).
//
// Replaces, bit for bit:
//   * the nn Graph of /root/reference/R/clustering.R:664-670
//       sparseMatrix(i = row, j = nn.ranked, x = 1, dims = c(n, n)), duplicates summed
//   * ComputeSNN, /root/reference/src/snn.cpp:14-38
//       A[i, nn(i,j)] += 1;  S = A * A^T;  v <- v / (k + (k - v));
//       v < prune -> 0 (strict);  exact zeros dropped;  column-compressed, sorted.
//
// K3 (transpose): histogram of neighbour ids -> scan -> scatter of row ids ->
//     per-column sort.  The sorted reverse lists R(c) = {i : c in N(i)} ARE
//     the nn Graph in CSC form.
// K4 (SNN rows): S[i, j] = sum over the k listed neighbours c of row i of the
//     multiplicity of j in R(c).  A row-group of threads walks the k reverse
//     lists (coalesced reads of sorted runs), counts every candidate j in a
//     shared-memory hash table (packed key|count words), keeps the entries
//     whose count passes the prune cut-off (a 2k-entry lookup table built from
//     the exact double expression), sorts the survivors by column and emits the
//     row.  The n x n pre-prune product is never materialised.  Because S is
//     symmetric the row-compressed arrays equal the column-compressed ones.
//
// Rows are binned by expanded work E_i = sum_c |R(c)| into classes that differ
// only in table size / threads per row; the "huge" class hashes in HBM.
#include <limits.h>
#include <stdlib.h>

#include <algorithm>

#include "internal.cuh"

namespace {

// ---------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ int warp_sum_i32(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_max_i32(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ long long warp_sum_i64(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Ascending sort of a[0..m) by `nthr` cooperating threads (thread id g).
// Bitonic network written with ascending comparators only (first step of each
// stage mirrors the partner), so that the virtual +inf padding up to the next
// power of two never has to move: comparators touching index >= m are skipped.
template <typename T, typename SyncFn>
__device__ __forceinline__ void bitonic_sort_asc(T *a, int m, int g, int nthr, SyncFn sync) {
  if (m <= 1) {
    return;
  }
  int P = 1;
  while (P < m) P <<= 1;
  const int half = P >> 1;
  for (int kk = 2; kk <= P; kk <<= 1) {
    // mirrored step
    {
      const int hk = kk >> 1;
      for (int t = g; t < half; t += nthr) {
        int blk = t / hk, off = t - blk * hk;
        int lo = blk * kk + off, hi = blk * kk + kk - 1 - off;
        if (hi < m) {
          T x = a[lo], y = a[hi];
          if (x > y) {
            a[lo] = y;
            a[hi] = x;
          }
        }
      }
      sync();
    }
    for (int j = kk >> 2; j > 0; j >>= 1) {
      for (int t = g; t < half; t += nthr) {
        int blk = t / j, off = t - blk * j;
        int lo = blk * 2 * j + off, hi = lo + j;
        if (hi < m) {
          T x = a[lo], y = a[hi];
          if (x > y) {
            a[lo] = y;
            a[hi] = x;
          }
        }
      }
      sync();
    }
  }
}

// ---------------------------------------------------------------------------
// K3: transpose
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_hist(const int32_t *__restrict__ idx0, int64_t total, int k, int64_t n, int32_t *__restrict__ indeg,
       DevInfo *__restrict__ info) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  int32_t c = idx0[t];
  if (c < 0 || (int64_t)c >= n)
    info->err_index_range = 1;
  else
    atomicAdd(&indeg[c], 1);
  // a row that lists the same neighbour twice (A_ic = 2): the SNN kernels must know before they start,
  // and every row counts (row j's duplicate changes S_ij of every row i)
  const int p = (int)(t % k);
  const int32_t *row = idx0 + (t - p);
  bool dup = false;
  for (int q = 0; q < p; ++q) dup |= row[q] == c;
  if (dup) info->has_dups = 1;
}

// longest reverse list (statistics; the work classes use the per-row sums)
__global__ void __launch_bounds__(256)
k_max_indeg(const int32_t *__restrict__ indeg, int64_t n, DevInfo *__restrict__ info) {
  int m = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = max(m, indeg[i]);
  m = warp_max_i32(m);
  if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(&info->max_indegree, m);
}

__global__ void __launch_bounds__(256)
k_scatter(const int32_t *__restrict__ idx0, int64_t total, int k, int64_t n,
          const int64_t *__restrict__ colptr, int32_t *__restrict__ cursor,
          int32_t *__restrict__ rev) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  int32_t c = idx0[t];
  if (c < 0 || (int64_t)c >= n) return;
  int pos = atomicAdd(&cursor[c], 1);
  rev[colptr[c] + pos] = (int32_t)(t / k);
}

constexpr int SORT_WARPS = 8;
constexpr int SORT_WARP_CAP = 1024;  // elements a warp sorts in shared memory

// one warp per column: sorted copy of the reverse list (the nn Graph's column), distinct rows counted.
// Out of place: the SNN kernels read the unsorted lists concurrently (they do not need an order).
template <typename T>  // int32_t: reverse lists (uniq counted); uint32_t: packed SNN words (uniq == nullptr)
__global__ void __launch_bounds__(SORT_WARPS * 32)
k_sort_columns(const T *__restrict__ rev, T *__restrict__ out, const int64_t *__restrict__ colptr,
               int64_t n, int32_t *__restrict__ uniq, int32_t *__restrict__ long_list,
               int32_t *__restrict__ n_long) {
  __shared__ T sbuf[SORT_WARPS][SORT_WARP_CAP];
  constexpr T TMAX = (T)(~(T)0) > 0 ? (T)(~(T)0) : (T)INT_MAX;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t c = (int64_t)blockIdx.x * SORT_WARPS + w;
  if (c >= n) return;
  const int64_t beg = colptr[c];
  const int64_t len64 = colptr[c + 1] - beg;
  const int len = (int)len64;
  if (len <= 1) {
    if (lane == 0) {
      if (uniq) uniq[c] = len;
      if (len == 1) out[beg] = rev[beg];
    }
    return;
  }
  int ndup = 0;
  if (len <= 32) {
    T v = lane < len ? rev[beg + lane] : TMAX;
#pragma unroll
    for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
      for (int j = kk >> 1; j > 0; j >>= 1) {
        T o = __shfl_xor_sync(0xffffffffu, v, j);
        bool up = (lane & kk) == 0, lower = (lane & j) == 0;
        v = (lower == up) ? min(v, o) : max(v, o);
      }
    }
    T prev = __shfl_up_sync(0xffffffffu, v, 1);
    if (lane < len) out[beg + lane] = v;
    ndup = (lane > 0 && lane < len && v == prev) ? 1 : 0;
  } else if (len <= SORT_WARP_CAP) {
    T *a = sbuf[w];
    for (int t = lane; t < len; t += 32) a[t] = rev[beg + t];
    __syncwarp();
    bitonic_sort_asc(a, len, lane, 32, [] { __syncwarp(); });
    for (int t = lane; t < len; t += 32) {
      T v = a[t];
      out[beg + t] = v;
      if (t > 0 && a[t - 1] == v) ndup++;
    }
  } else {
    if (lane == 0) {
      int pos = atomicAdd(n_long, 1);
      long_list[pos] = (int32_t)c;
    }
    return;  // uniq[c] is written by k_sort_long_columns
  }
  ndup = warp_sum_i32(ndup);
  if (lane == 0 && uniq) uniq[c] = len - ndup;
}

constexpr int LONG_THREADS = 1024;
constexpr int LONG_SMEM_CAP = 32768;  // elements (128 KB)

// one CTA per long column (hubs)
template <typename T>
__global__ void __launch_bounds__(LONG_THREADS)
k_sort_long_columns(const T *__restrict__ rev, T *__restrict__ out, const int64_t *__restrict__ colptr,
                    int32_t *__restrict__ uniq, const int32_t *__restrict__ long_list,
                    const int32_t *__restrict__ n_long_p) {
  extern __shared__ int32_t lbuf_raw[];
  T *lbuf = reinterpret_cast<T *>(lbuf_raw);
  __shared__ int red[32];
  const int n_long = *n_long_p;
  for (int li = blockIdx.x; li < n_long; li += gridDim.x) {
    const int64_t c = long_list[li];
    const int64_t beg = colptr[c];
    const int len = (int)(colptr[c + 1] - beg);
    T *a;
    if (len <= LONG_SMEM_CAP) {
      a = lbuf;
    } else {
      a = out + beg;  // sorted in HBM (block-scope visibility via __syncthreads)
    }
    for (int t = threadIdx.x; t < len; t += LONG_THREADS) a[t] = rev[beg + t];
    __syncthreads();
    bitonic_sort_asc(a, len, (int)threadIdx.x, LONG_THREADS, [] { __syncthreads(); });
    int ndup = 0;
    for (int t = threadIdx.x; t < len; t += LONG_THREADS) {
      T v = a[t];
      if (t > 0 && a[t - 1] == v) ndup++;
    }
    __syncthreads();
    if (len <= LONG_SMEM_CAP)
      for (int t = threadIdx.x; t < len; t += LONG_THREADS) out[beg + t] = a[t];
    ndup = warp_sum_i32(ndup);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ndup;
    __syncthreads();
    if (threadIdx.x < 32) {
      int v = warp_sum_i32(red[threadIdx.x]);
      if (threadIdx.x == 0 && uniq) uniq[c] = len - v;
    }
    __syncthreads();
  }
}

// nn Graph emission: column c gets its distinct rows and their multiplicities
__global__ void __launch_bounds__(256)
k_emit_nn_graph(const int32_t *__restrict__ rev, const int64_t *__restrict__ colptr,
                const int32_t *__restrict__ uniq, const int64_t *__restrict__ nn_p, int64_t n,
                int32_t *__restrict__ nn_i, double *__restrict__ nn_x) {
  const int lane = threadIdx.x & 31;
  const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= n) return;
  const int64_t beg = colptr[c];
  const int len = (int)(colptr[c + 1] - beg);
  const int64_t ob = nn_p[c];
  if (uniq[c] == len) {
    for (int t = lane; t < len; t += 32) {
      nn_i[ob + t] = rev[beg + t];
      nn_x[ob + t] = 1.0;
    }
  } else if (lane == 0) {  // duplicated neighbours: rare, sequential
    int64_t o = ob - 1;
    int32_t prev = -1;
    for (int t = 0; t < len; ++t) {
      int32_t v = rev[beg + t];
      if (t == 0 || v != prev) {
        ++o;
        nn_i[o] = v;
        nn_x[o] = 1.0;
      } else {
        nn_x[o] += 1.0;
      }
      prev = v;
    }
  }
}

// ---------------------------------------------------------------------------
// K4: SNN rows
// ---------------------------------------------------------------------------
struct SnnParams {
  const int32_t *idx0;     // [n x k] 0-based
  const int64_t *colptr;   // [n + 1]
  const int32_t *rev;      // reverse lists
  const uint8_t *keep;     // keep[count]
  const int32_t *row_list; // rows of this class (local row numbers)
  int32_t n_rows_class;
  int32_t k;
  int32_t cbits;           // low bits of a word hold the count
  int64_t row_begin;       // global row of local row 0
  const int64_t *tmp_off;  // [rows + 1]
  void *tmp;               // packed survivors
  int32_t *row_nnz;        // [rows]
  void *huge_tables;       // HBM hash tables (huge class)
  int32_t huge_log_t;
  int32_t sym;             // 1: only candidates j > i are counted (upper triangle; S is symmetric and
                           //    the lower triangle + diagonal are mirrored afterwards)
  int32_t sorted;          // 1: `rev` holds the sorted lists (nn Graph columns)
};

template <typename Word>
struct WordTraits;
template <>
struct WordTraits<uint32_t> {
  static constexpr uint32_t EMPTY = 0xffffffffu;
  typedef unsigned int atomic_t;
};
template <>
struct WordTraits<uint64_t> {
  static constexpr uint64_t EMPTY = 0xffffffffffffffffull;
  typedef unsigned long long atomic_t;
};

template <typename Word>
__device__ __forceinline__ void hash_insert(Word *tab, uint32_t mask, int log_t, int cbits,
                                            uint32_t j) {
  typedef typename WordTraits<Word>::atomic_t A;
  const Word EMPTY = WordTraits<Word>::EMPTY;
  uint32_t h = (j * 2654435761u) >> (32 - log_t);
  const Word fresh = ((Word)j << cbits) | (Word)1;
  volatile Word *vt = tab;
  while (true) {
    Word cur = vt[h];
    if (cur == EMPTY) {
      Word old = (Word)atomicCAS(reinterpret_cast<A *>(tab + h), (A)EMPTY, (A)fresh);
      if (old == EMPTY) return;
      cur = old;
    }
    if ((uint32_t)(cur >> cbits) == j) {
      atomicAdd(reinterpret_cast<A *>(tab + h), (A)1);
      return;
    }
    h = (h + 1) & mask;
  }
}

// Atomic-free insert for a table owned by ONE warp.  Requires that the (up to 32) keys
// inserted by one call are pairwise distinct -- true for the entries of one reverse list
// unless the input lists a neighbour twice (the caller then uses the atomic variant).
// An empty slot is claimed with a plain store and a read-back (one of the racing lanes
// wins, the others probe on); an existing key is bumped with a plain read-modify-write,
// which no other lane can touch.  Shared-memory atomics (ATOMS.CAS) were the bottleneck
// of this stage: ~9 SM-cycles per insert against ~1 for plain LDS/STS traffic.
template <typename Word>
__device__ __forceinline__ void hash_insert_warp(Word *tab, uint32_t mask, int log_t, int cbits,
                                                 uint32_t j, bool active) {
  const Word EMPTY = WordTraits<Word>::EMPTY;
  uint32_t h = (j * 2654435761u) >> (32 - log_t);
  const Word fresh = ((Word)j << cbits) | (Word)1;
  volatile Word *vt = tab;
  bool todo = active;
  do {
    // phase 1: every lane walks its own probe sequence with plain loads (cheap, divergent)
    Word cur = EMPTY;
    if (todo) {
      while (true) {
        cur = vt[h];
        if (cur == EMPTY || (uint32_t)(cur >> cbits) == j) break;
        h = (h + 1) & mask;
      }
    }
    // phase 2 (warp-uniform): bump an existing key, or claim the empty slot
    const bool claim = todo && cur == EMPTY;
    if (todo && !claim) vt[h] = cur + 1;
    if (claim) vt[h] = fresh;
    __syncwarp();
    // a lane whose claim was overwritten by another lane's (different) key probes on
    todo = claim && vt[h] != fresh;
    __syncwarp();
  } while (__any_sync(0xffffffffu, todo));
}

// G threads cooperate on one row.  LOG_T > 0: table of 2^LOG_T words in shared
// memory (+ a survivor buffer of half that size when G > 32);  LOG_T == 0: table
// in HBM, survivors written straight to their slot in `tmp` and sorted there.
template <typename Word, int LOG_T, int G, int RPB>
__global__ void __launch_bounds__(G *RPB) k_snn_rows(SnnParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const Word EMPTY = WordTraits<Word>::EMPTY;
  constexpr bool HUGE = (LOG_T == 0);
  constexpr bool INPLACE = (G == 32) && !HUGE;
  const int grp = threadIdx.x / G;
  const int g = threadIdx.x - grp * G;
  const int lane = threadIdx.x & 31;
  const int log_t = HUGE ? P.huge_log_t : LOG_T;
  const uint32_t T = 1u << log_t;
  const uint32_t mask = T - 1;

  Word *tab;
  Word *surv = nullptr;
  if (HUGE) {
    tab = reinterpret_cast<Word *>(P.huge_tables) + (size_t)blockIdx.x * T;
  } else {
    constexpr size_t per_group = ((size_t)1 << LOG_T) + (INPLACE ? 0 : ((size_t)1 << LOG_T) / 2);
    tab = reinterpret_cast<Word *>(smem_raw) + (size_t)grp * per_group;
    if (!INPLACE) surv = tab + T;
  }
  __shared__ int s_count[RPB];

  auto gsync = [&]() {
    if (G == 32) {
      __syncwarp();
    } else if (RPB == 1) {
      __syncthreads();
    } else {
      asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(G) : "memory");
    }
  };

  const int n_groups = gridDim.x * RPB;
  for (int r = blockIdx.x * RPB + grp; r < P.n_rows_class; r += n_groups) {
    const int row_l = P.row_list[r];
    const int64_t row = P.row_begin + row_l;
    // 1. clear
    for (uint32_t s = g; s < T; s += G) tab[s] = EMPTY;
    if (g == 0) s_count[grp] = 0;
    gsync();
    // 2. count every candidate j of the k reverse lists
    const int32_t *nbr = P.idx0 + row * P.k;
    for (int p = 0; p < P.k; ++p) {
      const int32_t c = nbr[p];
      const int64_t beg = P.colptr[c];
      const int len = (int)(P.colptr[c + 1] - beg);
      for (int t = g; t < len; t += G) {
        const int32_t j = P.rev[beg + t];
        if (!P.sym || (int64_t)j > row) hash_insert<Word>(tab, mask, log_t, P.cbits, (uint32_t)j);
      }
    }
    gsync();
    // 3. keep the entries that pass the prune cut-off
    const Word cmask = (((Word)1) << P.cbits) - 1;
    int m = 0;
    Word *out;
    if (INPLACE) {
      out = tab;
      for (uint32_t base = 0; base < T; base += 32) {
        Word wv = tab[base + lane];
        bool kp = (wv != EMPTY) && P.keep[(uint32_t)(wv & cmask)];
        unsigned bal = __ballot_sync(0xffffffffu, kp);
        __syncwarp();  // every lane has read its slot of this round
        if (kp) tab[m + __popc(bal & ((1u << lane) - 1))] = wv;
        m += __popc(bal);
      }
      __syncwarp();
    } else {
      out = HUGE ? reinterpret_cast<Word *>(P.tmp) + P.tmp_off[row_l] : surv;
      for (uint32_t base = 0; base < T; base += G) {
        uint32_t s = base + g;
        Word wv = s < T ? tab[s] : EMPTY;
        bool kp = (wv != EMPTY) && P.keep[(uint32_t)(wv & cmask)];
        unsigned bal = __ballot_sync(0xffffffffu, kp);
        int wbase = 0;
        if (lane == 0 && bal) wbase = atomicAdd(&s_count[grp], __popc(bal));
        wbase = __shfl_sync(0xffffffffu, wbase, 0);
        if (kp) out[wbase + __popc(bal & ((1u << lane) - 1))] = wv;
      }
      gsync();
      m = s_count[grp];
    }
    // 4. ascending column order (keys are distinct, so word order == key order)
    bitonic_sort_asc(out, m, g, G, gsync);
    // 5. emit
    if (!HUGE) {
      Word *dst = reinterpret_cast<Word *>(P.tmp) + P.tmp_off[row_l];
      for (int t = g; t < m; t += G) dst[t] = out[t];
    }
    if (g == 0) P.row_nnz[row_l] = m;
    gsync();
  }
}

// Warp-per-row variant for the common classes (table <= 8192 words).  Compared with the
// cooperative kernel above it needs no barriers, fetches the k list descriptors with one
// load per lane, and keeps four list loads in flight per lane before inserting.
template <typename Word, int LOG_T, int WARPS, bool ATOMIC>
__global__ void __launch_bounds__(WARPS * 32) k_snn_rows_warp(SnnParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const Word EMPTY = WordTraits<Word>::EMPTY;
  constexpr uint32_t T = 1u << LOG_T;
  constexpr uint32_t mask = T - 1;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Word *tab = reinterpret_cast<Word *>(smem_raw) + (size_t)w * T;
  const Word cmask = (((Word)1) << P.cbits) - 1;
  const int n_warps = gridDim.x * WARPS;
  for (int r = blockIdx.x * WARPS + w; r < P.n_rows_class; r += n_warps) {
    const int row_l = P.row_list[r];
    const int64_t row = P.row_begin + row_l;
    // 1. clear (16-byte stores)
    {
      constexpr int VEC = 16 / sizeof(Word);
      uint4 e4;
      e4.x = e4.y = e4.z = e4.w = 0xffffffffu;
      uint4 *t4 = reinterpret_cast<uint4 *>(tab);
      for (uint32_t s = lane; s < T / VEC; s += 32) t4[s] = e4;
    }
    __syncwarp();
    // 2. count every candidate of the k reverse lists
    const int32_t *nbr = P.idx0 + row * P.k;
    for (int p0 = 0; p0 < P.k; p0 += 32) {
      int64_t beg = 0;
      int len = 0;
      if (p0 + lane < P.k) {
        const int32_t c = nbr[p0 + lane];
        beg = P.colptr[c];
        len = (int)(P.colptr[c + 1] - beg);
      }
      const int np = (P.k - p0) < 32 ? (P.k - p0) : 32;
      for (int p = 0; p < np; ++p) {
        const int64_t b = __shfl_sync(0xffffffffu, beg, p);
        const int l = __shfl_sync(0xffffffffu, len, p);
        const int32_t *src = P.rev + b;
        for (int t = lane; ATOMIC && t < l; t += 128) {
          // up to four independent loads in flight, then the (serialising) inserts
          const int lowkey = P.sym ? (int)row : -1;
          int32_t v0 = src[t];
          int32_t v1 = (t + 32 < l) ? src[t + 32] : -1;
          int32_t v2 = (t + 64 < l) ? src[t + 64] : -1;
          int32_t v3 = (t + 96 < l) ? src[t + 96] : -1;
          if (v0 <= lowkey) v0 = -1;
          if (v1 <= lowkey) v1 = -1;
          if (v2 <= lowkey) v2 = -1;
          if (v3 <= lowkey) v3 = -1;
          if (ATOMIC) {
            if (v0 >= 0) hash_insert<Word>(tab, mask, LOG_T, P.cbits, (uint32_t)v0);
            if (v1 >= 0) hash_insert<Word>(tab, mask, LOG_T, P.cbits, (uint32_t)v1);
            if (v2 >= 0) hash_insert<Word>(tab, mask, LOG_T, P.cbits, (uint32_t)v2);
            if (v3 >= 0) hash_insert<Word>(tab, mask, LOG_T, P.cbits, (uint32_t)v3);
          }
        }
        if (!ATOMIC) {
          // warp-synchronous variant: the loop bound must be warp-uniform
          for (int t0 = 0; t0 < l; t0 += 128) {
            const int t = t0 + lane;
            const int lowkey = P.sym ? (int)row : -1;
            int32_t v0 = (t < l) ? src[t] : -1;
            int32_t v1 = (t + 32 < l) ? src[t + 32] : -1;
            int32_t v2 = (t + 64 < l) ? src[t + 64] : -1;
            int32_t v3 = (t + 96 < l) ? src[t + 96] : -1;
            if (v0 <= lowkey) v0 = -1;
            if (v1 <= lowkey) v1 = -1;
            if (v2 <= lowkey) v2 = -1;
            if (v3 <= lowkey) v3 = -1;
            hash_insert_warp<Word>(tab, mask, LOG_T, P.cbits, (uint32_t)v0, v0 >= 0);
            if (t0 + 32 < l) hash_insert_warp<Word>(tab, mask, LOG_T, P.cbits, (uint32_t)v1, v1 >= 0);
            if (t0 + 64 < l) hash_insert_warp<Word>(tab, mask, LOG_T, P.cbits, (uint32_t)v2, v2 >= 0);
            if (t0 + 96 < l) hash_insert_warp<Word>(tab, mask, LOG_T, P.cbits, (uint32_t)v3, v3 >= 0);
          }
        }
      }
    }
    __syncwarp();
    // 3. keep the entries that pass the prune cut-off (in-place compaction)
    int m = 0;
    for (uint32_t base = 0; base < T; base += 32) {
      Word wv = tab[base + lane];
      bool kp = (wv != EMPTY) && P.keep[(uint32_t)(wv & cmask)];
      unsigned bal = __ballot_sync(0xffffffffu, kp);
      if (bal) {  // warp-uniform
        __syncwarp();
        if (kp) tab[m + __popc(bal & ((1u << lane) - 1))] = wv;
        m += __popc(bal);
      }
    }
    __syncwarp();
    // 4. ascending column order, 5. emit
    bitonic_sort_asc(tab, m, lane, 32, [] { __syncwarp(); });
    Word *dst = reinterpret_cast<Word *>(P.tmp) + P.tmp_off[row_l];
    for (int t = lane; t < m; t += 32) dst[t] = tab[t];
    if (lane == 0) P.row_nnz[row_l] = m;
    __syncwarp();
  }
}

template <typename Word, int LOG_T, int WARPS>
int launch_class_warp(b200nn_ctx *ctx, SnnParams P, int blocks_per_sm, bool atomic) {
  if (P.n_rows_class <= 0) return B200NN_OK;
  size_t smem = ((size_t)1 << LOG_T) * sizeof(Word) * WARPS;
  int64_t want = ceil_div64(P.n_rows_class, WARPS);
  int64_t cap = (int64_t)ctx->sm_count * blocks_per_sm;
  int blocks = (int)std::min<int64_t>(want, cap);
  if (atomic) {
    CU_TRY(cudaFuncSetAttribute(k_snn_rows_warp<Word, LOG_T, WARPS, true>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_snn_rows_warp<Word, LOG_T, WARPS, true><<<blocks, WARPS * 32, smem, ctx->stream>>>(P);
  } else {
    CU_TRY(cudaFuncSetAttribute(k_snn_rows_warp<Word, LOG_T, WARPS, false>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_snn_rows_warp<Word, LOG_T, WARPS, false><<<blocks, WARPS * 32, smem, ctx->stream>>>(P);
  }
  KERNEL_CHECK(ctx);
  return B200NN_OK;
}

// ---------------------------------------------------------------------------
// Filtered warp-per-row kernel (prune cut-off needs an overlap of s_min >= 2).
// Only ~2% of the ~1000 distinct candidates of a row survive the cut-off, so exact
// hashing of every candidate occurrence is wasted work.  Pass A counts occurrences in
// a direct-mapped array of 4-bit counters that stop at s_min (no keys, no probing;
// colliding keys only ever over-count).  Pass B walks the lists again and sends to the exact hash table
// only the occurrences whose counter reached s_min -- true survivors plus a few
// collision victims (~10% of the occurrences), compacted into dense batches first.
// The exact table is small; a row that would overflow it is handed to the big-table
// kernel instead.  Counts, cut-off and output are exactly those of the plain kernel.
// ---------------------------------------------------------------------------
// exact table of 2^tx_log words; ~78 % of it may fill before the row is handed on
constexpr int filt_slots(int tx_log) { return (1 << tx_log) * 25 / 32; }      // distinct keys allowed
constexpr int filt_slot_cap(int tx_log) { return (1 << tx_log) * 26 / 32; }
// 2^cw_log four-bit counters (the cut-off only asks "count >= s_min", s_min <= F_SMIN_MAX)
constexpr int F_SMIN_MAX = 12;
constexpr int filt_bytes(int cw_log, int tx_log) {
  return (1 << cw_log) / 2 + (1 << tx_log) * 4 + filt_slot_cap(tx_log) * 2 + 64 * 4;
}

template <int CW_LOG>
__device__ __forceinline__ uint32_t hash_cnt(uint32_t j) { return (j * 0x9E3779B1u) >> (32 - CW_LOG); }

// exact insert of one batch (<= 32 occurrences, repeats of a key allowed).  A key that is
// already in the table costs one probe and one shared-memory atomic add; only the first
// occurrence of a key claims a slot with a compare-and-swap (~40 per row).  The claimed slots
// are recorded so that the table is read out and cleaned without scanning it.
template <int F_TX_LOG>
__device__ __forceinline__ void filt_insert_batch(uint32_t *tab, uint16_t *slots, int &nslots, int cbits,
                                                  uint32_t j, bool active, int lane) {
  constexpr uint32_t mask = (1u << F_TX_LOG) - 1;
  constexpr int F_SLOT_CAP = filt_slot_cap(F_TX_LOG);
  uint32_t h = (j * 2654435761u) >> (32 - F_TX_LOG);
  const uint32_t fresh = (j << cbits) | 1u;
  volatile uint32_t *vt = tab;
  bool won = false, todo = active;
  while (todo) {
    uint32_t cur = vt[h];
    if (cur == 0xffffffffu) {
      cur = atomicCAS(&tab[h], 0xffffffffu, fresh);
      if (cur == 0xffffffffu) {
        won = true;
        break;
      }
    }
    if ((cur >> cbits) == j) {
      atomicAdd(&tab[h], 1u);
      todo = false;
    } else {
      h = (h + 1) & mask;
    }
  }
  __syncwarp();
  const unsigned wb = __ballot_sync(0xffffffffu, won);
  if (won) {
    const int pos = nslots + __popc(wb & ((1u << lane) - 1));
    if (pos < F_SLOT_CAP) slots[pos] = (uint16_t)h;
  }
  nslots += __popc(wb);
  __syncwarp();
}

// log2 of the byte counters per row, log2 of the exact table's words, warps (rows) per CTA
template <int CW_LOG, int F_TX_LOG, int F_WARPS>
__global__ void __launch_bounds__(F_WARPS * 32)
k_snn_rows_filter(SnnParams P, int s_min, int32_t *__restrict__ ovf_list, DevInfo *__restrict__ info) {
  constexpr int F_BYTES = filt_bytes(CW_LOG, F_TX_LOG);
  constexpr int F_SLOTS = filt_slots(F_TX_LOG), F_SLOT_CAP = filt_slot_cap(F_TX_LOG);
  constexpr int F_CW_LOG = CW_LOG;
  static_assert((1 << CW_LOG) / 2 >= F_SLOT_CAP * 4, "the counter area doubles as the survivor buffer");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char *base = smem_raw + (size_t)w * F_BYTES;
  uint32_t *cnt = reinterpret_cast<uint32_t *>(base);  // eight 4-bit counters per word
  uint32_t *tab = reinterpret_cast<uint32_t *>(base + (1 << F_CW_LOG) / 2);
  uint16_t *slots = reinterpret_cast<uint16_t *>(tab + (1 << F_TX_LOG));
  uint32_t *pend = reinterpret_cast<uint32_t *>(slots + F_SLOT_CAP);
  const uint32_t cmask = (1u << P.cbits) - 1;
  for (int s = lane; s < (1 << F_TX_LOG); s += 32) tab[s] = 0xffffffffu;
  const int n_warps = gridDim.x * F_WARPS;
  for (int r = blockIdx.x * F_WARPS + w; r < P.n_rows_class; r += n_warps) {
    const int row_l = P.row_list[r];
    const int64_t row = P.row_begin + row_l;
    {
      uint4 z = make_uint4(0, 0, 0, 0);
      uint4 *c4 = reinterpret_cast<uint4 *>(cnt);
      for (int s = lane; s < (1 << F_CW_LOG) / 32; s += 32) c4[s] = z;
    }
    __syncwarp();
    const int32_t *nbr = P.idx0 + row * P.k;
    int nslots = 0, npend = 0;
    bool ovf = false, sat = false;
    for (int pass = 0; pass < 2 && !ovf; ++pass) {
      if (pass == 1) {
        __syncwarp();
        if (__any_sync(0xffffffffu, sat)) {
          ovf = true;
          break;
        }
      }
      for (int p0 = 0; p0 < P.k && !ovf; p0 += 32) {
        int64_t beg = 0;
        int len = 0;
        if (p0 + lane < P.k) {
          const int32_t c = nbr[p0 + lane];
          beg = P.colptr[c];
          len = (int)(P.colptr[c + 1] - beg);
        }
        const int np = (P.k - p0) < 32 ? (P.k - p0) : 32;
        // walk (list, 128-element block) pairs; the next block's loads are issued before the
        // current block is processed, so the L2/HBM latency of the gathers is overlapped
        int pl = 0, t0 = 0;
        int64_t b = __shfl_sync(0xffffffffu, beg, 0);
        int l = __shfl_sync(0xffffffffu, len, 0);
        int32_t v[4], nv[4];
        const int lowkey = P.sym ? (int)row : -1;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          v[u] = (32 * u + lane < l) ? P.rev[b + 32 * u + lane] : -1;
          if (v[u] <= lowkey) v[u] = -1;
        }
        while (pl < np && !ovf) {
          int pn = pl, tn = t0 + 128;
          int64_t bn = b;
          int ln = l;
          if (tn >= l) {
            pn = pl + 1;
            tn = 0;
            if (pn < np) {
              bn = __shfl_sync(0xffffffffu, beg, pn);
              ln = __shfl_sync(0xffffffffu, len, pn);
            }
          }
          if (pn < np) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              nv[u] = (tn + 32 * u + lane < ln) ? P.rev[bn + tn + 32 * u + lane] : -1;
              if (nv[u] <= lowkey) nv[u] = -1;
            }
          }
          uint32_t hc[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) hc[u] = hash_cnt<CW_LOG>((uint32_t)v[u]);
          if (pass == 0) {
            // eight 4-bit counters per 32-bit word, bumped with one shared-memory atomic each
            // (colliding lanes are serialised by the hardware; no conflict detection, no warp
            // synchronisation).  A counter that already reached s_min is left alone, so only
            // simultaneous bumps can push one past 15; the lane that wraps a counter sees it in
            // the returned word and the row is handed to the big-table kernel instead.
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (t0 + 32 * u >= l) break;  // warp-uniform
              if (v[u] >= 0) {
                const uint32_t sh = (hc[u] & 7u) * 4u;
                volatile uint32_t *wp = cnt + (hc[u] >> 3);
                if ((int)((*wp >> sh) & 15u) < s_min) {
                  const uint32_t old = atomicAdd(cnt + (hc[u] >> 3), 1u << sh);
                  sat |= ((old >> sh) & 15u) == 15u;
                }
              }
            }
          } else {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (t0 + 32 * u >= l) break;  // warp-uniform
              const bool pass_f = v[u] >= 0 && (int)((cnt[hc[u] >> 3] >> ((hc[u] & 7u) * 4u)) & 15u) >= s_min;
              const unsigned bal = __ballot_sync(0xffffffffu, pass_f);
              if (bal) {
                if (pass_f) pend[npend + __popc(bal & ((1u << lane) - 1))] = (uint32_t)v[u];
                npend += __popc(bal);
                __syncwarp();
                if (npend >= 32) {
                  filt_insert_batch<F_TX_LOG>(tab, slots, nslots, P.cbits, pend[lane], true, lane);
                  const uint32_t mv = (32 + lane < npend) ? pend[32 + lane] : 0;
                  __syncwarp();
                  if (32 + lane < npend) pend[lane] = mv;
                  npend -= 32;
                  __syncwarp();
                  if (nslots > F_SLOTS) ovf = true;
                }
              }
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) v[u] = nv[u];
          pl = pn;
          t0 = tn;
          b = bn;
          l = ln;
        }
      }
    }
    if (!ovf && npend > 0) {
      filt_insert_batch<F_TX_LOG>(tab, slots, nslots, P.cbits, lane < npend ? pend[lane] : 0, lane < npend, lane);
      if (nslots > F_SLOTS) ovf = true;
    }
    __syncwarp();
    if (ovf) {
      // hand the row to the big-table kernel; leave the table clean
      for (int s = lane; s < (1 << F_TX_LOG); s += 32) tab[s] = 0xffffffffu;
      if (lane == 0) {
        P.row_nnz[row_l] = 0;
        ovf_list[atomicAdd(&info->n_ovf, 1)] = row_l;
      }
      __syncwarp();
      continue;
    }
    // survivors: exact counts against the cut-off, compacted into the (now free) counter area
    uint32_t *out = reinterpret_cast<uint32_t *>(cnt);
    int m = 0;
    for (int t0 = 0; t0 < nslots; t0 += 32) {
      const int t = t0 + lane;
      uint32_t wv = 0;
      bool kp = false;
      if (t < nslots) {
        const int sl = slots[t];
        wv = tab[sl];
        tab[sl] = 0xffffffffu;
        kp = P.keep[wv & cmask] != 0;
      }
      const unsigned bal = __ballot_sync(0xffffffffu, kp);
      if (kp) out[m + __popc(bal & ((1u << lane) - 1))] = wv;
      m += __popc(bal);
    }
    __syncwarp();
    bitonic_sort_asc(out, m, lane, 32, [] { __syncwarp(); });
    uint32_t *dst = reinterpret_cast<uint32_t *>(P.tmp) + P.tmp_off[row_l];
    for (int t = lane; t < m; t += 32) dst[t] = out[t];
    if (lane == 0) P.row_nnz[row_l] = m;
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------
// Walker over the k reverse lists of a row (k <= 32: lane p holds the descriptor of list p), in
// 128-element blocks; offsets are 32-bit (the transpose has fewer than 2^31 entries).
// ---------------------------------------------------------------------------
struct ListWalk {
  uint32_t beg;  // per lane: start of list `lane` in rev
  int len;       // per lane: its length
  int k;
  // warp-uniform cursor
  int pl, t0, l;
  uint32_t b;
  __device__ __forceinline__ void start() {
    pl = -1;
    t0 = 0;
    l = 0;
    b = 0;
  }
  // next 128-element block; false when every list has been walked
  __device__ __forceinline__ bool advance() {
    t0 += 128;
    while (t0 >= l) {
      if (++pl >= k) return false;
      t0 = 0;
      b = __shfl_sync(0xffffffffu, beg, pl);
      l = __shfl_sync(0xffffffffu, len, pl);
    }
    return true;
  }
  __device__ __forceinline__ void load(const int32_t *__restrict__ rev, int lane, int lowkey, int32_t (&v)[4]) const {
    const int32_t *src = rev + b + t0 + lane;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      v[u] = (t0 + 32 * u + lane < l) ? src[32 * u] : -1;
      if (v[u] <= lowkey) v[u] = -1;  // symmetric mode on unsorted lists: only candidates j > i
    }
  }
};

// ---------------------------------------------------------------------------
// Two-walk filtered kernel for k <= 32 (the default): pass A (4-bit saturating counters) and pass B
// (exact table for the occurrences whose counter reached s_min) as in k_snn_rows_filter.
// ncu (profiles/r02_ncu_k_snn_rows_filter2.txt): the kernel is instruction-issue bound (73 % issue
// utilisation, ~80 warp instructions per 32 occurrences and pass), not bound by the shared-memory
// atomics or by bandwidth, so what counts is how many occurrences are walked:
//   * symmetric mode (P.sym): only candidates j > i are counted.  With P.sorted the reverse lists
//     are sorted (the nn Graph's columns), and since row i is itself a member of each of its k
//     lists, a binary search for i gives the start of the suffix j > i: half of every list is
//     neither loaded nor looped over.
// ---------------------------------------------------------------------------
template <int CW_LOG, int F_TX_LOG, int F_WARPS>
__global__ void __launch_bounds__(F_WARPS * 32)
k_snn_rows_filter2(SnnParams P, int s_min, int32_t *__restrict__ ovf_list, DevInfo *__restrict__ info) {
  constexpr int F_BYTES = filt_bytes(CW_LOG, F_TX_LOG);
  constexpr int F_SLOTS = filt_slots(F_TX_LOG), F_SLOT_CAP = filt_slot_cap(F_TX_LOG);
  static_assert((1 << CW_LOG) / 2 >= F_SLOT_CAP * 4, "the counter area doubles as the survivor buffer");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char *base = smem_raw + (size_t)w * F_BYTES;
  uint32_t *cnt = reinterpret_cast<uint32_t *>(base);  // eight 4-bit counters per word
  uint32_t *tab = reinterpret_cast<uint32_t *>(base + (1 << CW_LOG) / 2);
  uint16_t *slots = reinterpret_cast<uint16_t *>(tab + (1 << F_TX_LOG));
  uint32_t *pend = reinterpret_cast<uint32_t *>(slots + F_SLOT_CAP);
  const uint32_t cmask = (1u << P.cbits) - 1;
  const int k = P.k;
  for (int s = lane; s < (1 << F_TX_LOG); s += 32) tab[s] = 0xffffffffu;
  const int n_warps = gridDim.x * F_WARPS;
  for (int r = blockIdx.x * F_WARPS + w; r < P.n_rows_class; r += n_warps) {
    const int row_l = P.row_list[r];
    const int64_t row = P.row_begin + row_l;
    // per-element filter only when the lists are unsorted; sorted lists are cut at the row itself
    const int lowkey = (P.sym && !P.sorted) ? (int)row : -1;
    {
      uint4 z = make_uint4(0, 0, 0, 0);
      uint4 *c4 = reinterpret_cast<uint4 *>(cnt);
      for (int s = lane; s < (1 << CW_LOG) / 32; s += 32) c4[s] = z;
    }
    __syncwarp();
    ListWalk W;
    W.k = k;
    W.beg = 0;
    W.len = 0;
    if (lane < k) {
      const int32_t c = P.idx0[row * k + lane];
      const int64_t cb = P.colptr[c];
      int len = (int)(P.colptr[c + 1] - cb);
      uint32_t beg = (uint32_t)cb;
      if (P.sym && P.sorted) {
        // first position whose row id exceeds `row` (the list is ascending and contains `row`)
        int lo = 0, hi = len;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (P.rev[beg + mid] <= (int32_t)row) lo = mid + 1;
          else hi = mid;
        }
        beg += (uint32_t)lo;
        len -= lo;
      }
      W.beg = beg;
      W.len = len;
    }
    int nslots = 0, npend = 0;
    bool ovf = false, sat = false;
    int32_t v[4], nv[4];
    // ---- pass A: count ----
    W.start();
    bool have = W.advance();
    if (have) W.load(P.rev, lane, lowkey, v);
    while (have) {
      const int t0 = W.t0, l = W.l;
      have = W.advance();
      if (have) W.load(P.rev, lane, lowkey, nv);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (t0 + 32 * u >= l) break;  // warp-uniform
        if (v[u] >= 0) {
          const uint32_t hc = hash_cnt<CW_LOG>((uint32_t)v[u]);
          const uint32_t sh = (hc & 7u) * 4u;
          volatile uint32_t *wp = cnt + (hc >> 3);
          if ((int)((*wp >> sh) & 15u) < s_min) {
            const uint32_t old = atomicAdd(cnt + (hc >> 3), 1u << sh);
            sat |= ((old >> sh) & 15u) == 15u;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = nv[u];
    }
    __syncwarp();
    if (__any_sync(0xffffffffu, sat)) ovf = true;  // a counter wrapped: the big-table kernel takes the row
    // ---- pass B: exact counts of the occurrences whose counter reached s_min ----
    if (!ovf) {
      W.start();
      have = W.advance();
      if (have) W.load(P.rev, lane, lowkey, v);
      while (have && !ovf) {
        const int t0 = W.t0, l = W.l;
        have = W.advance();
        if (have) W.load(P.rev, lane, lowkey, nv);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (t0 + 32 * u >= l) break;  // warp-uniform
          bool pass_f = false;
          if (v[u] >= 0) {
            const uint32_t hc = hash_cnt<CW_LOG>((uint32_t)v[u]);
            pass_f = (int)((cnt[hc >> 3] >> ((hc & 7u) * 4u)) & 15u) >= s_min;
          }
          const unsigned bal = __ballot_sync(0xffffffffu, pass_f);
          if (bal) {
            if (pass_f) pend[npend + __popc(bal & ((1u << lane) - 1))] = (uint32_t)v[u];
            npend += __popc(bal);
            __syncwarp();
            if (npend >= 32) {
              filt_insert_batch<F_TX_LOG>(tab, slots, nslots, P.cbits, pend[lane], true, lane);
              const uint32_t mv = (32 + lane < npend) ? pend[32 + lane] : 0;
              __syncwarp();
              if (32 + lane < npend) pend[lane] = mv;
              npend -= 32;
              __syncwarp();
              if (nslots > F_SLOTS) ovf = true;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = nv[u];
      }
      if (!ovf && npend > 0) {
        filt_insert_batch<F_TX_LOG>(tab, slots, nslots, P.cbits, lane < npend ? pend[lane] : 0, lane < npend, lane);
        if (nslots > F_SLOTS) ovf = true;
      }
    }
    __syncwarp();
    if (ovf) {
      for (int s = lane; s < (1 << F_TX_LOG); s += 32) tab[s] = 0xffffffffu;
      if (lane == 0) {
        P.row_nnz[row_l] = 0;
        ovf_list[atomicAdd(&info->n_ovf, 1)] = row_l;
      }
      __syncwarp();
      continue;
    }
    // survivors: exact counts against the cut-off, compacted into the (now free) counter area
    uint32_t *out = reinterpret_cast<uint32_t *>(cnt);
    int m = 0;
    for (int t0 = 0; t0 < nslots; t0 += 32) {
      const int t = t0 + lane;
      uint32_t wv = 0;
      bool kp = false;
      if (t < nslots) {
        const int sl = slots[t];
        wv = tab[sl];
        tab[sl] = 0xffffffffu;
        kp = P.keep[wv & cmask] != 0;
      }
      const unsigned bal = __ballot_sync(0xffffffffu, kp);
      if (kp) out[m + __popc(bal & ((1u << lane) - 1))] = wv;
      m += __popc(bal);
    }
    __syncwarp();
    bitonic_sort_asc(out, m, lane, 32, [] { __syncwarp(); });
    uint32_t *dst = reinterpret_cast<uint32_t *>(P.tmp) + P.tmp_off[row_l];
    for (int t = lane; t < m; t += 32) dst[t] = out[t];
    if (lane == 0) P.row_nnz[row_l] = m;
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------
// Staged variant of k_snn_rows_filter2 (the default for k <= 32).  The profile of filter2 shows ~110 warp
// instructions per (list, pass) against ~25 per 32 occurrences of real work: the lists are short (45
// entries on average after the suffix cut), so per-list bookkeeping dominates and is paid in both passes.
// Here the row's occurrences are first copied into a flat shared-memory buffer (one lean copy loop per
// list); pass A and pass B then run over the flat array with full warps and no per-list logic.  Rows
// with more occurrences than the buffer holds are processed in chunks (and staged again for pass B).
// ---------------------------------------------------------------------------
constexpr int STAGE_CAP = 1024;  // occurrences staged per chunk (4 KB per warp)
constexpr int filt3_bytes(int cw_log, int tx_log) { return filt_bytes(cw_log, tx_log) + STAGE_CAP * 4; }

template <int CW_LOG, int F_TX_LOG, int F_WARPS>
__global__ void __launch_bounds__(F_WARPS * 32)
k_snn_rows_filter3(SnnParams P, int s_min, int32_t *__restrict__ ovf_list, DevInfo *__restrict__ info) {
  constexpr int F_BYTES = filt3_bytes(CW_LOG, F_TX_LOG);
  constexpr int F_SLOTS = filt_slots(F_TX_LOG), F_SLOT_CAP = filt_slot_cap(F_TX_LOG);
  static_assert((1 << CW_LOG) / 2 >= F_SLOT_CAP * 4, "the counter area doubles as the survivor buffer");
  static_assert(filt_bytes(CW_LOG, F_TX_LOG) % 16 == 0, "stage buffer alignment");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char *base = smem_raw + (size_t)w * F_BYTES;
  uint32_t *cnt = reinterpret_cast<uint32_t *>(base);  // eight 4-bit counters per word
  uint32_t *tab = reinterpret_cast<uint32_t *>(base + (1 << CW_LOG) / 2);
  uint16_t *slots = reinterpret_cast<uint16_t *>(tab + (1 << F_TX_LOG));
  uint32_t *pend = reinterpret_cast<uint32_t *>(slots + F_SLOT_CAP);
  int32_t *stage = reinterpret_cast<int32_t *>(base + filt_bytes(CW_LOG, F_TX_LOG));
  const uint32_t cmask = (1u << P.cbits) - 1;
  const int k = P.k;
  for (int s = lane; s < (1 << F_TX_LOG); s += 32) tab[s] = 0xffffffffu;
  const int n_warps = gridDim.x * F_WARPS;
  for (int r = blockIdx.x * F_WARPS + w; r < P.n_rows_class; r += n_warps) {
    const int row_l = P.row_list[r];
    const int64_t row = P.row_begin + row_l;
    // per-element filter only when the lists are unsorted; sorted lists are cut at the row itself
    const int lowkey = (P.sym && !P.sorted) ? (int)row : -1;
    {
      uint4 z = make_uint4(0, 0, 0, 0);
      uint4 *c4 = reinterpret_cast<uint4 *>(cnt);
      for (int s = lane; s < (1 << CW_LOG) / 32; s += 32) c4[s] = z;
    }
    // list descriptors: lane p < k holds (start, length) of the part of list p that is walked
    uint32_t beg = 0;
    int len = 0;
    if (lane < k) {
      const int32_t c = P.idx0[row * k + lane];
      const int64_t cb = P.colptr[c];
      len = (int)(P.colptr[c + 1] - cb);
      beg = (uint32_t)cb;
      if (P.sym && P.sorted) {
        // first position whose row id exceeds `row` (the list is ascending and contains `row`)
        int lo = 0, hi = len;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (P.rev[beg + mid] <= (int32_t)row) lo = mid + 1;
          else hi = mid;
        }
        beg += (uint32_t)lo;
        len -= lo;
      }
    }
    const int total = warp_sum_i32(len);
    const bool one_chunk = total <= STAGE_CAP;  // warp-uniform
    __syncwarp();
    int nslots = 0, npend = 0;
    bool ovf = false, sat = false;
    for (int pass = 0; pass < 2 && !ovf; ++pass) {
      if (pass == 1) {
        __syncwarp();
        if (__any_sync(0xffffffffu, sat)) {  // a counter wrapped: the big-table kernel takes the row
          ovf = true;
          break;
        }
      }
      int pl = 0, toff = 0;  // next list / offset inside it (warp-uniform)
      while (pl < k && !ovf) {
        // ---- stage up to STAGE_CAP occurrences (a one-chunk row keeps its pass-A staging for pass B)
        int cntc = 0;
        if (pass == 1 && one_chunk) {
          cntc = total;
          pl = k;
        } else {
          while (pl < k && cntc < STAGE_CAP) {
            const uint32_t b = __shfl_sync(0xffffffffu, beg, pl);
            const int l = __shfl_sync(0xffffffffu, len, pl);
            const int take = min(l - toff, STAGE_CAP - cntc);
            const int32_t *src = P.rev + b + toff;
            for (int t = lane; t < take; t += 32) {
              const int32_t x = src[t];
              stage[cntc + t] = x > lowkey ? x : -1;
            }
            cntc += take;
            toff += take;
            if (toff >= l) {
              ++pl;
              toff = 0;
            }
          }
          __syncwarp();
        }
        // ---- process the staged occurrences with full warps
        if (pass == 0) {
          for (int f0 = 0; f0 < cntc; f0 += 128) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int f = f0 + 32 * u + lane;
              if (f0 + 32 * u >= cntc) break;  // warp-uniform
              const int32_t x = f < cntc ? stage[f] : -1;
              if (x >= 0) {
                const uint32_t hc = hash_cnt<CW_LOG>((uint32_t)x);
                const uint32_t sh = (hc & 7u) * 4u;
                volatile uint32_t *wp = cnt + (hc >> 3);
                if ((int)((*wp >> sh) & 15u) < s_min) {
                  const uint32_t old = atomicAdd(cnt + (hc >> 3), 1u << sh);
                  sat |= ((old >> sh) & 15u) == 15u;
                }
              }
            }
          }
        } else {
          for (int f0 = 0; f0 < cntc && !ovf; f0 += 32) {
            const int f = f0 + lane;
            const int32_t x = f < cntc ? stage[f] : -1;
            bool pass_f = false;
            if (x >= 0) {
              const uint32_t hc = hash_cnt<CW_LOG>((uint32_t)x);
              pass_f = (int)((cnt[hc >> 3] >> ((hc & 7u) * 4u)) & 15u) >= s_min;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, pass_f);
            if (bal) {
              if (pass_f) pend[npend + __popc(bal & ((1u << lane) - 1))] = (uint32_t)x;
              npend += __popc(bal);
              __syncwarp();
              if (npend >= 32) {
                filt_insert_batch<F_TX_LOG>(tab, slots, nslots, P.cbits, pend[lane], true, lane);
                const uint32_t mv = (32 + lane < npend) ? pend[32 + lane] : 0;
                __syncwarp();
                if (32 + lane < npend) pend[lane] = mv;
                npend -= 32;
                __syncwarp();
                if (nslots > F_SLOTS) ovf = true;
              }
            }
          }
        }
        __syncwarp();  // the stage buffer may be refilled
      }
    }
    if (!ovf && npend > 0) {
      filt_insert_batch<F_TX_LOG>(tab, slots, nslots, P.cbits, lane < npend ? pend[lane] : 0, lane < npend, lane);
      if (nslots > F_SLOTS) ovf = true;
    }
    __syncwarp();
    if (ovf) {
      for (int s = lane; s < (1 << F_TX_LOG); s += 32) tab[s] = 0xffffffffu;
      if (lane == 0) {
        P.row_nnz[row_l] = 0;
        ovf_list[atomicAdd(&info->n_ovf, 1)] = row_l;
      }
      __syncwarp();
      continue;
    }
    // survivors: exact counts against the cut-off, compacted into the (now free) counter area
    uint32_t *out = reinterpret_cast<uint32_t *>(cnt);
    int m = 0;
    for (int t0 = 0; t0 < nslots; t0 += 32) {
      const int t = t0 + lane;
      uint32_t wv = 0;
      bool kp = false;
      if (t < nslots) {
        const int sl = slots[t];
        wv = tab[sl];
        tab[sl] = 0xffffffffu;
        kp = P.keep[wv & cmask] != 0;
      }
      const unsigned bal = __ballot_sync(0xffffffffu, kp);
      if (kp) out[m + __popc(bal & ((1u << lane) - 1))] = wv;
      m += __popc(bal);
    }
    __syncwarp();
    bitonic_sort_asc(out, m, lane, 32, [] { __syncwarp(); });
    uint32_t *dst = reinterpret_cast<uint32_t *>(P.tmp) + P.tmp_off[row_l];
    for (int t = lane; t < m; t += 32) dst[t] = out[t];
    if (lane == 0) P.row_nnz[row_l] = m;
    __syncwarp();
  }
}

template <int CW_LOG, int F_TX_LOG, int F_WARPS>
int launch_filter3(b200nn_ctx *ctx, const SnnParams &P, int s_min, int32_t *ovf_list, DevInfo *d_info) {
  auto kern = k_snn_rows_filter3<CW_LOG, F_TX_LOG, F_WARPS>;
  constexpr int bytes = filt3_bytes(CW_LOG, F_TX_LOG) * F_WARPS;
  static_assert(2 * (bytes + 1024) <= 228 * 1024, "two CTAs per SM must fit");
  CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  const int blocks = (int)std::min<int64_t>(ceil_div64(P.n_rows_class, F_WARPS), (int64_t)ctx->sm_count * 2);
  kern<<<blocks, F_WARPS * 32, bytes, ctx->stream>>>(P, s_min, ovf_list, d_info);
  KERNEL_CHECK(ctx);
  return B200NN_OK;
}

// expanded work, survivor bound and work class of every row
struct RowWorkParams {
  const int32_t *idx0;
  const int64_t *colptr;
  int64_t row_begin;
  int32_t n_rows;  // rows in [row_begin, row_end)
  int32_t k;
  int32_t s_min;   // smallest kept count (0: nothing is ever kept)
  int64_t n;
  int32_t class_lim[4];      // no row lists a neighbour twice
  int32_t class_lim_dup[4];  // some row does (info->has_dups, set by k_hist earlier on the stream)
  int32_t *row_bound;
  int32_t *class_list;  // 5 x n_rows
  DevInfo *info;
};

__global__ void __launch_bounds__(256) k_row_work(RowWorkParams P) {
  const int lane = threadIdx.x & 31;
  const int64_t rl = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool act = rl < P.n_rows;
  int64_t e = 0;
  if (act) {
    const int32_t *nbr = P.idx0 + (P.row_begin + rl) * P.k;
    for (int p = 0; p < P.k; ++p) {
      int32_t c = nbr[p];
      e += P.colptr[c + 1] - P.colptr[c];
    }
  }
  int64_t bound = 0;
  if (act && P.s_min > 0) {
    bound = e / P.s_min;
    if (bound > P.n) bound = P.n;
  }
  if (act) P.row_bound[rl] = (int32_t)bound;
  int cls = -1;
  const int32_t *lim = P.info->has_dups ? P.class_lim_dup : P.class_lim;
  if (act) cls = e <= lim[0] ? 0 : e <= lim[1] ? 1 : e <= lim[2] ? 2 : e <= lim[3] ? 3 : 4;
  // warp-aggregated append to the class lists
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    unsigned bal = __ballot_sync(0xffffffffu, cls == q);
    if (bal) {
      int base = 0;
      if (lane == __ffs(bal) - 1) base = atomicAdd(&P.info->class_count[q], __popc(bal));
      base = __shfl_sync(0xffffffffu, base, __ffs(bal) - 1);
      if (cls == q)
        P.class_list[(int64_t)q * P.n_rows + base + __popc(bal & ((1u << lane) - 1))] = (int32_t)rl;
    }
  }
  int emax = warp_max_i32((int)(e > INT_MAX ? INT_MAX : e));
  long long esum = warp_sum_i64(e), bsum = warp_sum_i64(bound);
  if (lane == 0) {
    if (emax > P.info->max_row_e) atomicMax(&P.info->max_row_e, emax);
    atomicAdd(reinterpret_cast<unsigned long long *>(&P.info->sum_row_e), (unsigned long long)esum);
    atomicAdd(reinterpret_cast<unsigned long long *>(&P.info->sum_bound), (unsigned long long)bsum);
  }
}

// ---------------------------------------------------------------------------
// Symmetric mode (all rows, no duplicated neighbours): the row kernels emit only the upper
// triangle (j > i); S = A A^T is symmetric, so the lower triangle is its mirror image and the
// diagonal is S_ii = k.  Half the candidate occurrences are walked.
// ---------------------------------------------------------------------------
// incoming (mirrored) entries per row
__global__ void __launch_bounds__(256)
k_low_count(const uint32_t *__restrict__ tmp, const int64_t *__restrict__ tmp_off, const int32_t *__restrict__ up_cnt,
            int cbits, int32_t n_rows, int32_t *__restrict__ low_cnt) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= n_rows) return;
  const uint32_t *src = tmp + tmp_off[r];
  const int m = up_cnt[r];
  for (int t = lane; t < m; t += 32) atomicAdd(&low_cnt[src[t] >> cbits], 1);
}

__global__ void __launch_bounds__(256)
k_row_total(const int32_t *__restrict__ low_cnt, const int32_t *__restrict__ up_cnt, int diag, int32_t n_rows,
            int32_t *__restrict__ total) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_rows) total[r] = low_cnt[r] + diag + up_cnt[r];
}

// upper entry (i -> j, s) becomes the word (i << cbits | s) in row j's (unsorted) lower segment
__global__ void __launch_bounds__(256)
k_low_scatter(const uint32_t *__restrict__ tmp, const int64_t *__restrict__ tmp_off, const int32_t *__restrict__ up_cnt,
              int cbits, int32_t n_rows, const int64_t *__restrict__ low_off, int32_t *__restrict__ low_cur,
              uint32_t *__restrict__ low_words) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= n_rows) return;
  const uint32_t *src = tmp + tmp_off[r];
  const int m = up_cnt[r];
  const uint32_t cmask = (1u << cbits) - 1;
  for (int t = lane; t < m; t += 32) {
    const uint32_t w = src[t];
    const uint32_t j = w >> cbits;
    const int pos = atomicAdd(&low_cur[j], 1);
    low_words[low_off[j] + pos] = ((uint32_t)r << cbits) | (w & cmask);
  }
}

// row i of the final CSR: sorted lower words, the diagonal, the upper words
__global__ void __launch_bounds__(256)
k_snn_fill_sym(const uint32_t *__restrict__ tmp, const int64_t *__restrict__ tmp_off, const int32_t *__restrict__ up_cnt,
               const uint32_t *__restrict__ low_sorted, const int64_t *__restrict__ low_off,
               const int64_t *__restrict__ snn_p, const double *__restrict__ jac, int cbits, int diag, int k,
               int32_t n_rows, int32_t *__restrict__ snn_i, double *__restrict__ snn_x) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= n_rows) return;
  const uint32_t cmask = (1u << cbits) - 1;
  int64_t ob = snn_p[r];
  const int64_t lb = low_off[r];
  const int lc = (int)(low_off[r + 1] - lb);
  for (int t = lane; t < lc; t += 32) {
    const uint32_t w = low_sorted[lb + t];
    snn_i[ob + t] = (int32_t)(w >> cbits);
    snn_x[ob + t] = jac[w & cmask];
  }
  ob += lc;
  if (diag) {
    if (lane == 0) {
      snn_i[ob] = (int32_t)r;
      snn_x[ob] = jac[k];
    }
    ob += 1;
  }
  const uint32_t *src = tmp + tmp_off[r];
  const int m = up_cnt[r];
  for (int t = lane; t < m; t += 32) {
    const uint32_t w = src[t];
    snn_i[ob + t] = (int32_t)(w >> cbits);
    snn_x[ob + t] = jac[w & cmask];
  }
}

// unpack survivors into the final CSR arrays
template <typename Word>
__global__ void __launch_bounds__(256)
k_snn_fill(const Word *__restrict__ tmp, const int64_t *__restrict__ tmp_off,
           const int64_t *__restrict__ snn_p, const double *__restrict__ jac, int cbits,
           int32_t n_rows, int32_t *__restrict__ snn_i, double *__restrict__ snn_x) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= n_rows) return;
  const int64_t ob = snn_p[r];
  const int m = (int)(snn_p[r + 1] - ob);
  const Word *src = tmp + tmp_off[r];
  const Word cmask = (((Word)1) << cbits) - 1;
  for (int t = lane; t < m; t += 32) {
    Word w = src[t];
    snn_i[ob + t] = (int32_t)(w >> cbits);
    snn_x[ob + t] = jac[(uint32_t)(w & cmask)];
  }
}

template <typename Word, int LOG_T, int G, int RPB>
int launch_class(b200nn_ctx *ctx, SnnParams P, int blocks_per_sm, int64_t max_blocks = 0) {
  if (P.n_rows_class <= 0) return B200NN_OK;
  constexpr bool HUGE = (LOG_T == 0);
  constexpr bool INPLACE = (G == 32) && !HUGE;
  size_t smem = 0;
  if (!HUGE) {
    size_t per_group = ((size_t)1 << LOG_T) + (INPLACE ? 0 : ((size_t)1 << LOG_T) / 2);
    smem = per_group * sizeof(Word) * RPB;
    CU_TRY(cudaFuncSetAttribute(k_snn_rows<Word, LOG_T, G, RPB>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  int64_t want = ceil_div64(P.n_rows_class, RPB);
  int64_t cap = (int64_t)ctx->sm_count * blocks_per_sm;
  if (max_blocks > 0 && cap > max_blocks) cap = max_blocks;
  int blocks = (int)std::min<int64_t>(want, cap);
  k_snn_rows<Word, LOG_T, G, RPB><<<blocks, G * RPB, smem, ctx->stream>>>(P);
  KERNEL_CHECK(ctx);
  return B200NN_OK;
}

int fetch_info(b200nn_ctx *ctx, const DevInfo *d_info);

// table sizes (log2 words) of the three warp-per-row classes and the block-per-row class
struct WordCfg32 {
  typedef uint32_t Word;
  static constexpr int L0 = 11, L1 = 12, L2 = 13, L3 = 15;
};
struct WordCfg64 {
  typedef uint64_t Word;
  static constexpr int L0 = 10, L1 = 11, L2 = 12, L3 = 14;
};
// a table of T words takes rows with at most 0.8 T expanded candidates (distinct keys <= that)
static inline int32_t class_limit(int log_t) { return (int32_t)(((int64_t)1 << log_t) * 4 / 5); }

constexpr int HUGE_BLOCKS_PER_SM = 1;

template <typename Cfg>
int run_snn_classes(b200nn_ctx *ctx, SnnParams P, const int32_t *class_list, int32_t n_rows, int64_t n_cols,
                    const int32_t *class_count, int max_row_e, bool has_dups, int filter_s_min,
                    int32_t *ovf_list, DevInfo *d_info) {
  typedef typename Cfg::Word Word;
  if (filter_s_min >= 2) {
    // classes 0-2 through the filtered kernel; overflowing rows are re-run below
    // class 0 rows (E <= 3.2 k) need only a 512-word exact table: 15 rows per CTA, 30 warps per SM;
    // classes 1-2: 8 rows per CTA, 16 warps per SM (shared memory bound)
    // (log2 counters, log2 exact-table words, rows per CTA) per class; two CTAs per SM.  These
    // kernels live on occupancy (dependent shared-memory round trips) and on few false
    // positives in the counters; measured best: class 0 (E <= 2.8 k) 8 k counters, 30 warps/SM;
    // class 1 (E <= 3.2 k) 16 k counters, 20 warps/SM; class 2 (E <= 6.5 k) 16 k counters, 16.
    constexpr int WA = 15, WB = 10, WC = 8;
    // k <= 32 (one reverse list per lane): k_snn_rows_filter2; B200NN_SNN_ROUND1=1 selects the round-1 kernel
    static const bool old_env = getenv("B200NN_SNN_ROUND1") != nullptr;  // round-1 kernel (any k)
    // the staged variant (flat shared-memory occurrence buffer) measured slower on config C
    // (15.1 ms vs 11.6 ms for the stage, gpurun_out r02e vs profiles/r02_bench_C.json): opt-in only
    static const bool staged_env = getenv("B200NN_SNN_STAGED") != nullptr;
    if (P.k <= 32 && !old_env && staged_env) {
      // (counters, exact table, warps per CTA): two CTAs per SM, 113 KB each
      for (int c = 0; c < 3; ++c) {
        P.row_list = class_list + c * (int64_t)n_rows;
        P.n_rows_class = class_count[c];
        if (P.n_rows_class <= 0) continue;
        if (c == 0)
          RC_TRY((launch_filter3<13, 9, 113 * 1024 / filt3_bytes(13, 9)>(ctx, P, filter_s_min, ovf_list, d_info)));
        else if (c == 1)
          RC_TRY((launch_filter3<14, 9, 113 * 1024 / filt3_bytes(14, 9)>(ctx, P, filter_s_min, ovf_list, d_info)));
        else
          RC_TRY((launch_filter3<14, 10, 113 * 1024 / filt3_bytes(14, 10)>(ctx, P, filter_s_min, ovf_list, d_info)));
      }
    } else if (P.k <= 32 && !old_env) {
      auto kA = k_snn_rows_filter2<13, 9, WA>;
      auto kB = k_snn_rows_filter2<14, 9, WB>;
      auto kC = k_snn_rows_filter2<14, 10, WC>;
      constexpr int bytesA = filt_bytes(13, 9) * WA, bytesB = filt_bytes(14, 9) * WB, bytesC = filt_bytes(14, 10) * WC;
      CU_TRY(cudaFuncSetAttribute(kA, cudaFuncAttributeMaxDynamicSharedMemorySize, bytesA));
      CU_TRY(cudaFuncSetAttribute(kB, cudaFuncAttributeMaxDynamicSharedMemorySize, bytesB));
      CU_TRY(cudaFuncSetAttribute(kC, cudaFuncAttributeMaxDynamicSharedMemorySize, bytesC));
      for (int c = 0; c < 3; ++c) {
        P.row_list = class_list + c * (int64_t)n_rows;
        P.n_rows_class = class_count[c];
        if (P.n_rows_class <= 0) continue;
        const int w = c == 0 ? WA : c == 1 ? WB : WC;
        const int blocks = (int)std::min<int64_t>(ceil_div64(P.n_rows_class, w), (int64_t)ctx->sm_count * 2);
        if (c == 0)
          kA<<<blocks, WA * 32, bytesA, ctx->stream>>>(P, filter_s_min, ovf_list, d_info);
        else if (c == 1)
          kB<<<blocks, WB * 32, bytesB, ctx->stream>>>(P, filter_s_min, ovf_list, d_info);
        else
          kC<<<blocks, WC * 32, bytesC, ctx->stream>>>(P, filter_s_min, ovf_list, d_info);
        KERNEL_CHECK(ctx);
      }
    } else {
      auto kA = k_snn_rows_filter<13, 9, WA>;
      auto kB = k_snn_rows_filter<14, 9, WB>;
      auto kC = k_snn_rows_filter<14, 10, WC>;
      constexpr int bytesA = filt_bytes(13, 9) * WA, bytesB = filt_bytes(14, 9) * WB, bytesC = filt_bytes(14, 10) * WC;
      CU_TRY(cudaFuncSetAttribute(kA, cudaFuncAttributeMaxDynamicSharedMemorySize, bytesA));
      CU_TRY(cudaFuncSetAttribute(kB, cudaFuncAttributeMaxDynamicSharedMemorySize, bytesB));
      CU_TRY(cudaFuncSetAttribute(kC, cudaFuncAttributeMaxDynamicSharedMemorySize, bytesC));
      for (int c = 0; c < 3; ++c) {
        P.row_list = class_list + c * (int64_t)n_rows;
        P.n_rows_class = class_count[c];
        if (P.n_rows_class <= 0) continue;
        const int w = c == 0 ? WA : c == 1 ? WB : WC;
        const int blocks = (int)std::min<int64_t>(ceil_div64(P.n_rows_class, w), (int64_t)ctx->sm_count * 2);
        if (c == 0)
          kA<<<blocks, WA * 32, bytesA, ctx->stream>>>(P, filter_s_min, ovf_list, d_info);
        else if (c == 1)
          kB<<<blocks, WB * 32, bytesB, ctx->stream>>>(P, filter_s_min, ovf_list, d_info);
        else
          kC<<<blocks, WC * 32, bytesC, ctx->stream>>>(P, filter_s_min, ovf_list, d_info);
        KERNEL_CHECK(ctx);
      }
    }
  } else {
    P.row_list = class_list + 0 * (int64_t)n_rows;
    P.n_rows_class = class_count[0];
    RC_TRY((launch_class_warp<Word, Cfg::L0, 8>(ctx, P, 3, has_dups)));
    P.row_list = class_list + 1 * (int64_t)n_rows;
    P.n_rows_class = class_count[1];
    RC_TRY((launch_class_warp<Word, Cfg::L1, 4>(ctx, P, 3, has_dups)));
    P.row_list = class_list + 2 * (int64_t)n_rows;
    P.n_rows_class = class_count[2];
    RC_TRY((launch_class_warp<Word, Cfg::L2, 2>(ctx, P, 3, has_dups)));
  }
  P.row_list = class_list + 3 * (int64_t)n_rows;
  P.n_rows_class = class_count[3];
  RC_TRY((launch_class<Word, Cfg::L3, 256, 1>(ctx, P, 1)));
  if (class_count[4] > 0) {
    // a row has at most n distinct candidates, however large its expanded work E is (hubs,
    // duplicated cells): 2 * min(E, n) words per table suffice
    const int64_t distinct_max = std::min<int64_t>((int64_t)max_row_e, n_cols);
    int log_t = 1;
    while (((int64_t)1 << log_t) < 2 * distinct_max) ++log_t;
    int64_t blocks = std::min<int64_t>(class_count[4], (int64_t)ctx->sm_count * HUGE_BLOCKS_PER_SM);
    Word *tables = nullptr;
    int rc_t = ws_get(ctx, WS_HUGE_TABLE, (size_t)blocks << log_t, &tables);
    while (rc_t == B200NN_ERR_NOMEM && blocks > 1) {  // fewer concurrent hub rows instead of failing
      blocks = (blocks + 1) / 2;
      rc_t = ws_get(ctx, WS_HUGE_TABLE, (size_t)blocks << log_t, &tables);
    }
    RC_TRY(rc_t);
    P.huge_tables = tables;
    P.huge_log_t = log_t;
    P.row_list = class_list + 4 * (int64_t)n_rows;
    P.n_rows_class = class_count[4];
    RC_TRY((launch_class<Word, 0, 256, 1>(ctx, P, HUGE_BLOCKS_PER_SM, blocks)));
  }
  if (filter_s_min >= 2) {
    // rows whose exact table overflowed in the filtered kernel (E <= 0.8 * 2^L2 <= 2^L3 / 2)
    RC_TRY(fetch_info(ctx, d_info));
    const int32_t n_ovf = ctx->h_info->n_ovf;
    if (n_ovf > 0) {
      P.row_list = ovf_list;
      P.n_rows_class = n_ovf;
      RC_TRY((launch_class<Word, Cfg::L3, 256, 1>(ctx, P, 1)));
    }
  }
  return B200NN_OK;
}

// Small read-backs go through a kernel that stores into the context's page-locked host block (device-
// visible under unified addressing) instead of cudaMemcpyAsync: a D2H copy would queue on the copy engine
// behind the multi-100-MB result transfers of the end-to-end path (measured: +6 ms per call).
__global__ void k_info_to_host(const DevInfo *__restrict__ d_info, DevInfo *__restrict__ h_info) {
  const int t = threadIdx.x;
  const int32_t *src = reinterpret_cast<const int32_t *>(d_info);
  volatile int32_t *dst = reinterpret_cast<volatile int32_t *>(h_info);
  for (int i = t; i < (int)(sizeof(DevInfo) / sizeof(int32_t)); i += blockDim.x) dst[i] = src[i];
  __threadfence_system();
}
__global__ void k_i64_to_host(const int64_t *__restrict__ src, long long *__restrict__ h_dst) {
  *reinterpret_cast<volatile long long *>(h_dst) = (long long)*src;
  __threadfence_system();
}

int fetch_info(b200nn_ctx *ctx, const DevInfo *d_info) {
  k_info_to_host<<<1, 32, 0, ctx->stream>>>(d_info, ctx->h_info);
  KERNEL_CHECK(ctx);
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  return B200NN_OK;
}

// one 64-bit device value -> host, on `stream` (slot 0 / 1 of the context's host scalars)
int fetch_i64(b200nn_ctx *ctx, const int64_t *d_val, int64_t *out, cudaStream_t stream, int slot) {
  k_i64_to_host<<<1, 1, 0, stream>>>(d_val, &ctx->h_scalars[slot]);
  KERNEL_CHECK(ctx);
  CU_TRY(cudaStreamSynchronize(stream));
  *out = (int64_t)ctx->h_scalars[slot];
  return B200NN_OK;
}

}  // namespace

// host-side evaluation of the reference's per-entry expression (snn.cpp:30-33,36)
static inline double jaccard_of(double s, double k) { return s / (k + (k - s)); }

int graphs_launch(b200nn_ctx *ctx, const int32_t *d_idx0, int64_t n_rows, int k, int64_t n,
                  double prune, bool want_nn_graph, bool compute_snn, int64_t row_begin,
                  int64_t row_end, b200nn_stats *stats, const NnGraphHostOut *early_nn) {
  cudaStream_t st = ctx->stream;
  ctx->nn_graph.valid = false;
  ctx->snn.valid = false;
  if (n_rows < 0 || n <= 0 || k <= 0) {
    b200nn_set_error("graphs: invalid shape n_rows=%lld n=%lld k=%d", (long long)n_rows,
                     (long long)n, k);
    return B200NN_ERR_INVALID;
  }
  const int64_t total = n_rows * (int64_t)k;
  if (total >= 0x7fffffffLL || n >= 0x7fffffffLL) {
    b200nn_set_error("nn Graph with %lld entries does not fit a dgCMatrix", (long long)total);
    return B200NN_ERR_NNZ_OVERFLOW;
  }
  if (compute_snn && n_rows != n) {
    b200nn_set_error("ComputeSNN needs a square neighbour matrix (n_rows=%lld, n=%lld)",
                     (long long)n_rows, (long long)n);
    return B200NN_ERR_INVALID;
  }
  if (row_begin < 0 || row_end > n_rows || row_begin > row_end) {
    b200nn_set_error("graphs: invalid row range [%lld, %lld)", (long long)row_begin,
                     (long long)row_end);
    return B200NN_ERR_INVALID;
  }

  cudaEvent_t e0 = ctx->ev[0], e1 = ctx->ev[1], e2 = ctx->ev[2];
  CU_TRY(cudaEventRecord(e0, st));

  DevInfo *d_info;
  int32_t *indeg, *cursor, *rev, *uniq, *long_list;
  int64_t *colptr;
  RC_TRY(ws_get(ctx, WS_INFO, 1, &d_info));
  RC_TRY(ws_get(ctx, WS_INDEG, (size_t)n + 1, &indeg));
  RC_TRY(ws_get(ctx, WS_CURSOR, (size_t)n + 1, &cursor));
  RC_TRY(ws_get(ctx, WS_COLPTR, (size_t)n + 1, &colptr));
  RC_TRY(ws_get(ctx, WS_REV, (size_t)total + 1, &rev));
  RC_TRY(ws_get(ctx, WS_UNIQ, (size_t)n + 1, &uniq));
  RC_TRY(ws_get(ctx, WS_LONG_LIST, (size_t)(total / SORT_WARP_CAP) + 2, &long_list));
  CU_TRY(cudaMemsetAsync(d_info, 0, sizeof(DevInfo), st));
  CU_TRY(cudaMemsetAsync(indeg, 0, sizeof(int32_t) * ((size_t)n + 1), st));
  CU_TRY(cudaMemsetAsync(cursor, 0, sizeof(int32_t) * ((size_t)n + 1), st));

  // ---- K3 ----
  // Main stream: histogram, scan, scatter -- all the SNN stage needs (unsorted reverse lists).
  // Side stream: the nn Graph proper (sorted columns, duplicates merged, emission, early copy to
  // the caller's buffers) runs concurrently with the SNN stage.
  if (total > 0) {
    k_hist<<<(unsigned)ceil_div64(total, 256), 256, 0, st>>>(d_idx0, total, k, n, indeg, d_info);
    KERNEL_CHECK(ctx);
  }
  RC_TRY(scan_exclusive_i32_to_i64(ctx, indeg, colptr, n));
  k_max_indeg<<<ctx->sm_count, 256, 0, st>>>(indeg, n, d_info);
  KERNEL_CHECK(ctx);
  if (total > 0) {
    k_scatter<<<(unsigned)ceil_div64(total, 256), 256, 0, st>>>(d_idx0, total, k, n, colptr, cursor,
                                                                 rev);
    KERNEL_CHECK(ctx);
  }
  int64_t *nn_p = nullptr;
  int32_t *nn_i = nullptr;
  double *nn_x = nullptr;
  int32_t *rev_sorted = nullptr;
  static const bool no_sym_env = getenv("B200NN_SNN_NO_SYM") != nullptr;
  // the symmetric SNN mode (all rows) cuts every sorted list at the row itself: it wants the sorted
  // columns too, and waits for them (only for them) before its row kernels start
  const bool sym_possible = compute_snn && row_begin == 0 && row_end == n && n > 1 && !no_sym_env;
  const bool need_sort = want_nn_graph || sym_possible;
  cudaStream_t ss = ctx->side_stream;
  if (need_sort) {
    RC_TRY(ws_get(ctx, WS_REV_SORTED, (size_t)total + 1, &rev_sorted));
    CU_TRY(cudaEventRecord(ctx->ev_side, st));
    CU_TRY(cudaStreamWaitEvent(ss, ctx->ev_side, 0));
    k_sort_columns<int32_t><<<(unsigned)ceil_div64(n, SORT_WARPS), SORT_WARPS * 32, 0, ss>>>(
        rev, rev_sorted, colptr, n, uniq, long_list, &d_info->n_long);
    KERNEL_CHECK(ctx);
    {
      size_t smem = sizeof(int32_t) * LONG_SMEM_CAP;
      CU_TRY(cudaFuncSetAttribute(k_sort_long_columns<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
      k_sort_long_columns<int32_t><<<ctx->sm_count, LONG_THREADS, smem, ss>>>(rev, rev_sorted, colptr, uniq,
                                                                              long_list, &d_info->n_long);
      KERNEL_CHECK(ctx);
    }
    CU_TRY(cudaEventRecord(ctx->ev_sorted, ss));
  }
  if (want_nn_graph) {
    RC_TRY(ws_get(ctx, WS_NN_P, (size_t)n + 1, &nn_p));
    // the graph has at most `total` entries (fewer only when a row lists a neighbour twice)
    RC_TRY(ws_get(ctx, WS_NN_I, (size_t)total + 1, &nn_i));
    RC_TRY(ws_get(ctx, WS_NN_X, (size_t)total + 1, &nn_x));
    RC_TRY(scan_exclusive_i32_to_i64(ctx, uniq, nn_p, n, ss, WS_SCAN_TMP2));
    k_emit_nn_graph<<<(unsigned)ceil_div64(n * 32, 256), 256, 0, ss>>>(rev_sorted, colptr, uniq, nn_p, n,
                                                                        nn_i, nn_x);
    KERNEL_CHECK(ctx);
    if (early_nn && (early_nn->p || early_nn->i || early_nn->x)) {
      // the caller's nn Graph buffers (capacity n_rows * k) are known already: they are filled on the
      // copy stream while the SNN stage runs; entries past nnz_nn are unspecified
      int32_t *p32 = nullptr;
      if (early_nn->p) {
        RC_TRY(ws_get(ctx, WS_OUT_P32, (size_t)n + 1, &p32));
        RC_TRY(narrow_i64_to_i32(ctx, nn_p, p32, n + 1, ss));
      }
      CU_TRY(cudaEventRecord(ctx->ev_copy, ss));
      CU_TRY(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_copy, 0));
      if (early_nn->p)
        CU_TRY(cudaMemcpyAsync(early_nn->p, p32, sizeof(int32_t) * ((size_t)n + 1), cudaMemcpyDeviceToHost,
                               ctx->copy_stream));
      if (early_nn->i && total)
        CU_TRY(cudaMemcpyAsync(early_nn->i, nn_i, sizeof(int32_t) * (size_t)total, cudaMemcpyDeviceToHost,
                               ctx->copy_stream));
      if (early_nn->x && total)
        CU_TRY(cudaMemcpyAsync(early_nn->x, nn_x, sizeof(double) * (size_t)total, cudaMemcpyDeviceToHost,
                               ctx->copy_stream));
    }
  }
  CU_TRY(cudaEventRecord(e1, st));
  DevInfo hi;
  memset(&hi, 0, sizeof(hi));

  // ---- K4 ----
  if (compute_snn) {
    const int64_t rows = row_end - row_begin;
    if (rows >= 0x7fffffffLL) return B200NN_ERR_INVALID;
    // lookup tables for BOTH count ranges (no duplicates: s <= k; duplicated neighbours: s <= k^2), so
    // that nothing has to be read back before the per-row work is known; huge k asks first
    DevInfo early;
    memset(&early, 0, sizeof(early));
    const bool lut_all = (int64_t)k * k <= 65536;
    if (!lut_all) {
      RC_TRY(fetch_info(ctx, d_info));
      early = *ctx->h_info;
    }
    const int64_t lut_max = (lut_all || early.has_dups) ? (int64_t)k * k : (int64_t)k;
    if (lut_max > (1 << 20)) {
      if (need_sort) CU_TRY(cudaStreamSynchronize(ctx->side_stream));
      b200nn_set_error("k = %d with duplicated neighbours exceeds the supported count range", k);
      return B200NN_ERR_INVALID;
    }
    // lookup tables from the exact double expression of the reference (host copies live in the
    // context so that the asynchronous upload needs no synchronisation)
    std::vector<uint8_t> &h_keep = ctx->h_keep;
    std::vector<double> &h_jac = ctx->h_jac;
    h_keep.assign((size_t)lut_max + 1, 0);
    h_jac.assign((size_t)lut_max + 1, 0.0);
    int32_t s_min = 0;
    for (int64_t s = 0; s <= lut_max; ++s) {
      double v = jaccard_of((double)s, (double)k);
      if (v < prune) v = 0;
      bool kp = (s > 0) && (v != 0);  // s == 0 entries are never stored by the product
      h_keep[(size_t)s] = kp ? 1 : 0;
      h_jac[(size_t)s] = v;
      if (kp && s_min == 0) s_min = (int32_t)s;
    }
    uint8_t *d_keep;
    double *d_jac;
    RC_TRY(ws_get(ctx, WS_LUT_KEEP, (size_t)lut_max + 1, &d_keep));
    RC_TRY(ws_get(ctx, WS_LUT_JAC, (size_t)lut_max + 1, &d_jac));
    CU_TRY(cudaMemcpyAsync(d_keep, h_keep.data(), h_keep.size(), cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(d_jac, h_jac.data(), h_jac.size() * sizeof(double),
                           cudaMemcpyHostToDevice, st));

    int32_t *row_bound, *class_list, *row_nnz;
    int64_t *tmp_off, *snn_p;
    RC_TRY(ws_get(ctx, WS_ROW_BOUND, (size_t)rows + 1, &row_bound));
    RC_TRY(ws_get(ctx, WS_CLASS_LIST, (size_t)rows * 5 + 8, &class_list));
    RC_TRY(ws_get(ctx, WS_ROW_NNZ, (size_t)rows + 1, &row_nnz));
    RC_TRY(ws_get(ctx, WS_TMP_OFF, (size_t)rows + 1, &tmp_off));
    RC_TRY(ws_get(ctx, WS_SNN_P, (size_t)rows + 1, &snn_p));

    RowWorkParams rw;
    rw.idx0 = d_idx0;
    rw.colptr = colptr;
    rw.row_begin = row_begin;
    rw.n_rows = (int32_t)rows;
    rw.k = k;
    rw.s_min = s_min;
    rw.n = n;
    // word format of the survivors: key << cbits | count.  The class limits are those of the 32-bit
    // configuration whenever the key fits; a row that lists a neighbour twice (known only after the
    // read-back below) may force 64-bit words, whose smaller tables take the same classes at half
    // the limits -- those are applied inside k_row_work through has_dups.
    int kbits = 1;
    while (((int64_t)1 << kbits) < n) ++kbits;
    const bool fits32_nodup = kbits < 32 && (int64_t)k <= (((int64_t)1 << (32 - kbits)) - 2);
    const bool fits32_dup = kbits < 32 && (int64_t)k * k <= (((int64_t)1 << (32 - kbits)) - 2);
    // the filtered kernels need 32-bit words, no duplicated neighbours, byte-sized counts
    // and a cut-off that actually removes single overlaps
    const int filter_if_nodup = (fits32_nodup && k <= 255 && s_min >= 2 && s_min <= F_SMIN_MAX) ? s_min : 0;
    // Class 0 is bounded by the 2^L0-word table of the plain warp kernel; the filtered kernels'
    // tables are independent of it (measured best split for them: E <= 2.8 k)
    for (int v = 0; v < 2; ++v) {  // v = 0: no duplicated neighbours, v = 1: some
      const bool u32 = v == 0 ? fits32_nodup : fits32_dup;
      int32_t *lim = v == 0 ? rw.class_lim : rw.class_lim_dup;
      lim[0] = (v == 0 && filter_if_nodup >= 2) ? 2800 : class_limit(u32 ? WordCfg32::L0 : WordCfg64::L0);
      lim[1] = class_limit(u32 ? WordCfg32::L1 : WordCfg64::L1);
      lim[2] = class_limit(u32 ? WordCfg32::L2 : WordCfg64::L2);
      // the block-per-row kernel keeps survivors in a buffer of T/2 words: E <= T/2 there
      lim[3] = (1 << (u32 ? WordCfg32::L3 : WordCfg64::L3)) / 2;
    }
    rw.row_bound = row_bound;
    rw.class_list = class_list;
    rw.info = d_info;
    if (rows > 0) {
      k_row_work<<<(unsigned)ceil_div64(rows, 256), 256, 0, st>>>(rw);
      KERNEL_CHECK(ctx);
    }
    RC_TRY(scan_exclusive_i32_to_i64(ctx, row_bound, tmp_off, rows));
    // the one read-back before the row kernels: flags of the transpose (index range, duplicated
    // neighbours), class sizes and the survivor-buffer bound
    RC_TRY(fetch_info(ctx, d_info));
    hi = *ctx->h_info;
    if (hi.err_index_range) {
      if (need_sort) CU_TRY(cudaStreamSynchronize(ctx->side_stream));
      b200nn_set_error("neighbour index out of range: every entry must lie in [1, %lld]", (long long)n);
      return B200NN_ERR_INDEX_RANGE;
    }
    const bool has_dups = hi.has_dups != 0;
    const int64_t smax = has_dups ? (int64_t)k * k : (int64_t)k;
    if (smax > lut_max) {
      if (need_sort) CU_TRY(cudaStreamSynchronize(ctx->side_stream));
      b200nn_set_error("k = %d with duplicated neighbours exceeds the supported count range", k);
      return B200NN_ERR_INVALID;
    }
    const bool use32 = has_dups ? fits32_dup : fits32_nodup;
    const int cbits = use32 ? 32 - kbits : 32;
    const size_t wbytes = use32 ? 4 : 8;
    const int filter_s_min = has_dups ? 0 : filter_if_nodup;
    void *tmp;
    RC_TRY(ws_reserve(ctx, WS_TMP, ((size_t)hi.sum_bound + 1) * wbytes, &tmp));
    CU_TRY(cudaMemsetAsync(row_nnz, 0, sizeof(int32_t) * ((size_t)rows + 1), st));

    SnnParams P;
    P.idx0 = d_idx0;
    P.colptr = colptr;
    P.rev = rev;
    P.keep = d_keep;
    P.row_list = nullptr;
    P.n_rows_class = 0;
    P.k = k;
    P.cbits = cbits;
    P.row_begin = row_begin;
    P.tmp_off = tmp_off;
    P.tmp = tmp;
    P.row_nnz = row_nnz;
    P.huge_tables = nullptr;
    P.huge_log_t = 0;
    // all rows, 0/1 adjacency, packed 32-bit words: count only j > i and mirror (B200NN_SNN_NO_SYM=1: off)
    P.sym = (sym_possible && !has_dups && use32) ? 1 : 0;
    P.sorted = 0;
    if (P.sym) {
      static const bool unsorted_env = getenv("B200NN_SNN_SYM_UNSORTED") != nullptr;  // per-element filter instead
      if (!unsorted_env) {
        CU_TRY(cudaStreamWaitEvent(st, ctx->ev_sorted, 0));
        P.rev = rev_sorted;
        P.sorted = 1;
      }
    }
    int32_t *ovf_list;
    RC_TRY(ws_get(ctx, WS_OVF_LIST, (size_t)rows + 1, &ovf_list));
    if (use32)
      RC_TRY(run_snn_classes<WordCfg32>(ctx, P, class_list, (int32_t)rows, n, hi.class_count,
                                        hi.max_row_e, has_dups, filter_s_min, ovf_list, d_info));
    else
      RC_TRY(run_snn_classes<WordCfg64>(ctx, P, class_list, (int32_t)rows, n, hi.class_count,
                                        hi.max_row_e, has_dups, 0, ovf_list, d_info));
    int64_t nnz = 0;
    int32_t *snn_i;
    double *snn_x;
    if (!P.sym) {
      RC_TRY(scan_exclusive_i32_to_i64(ctx, row_nnz, snn_p, rows));
      // read-back: nnz (sizes the result arrays)
      RC_TRY(fetch_i64(ctx, snn_p + rows, &nnz, st, 0));
      RC_TRY(ws_get(ctx, WS_SNN_I, (size_t)nnz + 1, &snn_i));
      RC_TRY(ws_get(ctx, WS_SNN_X, (size_t)nnz + 1, &snn_x));
      if (rows > 0) {
        unsigned blocks = (unsigned)ceil_div64(rows * 32, 256);
        if (use32)
          k_snn_fill<uint32_t><<<blocks, 256, 0, st>>>(reinterpret_cast<uint32_t *>(tmp), tmp_off,
                                                       snn_p, d_jac, cbits, (int32_t)rows, snn_i,
                                                       snn_x);
        else
          k_snn_fill<uint64_t><<<blocks, 256, 0, st>>>(reinterpret_cast<uint64_t *>(tmp), tmp_off,
                                                       snn_p, d_jac, cbits, (int32_t)rows, snn_i,
                                                       snn_x);
        KERNEL_CHECK(ctx);
      }
    } else {
      // the row kernels emitted the upper triangle: mirror it (count, scatter, sort the short
      // lower segments), add the diagonal S_ii = k, assemble the rows
      const uint32_t *tmp32 = reinterpret_cast<const uint32_t *>(tmp);
      int32_t *low_cnt = indeg, *low_cur = cursor, *total_row = row_bound;  // free since the transpose
      const int diag = h_keep[(size_t)k] ? 1 : 0;
      const unsigned wblocks = (unsigned)ceil_div64(rows * 32, 256);
      CU_TRY(cudaMemsetAsync(low_cnt, 0, sizeof(int32_t) * ((size_t)n + 1), st));
      CU_TRY(cudaMemsetAsync(low_cur, 0, sizeof(int32_t) * ((size_t)n + 1), st));
      k_low_count<<<wblocks, 256, 0, st>>>(tmp32, tmp_off, row_nnz, cbits, (int32_t)rows, low_cnt);
      KERNEL_CHECK(ctx);
      k_row_total<<<(unsigned)ceil_div64(rows, 256), 256, 0, st>>>(low_cnt, row_nnz, diag, (int32_t)rows, total_row);
      KERNEL_CHECK(ctx);
      int64_t *low_off;
      RC_TRY(ws_get(ctx, WS_LOW_OFF, (size_t)rows + 1, &low_off));
      RC_TRY(scan_exclusive_i32_to_i64(ctx, low_cnt, low_off, rows));
      RC_TRY(scan_exclusive_i32_to_i64(ctx, total_row, snn_p, rows));
      // read-back: nnz (sizes the result arrays)
      RC_TRY(fetch_i64(ctx, snn_p + rows, &nnz, st, 0));
      const int64_t nnz_low = (nnz - (int64_t)diag * rows) / 2;
      uint32_t *low_words, *low_sorted;
      int32_t *long_list2;
      RC_TRY(ws_get(ctx, WS_SNN_I, (size_t)nnz + 1, &snn_i));
      RC_TRY(ws_get(ctx, WS_SNN_X, (size_t)nnz + 1, &snn_x));
      RC_TRY(ws_get(ctx, WS_LOW_WORDS, (size_t)nnz_low + 1, &low_words));
      RC_TRY(ws_get(ctx, WS_LOW_SORTED, (size_t)nnz_low + 1, &low_sorted));
      RC_TRY(ws_get(ctx, WS_LONG_LIST2, (size_t)(nnz_low / SORT_WARP_CAP) + 2, &long_list2));
      k_low_scatter<<<wblocks, 256, 0, st>>>(tmp32, tmp_off, row_nnz, cbits, (int32_t)rows, low_off, low_cur,
                                             low_words);
      KERNEL_CHECK(ctx);
      k_sort_columns<uint32_t><<<(unsigned)ceil_div64(rows, SORT_WARPS), SORT_WARPS * 32, 0, st>>>(
          low_words, low_sorted, low_off, rows, nullptr, long_list2, &d_info->n_long2);
      KERNEL_CHECK(ctx);
      {
        size_t smem = sizeof(int32_t) * LONG_SMEM_CAP;
        CU_TRY(cudaFuncSetAttribute(k_sort_long_columns<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
        k_sort_long_columns<uint32_t><<<ctx->sm_count, LONG_THREADS, smem, st>>>(low_words, low_sorted, low_off,
                                                                                 nullptr, long_list2, &d_info->n_long2);
        KERNEL_CHECK(ctx);
      }
      k_snn_fill_sym<<<wblocks, 256, 0, st>>>(tmp32, tmp_off, row_nnz, low_sorted, low_off, snn_p, d_jac, cbits, diag,
                                              k, (int32_t)rows, snn_i, snn_x);
      KERNEL_CHECK(ctx);
    }
    ctx->snn.valid = true;
    ctx->snn.n_rows = rows;
    ctx->snn.nnz = nnz;
    ctx->snn.p = snn_p;
    ctx->snn.i = snn_i;
    ctx->snn.x = snn_x;
    if (stats) {
      stats->nnz_snn = nnz;
      stats->snn_expanded = hi.sum_row_e;
    }
  }
  if (!compute_snn) {  // nothing was read back yet
    RC_TRY(fetch_info(ctx, d_info));
    hi = *ctx->h_info;
    if (hi.err_index_range) {
      if (need_sort) CU_TRY(cudaStreamSynchronize(ctx->side_stream));
      b200nn_set_error("neighbour index out of range: every entry must lie in [1, %lld]", (long long)n);
      return B200NN_ERR_INDEX_RANGE;
    }
  }
  if (want_nn_graph) {
    // join the side stream: the nn Graph is complete (its early copies may still be in flight on the
    // copy stream; the host-buffer entry points wait for those)
    int64_t nnz_nn = 0;
    RC_TRY(fetch_i64(ctx, nn_p + n, &nnz_nn, ctx->side_stream, 1));
    ctx->nn_graph.valid = true;
    ctx->nn_graph.n_rows = n;
    ctx->nn_graph.nnz = nnz_nn;
    ctx->nn_graph.p = nn_p;
    ctx->nn_graph.i = nn_i;
    ctx->nn_graph.x = nn_x;
    if (stats) stats->nnz_nn = nnz_nn;
  }
  if (need_sort && !want_nn_graph) CU_TRY(cudaStreamSynchronize(ctx->side_stream));
  if (stats) stats->max_indegree = hi.max_indegree;
  CU_TRY(cudaEventRecord(e2, st));
  CU_TRY(cudaEventSynchronize(e2));
  if (stats) {
    // ms_nn_graph: the transpose the SNN stage waits for (histogram, scan, scatter); the sorted nn
    // Graph columns are produced on the side stream while the SNN rows are computed
    float ms = 0;
    CU_TRY(cudaEventElapsedTime(&ms, e0, e1));
    stats->ms_nn_graph = ms;
    CU_TRY(cudaEventElapsedTime(&ms, e1, e2));
    stats->ms_snn = ms;
  }
  return B200NN_OK;
}
