// knn_tc.cu -- exact kNN with tensor-core candidate generation (sm_100a).
//
// Replaces RANN::nn2 (call site /root/reference/R/clustering.R:1664-1667) for
// d <= 61 and k <= 32.  Stages, all on the device (G only for large reference sets):
//
//  P  prep      every point is centred (mean of a strided sample), scaled by a
//               power of two s so that |y|^2 <= 2^15, and rounded to FP16:
//               yh = fp16(s (x - mu)).  Queries become rows [yh, 1, 1, 1, 0..]
//               of the A operand, reference points rows
//               [-2 yh, h1, h2, h3, 0..] of the B operand where h1+h2+h3 is
//               the FP16 triple split of |yh|^2.  One 64-wide K block.
//  K1 candidates  persistent warp-specialised kernel.  TMA streams 128-row B
//               tiles (SWIZZLE_128B) through a 5-6-stage mbarrier ring;
//               one thread per resident query tile issues tcgen05.mma (M=128,
//               N=128, K=16 x KSTEPS, FP16 in, FP32 accumulate in TMEM) for the
//               CTA's two resident query tiles; the accumulator IS the selection key
//               key(q,r) = |yh_r|^2 - 2 yh_q.yh_r  (= |yh_q - yh_r|^2 - |yh_q|^2).
//               Eight epilogue warps (one thread per query row) read it back
//               with tcgen05.ld, reject 32-column chunks with a FMNMX3 min
//               tree against the running k'-th best, scan the 8-key groups that
//               hold a hit without inner branches into a pending buffer, and
//               merge that into a per-thread list in shared memory.  The
//               nq x nr key matrix never leaves the SM.
//  G  grouping  (nr >= 32768) coarse k-means partition of the reference set; operands are
//               stored group by group and each work item scans only the groups whose
//               ball can hold a key below its threshold bound (see the section above the
//               grouping kernels).  Results are unchanged; ~3 % of the key tiles are
//               visited on the 32-cluster benchmark mixture.
//  K2 re-rank   one warp per query: the k' candidate rows are gathered with
//               asynchronous copies, FP64 distances in the reference arithmetic
//               (sequential sum, no FMA), ranked by (d2, index); then the
//               exactness certificate
//                   s * d_k  <  sqrt(tau + |yh_q|^2 - eta) - e_q - E_max
//               (tau: smallest rejected key bound, e: FP16 rounding residual
//               norms, eta: bound on the FP32 accumulation error).  By the
//               triangle inequality every non-candidate is then strictly
//               farther than the k-th candidate, so the result equals the
//               exhaustive FP64 search.  Rows that fail go to knn_exact.cu (after a
//               grouped search: to its group-pruned scan, bounded by d_k).
#include <cuda.h>
#include <cuda_fp16.h>
#include <math_constants.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <utility>

#include "internal.cuh"

namespace {

// ---------------------------------------------------------------------------
// geometry
// ---------------------------------------------------------------------------
constexpr int BM = 128;             // queries per query tile (= TMEM lanes)
constexpr int QT = 2;               // query tiles resident per CTA
constexpr int BN = 128;             // reference rows per B tile
constexpr int BK = 64;              // K: one 128-byte swizzle row of FP16
constexpr int CTRL_WARPS = 4;       // TMA, MMA, TMEM alloc, second MMA issuer
constexpr int EPI_WARPS = 8;        // one thread per row of the two resident query tiles
constexpr int NUM_THREADS = (CTRL_WARPS + EPI_WARPS) * 32;
constexpr uint32_t TILE_BYTES = BM * BK * 2;  // 16 KB (A tile == B tile)
constexpr int MAX_D = BK - 3;       // 61
constexpr int MAX_KP = 48;          // shared-memory budget of the candidate lists (one list per query row)
// Larger k' (k > 32): the grouped search runs once per COLUMN SPLIT of the reference set -- split s of S
// holds tiles s, s + S, ... of every group -- each with its own list of k'/S entries per row and its
// own threshold; the row's candidates are the union, its threshold the minimum (every non-candidate
// was rejected by the list of its split, or pruned against that split's bound).  Every tile is still
// visited once in total.
constexpr int MAX_SPLITS = 16;      // k <= 256 (see plan_kp)
// FP32 accumulation error of one key: each of the `ksteps` K = 16 MMAs adds 16 products to the
// running sum; with a truncating aligned adder every partial result is within 2^-23 of the
// largest magnitude M involved, so |error| <= ksteps * 17 * 2^-23 * M (provable, not
// the measured 4 * 2^-24 * M)
constexpr double ETA_PER_KSTEP = 17.0 * 1.1920928955078125e-07;  // 17 * 2^-23
constexpr double NORM_TARGET = 32000.0;

struct TcMisc {            // small device-resident state of one kNN call
  double mu[BK];
  double scale;            // s (a power of two)
  double max_norm2;        // max |x - mu|^2 over reference and query points
  double e_max;            // max FP16 rounding residual norm over the reference set (scaled units)
  double rn2_max;          // max |yh_r|^2
  double qn2_max;
  int32_t n_fail;
  int32_t hang;            // set when an mbarrier wait timed out
  int32_t nonfinite;       // some input coordinate is NA / NaN / Inf (nn2 refuses those: .C(NAOK = FALSE))
  int32_t pad_;
  long long tiles_visited; // (item, tile) pairs of the last candidate launch
  long long tiles_total;
};

// ---------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// (a suspend-time hint on try_wait was measured: it frees issue slots but its wake-up
// latency slows the exhaustive scan by 10%, and the spin loops were not the bottleneck)
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug must end in a trap (CUDA error), never in a hung GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, TcMisc *misc) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  long long t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xffu) == 0) {
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > 8000000000LL) {  // ~4 s: beyond any legitimate wait
        misc->hang = 1;
        __threadfence();
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                           uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
// wait for outstanding tcgen05.ld; the registers are in/out operands so that no use of the
// loaded values can be scheduled above the wait
__device__ __forceinline__ void tc_wait_ld(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]),
                 "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]),
                 "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]),
                 "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]),
                 "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]),
                 "+r"(v[31])
               :
               : "memory");
}
__device__ __forceinline__ float min3f(float a, float b, float c) {
  float r;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (= 1, unused for swizzled K-major)
//   [32,46) SBO >> 4 (8 rows * 128 B = 1024) | [46,48) version = 1 | [61,64) layout = 2 (SW128)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffff) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor: D = F32, A = B = F16, both K-major, M = 128, N = 128
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

// ---------------------------------------------------------------------------
// per-thread candidate state in shared memory (one epilogue thread == one query row)
//   list : kp (key, id) pairs, unsorted, maximum tracked in registers (tau, maxpos)
//   pend : `pend` pairs appended by the filter, merged warp-synchronously
// slot e of thread t lives at word e * ls + t, ls = number of epilogue threads (conflict-free)
// ---------------------------------------------------------------------------
constexpr int PEND = 16;        // pending-hit buffer per thread
constexpr int GROUP_KEYS = 8;  // keys scanned between two capacity checks of the pending buffer

struct Sel {          // register-resident selection state of one query row
  float tau;          // current k'-th best key (+inf until the list is full)
  int cnt_list;       // entries in the list
  int maxpos;         // slot holding tau
  int cnt_pend;       // entries in the pending buffer
};

// recompute the maximum of the (full) list; 8 independent loads in flight
__device__ __forceinline__ void list_rescan(const float *lk, int kp, int ls, float &tau, int &maxpos) {
  float m = -CUDART_INF_F;
  int mp = 0;
  for (int e0 = 0; e0 < kp; e0 += 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = lk[(e0 + u) * ls];
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (v[u] > m) {
        m = v[u];
        mp = e0 + u;
      }
  }
  tau = m;
  maxpos = mp;
}

// Warp-synchronous merge of every lane's pending hits into its list.
// Deliberately not inlined: it is called from 17 sites and runs only ~100 times per
// query tile; the hot filter loop must stay small enough for the instruction cache.
__device__ __noinline__ Sel merge_pending(float *lk, int *li, const float *pk, const int *pi, int kp,
                                          int ls, Sel S) {
  int maxc = S.cnt_pend;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) maxc = max(maxc, __shfl_xor_sync(0xffffffffu, maxc, o));
  for (int e = 0; e < maxc; ++e) {
    bool rescan = false;
    if (e < S.cnt_pend) {
      const float key = pk[e * ls];
      const int id = pi[e * ls];
      if (S.cnt_list < kp) {
        lk[S.cnt_list * ls] = key;
        li[S.cnt_list * ls] = id;
        S.cnt_list++;
        rescan = S.cnt_list == kp;
      } else if (key < S.tau) {
        lk[S.maxpos * ls] = key;
        li[S.maxpos * ls] = id;
        rescan = true;
      }
    }
    if (__any_sync(0xffffffffu, rescan)) {
      // idempotent for lanes that did not change their list
      if (S.cnt_list == kp) list_rescan(lk, kp, ls, S.tau, S.maxpos);
    }
  }
  S.cnt_pend = 0;
  return S;
}

struct TcParams {
  int64_t nq, nr;
  int32_t n_items;     // ceil(nq / (QT * BM))
  int32_t n_btiles;    // ceil(nr / BN)
  int32_t ksteps;      // ceil((d + 3) / 16)
  int32_t kp;          // candidates kept per query
  int32_t stages;      // B ring depth (<= MAX_STAGES)
  int32_t *cand;       // row q's kp candidates at cand + q * cand_stride
  int64_t cand_stride; // >= kp (column splits of the reference set share one [nq x splits*kp] matrix)
  float *tau;          // [nq]
  float *dbg_keys;     // optional [QT*BM x BN]: keys of item 0 / B tile 0
  TcMisc *misc;
  // grouped (tile-pruned) search: item i visits the list_cnt[i] tiles stored at
  // tiles[i * list_stride ...]; null for the exhaustive scan
  const int32_t *list_cnt;
  int64_t list_stride;
  const int32_t *tiles;
  const int32_t *rperm;  // sorted reference position -> original row (-1: padding)
  const int32_t *qperm;  // sorted query position -> original row
  const double *q_n2;    // |yh_q|^2 by original query row (pass 1 only)
  float *item_T;         // pass 1: per-item max over rows of (tau + |yh_q|^2), via atomicMax
  int32_t pass1;         // 1: only item_T is produced
  const int32_t *order;  // items by decreasing list length (null: identity)
  const float *tau_init; // pass 2: per-row threshold found by pass 1 (null: start from +inf)
  int32_t *argmin_out;   // group assignment: only the reference row with the smallest key is wanted
};

// Item visited by CTA `b` in its round `i`: positions are dealt out in a snake so that the
// longest lists (order[] is sorted by decreasing length) spread evenly over the CTAs.
__device__ __forceinline__ int item_at(const TcParams &P, int i, int b, int grid) {
  const int pos = i * grid + ((i & 1) ? grid - 1 - b : b);
  if (pos >= P.n_items) return -1;
  return P.order ? P.order[pos] : pos;
}

// Scan of one 8-key group without inner branches: every key is stored to the next free
// pending slot and the slot is kept only if the key beats tau.  Before a group is scanned
// the warp makes sure that every lane has room for all of its keys (one warp-wide vote),
// so the scan needs neither bounds nor overflow handling.  Padding rows of the reference
// operand carry key = +inf and never pass the test.
#define TC_GROUP(V, G)                                                                     \
  if (__any_sync(0xffffffffu, gm[G] < S.tau)) {                                            \
    if (__any_sync(0xffffffffu, S.cnt_pend > PENDN - GROUP_KEYS))                          \
      S = merge_pending(lk, li, pk, pi, kp, LS, S);                                        \
    if (gm[G] < S.tau) {                                                                   \
      _Pragma("unroll") for (int j = 0; j < GROUP_KEYS; ++j) {                             \
        const float kv = __uint_as_float(V[(G)*GROUP_KEYS + j]);                           \
        pk[S.cnt_pend * LS] = kv;                                                          \
        pi[S.cnt_pend * LS] = col0 + (G)*GROUP_KEYS + j;                                   \
        S.cnt_pend += kv < S.tau ? 1 : 0;                                                  \
      }                                                                                    \
    }                                                                                      \
  }

// 32 keys of one query row: a FMNMX3 tree rejects the chunk for the whole warp in
// ~21 instructions; otherwise the group minima steer the scan to the groups holding hits.
// P1 && bucket (pass 1 of the grouped search, which only has to produce a threshold with at
// least k' keys below it): the minimum of every 8-key group stands for the group, so at most
// four values per chunk reach the pending buffer.
template <int LS, int PENDN, bool P1>
__device__ __forceinline__ void process_chunk(const uint32_t (&v)[32], Sel &S, float *lk, int *li,
                                              float *pk, int *pi, int kp, int col0, bool bucket) {
  static_assert(PENDN >= GROUP_KEYS, "pending buffer smaller than one group");
  float gm[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const float a = min3f(__uint_as_float(v[g * 8 + 0]), __uint_as_float(v[g * 8 + 1]), __uint_as_float(v[g * 8 + 2]));
    const float b = min3f(__uint_as_float(v[g * 8 + 3]), __uint_as_float(v[g * 8 + 4]), __uint_as_float(v[g * 8 + 5]));
    const float c = fminf(__uint_as_float(v[g * 8 + 6]), __uint_as_float(v[g * 8 + 7]));
    gm[g] = min3f(a, b, c);
  }
  const float m = min3f(gm[0], gm[1], fminf(gm[2], gm[3]));
  if (!__any_sync(0xffffffffu, m < S.tau)) return;
  // slow path (warp-uniform entry): some lane holds a key below its threshold
  if (P1 && bucket) {
    if (__any_sync(0xffffffffu, S.cnt_pend > PENDN - 4)) S = merge_pending(lk, li, pk, pi, kp, LS, S);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      pk[S.cnt_pend * LS] = gm[g];
      pi[S.cnt_pend * LS] = col0 + g * GROUP_KEYS;
      S.cnt_pend += gm[g] < S.tau ? 1 : 0;
    }
    return;
  }
  TC_GROUP(v, 0)
  TC_GROUP(v, 1)
  TC_GROUP(v, 2)
  TC_GROUP(v, 3)
}

constexpr int MAX_STAGES = 6;
constexpr int RING = 16;            // tile-id ring (power of two > stages + accumulator stages)

// ---------------------------------------------------------------------------
// K1: candidate generation
// ---------------------------------------------------------------------------
template <bool DBG, bool P1>
__global__ void __launch_bounds__(NUM_THREADS, 1)
k_knn_tc(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_r,
         TcParams P) {
  constexpr int LSTRIDE = EPI_WARPS * 32;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t *sm = smem_raw + (base - raw_u32);
  const int stages = P.stages;
  const uint32_t sA = base;                                     // QT tiles
  const uint32_t sB = sA + QT * TILE_BYTES;                     // `stages` tiles
  const uint32_t off_lists = (QT + stages) * TILE_BYTES;
  float *s_lk = reinterpret_cast<float *>(sm + off_lists);      // [kp][EPI_THREADS]
  int *s_li = reinterpret_cast<int *>(s_lk + P.kp * LSTRIDE);   // [kp][EPI_THREADS]
  float *s_pk = reinterpret_cast<float *>(s_li + P.kp * LSTRIDE);  // [PEND + 1][EPI_THREADS] (one spare slot)
  int *s_pi = reinterpret_cast<int *>(s_pk + (PEND + 1) * LSTRIDE);  // [PEND + 1][EPI_THREADS]
  const uint32_t sBar = base + off_lists + (uint32_t)(2 * P.kp + 2 * (PEND + 1)) * LSTRIDE * 4;
  const uint32_t bar_a_full = sBar + 0, bar_a_empty = sBar + 8;
  const uint32_t bar_b_full = sBar + 16;                        // [MAX_STAGES]
  const uint32_t bar_b_empty = bar_b_full + 8 * MAX_STAGES;     // [MAX_STAGES]
  const uint32_t bar_t_full = bar_b_empty + 8 * MAX_STAGES;     // [2 acc][QT]
  const uint32_t bar_t_empty = bar_t_full + 8 * 2 * QT;         // [2 acc][QT]
  const uint32_t s_tmem = bar_t_empty + 8 * 2 * QT;             // uint32
  // ring of tile ids (producer -> epilogue), indexed by the running B-tile counter
  volatile int32_t *s_ring = reinterpret_cast<volatile int32_t *>(sm + (s_tmem - base) + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(bar_a_full, 1);
    mbar_init(bar_a_empty, QT);          // one commit per MMA issuer
    for (int s = 0; s < stages; ++s) {
      mbar_init(bar_b_full + 8 * s, 1);
      mbar_init(bar_b_empty + 8 * s, QT);  // a stage is free once both issuers' MMAs retired
    }
    for (int i = 0; i < 2 * QT; ++i) {
      mbar_init(bar_t_full + 8 * i, 1);
      mbar_init(bar_t_empty + 8 * i, EPI_WARPS / QT);  // one arrival per epilogue warp of the tile
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_tmem),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(s_tmem));

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0, s = 0, ph = 0, bi = 0;
      for (int rnd = 0; rnd * (int)gridDim.x < P.n_items; ++rnd) {
        const int item = item_at(P, rnd, blockIdx.x, gridDim.x);
        if (item < 0) continue;
        mbar_wait(bar_a_empty, (it & 1) ^ 1, P.misc);
        mbar_expect_tx(bar_a_full, QT * TILE_BYTES);
        for (int t = 0; t < QT; ++t)
          tma_load_2d(sA + t * TILE_BYTES, &map_q, 0, (item * QT + t) * BM, bar_a_full);
        const int64_t l0 = (int64_t)item * P.list_stride;
        const int nt = P.list_cnt ? P.list_cnt[item] : P.n_btiles;
        // tile ids are fetched four at a time, one batch ahead of their use (list_stride % 4 == 0)
        const int4 *tl = reinterpret_cast<const int4 *>(P.tiles ? P.tiles + l0 : nullptr);
        int4 cur = make_int4(0, 0, 0, 0), nxt4 = make_int4(0, 0, 0, 0);
        if (P.tiles && nt > 0) cur = tl[0];
        if (P.tiles && nt > 4) nxt4 = tl[1];
        for (int jj = 0; jj < nt; ++jj, ++bi) {
          int j = jj;
          if (P.tiles) {
            const int u = jj & 3;
            j = u == 0 ? cur.x : u == 1 ? cur.y : u == 2 ? cur.z : cur.w;
            if (u == 3) {
              cur = nxt4;
              if (jj + 5 < nt) nxt4 = tl[(jj + 5) >> 2];
            }
          }
          mbar_wait(bar_b_empty + 8 * s, ph ^ 1, P.misc);
          s_ring[bi & (RING - 1)] = j;  // ordered before the epilogue's read by the barrier chain
          mbar_expect_tx(bar_b_full + 8 * s, TILE_BYTES);
          tma_load_2d(sB + s * TILE_BYTES, &map_r, 0, j * BN, bar_b_full + 8 * s);
          if (++s == (uint32_t)stages) { s = 0; ph ^= 1; }
        }
        ++it;
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ===================== MMA issuers: one thread per resident query tile ==============
    // The two query tiles advance independently (each has its own pair of accumulator
    // stages); a B stage is recycled when both issuers' MMAs on it have retired.
    if (lane == 0) {
      const int t = warp == 1 ? 0 : 1;
      const uint64_t da = make_desc_sw128(sA + t * TILE_BYTES);
      uint32_t it = 0, bi = 0, s = 0, ph = 0;
      for (int rnd = 0; rnd * (int)gridDim.x < P.n_items; ++rnd) {
        const int item = item_at(P, rnd, blockIdx.x, gridDim.x);
        if (item < 0) continue;
        mbar_wait(bar_a_full, it & 1, P.misc);
        tc_fence_after();
        const int nt = P.list_cnt ? P.list_cnt[item] : P.n_btiles;
        for (int j = 0; j < nt; ++j, ++bi) {
          const uint32_t acc = bi & 1, aph = (bi >> 1) & 1;
          mbar_wait(bar_b_full + 8 * s, ph, P.misc);
          mbar_wait(bar_t_empty + 8 * (acc * QT + t), aph ^ 1, P.misc);
          tc_fence_after();
          const uint64_t db = make_desc_sw128(sB + s * TILE_BYTES);
          const uint32_t d_tmem = tmem_base + (uint32_t)((acc * QT + t) * BN);
          for (int kk = 0; kk < P.ksteps; ++kk)
            tc_mma_f16(d_tmem, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), IDESC, kk > 0);
          tc_commit(bar_t_full + 8 * (acc * QT + t));
          tc_commit(bar_b_empty + 8 * s);
          if (++s == (uint32_t)stages) { s = 0; ph ^= 1; }
        }
        tc_commit(bar_a_empty);
        ++it;
      }
    }
  } else if (warp >= CTRL_WARPS) {
    // ===================== epilogue: running top-k' per query row =========================
    const int ew = warp - CTRL_WARPS;        // 0 .. 7
    const int t = ew / 4;                    // query tile of this warp
    const int quarter = warp & 3;            // TMEM lane quarter this warp may access (== ew & 3)
    const int row = quarter * 32 + lane;     // row inside the query tile
    const int et = ew * 32 + lane;           // list column of this thread
    float *lk = s_lk + et;
    int *li = s_li + et;
    float *pk = s_pk + et;
    int *pi = s_pi + et;
    const int kp = P.kp;
    uint32_t bi = 0;
    for (int rnd = 0; rnd * (int)gridDim.x < P.n_items; ++rnd) {
      const int item = item_at(P, rnd, blockIdx.x, gridDim.x);
      if (item < 0) continue;
      Sel S;
      S.tau = CUDART_INF_F;
      S.cnt_list = 0;
      S.maxpos = 0;
      S.cnt_pend = 0;
      const int64_t gq = ((int64_t)item * QT + t) * BM + row;
      const int64_t oq = gq < P.nq ? (P.qperm ? (int64_t)P.qperm[gq] : gq) : -1;  // original query row
      if (P.tau_init && oq >= 0) {
        // pass 2 starts from pass 1's threshold: the pass-1 tiles are visited again, so at least
        // k' keys lie strictly below the next float above it and the list still fills
        const float t1 = P.tau_init[oq];
        if (t1 < CUDART_INF_F) S.tau = nextafterf(t1, CUDART_INF_F);
      }
      const int nt = P.list_cnt ? P.list_cnt[item] : P.n_btiles;
      // >= 4 tiles of one group hold >= 3 tiles without padding = 48 >= MAX_KP eight-key buckets
      const bool bucket = P1 && nt >= 4;
      const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
      uint32_t va[32], vb[32];
      if (nt > 0) {  // prologue: first chunk of the first tile
        const uint32_t acc = bi & 1, aph = (bi >> 1) & 1;
        mbar_wait(bar_t_full + 8 * (acc * QT + t), aph, P.misc);
        tc_fence_after();
        tc_ld32(lane_addr + (uint32_t)((acc * QT + t) * BN), va);
      }
      for (int j = 0; j < nt; ++j, ++bi) {
        const uint32_t acc = bi & 1;
        const uint32_t taddr = lane_addr + (uint32_t)((acc * QT + t) * BN);
        const int colbase = s_ring[bi & (RING - 1)] * BN;
        float *dk = nullptr;
        if (DBG) dk = (item == 0 && j == 0) ? P.dbg_keys + (size_t)(t * BM + row) * BN : nullptr;
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          tc_wait_ld(va);
          tc_ld32(taddr + 64 * h + 32, vb);
          if (DBG && dk) {
#pragma unroll
            for (int c = 0; c < 32; ++c) dk[64 * h + c] = __uint_as_float(va[c]);
          }
          process_chunk<LSTRIDE, PEND, P1>(va, S, lk, li, pk, pi, kp, colbase + 64 * h, bucket);
          tc_wait_ld(vb);
          if (h == 0) {
            tc_ld32(taddr + 64, va);
          } else {
            // all four loads of this tile have landed: hand the accumulator back, and start
            // on the next tile (normally already computed) before scanning the last chunk
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_t_empty + 8 * (acc * QT + t));
            if (j + 1 < nt) {
              const uint32_t nb = bi + 1, nacc = nb & 1, naph = (nb >> 1) & 1;
              mbar_wait(bar_t_full + 8 * (nacc * QT + t), naph, P.misc);
              tc_fence_after();
              tc_ld32(lane_addr + (uint32_t)((nacc * QT + t) * BN), va);
            }
          }
          if (DBG && dk) {
#pragma unroll
            for (int c = 0; c < 32; ++c) dk[64 * h + 32 + c] = __uint_as_float(vb[c]);
          }
          process_chunk<LSTRIDE, PEND, P1>(vb, S, lk, li, pk, pi, kp, colbase + 64 * h + 32, bucket);
        }
      }
      // flush this thread's candidates
      S = merge_pending(lk, li, pk, pi, kp, LSTRIDE, S);
      if (oq >= 0 && P.argmin_out) {
        // nearest-centroid assignment: the smallest key of the list (ties: lower id), deterministic
        float bk = CUDART_INF_F;
        int bi = 0;
        for (int e = 0; e < S.cnt_list; ++e) {
          const float key = lk[e * LSTRIDE];
          const int id = li[e * LSTRIDE];
          if (key < bk || (key == bk && id < bi)) {
            bk = key;
            bi = id;
          }
        }
        P.argmin_out[oq] = bi;
      } else if (oq >= 0) {
        const float tq = S.cnt_list == kp ? S.tau : CUDART_INF_F;
        if (P.pass1) {
          P.tau[oq] = tq;
          // upper bound of tau + |yh_q|^2 (an approximate squared distance, >= 0 up to rounding);
          // the maximum over the column shares bounds the row's final threshold as well
          float T = tq + (float)P.q_n2[oq];
          T = T > 0.f ? T * 1.0001f + 1e-3f : 1e-3f;
          atomicMax(reinterpret_cast<int *>(P.item_T + item), __float_as_int(T));
        } else {
          int32_t *dst = P.cand + oq * P.cand_stride;
          for (int e = 0; e < kp; ++e) {
            int id = e < S.cnt_list ? li[e * LSTRIDE] : -1;
            if (id >= 0 && P.rperm) id = P.rperm[id];
            dst[e] = id;
          }
          P.tau[oq] = tq;
        }
      }
      __syncwarp();
    }
  }
  // teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------
// P: operand preparation
// ---------------------------------------------------------------------------
constexpr int MEAN_BLOCKS = 128;
constexpr int MEAN_ROWS = 256;  // sampled rows per block

__global__ void __launch_bounds__(64)
k_mean_partial(const double *__restrict__ X, int64_t n, int d, double *__restrict__ partial) {
  const int j = threadIdx.x;
  const int64_t m = (int64_t)MEAN_BLOCKS * MEAN_ROWS;
  const int64_t stride = n > m ? n / m : 1;
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  if (j < d) {
    for (int r = 0; r < MEAN_ROWS; r += 4) {
      int64_t i0 = ((int64_t)blockIdx.x * MEAN_ROWS + r) * stride;
      if (i0 < n) a0 += X[i0 * d + j];
      if (i0 + stride < n) a1 += X[(i0 + stride) * d + j];
      if (i0 + 2 * stride < n) a2 += X[(i0 + 2 * stride) * d + j];
      if (i0 + 3 * stride < n) a3 += X[(i0 + 3 * stride) * d + j];
    }
    partial[blockIdx.x * BK + j] = (a0 + a1) + (a2 + a3);
  }
}

__global__ void __launch_bounds__(64)
k_mean_final(const double *__restrict__ partial, int64_t n, int d, TcMisc *__restrict__ misc) {
  const int j = threadIdx.x;
  const int64_t m = (int64_t)MEAN_BLOCKS * MEAN_ROWS;
  const int64_t stride = n > m ? n / m : 1;
  int64_t cnt = (n + stride - 1) / stride;  // number of sampled rows that exist
  if (cnt > m) cnt = m;
  if (j < BK) {
    double s = 0;
    if (j < d)
      for (int b = 0; b < MEAN_BLOCKS; ++b) s += partial[b * BK + j];
    misc->mu[j] = (j < d && cnt > 0) ? s / (double)cnt : 0.0;
  }
  if (j == 0) {
    misc->max_norm2 = 0.0;
    misc->e_max = 0.0;
    misc->rn2_max = 0.0;
    misc->qn2_max = 0.0;
    misc->n_fail = 0;
    misc->hang = 0;
    misc->nonfinite = 0;
    misc->tiles_visited = 0;
    misc->tiles_total = 0;
  }
}

__device__ __forceinline__ void atomic_max_nonneg(double *addr, double v) {
  // non-negative doubles order like their bit patterns
  atomicMax(reinterpret_cast<unsigned long long *>(addr), (unsigned long long)__double_as_longlong(v));
}

__global__ void __launch_bounds__(256)
k_max_norm(const double *__restrict__ X, int64_t n, int d, TcMisc *__restrict__ misc) {
  __shared__ double mu[BK];
  if (threadIdx.x < BK) mu[threadIdx.x] = misc->mu[threadIdx.x];
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double n2 = 0.0;
  if (i < n) {
    const double *x = X + i * d;
    for (int j = 0; j < d; ++j) {
      double y = x[j] - mu[j];
      n2 += y * y;
    }
    if (!isfinite(n2)) misc->nonfinite = 1;  // any NaN / Inf coordinate (or mean) ends up here
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n2 = fmax(n2, __shfl_xor_sync(0xffffffffu, n2, o));
  if ((threadIdx.x & 31) == 0 && n2 > 0.0) atomic_max_nonneg(&misc->max_norm2, n2);
}

__global__ void k_pick_scale(TcMisc *misc) {
  double m = misc->max_norm2;
  double s = 1.0;
  if (m > 0.0 && isfinite(m)) {
    // largest power of two with s^2 * m <= NORM_TARGET
    int e = (int)floor(0.5 * log2(NORM_TARGET / m));
    if (e > 100) e = 100;
    if (e < -100) e = -100;
    s = ldexp(1.0, e);
    while (s * s * m > NORM_TARGET) s *= 0.5;
  }
  misc->scale = s;
}

// one thread per point: FP16 operands rows + rounding residuals.
// a_out: query operand [n x 64] (may be null), b_out: reference operand (may be null)
__global__ void __launch_bounds__(128)
k_convert(const double *__restrict__ X, int64_t n, int d, __half *__restrict__ a_out,
          __half *__restrict__ b_out, double *__restrict__ q_n2, double *__restrict__ q_err,
          TcMisc *__restrict__ misc, uint2 *__restrict__ split_out, int is_ref) {
  __shared__ double mu[BK];
  if (threadIdx.x < BK) mu[threadIdx.x] = misc->mu[threadIdx.x];
  __syncthreads();
  const double s = misc->scale;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double n2 = 0.0, e2 = 0.0;
  if (i < n) {
    const double *x = X + i * d;
    __half *a = a_out ? a_out + i * BK : nullptr;
    __half *b = b_out ? b_out + i * BK : nullptr;
    for (int j0 = 0; j0 < BK; j0 += 8) {
      __align__(16) __half ha[8];
      __align__(16) __half hb[8];
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int j = j0 + jj;
        __half h = __float2half_rn(0.f);
        if (j < d) {
          double y = s * (x[j] - mu[j]);
          h = __double2half(y);
          double hv = (double)__half2float(h);
          double del = hv - y;
          e2 += del * del;
          n2 += hv * hv;
        }
        ha[jj] = h;
        hb[jj] = __hmul(h, __float2half_rn(-2.0f));  // exact
      }
      if (j0 + 8 <= d) {
        if (a) *reinterpret_cast<uint4 *>(a + j0) = *reinterpret_cast<uint4 *>(ha);
        if (b) *reinterpret_cast<uint4 *>(b + j0) = *reinterpret_cast<uint4 *>(hb);
      } else {
        // the chunk holding columns d, d+1, d+2 (and possibly the one after it)
        double rem = n2;  // complete only once j0 + 8 > d, which holds here
        __half h1 = __double2half(rem);
        double r1 = rem - (double)__half2float(h1);
        __half h2 = __double2half(r1);
        double r2 = r1 - (double)__half2float(h2);
        __half h3 = __double2half(r2);
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int j = j0 + jj;
          if (j == d) { ha[jj] = __float2half_rn(1.f); hb[jj] = h1; }
          else if (j == d + 1) { ha[jj] = __float2half_rn(1.f); hb[jj] = h2; }
          else if (j == d + 2) { ha[jj] = __float2half_rn(1.f); hb[jj] = h3; }
          else if (j > d + 2) { ha[jj] = __float2half_rn(0.f); hb[jj] = __float2half_rn(0.f); }
        }
        if (a) *reinterpret_cast<uint4 *>(a + j0) = *reinterpret_cast<uint4 *>(ha);
        if (b) *reinterpret_cast<uint4 *>(b + j0) = *reinterpret_cast<uint4 *>(hb);
      }
    }
    if (q_n2) q_n2[i] = n2;
    if (q_err) q_err[i] = sqrt(e2);
    if (split_out) {
      __half h1 = __double2half(n2);
      double r1 = n2 - (double)__half2float(h1);
      __half h2 = __double2half(r1);
      double r2 = r1 - (double)__half2float(h2);
      __half h3 = __double2half(r2);
      uint2 pk;
      pk.x = (uint32_t)__half_as_ushort(h1) | ((uint32_t)__half_as_ushort(h2) << 16);
      pk.y = (uint32_t)__half_as_ushort(h3);
      split_out[i] = pk;
    }
  }
  // maxima over the block's points
  double em = (i < n) ? sqrt(e2) : 0.0, nm = (i < n) ? n2 : 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    em = fmax(em, __shfl_xor_sync(0xffffffffu, em, o));
    nm = fmax(nm, __shfl_xor_sync(0xffffffffu, nm, o));
  }
  if ((threadIdx.x & 31) == 0) {
    if (b_out || is_ref) {
      if (em > 0.0) atomic_max_nonneg(&misc->e_max, em);
      if (nm > 0.0) atomic_max_nonneg(&misc->rn2_max, nm);
    }
    if (a_out && nm > 0.0) atomic_max_nonneg(&misc->qn2_max, nm);
  }
}

// ---------------------------------------------------------------------------
// G: grouping for the tile-pruned search.  Reference points are partitioned by a
// coarse k-means (any partition is valid; quality only affects how much can be
// skipped), stored group by group (groups padded to whole B tiles with rows whose
// key is +inf), and every work item (256 consecutive sorted queries) visits only
// the groups whose ball can intersect its own:
//     skip g  <=>  (|c_item - c_g| - rho_item - rho_g)^2  >  T_item
// where T_item bounds tau + |yh_q|^2 for every row of the item (from pass 1 over the
// item's nearest group).  By the triangle inequality every point of a skipped group
// has a key above the row's final tau, exactly what the re-rank certificate assumes
// of a non-candidate, so the result stays exact.
// ---------------------------------------------------------------------------
constexpr int KM_SAMPLE = 32768;
constexpr int KM_ITERS = 6;
constexpr int KM_CH = 32;  // centroids staged per shared-memory chunk

__device__ __forceinline__ void load_y_row(const __half *__restrict__ y16, int64_t i, int d, float (&y)[BK]) {
  const uint4 *src = reinterpret_cast<const uint4 *>(y16 + i * BK);
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint4 v = src[c];
    const __half2 *h = reinterpret_cast<const __half2 *>(&v);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float2 f = __half22float2(h[u]);
      y[c * 8 + 2 * u] = (c * 8 + 2 * u) < d ? f.x : 0.f;
      y[c * 8 + 2 * u + 1] = (c * 8 + 2 * u + 1) < d ? f.y : 0.f;
    }
  }
}

__global__ void k_km_init(const __half *__restrict__ y16, int64_t n, int d, int G, float *__restrict__ cent,
                          float *__restrict__ cnorm) {
  const int g = blockIdx.x, j = threadIdx.x;  // 64 threads
  const int64_t i = (int64_t)g * (n / G) + (n / G) / 2;
  float v = j < d ? __half2float(y16[i * BK + j]) : 0.f;
  cent[g * BK + j] = v;
  float s = v * v;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ float part[2];
  if ((j & 31) == 0) part[j >> 5] = s;
  __syncthreads();
  if (j == 0) cnorm[g] = part[0] + part[1];
}

// nearest centroid of rows i = r * stride (r < m); optional k-means accumulation
__global__ void __launch_bounds__(128)
k_km_assign(const __half *__restrict__ y16, int64_t m, int64_t stride, int d, int G,
            const float *__restrict__ cent, const float *__restrict__ cnorm, int32_t *__restrict__ gid,
            int32_t *__restrict__ counts, long long *__restrict__ sums) {
  __shared__ __align__(16) float cs[KM_CH * BK];
  __shared__ float cn[KM_CH];
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool act = r < m;
  float y[BK];
  if (act) load_y_row(y16, r * stride, d, y);
  else {
#pragma unroll
    for (int j = 0; j < BK; ++j) y[j] = 0.f;
  }
  float best = CUDART_INF_F;
  int bg = 0;
  for (int g0 = 0; g0 < G; g0 += KM_CH) {
    __syncthreads();
    for (int e = threadIdx.x; e < KM_CH * BK; e += blockDim.x) {
      int g = g0 + e / BK;
      cs[e] = g < G ? cent[(size_t)g * BK + (e % BK)] : 0.f;
    }
    if (threadIdx.x < KM_CH) cn[threadIdx.x] = (g0 + threadIdx.x) < G ? cnorm[g0 + threadIdx.x] : CUDART_INF_F;
    __syncthreads();
    for (int c = 0; c < KM_CH; ++c) {
      const float4 *cv = reinterpret_cast<const float4 *>(cs + c * BK);
      float dot0 = 0.f, dot1 = 0.f;
#pragma unroll
      for (int q4 = 0; q4 < BK / 4; q4 += 2) {
        float4 a = cv[q4], b = cv[q4 + 1];
        dot0 = fmaf(y[q4 * 4 + 0], a.x, dot0);
        dot0 = fmaf(y[q4 * 4 + 1], a.y, dot0);
        dot0 = fmaf(y[q4 * 4 + 2], a.z, dot0);
        dot0 = fmaf(y[q4 * 4 + 3], a.w, dot0);
        dot1 = fmaf(y[q4 * 4 + 4], b.x, dot1);
        dot1 = fmaf(y[q4 * 4 + 5], b.y, dot1);
        dot1 = fmaf(y[q4 * 4 + 6], b.z, dot1);
        dot1 = fmaf(y[q4 * 4 + 7], b.w, dot1);
      }
      const float sc = cn[c] - 2.f * (dot0 + dot1);
      if (sc < best) {
        best = sc;
        bg = g0 + c;
      }
    }
  }
  if (act) {
    if (gid) gid[r] = bg;
    if (counts) atomicAdd(&counts[bg], 1);
    if (sums) {
#pragma unroll
      for (int j = 0; j < BK; ++j)
        if (j < d)  // exact fixed-point sum: an FP16 value times 2^24 is an integer, and integer addition is
                    // associative, so the centroids do not depend on the order of the atomics (every
                    // device / rank derives the SAME grouping from the same data)
          atomicAdd(reinterpret_cast<unsigned long long *>(&sums[(size_t)bg * BK + j]),
                    (unsigned long long)(long long)(y[j] * 16777216.f));
    }
  }
}

__global__ void k_km_update(long long *__restrict__ sums, int32_t *__restrict__ counts, int d,
                            float *__restrict__ cent, float *__restrict__ cnorm) {
  const int g = blockIdx.x, j = threadIdx.x;  // 64 threads
  const int c = counts[g];
  float v = cent[g * BK + j];
  if (c > 0) v = j < d ? (float)((double)sums[(size_t)g * BK + j] * (1.0 / 16777216.0) / (double)c) : 0.f;
  cent[g * BK + j] = v;
  sums[(size_t)g * BK + j] = 0;
  float s = v * v;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ float part[2];
  if ((j & 31) == 0) part[j >> 5] = s;
  __syncthreads();
  if (j == 0) {
    cnorm[g] = part[0] + part[1];
    counts[g] = 0;
  }
}

// B operand of the centroids for the tensor-core assignment: rows [-2 ch, h1, h2, h3, 0..] with
// ch = fp16(c) and h1 + h2 + h3 the FP16 split of |ch|^2; rows g >= G carry key = +inf.  One warp per row.
__global__ void __launch_bounds__(256)
k_cent_operand(const float *__restrict__ cent, int G, int G_pad, int d, __half *__restrict__ c16) {
  const int lane = threadIdx.x & 31;
  const int g = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (g >= G_pad) return;
  __half h0 = __float2half_rn(0.f), h1 = h0;
  double n2 = 0.0;
  if (g < G) {
    const int j0 = lane, j1 = lane + 32;
    if (j0 < d) h0 = __float2half_rn(cent[(size_t)g * BK + j0]);
    if (j1 < d) h1 = __float2half_rn(cent[(size_t)g * BK + j1]);
    const double a = (double)__half2float(h0), b = (double)__half2float(h1);
    n2 = a * a + b * b;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, o);
  const __half s1 = __double2half(n2);
  const double r1 = n2 - (double)__half2float(s1);
  const __half s2 = __double2half(r1);
  const __half s3 = __double2half(r1 - (double)__half2float(s2));
  for (int j = lane; j < BK; j += 32) {
    __half o = __float2half_rn(0.f);
    if (g >= G) o = j == d ? __ushort_as_half(0x7c00) : o;
    else if (j < d) o = __hmul(j < 32 ? h0 : h1, __float2half_rn(-2.0f));
    else if (j == d) o = s1;
    else if (j == d + 1) o = s2;
    else if (j == d + 2) o = s3;
    c16[(size_t)g * BK + j] = o;
  }
}

// rows per group (shared-memory histogram per block, G <= 1024)
__global__ void __launch_bounds__(256)
k_count_gid(const int32_t *__restrict__ gid, int64_t n, int G, int32_t *__restrict__ counts) {
  __shared__ int s_cnt[1024];
  for (int g = threadIdx.x; g < G; g += blockDim.x) s_cnt[g] = 0;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    atomicAdd(&s_cnt[min(max(gid[i], 0), G - 1)], 1);
  __syncthreads();
  for (int g = threadIdx.x; g < G; g += blockDim.x)
    if (s_cnt[g] > 0) atomicAdd(&counts[g], s_cnt[g]);
}

// Lloyd step: fixed-point sums of the sampled rows per assigned group (exact, order-independent)
__global__ void __launch_bounds__(128)
k_km_accum(const __half *__restrict__ y16, int64_t m, int64_t stride, int d, int G, const int32_t *__restrict__ gid,
           int32_t *__restrict__ counts, long long *__restrict__ sums) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= m) return;
  float y[BK];
  load_y_row(y16, r * stride, d, y);
  const int g = min(max(gid[r], 0), G - 1);
  atomicAdd(&counts[g], 1);
#pragma unroll
  for (int j = 0; j < BK; ++j)
    if (j < d)
      atomicAdd(reinterpret_cast<unsigned long long *>(&sums[(size_t)g * BK + j]),
                (unsigned long long)(long long)(y[j] * 16777216.f));
}

// padded reference offsets (multiples of BN * rtiles) and query offsets (multiples of an item); zeroes the cursors
__global__ void k_group_offsets(const int32_t *__restrict__ rcount, const int32_t *__restrict__ qcount, int G,
                                int32_t *__restrict__ goff, int32_t *__restrict__ qoff,
                                int32_t *__restrict__ rcursor, int32_t *__restrict__ qcursor, int rtiles) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int ro = 0, qo = 0;
  const int rm = BN * rtiles;  // column splits: every group holds the same number of tiles of every split
  for (int g = 0; g < G; ++g) {
    goff[g] = ro;
    qoff[g] = qo;
    ro += ((rcount[g] + rm - 1) / rm) * rm;
    qo += ((qcount[g] + QT * BM - 1) / (QT * BM)) * (QT * BM);  // items never straddle two groups
    rcursor[g] = 0;
    qcursor[g] = 0;
  }
  goff[G] = ro;
  qoff[G] = qo;
}

// scatter of row ids into their group's range; cursors are bumped once per (block, group)
// from a shared-memory histogram instead of once per row (G <= 1024)
__global__ void __launch_bounds__(256)
k_group_scatter(const int32_t *__restrict__ gid, int64_t n, int G, const int32_t *__restrict__ off,
                int32_t *__restrict__ cursor, int32_t *__restrict__ perm, const int32_t *__restrict__ owner = nullptr,
                int rank = 0) {
  __shared__ int s_cnt[1024];
  __shared__ int s_base[1024];
  for (int g = threadIdx.x; g < G; g += blockDim.x) s_cnt[g] = 0;
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int g = -1, local = 0;
  if (i < n) {
    g = gid[i];
    if (owner && owner[g] != rank) g = -1;  // a query row of another share
    if (g >= 0) local = atomicAdd(&s_cnt[g], 1);
  }
  __syncthreads();
  for (int q = threadIdx.x; q < G; q += blockDim.x)
    if (s_cnt[q] > 0) s_base[q] = atomicAdd(&cursor[q], s_cnt[q]);
  __syncthreads();
  if (g >= 0) perm[off[g] + s_base[g] + local] = (int32_t)i;
}

// ---------------------------------------------------------------------------
// Hub-first order inside a group.  With c any fixed point,
//     key(q, r) = |r - c|^2 - 2 (q - c).(r - c) + const(q):
// reference points close to their group's centre are near neighbours of EVERY query that
// visits the group (the "hubs" of high-dimensional data).  Storing each group by increasing
// a_r = |yh_r - c_g|^2 makes the running thresholds of the candidate kernel fall to their
// final values within the first tiles of a group, after which whole 32-key chunks are
// rejected by the min-tree; it also gives every 128-row tile an annulus
// lo_t <= |r - c_g| <= hi_t that prunes tiles inside a kept group (k_build_lists).
// The order is a counting sort over NB z-score buckets of a_r per group (any order is valid).
// ---------------------------------------------------------------------------
constexpr int FINE_NB = 64;

// a_i = |yh_i - c_gid(i)|^2 per row, and per group sum / sum of squares of a (shared-memory
// partial sums per block, G <= 1024)
__global__ void __launch_bounds__(128)
k_row_anorm(const __half *__restrict__ y16, const int32_t *__restrict__ gid, int64_t n, int d, int G,
            const float *__restrict__ cent, float *__restrict__ anorm, float *__restrict__ gstat) {
  __shared__ float s_sum[1024];
  __shared__ float s_sq[1024];
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    s_sum[g] = 0.f;
    s_sq[g] = 0.f;
  }
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float y[BK];
    load_y_row(y16, i, d, y);
    const int g = gid[i];
    const float4 *cv = reinterpret_cast<const float4 *>(cent + (size_t)g * BK);
    float a = 0.f;
#pragma unroll
    for (int q4 = 0; q4 < BK / 4; ++q4) {
      const float4 c = cv[q4];
      const float t0 = y[q4 * 4] - c.x, t1 = y[q4 * 4 + 1] - c.y, t2 = y[q4 * 4 + 2] - c.z, t3 = y[q4 * 4 + 3] - c.w;
      a = fmaf(t0, t0, a);
      a = fmaf(t1, t1, a);
      a = fmaf(t2, t2, a);
      a = fmaf(t3, t3, a);
    }
    anorm[i] = a;
    if (gstat) {
      atomicAdd(&s_sum[g], a);
      atomicAdd(&s_sq[g], a * a);
    }
  }
  if (gstat) {
    __syncthreads();
    for (int g = threadIdx.x; g < G; g += blockDim.x)
      if (s_sum[g] != 0.f) {
        atomicAdd(&gstat[2 * g], s_sum[g]);
        atomicAdd(&gstat[2 * g + 1], s_sq[g]);
      }
  }
}

// (sum, sum of squares) -> (mean, NB / (6 sd)) per group
__global__ void k_group_stat_final(float *__restrict__ gstat, const int32_t *__restrict__ rcount, int G) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  const float c = (float)rcount[g];
  float mean = 0.f, inv = 0.f;
  if (c > 0.f) {
    mean = gstat[2 * g] / c;
    const float var = gstat[2 * g + 1] / c - mean * mean;
    inv = var > 1e-12f * mean * mean && var > 0.f ? (float)FINE_NB / (6.f * sqrtf(var)) : 0.f;
  }
  gstat[2 * g] = mean;
  gstat[2 * g + 1] = inv;
}

__device__ __forceinline__ int fine_bucket(float a, float mean, float inv) {
  const float z = (a - mean) * inv + 0.5f * FINE_NB;
  int b = z > 0.f ? (z < (float)FINE_NB ? (int)z : FINE_NB - 1) : 0;  // NaN -> 0
  return b;
}

__global__ void __launch_bounds__(256)
k_fine_count(const int32_t *__restrict__ gid, const float *__restrict__ anorm, int64_t n,
             const float *__restrict__ gstat, int32_t *__restrict__ fcount) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int g = gid[i];
  atomicAdd(&fcount[g * FINE_NB + fine_bucket(anorm[i], gstat[2 * g], gstat[2 * g + 1])], 1);
}

// one block per group: bucket offsets inside the group's (padded) range; cursors zeroed
__global__ void __launch_bounds__(FINE_NB)
k_fine_offsets(const int32_t *__restrict__ fcount, const int32_t *__restrict__ off, int32_t *__restrict__ foff,
               int32_t *__restrict__ fcursor) {
  __shared__ int s[FINE_NB];
  const int g = blockIdx.x, b = threadIdx.x;
  s[b] = fcount[g * FINE_NB + b];
  __syncthreads();
  int pre = 0;
  for (int u = 0; u < b; ++u) pre += s[u];
  foff[g * FINE_NB + b] = off[g] + pre;
  fcursor[g * FINE_NB + b] = 0;
}

__global__ void __launch_bounds__(256)
k_fine_scatter(const int32_t *__restrict__ gid, const float *__restrict__ anorm, int64_t n,
               const float *__restrict__ gstat, const int32_t *__restrict__ foff, int32_t *__restrict__ fcursor,
               int32_t *__restrict__ perm, const int32_t *__restrict__ owner, int rank,
               const int32_t *__restrict__ goff = nullptr, int deal = 1) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int g = gid[i];
  if (owner && owner[g] != rank) return;  // a query row of another share
  const int f = g * FINE_NB + fine_bucket(anorm[i], gstat[2 * g], gstat[2 * g + 1]);
  int p = foff[f] + atomicAdd(&fcursor[f], 1);
  if (deal > 1) {
    // column splits: the hub-first sequence of the group is dealt round-robin into `deal` sub-sequences
    // stored one after the other (each a whole number of tiles, itself hub-first), so that every split
    // is a uniform sample of the group -- a row's near neighbours, which crowd into the first tiles
    // of the hub-first order, then fall into the splits binomially (see plan_kp)
    const int g0 = goff[g], sub = (goff[g + 1] - g0) / deal, pl = p - g0;
    p = g0 + (pl % deal) * sub + pl / deal;
  }
  perm[p] = (int32_t)i;
}

// annulus [lo, hi] of |yh_r - c_g| over the rows of every 128-row tile (one warp per tile);
// rounded outwards.  Padding-only tiles get lo = +inf.
__global__ void __launch_bounds__(256)
k_tile_annulus(const int32_t *__restrict__ rperm, const float *__restrict__ anorm, int64_t n_tiles,
               float *__restrict__ tile_lo, float *__restrict__ tile_hi) {
  const int lane = threadIdx.x & 31;
  const int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (t >= n_tiles) return;
  float mn = CUDART_INF_F, mx = 0.f;
  for (int u = lane; u < BN; u += 32) {
    const int i = rperm[t * BN + u];
    if (i >= 0) {
      const float a = anorm[i];
      mn = fminf(mn, a);
      mx = fmaxf(mx, a);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if (lane == 0) {
    tile_lo[t] = mn < CUDART_INF_F ? fmaxf(sqrtf(mn) * 0.9999f - 1e-4f, 0.f) : CUDART_INF_F;
    tile_hi[t] = sqrtf(mx) * 1.0001f + 1e-4f;
  }
}

// Sharded self search (opts.shard_world > 1): whole groups are dealt to the shares, largest first,
// each to the share with the fewest query rows so far (LPT).  rcount, hence the deal, is identical on
// every device that runs the same data (the grouping is deterministic), so the shares partition the
// rows.  qcount[g] = rcount[g] for the groups of `rank`, 0 for the others.
__global__ void __launch_bounds__(1024)
k_group_owner(const int32_t *__restrict__ rcount, int G, int world, int rank, int32_t *__restrict__ owner,
              int32_t *__restrict__ qcount) {
  __shared__ int s_order[1024];
  __shared__ int s_owner[1024];
  const int g = threadIdx.x;
  if (g < G) {
    const int c = rcount[g];
    int pos = 0;
    for (int h = 0; h < G; ++h) {
      const int ch = rcount[h];
      pos += (ch > c || (ch == c && h < g)) ? 1 : 0;
    }
    s_order[pos] = g;
  }
  __syncthreads();
  if (g == 0) {
    long long load[64];
    for (int r = 0; r < world; ++r) load[r] = 0;
    for (int p = 0; p < G; ++p) {
      const int gg = s_order[p];
      int best = 0;
      for (int r = 1; r < world; ++r)
        if (load[r] < load[best]) best = r;
      s_owner[gg] = best;
      load[best] += rcount[gg];
    }
  }
  __syncthreads();
  if (g < G) {
    owner[g] = s_owner[g];
    qcount[g] = s_owner[g] == rank ? rcount[g] : 0;
  }
}

// padding rows [n, n_pad) of an ungrouped reference operand: key = +inf for every query
__global__ void __launch_bounds__(128)
k_pad_ref(__half *__restrict__ r16, int64_t n, int64_t n_pad, int d) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t p = n + e / BK;
  const int j = (int)(e % BK);
  if (p < n_pad) r16[p * BK + j] = j == d ? __ushort_as_half(0x7c00) : __float2half_rn(0.f);
}

// B operand rows in sorted, padded order; padding rows get key = +inf
__global__ void __launch_bounds__(256)
k_gather_ref(const __half *__restrict__ y16, const uint2 *__restrict__ split, const int32_t *__restrict__ rperm,
             int64_t nr_pad, int d, __half *__restrict__ r16) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t p = e >> 3;
  const int c = (int)(e & 7);
  if (p >= nr_pad) return;
  const int i = rperm[p];
  __align__(16) __half out[8];
  if (i < 0) {
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) out[jj] = (c * 8 + jj) == d ? __ushort_as_half(0x7c00) : __float2half_rn(0.f);
  } else {
    uint4 v = *reinterpret_cast<const uint4 *>(y16 + (int64_t)i * BK + c * 8);
    const __half *h = reinterpret_cast<const __half *>(&v);
    const uint2 sp = split[i];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const int j = c * 8 + jj;
      __half o;
      if (j < d) o = __hmul(h[jj], __float2half_rn(-2.0f));
      else if (j == d) o = __ushort_as_half((unsigned short)(sp.x & 0xffff));
      else if (j == d + 1) o = __ushort_as_half((unsigned short)(sp.x >> 16));
      else if (j == d + 2) o = __ushort_as_half((unsigned short)(sp.y & 0xffff));
      else o = __float2half_rn(0.f);
      out[jj] = o;
    }
  }
  *reinterpret_cast<uint4 *>(r16 + p * BK + c * 8) = *reinterpret_cast<uint4 *>(out);
}

// A operand rows in sorted order (plain row gather)
__global__ void __launch_bounds__(256)
k_gather_qry(const __half *__restrict__ y16, const int32_t *__restrict__ qperm, int64_t nq,
             __half *__restrict__ q16) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t p = e >> 3;
  const int c = (int)(e & 7);
  if (p >= nq) return;
  const int64_t i = qperm[p];
  uint4 v = make_uint4(0, 0, 0, 0);  // padding rows are all zero and never flushed
  if (i >= 0) v = *reinterpret_cast<const uint4 *>(y16 + i * BK + c * 8);
  *reinterpret_cast<uint4 *>(q16 + p * BK + c * 8) = v;
}

// radius of every group around its centroid (rounded up)
__global__ void __launch_bounds__(128)
k_group_radius(const __half *__restrict__ y16, const int32_t *__restrict__ rperm, const int32_t *__restrict__ goff,
               int d, const float *__restrict__ cent, float *__restrict__ grad) {
  // grid (G, RAD_SLICES): every block covers a strided share of the group's rows and folds
  // its maximum into grad[g] (pre-set to the rounding slack) with an integer atomicMax
  __shared__ float cs[BK];
  __shared__ float red[4];
  const int g = blockIdx.x;
  if (threadIdx.x < BK) cs[threadIdx.x] = cent[(size_t)g * BK + threadIdx.x];
  __syncthreads();
  float mx = 0.f;
  for (int p = goff[g] + blockIdx.y * blockDim.x + threadIdx.x; p < goff[g + 1]; p += blockDim.x * gridDim.y) {
    const int i = rperm[p];
    if (i < 0) continue;
    float y[BK];
    load_y_row(y16, i, d, y);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < BK; ++j) {
      float t = y[j] - cs[j];
      s = fmaf(t, t, s);
    }
    mx = fmaxf(mx, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    atomicMax(reinterpret_cast<int *>(grad + g), __float_as_int(sqrtf(m) * 1.0001f + 1e-4f));  // positive floats
  }
}

// centroid and radius of every work item (QT*BM consecutive sorted queries)
__global__ void __launch_bounds__(64)
k_item_stats(const __half *__restrict__ q16, const int32_t *__restrict__ qperm, int64_t nq, int d,
             float *__restrict__ item_c, float *__restrict__ item_rad) {
  __shared__ float cs[BK];
  __shared__ float red[2];
  __shared__ int s_rows;
  const int item = blockIdx.x, j = threadIdx.x;
  const int64_t r0 = (int64_t)item * QT * BM;
  if (j == 0) s_rows = 0;
  __syncthreads();
  {  // valid rows form a prefix of the item (the group's queries, then padding)
    int c = 0;
    for (int r = j; r < QT * BM; r += 64)
      if (r0 + r < nq && qperm[r0 + r] >= 0) c++;
    atomicAdd(&s_rows, c);
  }
  __syncthreads();
  const int rows = s_rows;
  float s = 0.f;
  if (j < d)
    for (int r = 0; r < rows; ++r) s += __half2float(q16[(r0 + r) * BK + j]);
  cs[j] = (j < d && rows > 0) ? s / (float)rows : 0.f;
  item_c[(size_t)item * BK + j] = cs[j];
  __syncthreads();
  float mx = 0.f;
  for (int r = j; r < rows; r += 64) {
    float y[BK];
    load_y_row(q16, r0 + r, d, y);
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < BK; ++c) {
      float t = y[c] - cs[c];
      q = fmaf(t, t, q);
    }
    mx = fmaxf(mx, q);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((j & 31) == 0) red[j >> 5] = mx;
  __syncthreads();
  if (j == 0) item_rad[item] = rows > 0 ? sqrtf(fmaxf(red[0], red[1])) * 1.0001f + 1e-4f : -1.f;
}

// tile list of every item.  pass 1: the tiles of the nearest group (to establish T_item);
// pass 2: every group whose ball may hold a key below the item's bound, plus the pass-1 group.
constexpr int LIST_THREADS = 256;
constexpr int PASS1_TILES = 8;
constexpr int LIST_NBK = 64;  // score buckets of the pass-2 visiting order
__global__ void __launch_bounds__(LIST_THREADS)
k_build_lists(const float *__restrict__ item_c, const float *__restrict__ item_rad,
              const float *__restrict__ item_T, int G, const float *__restrict__ cent,
              const float *__restrict__ grad, const int32_t *__restrict__ goff,
              const float *__restrict__ tile_lo, const float *__restrict__ tile_hi, int pass1,
              int64_t list_stride, int32_t *__restrict__ tiles, int32_t *__restrict__ list_cnt,
              int nsplit, int split, const int32_t *__restrict__ rcount, int min_rows) {
  __shared__ float cs[BK];
  __shared__ float s_min[LIST_THREADS / 32];
  __shared__ int s_scan[LIST_THREADS / 32];
  __shared__ float s_big[LIST_THREADS / 32];
  __shared__ float s_dmin, s_dbig;
  __shared__ int s_hist[LIST_NBK];
  __shared__ int s_cur[LIST_NBK];
  __shared__ int s_total;
  const int item = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (item_rad[item] < 0.f) {  // padding-only item
    if (tid == 0) list_cnt[item] = 0;
    return;
  }
  if (tid < BK) cs[tid] = item_c[(size_t)item * BK + tid];
  __syncthreads();
  constexpr int GPT = 4;  // groups per thread (G <= 1024)
  float dist[GPT];
  bool big[GPT];
  // The ANCHOR group of an item -- pass 1 scans its first tiles to get a threshold, pass 2 always visits
  // them again -- is the nearest group that can fill a k'-entry list: at least ANCHOR_TILES tiles of this
  // split (pass 1 then works on 8-key bucket minima of >= 3 full tiles) or at least min_rows = 2 k' rows per
  // split (plain keys).  A tiny group (k-means leaves a few with a handful of points) cannot: the item's
  // bound would stay +inf and the item would visit EVERY tile -- one such item of 7 942 tiles among items of
  // ~250 made its CTA the tail of the whole launch (pass 2 of config C: 16.0 -> 12.3 ms once it was gone).
  // Without any sufficient group the nearest group of all is the anchor.
  constexpr int ANCHOR_TILES = 4;
  float dmin = CUDART_INF_F, dbig = CUDART_INF_F;
#pragma unroll
  for (int u = 0; u < GPT; ++u) {
    const int g = tid * GPT + u;
    dist[u] = CUDART_INF_F;
    big[u] = false;
    if (g < G && goff[g + 1] > goff[g]) {
      const float4 *cv = reinterpret_cast<const float4 *>(cent + (size_t)g * BK);
      float s = 0.f;
#pragma unroll
      for (int q4 = 0; q4 < BK / 4; ++q4) {
        float4 a = cv[q4];
        float t0 = a.x - cs[q4 * 4], t1 = a.y - cs[q4 * 4 + 1], t2 = a.z - cs[q4 * 4 + 2], t3 = a.w - cs[q4 * 4 + 3];
        s = fmaf(t0, t0, s);
        s = fmaf(t1, t1, s);
        s = fmaf(t2, t2, s);
        s = fmaf(t3, t3, s);
      }
      dist[u] = sqrtf(s);
      dmin = fminf(dmin, dist[u]);
      big[u] = (goff[g + 1] - goff[g]) / (BN * nsplit) >= ANCHOR_TILES || rcount[g] / nsplit >= min_rows;
      if (big[u]) dbig = fminf(dbig, dist[u]);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    dmin = fminf(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
    dbig = fminf(dbig, __shfl_xor_sync(0xffffffffu, dbig, o));
  }
  if (lane == 0) {
    s_min[w] = dmin;
    s_big[w] = dbig;
  }
  __syncthreads();
  if (tid == 0) {
    float m = s_min[0], b = s_big[0];
    for (int i = 1; i < LIST_THREADS / 32; ++i) {
      m = fminf(m, s_min[i]);
      b = fminf(b, s_big[i]);
    }
    s_dmin = m;
    s_dbig = b;
  }
  __syncthreads();
  const bool any_big = s_dbig < CUDART_INF_F;
  dmin = any_big ? s_dbig : s_dmin;  // distance of the anchor group
  const float T = pass1 ? 0.f : item_T[item];
  const float ri = item_rad[item];
  int ntile[GPT], mine = 0;
#pragma unroll
  for (int u = 0; u < GPT; ++u) {
    const int g = tid * GPT + u;
    ntile[u] = 0;
    if (dist[u] < CUDART_INF_F) {
      bool keep = dist[u] <= dmin && (big[u] || !any_big);  // the anchor group is always visited (pass 1 == subset of pass 2)
      if (!pass1 && !keep) {
        float lb = dist[u] * 0.99999f - ri - grad[g];
        keep = !(lb > 0.f && lb * lb > T);  // also keeps everything when T is +inf / NaN
      }
      // column split `split` of `nsplit` (k' beyond one shared-memory list): the split-th of the group's
      // nsplit equally long runs of tiles (k_fine_scatter deals the rows; groups hold multiples of nsplit tiles)
      if (keep) ntile[u] = (goff[g + 1] - goff[g]) / (BN * nsplit);
      // pass 1 only needs an upper bound of tau: a few tiles of the own group suffice
      if (pass1 && ntile[u] > PASS1_TILES) ntile[u] = PASS1_TILES;
    }
    mine += ntile[u];
  }
  // block exclusive scan of `mine`
  int inc = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_scan[w] = inc;
  __syncthreads();
  int woff = 0, total = 0;
  for (int i = 0; i < LIST_THREADS / 32; ++i) {
    if (i < w) woff += s_scan[i];
    total += s_scan[i];
  }
  int32_t *dst = tiles + (int64_t)item * list_stride;
  if (pass1) {
    int pos = woff + inc - mine;
#pragma unroll
    for (int u = 0; u < GPT; ++u) {
      const int g = tid * GPT + u;
      if (ntile[u] > 0) {
        const int t0 = goff[g] / BN + split * ((goff[g + 1] - goff[g]) / (BN * nsplit));
        for (int t = 0; t < ntile[u]; ++t) dst[pos + t] = t0 + t;
      }
      pos += ntile[u];
    }
  } else {
    // pass 2.  Inside a kept group every tile is tested against its annulus: for r in tile t,
    // |q - r| >= | |q - c_g| - |r - c_g| | with |q - c_g| in [D - rho_item, D + rho_item].
    // Kept tiles are visited by increasing expected squared distance D^2 + lo_t^2 (64 score
    // buckets, counting sort), so the thresholds reach their final values within the first
    // tiles and later tiles yield almost no hits.  The set of visited tiles -- hence the
    // result -- does not depend on the order.  The pass-1 tiles are always visited again.
    for (int b = tid; b < LIST_NBK; b += LIST_THREADS) s_hist[b] = 0;
    __syncthreads();
    const float s0 = dmin * dmin;
    const float inv = (T > 0.f && T < CUDART_INF_F) ? (float)LIST_NBK / (4.f * T) : 0.f;
    for (int phase = 0; phase < 2; ++phase) {
#pragma unroll
      for (int u = 0; u < GPT; ++u) {
        const int g = tid * GPT + u;
        if (ntile[u] <= 0) continue;
        const int t0 = goff[g] / BN + split * ((goff[g + 1] - goff[g]) / (BN * nsplit));
        const float D = dist[u];
        const bool nearest = D <= dmin && (big[u] || !any_big);
        for (int t = 0; t < ntile[u]; ++t) {
          const int tile = t0 + t;
          const float lo = tile_lo[tile], hi = tile_hi[tile];
          bool keep_t = true;
          if (!(nearest && t < PASS1_TILES)) {
            const float lb = fmaxf(lo - (D * 1.00001f + ri), (D * 0.99999f - ri) - hi);
            keep_t = !(lb > 0.f && lb * lb > T);
          }
          if (!keep_t) continue;
          const float z = (D * D + lo * lo - s0) * inv;
          const int b = z > 0.f ? (z < (float)LIST_NBK ? (int)z : LIST_NBK - 1) : 0;
          if (phase == 0)
            atomicAdd(&s_hist[b], 1);
          else
            dst[atomicAdd(&s_cur[b], 1)] = tile;
        }
      }
      __syncthreads();
      if (phase == 0) {
        if (tid == 0) {
          int run = 0;
          for (int b = 0; b < LIST_NBK; ++b) {
            s_cur[b] = run;
            run += s_hist[b];
          }
          s_total = run;
        }
        __syncthreads();
      }
    }
    total = s_total;
  }
  if (tid == 0) list_cnt[item] = total;
}

// items ordered by decreasing list length (counting sort; one block of 1024 threads)
__global__ void __launch_bounds__(1024)
k_order_items(const int32_t *__restrict__ list_cnt, int n_items, int max_cnt, int32_t *__restrict__ bins,
              int32_t *__restrict__ order, TcMisc *__restrict__ misc, int accumulate) {
  __shared__ int s_part[1024];
  __shared__ long long s_vis[32];
  const int tid = threadIdx.x;
  for (int b = tid; b <= max_cnt; b += blockDim.x) bins[b] = 0;
  __syncthreads();
  for (int i = tid; i < n_items; i += blockDim.x) atomicAdd(&bins[list_cnt[i]], 1);
  __syncthreads();
  // descending exclusive scan over bins: thread t owns a contiguous chunk, highest bins first
  const int per = (max_cnt + 1 + 1023) / 1024;
  const int hi = max_cnt - tid * per;            // first (largest) bin of this thread
  int local = 0;
  long long vis = 0;
  for (int u = 0; u < per; ++u) {
    const int b = hi - u;
    if (b >= 0) {
      local += bins[b];
      vis += (long long)b * bins[b];
    }
  }
  s_part[tid] = local;
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int t = 0; t < 1024; ++t) {
      int c = s_part[t];
      s_part[t] = run;
      run += c;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) vis += __shfl_xor_sync(0xffffffffu, vis, o);
  if ((tid & 31) == 0) s_vis[tid >> 5] = vis;
  __syncthreads();
  int run = s_part[tid];
  for (int u = 0; u < per; ++u) {
    const int b = hi - u;
    if (b >= 0) {
      const int c = bins[b];
      bins[b] = run;
      run += c;
    }
  }
  if (tid == 0) {
    long long v = 0;
    for (int w = 0; w < 32; ++w) v += s_vis[w];
    misc->tiles_visited = accumulate ? misc->tiles_visited + v : v;
    misc->tiles_total = (long long)n_items * max_cnt;
  }
  __syncthreads();
  for (int i = tid; i < n_items; i += blockDim.x) order[atomicAdd(&bins[list_cnt[i]], 1)] = i;
}

// ---------------------------------------------------------------------------
// K2: FP64 re-rank + certificate.  One warp per query.
// ---------------------------------------------------------------------------
constexpr int RR_WARPS = 4;

// shared-memory row stride (doubles) of the gathered candidate rows.  Even d: rows are moved
// and read 16 bytes at a time, and a stride of an odd number of 16-byte units keeps the
// per-lane row reads conflict-free; odd d: 8-byte accesses with an odd stride.
__host__ __device__ inline int rr_stride(int d) { return (d & 1) ? d + 2 : ((d & 3) == 2 ? d : d + 2); }
__host__ __device__ inline int rr_kp_round(int kp) { return kp <= 64 ? 64 : (kp + 31) & ~31; }
__host__ __device__ inline size_t rr_per_warp(int d, int kp) {
  size_t b = sizeof(double) * ((size_t)32 * rr_stride(d) + ((d + 1) & ~1)) + (size_t)rr_kp_round(kp) * (8 + 4);
  return (b + 15) & ~(size_t)15;
}

__device__ __forceinline__ void cp_async8(void *s, const void *g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(s)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async16(void *s, const void *g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(s)), "l"(g) : "memory");
}

template <bool VEC>  // VEC: d even and both matrices 16-byte aligned
__global__ void __launch_bounds__(RR_WARPS * 32)
k_rerank(const double *__restrict__ ref, const double *__restrict__ qry, int64_t nq, int d, int k,
         int kp, const int32_t *__restrict__ cand, const float *__restrict__ tau,
         const double *__restrict__ q_n2, const double *__restrict__ q_err,
         TcMisc *__restrict__ misc, int32_t *__restrict__ idx0, double *__restrict__ dist,
         int32_t *__restrict__ fail_list, double *__restrict__ fail_tau,
         const int32_t *__restrict__ visit, int64_t n_visit, double eta_coef) {
  extern __shared__ __align__(16) unsigned char rr_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ds = rr_stride(d);
  // per warp: rows[32][ds] doubles, q[d] doubles, key[kp rounded] u64, id[kp rounded] i32
  unsigned char *wbase = rr_smem + (size_t)warp * rr_per_warp(d, kp);
  double *rows = reinterpret_cast<double *>(wbase);
  double *qv = rows + 32 * ds;
  unsigned long long *skey = reinterpret_cast<unsigned long long *>(qv + ((d + 1) & ~1));
  int32_t *sid = reinterpret_cast<int32_t *>(skey + rr_kp_round(kp));

  // queries are visited in the group-sorted order of the candidate stage when available:
  // neighbouring warps then gather overlapping candidate rows (L2 hits instead of HBM)
  const int64_t pos = (int64_t)blockIdx.x * RR_WARPS + warp;
  if (pos >= n_visit) return;
  const int64_t q = visit ? (int64_t)visit[pos] : pos;
  if (q < 0 || q >= nq) return;
  if (VEC) {
    if (lane < (d >> 1)) cp_async16(qv + 2 * lane, qry + q * d + 2 * lane);
    for (int j = 64 + 2 * lane; j < d; j += 64) cp_async16(qv + j, qry + q * d + j);
  } else {
    for (int j = lane; j < d; j += 32) cp_async8(qv + j, qry + q * d + j);
  }
  for (int half = 0; half * 32 < kp; ++half) {
    const int c = half * 32 + lane;
    const int id = c < kp ? cand[q * kp + c] : -1;
    // cooperative, coalesced gather of the 32 candidate rows by asynchronous copies: every row
    // is in flight at once (one L2 round trip per query instead of one per candidate)
#pragma unroll 4
    for (int cc = 0; cc < 32; ++cc) {
      const int idc = __shfl_sync(0xffffffffu, id, cc);
      if (idc >= 0) {
        const double *src = ref + (int64_t)idc * d;
        double *dst = rows + cc * ds;
        if (VEC) {
          if (lane < (d >> 1)) cp_async16(dst + 2 * lane, src + 2 * lane);
          for (int j = 64 + 2 * lane; j < d; j += 64) cp_async16(dst + j, src + j);
        } else {
          for (int j = lane; j < d; j += 32) cp_async8(dst + j, src + j);
        }
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    // the reference's arithmetic: differences squared and summed in column order, no FMA
    double d2 = CUDART_INF;
    if (id >= 0) {
      d2 = 0.0;
      const double *rr = rows + lane * ds;
      if (VEC) {
        for (int j = 0; j < d; j += 2) {
          const double2 a = *reinterpret_cast<const double2 *>(qv + j);
          const double2 b = *reinterpret_cast<const double2 *>(rr + j);
          const double t0 = __dsub_rn(a.x, b.x), t1 = __dsub_rn(a.y, b.y);
          d2 = __dadd_rn(d2, __dmul_rn(t0, t0));
          d2 = __dadd_rn(d2, __dmul_rn(t1, t1));
        }
      } else {
        for (int j = 0; j < d; ++j) {
          const double tt = __dsub_rn(qv[j], rr[j]);
          d2 = __dadd_rn(d2, __dmul_rn(tt, tt));
        }
      }
    }
    skey[c] = (unsigned long long)__double_as_longlong(d2);  // d2 >= 0: bit order == value order
    sid[c] = id >= 0 ? id : 0x7fffffff;
    __syncwarp();
  }
  // rank of every candidate under (d2, index); ranks below k are the result rows.  Missing
  // candidates share the key (+inf, INT_MAX): they rank after every real one.
  const int kp_r = ((kp + 31) / 32) * 32;
  unsigned long long kth = (unsigned long long)__double_as_longlong(CUDART_INF);
  for (int c = lane; c < kp_r; c += 32) {
    const unsigned long long mk = skey[c];
    const int mi = sid[c];
    int rank = 0;
#pragma unroll 8
    for (int o = 0; o < kp_r; ++o) {
      const unsigned long long ok = skey[o];
      const int oi = sid[o];
      rank += (ok < mk || (ok == mk && oi < mi)) ? 1 : 0;
    }
    if (rank < k && mi != 0x7fffffff) {
      idx0[q * k + rank] = mi;
      dist[q * k + rank] = sqrt(__longlong_as_double((long long)mk));
    }
    const unsigned hit = __ballot_sync(0xffffffffu, rank == k - 1);
    if (hit) kth = __shfl_sync(0xffffffffu, mk, __ffs(hit) - 1);
  }
  // certificate
  const double s = misc->scale;
  const double qn = q_n2[q], eq = q_err[q];
  const double M = misc->rn2_max + 2.0 * sqrt(qn) * sqrt(misc->rn2_max) + 1.0;
  const double eta = eta_coef * M + 1e-4;
  // every non-candidate has a key >= the threshold of the list that rejected it
  const double tq = (double)tau[q];
  const double dk2 = __longlong_as_double((long long)kth);
  bool ok = false;
  if (isfinite(tq) && isfinite(dk2)) {
    double lb2 = tq + qn - eta;
    double lb = (lb2 > 0.0 ? sqrt(lb2) : 0.0) - eq - misc->e_max;
    ok = s * sqrt(dk2) * (1.0 + 1e-9) < lb * (1.0 - 1e-9);
  }
  if (!ok && lane == 0) {
    int pos = atomicAdd(&misc->n_fail, 1);
    fail_list[pos] = (int32_t)q;
    // the k-th candidate is a real point: its squared distance (reference arithmetic) bounds
    // the true k-th squared distance from above and lets the exhaustive scan skip nearly all rows
    fail_tau[pos] = isfinite(dk2) ? dk2 : CUDART_INF;
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                    const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                    const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_operand_map(CUtensorMap *map, const void *ptr, int64_t rows, int64_t row_stride = 1) {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    CU_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr));
    if (!p || qr != cudaDriverEntryPointSuccess) {
      b200nn_set_error("cuTensorMapEncodeTiled is not available from the driver");
      return B200NN_ERR_CUDA;
    }
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  cuuint64_t gdim[2] = {(cuuint64_t)BK, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)BK * 2 * (cuuint64_t)row_stride};  // row_stride > 1: every stride-th row
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(ptr), gdim, gstr, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    b200nn_set_error("cuTensorMapEncodeTiled failed with code %d", (int)r);
    return B200NN_ERR_CUDA;
  }
  return B200NN_OK;
}

int pick_kp(int k, int n_cand) {
  int kp = n_cand > 0 ? n_cand : k + (k / 2 > 8 ? k / 2 : 8);
  if (kp < k + 1) kp = k + 1;
  kp = (kp + 7) & ~7;
  return kp;
}

// k' candidates per row as `splits` lists of `per` entries.  A row's certificate needs about `want`
// (~1.5 k) keys below its threshold, and the threshold is the SMALLEST of the splits' thresholds: the
// want keys fall into the splits binomially (mean m = want / S, deviation ~ sqrt(m)), so every list
// gets four deviations of head room -- measured with per = want / S: 64 % of the rows of a 1M-cell
// input lost their certificate at k = 100 to the one unlucky split.  splits == 0: no plan.
struct KpPlan {
  int splits, per, total;
};
KpPlan plan_kp(int want) {
  KpPlan p = {1, want, want};
  if (want <= MAX_KP) return p;
  for (int S = 2; S <= MAX_SPLITS; ++S) {
    const double m = (double)want / S;
    const int per = ((int)ceil(m + 4.0 * sqrt(m)) + 7) & ~7;
    if (per <= MAX_KP) {
      p.splits = S;
      p.per = per;
      p.total = S * per;
      return p;
    }
  }
  p.splits = 0;
  return p;
}

// tau[q] = min over the splits' thresholds tau[s * nq + q]
__global__ void k_tau_min(float *__restrict__ tau, int64_t nq, int splits) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  float m = tau[q];
  for (int s = 1; s < splits; ++s) m = fminf(m, tau[(int64_t)s * nq + q]);
  tau[q] = m;
}

size_t tc_smem_bytes(int kp, int stages) {
  return 1024 + (size_t)(QT + stages) * TILE_BYTES +
         (size_t)(2 * kp + 2 * (PEND + 1)) * EPI_WARPS * 32 * 4 +
         8 * (2 + 2 * MAX_STAGES + 4 * QT) + 32 + 4 * RING;
}

int pick_stages(int kp) {
  int st = MAX_STAGES;
  while (st > 2 && tc_smem_bytes(kp, st) > 227 * 1024) --st;
  return st;
}

}  // namespace

// what the candidate stage hands to the re-rank / fallback stage (device pointers into the
// context's workspace; valid until the next kNN call on the context)
struct TcCandidates {
  int32_t *cand = nullptr;     // [nq x kp] candidate rows (-1: none)
  float *tau = nullptr;        // [nq] threshold below which every key is a candidate
  double *q_n2 = nullptr;      // [nq] |yh_q|^2
  double *q_err = nullptr;     // [nq] |y_q - yh_q|
  TcMisc *misc = nullptr;
  const int32_t *visit = nullptr;  // grouped search: queries in group-sorted order (-1: padding)
  int64_t n_visit = 0;
  bool have_groups = false;    // grouped search: the clustering of the reference rows
  GroupedRefs groups;
};

bool knn_tc_grouped(int64_t nr, int scan_mode);

bool knn_tc_applicable(int64_t nr, int64_t nq, int d, int k, int scan_mode) {
  if (d > MAX_D || d < 1) return false;
  const int kp = pick_kp(k, 0);
  // one list per row covers k <= 32; beyond that the grouped search runs per column split
  if (kp > MAX_KP && !(knn_tc_grouped(nr, scan_mode) && plan_kp(kp).splits > 0)) return false;
  if (nr < 2048 || nq < 1) return false;          // tiny problems: the exact scan is faster
  if (nr >= 0x7fffffffLL - BN || nq >= 0x7fffffffLL - QT * BM) return false;
  return true;
}

// launch helper for the candidate kernel
static int launch_tc(b200nn_ctx *ctx, const CUtensorMap &map_q, const CUtensorMap &map_r, const TcParams &P,
                     bool dbg) {
  size_t smem = tc_smem_bytes(P.kp, P.stages);
  if (smem > 227 * 1024 || P.stages < 3) {
    b200nn_set_error("candidate list of %d entries does not fit shared memory", P.kp);
    return B200NN_ERR_INVALID;
  }
  int grid = P.n_items < ctx->sm_count ? P.n_items : ctx->sm_count;
  if (grid <= 0) return B200NN_OK;
  auto kern = dbg ? k_knn_tc<true, false> : P.pass1 ? k_knn_tc<false, true> : k_knn_tc<false, false>;
  CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, NUM_THREADS, smem, ctx->stream>>>(map_q, map_r, P);
  KERNEL_CHECK(ctx);
  return B200NN_OK;
}

// B200NN_DEBUG=1: print the tile-list statistics of the grouped search to stderr
static int debug_list_stats(b200nn_ctx *ctx, const char *what, const int32_t *d_cnt, int n_items,
                            const float *dbg_T = nullptr, const float *dbg_R = nullptr) {
  static const bool on = getenv("B200NN_DEBUG") != nullptr;
  if (!on) return B200NN_OK;
  std::vector<int32_t> h((size_t)n_items);
  CU_TRY(cudaMemcpyAsync(h.data(), d_cnt, sizeof(int32_t) * (size_t)n_items, cudaMemcpyDeviceToHost, ctx->stream));
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  long long sum = 0;
  int mx = 0, mn = 1 << 30;
  std::vector<long long> per_cta((size_t)ctx->sm_count, 0);
  for (int i = 0; i < n_items; ++i) {
    sum += h[i];
    mx = h[i] > mx ? h[i] : mx;
    mn = h[i] < mn ? h[i] : mn;
    per_cta[(size_t)(i % ctx->sm_count)] += h[i];
  }
  long long cmax = 0;
  for (long long v : per_cta) cmax = v > cmax ? v : cmax;
  int big = 0;
  for (int i = 0; i < n_items; ++i) big += h[i] > 2 * (sum / n_items + 1);
  fprintf(stderr, "[b200nn] %s: %d items, tiles/item min %d mean %.1f max %d, total %lld, busiest CTA %lld (ideal %.0f), "
          "%d items above twice the mean\n",
          what, n_items, mn, (double)sum / n_items, mx, sum, cmax, (double)sum / ctx->sm_count, big);
  if (dbg_T) {
    std::vector<float> T((size_t)n_items), R((size_t)n_items);
    CU_TRY(cudaMemcpy(T.data(), dbg_T, sizeof(float) * (size_t)n_items, cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(R.data(), dbg_R, sizeof(float) * (size_t)n_items, cudaMemcpyDeviceToHost));
    int shown = 0;
    for (int i = 0; i < n_items && shown < 12; ++i)
      if (h[i] > 2 * (sum / n_items + 1)) {
        fprintf(stderr, "   item %d: %d tiles, T = %g, radius = %g\n", i, h[i], T[i], R[i]);
        shown++;
      }
    double tm = 0; int tc = 0;
    for (int i = 0; i < n_items; ++i) if (h[i] > 0 && h[i] <= 2 * (sum / n_items + 1)) { tm += T[i]; tc++; }
    fprintf(stderr, "   typical T = %g\n", tc ? tm / tc : 0.0);
  }
  return B200NN_OK;
}

static int pick_groups(int64_t nr) {
  int64_t want = nr / 4096;
  int G = 16;
  while (G < want && G < 1024) G <<= 1;
  return G;
}

// candidate stage alone (also used by the debug / profiling entry point).
// mode: 0 = auto, 1 = exhaustive scan of every tile, 2 = grouped (tile-pruned)
bool knn_tc_grouped(int64_t nr, int scan_mode) { return scan_mode == 2 || (scan_mode == 0 && nr >= 32768); }

// Reference-side state of the grouped search kept by a resident index (b200nn_index):
// centring / scale, the group-sorted FP16 operand, the grouping with its balls and tile annuli.
// One device blob; the segments are copied back into the workspace at the start of a later call,
// which then only prepares the queries.
struct TcRefCache {
  bool valid = false;
  int64_t nr = 0;
  int d = 0, G = 0, splits = 1;
  void *blob = nullptr;
  size_t bytes = 0;
};
TcRefCache *tc_ref_cache_create() { return new (std::nothrow) TcRefCache(); }
void tc_ref_cache_destroy(TcRefCache *c) {
  if (!c) return;
  if (c->blob) cudaFree(c->blob);
  delete c;
}

// per-call fields of the restored prep state
__global__ void k_misc_reset(TcMisc *misc) {
  misc->qn2_max = 0.0;
  misc->n_fail = 0;
  misc->hang = 0;
  misc->nonfinite = 0;
  misc->tiles_visited = 0;
  misc->tiles_total = 0;
}

static int knn_tc_candidates(b200nn_ctx *ctx, const double *d_ref, int64_t nr, const double *d_qry,
                             int64_t nq, int d, KpPlan kpl, int mode, TcCandidates *out,
                             float *d_dbg_keys, b200nn_stats *stats, int shard_rank = 0, int shard_world = 0,
                             TcRefCache *cache = nullptr) {
  cudaStream_t st = ctx->stream;
  bool timed_pass2 = false;
  const bool grouped = d_dbg_keys == nullptr && knn_tc_grouped(nr, mode);
  const bool sharded = shard_world > 1;
  if (!grouped || sharded) cache = nullptr;  // only the grouped search has reference-side state worth keeping
  // (the layout of the prepared reference side depends on the number of column splits)
  const bool cached = cache && cache->valid && cache->nr == nr && cache->d == d && cache->splits == kpl.splits;
  // with restored reference-side state the queries are prepared on their own, also for a self search
  const bool self = (d_qry == d_ref) && (nq == nr) && !cached;
  if (sharded && !(grouped && self && shard_world <= 64 && shard_rank >= 0 && shard_rank < shard_world)) {
    b200nn_set_error("sharded tensor search needs a self query in grouped mode (rank %d of %d)", shard_rank,
                     shard_world);
    return B200NN_ERR_INVALID;
  }
  *out = TcCandidates();
  TcMisc *misc;
  double *partial, *q_n2, *q_err;
  __half *r16, *q16;
  int32_t *cand;
  float *tau;
  RC_TRY(ws_get(ctx, WS_TC_MISC, 1, &misc));
  RC_TRY(ws_get(ctx, WS_SCAN_TMP, (size_t)MEAN_BLOCKS * BK, &partial));
  RC_TRY(ws_get(ctx, WS_TC_QRY16, (size_t)nq * BK, &q16));
  RC_TRY(ws_get(ctx, WS_TC_QRYERR, (size_t)nq * 2, &q_n2));
  q_err = q_n2 + nq;
  const int kp = kpl.total;
  if (kpl.splits < 1 || kpl.splits > MAX_SPLITS || kpl.per > MAX_KP || kpl.total != kpl.splits * kpl.per ||
      (kpl.splits > 1 && !grouped)) {
    b200nn_set_error("candidate lists of %d entries need %d column splits of the grouped search", kp, kpl.splits);
    return B200NN_ERR_INVALID;
  }
  RC_TRY(ws_get(ctx, WS_TC_CAND, (size_t)nq * kp, &cand));
  RC_TRY(ws_get(ctx, WS_TC_CANDKEY, (size_t)nq * kpl.splits, &tau));

  cudaEvent_t e0 = ctx->ev[8], e1 = ctx->ev[9], e2 = ctx->ev[10];
  CU_TRY(cudaEventRecord(e0, st));
  if (!cached) {
    k_mean_partial<<<MEAN_BLOCKS, 64, 0, st>>>(d_ref, nr, d, partial);
    KERNEL_CHECK(ctx);
    k_mean_final<<<1, 64, 0, st>>>(partial, nr, d, misc);
    KERNEL_CHECK(ctx);
    k_max_norm<<<(unsigned)ceil_div64(nr, 256), 256, 0, st>>>(d_ref, nr, d, misc);
    KERNEL_CHECK(ctx);
    // a resident index fixes centre and scale from the reference set alone: later query batches must
    // see the same operands (a far-away query only costs that row its certificate)
    if (!self && !cache) {
      k_max_norm<<<(unsigned)ceil_div64(nq, 256), 256, 0, st>>>(d_qry, nq, d, misc);
      KERNEL_CHECK(ctx);
    }
    k_pick_scale<<<1, 1, 0, st>>>(misc);
    KERNEL_CHECK(ctx);
    if (!self && cache) {  // non-finite queries must still be refused
      k_max_norm<<<(unsigned)ceil_div64(nq, 256), 256, 0, st>>>(d_qry, nq, d, misc);
      KERNEL_CHECK(ctx);
    }
  }
  // segments of the reference-side state, in a fixed order (saved after the first call on an index,
  // copied back at the start of the later ones)
  std::vector<std::pair<void *, size_t>> segs;
  segs.emplace_back((void *)misc, sizeof(TcMisc));

  TcParams P;
  memset(&P, 0, sizeof(P));
  P.nq = nq;
  P.nr = nr;
  P.n_items = (int32_t)ceil_div64(nq, QT * BM);
  P.ksteps = (d + 3 + 15) / 16;
  P.kp = kpl.per;
  P.stages = pick_stages(kpl.per);
  P.cand = cand;
  P.cand_stride = kp;
  P.tau = tau;
  P.dbg_keys = d_dbg_keys;
  P.misc = misc;
  P.q_n2 = q_n2;
  CUtensorMap map_q, map_r;

  if (!grouped) {
    const int64_t nr_pad = ceil_div64(nr, BN) * BN;  // whole tiles; the tail rows never qualify
    RC_TRY(ws_get(ctx, WS_TC_REF16, (size_t)nr_pad * BK, &r16));
    if (nr_pad > nr) {
      k_pad_ref<<<(unsigned)ceil_div64((nr_pad - nr) * BK, 128), 128, 0, st>>>(r16, nr, nr_pad, d);
      KERNEL_CHECK(ctx);
    }
    if (self) {
      k_convert<<<(unsigned)ceil_div64(nr, 128), 128, 0, st>>>(d_ref, nr, d, q16, r16, q_n2, q_err, misc,
                                                                nullptr, 1);
      KERNEL_CHECK(ctx);
    } else {
      k_convert<<<(unsigned)ceil_div64(nr, 128), 128, 0, st>>>(d_ref, nr, d, nullptr, r16, nullptr,
                                                                nullptr, misc, nullptr, 1);
      KERNEL_CHECK(ctx);
      k_convert<<<(unsigned)ceil_div64(nq, 128), 128, 0, st>>>(d_qry, nq, d, q16, nullptr, q_n2, q_err,
                                                                misc, nullptr, 0);
      KERNEL_CHECK(ctx);
    }
    CU_TRY(cudaEventRecord(e1, st));
    RC_TRY(make_operand_map(&map_q, q16, nq));
    RC_TRY(make_operand_map(&map_r, r16, nr_pad));
    P.n_btiles = (int32_t)(nr_pad / BN);
    RC_TRY(launch_tc(ctx, map_q, map_r, P, d_dbg_keys != nullptr));
  } else {
    // ---- grouped search ------------------------------------------------------------
    const int G = pick_groups(nr);
    const int64_t nr_pad = ((nr + BN - 1) / BN + (int64_t)G * kpl.splits) * BN;  // upper bound of the padded length
    const int64_t n_tiles_max = nr_pad / BN;
    const int64_t nq_pad = ((nq + QT * BM - 1) / (QT * BM) + G) * (QT * BM);
    const int64_t list_stride = (n_tiles_max + 3) & ~(int64_t)3;
    P.n_items = (int32_t)(nq_pad / (QT * BM));
    RC_TRY(ws_get(ctx, WS_TC_QRY16, (size_t)nq_pad * BK, &q16));
    // scratch: FP16 rows in original order (A format) for both sets, splits, group ids
    __half *yr16, *yq16;
    uint2 *split;
    int32_t *gid_r, *gid_q, *rperm, *qperm, *ibuf;
    float *fbuf;
    RC_TRY(ws_get(ctx, WS_TC_Y16R, (size_t)nr * BK, &yr16));
    RC_TRY(ws_get(ctx, WS_TC_SPLIT, (size_t)nr, &split));
    RC_TRY(ws_get(ctx, WS_TC_REF16, (size_t)nr_pad * BK, &r16));
    // int scratch: gid_r[nr] gid_q[nq] rperm[nr_pad] qperm[nq] + 8 small arrays of (G+1)
    const size_t nF = (size_t)G * FINE_NB;  // (group, bucket) bins of the hub-first order
    RC_TRY(ws_get(ctx, WS_TC_IBUF,
                  (size_t)nr + nq + nr_pad + nq_pad + 8 * (size_t)(G + 1) + 2 * (size_t)P.n_items +
                      (size_t)list_stride + 6 * nF + (size_t)(G + 1) + 64,
                  &ibuf));
    gid_r = ibuf;
    gid_q = gid_r + nr;
    rperm = gid_q + nq;
    qperm = rperm + nr_pad;
    int32_t *rcount = qperm + nq_pad, *qcount = rcount + (G + 1), *goff = qcount + (G + 1),
            *qoff = goff + (G + 1), *rcursor = qoff + (G + 1), *qcursor = rcursor + (G + 1),
            *kcount = qcursor + (G + 1), *list_cnt = kcount + (G + 1) * 2, *order = list_cnt + P.n_items,
            *bins = order + P.n_items;
    int32_t *fcount_r = bins + list_stride + 32, *fcount_q = fcount_r + nF, *foff_r = fcount_q + nF,
            *foff_q = foff_r + nF, *fcur_r = foff_q + nF, *fcur_q = fcur_r + nF, *owner_buf = fcur_q + nF;
    const int32_t *owner = sharded ? owner_buf : nullptr;
    // float scratch: cent[G*64] cnorm[G] sums[G*64] grad[G] item_c[nI*64] item_rad[nI] item_T[nI]
    //                gstat[2G] tile_lo/hi[n_tiles] anorm_r[nr] anorm_q[nq]
    const size_t nI = (size_t)P.n_items;
    RC_TRY(ws_get(ctx, WS_TC_FBUF,
                  (size_t)G * BK * 3 + 2 * (size_t)G + nI * BK + 2 * nI + 2 * (size_t)G + 2 * (size_t)n_tiles_max +
                      (size_t)nr + (size_t)nq + 64,
                  &fbuf));
    float *cent = fbuf, *cnorm = cent + (size_t)G * BK * 3, *grad = cnorm + G,
          *item_c = grad + G, *item_rad = item_c + nI * BK, *item_T = item_rad + nI;
    float *gstat = item_T + nI, *tile_lo = gstat + 2 * (size_t)G, *tile_hi = tile_lo + n_tiles_max,
          *anorm_r = tile_hi + n_tiles_max, *anorm_q = anorm_r + nr;
    int32_t *tiles;
    RC_TRY(ws_get(ctx, WS_TC_TILES, nI * (size_t)list_stride, &tiles));

    segs.emplace_back((void *)r16, sizeof(__half) * (size_t)nr_pad * BK);
    segs.emplace_back((void *)rperm, sizeof(int32_t) * (size_t)nr_pad);
    segs.emplace_back((void *)rcount, sizeof(int32_t) * (size_t)(G + 1));
    segs.emplace_back((void *)cent, sizeof(float) * (size_t)G * BK);
    segs.emplace_back((void *)grad, sizeof(float) * (size_t)G);
    segs.emplace_back((void *)gstat, sizeof(float) * 2 * (size_t)G);
    segs.emplace_back((void *)tile_lo, sizeof(float) * (size_t)n_tiles_max);
    segs.emplace_back((void *)tile_hi, sizeof(float) * (size_t)n_tiles_max);
    CU_TRY(cudaMemsetAsync(rcount, 0, sizeof(int32_t) * 8 * (size_t)(G + 1), st));
    if (cached) {
      if (cache->G != G) {
        b200nn_set_error("resident index: cached grouping does not match (%d groups, expected %d)", cache->G, G);
        return B200NN_ERR_STATE;
      }
      const char *src = static_cast<const char *>(cache->blob);
      for (auto &sg : segs) {
        CU_TRY(cudaMemcpyAsync(sg.first, src, sg.second, cudaMemcpyDeviceToDevice, st));
        src += (sg.second + 255) & ~(size_t)255;
      }
      k_misc_reset<<<1, 1, 0, st>>>(misc);
      KERNEL_CHECK(ctx);
      k_max_norm<<<(unsigned)ceil_div64(nq, 256), 256, 0, st>>>(d_qry, nq, d, misc);  // refuses NaN / Inf queries
      KERNEL_CHECK(ctx);
    }

    // 1. FP16 rows (original order), residuals, norms
    if (!cached) {
      k_convert<<<(unsigned)ceil_div64(nr, 128), 128, 0, st>>>(d_ref, nr, d, yr16, nullptr, self ? q_n2 : nullptr,
                                                                self ? q_err : nullptr, misc, split, 1);
      KERNEL_CHECK(ctx);
    }
    if (self) {
      yq16 = yr16;
    } else {
      RC_TRY(ws_get(ctx, WS_TC_Y16Q, (size_t)nq * BK, &yq16));
      k_convert<<<(unsigned)ceil_div64(nq, 128), 128, 0, st>>>(d_qry, nq, d, yq16, nullptr, q_n2, q_err, misc,
                                                                nullptr, 0);
      KERNEL_CHECK(ctx);
    }
    // 2. coarse k-means on a strided sample of the reference rows
    const int64_t m = nr < KM_SAMPLE ? nr : KM_SAMPLE;
    const int64_t stride = nr / m;
    long long *sums = reinterpret_cast<long long *>(cent + (size_t)G * BK);  // [G x 64] fixed-point sums
    if (!cached) {
      CU_TRY(cudaMemsetAsync(sums, 0, sizeof(long long) * (size_t)G * BK, st));
      k_km_init<<<G, 64, 0, st>>>(yr16, m * stride, d, G, cent, cnorm);
      KERNEL_CHECK(ctx);
    }
    static const bool assign_fma = getenv("B200NN_KM_ASSIGN_FMA") != nullptr;
    const int G_pad = ((G + BN - 1) / BN) * BN;
    __half *c16 = nullptr;
    if (!assign_fma || cached) RC_TRY(ws_get(ctx, WS_TC_CENT16, (size_t)G_pad * BK, &c16));
    // nearest centroid of n_rows rows (every row_stride-th row of y16) on the tensor cores
    auto assign_tc = [&](const __half *y16, int64_t n_rows, int64_t row_stride, int32_t *gid_out, int32_t *counts) -> int {
      k_cent_operand<<<(unsigned)ceil_div64((int64_t)G_pad * 32, 256), 256, 0, st>>>(cent, G, G_pad, d, c16);
      KERNEL_CHECK(ctx);
      CUtensorMap mq, mr;
      RC_TRY(make_operand_map(&mq, y16, n_rows, row_stride));
      RC_TRY(make_operand_map(&mr, c16, G_pad));
      TcParams A;
      memset(&A, 0, sizeof(A));
      A.nq = n_rows;
      A.nr = G_pad;
      A.n_items = (int32_t)ceil_div64(n_rows, QT * BM);
      A.n_btiles = G_pad / BN;
      A.ksteps = P.ksteps;
      A.kp = 8;
      A.stages = pick_stages(8);
      A.cand = cand;
      A.cand_stride = 8;
      A.tau = tau;
      A.misc = misc;
      A.argmin_out = gid_out;
      RC_TRY(launch_tc(ctx, mq, mr, A, false));
      if (counts) {
        k_count_gid<<<ctx->sm_count * 4, 256, 0, st>>>(gid_out, n_rows, G, counts);
        KERNEL_CHECK(ctx);
      }
      return B200NN_OK;
    };
    for (int itn = 0; itn < (cached ? 0 : KM_ITERS); ++itn) {
      if (assign_fma) {
        k_km_assign<<<(unsigned)ceil_div64(m, 128), 128, 0, st>>>(yr16, m, stride, d, G, cent, cnorm, nullptr,
                                                                   kcount, sums);
        KERNEL_CHECK(ctx);
      } else {
        // the Lloyd assignment of the sample also runs through the candidate kernel
        RC_TRY(assign_tc(yr16, m, stride, gid_r, nullptr));
        k_km_accum<<<(unsigned)ceil_div64(m, 128), 128, 0, st>>>(yr16, m, stride, d, G, gid_r, kcount, sums);
        KERNEL_CHECK(ctx);
      }
      k_km_update<<<G, 64, 0, st>>>(sums, kcount, d, cent, cnorm);
      KERNEL_CHECK(ctx);
    }
    // 3. assign every reference / query row to its nearest centroid, sort by group.  The assignment is
    // itself a nearest-neighbour search of n rows against G centroids: it runs on the tensor cores
    // through the candidate kernel (2 - 8 centroid tiles per work item), ~6x faster than the FP32
    // FMA kernel (k_km_assign; B200NN_KM_ASSIGN_FMA=1 keeps it for comparison).  Any assignment is
    // valid (only the pruning depends on it); this one is deterministic.
    if (cached) {
      // the reference rows keep their groups
    } else if (assign_fma) {
      k_km_assign<<<(unsigned)ceil_div64(nr, 128), 128, 0, st>>>(yr16, nr, 1, d, G, cent, cnorm, gid_r, rcount,
                                                                  nullptr);
      KERNEL_CHECK(ctx);
    } else {
      RC_TRY(assign_tc(yr16, nr, 1, gid_r, rcount));
    }
    if (self) {
      if (sharded) {  // this share takes whole groups
        k_group_owner<<<1, 1024, 0, st>>>(rcount, G, shard_world, shard_rank, owner_buf, qcount);
        KERNEL_CHECK(ctx);
      } else {
        CU_TRY(cudaMemcpyAsync(qcount, rcount, sizeof(int32_t) * G, cudaMemcpyDeviceToDevice, st));
      }
      gid_q = gid_r;
    } else if (assign_fma && !cached) {
      k_km_assign<<<(unsigned)ceil_div64(nq, 128), 128, 0, st>>>(yq16, nq, 1, d, G, cent, cnorm, gid_q, qcount,
                                                                  nullptr);
      KERNEL_CHECK(ctx);
    } else {
      RC_TRY(assign_tc(yq16, nq, 1, gid_q, qcount));
    }
    k_group_offsets<<<1, 32, 0, st>>>(rcount, qcount, G, goff, qoff, rcursor, qcursor, kpl.splits);
    KERNEL_CHECK(ctx);
    if (cached)  // rperm was restored
      CU_TRY(cudaMemsetAsync(qperm, 0xff, sizeof(int32_t) * (size_t)nq_pad, st));
    else
      CU_TRY(cudaMemsetAsync(rperm, 0xff, sizeof(int32_t) * (size_t)(nr_pad + nq_pad), st));  // rperm and qperm
    // 3b. squared distance of every row to its group's centre; groups are stored hub-first
    static const bool no_hub = getenv("B200NN_NO_HUB") != nullptr;  // arrival order (for comparison)
    if (!cached) {
      CU_TRY(cudaMemsetAsync(gstat, 0, sizeof(float) * 2 * (size_t)G, st));
      k_row_anorm<<<(unsigned)ceil_div64(nr, 128), 128, 0, st>>>(yr16, gid_r, nr, d, G, cent, anorm_r, gstat);
      KERNEL_CHECK(ctx);
    }
    if (self) {
      anorm_q = anorm_r;
    } else {
      k_row_anorm<<<(unsigned)ceil_div64(nq, 128), 128, 0, st>>>(yq16, gid_q, nq, d, G, cent, anorm_q, nullptr);
      KERNEL_CHECK(ctx);
    }
    if (no_hub && kpl.splits == 1) {
      if (!cached) {
        k_group_scatter<<<(unsigned)ceil_div64(nr, 256), 256, 0, st>>>(gid_r, nr, G, goff, rcursor, rperm);
        KERNEL_CHECK(ctx);
      }
      k_group_scatter<<<(unsigned)ceil_div64(nq, 256), 256, 0, st>>>(gid_q, nq, G, qoff, qcursor, qperm, owner,
                                                                      shard_rank);
      KERNEL_CHECK(ctx);
    } else {
      CU_TRY(cudaMemsetAsync(fcount_r, 0, sizeof(int32_t) * 2 * nF, st));  // fcount_r and fcount_q
      if (!cached) {
        k_group_stat_final<<<(G + 127) / 128, 128, 0, st>>>(gstat, rcount, G);
        KERNEL_CHECK(ctx);
        k_fine_count<<<(unsigned)ceil_div64(nr, 256), 256, 0, st>>>(gid_r, anorm_r, nr, gstat, fcount_r);
        KERNEL_CHECK(ctx);
        k_fine_offsets<<<G, FINE_NB, 0, st>>>(fcount_r, goff, foff_r, fcur_r);
        KERNEL_CHECK(ctx);
      }
      if (self) {
        k_fine_offsets<<<G, FINE_NB, 0, st>>>(fcount_r, qoff, foff_q, fcur_q);
        KERNEL_CHECK(ctx);
      } else {
        k_fine_count<<<(unsigned)ceil_div64(nq, 256), 256, 0, st>>>(gid_q, anorm_q, nq, gstat, fcount_q);
        KERNEL_CHECK(ctx);
        k_fine_offsets<<<G, FINE_NB, 0, st>>>(fcount_q, qoff, foff_q, fcur_q);
        KERNEL_CHECK(ctx);
      }
      if (!cached) {
        k_fine_scatter<<<(unsigned)ceil_div64(nr, 256), 256, 0, st>>>(gid_r, anorm_r, nr, gstat, foff_r, fcur_r, rperm,
                                                                       nullptr, 0, goff, kpl.splits);
        KERNEL_CHECK(ctx);
      }
      k_fine_scatter<<<(unsigned)ceil_div64(nq, 256), 256, 0, st>>>(gid_q, anorm_q, nq, gstat, foff_q, fcur_q, qperm,
                                                                     owner, shard_rank);
      KERNEL_CHECK(ctx);
    }
    // 4. operands in sorted order
    if (!cached) {
      k_gather_ref<<<(unsigned)ceil_div64(nr_pad * 8, 256), 256, 0, st>>>(yr16, split, rperm, nr_pad, d, r16);
      KERNEL_CHECK(ctx);
    }
    k_gather_qry<<<(unsigned)ceil_div64(nq_pad * 8, 256), 256, 0, st>>>(yq16, qperm, nq_pad, q16);
    KERNEL_CHECK(ctx);
    // 5. balls
    if (!cached) {
      CU_TRY(cudaMemsetAsync(grad, 0, sizeof(float) * (size_t)G, st));
      k_group_radius<<<dim3(G, 8), 128, 0, st>>>(yr16, rperm, goff, d, cent, grad);
      KERNEL_CHECK(ctx);
      k_tile_annulus<<<(unsigned)ceil_div64(n_tiles_max * 32, 256), 256, 0, st>>>(rperm, anorm_r, n_tiles_max, tile_lo,
                                                                                   tile_hi);
      KERNEL_CHECK(ctx);
    }
    k_item_stats<<<P.n_items, 64, 0, st>>>(q16, qperm, nq_pad, d, item_c, item_rad);
    KERNEL_CHECK(ctx);
    CU_TRY(cudaMemsetAsync(item_T, 0, sizeof(float) * nI, st));
    CU_TRY(cudaEventRecord(e1, st));
    // 6. pass 1: nearest group only -> T_item;  pass 2: pruned lists
    RC_TRY(make_operand_map(&map_q, q16, nq_pad));
    RC_TRY(make_operand_map(&map_r, r16, nr_pad));  // rows past the last group are never listed
    P.nq = nq_pad;
    P.nr = nr_pad;
    P.n_btiles = (int32_t)n_tiles_max;
    P.list_cnt = list_cnt;
    P.list_stride = list_stride;
    P.tiles = tiles;
    P.rperm = rperm;
    P.qperm = qperm;
    P.item_T = item_T;
    // one round per column split (a single round for k' <= MAX_KP): pass 1, lists, pass 2
    for (int sp = 0; sp < kpl.splits; ++sp) {
    if (sp > 0) CU_TRY(cudaMemsetAsync(item_T, 0, sizeof(float) * nI, st));
    P.cand = cand + (size_t)sp * kpl.per;
    P.tau = tau + (size_t)sp * nq;
    P.tau_init = nullptr;
    P.order = nullptr;
    k_build_lists<<<P.n_items, LIST_THREADS, 0, st>>>(item_c, item_rad, item_T, G, cent, grad, goff, tile_lo, tile_hi,
                                                      1, P.list_stride, tiles, list_cnt, kpl.splits, sp, rcount, 2 * kpl.per);
    KERNEL_CHECK(ctx);
    P.pass1 = 1;
    RC_TRY(debug_list_stats(ctx, "pass 1", list_cnt, P.n_items));
    RC_TRY(launch_tc(ctx, map_q, map_r, P, false));
    k_build_lists<<<P.n_items, LIST_THREADS, 0, st>>>(item_c, item_rad, item_T, G, cent, grad, goff, tile_lo, tile_hi,
                                                      0, P.list_stride, tiles, list_cnt, kpl.splits, sp, rcount, 2 * kpl.per);
    KERNEL_CHECK(ctx);
    out->visit = qperm;
    out->n_visit = nq_pad;
    out->groups.G = G;
    out->groups.ld = BK;
    out->groups.cent = cent;
    out->groups.grad = grad;
    out->groups.goff = goff;
    out->groups.rperm = rperm;
    out->groups.yq16 = reinterpret_cast<const uint16_t *>(yq16);
    out->groups.q_err = q_err;
    out->groups.scale = &misc->scale;
    out->groups.e_max = &misc->e_max;
    out->have_groups = true;
    P.pass1 = 0;
    P.tau_init = P.tau;  // written by pass 1, read at the start of every pass-2 row
    k_order_items<<<1, 1024, 0, st>>>(list_cnt, P.n_items, (int)n_tiles_max, bins, order, misc, sp > 0);
    KERNEL_CHECK(ctx);
    P.order = order;
    RC_TRY(debug_list_stats(ctx, "pass 2", list_cnt, P.n_items, item_T, item_rad));
    if (sp == 0) CU_TRY(cudaEventRecord(ctx->ev[14], st));  // several splits: from the first pass 2 to the last
    RC_TRY(launch_tc(ctx, map_q, map_r, P, false));
    if (sp + 1 == kpl.splits) CU_TRY(cudaEventRecord(ctx->ev[15], st));
    }  // column splits
    if (kpl.splits > 1) {
      k_tau_min<<<(unsigned)ceil_div64(nq, 256), 256, 0, st>>>(tau, nq, kpl.splits);
      KERNEL_CHECK(ctx);
    }
    timed_pass2 = true;
    if (cache && !cached) {
      // first call on a resident index: keep the reference-side state (stream-ordered copies)
      size_t total = 0;
      for (auto &sg : segs) total += (sg.second + 255) & ~(size_t)255;
      if (cache->blob && cache->bytes < total) {
        cudaFree(cache->blob);
        cache->blob = nullptr;
      }
      if (!cache->blob) {
        if (cudaMalloc(&cache->blob, total) != cudaSuccess) {
          (void)cudaGetLastError();
          cache->blob = nullptr;  // not fatal: the index simply stays uncached
        }
        cache->bytes = cache->blob ? total : 0;
      }
      if (cache->blob) {
        char *dst = static_cast<char *>(cache->blob);
        for (auto &sg : segs) {
          CU_TRY(cudaMemcpyAsync(dst, sg.first, sg.second, cudaMemcpyDeviceToDevice, st));
          dst += (sg.second + 255) & ~(size_t)255;
        }
        cache->nr = nr;
        cache->d = d;
        cache->G = G;
        cache->splits = kpl.splits;
        cache->valid = true;
      }
    }
  }
  CU_TRY(cudaEventRecord(e2, st));
  CU_TRY(cudaEventSynchronize(e2));
  if (stats) {
    float ms = 0;
    CU_TRY(cudaEventElapsedTime(&ms, e0, e1));
    stats->ms_prep += ms;
    CU_TRY(cudaEventElapsedTime(&ms, e1, e2));
    stats->ms_knn_cand += ms;
    if (timed_pass2) {
      CU_TRY(cudaEventElapsedTime(&ms, ctx->ev[14], ctx->ev[15]));
      stats->ms_knn_pass2 += ms;
    } else {
      stats->ms_knn_pass2 += ms;  // exhaustive scan: the one candidate launch
    }
  }
  out->cand = cand;
  out->tau = tau;
  out->q_n2 = q_n2;
  out->q_err = q_err;
  out->misc = misc;
  return B200NN_OK;
}

int knn_tc_launch(b200nn_ctx *ctx, const double *d_ref, int64_t nr, const double *d_qry, int64_t nq,
                  int d, int k, int n_cand, int scan_mode, int32_t *d_idx0, double *d_dist,
                  b200nn_stats *stats, int shard_rank, int shard_world, TcRefCache *cache) {
  cudaStream_t st = ctx->stream;
  int want = pick_kp(k, n_cand);
  if (!knn_tc_grouped(nr, scan_mode) && want > MAX_KP) want = MAX_KP;  // one list per row is all there is
  KpPlan kpl = plan_kp(want);
  if (kpl.splits == 0) kpl = plan_kp(pick_kp(k, 0));  // an oversized n_cand: the default width
  const int kp = kpl.total;
  if (kpl.splits == 0 || kp <= k) {
    b200nn_set_error("k = %d leaves no room for a widened candidate set", k);
    return B200NN_ERR_INVALID;
  }
  TcCandidates C;
  RC_TRY(knn_tc_candidates(ctx, d_ref, nr, d_qry, nq, d, kpl, scan_mode, &C, nullptr, stats, shard_rank,
                           shard_world, cache));
  if (shard_world > 1) {  // the rows this share wrote (the group-sorted query order, -1 = padding)
    ctx->owned_rows = C.visit;
    ctx->owned_slots = C.n_visit;
  }
  int32_t *cand = C.cand;
  float *tau = C.tau;
  double *q_n2 = C.q_n2, *q_err = C.q_err;
  TcMisc *misc = C.misc;
  int32_t *flags;
  RC_TRY(ws_get(ctx, WS_TC_FLAGS, (size_t)nq + 4, &flags));
  int32_t *fail_list = flags + 1;  // flags[0] receives the count (knn_exact reads list[-1])
  double *fail_tau;
  RC_TRY(ws_get(ctx, WS_TC_FAILTAU, (size_t)nq, &fail_tau));

  cudaEvent_t e0 = ctx->ev[11], e1 = ctx->ev[12], e2 = ctx->ev[13];
  CU_TRY(cudaEventRecord(e0, st));
  const size_t smem = rr_per_warp(d, kp) * RR_WARPS;
  const int64_t n_visit = C.visit ? C.n_visit : nq;
  const bool vec = (d % 2 == 0) && (reinterpret_cast<uintptr_t>(d_ref) % 16 == 0) &&
                   (reinterpret_cast<uintptr_t>(d_qry) % 16 == 0);
  auto kern = vec ? k_rerank<true> : k_rerank<false>;
  CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(unsigned)ceil_div64(n_visit, RR_WARPS), RR_WARPS * 32, smem, st>>>(
      d_ref, d_qry, nq, d, k, kp, cand, tau, q_n2, q_err, misc, d_idx0, d_dist, fail_list, fail_tau,
      C.visit, n_visit, ETA_PER_KSTEP * (double)((d + 3 + 15) / 16));
  KERNEL_CHECK(ctx);
  CU_TRY(cudaEventRecord(e1, st));
  TcMisc h;
  CU_TRY(cudaMemcpyAsync(&h, misc, sizeof(TcMisc), cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  if (h.hang) {
    b200nn_set_error("tensor-core kNN kernel: mbarrier wait timed out (pipeline protocol error)");
    return B200NN_ERR_CUDA;
  }
  if (h.nonfinite) {
    b200nn_set_error("NA/NaN/Inf in foreign function call (the embedding holds a non-finite value)");
    return B200NN_ERR_INVALID;
  }
  if (h.n_fail > 0) {
    CU_TRY(cudaMemcpyAsync(flags, &misc->n_fail, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    if (C.have_groups)
      RC_TRY(knn_exact_grouped_launch(ctx, d_ref, d_qry, d, k, fail_list, h.n_fail, fail_tau, C.groups, d_idx0,
                                      d_dist));
    else
      RC_TRY(knn_exact_launch(ctx, d_ref, nr, d_qry, nq, d, k, fail_list, h.n_fail, d_idx0, d_dist, fail_tau));
  }
  CU_TRY(cudaEventRecord(e2, st));
  CU_TRY(cudaEventSynchronize(e2));
  if (stats) {
    float ms = 0;
    CU_TRY(cudaEventElapsedTime(&ms, e0, e1));
    stats->ms_knn_rerank += ms;
    CU_TRY(cudaEventElapsedTime(&ms, e1, e2));
    stats->ms_knn_fallback += ms;
    stats->n_uncertified = h.n_fail;
    const long long all = (long long)ceil_div64(nq, QT * BM) * ceil_div64(nr, BN);
    stats->knn_tiles_visited = h.tiles_total ? h.tiles_visited : all;
    stats->knn_tiles_total = all;
  }
  return B200NN_OK;
}

// diagnostic entry point: raw keys of (item 0, B tile 0) + candidates + prep state
int knn_tc_debug(b200nn_ctx *ctx, const double *d_ref, int64_t nr, const double *d_qry, int64_t nq,
                 int d, int kp, int scan_mode, float *d_keys_out, int32_t *d_cand_out, float *d_tau_out,
                 double *h_scale_mu /* [1 + 64] */) {
  TcCandidates C;
  const KpPlan one = {1, kp, kp};
  RC_TRY(knn_tc_candidates(ctx, d_ref, nr, d_qry ? d_qry : d_ref, nq, d, one, d_keys_out ? 1 : scan_mode, &C,
                           d_keys_out, nullptr));
  int32_t *cand = C.cand;
  float *tau = C.tau;
  void *misc_v = C.misc;
  cudaStream_t st = ctx->stream;
  if (d_cand_out)
    CU_TRY(cudaMemcpyAsync(d_cand_out, cand, sizeof(int32_t) * (size_t)nq * kp,
                           cudaMemcpyDeviceToDevice, st));
  if (d_tau_out)
    CU_TRY(cudaMemcpyAsync(d_tau_out, tau, sizeof(float) * (size_t)nq, cudaMemcpyDeviceToDevice, st));
  TcMisc h;
  CU_TRY(cudaMemcpyAsync(&h, misc_v, sizeof(TcMisc), cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  if (h_scale_mu) {
    h_scale_mu[0] = h.scale;
    for (int j = 0; j < BK; ++j) h_scale_mu[1 + j] = h.mu[j];
  }
  if (h.hang) {
    b200nn_set_error("tensor-core kNN kernel: mbarrier wait timed out");
    return B200NN_ERR_CUDA;
  }
  return B200NN_OK;
}
