// topk_staged.cuh -- stand-alone top-k + lazy seen filter over a MATERIALISED score matrix [rows][n_items]
// (reference: `_filter_seen` replay/models/base_rec.py:417-464 followed by `get_top_k` replay/utils.py:100-109,
// here as one HBM-bound pass: 4 bytes read per (user, item) pair).
//
// Design (k <= 32).  Persistent CTAs (6 per SM) walk rows round-robin; warp-specialised:
//  * producer warp: streams the row, cut into chunks of <= CH floats, into a STAGES-deep shared-memory ring
//    with 1-D bulk copies (TMA, cp.async.bulk + full/empty mbarriers) -- loads never wait for the selection;
//  * 4 worker warps, pass 1: every thread keeps the three best float4 maxima of its strided share of the
//    WHOLE row, where the best two are, and a copy of those two float4s in shared memory; nothing else happens per
//    chunk.  At the end of the row 4-lane groups reduce to 32 group maxima;
//  * selector warp: sorts the 32 group maxima; the (k + spare)-th largest is a threshold T reached by at
//    least k + spare elements;
//  * workers, pass 2: only threads whose maximum reaches T revisit their best (two) float4s (from the shared-memory
//    stash; their whole share, from global memory, only if the third-best reaches T too), resolve the item id, drop seen items (binary search
//    in a shared-memory copy of the user's seen list -- sampled, plus one window of global memory, for lists over
//    512 items) and push (score, item) candidates -- about k + 8 per row;
//  * selector: places the candidates by rank (no serial insertion on the common path) and verifies: if the
//    k-th entry reaches T nothing below T can belong to the top-k; otherwise (seen items ate the candidates)
//    T is lowered and only the band [T_new, T_old) is collected in another round.  Candidate overflow
//    (massive ties) falls back to an exact sequential scan of the band by the selector.
//  * the next row's user id, CSR bounds and seen list (cp.async) are prefetched one row ahead.
// Hand-over between the roles uses named barriers (bar.arrive / bar.sync producer-consumer idiom).
// Order: relevance desc, then item id asc -- bit-identical to the other top-k kernels (score.cuh).
#pragma once
#include "score.cuh"
#include "tc_common.cuh"

namespace cql {

constexpr int TKS_WORKERS = 128;      // 4 worker warps + selector warp + producer warp
constexpr int TKS_CTAS_PER_SM = 6;
constexpr int TKS_GROUP = TKS_WORKERS / 32;   // lanes per group maximum (32 groups per CTA)
constexpr int TKS_SEL_WARP = TKS_WORKERS / 32;
constexpr int TKS_PROD_WARP = TKS_SEL_WARP + 1;
constexpr int TKS_THREADS = TKS_WORKERS + 64;
constexpr int TKS_SYNC = TKS_WORKERS + 32;   // participants of the worker <-> selector barriers
constexpr int TKS_STAGES = 2;
constexpr int TKS_CH = 3072;          // floats per stage: 2 x 12 KB ring, 6 CTAs per SM (measured 5.03 TB/s; 2 x 16 KB x 5 CTAs: 4.72; 3 x 16 KB x 4: 4.54)
constexpr int TKS_SEEN = 512;         // ints per seen cache (two caches: current row / next row)
constexpr int TKS_CAP = 256;          // candidates per collection round

struct TksCtl {
  float t_lo, t_hi;
  int ncand[2], done;      // candidate counters, double-buffered by row parity
  int seen_n;        // current row: length of the seen list, -1 = no filter
  int seen_stride;   // 1: the whole list sits in the row's shared-memory cache; > 1: every stride-th element does
  int seen_w;        // number of cached entries
  long long seen_lo; // current row: offset of the list in seen_items
  int cache_sel;     // which of the two caches
};

constexpr size_t TKS_OFF_SEEN = (size_t)TKS_STAGES * TKS_CH * 4;
constexpr size_t TKS_OFF_CS = TKS_OFF_SEEN + 2 * TKS_SEEN * 4;
constexpr size_t TKS_OFF_CI = TKS_OFF_CS + TKS_CAP * 4;
constexpr size_t TKS_OFF_G = TKS_OFF_CI + TKS_CAP * 4;       // gmax[32] | sorted[32] | listS[32] | listI[32]
constexpr size_t TKS_OFF_BEST = TKS_OFF_G + 128 * 4;   // float4 best[2][TKS_WORKERS]: every worker's best two float4s of the row
constexpr size_t TKS_OFF_CTL = TKS_OFF_BEST + 2 * TKS_WORKERS * 16;
constexpr size_t TKS_OFF_BAR = TKS_OFF_CTL + 64;
constexpr size_t TKS_SMEM = TKS_OFF_BAR + 2 * TKS_STAGES * 8;

struct TksArgs {
  const float* scores;
  int64_t n_rows, n_items;
  const int32_t* users;
  const int32_t* items;
  const int64_t* seen_indptr;
  const int32_t* seen_items;
  int k, chunk_len, nchunks, aligned;
  float* out_s;
  int* out_i;
};

// Seen list of the current row.  Lists of up to TKS_SEEN items are cached whole (stride 1).  A longer list is cached
// SAMPLED: entry i is the last (largest) element of window i = [i * stride, min((i + 1) * stride, n)), stride =
// ceil(n / TKS_SEEN).  A search is the branch-free lower bound over the cached entries plus, for a sampled list, a
// binary search inside ONE window of global memory (stride - 1 ints, one or two sectors: one L2/DRAM round trip instead
// of ~12 dependent ones -- 3 % of the ML-20M-shaped users have more than 512 seen items, and a CTA that met two of
// them finished 30 us after the others).
struct TksSeen {
  const int32_t* g;   // global list
  const int32_t* s;   // shared-memory cache (whole or sampled)
  int64_t lo;         // offset of the list in g
  int n, stride, w;   // length (< 0: no filter), sampling stride, number of cached entries
};
__device__ __forceinline__ int tks_stride(int64_t n) { return n <= TKS_SEEN ? 1 : (int)((n + TKS_SEEN - 1) / TKS_SEEN); }
__device__ __forceinline__ int tks_entries(int64_t n, int stride) { return (int)((n + stride - 1) / stride); }
__device__ __forceinline__ int64_t tks_entry_index(int i, int stride, int64_t n) {   // list index cached as entry i
  return min((int64_t)(i + 1) * stride, n) - 1;
}
__device__ __forceinline__ bool tks_is_seen(const TksSeen& v, int item) {
  const int w = v.w;
  int lo = 0;
#pragma unroll
  for (int step = TKS_SEEN / 2; step >= 1; step >>= 1) {     // fixed trip count: searching lanes stay converged
    const int p = lo + step;
    if (p <= w && v.s[p - 1] < item) lo = p;
  }
  if (lo >= w) return false;
  if (v.s[lo] == item) return true;
  if (v.stride == 1) return false;
  const int64_t wlo = v.lo + (int64_t)lo * v.stride;
  return is_seen(v.g, wlo, v.lo + tks_entry_index(lo, v.stride, v.n), item);   // the window without its last element
}

// named barriers: producers `arrive`, consumers `sync` (PTX producer/consumer idiom)
enum { TKS_BAR_GM = 1, TKS_BAR_TR = 2, TKS_BAR_CD = 3, TKS_BAR_WK = 4 };
__device__ __forceinline__ void nbar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void nbar_arrive(int id, int n) {
  __threadfence_block();
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory");
}

__global__ void __launch_bounds__(TKS_THREADS, TKS_CTAS_PER_SM) k_topk_filter_staged(const TksArgs a) {
  extern __shared__ __align__(128) uint8_t tks_sm[];
  float* ring = reinterpret_cast<float*>(tks_sm);
  int32_t* seen_cache = reinterpret_cast<int32_t*>(tks_sm + TKS_OFF_SEEN);
  float* candS = reinterpret_cast<float*>(tks_sm + TKS_OFF_CS);
  int* candI = reinterpret_cast<int*>(tks_sm + TKS_OFF_CI);
  float* gmax = reinterpret_cast<float*>(tks_sm + TKS_OFF_G);
  float* gsorted = gmax + 32;
  float* listS = gmax + 64;
  int* listI = reinterpret_cast<int*>(gmax + 96);
  volatile TksCtl* ctl = reinterpret_cast<volatile TksCtl*>(tks_sm + TKS_OFF_CTL);
  uint64_t* full = reinterpret_cast<uint64_t*>(tks_sm + TKS_OFF_BAR);
  uint64_t* empty = full + TKS_STAGES;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int k = a.k;
  const int64_t nrows_cta = a.n_rows > blockIdx.x ? (a.n_rows - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t total = nrows_cta * a.nchunks;
  if (total == 0) return;
  auto row_of = [&](int64_t j) { return (int64_t)blockIdx.x + j * gridDim.x; };
  const bool items_al = (reinterpret_cast<uintptr_t>(a.items) & 15) == 0;
  const int rank0 = min(32, k + 4 + (k >> 1));     // spare candidates so that seen items rarely force another round

  if (tid == 0) {
    for (int s = 0; s < TKS_STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], TKS_WORKERS / 32); }
    tc::fence_mbar_init();
    ctl->ncand[0] = 0;
    ctl->ncand[1] = 0;
    ctl->done = 0;
  }
  __syncthreads();

  if (warp == TKS_PROD_WARP) {
    // =========================== producer: bulk copies into the ring ===========================
    if (!a.aligned) return;
    int64_t j = 0;
    int c = 0;
    for (int64_t q = 0; q < total; ++q) {
      const int s = (int)(q % TKS_STAGES);
      if (q >= TKS_STAGES) tc::mbar_wait(&empty[s], (uint32_t)(((q / TKS_STAGES) - 1) & 1));
      if (lane == 0) {
        const int64_t c0 = (int64_t)c * a.chunk_len;
        const uint32_t len = (uint32_t)min((int64_t)a.chunk_len, a.n_items - c0);
        tc::mbar_arrive_expect_tx(&full[s], len * 4);
        tc::bulk_g2s(ring + (size_t)s * TKS_CH, a.scores + (size_t)row_of(j) * a.n_items + c0, len * 4, &full[s]);
      }
      __syncwarp();
      if (++c == a.nchunks) { c = 0; ++j; }
    }
    return;
  }

  if (warp < TKS_SEL_WARP) {
    // =========================== workers: pass 1 (maxima) and pass 2 (candidates) ===========================
    auto fetch_ids = [&](int64_t gv, int* it) {                              // item ids of float4 `gv` of a row
      const int64_t col0 = gv * 4;
      if (a.items != nullptr && items_al && col0 + 3 < a.n_items) {
        const int4 y = __ldg(reinterpret_cast<const int4*>(a.items) + gv);
        it[0] = y.x; it[1] = y.y; it[2] = y.z; it[3] = y.w;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          it[e] = a.items ? (col0 + e < a.n_items ? __ldg(a.items + col0 + e) : -1) : (int)(col0 + e);
      }
    };
    auto fetch4 = [&](const float* rowp, int64_t gv, float* sc, int* it) {   // scores and item ids of float4 `gv` of a row
      const int64_t col0 = gv * 4;
      if (a.aligned) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(rowp) + gv);
        sc[0] = x.x; sc[1] = x.y; sc[2] = x.z; sc[3] = x.w;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) sc[e] = col0 + e < a.n_items ? __ldg(rowp + col0 + e) : -INFINITY;
      }
      fetch_ids(gv, it);
    };
    // the best two float4s of the share are stashed in shared memory as they are found (a handful of stores per
    // row): re-reading them from global memory after the row missed L2 (ncu: +6.5 % DRAM reads) and put a DRAM
    // round trip on the row's hand-over
    float4* best = reinterpret_cast<float4*>(tks_sm + TKS_OFF_BEST) + tid;
    int64_t q = 0;
    for (int64_t j = 0; j < nrows_cta; ++j) {
      const float* rowp = a.scores + (size_t)row_of(j) * a.n_items;
      // pass 1 over the whole row: best and second-best float4 maximum of this thread's strided share
      float m = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;   // best / second / third float4 maximum of the share
      int gvbest = 0, gv2 = 0;                                // float4 index (within the row) of the best two
      int sb = 0;                                             // stash slot of the best (the runner-up sits in the other)
      for (int c = 0; c < a.nchunks; ++c, ++q) {
        const int64_t c0 = (int64_t)c * a.chunk_len;
        const int len = (int)min((int64_t)a.chunk_len, a.n_items - c0);
        const int nvec = (len + 3) >> 2;
        const int gv0 = (int)(c0 >> 2);
        float* buf = ring + (size_t)(q % TKS_STAGES) * TKS_CH;
        if (a.aligned) {
          tc::mbar_wait(&full[q % TKS_STAGES], (uint32_t)((q / TKS_STAGES) & 1));
        } else {                                           // unaligned rows: plain loads, tail padded with -inf
          for (int e = tid; e < nvec * 4; e += TKS_WORKERS) buf[e] = e < len ? __ldcs(rowp + c0 + e) : -INFINITY;
          nbar_sync(TKS_BAR_WK, TKS_WORKERS);
        }
        const float4* b4 = reinterpret_cast<const float4*>(buf);
#pragma unroll 4
        for (int v = tid; v < nvec; v += TKS_WORKERS) {
          const float4 x = b4[v];
          const float mx = fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w));
          if (mx > m) { m3 = m2; m2 = m; gv2 = gvbest; m = mx; gvbest = gv0 + v; sb ^= 1; best[sb * TKS_WORKERS] = x; }
          else if (mx > m2) { m3 = m2; m2 = mx; gv2 = gv0 + v; best[(sb ^ 1) * TKS_WORKERS] = x; }
          else m3 = fmaxf(m3, mx);
        }
        if (a.aligned) {
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&empty[q % TKS_STAGES]);
        }
      }
      // the best two float4s come back from the stash (slots never written hold garbage, but then m / m2 = -inf and
      // they are not offered); their item ids are read speculatively (L2 hits), overlapped with the threshold hand-over
      float sc[8];
      int it[8];
      {
        const float4 x = best[sb * TKS_WORKERS], y = best[(sb ^ 1) * TKS_WORKERS];
        sc[0] = x.x; sc[1] = x.y; sc[2] = x.z; sc[3] = x.w;
        sc[4] = y.x; sc[5] = y.y; sc[6] = y.z; sc[7] = y.w;
      }
      fetch_ids(gvbest, it);
      fetch_ids(gv2, it + 4);
      float g = m;
#pragma unroll
      for (int o = 1; o < TKS_GROUP; o <<= 1) g = fmaxf(g, __shfl_xor_sync(0xffffffffu, g, o));
      if ((lane & (TKS_GROUP - 1)) == 0) gmax[warp * (32 / TKS_GROUP) + lane / TKS_GROUP] = g;
      nbar_sync(TKS_BAR_GM, TKS_SYNC);                    // group maxima (and the selector's row info) are out
      // round 0: every warp derives the threshold itself = the rank0-th largest group maximum (rank by counting)
      float t_lo, t_hi = INFINITY;
      {
        const float mine = gmax[lane];
        int pos = 0;
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          const float o = gmax[t];
          pos += (o > mine || (o == mine && t < lane)) ? 1 : 0;
        }
        const unsigned hit = __ballot_sync(0xffffffffu, pos == rank0 - 1);
        t_lo = __shfl_sync(0xffffffffu, mine, __ffs(hit) - 1);
      }
      volatile int* ncand = &ctl->ncand[j & 1];
      for (int round = 0;; ++round) {
        if (round > 0) {
          nbar_sync(TKS_BAR_TR, TKS_SYNC);                // the selector's verdict, or the next band
          if (ctl->done) break;
          t_lo = ctl->t_lo;
          t_hi = ctl->t_hi;
        }
        if (m >= t_lo) {
          // pass 2: normally only the best float4 holds anything >= t_lo; rescan the share when the runner-up does too
          const int seen_n = ctl->seen_n;
          TksSeen sv{a.seen_items, seen_cache, 0, seen_n, 1, 0};
          if (seen_n >= 0) {
            sv.lo = ctl->seen_lo;
            sv.stride = ctl->seen_stride;
            sv.w = ctl->seen_w;
            sv.s = seen_cache + ctl->cache_sel * TKS_SEEN;
          }
          if (m3 < t_lo) {
            // in-band elements of the best two float4s as a bit mask, then ONE search per set bit: searching inside a
            // loop over the 8 elements made every warp run the 9-step seen search up to 8 times in a row (the lanes'
            // hits sit at different positions), ~1.5 us per row that the 2-stage ring could not hide
            unsigned band = 0;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int64_t col = (int64_t)(e < 4 ? gvbest : gv2) * 4 + (e & 3);
              const bool in = col < a.n_items && sc[e] >= t_lo && sc[e] < t_hi && (e < 4 || m2 >= t_lo);
              band |= in ? 1u << e : 0u;
            }
            while (band) {
              const int e = __ffs(band) - 1;
              band &= band - 1;
              float s = sc[0];
              int item = it[0];
#pragma unroll
              for (int u = 1; u < 8; ++u) {
                s = u == e ? sc[u] : s;
                item = u == e ? it[u] : item;
              }
              if (seen_n >= 0 && tks_is_seen(sv, item)) continue;
              const int p = atomicAdd(const_cast<int*>(ncand), 1);
              if (p < TKS_CAP) { candS[p] = s; candI[p] = item; }
            }
          }
        }
        // three or more float4s of a thread's share reach the band (a few % of the rows): the warp rescans that
        // share together -- chunk starts are multiples of TKS_WORKERS float4s, so the share is gv = tid + i * TKS_WORKERS
        __syncwarp();
        unsigned need = __ballot_sync(0xffffffffu, m >= t_lo && m3 >= t_lo);
        if (need) {
          const int seen_n = ctl->seen_n;
          TksSeen sv{a.seen_items, seen_cache, 0, seen_n, 1, 0};
          if (seen_n >= 0) {
            sv.lo = ctl->seen_lo;
            sv.stride = ctl->seen_stride;
            sv.w = ctl->seen_w;
            sv.s = seen_cache + ctl->cache_sel * TKS_SEEN;
          }
          const int64_t nvec_row = (a.n_items + 3) >> 2;
          while (need) {
            const int src = __ffs(need) - 1;
            need &= need - 1;
            for (int64_t gv = warp * 32 + src + (int64_t)lane * TKS_WORKERS; gv < nvec_row; gv += 32 * TKS_WORKERS) {
              float s4[4];
              int i4[4];
              fetch4(rowp, gv, s4, i4);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                if (gv * 4 + e >= a.n_items || !(s4[e] >= t_lo && s4[e] < t_hi)) continue;
                if (seen_n >= 0 && tks_is_seen(sv, i4[e])) continue;
                const int p = atomicAdd(const_cast<int*>(ncand), 1);
                if (p < TKS_CAP) { candS[p] = s4[e]; candI[p] = i4[e]; }
              }
            }
          }
        }
        nbar_sync(TKS_BAR_CD, TKS_SYNC);                  // candidates of this round are out
        // k unseen candidates at or above t_lo settle the row (nothing below t_lo can enter the top-k): in that case
        // nobody waits for the selector, which places the candidates while the next row streams in
        const int n = *ncand;
        if (round == 0 && n >= k && n <= TKS_CAP) break;
      }
    }
    return;
  }

  // =========================== selector warp: thresholds, placement, verification, output ===========================
  const bool has_seen = a.seen_indptr != nullptr;
  auto user_of = [&](int64_t j) -> int {
    if (j >= nrows_cta) return 0;
    const int64_t r = row_of(j);
    return a.users ? __ldg(a.users + r) : (int)r;
  };
  // row metadata pipeline: (lo0,hi0) current row, (lo1,hi1) next row, u2 = user of the row after that
  int64_t lo0 = 0, hi0 = 0, lo1 = 0, hi1 = 0;
  int u2 = 0;
  if (has_seen) {
    const int u0 = user_of(0);
    lo0 = __ldg(a.seen_indptr + u0);
    hi0 = __ldg(a.seen_indptr + u0 + 1);
    {
      const int64_t n0 = hi0 - lo0;
      const int st0 = tks_stride(n0), w0 = tks_entries(n0, st0);
      for (int i = lane; i < w0; i += 32) seen_cache[i] = __ldg(a.seen_items + lo0 + tks_entry_index(i, st0, n0));
    }
    if (nrows_cta > 1) {
      const int u1 = user_of(1);
      lo1 = __ldg(a.seen_indptr + u1);
      hi1 = __ldg(a.seen_indptr + u1 + 1);
    }
    u2 = user_of(2);
    __syncwarp();
  }

  // the row's list: lane e holds the e-th best; lanes >= k stay (-inf, -1)
  float ms = -INFINITY, thr_s = -INFINITY;
  int mi = -1, thr_i = -1;
  auto insert = [&](float s2, int i2) {   // converged warp; caller checked better(s2, i2, thr_s, thr_i)
    const int pos = __popc(__ballot_sync(0xffffffffu, better(ms, mi, s2, i2)));
    const float us = __shfl_up_sync(0xffffffffu, ms, 1);
    const int ui = __shfl_up_sync(0xffffffffu, mi, 1);
    if (lane == pos) { ms = s2; mi = i2; }
    else if (lane > pos && lane < k) { ms = us; mi = ui; }
    thr_s = __shfl_sync(0xffffffffu, ms, k - 1);
    thr_i = __shfl_sync(0xffffffffu, mi, k - 1);
  };

  for (int64_t j = 0; j < nrows_cta; ++j) {
    const int64_t row = row_of(j);
    const float* rowp = a.scores + (size_t)row * a.n_items;
    // prefetch: next row's seen list (cp.async into the other cache), CSR bounds of the row after, user id 3 ahead
    int64_t lo2 = 0, hi2 = 0;
    int u3 = 0;
    if (has_seen) {
      const int64_t n1 = j + 1 < nrows_cta ? hi1 - lo1 : 0;
      const int st1 = tks_stride(n1), w1 = tks_entries(n1, st1);
      int32_t* nc = seen_cache + ((j + 1) & 1) * TKS_SEEN;
      for (int e = lane; e < w1; e += 32)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(tc::smem_u32(nc + e)),
                     "l"(a.seen_items + lo1 + tks_entry_index(e, st1, n1)) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (j + 2 < nrows_cta) {
        lo2 = __ldg(a.seen_indptr + u2);
        hi2 = __ldg(a.seen_indptr + u2 + 1);
      }
      u3 = user_of(j + 3);
    }
    const int st0 = tks_stride(hi0 - lo0);
    const TksSeen sv{a.seen_items, seen_cache + (j & 1) * TKS_SEEN, lo0, has_seen ? (int)(hi0 - lo0) : -1, st0, tks_entries(hi0 - lo0, st0)};
    if (lane == 0) {                        // row info for the workers (read after the row's first TR barrier)
      ctl->seen_n = sv.n;
      ctl->seen_stride = sv.stride;
      ctl->seen_w = sv.w;
      ctl->seen_lo = lo0;
      ctl->cache_sel = (int)(j & 1);
    }
    ms = -INFINITY; mi = -1; thr_s = -INFINITY; thr_i = -1;
    nbar_sync(TKS_BAR_GM, TKS_SYNC);
    // sort the 32 group maxima descending by counting (independent broadcast reads, no dependent chain)
    {
      const float mine = gmax[lane];
      int pos = 0;
#pragma unroll
      for (int t = 0; t < 32; ++t) {
        const float o = gmax[t];
        pos += (o > mine || (o == mine && t < lane)) ? 1 : 0;
      }
      gsorted[pos] = mine;
      __syncwarp();
    }
    if (lane == 0) ctl->ncand[(j + 1) & 1] = 0;          // next row's counter (every worker has read its last value)
    volatile int* ncand = &ctl->ncand[j & 1];
    int rank = rank0;
    float t_hi = INFINITY, t_lo = gsorted[rank - 1];     // round 0: the workers derived the same band themselves
    bool settled = false;
    for (int round = 0;; ++round) {
      if (round > 0) {
        if (lane == 0) { ctl->t_lo = t_lo; ctl->t_hi = t_hi; *ncand = 0; ctl->done = 0; }
        nbar_arrive(TKS_BAR_TR, TKS_SYNC);
      }
      nbar_sync(TKS_BAR_CD, TKS_SYNC);
      const int n = *ncand;
      if (round == 0 && n >= k && n <= TKS_CAP) settled = true;     // same predicate as the workers: no verdict needed
      if (n > TKS_CAP) {
        // overflow (massive ties): exact sequential scan of the band, straight from global memory
        for (int64_t base = 0; base < a.n_items; base += 128) {
          float sv4[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int64_t e = base + u * 32 + lane;
            sv4[u] = e < a.n_items ? __ldg(rowp + e) : -INFINITY;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int64_t e = base + u * 32 + lane;
            const float s = sv4[u];
            unsigned mm = __ballot_sync(0xffffffffu, e < a.n_items && s >= t_lo && s < t_hi && s >= thr_s);
            while (mm) {
              const int src = __ffs(mm) - 1;
              mm &= mm - 1;
              const float s2 = __shfl_sync(0xffffffffu, s, src);
              const int64_t col = base + u * 32 + src;
              const int i2 = a.items ? __ldg(a.items + col) : (int)col;
              if (!better(s2, i2, thr_s, thr_i)) continue;
              if (has_seen && tks_is_seen(sv, i2)) continue;
              insert(s2, i2);
            }
          }
        }
      } else if (round == 0 && n <= 32) {
        // common path: the list is empty -- place the candidates by rank (no serial insertion chain)
        const float s = lane < n ? candS[lane] : -INFINITY;
        const int item = lane < n ? candI[lane] : -1;
        int r = 0;
#pragma unroll 4
        for (int t = 0; t < n; ++t) r += better(candS[t], candI[t], s, item) ? 1 : 0;
        listS[lane] = -INFINITY;
        listI[lane] = -1;
        __syncwarp();
        if (lane < n && r < k) { listS[r] = s; listI[r] = item; }
        __syncwarp();
        ms = listS[lane];
        mi = listI[lane];
        thr_s = __shfl_sync(0xffffffffu, ms, k - 1);
        thr_i = __shfl_sync(0xffffffffu, mi, k - 1);
      } else {
        for (int base = 0; base < n; base += 32) {        // candidates are unseen and carry their item id
          const int e = base + lane;
          const float s = e < n ? candS[e] : -INFINITY;
          const int item = e < n ? candI[e] : -1;
          unsigned mm = __ballot_sync(0xffffffffu, e < n && better(s, item, thr_s, thr_i));
          while (mm) {
            const int src = __ffs(mm) - 1;
            mm &= mm - 1;
            const float s2 = __shfl_sync(0xffffffffu, s, src);
            const int i2 = __shfl_sync(0xffffffffu, item, src);
            if (better(s2, i2, thr_s, thr_i)) insert(s2, i2);
          }
        }
      }
      // nothing below t_lo can belong to the top-k once the k-th entry reaches t_lo
      if ((thr_i >= 0 && thr_s >= t_lo) || t_lo == -INFINITY) break;
      // seen items ate the candidates: lower the threshold, collect only the new band next round
      const int n_ok = __popc(__ballot_sync(0xffffffffu, mi >= 0 && ms >= t_lo));
      rank += max(2, 2 * (k - n_ok));
      while (rank <= 32 && gsorted[rank - 1] >= t_lo) ++rank;
      t_hi = t_lo;
      t_lo = rank <= 32 ? gsorted[rank - 1] : -INFINITY;
    }
    if (!settled) {
      if (lane == 0) ctl->done = 1;
      nbar_arrive(TKS_BAR_TR, TKS_SYNC);
    }
    if (lane < k) { a.out_s[row * k + lane] = ms; a.out_i[row * k + lane] = mi; }
    if (has_seen) {
      asm volatile("cp.async.wait_all;" ::: "memory");
      __syncwarp();
      lo0 = lo1; hi0 = hi1; lo1 = lo2; hi1 = hi2; u2 = u3;
    }
  }
}

}  // namespace cql
