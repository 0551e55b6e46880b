// metrics.cuh -- ranking metrics over the U x k output of predict, on the GPU (the step right
// after the hot path; lets optimize() trials stay off Spark).  Per-user definitions follow the reference
// `_get_metric_value_by_user` of replay/metrics/{ndcg.py:51-61, hitrate.py, map.py, mrr.py, precision.py, recall.py}
// and the right-join semantics of get_enriched_recommendations (base_metric.py:102-140): every ground-truth user
// counts, users without recommendations score 0.  One thread per user (k is tiny), double arithmetic in the
// reference's summation order; the mean over users is a fixed-order tree.
#pragma once
#include "common.cuh"

namespace cql {

enum { MET_NDCG = 0, MET_HITRATE, MET_MAP, MET_MRR, MET_PRECISION, MET_RECALL, MET_COUNT };
constexpr int MET_MAX_KS = 8;

struct MetricArgs {
  const int32_t* rec_items;    // [U][k_rec], best first, padded with -1
  const int32_t* users;        // [U] user id of each row
  const int64_t* gt_indptr;    // CSR over user id
  const int32_t* gt_items;     // sorted ascending per user
  int64_t n_users;
  int k_rec, n_ks;
  int ks[MET_MAX_KS];
  double* per_user;            // [MET_COUNT][n_ks][U]
};

__device__ __forceinline__ bool in_sorted(const int32_t* __restrict__ a, int64_t lo, int64_t hi, int v) {
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    const int x = __ldg(a + mid);
    if (x == v) return true;
    if (x < v) lo = mid + 1; else hi = mid;
  }
  return false;
}

__global__ void k_rank_metrics(const MetricArgs a) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= a.n_users) return;
  const int user = a.users[u];
  const int64_t lo = a.gt_indptr[user], hi = a.gt_indptr[user + 1];
  const int64_t n_gt = hi - lo;
  const int32_t* rec = a.rec_items + u * a.k_rec;
  int n_pred = 0;
  while (n_pred < a.k_rec && rec[n_pred] >= 0) ++n_pred;
  for (int q = 0; q < a.n_ks; ++q) {
    const int k = a.ks[q];
    const int len = min(k, n_pred);
    double dcg = 0.0, ap = 0.0;
    int hits = 0, first = -1;
    for (int i = 0; i < len; ++i) {
      if (!in_sorted(a.gt_items, lo, hi, rec[i])) continue;
      ++hits;
      if (first < 0) first = i;
      dcg += 1.0 / log2((double)(i + 2));
      ap += (double)hits / (double)(i + 1);
    }
    double idcg = 0.0;
    const int gl = (int)min((int64_t)k, n_gt);
    for (int i = 0; i < gl; ++i) idcg += 1.0 / log2((double)(i + 2));
    const bool empty = n_pred == 0 || n_gt == 0;
    double* o = a.per_user + (size_t)q * a.n_users + u;
    const size_t ms = (size_t)a.n_ks * a.n_users;
    o[MET_NDCG * ms] = empty ? 0.0 : dcg / idcg;
    o[MET_HITRATE * ms] = hits > 0 ? 1.0 : 0.0;
    o[MET_MAP * ms] = empty ? 0.0 : ap / (double)k;
    o[MET_MRR * ms] = first >= 0 ? 1.0 / (double)(1 + first) : 0.0;
    o[MET_PRECISION * ms] = n_pred == 0 ? 0.0 : (double)hits / (double)k;
    o[MET_RECALL * ms] = n_gt == 0 ? 0.0 : (double)hits / (double)n_gt;
  }
}

// mean over users of each (metric, k) series: one block per series, strided partial sums + fixed-order tree
__global__ void __launch_bounds__(256) k_metric_mean(const double* __restrict__ per_user, int64_t n_users, double* __restrict__ out) {
  __shared__ double part[256];
  const double* p = per_user + (size_t)blockIdx.x * n_users;
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n_users; i += 256) s += p[i];
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = n_users > 0 ? part[0] / (double)n_users : 0.0;
}

}  // namespace cql
