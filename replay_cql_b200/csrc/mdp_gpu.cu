// mdp_gpu.cu -- MDP builder on the GPU (SURVEY.md section 8 row a7 / "next" item 1).
//
// Replaces the wrapper's two Spark window sorts + global orderBy + toPandas ([EXT] MdpDatasetBuilder.build) with
// stable LSD radix sorts on the device (CUB: library code, like cuBLAS -- sorting is not the hot path's arithmetic):
//   order      = stable sort by (user, timestamp)                 : sort by ts, then by user
//   rank order = stable sort by (user, relevance desc, ts desc)   : sort by ts desc, rel desc, user
// followed by one pass that writes the 32-byte transition rows (reward = rank-in-user < top_k, terminal = last
// row of the user, action = relevance + noise, next_obs = next row of the same user).  Ties keep input order,
// exactly like the host builder (replay_cql_b200/mdp.py), which the tests hold it bit-exact against.
#include <algorithm>
#include <initializer_list>
#include <vector>
#include <cub/cub.cuh>
#include "engine.cuh"

using namespace cql;

namespace {

__global__ void k_iota(uint32_t* p, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (uint32_t)i;
}
template <typename T>
__global__ void k_gather(const T* __restrict__ src, const uint32_t* __restrict__ idx, int64_t n, T* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}
// head position of every run of equal users in a user-sorted sequence (0 elsewhere) -> max-scan = group start
__global__ void k_heads(const int32_t* __restrict__ u_sorted, int64_t n, int64_t* __restrict__ head) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) head[i] = (i == 0 || u_sorted[i] != u_sorted[i - 1]) ? i : 0;
}
__global__ void k_rewarded(const uint32_t* __restrict__ rank_order, const int64_t* __restrict__ group_start, int64_t n,
                           int top_k, uint8_t* __restrict__ rewarded) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) rewarded[rank_order[i]] = (i - group_start[i]) < top_k ? 1 : 0;
}
struct MaxOp {
  __device__ __forceinline__ int64_t operator()(int64_t a, int64_t b) const { return a > b ? a : b; }
};

__global__ void k_emit(const uint32_t* __restrict__ order, const int32_t* __restrict__ user, const int32_t* __restrict__ item,
                       const double* __restrict__ rel, const double* __restrict__ noise, const uint8_t* __restrict__ rewarded,
                       int64_t n, float noise_scale, uint64_t seed, float4* __restrict__ table, float* __restrict__ obs_o,
                       float* __restrict__ act_o, float* __restrict__ rew_o, float* __restrict__ term_o) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t row = order[i];
  const int32_t u = user[row];
  const bool last = (i == n - 1) || user[order[i + 1]] != u;
  double nz;
  if (noise) {
    nz = noise[row];
  } else {                                   // seeded N(0,1) * scale per ORIGINAL row
    uint32_t r[4];
    Philox::gen(seed ^ 0xA0761D6478BD642Full, 0x4D4450ull, (uint64_t)row, r);
    nz = (double)(sqrtf(-2.f * logf(u01(r[0]))) * cospif(2.f * u01(r[1])) * noise_scale);
  }
  const float act = (float)((double)(float)rel[row] + nz);
  const float rw = rewarded[row] ? 1.f : 0.f;
  const float uf = (float)u, itf = (float)item[row];
  float nu = 0.f, ni = 0.f;
  if (!last) { nu = uf; ni = (float)item[order[i + 1]]; }
  table[2 * i] = make_float4(uf, itf, act, rw);
  table[2 * i + 1] = make_float4(nu, ni, last ? 1.f : 0.f, 0.f);
  if (obs_o) { obs_o[2 * i] = uf; obs_o[2 * i + 1] = itf; act_o[i] = act; rew_o[i] = rw; term_o[i] = last ? 1.f : 0.f; }
}

template <typename T>
T* dmalloc(size_t n, std::vector<void*>& pool) {
  void* p = nullptr;
  CQL_CUDA(cudaMalloc(&p, std::max<size_t>(1, n) * sizeof(T)));
  pool.push_back(p);
  return reinterpret_cast<T*>(p);
}

}  // namespace

// returns the number of kernels launched (for the handle's launch counter)
int64_t mdp_build_on_device(Handle& h, const int32_t* user_h, const int32_t* item_h, const int64_t* ts_h, const double* rel_h,
                            const double* noise_h, int64_t n, int top_k, float noise_scale, float* obs_out, float* act_out,
                            float* rew_out, float* term_out, int64_t* order_out) {
  CQL_REQUIRE(n >= 1 && n < (1ll << 31), "cql_build_mdp: n must be in 1..2^31-1");
  CQL_REQUIRE(user_h && item_h && ts_h && rel_h, "cql_build_mdp: NULL column");
  cudaStream_t st = h.own_stream;
  std::vector<void*> pool;
  int64_t launches = 0;
  try {
    int32_t* user = dmalloc<int32_t>(n, pool);
    int32_t* item = dmalloc<int32_t>(n, pool);
    int64_t* ts = dmalloc<int64_t>(n, pool);
    double* rel = dmalloc<double>(n, pool);
    double* noise = noise_h ? dmalloc<double>(n, pool) : nullptr;
    CQL_CUDA(cudaMemcpyAsync(user, user_h, n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CQL_CUDA(cudaMemcpyAsync(item, item_h, n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CQL_CUDA(cudaMemcpyAsync(ts, ts_h, n * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    CQL_CUDA(cudaMemcpyAsync(rel, rel_h, n * sizeof(double), cudaMemcpyHostToDevice, st));
    if (noise) CQL_CUDA(cudaMemcpyAsync(noise, noise_h, n * sizeof(double), cudaMemcpyHostToDevice, st));

    uint32_t* ia = dmalloc<uint32_t>(n, pool);
    uint32_t* ib = dmalloc<uint32_t>(n, pool);
    int64_t* k64a = dmalloc<int64_t>(n, pool);
    int64_t* k64b = dmalloc<int64_t>(n, pool);
    double* kda = dmalloc<double>(n, pool);
    double* kdb = dmalloc<double>(n, pool);
    int32_t* k32a = dmalloc<int32_t>(n, pool);
    int32_t* k32b = dmalloc<int32_t>(n, pool);
    uint32_t* order = dmalloc<uint32_t>(n, pool);
    uint8_t* rewarded = dmalloc<uint8_t>(n, pool);
    int64_t* head = k64a;            // reused after the timestamp sorts
    int64_t* gstart = k64b;

    // one temp buffer big enough for every CUB call below
    size_t t1 = 0, t2 = 0, t3 = 0, t4 = 0, t5 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t1, ts, k64a, ia, ib, (int)n, 0, 64, st);
    cub::DeviceRadixSort::SortPairs(nullptr, t2, k32a, k32b, ia, ib, (int)n, 0, 32, st);
    cub::DeviceRadixSort::SortPairsDescending(nullptr, t3, ts, k64a, ia, ib, (int)n, 0, 64, st);
    cub::DeviceRadixSort::SortPairsDescending(nullptr, t4, kda, kdb, ia, ib, (int)n, 0, 64, st);
    cub::DeviceScan::InclusiveScan(nullptr, t5, head, gstart, MaxOp(), (int)n, st);
    size_t tbytes = t1;
    for (size_t t : {t2, t3, t4, t5}) tbytes = t > tbytes ? t : tbytes;
    void* temp = dmalloc<uint8_t>(tbytes, pool);
    const unsigned nb = (unsigned)((n + 255) / 256);
    size_t tb;

    // ---- order: (user, ts, input order)
    k_iota<<<nb, 256, 0, st>>>(ia, n);
    tb = tbytes; CQL_CUDA(cub::DeviceRadixSort::SortPairs(temp, tb, ts, k64a, ia, ib, (int)n, 0, 64, st));
    k_gather<int32_t><<<nb, 256, 0, st>>>(user, ib, n, k32a);
    tb = tbytes; CQL_CUDA(cub::DeviceRadixSort::SortPairs(temp, tb, k32a, k32b, ib, order, (int)n, 0, 32, st));
    // ---- rank order: (user, rel desc, ts desc, input order)
    k_iota<<<nb, 256, 0, st>>>(ia, n);
    tb = tbytes; CQL_CUDA(cub::DeviceRadixSort::SortPairsDescending(temp, tb, ts, k64a, ia, ib, (int)n, 0, 64, st));
    k_gather<double><<<nb, 256, 0, st>>>(rel, ib, n, kda);
    tb = tbytes; CQL_CUDA(cub::DeviceRadixSort::SortPairsDescending(temp, tb, kda, kdb, ib, ia, (int)n, 0, 64, st));
    k_gather<int32_t><<<nb, 256, 0, st>>>(user, ia, n, k32a);
    tb = tbytes; CQL_CUDA(cub::DeviceRadixSort::SortPairs(temp, tb, k32a, k32b, ia, ib, (int)n, 0, 32, st));   // ib = rank order
    k_heads<<<nb, 256, 0, st>>>(k32b, n, head);
    tb = tbytes; CQL_CUDA(cub::DeviceScan::InclusiveScan(temp, tb, head, gstart, MaxOp(), (int)n, st));
    k_rewarded<<<nb, 256, 0, st>>>(ib, gstart, n, top_k, rewarded);
    launches += 20;

    // ---- emit transition rows straight into the replay table
    if (h.table) { CQL_CUDA(cudaFree(h.table)); h.table = nullptr; h.n_trans = 0; }
    CQL_CUDA(cudaMalloc(&h.table, (size_t)n * 8 * sizeof(float)));
    float *d_obs = nullptr, *d_act = nullptr, *d_rew = nullptr, *d_term = nullptr;
    if (obs_out) {
      CQL_REQUIRE(act_out && rew_out && term_out, "cql_build_mdp: give all four output columns or none");
      d_obs = dmalloc<float>(2 * n, pool); d_act = dmalloc<float>(n, pool);
      d_rew = dmalloc<float>(n, pool); d_term = dmalloc<float>(n, pool);
    }
    k_emit<<<nb, 256, 0, st>>>(order, user, item, rel, noise, rewarded, n, noise_scale, h.cfg.seed,
                               reinterpret_cast<float4*>(h.table), d_obs, d_act, d_rew, d_term);
    CQL_CUDA(cudaGetLastError());
    launches += 1;
    if (obs_out) {
      CQL_CUDA(cudaMemcpyAsync(obs_out, d_obs, 2 * n * sizeof(float), cudaMemcpyDeviceToHost, st));
      CQL_CUDA(cudaMemcpyAsync(act_out, d_act, n * sizeof(float), cudaMemcpyDeviceToHost, st));
      CQL_CUDA(cudaMemcpyAsync(rew_out, d_rew, n * sizeof(float), cudaMemcpyDeviceToHost, st));
      CQL_CUDA(cudaMemcpyAsync(term_out, d_term, n * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    if (order_out) {
      // widen on the host: order is uint32 on the device
      std::vector<uint32_t> tmp((size_t)n);
      CQL_CUDA(cudaMemcpyAsync(tmp.data(), order, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
      CQL_CUDA(cudaStreamSynchronize(st));
      for (int64_t i = 0; i < n; ++i) order_out[i] = tmp[(size_t)i];
    }
    CQL_CUDA(cudaStreamSynchronize(st));
    h.n_trans = n;
  } catch (...) {
    for (void* p : pool) cudaFree(p);
    throw;
  }
  for (void* p : pool) cudaFree(p);
  return launches;
}
