// engine.cuh -- the per-GPU handle: configuration, HBM-resident state, scratch.
#pragma once
#include <vector>
#include <string>
#include "common.cuh"
#include "mlp_simt.cuh"
#include "dp_peer.cuh"
#include <nvtx3/nvToolsExt.h>

namespace cql {

constexpr int N_NOISE_PARTS = 8;

struct StepInfo {          // written by k_step_begin, read by the Adam kernels
  double bc1, bc2_sqrt;    // 1 - beta1^t, sqrt(1 - beta2^t)
  long long step;          // t (1-based) of the update in flight
};

// NVTX range (SURVEY.md section 5 "Tracing / profiling"): the phases of an update, scoring and ingestion show up by name
// on an Nsight Systems / ncu --nvtx timeline.  Header-only NVTX3: a no-op unless a tool is attached.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

struct MdpSession;                   // chunked ingestion session (mdp_gpu.cu)
// seen-items CSR of a log on the device (mdp_gpu.cu); returns the number of (user, item) entries written
int64_t seen_csr_on_device(struct Handle& h, const int32_t* users_h, const int32_t* items_h, int64_t n, int64_t n_users,
                           const uint8_t* wanted_h, int64_t* indptr_d, int32_t* seen_d, cudaStream_t st);
void mdp_session_free(MdpSession*);

struct Handle {
  cql_config cfg{};
  std::string err;
  cudaStream_t own_stream = nullptr;
  cudaStream_t side_stream = nullptr;            // fork/join branch for independent small launches (captured into the graphs)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int num_sms = 148;
  int64_t launches = 0;

  // ---- learner state (flat layout, see common.cuh) ----
  float* params = nullptr;   // state_floats(C)
  float* adam_m = nullptr;   // same layout (trainable part used)
  float* adam_v = nullptr;
  float* grads = nullptr;    // grad_floats(C): [actor | critics | scalars]
  long long* step_dev = nullptr;   // completed updates
  StepInfo* stepinfo = nullptr;
  float* metrics = nullptr;  // 8 floats: temp_loss,temp,alpha_loss,alpha,critic_loss,actor_loss,td_loss,-
  float* metrics_host = nullptr;   // pinned

  // ---- replay table ----
  float* table = nullptr;    // [n_trans][8]: obs.x obs.y act rew nobs.x nobs.y term pad
  int64_t n_trans = 0;
  int64_t table_cap = 0;     // rows the table allocation can hold (re-ingesting a log of the same size re-uses it)
  bool table_sharded = false;   // data parallel: this rank holds only its own users' episodes (cql_set_table_sharded)
  int64_t sample_pos_host = 0;  // position in the epoch stream for the stand-alone sampler
  long long* sample_pos = nullptr;   // device: next position in the permutation stream

  // ---- per-step scratch (B = batch, n = action samples, C = critics) ----
  int B = 0, n = 0, C = 0;
  int rsA = 0, rsC = 0;      // rows per batch element: alpha job (3n), critic job (3n+1)
  float* batch = nullptr;    // [B][8] sampled transition rows (same row format as table)
  float* batch_host = nullptr;   // pinned staging for cql_update_batch
  float* noise = nullptr;    // packed, see header
  float* noise_host = nullptr;
  int64_t noise_floats = 0;
  float4* XA = nullptr;      // [2B] actor inputs: s rows then s' rows
  float* outA = nullptr;     // [2B][2] mu, raw logstd
  float* h2A = nullptr;      // [tiles(B)][256][64] actor H2 of the s rows
  float4* XAl = nullptr;     // [B*rsA] alpha-step critic rows
  float* offAl = nullptr;    // [B*rsA] log-prob offsets
  float* QAl = nullptr;      // [C][B*rsA]
  float4* XC = nullptr;      // [B*rsC] critic-step rows (3n samples + data row)
  float* offC = nullptr;     // [B*rsC]
  float* QC = nullptr;       // [C][B*rsC]
  float* h2C = nullptr;      // [C][tiles][256][64]
  float* dQ = nullptr;       // [C][B*rsC]
  float4* XT = nullptr;      // [B] target rows (s', tanh(mu(s')))
  float* QT = nullptr;       // [C][B]
  float4* XP = nullptr;      // [B] actor-step rows (s, a_pi)
  float* QP = nullptr;       // [C][B]
  float* h2P = nullptr;      // [C][tiles(B)][256][64]
  float* dQP = nullptr;      // [C][B]
  float4* dXP = nullptr;     // [C][B]
  float* dOutA = nullptr;    // [B][2]
  float* perb = nullptr;     // [B][4]: temp term, logp_pi, a_pi raw, -
  float* pairv = nullptr;    // [C][B] PairVals {lse_alpha, lse_critic, q_data, td_err}
  float* loss_sums = nullptr;  // [16] reduced sums / conservative coefficient; [14], [15] = block tickets (k_actor_dout, k_lse)
  float* actor_part = nullptr; // [ceil(B / 128) * 4] warp sums of the actor-loss metric (k_actor_dout)
  float* smallC = nullptr;   // [C][tilesC][SMALL_STRIDE]
  float* smallA = nullptr;   // [tilesB][SMALL_STRIDE]
  float* pw2C = nullptr;     // [C][splitsC][H*H]
  float* pw2A = nullptr;     // [splitsA][H*H]
  int splitsC = 1, splitsA = 1;

  // ---- tensor-core path (precision != FP32): packed W2 per slot, layer-3 partial sums ----
  uint8_t* packed_fwd = nullptr;   // [(2+2C) slots][PACKED_NET_BYTES]  B[n][k] = W2[n][k]
  size_t packed_net_bytes = 0;     // stride of packed_fwd
  size_t packed_net_bytes_bwd = 0; // stride of packed_bwd
  float* part = nullptr;           // scratch for [n_nets][SLICES][rows][OUT] partials
  size_t part_floats = 0;
  uint8_t* packed_bwd = nullptr;   // [(1+C) slots: actor, critics][PACKED_NET_BYTES]  B[n][k] = W2[k][n]
  float* small1 = nullptr;         // [C][slots1][SMALL_STRIDE] bwd1 partials (W1 | b1), slots1 = 4 * num_sms
  float* small2 = nullptr;         // [C][splits_tc][SMALL_STRIDE] bwd2 partials (b2 | W3 | b3)
  float* pw2_tc = nullptr;         // [C][splits_tc][H*H]
  float4* dX_part = nullptr;       // [C*SLICES][B]
  unsigned int* b2_tickets = nullptr;   // [C][64] group tickets of the f16x3 bwd2 kernel
  int slots1 = 0, splits_tc = 0, tc_slices = 1;
  int last_dx_parts = 0;           // parts (nets x column slices) of dX_part written by the last dx launch
  // CTA-pair (cta_group::2) f16x3 kernels: W2 of every slot in the pair layout (mlp_tc_h2.cuh)
  uint8_t* packed_fwd2 = nullptr;  // [(2+2C) slots][H2Cfg::PACKED_NET_BYTES]  B[n][k] = W2[n][k]
  uint8_t* packed_bwd2 = nullptr;  // [(1+C) slots]                           B[n][k] = W2[k][n]
  size_t packed_net_bytes2 = 0;
  DpPeer dp;                       // data-parallel peers (cql_dp_attach); world <= 1: single GPU
  bool dp_fused = false;           // the gradient exchange happens INSIDE the update kernels (f16x3 path, dp_peer.cuh)
  int* w2max = nullptr;            // [(2+2C) slots][4]: rotating max|W2| slots of the fused Adam + pack kernel (adam_pack.cuh)
  int pair_swap_b = 0;             // which cluster rank holds the first half of the operand rows (probed at create)

  // optional event marks for cql_timed_update
  bool timing = false;
  cudaEvent_t ev[16] = {};

  MdpSession* mdp = nullptr;       // between cql_mdp_begin and cql_mdp_finish
  uint8_t* mdp_ring = nullptr;     // pinned staging ring of the ingestion sessions
  std::vector<void*> allocs;

  template <typename T>
  T* dalloc(size_t count) {
    void* p = nullptr;
    CQL_CUDA(cudaMalloc(&p, count * sizeof(T)));
    CQL_CUDA(cudaMemset(p, 0, count * sizeof(T)));
    allocs.push_back(p);
    return reinterpret_cast<T*>(p);
  }
  void free_all() {
    for (void* p : allocs) cudaFree(p);
    allocs.clear();
    if (mdp) { mdp_session_free(mdp); mdp = nullptr; }
    if (mdp_ring) { cudaFreeHost(mdp_ring); mdp_ring = nullptr; }
    if (table) { cudaFree(table); table = nullptr; table_cap = 0; }
    if (metrics_host) { cudaFreeHost(metrics_host); metrics_host = nullptr; }
    if (batch_host) { cudaFreeHost(batch_host); batch_host = nullptr; }
    if (noise_host) { cudaFreeHost(noise_host); noise_host = nullptr; }
    if (own_stream) { cudaStreamDestroy(own_stream); own_stream = nullptr; }
    if (side_stream) { cudaStreamDestroy(side_stream); side_stream = nullptr; }
    if (ev_fork) { cudaEventDestroy(ev_fork); ev_fork = nullptr; }
    if (ev_join) { cudaEventDestroy(ev_join); ev_join = nullptr; }
    for (auto& e : ev) if (e) { cudaEventDestroy(e); e = nullptr; }
  }

  float* net_params(int slot) const { return params + (size_t)slot * NET_STRIDE; }
  float* scalars() const { return params + scalars_off(C); }
  // grads buffer parts
  float* g_actor() const { return grads; }
  float* g_critics() const { return grads + NET_STRIDE; }
  float* g_scalars() const { return grads + (size_t)(1 + C) * NET_STRIDE; }
};

void mdp_begin(Handle& h, int64_t n);
void mdp_append(Handle& h, int col, int dtype, const void* host, int64_t count);
int64_t mdp_finish(Handle& h, int top_k, float noise_scale, float* obs_out, float* act_out, float* rew_out, float* term_out,
                   int64_t* order_out);

// replay table for n rows: the existing allocation is kept when it is large enough (a 640 MB cudaFree + cudaMalloc per
// ingestion synchronises the device and costs milliseconds to tens of milliseconds)
inline void ensure_table(Handle& h, int64_t n) {
  h.n_trans = 0;
  if (h.table != nullptr && h.table_cap >= n) return;
  if (h.table) { CQL_CUDA(cudaFree(h.table)); h.table = nullptr; h.table_cap = 0; }
  CQL_CUDA(cudaMalloc(&h.table, (size_t)n * 8 * sizeof(float)));
  h.table_cap = n;
}

inline void mark(Handle* h, cudaStream_t st, int i) {
  if (h->timing) CQL_CUDA(cudaEventRecord(h->ev[i], st));
}

inline cudaStream_t pick_stream(Handle* h, void* stream) {
  return stream ? reinterpret_cast<cudaStream_t>(stream) : h->own_stream;
}

#define CQL_LAUNCH_CHECK(h)                 \
  do {                                      \
    (h)->launches++;                        \
    CQL_CUDA(cudaGetLastError());           \
  } while (0)

}  // namespace cql
