// dp_peer.cuh -- one-shot all-reduce(mean) of a gradient group over NVLink peer memory (data-parallel training).
//
// The three collectives of a CQL update are sub-MB and latency-bound.  Instead of an NCCL call, every rank
// launches ONE kernel (k_dp_exchange) that
//   publishes : copies its gradient group into ITS symmetric staging buffer (double-buffered by an epoch parity);
//               the last block fences system-wide and stores the new epoch into every peer's signal pad;
//   reduces   : waits until its own signal pad shows that epoch from every peer, then reads all ranks' staging
//               buffers directly (P2P loads through NVLink / NVSwitch), sums them in RANK ORDER, divides by the
//               world size and writes the local gradient buffer -- every rank computes bit-identical averages, so
//               the Adam steps that follow keep the replicas bit-identical.
// Staging[parity] is rewritten two epochs later; a rank can only get there after it has seen every peer's NEXT
// epoch, which a peer publishes only after it finished reading this one -- no second barrier is needed.
// The buffers come from torch.distributed._symmetric_memory (host side: parallel.PeerGradExchange); the library only
// sees raw device pointers.  A wait that does not complete within ~2 s sets an error flag instead of hanging.
#pragma once
#include "common.cuh"

namespace cql {

constexpr int DP_MAX_WORLD = 16;
constexpr int DP_GROUPS = 4;

struct DpPeer {
  int world = 0, rank = 0;
  float* stage[DP_MAX_WORLD] = {};                 // every rank's staging buffer [2][stage_floats] (mine at [rank])
  unsigned long long* sig[DP_MAX_WORLD] = {};      // every rank's signal pad: slot sender * DP_GROUPS + group
  long long stage_floats = 0;
  unsigned long long* epoch = nullptr;             // local [DP_GROUPS] completed epochs
  unsigned int* ticket = nullptr;                  // local [2 * DP_GROUPS] block tickets
  int* error = nullptr;                            // local: set when a wait timed out
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_volatile4(const float* p) {   // bypasses L1: peer data changes between launches
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// ---- fused mode (f16x3 path): no exchange launch at all.  The kernel that PRODUCES a gradient group (k_scalar_reduce,
// k_reduce_grads_tc) also writes it into this rank's staging buffer and -- its last block -- signals the peers
// (dp_publish_done); the kernel that CONSUMES the group (k_scalar_adam_dq, k_adam_pack) waits for every peer's signal
// (dp_wait_peers, one thread per block) and reads the mean straight from the peers' staging buffers over NVLink
// (dp_mean4 / dp_mean1), in rank order, so every rank applies bit-identical gradients; its last block advances the
// epoch (dp_consume_done).  Same double-buffering argument as above: a rank rewrites staging[parity] two epochs later,
// after it has seen every peer's next signal, which a peer sends only after its consumer kernel of this epoch ended.
__device__ __forceinline__ long long dp_base(const DpPeer& p, int group, long long off) {
  return (long long)(p.epoch[group] & 1ull) * p.stage_floats + off;
}
// call by ONE thread after all of this rank's staging writes of the group are visible device-wide
__device__ __forceinline__ void dp_signal_peers(const DpPeer& p, int group) {
  const unsigned long long e = p.epoch[group];
  __threadfence_system();
  for (int r = 0; r < p.world; ++r)
    if (r != p.rank) st_release_sys(p.sig[r] + p.rank * DP_GROUPS + group, e + 1);
}
// multi-block producer: every block calls this after its staging writes; the last one signals
__device__ __forceinline__ void dp_publish_done(const DpPeer& p, int group, unsigned int n_blocks) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(&p.ticket[group], 1u);
    if (t == n_blocks - 1) {
      p.ticket[group] = 0;
      dp_signal_peers(p, group);
    }
  }
}
// every block of a consumer kernel, before its first staging read (contains a __syncthreads)
__device__ __forceinline__ void dp_wait_peers(const DpPeer& p, int group) {
  if (threadIdx.x == 0) {
    const unsigned long long e = p.epoch[group];
    for (int r = 0; r < p.world; ++r) {
      if (r == p.rank) continue;
      const unsigned long long* f = p.sig[p.rank] + r * DP_GROUPS + group;
      long long spins = 0;
      while (ld_acquire_sys(f) < e + 1) {
        if (++spins > (1ll << 24)) { *p.error = 1; break; }      // tens of seconds: give up instead of hanging the GPU
        __nanosleep(64);
      }
    }
  }
  __syncthreads();
}
// (loads of up to 8 ranks are ISSUED together before the first sum: each is a round trip over NVLink)
__device__ __forceinline__ float4 dp_mean4(const DpPeer& p, long long base_i) {
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r0 = 0; r0 < p.world; r0 += 8) {
    float4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (r0 + i < p.world) v[i] = ld_volatile4(p.stage[r0 + i] + base_i);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (r0 + i < p.world) { s.x += v[i].x; s.y += v[i].y; s.z += v[i].z; s.w += v[i].w; }
  }
  const float w = (float)p.world;
  return make_float4(s.x / w, s.y / w, s.z / w, s.w / w);
}
__device__ __forceinline__ float ld_volatile1(const float* ptr) {
  float v;
  asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(ptr) : "memory");
  return v;
}
__device__ __forceinline__ float dp_mean1(const DpPeer& p, long long base_i) {
  float s = 0.f;
  for (int r0 = 0; r0 < p.world; r0 += 8) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (r0 + i < p.world) v[i] = ld_volatile1(p.stage[r0 + i] + base_i);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (r0 + i < p.world) s += v[i];
  }
  return s / (float)p.world;
}
// multi-block consumer: every block calls this at its end; the last one advances the epoch
__device__ __forceinline__ void dp_consume_done(const DpPeer& p, int group, unsigned int n_blocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(&p.ticket[DP_GROUPS + group], 1u);
    if (t == n_blocks - 1) {
      p.ticket[DP_GROUPS + group] = 0;
      p.epoch[group] = p.epoch[group] + 1;
    }
  }
}

// One kernel per exchange; grid <= number of SMs, so every block is resident while it waits for the peers.
// n is a multiple of 4; buffer / staging offsets are 16-byte aligned.  A block reduces exactly the elements it
// published itself, so there is no dependency between the blocks of one rank.
static __global__ void __launch_bounds__(256) k_dp_exchange(const DpPeer p, float* __restrict__ buf, long long off, long long n, int group) {
  __shared__ int ok;
  const unsigned long long e = p.epoch[group];
  const long long base = (long long)(e & 1) * p.stage_floats + off;
  const long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4, di = (long long)gridDim.x * blockDim.x * 4;
  // ---- publish: my gradients -> my staging buffer, then (last block) the new epoch into every peer's signal pad
  float* mine = p.stage[p.rank] + base;
  for (long long i = i0; i < n; i += di) *reinterpret_cast<float4*>(mine + i) = *reinterpret_cast<const float4*>(buf + i);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(&p.ticket[group], 1u);
    if (t == gridDim.x - 1) {                        // every block's copy is visible device-wide
      p.ticket[group] = 0;
      __threadfence_system();
      for (int r = 0; r < p.world; ++r)
        if (r != p.rank) st_release_sys(p.sig[r] + p.rank * DP_GROUPS + group, e + 1);
    }
    // ---- wait until every peer has published this epoch
    int good = 1;
    for (int r = 0; r < p.world && good; ++r) {
      if (r == p.rank) continue;
      const unsigned long long* f = p.sig[p.rank] + r * DP_GROUPS + group;
      long long spins = 0;
      while (ld_acquire_sys(f) < e + 1) {
        if (++spins > (1ll << 24)) { good = 0; break; }      // ~2 s: give up instead of hanging the GPU
        __nanosleep(64);
      }
    }
    if (!good) *p.error = 1;
    ok = good;
  }
  __syncthreads();
  // ---- one-shot reduce in rank order (P2P loads), mean, back into the local gradient buffer
  const float w = (float)p.world;
  for (long long i = i0; i < n; i += di) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < p.world; ++r) {
      const float4 v = ld_volatile4(p.stage[r] + base + i);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    if (ok) *reinterpret_cast<float4*>(buf + i) = make_float4(s.x / w, s.y / w, s.z / w, s.w / w);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(&p.ticket[DP_GROUPS + group], 1u);
    if (t == gridDim.x - 1) {
      p.ticket[DP_GROUPS + group] = 0;
      p.epoch[group] = e + 1;
    }
  }
}

}  // namespace cql
