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
  int debug = 0;                                   // CQL_DP_DEBUG (timing experiments only, results wrong): 1 = no wait for
                                                   // scalar group, 4 = own values instead of the peers' packets (no polling), 8 = nothing sent
};

__device__ __forceinline__ float4 ld_volatile4(const float* p) {   // bypasses L1: peer data changes between launches
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// ---- fused mode (f16x3 path): no exchange launch at all, and no signal either for the two big groups.
// PUSH with the flag inside the data ("LL" packets): the kernel that PRODUCES a gradient group (k_reduce_grads_tc)
// stores every value it has summed straight into EVERY peer's staging buffer as 8-byte {value, epoch tag} pairs (two
// pairs per 16-byte store; NVLink keeps an aligned 8-byte store whole); the kernel that CONSUMES the group (k_adam_pack)
// polls its OWN staging buffer -- local memory -- until the tags of the pairs it needs carry this epoch, takes its own
// contribution from its gradient buffer, and sums in RANK ORDER, so every rank applies bit-identical gradients.  One
// NVLink one-way latency per exchange.  History (2 GPUs, us per update; 1 GPU = 213): publish + fence + signal, then the
// consumer reads the peers' buffers over NVLink: 251 (of which the peer reads 15, signal + fences + store
// acknowledgement ~20; waiting itself 0: the ranks run in lock step); one fence per signal round and relaxed polling,
// peer loads batched: 248; LL push: see profiles/.  Staging[parity][source rank] is rewritten two epochs later; a rank
// gets there only after consuming the next epoch, whose packets a peer sends only after its own consumer of THIS epoch
// has ended -- no second barrier.  The consumer's last block advances the epoch (dp_consume_done).
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// LL region of a staging buffer: after the pull region [2][stage_floats]; [parity][source rank][stage_floats] PAIRS
__device__ __forceinline__ uint2* dp_ll_slot(const DpPeer& p, int dst_rank, int src_rank, int group) {
  return reinterpret_cast<uint2*>(p.stage[dst_rank] + 2 * p.stage_floats) +
         ((long long)(p.epoch[group] & 1ull) * p.world + src_rank) * p.stage_floats;
}
__device__ __forceinline__ unsigned int dp_ll_tag(const DpPeer& p, int group) { return (unsigned int)(p.epoch[group] + 1); }
// producer: four consecutive values -> every peer (index i counts floats inside the gradient layout)
__device__ __forceinline__ void dp_ll_push4(const DpPeer& p, int group, long long i, float4 v) {
  if (p.debug & 8) return;
  const unsigned int tag = dp_ll_tag(p, group);
  for (int r = 0; r < p.world; ++r) {
    if (r == p.rank) continue;
    uint2* dst = dp_ll_slot(p, r, p.rank, group) + i;
    asm volatile("st.relaxed.sys.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(__float_as_uint(v.x)), "r"(tag),
                 "r"(__float_as_uint(v.y)), "r"(tag) : "memory");
    asm volatile("st.relaxed.sys.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 2), "r"(__float_as_uint(v.z)), "r"(tag),
                 "r"(__float_as_uint(v.w)), "r"(tag) : "memory");
  }
}
// consumer: mean over ranks of EIGHT consecutive values (i .. i + 7); `own` = this rank's values.  All packets of all
// peers are requested before the first tag is checked; stale ones are polled again (timeout -> error flag).
__device__ __forceinline__ void dp_ll_mean8(const DpPeer& p, int group, long long i, const float4& own0, const float4& own1,
                                            float4& m0, float4& m1) {
  const unsigned int tag = dp_ll_tag(p, group);
  float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
  for (int r0 = 0; r0 < p.world; r0 += 8) {
    uint4 q[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = r0 + j;
      if (r < p.world && r != p.rank && !(p.debug & 4)) {
        const uint4* src = reinterpret_cast<const uint4*>(dp_ll_slot(p, p.rank, r, group) + i);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          asm volatile("ld.relaxed.sys.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q[j][c].x), "=r"(q[j][c].y), "=r"(q[j][c].z), "=r"(q[j][c].w) : "l"(src + c));
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = r0 + j;
      if (r >= p.world) continue;
      if (r == p.rank || (p.debug & 4)) {
        s0.x += own0.x; s0.y += own0.y; s0.z += own0.z; s0.w += own0.w;
        s1.x += own1.x; s1.y += own1.y; s1.z += own1.z; s1.w += own1.w;
        continue;
      }
      const uint4* src = reinterpret_cast<const uint4*>(dp_ll_slot(p, p.rank, r, group) + i);
      long long spins = 0;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        while (q[j][c].y != tag || q[j][c].w != tag) {
          if (++spins > (1ll << 22)) { *p.error = 1; break; }     // seconds: give up instead of hanging the GPU
          __nanosleep(20);
          asm volatile("ld.relaxed.sys.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q[j][c].x), "=r"(q[j][c].y), "=r"(q[j][c].z), "=r"(q[j][c].w) : "l"(src + c) : "memory");
        }
      }
      s0.x += __uint_as_float(q[j][0].x); s0.y += __uint_as_float(q[j][0].z);
      s0.z += __uint_as_float(q[j][1].x); s0.w += __uint_as_float(q[j][1].z);
      s1.x += __uint_as_float(q[j][2].x); s1.y += __uint_as_float(q[j][2].z);
      s1.z += __uint_as_float(q[j][3].x); s1.w += __uint_as_float(q[j][3].z);
    }
  }
  const float n = (float)p.world;
  m0 = make_float4(s0.x / n, s0.y / n, s0.z / n, s0.w / n);
  m1 = make_float4(s1.x / n, s1.y / n, s1.z / n, s1.w / n);
}

// ---- the scalar group (two floats: d loss / d log_temp, d loss / d log_alpha) travels INSIDE the signal words: word of
// group 0 = (epoch + 1) << 32 | bits(g_temp), word of group 3 = (epoch + 1) << 32 | bits(g_alpha).  No staging write, no
// fence (the data is the flag), no peer read: the consumer polls its own pad.
constexpr int DP_SCALAR_SLOT2 = 3;
__device__ __forceinline__ void dp_signal_scalars(const DpPeer& p, float g_t, float g_a) {
  if (p.debug & 8) return;
  const unsigned long long tag = (p.epoch[0] + 1) << 32;
  for (int r = 0; r < p.world; ++r)
    if (r != p.rank) {
      st_relaxed_sys(p.sig[r] + p.rank * DP_GROUPS + 0, tag | (unsigned long long)__float_as_uint(g_t));
      st_relaxed_sys(p.sig[r] + p.rank * DP_GROUPS + DP_SCALAR_SLOT2, tag | (unsigned long long)__float_as_uint(g_a));
    }
}
// ONE thread: waits for every peer's two words of this epoch and returns the means (rank order; own values passed in)
__device__ __forceinline__ void dp_wait_scalars(const DpPeer& p, float own_t, float own_a, float& mean_t, float& mean_a) {
  const unsigned long long want = (p.epoch[0] + 1) & 0xffffffffull;
  float st = 0.f, sa = 0.f;
  for (int r = 0; r < p.world; ++r) {
    float gt = own_t, ga = own_a;
    if (r != p.rank && !(p.debug & 1)) {
      const unsigned long long* f = p.sig[p.rank] + r * DP_GROUPS;
      unsigned long long w0 = 0, w1 = 0;
      long long spins = 0;
      for (;;) {
        w0 = ld_relaxed_sys(f + 0);
        w1 = ld_relaxed_sys(f + DP_SCALAR_SLOT2);
        if ((w0 >> 32) >= want && (w1 >> 32) >= want) break;
        if (++spins > (1ll << 24)) { *p.error = 1; break; }
        __nanosleep(32);
      }
      gt = __uint_as_float((unsigned int)(w0 & 0xffffffffull));
      ga = __uint_as_float((unsigned int)(w1 & 0xffffffffull));
    }
    st += gt;
    sa += ga;
  }
  mean_t = st / (float)p.world;
  mean_a = sa / (float)p.world;
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
      __threadfence_system();                        // one release fence, then relaxed stores issued back to back
      for (int r = 0; r < p.world; ++r)
        if (r != p.rank) st_relaxed_sys(p.sig[r] + p.rank * DP_GROUPS + group, e + 1);
    }
    // ---- wait until every peer has published this epoch
    int good = 1;
    for (int r = 0; r < p.world && good; ++r) {
      if (r == p.rank) continue;
      const unsigned long long* f = p.sig[p.rank] + r * DP_GROUPS + group;
      long long spins = 0;
      while (ld_relaxed_sys(f) < e + 1) {
        if (++spins > (1ll << 24)) { good = 0; break; }      // ~2 s: give up instead of hanging the GPU
        __nanosleep(32);
      }
    }
    __threadfence_system();                          // acquire side
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
