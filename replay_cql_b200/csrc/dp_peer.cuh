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

// ---- fused mode (f16x3 path): no exchange launch at all, and no signal or fence either for the two big groups.
// PUSH with the flag inside the data ("LL" packets): values travel as 8-byte {value, epoch tag} pairs (two pairs per
// 16-byte store; NVLink keeps an aligned 8-byte store whole) written straight into a peer's staging buffer; whoever
// needs a value polls its OWN staging buffer -- local memory -- until the pair carries this epoch's tag.  The schedule
// is two-shot (see DpSlices below): k_reduce_grads_tc pushes slices to their owners, k_adam_pack's owner phase sums in
// RANK ORDER and pushes the mean, every consumer thread polls the means it needs -- every rank applies the owner's bits.
// History (us per update; independent replicas 217.5): publish + fence + signal, consumers read the peers' buffers over
// NVLink: 251 at 2 GPUs (peer reads 11-15, signal + fences + store acknowledgement ~20; waiting itself 0: the ranks run
// in lock step); one fence per signal round, relaxed polls, batched peer loads: 248; one-shot push: 230.5 at 2 GPUs but
// 270 at 8 (sends +28, polls +24); two-shot push: 231 / 240 (profiles/r02_dp_decomposition.txt).
// Slot [parity][source rank] is rewritten two epochs later; a rank gets there only after consuming the next epoch,
// whose packets a peer sends only after its own consumer of THIS epoch has ended -- no second barrier.  The consumer's
// last block advances the epoch (dp_consume_done).
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
// Position (in 16-byte packets) of the two packets of the 16-byte value group q (= float index / 4) inside a source
// rank's slot: groups are tiled by 32 -- [tile][half][lane] -- so that a warp of the producer, whose lanes hold 32
// consecutive groups, writes 512 contiguous bytes with each of its two stores: full 128-byte lines on NVLink instead
// of 16 bytes out of every 32.
__device__ __forceinline__ long long dp_ll_pkt(long long q, int half) { return (q >> 5) * 64 + half * 32 + (q & 31); }
// ---- two-shot schedule of a big group (critics: 134 k floats, actor: 67 k).  One-shot (every rank pushes everything to
// every peer) moves 7 x 2 x 0.8 MB per rank and update at 8 GPUs: measured +28 us for the sends and +24 us for the
// consumers' polls (profiles/r02_dp_decomposition.txt).  Here the group's 16-byte value groups are cut into `world`
// contiguous slices; slice o is OWNED by rank o:
//   A  (k_reduce_grads_tc)  a rank pushes its sums of slice o to rank o only            (slot [source = me] at rank o)
//   B  (k_adam_pack, first) the owner polls the world - 1 contributions of its slice, adds its own IN RANK ORDER, divides
//                           and pushes the mean to every peer and to itself             (slot [source = owner] everywhere)
//   C  (k_adam_pack)        every consumer polls the owner's packets of the values it needs.
// 2 x (world - 1) / world of the group crosses NVLink per rank instead of world - 1 times the group, every rank applies
// the owner's bits (replicas bit-identical; at world 2 the mean is still (g0 + g1) / 2 = the NCCL result), and phase B
// depends on phase A packets only -- sent by the previous kernel of every rank -- so no rank's B waits for another
// rank's B or C.  Contributions to owner o and results from owner o use the same slot space: disjoint value ranges.
struct DpSlices {
  long long q_lo;      // first 16-byte value group of the gradient group (absolute index in the gradient layout / 4)
  long long n_q;       // value groups in the gradient group
  long long per;       // value groups per slice (even: an 8-float chunk never straddles two owners)
};
__device__ __forceinline__ DpSlices dp_slices(const DpPeer& p, long long off_floats, long long n_floats) {
  DpSlices sl;
  sl.q_lo = off_floats >> 2;
  sl.n_q = n_floats >> 2;
  sl.per = ((sl.n_q + p.world - 1) / p.world + 1) & ~1ll;
  return sl;
}
__device__ __forceinline__ int dp_owner(const DpSlices& sl, long long q_abs) { return (int)((q_abs - sl.q_lo) / sl.per); }
__device__ __forceinline__ void dp_ll_store(uint4* slot, long long q_abs, float4 v, unsigned int tag) {
  asm volatile("st.relaxed.sys.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(slot + dp_ll_pkt(q_abs, 0)), "r"(__float_as_uint(v.x)), "r"(tag),
               "r"(__float_as_uint(v.y)), "r"(tag) : "memory");
  asm volatile("st.relaxed.sys.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(slot + dp_ll_pkt(q_abs, 1)), "r"(__float_as_uint(v.z)), "r"(tag),
               "r"(__float_as_uint(v.w)), "r"(tag) : "memory");
}
// both packets of value group q_abs; polls until they carry `tag` (timeout -> error flag); `first` = already requested
__device__ __forceinline__ float4 dp_ll_wait(const DpPeer& p, const uint4* slot, long long q_abs, unsigned int tag) {
  uint4 a, b;
  long long spins = 0;
  for (;;) {
    asm volatile("ld.relaxed.sys.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(slot + dp_ll_pkt(q_abs, 0)) : "memory");
    asm volatile("ld.relaxed.sys.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(slot + dp_ll_pkt(q_abs, 1)) : "memory");
    if (a.y == tag && a.w == tag && b.y == tag && b.w == tag) break;
    if (++spins > (1ll << 22)) { *p.error = 1; break; }         // seconds: give up instead of hanging the GPU
    __nanosleep(20);
  }
  return make_float4(__uint_as_float(a.x), __uint_as_float(a.z), __uint_as_float(b.x), __uint_as_float(b.z));
}
// phase A -- producer: the sums of value group q (absolute float index i = 4 q) go to the slice's owner
__device__ __forceinline__ void dp_ll_push4(const DpPeer& p, int group, const DpSlices& sl, long long i, float4 v) {
  if (p.debug & 8) return;
  const int o = dp_owner(sl, i >> 2);
  if (o == p.rank) return;                                   // the owner takes its own sums from its gradient buffer
  dp_ll_store(reinterpret_cast<uint4*>(dp_ll_slot(p, o, p.rank, group)), i >> 2, v, dp_ll_tag(p, group));
}
// phase B -- owner: thread `gid` of `n_threads` reduces value groups gid, gid + n_threads, ... of this rank's slice.
// `own` = this rank's gradient buffer at the group's first float.
__device__ __forceinline__ void dp_ll_owner_reduce(const DpPeer& p, int group, const DpSlices& sl, const float* own,
                                                   long long gid, long long n_threads) {
  const unsigned int tag = dp_ll_tag(p, group);
  const long long lo = sl.q_lo + (long long)p.rank * sl.per;
  const long long hi = min(sl.q_lo + sl.n_q, lo + sl.per);
  for (long long q = lo + gid; q < hi; q += n_threads) {
    const float4 mine = *reinterpret_cast<const float4*>(own + ((q - sl.q_lo) << 2));
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r0 = 0; r0 < p.world; r0 += 8) {                // (all packets of up to 8 ranks requested before the first check)
      uint4 a[8], b[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (r0 + j < p.world && r0 + j != p.rank && !(p.debug & 4)) {
          const uint4* slot = reinterpret_cast<const uint4*>(dp_ll_slot(p, p.rank, r0 + j, group));
          asm volatile("ld.relaxed.sys.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a[j].x), "=r"(a[j].y), "=r"(a[j].z), "=r"(a[j].w) : "l"(slot + dp_ll_pkt(q, 0)));
          asm volatile("ld.relaxed.sys.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(b[j].x), "=r"(b[j].y), "=r"(b[j].z), "=r"(b[j].w) : "l"(slot + dp_ll_pkt(q, 1)));
        }
#pragma unroll
      for (int j = 0; j < 8; ++j) {                          // rank order
        if (r0 + j >= p.world) continue;
        float4 v = mine;
        if (r0 + j != p.rank && !(p.debug & 4)) {
          if (a[j].y == tag && a[j].w == tag && b[j].y == tag && b[j].w == tag)
            v = make_float4(__uint_as_float(a[j].x), __uint_as_float(a[j].z), __uint_as_float(b[j].x), __uint_as_float(b[j].z));
          else
            v = dp_ll_wait(p, reinterpret_cast<const uint4*>(dp_ll_slot(p, p.rank, r0 + j, group)), q, tag);
        }
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
    }
    const float n = (float)p.world;
    const float4 m = make_float4(s.x / n, s.y / n, s.z / n, s.w / n);
    for (int r = 0; r < p.world; ++r)
      if (r == p.rank || !(p.debug & 8)) dp_ll_store(reinterpret_cast<uint4*>(dp_ll_slot(p, r, p.rank, group)), q, m, tag);
  }
}
// phase C -- consumer: means of EIGHT consecutive values (float index i .. i + 7) from their owner's packets
__device__ __forceinline__ void dp_ll_mean8(const DpPeer& p, int group, const DpSlices& sl, long long i, float4& m0, float4& m1) {
  const unsigned int tag = dp_ll_tag(p, group);
  const int o = dp_owner(sl, i >> 2);
  if ((p.debug & 8) && o != p.rank) return;                  // (timing experiments: nothing was sent; keep the own values)
  const uint4* slot = reinterpret_cast<const uint4*>(dp_ll_slot(p, p.rank, o, group));
  m0 = dp_ll_wait(p, slot, i >> 2, tag);
  m1 = dp_ll_wait(p, slot, (i >> 2) + 1, tag);
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
