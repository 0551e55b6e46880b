// mma_bench.cuh -- measurement aid: issue rate of tcgen05.mma kind::f16 (K = 16) for the operand placements and
// shapes the f16x3 kernels can choose from.  One CTA (cta_group::1) or one CTA pair (cta_group::2); one thread issues
// `iters` back-to-back MMAs into the same accumulator, commits, waits; clock64 around it.  Operand contents are
// irrelevant (whatever shared / tensor memory holds).  Used to decide between N = 128 (two accumulators + A in TMEM)
// and N = 256 (A in shared memory) -- DESIGN.md section 3.
#pragma once
#include "mlp_tc_h2.cuh"

namespace cql {
namespace tc {

// mode bits: 0 = A in TMEM (TS) else shared memory (SS); 1 = N 256 else 128; 2 = cta_group::2 (M = 256) else ::1 (M = 128)
// bit 3: issue the 3-term pattern (alternating two A / two B operand addresses) instead of one address pair
template <bool PAIR>
__device__ __forceinline__ void mma_bench_body(int mode, int iters, long long* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 64 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const bool ts = mode & 1, n256 = mode & 2, three = mode & 8;
  constexpr bool pair = PAIR;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank = 0;
  if constexpr (PAIR) rank = cluster_ctarank();
  for (int i = threadIdx.x; i < 16 * 1024; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;   // fp16 1.0
  fence_proxy_async_smem();
  if (warp == 0) {
    if constexpr (PAIR) tmem_alloc2(slot, 512); else tmem_alloc(slot, 512);
    if (lane == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 0 && rank == 0) {
    const int N = n256 ? 256 : 128;
    const uint32_t idesc = instr_desc(FMT_F16, pair ? 256 : 128, N);
    const int brows = pair ? N / 2 : N;                       // operand rows held by this CTA
    const uint64_t a0 = smem_desc(smem_u32(sm), 128 * 16, 128), a1 = smem_desc(smem_u32(sm) + 8192, 128 * 16, 128);
    const uint64_t b0 = smem_desc(smem_u32(sm) + 16384, brows * 16, 128), b1 = smem_desc(smem_u32(sm) + 32768, brows * 16, 128);
    const uint32_t ta0 = tmem + 256, ta1 = tmem + 264;
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const bool alt = three && (i % 3 == 0);
        const bool altb = three && (i % 3 == 1);
        if constexpr (PAIR) {
          if (ts) umma_ts2(tmem, alt ? ta1 : ta0, altb ? b1 : b0, idesc, 1u);
          else {
            const uint32_t z = 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(tmem),
                         "l"(alt ? a1 : a0), "l"(altb ? b1 : b0), "r"(idesc), "r"(1u), "r"(z)
                         : "memory");
          }
        } else {
          if (ts) umma_ts<false>(tmem, alt ? ta1 : ta0, altb ? b1 : b0, idesc, 1u);
          else umma<false>(tmem, alt ? a1 : a0, altb ? b1 : b0, idesc, 1u);
        }
      }
      if constexpr (PAIR) umma_commit2(bar, 1); else umma_commit(bar);
      t1 = clock64();
    }
    __syncwarp();
    mbar_wait(bar, 0);
    const long long t2 = clock64();
    if (elect_one()) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();
  if (warp == 0) { if constexpr (PAIR) tmem_dealloc2(tmem, 512); else tmem_dealloc(tmem, 512); }
}
__global__ void __launch_bounds__(128, 1) mma_bench_kernel(int mode, int iters, long long* __restrict__ out) {
  mma_bench_body<false>(mode, iters, out);
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) mma_bench_pair_kernel(int mode, int iters, long long* __restrict__ out) {
  mma_bench_body<true>(mode, iters, out);
}

}  // namespace tc
}  // namespace cql
