// tc_selftest.cuh -- smallest possible tcgen05 program: D[128 x N] = A[128 x K] * B[N x K]^T.
// Exercises exactly the operand layout, descriptors, TMEM allocation, commit/mbarrier and
// tcgen05.ld paths the fused kernels rely on; checked against a float64 product by the tests.
#pragma once
#include "tc_common.cuh"

namespace cql {
namespace tc {

template <bool TF32>
__device__ __forceinline__ void fill_operand(const float* __restrict__ src, int rows, int K, uint8_t* hi, uint8_t* lo) {
  constexpr int ES = TF32 ? 4 : 2, EPC = 16 / ES;
  const int chunks = rows * (K / EPC);
  for (int c = threadIdx.x; c < chunks; c += blockDim.x) {
    const int row = c % rows, kc = c / rows;
    const float* p = src + (size_t)row * K + kc * EPC;
    const uint32_t off = chunk_off(rows, row, kc);
    if (TF32) {
      float4 h, l;
      split_tf32(p[0], h.x, l.x); split_tf32(p[1], h.y, l.y); split_tf32(p[2], h.z, l.z); split_tf32(p[3], h.w, l.w);
      *reinterpret_cast<float4*>(hi + off) = h;
      *reinterpret_cast<float4*>(lo + off) = l;
    } else {
      __nv_bfloat162 v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = __floats2bfloat162_rn(p[2 * i], p[2 * i + 1]);
      *reinterpret_cast<uint4*>(hi + off) = *reinterpret_cast<uint4*>(v);
    }
  }
}

template <bool TF32>
__global__ void __launch_bounds__(128) umma_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                            float* __restrict__ D, int N, int K) {
  extern __shared__ __align__(1024) uint8_t sm[];
  constexpr int ES = TF32 ? 4 : 2, TERMS = TF32 ? 2 : 1, UK = 32 / ES;
  const uint32_t a_bytes = 128u * K * ES, b_bytes = (uint32_t)N * K * ES;
  uint8_t* As = sm;
  uint8_t* Bs = As + TERMS * a_bytes;
  uint64_t* bar = reinterpret_cast<uint64_t*>(Bs + TERMS * b_bytes);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t ncols = 32;
  while (ncols < (uint32_t)N) ncols <<= 1;

  fill_operand<TF32>(A, 128, K, As, As + a_bytes);
  fill_operand<TF32>(B, N, K, Bs, Bs + b_bytes);
  fence_proxy_async_smem();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
  if (warp == 0) tmem_alloc(slot, ncols);
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (tid == 0) {
    const uint32_t idesc = instr_desc(TF32 ? FMT_TF32 : FMT_BF16, 128, N);
    const uint32_t a_lbo = 128 * 16, b_lbo = N * 16;
    uint32_t acc = 0;
    for (int j = 0; j < K / UK; ++j) {
      const uint64_t a_hi = smem_desc(smem_u32(As) + 2 * j * a_lbo, a_lbo, 128);
      const uint64_t b_hi = smem_desc(smem_u32(Bs) + 2 * j * b_lbo, b_lbo, 128);
      if (TF32) {
        const uint64_t a_lo = smem_desc(smem_u32(As + a_bytes) + 2 * j * a_lbo, a_lbo, 128);
        const uint64_t b_lo = smem_desc(smem_u32(Bs + b_bytes) + 2 * j * b_lbo, b_lbo, 128);
        umma<TF32>(tmem, a_lo, b_hi, idesc, acc); acc = 1;
        umma<TF32>(tmem, a_hi, b_lo, idesc, acc);
        umma<TF32>(tmem, a_hi, b_hi, idesc, acc);
      } else {
        umma<TF32>(tmem, a_hi, b_hi, idesc, acc); acc = 1;
      }
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (c0 + i < N) D[(size_t)(warp * 32 + lane) * N + c0 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, ncols);
}

// Same product with the A operand staged in TENSOR MEMORY (tcgen05.st, thread = row) instead of shared memory --
// the layout the fused kernels use so that operand A costs no shared-memory bandwidth.
// TMEM columns: [0, KA) A hi, [KA, 2KA) A lo (tf32 only), then D; KA = K (tf32) or K/2 (bf16, two per column).
template <bool TF32>
__global__ void __launch_bounds__(128) umma_ts_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                               float* __restrict__ D, int N, int K) {
  extern __shared__ __align__(1024) uint8_t sm[];
  constexpr int ES = TF32 ? 4 : 2, TERMS = TF32 ? 2 : 1, UK = 32 / ES;
  const uint32_t b_bytes = (uint32_t)N * K * ES;
  uint8_t* Bs = sm;
  uint64_t* bar = reinterpret_cast<uint64_t*>(Bs + TERMS * b_bytes);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KA = TF32 ? K : K / 2;
  const int d_col = TERMS * KA;

  fill_operand<TF32>(B, N, K, Bs, Bs + b_bytes);
  fence_proxy_async_smem();
  if (warp == 0) tmem_alloc(slot, 512);
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  {  // each thread stages its own row of A
    const float* arow = A + (size_t)tid * K;
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < KA; c0 += 8) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (TF32) {
          float h, l;
          split_tf32(arow[c0 + i], h, l);
          hi[i] = __float_as_uint(h); lo[i] = __float_as_uint(l);
        } else {
          __nv_bfloat162 p = __floats2bfloat162_rn(arow[2 * (c0 + i)], arow[2 * (c0 + i) + 1]);
          hi[i] = *reinterpret_cast<uint32_t*>(&p); lo[i] = 0;
        }
      }
      tmem_st8(lane_base + c0, hi);
      if (TF32) tmem_st8(lane_base + KA + c0, lo);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) {
    const uint32_t idesc = instr_desc(TF32 ? FMT_TF32 : FMT_BF16, 128, N);
    const uint32_t b_lbo = N * 16;
    if (elect_one()) {
      uint32_t acc = 0;
      for (int j = 0; j < K / UK; ++j) {
        const uint64_t b_hi = smem_desc(smem_u32(Bs) + 2 * j * b_lbo, b_lbo, 128);
        const uint32_t a_hi = tmem + j * 8;               // 8 columns per MMA in both formats
        if (TF32) {
          const uint64_t b_lo = smem_desc(smem_u32(Bs + b_bytes) + 2 * j * b_lbo, b_lbo, 128);
          umma_ts<TF32>(tmem + d_col, a_hi + KA, b_hi, idesc, acc); acc = 1;
          umma_ts<TF32>(tmem + d_col, a_hi, b_lo, idesc, acc);
          umma_ts<TF32>(tmem + d_col, a_hi, b_hi, idesc, acc);
        } else {
          umma_ts<TF32>(tmem + d_col, a_hi, b_hi, idesc, acc); acc = 1;
        }
      }
      umma_commit(bar);
    }
    __syncwarp();
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + d_col + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (c0 + i < N) D[(size_t)(warp * 32 + lane) * N + c0 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

inline size_t selftest_smem(bool tf32, int N, int K) {
  const size_t es = tf32 ? 4 : 2, terms = tf32 ? 2 : 1;
  return terms * (128 + (size_t)N) * K * es + 64;
}

}  // namespace tc
}  // namespace cql
