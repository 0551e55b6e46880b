// mlp_simt.cuh -- FP32 CUDA-core kernels for the 3-layer MLPs (in -> 256 -> 256 -> out).
//
// This is the exact-order FP32 path (CQL_PREC_FP32): every contraction is an
// FP32 FMA chain, so it tracks the CPU oracle to rounding.  Three kernels carry
// all the FLOPs of a CQL update (DESIGN.md "kernels"):
//   fwd   : out = W3 relu(W2 relu(W1 x + b1) + b2) + b3         (rows x 256 x 256)
//   bwd1  : dZ1 = (dZ2 W2) * relu'(Z1) ; small grads ; optional dx (rows x 256 x 256)
//   bwd2  : dW2 = dZ2^T H1   (256 x 256 x rows, split over row tiles)
// Row tiles are 64 rows.  H1 is never stored (3 FMAs to recompute); H2 is stored
// by fwd in a tile-blocked transposed layout h2[tile][j][r] so that both
// backward kernels read it coalesced *and* write shared memory conflict-free.
#pragma once
#include "common.cuh"

namespace cql {

constexpr int BM = 64;    // rows per tile
constexpr int BK = 16;    // reduction chunk
constexpr int NT = 256;   // threads per CTA

struct FwdJob {
  const float4* X;      // [rows] (x0, x1, x2, -)
  const float* params;  // first network slot
  float* out;           // [n_nets][rows][OUT]
  float* h2;            // [n_nets][tiles][256][64] or nullptr
  int rows, n_nets, tile_begin;
};
struct FwdJobs {
  FwdJob j[4];
  int n, total_tiles;
};

__host__ __device__ inline int tiles_of(int rows) { return (rows + BM - 1) / BM; }

constexpr int FWD_SMEM = (H * BM + 2 * BK * H) * 4 + BM * 16;  // As + Bs[2] + Xs

// ---- shared building blocks -------------------------------------------------
// H1^T tile: As[k*64 + r] = relu(W1[k].x[r] + b1[k]); warp w owns k in [32w,32w+32), lanes own rows.
template <int IN>
__device__ __forceinline__ void build_h1_tile(const float* __restrict__ net, const float4* Xs, float* As,
                                              int k_begin, int k_count, int stride) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* W1 = net + off_W1(IN);
  const float* b1 = net + off_b1(IN);
  const float4 xa = Xs[lane], xb = Xs[lane + 32];
  const int per_warp = k_count / (NT / 32);
  for (int kk = 0; kk < per_warp; ++kk) {
    const int kl = warp * per_warp + kk;
    const int k = k_begin + kl;
    const float w0 = __ldg(W1 + k * IN), w1 = __ldg(W1 + k * IN + 1);
    const float w2 = IN == 3 ? __ldg(W1 + k * IN + 2) : 0.f;
    const float bb = __ldg(b1 + k);
    float za = fmaf(xa.y, w1, xa.x * w0), zb = fmaf(xb.y, w1, xb.x * w0);
    if (IN == 3) { za = fmaf(xa.z, w2, za); zb = fmaf(xb.z, w2, zb); }
    As[kl * stride + lane] = fmaxf(za + bb, 0.f);
    As[kl * stride + lane + 32] = fmaxf(zb + bb, 0.f);
  }
}

__device__ __forceinline__ void fma_8x8(float (&acc)[8][8], const float4& a0, const float4& a1,
                                        const float4& b0, const float4& b1) {
  const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
  const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
}

// acc[r][c] (+)= sum_k As[k][rows of ty] * Bsrc[k][cols of tx], Bsrc streamed from global.
// TRANSPOSED_B: global is [n][k] (W2 as stored, reduction index minor)  -> forward
// otherwise   : global is [k][n] (W2 rows are the reduction index)      -> backward dH1
template <bool TRANSPOSED_B>
__device__ __forceinline__ void gemm_tile_64x256(const float* __restrict__ Bg, const float* As, float* Bs,
                                                 float (&acc)[8][8]) {
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  float4 pf[4];
  auto load_chunk = [&](int c) {
    if (TRANSPOSED_B) {
      const float4* src = reinterpret_cast<const float4*>(Bg + (size_t)tid * H + c * BK);
#pragma unroll
      for (int i = 0; i < 4; ++i) pf[i] = __ldg(src + i);
    } else {
      const float4* src = reinterpret_cast<const float4*>(Bg + (size_t)c * BK * H);
#pragma unroll
      for (int i = 0; i < 4; ++i) pf[i] = __ldg(src + tid + i * NT);
    }
  };
  auto store_chunk = [&](int buf) {
    float* dst = Bs + buf * BK * H;
    if (TRANSPOSED_B) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        dst[(i * 4 + 0) * H + tid] = pf[i].x;
        dst[(i * 4 + 1) * H + tid] = pf[i].y;
        dst[(i * 4 + 2) * H + tid] = pf[i].z;
        dst[(i * 4 + 3) * H + tid] = pf[i].w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(dst)[tid + i * NT] = pf[i];
    }
  };
  load_chunk(0);
  store_chunk(0);
  __syncthreads();
  constexpr int NCHUNK = H / BK;
  for (int c = 0; c < NCHUNK; ++c) {
    if (c + 1 < NCHUNK) load_chunk(c + 1);
    const float* Bc = Bs + (c & 1) * BK * H;
    const float* Ac = As + c * BK * BM;
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(Ac + kk * BM + ty * 8);
      const float4 a1 = *reinterpret_cast<const float4*>(Ac + kk * BM + ty * 8 + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(Bc + kk * H + tx * 4);
      const float4 b1 = *reinterpret_cast<const float4*>(Bc + kk * H + 128 + tx * 4);
      fma_8x8(acc, a0, a1, b0, b1);
    }
    if (c + 1 < NCHUNK) store_chunk((c + 1) & 1);
    __syncthreads();
  }
}

__device__ __forceinline__ int col_of(int tx, int j) { return (j < 4 ? 0 : 128) + tx * 4 + (j & 3); }

// One 64-row tile through layers 1-3.  Xs: 64 float4 in smem (already synced).
// Returns per-row outputs in outs[r*OUT+o] (smem, valid after the trailing sync).
template <int IN, int OUT>
__device__ __forceinline__ void fwd_tile(const float* __restrict__ net, const float4* Xs, float* As, float* Bs,
                                         float* outs, float* h2_tile /*global [256][64] or null*/) {
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  build_h1_tile<IN>(net, Xs, As, 0, H, BM);
  __syncthreads();
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  gemm_tile_64x256<true>(net + off_W2(IN), As, Bs, acc);
  // epilogue: bias + relu, layer 3, optional H2 store
  const float* b2 = net + off_b2(IN);
  const float* W3 = net + off_W3(IN);
  float bias[8], w3[OUT][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = col_of(tx, j);
    bias[j] = __ldg(b2 + c);
#pragma unroll
    for (int o = 0; o < OUT; ++o) w3[o][j] = __ldg(W3 + o * H + c);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = fmaxf(acc[i][j] + bias[j], 0.f);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int o = 0; o < OUT; ++o) {
      float p = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) p = fmaf(acc[i][j], w3[o][j], p);
      p = warp_sum(p);
      if (tx == 0) outs[(ty * 8 + i) * OUT + o] = p + __ldg(net + off_b3(IN, OUT) + o);
    }
  }
  if (h2_tile != nullptr) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float* dst = h2_tile + col_of(tx, j) * BM + ty * 8;
      *reinterpret_cast<float4*>(dst) = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[4][j], acc[5][j], acc[6][j], acc[7][j]);
    }
  }
  __syncthreads();
}

// ---- forward kernel -----------------------------------------------------------
template <int IN, int OUT>
__global__ void __launch_bounds__(NT, 2) mlp_fwd_kernel(const FwdJobs jobs) {
  extern __shared__ __align__(16) float smem[];
  float* As = smem;
  float* Bs = As + H * BM;
  float4* Xs = reinterpret_cast<float4*>(Bs + 2 * BK * H);
  __shared__ float outs[BM * OUT];
  int ji = 0;
  while (ji + 1 < jobs.n && (int)blockIdx.x >= jobs.j[ji + 1].tile_begin) ++ji;
  const FwdJob& jb = jobs.j[ji];
  const int tiles = tiles_of(jb.rows);
  const int local = blockIdx.x - jb.tile_begin;
  const int net_i = local / tiles, tile = local % tiles;
  const float* net = jb.params + (size_t)net_i * NET_STRIDE;
  const int row0 = tile * BM;
  if (threadIdx.x < BM) {
    const int r = row0 + threadIdx.x;
    Xs[threadIdx.x] = r < jb.rows ? jb.X[r] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  float* h2_tile = jb.h2 ? jb.h2 + ((size_t)net_i * tiles + tile) * H * BM : nullptr;
  fwd_tile<IN, OUT>(net, Xs, As, Bs, outs, h2_tile);
  if (threadIdx.x < BM * OUT) {
    const int r = row0 + threadIdx.x / OUT;
    if (r < jb.rows) jb.out[((size_t)net_i * jb.rows + r) * OUT + threadIdx.x % OUT] = outs[threadIdx.x];
  }
}

// ---- backward 1: dZ2 -> dZ1, small grads, dx ---------------------------------------
struct BwdJob {
  const float4* X;      // [rows]
  const float* dOut;    // [n_nets][rows][OUT]
  const float* h2;      // [n_nets][tiles][256][64]
  const float* params;  // first network slot
  float* small;         // [n_nets][tiles][SMALL_STRIDE] partial grads (W1|b1|b2|W3|b3) or nullptr
  float4* dX;           // [n_nets][rows] or nullptr
  float* pw2;           // [n_nets][splits][H*H] partial dW2 (bwd2)
  int rows, n_nets, splits;
  // f16x3 dx-only launch of the actor step: dOut is taken from the forward's partial sums (tc::Bwd1Job::q_part)
  const float* q_part = nullptr;
  int q_parts = 0;
  float dq_scale = 0.f;
};

// dZ2^T tile into As[jl*stride + r] for j in [j_begin, j_begin+j_count); optional small grads.
template <int OUT, bool WGRADS>
__device__ __forceinline__ void build_dz2_tile(const float* __restrict__ net_W3, const float* __restrict__ h2_tile,
                                               const float* dOs, float* As, int j_begin, int j_count, int stride,
                                               float* sm_db2, float* sm_dW3) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float da[OUT], db[OUT];
#pragma unroll
  for (int o = 0; o < OUT; ++o) { da[o] = dOs[lane * OUT + o]; db[o] = dOs[(lane + 32) * OUT + o]; }
  const int per_warp = j_count / (NT / 32);
  for (int jj = 0; jj < per_warp; ++jj) {
    const int jl = warp * per_warp + jj, j = j_begin + jl;
    const float ha = __ldg(h2_tile + j * BM + lane), hb = __ldg(h2_tile + j * BM + lane + 32);
    float ga = 0.f, gb = 0.f;
#pragma unroll
    for (int o = 0; o < OUT; ++o) {
      const float w = __ldg(net_W3 + o * H + j);
      ga = fmaf(da[o], w, ga);
      gb = fmaf(db[o], w, gb);
    }
    const float za = ha > 0.f ? ga : 0.f, zb = hb > 0.f ? gb : 0.f;
    As[jl * stride + lane] = za;
    As[jl * stride + lane + 32] = zb;
    if (WGRADS) {
      const float s = warp_sum(za + zb);
      if (lane == 0) sm_db2[j] = s;
#pragma unroll
      for (int o = 0; o < OUT; ++o) {
        const float t = warp_sum(fmaf(da[o], ha, db[o] * hb));
        if (lane == 0) sm_dW3[o * H + j] = t;
      }
    }
  }
}

template <int IN, int OUT, bool WGRADS, bool DX>
__global__ void __launch_bounds__(NT, 2) mlp_bwd1_kernel(const BwdJob jb) {
  extern __shared__ __align__(16) float smem[];
  float* As = smem;
  float* Bs = As + H * BM;
  float4* Xs = reinterpret_cast<float4*>(Bs + 2 * BK * H);
  __shared__ float dOs[BM * OUT];
  __shared__ float sm_db2[H];
  __shared__ float sm_dW3[OUT * H];
  const int tiles = tiles_of(jb.rows);
  const int net_i = blockIdx.x / tiles, tile = blockIdx.x % tiles;
  const float* net = jb.params + (size_t)net_i * NET_STRIDE;
  const int row0 = tile * BM, tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  if (tid < BM) {
    const int r = row0 + tid;
    Xs[tid] = r < jb.rows ? jb.X[r] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (tid < BM * OUT) {
    const int r = row0 + tid / OUT;
    dOs[tid] = r < jb.rows ? jb.dOut[((size_t)net_i * jb.rows + r) * OUT + tid % OUT] : 0.f;
  }
  __syncthreads();
  const float* h2_tile = jb.h2 + ((size_t)net_i * tiles + tile) * H * BM;
  build_dz2_tile<OUT, WGRADS>(net + off_W3(IN), h2_tile, dOs, As, 0, H, BM, sm_db2, sm_dW3);
  __syncthreads();
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  gemm_tile_64x256<false>(net + off_W2(IN), As, Bs, acc);
  // epilogue: mask with relu'(Z1), small grads, dx.  As is free now -> scratch.
  const float* W1 = net + off_W1(IN);
  const float* b1 = net + off_b1(IN);
  float4 xr[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) xr[i] = Xs[ty * 8 + i];
  float dxa[8][IN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int c = 0; c < IN; ++c) dxa[i][c] = 0.f;
  float* scr = As;  // [ty][k][4]
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = col_of(tx, j);
    const float w0 = __ldg(W1 + k * IN), w1 = __ldg(W1 + k * IN + 1);
    const float w2 = IN == 3 ? __ldg(W1 + k * IN + 2) : 0.f;
    const float bb = __ldg(b1 + k);
    float s_b = 0.f, s_w0 = 0.f, s_w1 = 0.f, s_w2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float z = fmaf(xr[i].y, w1, xr[i].x * w0);
      if (IN == 3) z = fmaf(xr[i].z, w2, z);
      z += bb;
      const float d = z > 0.f ? acc[i][j] : 0.f;
      s_b += d;
      s_w0 = fmaf(d, xr[i].x, s_w0);
      s_w1 = fmaf(d, xr[i].y, s_w1);
      if (IN == 3) s_w2 = fmaf(d, xr[i].z, s_w2);
      if (DX) {
        dxa[i][0] = fmaf(d, w0, dxa[i][0]);
        dxa[i][1] = fmaf(d, w1, dxa[i][1]);
        if (IN == 3) dxa[i][2] = fmaf(d, w2, dxa[i][2]);
      }
    }
    if (WGRADS) *reinterpret_cast<float4*>(scr + ((size_t)ty * H + k) * 4) = make_float4(s_b, s_w0, s_w1, s_w2);
  }
  if (DX) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < IN; ++c) v[c] = warp_sum(dxa[i][c]);
      const int r = row0 + ty * 8 + i;
      if (tx == 0 && r < jb.rows) jb.dX[(size_t)net_i * jb.rows + r] = make_float4(v[0], v[1], v[2], 0.f);
    }
  }
  if (WGRADS) {
    __syncthreads();
    float* out = jb.small + ((size_t)net_i * tiles + tile) * SMALL_STRIDE;
    {  // thread k: sum the 8 row-groups in fixed order
      const int k = tid;
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int g = 0; g < NT / 32; ++g) {
        const float4 v = *reinterpret_cast<const float4*>(scr + ((size_t)g * H + k) * 4);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
      out[k * IN + 0] = s.y;
      out[k * IN + 1] = s.z;
      if (IN == 3) out[k * IN + 2] = s.w;
      out[H * IN + k] = s.x;                 // b1
      out[H * IN + H + k] = sm_db2[k];       // b2
#pragma unroll
      for (int o = 0; o < OUT; ++o) out[H * IN + 2 * H + o * H + k] = sm_dW3[o * H + k];
    }
    if (tid < OUT) {  // b3
      float s = 0.f;
      for (int r = 0; r < BM; ++r) s += dOs[r * OUT + tid];
      out[H * IN + 2 * H + OUT * H + tid] = s;
    }
  }
}
constexpr int BWD1_SMEM = FWD_SMEM;

// ---- backward 2: dW2 = dZ2^T H1 --------------------------------------------------------
// grid (4 output blocks of 128x128, splits, n_nets); each CTA walks tiles split, split+S, ...
constexpr int B2S = 68;  // padded row stride (floats) of the [128][64] operand tiles
constexpr int BWD2_SMEM = 2 * 128 * B2S * 4 + BM * 16;

template <int IN, int OUT>
__global__ void __launch_bounds__(NT, 2) mlp_bwd2_kernel(const BwdJob jb) {
  extern __shared__ __align__(16) float smem[];
  float* As = smem;               // dZ2^T block [128 j][68]
  float* Hs = As + 128 * B2S;     // H1^T  block [128 k][68]
  float4* Xs = reinterpret_cast<float4*>(Hs + 128 * B2S);
  __shared__ float dOs[BM * OUT];
  const int jb_i = blockIdx.x >> 1, kb_i = blockIdx.x & 1;
  const int split = blockIdx.y, net_i = blockIdx.z;
  const float* net = jb.params + (size_t)net_i * NET_STRIDE;
  const int tiles = tiles_of(jb.rows);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = lane & 15, ty = warp * 2 + (lane >> 4);
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  for (int tile = split; tile < tiles; tile += jb.splits) {
    const int row0 = tile * BM;
    if (tid < BM) {
      const int r = row0 + tid;
      Xs[tid] = r < jb.rows ? jb.X[r] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (tid < BM * OUT) {
      const int r = row0 + tid / OUT;
      dOs[tid] = r < jb.rows ? jb.dOut[((size_t)net_i * jb.rows + r) * OUT + tid % OUT] : 0.f;
    }
    __syncthreads();
    const float* h2_tile = jb.h2 + ((size_t)net_i * tiles + tile) * H * BM;
    build_dz2_tile<OUT, false>(net + off_W3(IN), h2_tile, dOs, As, jb_i * 128, 128, B2S, nullptr, nullptr);
    build_h1_tile<IN>(net, Xs, Hs, kb_i * 128, 128, B2S);
    __syncthreads();
#pragma unroll 2
    for (int r = 0; r < BM; r += 4) {
      float4 a[8], b[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4*>(As + (i * 16 + ty) * B2S + r);
#pragma unroll
      for (int i = 0; i < 8; ++i) b[i] = *reinterpret_cast<const float4*>(Hs + (i * 16 + tx) * B2S + r);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
          acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
          acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
          acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
        }
    }
    __syncthreads();
  }
  float* out = jb.pw2 + ((size_t)net_i * jb.splits + split) * H * H;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j)
      out[(size_t)(jb_i * 128 + i * 16 + ty) * H + kb_i * 128 + j * 16 + tx] = acc[i][j];
}

}  // namespace cql
