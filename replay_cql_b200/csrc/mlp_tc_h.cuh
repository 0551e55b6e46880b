// mlp_tc_h.cuh -- FP32-grade hidden-layer contraction on the fp16 tensor-core path ("f16x3").
//
// Same idea as the tf32 3-term split (mlp_tc_ts.cuh): x = x_hi + x_lo with 11-bit significands each, and
// a_lo*w_hi + a_hi*w_lo + a_hi*w_hi accumulated in FP32 tensor memory.  fp16 has the SAME significand width as
// tf32 (10 explicit bits), so the split is equally exact -- but kind::f16 runs at twice the kind::tf32 rate and its
// operands are half as wide, so a 128-column W2 slice (hi + lo) fits in the 128 KB resident operand buffer: a row
// tile is visited by 2 (net, slice) work items instead of 4, halving the producers' layer-1 work as well.
//
// What fp16 lacks is exponent range (5 bits).  Every operand row therefore carries an exact power-of-two scale:
//   * A (activations, one row per thread): s_m = 2^e from a bound on the row, |H1[m][k]| <= sum_c |x_c| max_k|W1[k][c]|
//     + max_k|b1[k]|, chosen so that the row stays below 2^15;
//   * B (W2 rows, packed once per Adam step): s_n = 2^e from the row maximum, stored with the packed operand;
// the epilogue multiplies the accumulator by 1/(s_m s_n) -- exact.  Small entries of a scaled row fall into fp16
// subnormals, whose spacing 2^-24 is still 2^-39 of the row maximum: the representation error is
// max(2^-22 |x|, 2^-39 max|row|), far inside the 1e-4 gates.
//
// TMEM columns: [0,256) two 128-column accumulators; [256,512) four A stages of 64 K: hi[32 cols] | lo[32 cols]
// (two fp16 per 32-bit column).  warps 0-3 epilogue, 4-19 producers, 20 MMA issuer (as in mlp_tc_ts.cuh).
#pragma once
#include <cuda_fp16.h>
#include "mlp_tc.cuh"

namespace cql {
namespace tc {

struct HCfg {
  static constexpr int ES = 2, EPC = 8, UK = 16;
  static constexpr int NS = 128;                      // output columns per work item
  static constexpr int SLICES = H / NS;               // 2
  static constexpr int NPW = 16;
  static constexpr int KC = 64;                       // K per stage
  static constexpr int STAGES = 4;
  static constexpr int NCHUNK = H / KC;               // 4
  static constexpr int KPW = KC / (NPW / 4);          // K elements per producer warp per stage: 16 = 8 columns
  static constexpr int MMA_WARP = 4 + NPW;
  static constexpr int THREADS = (5 + NPW) * 32;      // 672
  static constexpr int PROD_THREADS = NPW * 32;
  static constexpr uint32_t B_TERM_BYTES = NS * H * ES;            // 64 KB
  static constexpr uint32_t B_BYTES = 2 * B_TERM_BYTES;            // 128 KB (hi | lo)
  static constexpr size_t META_OFF = (size_t)SLICES * B_BYTES;     // float inv_s[256] | float wmax[8]
  static constexpr size_t PACKED_NET_BYTES = META_OFF + 2048;
  static constexpr uint32_t A_COL0 = 2 * NS;          // 256
  static constexpr uint32_t A_STAGE_COLS = KC;        // 32 hi + 32 lo columns
  static constexpr uint32_t A_LO_COLS = KC / 2;
  static constexpr uint32_t TMEM_ALLOC = 512;
  static constexpr uint32_t OFF_B = 0;
  static constexpr uint32_t OFF_W1 = B_BYTES;                  // float4[256] pair-packed W1|b1
  static constexpr uint32_t OFF_EB = OFF_W1 + H * 16;          // float4[NS]
  static constexpr uint32_t OFF_BAR = OFF_EB + NS * 16;
  static constexpr uint32_t N_BARS = 2 * STAGES + 4 + 2;
  static constexpr uint32_t OFF_SLOT = OFF_BAR + N_BARS * 8;
  static constexpr uint32_t SMEM_BYTES = OFF_SLOT + 16;
};

// meta block of a packed net
struct HMeta {
  float inv_s[H];     // 1 / scale of operand row n
  float wmax[8];      // max|W1[:,0]|, max|W1[:,1]|, max|W1[:,2]|, max|b1|, max|W3[0,:]|, max|W3[1,:]|, -, -
};

// power-of-two scale that brings `bound` below 2^15, and its inverse (bit arithmetic on the exponent; exact)
__device__ __forceinline__ void pow2_scale(float bound, float& s, float& inv_s) {
  int e = (int)((__float_as_uint(bound) >> 23) & 0xffu);      // bound in [2^(e-127), 2^(e-126))
  e = max(e, 20);                                             // zero / denormal rows: any finite scale works
  s = __uint_as_float((uint32_t)(268 - e) << 23);             // 2^(141-e):  bound * s < 2^15
  inv_s = __uint_as_float((uint32_t)(e - 14) << 23);          // 2^(e-141)
}

// x (already scaled) -> fp16 hi and fp16 lo = rn(x - hi), two values per 32-bit word (low half = first value)
__device__ __forceinline__ void split_h2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x0, x1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// packed fp32x2 multiply
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "mul.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}

// Packed weights, per net: per slice [term hi|lo][kchunk16 (8 fp16)][n_local/8][8][16 B] (chunk_off with rows = NS),
// then HMeta.  One WARP per operand row n (32 lanes = the 32 K-chunks), so the row maximum is a warp reduction.
// transpose=0: B[n][k] = W2[n][k] (forward);  1: B[n][k] = W2[k][n] (backward dH1 = dZ2 W2).
__global__ void __launch_bounds__(256) k_pack_multi_h(const PackJobs jobs, int out_dim_actor) {
  using C = HCfg;
  const PackJobs::J jb = jobs.j[blockIdx.y];
  const int c = blockIdx.x * blockDim.x + threadIdx.x;        // 16-byte chunk id: n * 32 + kc
  const int n = c >> 5, kc = c & 31;
  const float* W2 = jb.net + off_W2(jb.in_dim);
  float v[8];
  float mx = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = kc * 8 + i;
    v[i] = jb.transpose ? W2[(size_t)k * H + n] : W2[(size_t)n * H + k];
    mx = fmaxf(mx, fabsf(v[i]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float s, inv_s;
  pow2_scale(mx, s, inv_s);
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split_h2(v[2 * i] * s, v[2 * i + 1] * s, hi[i], lo[i]);
  const int slice = n / C::NS, nl = n % C::NS;
  uint8_t* base = jb.dst + (size_t)slice * C::B_BYTES + chunk_off(C::NS, nl, kc);
  *reinterpret_cast<uint4*>(base) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(base + C::B_TERM_BYTES) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  HMeta* meta = reinterpret_cast<HMeta*>(jb.dst + C::META_OFF);
  if (kc == 0) meta->inv_s[n] = inv_s;
  if (blockIdx.x == 0) {                                      // maxima of the small layers (scales of the A rows)
    __shared__ int wm[8];
    if (threadIdx.x < 8) wm[threadIdx.x] = 0;
    __syncthreads();
    const int out_dim = jb.in_dim == 2 ? out_dim_actor : 1;
    const int k = threadIdx.x;                                // 256 threads = 256 hidden units
    const float* W1 = jb.net + off_W1(jb.in_dim);
    for (int cc = 0; cc < jb.in_dim; ++cc) atomicMax(&wm[cc], __float_as_int(fabsf(W1[k * jb.in_dim + cc])));
    atomicMax(&wm[3], __float_as_int(fabsf(jb.net[off_b1(jb.in_dim) + k])));
    for (int o = 0; o < out_dim; ++o) atomicMax(&wm[4 + o], __float_as_int(fabsf(jb.net[off_W3(jb.in_dim) + o * H + k])));
    __syncthreads();
    if (threadIdx.x < 8) meta->wmax[threadIdx.x] = __int_as_float(wm[threadIdx.x]);
  }
}

struct HItem { int job, net, slice, tile, pair_id; };
__device__ __forceinline__ HItem decode_item_h(const TcFwdJobs& jobs, int item) {
  HItem it;
  it.job = 0;
  while (it.job + 1 < jobs.n && item >= jobs.item_begin[it.job + 1]) ++it.job;
  const int local = item - jobs.item_begin[it.job];
  const int tiles = (jobs.j[it.job].rows + TM - 1) / TM;
  const int pair = local / tiles;
  it.tile = local % tiles;
  it.net = pair / HCfg::SLICES;
  it.slice = pair % HCfg::SLICES;
  it.pair_id = it.job * 64 + pair;
  return it;
}

// bound on |H1[row][:]| from the row's input and the per-net maxima of |W1| columns and |b1|
__device__ __forceinline__ float h1_row_bound(const float4& x, const float4& wm) {
  return fmaf(fabsf(x.x), wm.x, fmaf(fabsf(x.y), wm.y, fmaf(fabsf(x.z), wm.z, wm.w)));
}

template <int IN, int OUT>
__global__ void __launch_bounds__(HCfg::THREADS, 1) tc_fwd_h_kernel(const TcFwdJobs jobs) {
  using C = HCfg;
  extern __shared__ __align__(1024) uint8_t sm[];
  uint8_t* Bs = sm + C::OFF_B;
  float4* w1p = reinterpret_cast<float4*>(sm + C::OFF_W1);   // [k/2][2]: {wx_k,wx_k1,wy_k,wy_k1}, {wz_k,wz_k1,b_k,b_k1}
  float4* ebs = reinterpret_cast<float4*>(sm + C::OFF_EB);   // [NS]: b2, w3_0, w3_1, 1/s_n
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + C::OFF_BAR);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::STAGES;
  uint64_t* tfull = bars + 2 * C::STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* bload = tempty + 2;
  uint64_t* drain = bload + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(sm + C::OFF_SLOT);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int total = jobs.item_begin[jobs.n];
  const int item_lo = (int)((long long)total * blockIdx.x / gridDim.x);
  const int item_hi = (int)((long long)total * (blockIdx.x + 1) / gridDim.x);

  if (warp == C::MMA_WARP) {
    tmem_alloc(slot, C::TMEM_ALLOC);
    if (lane == 0) {
      for (int s = 0; s < C::STAGES; ++s) { mbar_init(&full[s], C::NPW); mbar_init(&empty[s], 1); }
      for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
      mbar_init(bload, 1);
      mbar_init(drain, 1);
      fence_mbar_init();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;

  if (warp == C::MMA_WARP) {
    // =============================== MMA issuer ===============================
    const uint32_t idesc = instr_desc(FMT_F16, TM, C::NS);
    const uint32_t b_lbo = C::NS * 16;
    const uint32_t b_base = smem_u32(Bs);
    int cur_pair = -1;
    uint32_t it = 0, nb = 0, nd = 0, tcount = 0;
    for (int item = item_lo; item < item_hi; ++item) {
      const HItem ii = decode_item_h(jobs, item);
      if (ii.pair_id != cur_pair) {
        if (cur_pair >= 0) {
          if (elect_one()) umma_commit(drain);
          __syncwarp();
          mbar_wait(drain, nd & 1);
          ++nd;
        }
        const TcFwdJob& jb = jobs.j[ii.job];
        const uint8_t* src = jb.packed + (size_t)ii.net * C::PACKED_NET_BYTES + (size_t)ii.slice * C::B_BYTES;
        if (elect_one()) {
          mbar_arrive_expect_tx(bload, C::B_BYTES);
          for (uint32_t o = 0; o < C::B_BYTES; o += 32768) bulk_g2s(Bs + o, src + o, 32768, bload);
        }
        __syncwarp();
        mbar_wait(bload, nb & 1);
        ++nb;
        cur_pair = ii.pair_id;
      }
      const uint32_t acc = tcount & 1;
      mbar_wait(&tempty[acc], ((tcount >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem + acc * C::NS;
      for (int c = 0; c < C::NCHUNK; ++c, ++it) {
        const uint32_t s = it % C::STAGES;
        mbar_wait(&full[s], (it / C::STAGES) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_stage = tmem + C::A_COL0 + s * C::A_STAGE_COLS;
#pragma unroll
          for (int j = 0; j < C::KC / C::UK; ++j) {
            const uint32_t g = c * (C::KC / C::UK) + j;
            const uint64_t b_hi = smem_desc(b_base + 2 * g * b_lbo, b_lbo, 128);
            const uint64_t b_lo = smem_desc(b_base + C::B_TERM_BYTES + 2 * g * b_lbo, b_lbo, 128);
            const uint32_t a_hi = a_stage + j * 8, a_lo = a_hi + C::A_LO_COLS;     // 8 columns = 16 fp16 per MMA
            umma_ts<false>(d_tmem, a_lo, b_hi, idesc, (c == 0 && j == 0) ? 0u : 1u);
            umma_ts<false>(d_tmem, a_hi, b_lo, idesc, 1u);
            umma_ts<false>(d_tmem, a_hi, b_hi, idesc, 1u);
          }
          umma_commit(&empty[s]);
          if (c == C::NCHUNK - 1) umma_commit(&tfull[acc]);
        }
        __syncwarp();
      }
      ++tcount;
    }
  } else if (warp >= 4) {
    // =============================== producers: layer 1 -> scaled fp16 hi|lo -> TMEM ===============================
    const int pw = warp - 4, ptid = tid - 128;
    const int kq = pw >> 2;                                        // which K quarter of every stage
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + C::A_COL0 + kq * (C::KPW / 2);
    int cur_netkey = -1;
    float4 wm = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t it = 0;
    for (int item = item_lo; item < item_hi; ++item) {
      const HItem ii = decode_item_h(jobs, item);
      const TcFwdJob& jb = jobs.j[ii.job];
      const int netkey = ii.job * 64 + ii.net;
      if (netkey != cur_netkey) {
        cur_netkey = netkey;
        asm volatile("bar.sync 1, %0;" ::"n"(C::PROD_THREADS));
        const float* net = jb.params + (size_t)ii.net * NET_STRIDE;
        for (int pr = ptid; pr < H / 2; pr += C::PROD_THREADS) {
          const int k = 2 * pr;
          const float* wa = net + off_W1(IN) + k * IN;
          const float* wb = wa + IN;
          w1p[2 * pr] = make_float4(wa[0], wb[0], wa[1], wb[1]);
          w1p[2 * pr + 1] = make_float4(IN == 3 ? wa[2] : 0.f, IN == 3 ? wb[2] : 0.f, net[off_b1(IN) + k], net[off_b1(IN) + k + 1]);
        }
        const HMeta* meta = reinterpret_cast<const HMeta*>(jb.packed + (size_t)ii.net * C::PACKED_NET_BYTES + C::META_OFF);
        wm = make_float4(__ldg(&meta->wmax[0]), __ldg(&meta->wmax[1]), IN == 3 ? __ldg(&meta->wmax[2]) : 0.f, __ldg(&meta->wmax[3]));
        asm volatile("bar.sync 1, %0;" ::"n"(C::PROD_THREADS));
      }
      const int r = ii.tile * TM + (warp & 3) * 32 + lane;
      const float4 x = r < jb.rows ? __ldg(jb.X + r) : make_float4(0.f, 0.f, 0.f, 0.f);
      float sa, inv_sa;
      pow2_scale(h1_row_bound(x, wm), sa, inv_sa);
      const float2 xx = make_float2(x.x, x.x), xy = make_float2(x.y, x.y), xz = make_float2(x.z, x.z), ss = make_float2(sa, sa);
      for (int c = 0; c < C::NCHUNK; ++c, ++it) {
        const uint32_t s = it % C::STAGES;
        uint32_t hi[C::KPW / 2], lo[C::KPW / 2];
#pragma unroll
        for (int pp = 0; pp < C::KPW / 2; ++pp) {
          const int pr = (c * C::KC + kq * C::KPW) / 2 + pp;
          const float4 wA = w1p[2 * pr], wB = w1p[2 * pr + 1];
          float2 v = ffma2(xx, make_float2(wA.x, wA.y), make_float2(wB.z, wB.w));   // chain starts from the bias
          v = ffma2(xy, make_float2(wA.z, wA.w), v);
          if (IN == 3) v = ffma2(xz, make_float2(wB.x, wB.y), v);
          v = fmul2(make_float2(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f)), ss);             // exact: power-of-two scale
          split_h2(v.x, v.y, hi[pp], lo[pp]);
        }
        mbar_wait(&empty[s], ((it / C::STAGES) & 1) ^ 1);          // values are ready before the slot is: wait late
        tc_fence_after();
        tmem_st8(lane_base + s * C::A_STAGE_COLS, hi);
        tmem_st8(lane_base + s * C::A_STAGE_COLS + C::A_LO_COLS, lo);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[s]);
      }
    }
  } else {
    // =============================== epilogue: TMEM -> unscale -> layer 3 (+ H2) ===============================
    int cur_pair = -1;
    uint32_t tcount = 0;
    const int row_in_tile = warp * 32 + lane;
    float4 wm = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int item = item_lo; item < item_hi; ++item) {
      const HItem ii = decode_item_h(jobs, item);
      const TcFwdJob& jb = jobs.j[ii.job];
      if (ii.pair_id != cur_pair) {
        cur_pair = ii.pair_id;
        asm volatile("bar.sync 2, 128;");
        const float* net = jb.params + (size_t)ii.net * NET_STRIDE;
        const HMeta* meta = reinterpret_cast<const HMeta*>(jb.packed + (size_t)ii.net * C::PACKED_NET_BYTES + C::META_OFF);
        for (int cidx = tid; cidx < C::NS; cidx += 128) {
          const int col = ii.slice * C::NS + cidx;
          ebs[cidx] = make_float4(net[off_b2(IN) + col], net[off_W3(IN) + col],
                                  OUT == 2 ? net[off_W3(IN) + H + col] : 0.f, __ldg(&meta->inv_s[col]));
        }
        wm = make_float4(__ldg(&meta->wmax[0]), __ldg(&meta->wmax[1]), IN == 3 ? __ldg(&meta->wmax[2]) : 0.f, __ldg(&meta->wmax[3]));
        asm volatile("bar.sync 2, 128;");
      }
      const uint32_t acc = tcount & 1;
      const int row = ii.tile * TM + row_in_tile;
      const float4 x = row < jb.rows ? __ldg(jb.X + row) : make_float4(0.f, 0.f, 0.f, 0.f);
      float sa, inv_sa;
      pow2_scale(h1_row_bound(x, wm), sa, inv_sa);                 // the producers' scale of this row, recomputed
      mbar_wait(&tfull[acc], (tcount >> 1) & 1);
      tc_fence_after();
      const bool store_h2 = jb.h2 != nullptr;
      const int tiles64 = (jb.rows + 63) / 64;
      float* h2row = store_h2 ? jb.h2 + (((size_t)ii.net * tiles64 + (row >> 6)) * H + ii.slice * C::NS) * 64 + (row & 63)
                              : nullptr;
      float q0 = 0.f, q1 = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < C::NS; c0 += 32) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + acc * C::NS + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float4 e = ebs[c0 + i];
          const float hv = fmaxf(fmaf(v[i] * inv_sa, e.w, e.x), 0.f);
          v[i] = hv;
          q0 = fmaf(hv, e.y, q0);
          if (OUT == 2) q1 = fmaf(hv, e.z, q1);
        }
        if (store_h2 && (row >> 6) < tiles64) {
#pragma unroll
          for (int i = 0; i < 32; ++i) h2row[(size_t)(c0 + i) * 64] = v[i];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (row < jb.rows) {
        float* o = jb.out_part + (((size_t)ii.net * C::SLICES + ii.slice) * jb.rows + row) * OUT;
        o[0] = q0;
        if (OUT == 2) o[1] = q1;
      }
      ++tcount;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == C::MMA_WARP) tmem_dealloc(tmem, C::TMEM_ALLOC);
}

}  // namespace tc
}  // namespace cql
