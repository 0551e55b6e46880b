// mlp_tc_ts.cuh -- FP32-grade (tf32 3-term split) fused forward with operand A staged in TENSOR MEMORY.
//
// Why: with a 64-column W2 slice every tcgen05.mma (M=128, N=64, K=8) reads 4 KB of A and 2 KB of B per
// 32 tensor cycles from shared memory = 192 B/clk, above the SM's 128 B/clk shared-memory port, and the
// producers' operand stores compete for the same port (measured: tensor pipe 25 % busy, independent of
// producer count / ring depth / issue path).  Here the producers write H1 (hi and lo terms) straight
// from registers into TMEM with tcgen05.st -- a thread owns one row = one TMEM lane -- and the MMA takes
// A from TMEM ("TS" form), so shared memory only serves the resident W2 slice: 64 B/clk.
//
// TMEM columns: [0,128) two 64-column accumulators; [128,512) three A stages of 64 K: hi[64] | lo[64].
// warps 0-3 epilogue, 4-19 producers (lane quarter = warp % 4, K eighth = (warp-4) / 4), 20 MMA issuer.
#pragma once
#include "mlp_tc.cuh"

namespace cql {
namespace tc {

struct TsCfg : Cfg<true> {
  static constexpr int NPW = 16;
  static constexpr int KC = 64;                       // K per stage
  static constexpr int STAGES = 3;                    // 128 accumulator + 3 x 128 operand columns = all 512
  static constexpr int NCHUNK = H / KC;               // 8
  static constexpr int KPW = KC / (NPW / 4);          // K elements per producer warp per stage: 8
  static constexpr int MMA_WARP = 4 + NPW;
  static constexpr int THREADS = (5 + NPW) * 32;      // 672
  static constexpr int PROD_THREADS = NPW * 32;
  static constexpr uint32_t A_COL0 = 2 * NS;          // 128
  static constexpr uint32_t A_STAGE_COLS = 2 * KC;    // 64
  static constexpr uint32_t TMEM_ALLOC = 512;
  static constexpr uint32_t OFF_B = 0;
  static constexpr uint32_t OFF_W1 = B_BYTES;                  // float4[256] pair-packed W1|b1
  static constexpr uint32_t OFF_EB = OFF_W1 + H * 16;          // float4[NS]
  static constexpr uint32_t OFF_BAR = OFF_EB + NS * 16;
  static constexpr uint32_t N_BARS = 2 * STAGES + 4 + 2;
  static constexpr uint32_t OFF_SLOT = OFF_BAR + N_BARS * 8;
  static constexpr uint32_t SMEM_BYTES = OFF_SLOT + 16;
};

template <int IN, int OUT>
__global__ void __launch_bounds__(TsCfg::THREADS, 1) tc_fwd_ts_kernel(const TcFwdJobs jobs) {
  using C = TsCfg;
  extern __shared__ __align__(1024) uint8_t sm[];
  uint8_t* Bs = sm + C::OFF_B;
  float4* w1p = reinterpret_cast<float4*>(sm + C::OFF_W1);   // [k/2][2]: {wx_k,wx_k1,wy_k,wy_k1}, {wz_k,wz_k1,b_k,b_k1}
  float4* ebs = reinterpret_cast<float4*>(sm + C::OFF_EB);
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
    const uint32_t idesc = instr_desc(FMT_TF32, TM, C::NS);
    const uint32_t b_lbo = C::NS * 16;
    const uint32_t b_base = smem_u32(Bs);
    int cur_pair = -1;
    uint32_t it = 0, nb = 0, nd = 0, tcount = 0;
    for (int item = item_lo; item < item_hi; ++item) {
      const ItemInfo ii = decode_item<true>(jobs, item);
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
            const uint32_t a_hi = a_stage + j * C::UK, a_lo = a_hi + C::KC;
            umma_ts<true>(d_tmem, a_lo, b_hi, idesc, (c == 0 && j == 0) ? 0u : 1u);
            umma_ts<true>(d_tmem, a_hi, b_lo, idesc, 1u);
            umma_ts<true>(d_tmem, a_hi, b_hi, idesc, 1u);
          }
          umma_commit(&empty[s]);
          if (c == C::NCHUNK - 1) umma_commit(&tfull[acc]);
        }
        __syncwarp();
      }
      ++tcount;
    }
  } else if (warp >= 4) {
    // =============================== producers: layer 1 -> TMEM ===============================
    const int pw = warp - 4, ptid = tid - 128;
    const int kq = pw >> 2;                                        // which K eighth of every stage
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + C::A_COL0 + kq * C::KPW;
    int cur_netkey = -1;
    uint32_t it = 0;
    for (int item = item_lo; item < item_hi; ++item) {
      const ItemInfo ii = decode_item<true>(jobs, item);
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
        asm volatile("bar.sync 1, %0;" ::"n"(C::PROD_THREADS));
      }
      const int r = ii.tile * TM + (warp & 3) * 32 + lane;
      const float4 x = r < jb.rows ? __ldg(jb.X + r) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float2 xx = make_float2(x.x, x.x), xy = make_float2(x.y, x.y), xz = make_float2(x.z, x.z);
      for (int c = 0; c < C::NCHUNK; ++c, ++it) {
        const uint32_t s = it % C::STAGES;
        uint32_t hi[C::KPW], lo[C::KPW];
#pragma unroll
        for (int pp = 0; pp < C::KPW / 2; ++pp) {
          const int pr = (c * C::KC + kq * C::KPW) / 2 + pp;
          const float4 wA = w1p[2 * pr], wB = w1p[2 * pr + 1];
          float2 v = ffma2(xx, make_float2(wA.x, wA.y), make_float2(wB.z, wB.w));   // chain starts from the bias
          v = ffma2(xy, make_float2(wA.z, wA.w), v);
          if (IN == 3) v = ffma2(xz, make_float2(wB.x, wB.y), v);
          float h0, l0, h1, l1;
          split_tf32_fast(fmaxf(v.x, 0.f), h0, l0);
          split_tf32_fast(fmaxf(v.y, 0.f), h1, l1);
          hi[2 * pp] = __float_as_uint(h0); lo[2 * pp] = __float_as_uint(l0);
          hi[2 * pp + 1] = __float_as_uint(h1); lo[2 * pp + 1] = __float_as_uint(l1);
        }
        mbar_wait(&empty[s], ((it / C::STAGES) & 1) ^ 1);          // values are ready before the slot is: wait late
        tc_fence_after();
#pragma unroll
        for (int q = 0; q < C::KPW; q += 8) {
          tmem_st8(lane_base + s * C::A_STAGE_COLS + q, *reinterpret_cast<const uint32_t(*)[8]>(&hi[q]));
          tmem_st8(lane_base + s * C::A_STAGE_COLS + C::KC + q, *reinterpret_cast<const uint32_t(*)[8]>(&lo[q]));
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[s]);
      }
    }
  } else {
    // =============================== epilogue: TMEM -> layer 3 (+ H2) ===============================
    int cur_pair = -1;
    uint32_t tcount = 0;
    const int row_in_tile = warp * 32 + lane;
    for (int item = item_lo; item < item_hi; ++item) {
      const ItemInfo ii = decode_item<true>(jobs, item);
      const TcFwdJob& jb = jobs.j[ii.job];
      if (ii.pair_id != cur_pair) {
        cur_pair = ii.pair_id;
        asm volatile("bar.sync 2, 128;");
        const float* net = jb.params + (size_t)ii.net * NET_STRIDE;
        for (int cidx = tid; cidx < C::NS; cidx += 128) {
          const int col = ii.slice * C::NS + cidx;
          ebs[cidx] = make_float4(net[off_b2(IN) + col], net[off_W3(IN) + col],
                                  OUT == 2 ? net[off_W3(IN) + H + col] : 0.f, 0.f);
        }
        asm volatile("bar.sync 2, 128;");
      }
      const uint32_t acc = tcount & 1;
      mbar_wait(&tfull[acc], (tcount >> 1) & 1);
      tc_fence_after();
      const int row = ii.tile * TM + row_in_tile;
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
          const float hv = fmaxf(v[i] + e.x, 0.f);
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
