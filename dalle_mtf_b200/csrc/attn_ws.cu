// K4 (warp-specialised) — causal flash attention forward on tcgen05 / TMEM / TMA for sm_100a.
//
// Replaces mtf_transformer.attention.attention + the [S,S] additive -1e10 mask the reference materialises
// (src/dalle_mtf/models.py:221-227, 287-299): logits fp32 (TMEM accumulators), softmax fp32, P rounded to bf16 for the
// P.V product (what mtf does when it casts the weights to v's dtype), nothing of size S x S reaches HBM or even shared
// memory.  `scale` multiplies q.k (reference: 1.0 — mtf folds 1/sqrt(dh) into the q initialiser).
//
// Layout: qkv bf16 [B][S][3][H][dh] (the fused q|k|v projection output), out bf16 [B][S][H][dh],
//         lse f32 [B][H][S] (natural log of sum exp(scale*s)).
//
// One CTA = two neighbouring 128-query tiles of one (batch, head) that share every K / V block they both need
// ("ping-pong": while one tile's softmax warpgroup exponentiates, the tensor pipe works for the other tile).
//   warps 0-3  softmax warpgroup of tile 0: thread r owns query row r = TMEM lane r (no cross-thread reductions);
//   warps 4-7  softmax warpgroup of tile 1;
//   warp  8    TMA producer (one lane): Q tiles once, then a ring of K / V blocks (128 keys each);
//   warp  9    MMA issuer (one lane):  S_t = Q_t K_j^T (SS form)  and  O_t += P_t V_j  (TS form: P is read from TMEM).
// TMEM (512 columns): S_0 | S_1 (128 fp32 columns each) | O_0 | O_1 (dh columns each).  P_t (bf16, 64 columns) is
// written by the softmax threads over the first half of S_t with tcgen05.st and consumed in place as the A operand of
// the P.V product, so S_t's next logits can only be issued behind that product — the tensor pipe executes in issue
// order, which is exactly the dependency needed.  Issue order: S0_0 S1_0 | PV0_0 S0_1 PV1_0 S1_1 | PV0_1 S0_2 ...
// The output stays in TMEM across key blocks; a row is rescaled (tcgen05.ld -> scale -> tcgen05.st by its own softmax
// thread) only when its running maximum has moved by more than 2^8 since the scale it uses (lazy rescaling).
// All hand-offs are mbarriers: TMA -> MMA (k_full / v_full), MMA -> TMA (tcgen05.commit on k_empty / v_empty),
// MMA -> softmax (commit on s_ready, o_done), softmax -> MMA (p_ready, one arrive per warp).
#include "common.cuh"
#include "ptx.cuh"

namespace db200 {

namespace {

constexpr float LOG2E_F = 1.4426950408889634f;

// MUFU.EX2 directly: arguments are <= 8 after the running-max subtraction; ex2.approx(-inf) = +0.
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int DH>
__device__ __forceinline__ void ws_load_tile(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int chan, int row0,
                                             int b) {
#pragma unroll
  for (int t = 0; t < DH / 64; ++t) tma_load_4d(dst + t * (128 * 128), tm, bar, 64 * t, chan, row0, b);
}

template <int DH>
struct FwdWs {
  static constexpr int NS = (DH == 128) ? 2 : 4;           // K / V ring depth (128-key blocks)
  static constexpr uint32_t TILE = 128 * DH * 2;           // one [128][DH] bf16 operand tile
  static constexpr uint32_t BAR_BYTES = 256;
  static constexpr size_t SMEM = 1024 + 2 * TILE + 2 * NS * TILE + BAR_BYTES;
  static constexpr int THREADS = 320;
};

}  // namespace

template <int DH>
__global__ void __launch_bounds__(320, 1)
attn_fwd_ws_kernel(const __grid_constant__ CUtensorMap tmQKV, bf16* __restrict__ out, float* __restrict__ lse_out,
                   int S, int H, float scale) {
  using C = FwdWs<DH>;
  constexpr int NS = C::NS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t sQ = base, sK = sQ + 2 * C::TILE, sV = sK + NS * C::TILE;
  const uint32_t bars = sV + NS * C::TILE;
  // barrier map (8 bytes each)
  const uint32_t q_full = bars;                    // [2]
  const uint32_t k_full = bars + 16;               // [NS]
  const uint32_t v_full = k_full + 8 * NS;         // [NS]
  const uint32_t k_empty = v_full + 8 * NS;        // [NS]
  const uint32_t v_empty = k_empty + 8 * NS;       // [NS]
  const uint32_t s_ready = v_empty + 8 * NS;       // [2]
  const uint32_t p_ready = s_ready + 16;           // [2]
  const uint32_t o_done = p_ready + 16;            // [2]
  const uint32_t tmem_slot = o_done + 16;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // 1-D grid ordered by work: CTAs are dispatched in linear order, so ALL (batch, head) instances of the heaviest tile
  // pair (the latest queries: most key blocks) come first and the light ones fill the tail (LPT scheduling)
  const int n_qt = (S + 127) >> 7;
  const int n_bh = gridDim.x / ((n_qt + 1) >> 1);
  const int pair = ((n_qt + 1) >> 1) - 1 - (int)blockIdx.x / n_bh;
  const int h = ((int)blockIdx.x % n_bh) % H, b = ((int)blockIdx.x % n_bh) / H;
  // tile t of this CTA = query tile 2*pair + t; it needs key blocks 0 .. 2*pair + t (the last one is its diagonal)
  const int n_kv0 = 2 * pair + 1;
  const int n_kv1 = (2 * pair + 1 < n_qt) ? 2 * pair + 2 : 0;
  const int n_j = n_kv1 > n_kv0 ? n_kv1 : n_kv0;
  const int last_user = n_kv1 > 0 ? 1 : 0;  // the tile that issues the last product on every ring slot it shares

  if (tid == 0) {
    tma_prefetch_desc(&tmQKV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(q_full + 8 * i, 1);
      mbar_init(s_ready + 8 * i, 1);
      mbar_init(p_ready + 8 * i, 4);  // one arrive per softmax warp
      mbar_init(o_done + 8 * i, 1);
    }
    for (int i = 0; i < NS; ++i) {
      mbar_init(k_full + 8 * i, 1); mbar_init(v_full + 8 * i, 1);
      mbar_init(k_empty + 8 * i, 1); mbar_init(v_empty + 8 * i, 1);
    }
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;

  if (warp == 8) {
    // ------------------------------------------------------------------------------------------- TMA producer
    if (lane == 0) {
      mbar_expect_tx(q_full, C::TILE);
      ws_load_tile<DH>(sQ, &tmQKV, q_full, 0 * H + h, (2 * pair) * 128, b);
      if (n_kv1 > 0) {
        mbar_expect_tx(q_full + 8, C::TILE);
        ws_load_tile<DH>(sQ + C::TILE, &tmQKV, q_full + 8, 0 * H + h, (2 * pair + 1) * 128, b);
      }
      for (int j = 0; j < n_j; ++j) {
        const int st = j % NS;
        const uint32_t ph = (uint32_t)(j / NS) & 1u;
        mbar_wait(k_empty + 8 * st, ph ^ 1u);
        mbar_expect_tx(k_full + 8 * st, C::TILE);
        ws_load_tile<DH>(sK + st * C::TILE, &tmQKV, k_full + 8 * st, 1 * H + h, j * 128, b);
        mbar_wait(v_empty + 8 * st, ph ^ 1u);
        mbar_expect_tx(v_full + 8 * st, C::TILE);
        ws_load_tile<DH>(sV + st * C::TILE, &tmQKV, v_full + 8 * st, 2 * H + h, j * 128, b);
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);  // S = Q K^T : both K-major (K = dh)
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, DH, 0, 1);   // O = P V   : A from TMEM, B MN-major (K = keys)
      auto issue_s = [&](int t, int j) {
        const int st = j % NS;
        if (j == 0) mbar_wait(q_full + 8 * t, 0);
        mbar_wait(k_full + 8 * st, (uint32_t)(j / NS) & 1u);
        tc_fence_after();
        const uint32_t qb = sQ + t * C::TILE, kb = sK + st * C::TILE;
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk) {
          const uint32_t o = (kk / 4) * (128 * 128) + (kk % 4) * 32;
          umma_bf16_ss(tmem + t * 128, umma_smem_desc_sw128(qb + o, 0, 1024), umma_smem_desc_sw128(kb + o, 0, 1024),
                       idesc_s, kk > 0);
        }
        umma_commit(s_ready + 8 * t);
        if (t == last_user) umma_commit(k_empty + 8 * st);
      };
      auto issue_pv = [&](int t, int j) {
        const int st = j % NS;
        mbar_wait(v_full + 8 * st, (uint32_t)(j / NS) & 1u);
        mbar_wait(p_ready + 8 * t, (uint32_t)j & 1u);
        tc_fence_after();
        const uint32_t vb = sV + st * C::TILE;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)  // K = 128 keys, 16 per instruction = 8 packed TMEM columns of P
          umma_bf16_ts(tmem + 256 + t * DH, tmem + t * 128 + kk * 8, umma_smem_desc_sw128(vb + kk * 2048, 128 * 128, 1024),
                       idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
        umma_commit(o_done + 8 * t);
        if (t == last_user) umma_commit(v_empty + 8 * st);
      };
      issue_s(0, 0);
      if (n_kv1 > 0) issue_s(1, 0);
      for (int j = 0; j < n_j; ++j) {
        if (j < n_kv0) issue_pv(0, j);
        if (j + 1 < n_kv0) issue_s(0, j + 1);
        if (j < n_kv1) issue_pv(1, j);
        if (j + 1 < n_kv1) issue_s(1, j + 1);
      }
    }
  } else {
    // ------------------------------------------------------------------------------------------- softmax warpgroups
    const int t = warp >> 2;                       // tile of this warpgroup
    const int n_kv = t == 0 ? n_kv0 : n_kv1;
    if (n_kv > 0) {
      const int row = tid & 127;                   // query row inside the tile = TMEM lane
      const int qt = 2 * pair + t;
      const int qi = qt * 128 + row;
      const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;
      const uint32_t tS = tmem + t * 128 + lane_off, tO = tmem + 256 + t * DH + lane_off;
      const float c1 = scale * LOG2E_F;
      float m_used = -INFINITY;  // the maximum the accumulated P / O / l are scaled by
      float l_run = 0.f;
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(s_ready + 8 * t, (uint32_t)j & 1u);
        tc_fence_after();
        uint32_t sv[128];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld_x32(tS + c * 32, sv + c * 32);
        tmem_ld_wait();
        if (j == qt) {  // diagonal block: keys after the query are masked (src/dalle_mtf/models.py:221-227)
#pragma unroll
          for (int c = 0; c < 128; ++c)
            if (c > row) sv[c] = 0xff800000u;  // -inf
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int c = 0; c < 128; c += 4) {
          mx0 = fmaxf(mx0, __uint_as_float(sv[c]));
          mx1 = fmaxf(mx1, __uint_as_float(sv[c + 1]));
          mx2 = fmaxf(mx2, __uint_as_float(sv[c + 2]));
          mx3 = fmaxf(mx3, __uint_as_float(sv[c + 3]));
        }
        const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));  // finite: key 0 is visible to every query
        if (j == 0) {
          m_used = mx;
        } else {
          const bool need = (mx - m_used) * c1 > 8.f;
          if (__any_sync(0xffffffffu, need)) {  // tcgen05.ld / st are warp-collective
            // the previous P.V of this tile must have landed in O before it is rescaled
            mbar_wait(o_done + 8 * t, (uint32_t)(j - 1) & 1u);
            tc_fence_after();
            const float alpha = need ? ex2f((m_used - mx) * c1) : 1.f;
#pragma unroll
            for (int c = 0; c < DH / 32; ++c) {
              uint32_t r[32];
              tmem_ld_x32(tO + c * 32, r);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
              tmem_st_x32(tO + c * 32, r);
            }
            if (need) {
              l_run *= alpha;
              m_used = mx;
            }
          }
        }
        // P_j = 2^(c1 (s - m_used)) -> bf16 pairs -> the first 64 columns of S_t (A operand of the P.V product)
        const float mc = m_used * c1;
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = ex2f(fmaf(__uint_as_float(sv[c * 32 + i]), c1, -mc));
            const float p1 = ex2f(fmaf(__uint_as_float(sv[c * 32 + i + 1]), c1, -mc));
            l0 += p0;
            l1 += p1;
            pk[i >> 1] = pack_bf16x2(p0, p1);
          }
          tmem_st_x16(tS + c * 16, pk);
        }
        l_run += l0 + l1;
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_ready + 8 * t);
      }
      // ---- epilogue: O / l -> bf16, lse
      mbar_wait(o_done + 8 * t, (uint32_t)(n_kv - 1) & 1u);
      tc_fence_after();
      const float inv = 1.f / l_run;
      bf16* op = out + (((long long)b * S + qi) * H + h) * DH;
#pragma unroll
      for (int c = 0; c < DH / 32; ++c) {
        uint32_t r[32];
        tmem_ld_x32(tO + c * 32, r);
        tmem_ld_wait();
        if (qi < S) {
#pragma unroll
          for (int e = 0; e < 32; e += 8) {
            uint4 q;
            q.x = pack_bf16x2(__uint_as_float(r[e]) * inv, __uint_as_float(r[e + 1]) * inv);
            q.y = pack_bf16x2(__uint_as_float(r[e + 2]) * inv, __uint_as_float(r[e + 3]) * inv);
            q.z = pack_bf16x2(__uint_as_float(r[e + 4]) * inv, __uint_as_float(r[e + 5]) * inv);
            q.w = pack_bf16x2(__uint_as_float(r[e + 6]) * inv, __uint_as_float(r[e + 7]) * inv);
            *reinterpret_cast<uint4*>(op + c * 32 + e) = q;
          }
        }
      }
      if (qi < S) lse_out[((long long)b * H + h) * S + qi] = m_used * scale + logf(l_run);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// backward (warp-specialised).  Two launches of ONE kernel template, no atomics:
//   MODE 0  dK / dV : CTA = 128 keys (TMEM lanes = keys), loops over 64-query blocks i >= the first block that sees
//                     them, in the TRANSPOSED orientation:  S^T = K Q_i^T,  dP^T = V dO_i^T  (128 x 64, SS form);
//                     P^T / dS^T are written by the gradient warpgroups as bf16 over the first halves of S^T / dP^T and
//                     consumed from TMEM:  dV += P^T dO_i,  dK += dS^T Q_i  (TS form, B = the same Q_i / dO_i tiles
//                     read as MN-major operands).  lse / delta of the block's 64 queries are staged in shared memory.
//   MODE 1  dQ      : CTA = 128 queries (lanes = queries), loops over 64-key blocks j:  S = Q K_j^T,  dP = dO V_j^T,
//                     dS over dP in TMEM,  dQ += dS K_j  (TS form, K_j read MN-major).
//   p  = exp(scale*s - lse)   (0 where key > query or the query is out of range)
//   ds = p * (dp - delta) * scale
// Roles (384 threads): warps 0-3 / 4-7 = two gradient warpgroups that take alternate blocks (each owns one of the two
// S / dP buffers in TMEM), warp 8 = TMA producer (resident tiles, then a 4-deep ring of block tiles), warp 9 = MMA
// issuer, warp 10 = lse / delta stager (MODE 0).  Issue order  L0 L1 | G0 L2 | G1 L3 | ...  (L = the two logit
// products of a block, G = its gradient products): the tensor pipe computes block i+1's logits while a warpgroup
// turns block i's into P / dS.  TMEM: accumulators [0, 2 dh), buffers at 256 + 128 buf (S at +0, dP at +64).
// ------------------------------------------------------------------------------------------------------------------
namespace {
template <int DH>
struct BwdWs {
  static constexpr int NS = 4;
  static constexpr uint32_t TILE = 128 * DH * 2;  // resident [128][DH] tile
  static constexpr uint32_t HT = 64 * DH * 2;     // ring [64][DH] tile
  static constexpr uint32_t STAT_BYTES = NS * 128 * 4;
  static constexpr uint32_t BAR_BYTES = 256;
  static constexpr size_t SMEM = 1024 + 2 * TILE + NS * 2 * HT + STAT_BYTES + BAR_BYTES;
  static constexpr int THREADS = 384;
};
constexpr uint32_t T64B = 64 * 128;    // bytes of one [64 rows][64] swizzled sub-tile
constexpr uint32_t T128B = 128 * 128;  // bytes of one [128 rows][64] swizzled sub-tile

template <int DH>
__device__ __forceinline__ void ws_load_tile64(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int chan, int row0,
                                               int b) {
#pragma unroll
  for (int t = 0; t < DH / 64; ++t) tma_load_4d(dst + t * T64B, tm, bar, 64 * t, chan, row0, b);
}
}  // namespace

template <int DH, int MODE>
__global__ void __launch_bounds__(384, 1)
attn_bwd_ws_kernel(const __grid_constant__ CUtensorMap tmQKV128, const __grid_constant__ CUtensorMap tmQKV64,
                   const __grid_constant__ CUtensorMap tmDO128, const __grid_constant__ CUtensorMap tmDO64,
                   const float* __restrict__ lse, const float* __restrict__ delta, bf16* __restrict__ dqkv, int S,
                   int H, float scale) {
  using C = BwdWs<DH>;
  constexpr int NS = C::NS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t sX = base, sY = sX + C::TILE, sR = sY + C::TILE;  // ring stage st: B1 at sR + st*2*HT, B2 right behind
  const uint32_t sStat = sR + NS * 2 * C::HT;
  const uint32_t bars = sStat + C::STAT_BYTES;
  const uint32_t x_full = bars;                 // resident tiles
  const uint32_t r_full = bars + 8;             // [NS]
  const uint32_t r_empty = r_full + 8 * NS;     // [NS]
  const uint32_t stat_full = r_empty + 8 * NS;  // [NS]
  const uint32_t s_ready = stat_full + 8 * NS;  // [2]
  const uint32_t p_ready = s_ready + 16;        // [2]
  const uint32_t acc_done = p_ready + 16;
  const uint32_t tmem_slot = acc_done + 8;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  const float* stat = reinterpret_cast<const float*>(smem_raw + (sStat - raw));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // 1-D grid ordered by work (LPT): MODE 0: key block 0 sees every query block -> all its (batch, head) instances first;
  // MODE 1: the last query tile has the most key blocks -> first
  const int n_blk = (S + 127) >> 7;
  const int n_bh = gridDim.x / n_blk;
  const int blk = MODE == 0 ? (int)blockIdx.x / n_bh : n_blk - 1 - (int)blockIdx.x / n_bh;
  const int h = ((int)blockIdx.x % n_bh) % H, b = ((int)blockIdx.x % n_bh) / H;
  const int r0 = blk * 128;  // first key (MODE 0) / query (MODE 1) of this CTA
  const int it0 = MODE == 0 ? 2 * blk : 0;                                               // first 64-row block of the loop
  const int n_it = MODE == 0 ? (S + 63) / 64 - it0 : (min(S, r0 + 128) + 63) / 64;       // >= 1
  const long long bh = (long long)b * H + h;

  if (tid == 0) {
    tma_prefetch_desc(&tmQKV128);
    tma_prefetch_desc(&tmQKV64);
    tma_prefetch_desc(&tmDO128);
    tma_prefetch_desc(&tmDO64);
    mbar_init(x_full, 1);
    for (int i = 0; i < NS; ++i) {
      mbar_init(r_full + 8 * i, 1);
      mbar_init(r_empty + 8 * i, 1);
      mbar_init(stat_full + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(s_ready + 8 * i, 1);
      mbar_init(p_ready + 8 * i, 4);  // one arrive per warp of the warpgroup
    }
    mbar_init(acc_done, 1);
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;

  if (warp == 8) {
    // ------------------------------------------------------------------------------------------- TMA producer
    if (lane == 0) {
      mbar_expect_tx(x_full, 2 * C::TILE);
      if (MODE == 0) {
        ws_load_tile<DH>(sX, &tmQKV128, x_full, 1 * H + h, r0, b);   // K
        ws_load_tile<DH>(sY, &tmQKV128, x_full, 2 * H + h, r0, b);   // V
      } else {
        ws_load_tile<DH>(sX, &tmQKV128, x_full, 0 * H + h, r0, b);   // Q
        ws_load_tile<DH>(sY, &tmDO128, x_full, h, r0, b);            // dO
      }
      for (int it = 0; it < n_it; ++it) {
        const int st = it % NS;
        mbar_wait(r_empty + 8 * st, ((uint32_t)(it / NS) & 1u) ^ 1u);
        const uint32_t fb = r_full + 8 * st, d1 = sR + st * 2 * C::HT, d2 = d1 + C::HT;
        const int c0 = (it0 + it) * 64;
        mbar_expect_tx(fb, 2 * C::HT);
        if (MODE == 0) {
          ws_load_tile64<DH>(d1, &tmQKV64, fb, 0 * H + h, c0, b);    // Q_i
          ws_load_tile64<DH>(d2, &tmDO64, fb, h, c0, b);             // dO_i
        } else {
          ws_load_tile64<DH>(d1, &tmQKV64, fb, 1 * H + h, c0, b);    // K_j
          ws_load_tile64<DH>(d2, &tmQKV64, fb, 2 * H + h, c0, b);    // V_j
        }
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_l = umma_idesc_bf16(128, 64, 0, 0);   // logits: both operands K-major (K = dh)
      constexpr uint32_t idesc_g = umma_idesc_bf16(128, DH, 0, 1);   // gradients: A from TMEM, B MN-major (K = 64 rows)
      auto issue_l = [&](int it) {
        const int st = it % NS, buf = it & 1;
        mbar_wait(r_full + 8 * st, (uint32_t)(it / NS) & 1u);
        tc_fence_after();
        const uint32_t b1 = sR + st * 2 * C::HT, b2 = b1 + C::HT;
        const uint32_t tS = tmem + 256 + buf * 128, tdP = tS + 64;
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk)
          umma_bf16_ss(tS, umma_smem_desc_sw128(sX + (kk / 4) * T128B + (kk % 4) * 32, 0, 1024),
                       umma_smem_desc_sw128(b1 + (kk / 4) * T64B + (kk % 4) * 32, 0, 1024), idesc_l, kk > 0);
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk)
          umma_bf16_ss(tdP, umma_smem_desc_sw128(sY + (kk / 4) * T128B + (kk % 4) * 32, 0, 1024),
                       umma_smem_desc_sw128(b2 + (kk / 4) * T64B + (kk % 4) * 32, 0, 1024), idesc_l, kk > 0);
        umma_commit(s_ready + 8 * buf);
      };
      auto issue_g = [&](int it) {
        const int st = it % NS, buf = it & 1;
        mbar_wait(p_ready + 8 * buf, (uint32_t)(it >> 1) & 1u);
        tc_fence_after();
        const uint32_t b1 = sR + st * 2 * C::HT, b2 = b1 + C::HT;
        const uint32_t tS = tmem + 256 + buf * 128, tdP = tS + 64;
        const uint32_t acc = (it > 0) ? 1u : 0u;
        if (MODE == 0) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)  // dV += P^T dO_i
            umma_bf16_ts(tmem, tS + kk * 8, umma_smem_desc_sw128(b2 + kk * 2048, T64B, 1024), idesc_g, acc | (kk > 0));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)  // dK += dS^T Q_i
            umma_bf16_ts(tmem + DH, tdP + kk * 8, umma_smem_desc_sw128(b1 + kk * 2048, T64B, 1024), idesc_g,
                         acc | (kk > 0));
        } else {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)  // dQ += dS K_j
            umma_bf16_ts(tmem, tdP + kk * 8, umma_smem_desc_sw128(b1 + kk * 2048, T64B, 1024), idesc_g, acc | (kk > 0));
        }
        umma_commit(r_empty + 8 * st);
        if (it == n_it - 1) umma_commit(acc_done);
      };
      mbar_wait(x_full, 0);
      issue_l(0);
      if (n_it > 1) issue_l(1);
      for (int it = 0; it < n_it; ++it) {
        issue_g(it);
        if (it + 2 < n_it) issue_l(it + 2);
      }
    }
  } else if (warp == 10) {
    // ------------------------------------------------------------------------------------------- lse / delta stager
    if (MODE == 0) {
      float* stat_w = reinterpret_cast<float*>(smem_raw + (sStat - raw));
      for (int it = 0; it < n_it; ++it) {
        const int st = it % NS;
        mbar_wait(r_empty + 8 * st, ((uint32_t)(it / NS) & 1u) ^ 1u);
        const int c0 = (it0 + it) * 64;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int q = c0 + lane + 32 * u;
          float l2 = INFINITY, d = 0.f;  // out-of-range query: p = 2^(-inf) = 0
          if (q < S) {
            l2 = lse[bh * S + q] * LOG2E_F;
            d = delta[bh * S + q];
          }
          stat_w[st * 128 + lane + 32 * u] = l2;
          stat_w[st * 128 + 64 + lane + 32 * u] = d;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(stat_full + 8 * st);
      }
    }
  } else if (warp < 8) {
    // ------------------------------------------------------------------------------------------- gradient warpgroups
    const int g = warp >> 2;
    const int row = tid & 127;                 // TMEM lane: key (MODE 0) / query (MODE 1) inside the tile
    const int ri = r0 + row;
    const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;
    const uint32_t tS = tmem + 256 + g * 128 + lane_off, tdP = tS + 64;
    const float c1 = scale * LOG2E_F;
    float lse2 = INFINITY, dl = 0.f;           // MODE 1: this query's statistics
    if (MODE == 1 && ri < S) {
      lse2 = lse[bh * S + ri] * LOG2E_F;
      dl = delta[bh * S + ri];
    }
    for (int it = g; it < n_it; it += 2) {
      const int st = it % NS;
      const int c0 = (it0 + it) * 64;
      mbar_wait(s_ready + 8 * g, (uint32_t)(it >> 1) & 1u);
      tc_fence_after();
      if (MODE == 0) mbar_wait(stat_full + 8 * st, (uint32_t)(it / NS) & 1u);
      // causal mask (src/dalle_mtf/models.py:221-227): the pair is dropped where key > query.
      //   MODE 0: key = ri, query = c0 + c  -> dropped where c < ri - c0;   MODE 1: query = ri, key = c0 + c -> c > ri - c0
      const int dgl = ri - c0;
      const bool diag = MODE == 0 ? (dgl > 0) : (dgl < 63);
      const float4* sl = reinterpret_cast<const float4*>(stat + st * 128);
      const float4* sd = reinterpret_cast<const float4*>(stat + st * 128 + 64);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t rs[32], rd[32], pk[16], dk[16];
        tmem_ld_x32(tS + c * 32, rs);
        tmem_ld_x32(tdP + c * 32, rd);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          float l2[4], dd[4];
          if (MODE == 0) {
            const float4 a = sl[(c * 32 + i) >> 2], d4 = sd[(c * 32 + i) >> 2];
            l2[0] = a.x; l2[1] = a.y; l2[2] = a.z; l2[3] = a.w;
            dd[0] = d4.x; dd[1] = d4.y; dd[2] = d4.z; dd[3] = d4.w;
          } else {
            l2[0] = l2[1] = l2[2] = l2[3] = lse2;
            dd[0] = dd[1] = dd[2] = dd[3] = dl;
          }
          float p[4], ds[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            p[e] = ex2f(fmaf(__uint_as_float(rs[i + e]), c1, -l2[e]));
            if (diag) {
              const int cc = c * 32 + i + e;
              if (MODE == 0 ? (cc < dgl) : (cc > dgl)) p[e] = 0.f;
            }
            ds[e] = (p[e] * scale) * (__uint_as_float(rd[i + e]) - dd[e]);
          }
          pk[i >> 1] = pack_bf16x2(p[0], p[1]);
          pk[(i >> 1) + 1] = pack_bf16x2(p[2], p[3]);
          dk[i >> 1] = pack_bf16x2(ds[0], ds[1]);
          dk[(i >> 1) + 1] = pack_bf16x2(ds[2], ds[3]);
        }
        if (MODE == 0) tmem_st_x16(tS + c * 16, pk);
        tmem_st_x16(tdP + c * 16, dk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready + 8 * g);
    }
    // ---- epilogue: accumulators -> bf16 rows of dqkv.  MODE 0: warpgroup 0 writes dV, 1 writes dK; MODE 1: half of dQ each
    mbar_wait(acc_done, 0);
    tc_fence_after();
    constexpr int NCH = MODE == 0 ? DH / 32 : DH / 64;
    const int which = MODE == 0 ? (g == 0 ? 2 : 1) : 0;
    const uint32_t tsrc = tmem + lane_off + (MODE == 0 ? g * DH : g * (DH / 2));
    bf16* dst = dqkv + ((((long long)b * S + ri) * 3 + which) * H + h) * DH + (MODE == 0 ? 0 : g * (DH / 2));
#pragma unroll 1
    for (int c = 0; c < NCH; ++c) {
      uint32_t r[32];
      tmem_ld_x32(tsrc + c * 32, r);
      tmem_ld_wait();
      if (ri < S) {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          uint4 q;
          q.x = pack_bf16x2(__uint_as_float(r[q4 * 8 + 0]), __uint_as_float(r[q4 * 8 + 1]));
          q.y = pack_bf16x2(__uint_as_float(r[q4 * 8 + 2]), __uint_as_float(r[q4 * 8 + 3]));
          q.z = pack_bf16x2(__uint_as_float(r[q4 * 8 + 4]), __uint_as_float(r[q4 * 8 + 5]));
          q.w = pack_bf16x2(__uint_as_float(r[q4 * 8 + 6]), __uint_as_float(r[q4 * 8 + 7]));
          *reinterpret_cast<uint4*>(dst + c * 32 + q4 * 8) = q;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
static int ws_qkv_map(CUtensorMap* tm, const void* qkv, int B, int S, int H, int dh, uint32_t box_rows) {
  uint64_t dims[4] = {(uint64_t)dh, (uint64_t)3 * H, (uint64_t)S, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)dh * 2, (uint64_t)3 * H * dh * 2, (uint64_t)S * 3 * H * dh * 2};
  uint32_t box[4] = {64, 1, box_rows, 1};
  return make_tmap_bf16(tm, qkv, 4, dims, strides, box);
}

template <int DH>
static int fwd_ws_launch_t(cudaStream_t stream, const void* qkv, void* out, float* lse, int B, int S, int H,
                           float scale) {
  using C = FwdWs<DH>;
  CUtensorMap tm;
  int rc = ws_qkv_map(&tm, qkv, B, S, H, DH, 128);
  if (rc != DB200_OK) return rc;
  static const cudaError_t attr = cudaFuncSetAttribute(attn_fwd_ws_kernel<DH>,
                                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
  DB200_CUDA(attr);
  const int n_qt = (S + 127) / 128;
  dim3 grid(((n_qt + 1) / 2) * H * B);
  attn_fwd_ws_kernel<DH><<<grid, C::THREADS, C::SMEM, stream>>>(tm, (bf16*)out, lse, S, H, scale);
  return check_launch("attn_fwd_ws_kernel");
}

int attn_fwd_ws_launch(cudaStream_t stream, const void* qkv, void* out, float* lse, int B, int S, int H, int dh,
                       float scale) {
  if (dh == 128) return fwd_ws_launch_t<128>(stream, qkv, out, lse, B, S, H, scale);
  return fwd_ws_launch_t<64>(stream, qkv, out, lse, B, S, H, scale);
}

static int ws_o_map(CUtensorMap* tm, const void* o, int B, int S, int H, int dh, uint32_t box_rows) {
  uint64_t dims[4] = {(uint64_t)dh, (uint64_t)H, (uint64_t)S, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)dh * 2, (uint64_t)H * dh * 2, (uint64_t)S * H * dh * 2};
  uint32_t box[4] = {64, 1, box_rows, 1};
  return make_tmap_bf16(tm, o, 4, dims, strides, box);
}

template <int DH>
static int bwd_ws_launch_t(cudaStream_t stream, const void* qkv, const void* dout, const float* lse,
                           const float* delta, void* dqkv, int B, int S, int H, float scale) {
  using C = BwdWs<DH>;
  CUtensorMap q128, q64, o128, o64;
  int rc = ws_qkv_map(&q128, qkv, B, S, H, DH, 128);
  if (rc == DB200_OK) rc = ws_qkv_map(&q64, qkv, B, S, H, DH, 64);
  if (rc == DB200_OK) rc = ws_o_map(&o128, dout, B, S, H, DH, 128);
  if (rc == DB200_OK) rc = ws_o_map(&o64, dout, B, S, H, DH, 64);
  if (rc != DB200_OK) return rc;
  static const cudaError_t a0 = cudaFuncSetAttribute(attn_bwd_ws_kernel<DH, 0>,
                                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
  static const cudaError_t a1 = cudaFuncSetAttribute(attn_bwd_ws_kernel<DH, 1>,
                                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
  DB200_CUDA(a0);
  DB200_CUDA(a1);
  dim3 grid(((S + 127) / 128) * H * B);
  attn_bwd_ws_kernel<DH, 0><<<grid, C::THREADS, C::SMEM, stream>>>(q128, q64, o128, o64, lse, delta, (bf16*)dqkv, S, H,
                                                                 scale);
  rc = check_launch("attn_bwd_ws_kernel<dkdv>");
  if (rc != DB200_OK) return rc;
  attn_bwd_ws_kernel<DH, 1><<<grid, C::THREADS, C::SMEM, stream>>>(q128, q64, o128, o64, lse, delta, (bf16*)dqkv, S, H,
                                                                 scale);
  return check_launch("attn_bwd_ws_kernel<dq>");
}

int attn_bwd_ws_launch(cudaStream_t stream, const void* qkv, const void* dout, const float* lse, const float* delta,
                       void* dqkv, int B, int S, int H, int dh, float scale) {
  if (dh == 128) return bwd_ws_launch_t<128>(stream, qkv, dout, lse, delta, dqkv, B, S, H, scale);
  return bwd_ws_launch_t<64>(stream, qkv, dout, lse, delta, dqkv, B, S, H, scale);
}

}  // namespace db200
