// K4 — causal flash attention, forward and backward, on tcgen05 / TMEM for sm_100a.
//
// Replaces mtf_transformer.attention.attention + the [S,S] additive -1e10 mask the reference materialises
// (src/dalle_mtf/models.py:221-227, 287-299).  Logits are fp32 (TMEM accumulators), softmax is fp32, P is rounded to
// bf16 for the PV product (what mtf does when it casts the weights to v's dtype), nothing of size S x S reaches HBM.
// `scale` multiplies q.k (reference: 1.0 — mtf folds 1/sqrt(dh) into the q initialiser).
//
// Layout: qkv bf16 [B][S][3][H][dh] (the fused q|k|v projection output), out / dout bf16 [B][S][H][dh],
//         lse f32 [B][H][S] (natural log of sum exp(scale*s)), dqkv like qkv.
//
// Design (all three kernels): 256 threads; threads t and t+128 share TMEM lane t (= one query row of the 128-row tile,
// or one key row of the dK/dV accumulators) and each handle half of its columns.  Operand tiles are brought by TMA as [rows][64] bf16 sub-tiles with the
// 128-byte swizzle, which serves both as a K-major UMMA operand (K = head dim) and as an MN-major one (K = rows), so
// Q, K, V, dO are each loaded once and used for every product they appear in.  P / dS are written by the softmax
// threads straight into the same swizzled layout and consumed by the next tcgen05.mma.
//   fwd      : CTA per (q block, head, batch), loop over kv blocks; S and P.V products in TMEM, running max/sum
//              and the output accumulator in registers.  2 CTAs/SM overlap one CTA's softmax with the other's MMAs.
//   bwd dK/dV: CTA per (kv block, head, batch), loop over q blocks >= kv block; dV += P^T dO, dK += dS^T Q in TMEM.
//   bwd dQ   : CTA per (q block, head, batch), loop over kv blocks <= q block; dQ += dS K in TMEM (no atomics).
#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"

namespace db200 {

constexpr float LOG2E = 1.4426950408889634f;

// MUFU.EX2 directly (exp2f() adds range handling the softmax does not need: arguments are <= 0 or the result is
// masked / multiplied by 0).  ex2.approx(-inf) = +0.
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// Load a [ROWS][DH] tile (DH/64 swizzled sub-tiles of ROWS*128 bytes) from a rank-4 map {dh, chan, seq, batch}.
template <int DH>
__device__ __forceinline__ void tma_prefetch_tile(const CUtensorMap* tm, int chan, int row0, int b) {
#pragma unroll
  for (int t = 0; t < DH / 64; ++t) tma_prefetch_4d(tm, 64 * t, chan, row0, b);
}

template <int DH>
__device__ __forceinline__ void tma_load_tile(uint32_t dst, uint32_t rows_bytes, const CUtensorMap* tm, uint32_t bar,
                                              int chan, int row0, int b) {
#pragma unroll
  for (int t = 0; t < DH / 64; ++t) tma_load_4d(dst + t * rows_bytes, tm, bar, 64 * t, chan, row0, b);
}

// ------------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------------
template <int DH>
struct FwdCfg {
  static constexpr int BNK = (DH == 128) ? 64 : 128;  // keys per inner block (keeps smem <= ~80 KiB -> 2 CTAs/SM)
  static constexpr uint32_t Q_BYTES = 128 * DH * 2;
  static constexpr uint32_t KV_BYTES = BNK * DH * 2;
  static constexpr uint32_t P_BYTES = 128 * BNK * 2;
  static constexpr size_t SMEM = 1024 + Q_BYTES + 2 * KV_BYTES + P_BYTES + 64 + 2 * 128 * 4 + 64;
};

template <int DH>
__global__ void __launch_bounds__(256, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                bf16* __restrict__ out, float* __restrict__ lse_out, int S, int H, float scale) {
  using C = FwdCfg<DH>;
  constexpr int BNK = C::BNK;
  constexpr int HC = BNK / 2;  // key columns of S handled by one thread (two threads share a query row)
  constexpr int HD = DH / 2;   // output columns accumulated by one thread
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t sQ = base, sK = sQ + C::Q_BYTES, sV = sK + C::KV_BYTES, sP = sV + C::KV_BYTES;
  const uint32_t bars = sP + C::P_BYTES;
  const uint32_t bar_q = bars, bar_k = bars + 8, bar_v = bars + 16, bar_s = bars + 24, bar_o = bars + 32;
  const uint32_t tmem_slot = bars + 40;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  float* xch = reinterpret_cast<float*>(smem_raw + (bars + 64 - raw));  // [2][128] row-max exchange

  const int tid = threadIdx.x, warp = tid >> 5;
  const int rowi = tid & 127, half = tid >> 7;  // TMEM lane (query row in the tile), column half
  const int qb = gridDim.x - 1 - blockIdx.x;    // heavy (late) query blocks first
  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = qb * 128;
  const int kv_len = min(S, q0 + 128);
  const int n_kv = (kv_len + BNK - 1) / BNK;

  if (tid == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    mbar_init(bar_q, 1); mbar_init(bar_k, 1); mbar_init(bar_v, 1); mbar_init(bar_s, 1); mbar_init(bar_o, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const uint32_t tS = tmem, tO = tmem + 128;
  const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;

  if (tid == 0) {
    mbar_expect_tx(bar_q, C::Q_BYTES);
    tma_load_tile<DH>(sQ, 128 * 128, &tmQ, bar_q, /*chan=*/0 * H + h, q0, b);
    mbar_expect_tx(bar_k, C::KV_BYTES);
    tma_load_tile<DH>(sK, BNK * 128, &tmKV, bar_k, 1 * H + h, 0, b);
    mbar_expect_tx(bar_v, C::KV_BYTES);
    tma_load_tile<DH>(sV, BNK * 128, &tmKV, bar_v, 2 * H + h, 0, b);
  }

  constexpr uint32_t idesc_s = umma_idesc_bf16(128, BNK, 0, 0);  // S = Q K^T : both K-major (K = dh)
  constexpr uint32_t idesc_o = umma_idesc_bf16(128, DH, 0, 1);   // O = P V   : A K-major (K = keys), B MN-major

  const int qi = q0 + rowi;  // this thread's query index
  const float c1 = scale * LOG2E;
  float m_run = -INFINITY, l_run = 0.f;  // l_run: partial sum over this thread's key columns
  float o_acc[HD];
#pragma unroll
  for (int e = 0; e < HD; ++e) o_acc[e] = 0.f;

  mbar_wait(bar_q, 0);

  for (int j = 0; j < n_kv; ++j) {
    const uint32_t ph = j & 1;
    if (tid == 0) {
      mbar_wait(bar_k, ph);
      tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < DH / 16; ++kk) {
        const uint64_t ad = umma_smem_desc_sw128(sQ + (kk / 4) * (128 * 128) + (kk % 4) * 32, 0, 1024);
        const uint64_t bd = umma_smem_desc_sw128(sK + (kk / 4) * (BNK * 128) + (kk % 4) * 32, 0, 1024);
        umma_bf16_ss(tS, ad, bd, idesc_s, kk > 0);
      }
      umma_commit(bar_s);
    }
    mbar_wait(bar_s, ph);
    tc_fence_after();
    if (tid == 0 && j + 1 < n_kv) {  // K buffer is free: prefetch the next key block under the softmax
      mbar_expect_tx(bar_k, C::KV_BYTES);
      tma_load_tile<DH>(sK, BNK * 128, &tmKV, bar_k, 1 * H + h, (j + 1) * BNK, b);
    }
    // ---- softmax: this thread's half of the row, kept in registers
    const int k0 = j * BNK + half * HC;
    const bool need_mask = (j * BNK + BNK - 1) > q0;
    float sv[HC];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < HC / 32; ++c) {
      uint32_t r[32];
      tmem_ld_x32(tS + lane_off + half * HC + c * 32, r);
      tmem_ld_wait();
      if (need_mask && (k0 + c * 32 + 31) > qi) {  // chunk crosses the diagonal
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float s = __uint_as_float(r[i]);
          if ((k0 + c * 32 + i) > qi) s = -INFINITY;
          sv[c * 32 + i] = s;
          mx = fmaxf(mx, s);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float s = __uint_as_float(r[i]);
          sv[c * 32 + i] = s;
          mx = fmaxf(mx, s);
        }
      }
    }
    xch[half * 128 + rowi] = mx;
    tc_fence_before();
    __syncthreads();  // (1) both halves' maxima visible; every thread has finished reading S from TMEM
    const float m_new = fmaxf(m_run, fmaxf(mx, xch[(half ^ 1) * 128 + rowi]));  // finite: key 0 is always visible
    const float alpha = ex2((m_run - m_new) * c1);
    const float mc = m_new * c1;
    float lsum = 0.f;
#pragma unroll
    for (int c = 0; c < HC / 32; ++c) {
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const float p0 = ex2(fmaf(sv[c * 32 + i], c1, -mc)), p1 = ex2(fmaf(sv[c * 32 + i + 1], c1, -mc));  // ex2(-inf) = 0
        lsum += p0 + p1;
        pk[i >> 1] = pack_bf16x2(p0, p1);
      }
      const int col = half * HC + c * 32;  // column inside the [128][BNK] P tile
      const uint32_t sub = sP + (col / 64) * (128 * 128);
      const int cc = col % 64;
#pragma unroll
      for (int g = 0; g < 4; ++g)
        st_shared_v4(sub + sw128_offset(rowi, cc + g * 8), pk[g * 4], pk[g * 4 + 1], pk[g * 4 + 2], pk[g * 4 + 3]);
    }
    l_run = l_run * alpha + lsum;
    m_run = m_new;
    fence_proxy_async_smem();  // P (generic-proxy stores) -> visible to tcgen05.mma (async proxy)
    __syncthreads();           // (2) P complete
    // ---- O_j = P V_j  (fresh accumulator; the running output lives in registers)
    if (tid == 0) {
      mbar_wait(bar_v, ph);
      tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < BNK / 16; ++kk) {
        const uint64_t ad = umma_smem_desc_sw128(sP + (kk / 4) * (128 * 128) + (kk % 4) * 32, 0, 1024);
        const uint64_t bd = umma_smem_desc_sw128(sV + kk * 2048, BNK * 128, 1024);
        umma_bf16_ss(tO, ad, bd, idesc_o, kk > 0);
      }
      umma_commit(bar_o);
    }
#pragma unroll
    for (int e = 0; e < HD; ++e) o_acc[e] *= alpha;
    mbar_wait(bar_o, ph);
    tc_fence_after();
    if (tid == 0 && j + 1 < n_kv) {
      mbar_expect_tx(bar_v, C::KV_BYTES);
      tma_load_tile<DH>(sV, BNK * 128, &tmKV, bar_v, 2 * H + h, (j + 1) * BNK, b);
    }
#pragma unroll
    for (int c = 0; c < HD / 32; ++c) {
      uint32_t r[32];
      tmem_ld_x32(tO + lane_off + half * HD + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o_acc[c * 32 + i] += __uint_as_float(r[i]);
    }
    tc_fence_before();
    __syncthreads();  // (3) O consumed before the next iteration's MMAs overwrite S / O; xch reusable
    tc_fence_after();
  }

  // combine the two halves' partial row sums
  xch[half * 128 + rowi] = l_run;
  __syncthreads();
  const float l_tot = l_run + xch[(half ^ 1) * 128 + rowi];
  if (qi < S) {
    const float inv = 1.f / l_tot;
    bf16* op = out + (((long long)b * S + qi) * H + h) * DH + half * HD;
#pragma unroll
    for (int e = 0; e < HD; e += 8) {
      uint4 q;
      q.x = pack_bf16x2(o_acc[e] * inv, o_acc[e + 1] * inv);
      q.y = pack_bf16x2(o_acc[e + 2] * inv, o_acc[e + 3] * inv);
      q.z = pack_bf16x2(o_acc[e + 4] * inv, o_acc[e + 5] * inv);
      q.w = pack_bf16x2(o_acc[e + 6] * inv, o_acc[e + 7] * inv);
      *reinterpret_cast<uint4*>(op + e) = q;
    }
    if (half == 0) lse_out[((long long)b * H + h) * S + qi] = m_run * scale + logf(l_tot);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// forward, software-pipelined variant (EXPERIMENTAL: selected with DB200_ATTN_V2=1, not yet validated on hardware —
// the default path is attn_fwd_kernel above).  Differences to attn_fwd_kernel:
//   * S is double-buffered in TMEM: S_{j+1} = Q K_{j+1}^T is issued as soon as every thread has read S_j, so the
//     tensor pipe computes the next logits while the CTA exponentiates the current ones;
//   * the output accumulates in TMEM across key blocks (P V with the accumulate flag) instead of being pulled into
//     registers every block; rows are rescaled in TMEM (tcgen05.ld -> scale -> tcgen05.st) only when their running
//     maximum moved by more than 2^8 since the scale they use ("lazy rescaling"), which is rare after the first blocks;
//   * V is double-buffered in shared memory; K and P stay single-buffered (their consumers are a full softmax behind).
// 64 keys per block for both head sizes: 2 x 64 (S) + DH (O) <= 256 TMEM columns and <= 96 KiB smem -> 2 CTAs / SM.
// ------------------------------------------------------------------------------------------------------------------
template <int DH>
struct Fwd2Cfg {
  static constexpr int BNK = 64;
  static constexpr uint32_t Q_BYTES = 128 * DH * 2;
  static constexpr uint32_t KV_BYTES = BNK * DH * 2;
  static constexpr uint32_t P_BYTES = 128 * BNK * 2;
  static constexpr size_t SMEM = 1024 + Q_BYTES + 3 * KV_BYTES + P_BYTES + 64 + 2 * 128 * 4 + 64;
};

template <int DH>
__global__ void __launch_bounds__(256, 2)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                 bf16* __restrict__ out, float* __restrict__ lse_out, int S, int H, float scale) {
  using C = Fwd2Cfg<DH>;
  constexpr int BNK = C::BNK;
  constexpr int HC = BNK / 2;  // 32 key columns of S per thread
  constexpr int HD = DH / 2;   // output columns per thread
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t sQ = base, sK = sQ + C::Q_BYTES, sV = sK + C::KV_BYTES, sP = sV + 2 * C::KV_BYTES;
  const uint32_t bars = sP + C::P_BYTES;
  const uint32_t bar_q = bars, bar_k = bars + 8, bar_v = bars + 16 /*2*/, bar_s = bars + 32 /*2*/, bar_o = bars + 48;
  const uint32_t tmem_slot = bars + 56;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  float* xch = reinterpret_cast<float*>(smem_raw + (bars + 64 - raw));  // [2][128] row-max exchange

  const int tid = threadIdx.x, warp = tid >> 5;
  const int rowi = tid & 127, half = tid >> 7;
  const int qb = gridDim.x - 1 - blockIdx.x;  // heavy (late) query blocks first
  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = qb * 128;
  const int kv_len = min(S, q0 + 128);
  const int n_kv = (kv_len + BNK - 1) / BNK;

  if (tid == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    mbar_init(bar_q, 1); mbar_init(bar_k, 1); mbar_init(bar_v, 1); mbar_init(bar_v + 8, 1);
    mbar_init(bar_s, 1); mbar_init(bar_s + 8, 1); mbar_init(bar_o, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const uint32_t tS = tmem /* 2 x 64 columns */, tO = tmem + 128;
  const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;

  constexpr uint32_t idesc_s = umma_idesc_bf16(128, BNK, 0, 0);
  constexpr uint32_t idesc_o = umma_idesc_bf16(128, DH, 0, 1);

  auto issue_s = [&](int buf) {  // S = Q K^T into tS[buf]; K from sK
#pragma unroll
    for (int kk = 0; kk < DH / 16; ++kk) {
      const uint64_t ad = umma_smem_desc_sw128(sQ + (kk / 4) * (128 * 128) + (kk % 4) * 32, 0, 1024);
      const uint64_t bd = umma_smem_desc_sw128(sK + (kk / 4) * (BNK * 128) + (kk % 4) * 32, 0, 1024);
      umma_bf16_ss(tS + buf * BNK, ad, bd, idesc_s, kk > 0);
    }
    umma_commit(bar_s + 8 * buf);
  };

  if (tid == 0) {
    mbar_expect_tx(bar_q, C::Q_BYTES);
    tma_load_tile<DH>(sQ, 128 * 128, &tmQ, bar_q, 0 * H + h, q0, b);
    mbar_expect_tx(bar_k, C::KV_BYTES);
    tma_load_tile<DH>(sK, BNK * 128, &tmKV, bar_k, 1 * H + h, 0, b);
    mbar_expect_tx(bar_v, C::KV_BYTES);
    tma_load_tile<DH>(sV, BNK * 128, &tmKV, bar_v, 2 * H + h, 0, b);
    mbar_wait(bar_q, 0);
    mbar_wait(bar_k, 0);
    tc_fence_after();
    issue_s(0);
  }

  const int qi = q0 + rowi;
  const float c1 = scale * LOG2E;
  float m_run = -INFINITY;   // true running row maximum (of the raw logits)
  float m_used = -INFINITY;  // the maximum the accumulated P / O / l are scaled by
  float l_run = 0.f;         // partial row sum over this thread's key columns, relative to m_used

  for (int j = 0; j < n_kv; ++j) {
    const int sb = j & 1;
    // ---- (A) S_j is in tS[sb]
    mbar_wait(bar_s + 8 * sb, (j >> 1) & 1);
    tc_fence_after();
    if (tid == 0 && j + 1 < n_kv) {  // sK is free: S_j has consumed K_j
      mbar_expect_tx(bar_k, C::KV_BYTES);
      tma_load_tile<DH>(sK, BNK * 128, &tmKV, bar_k, 1 * H + h, (j + 1) * BNK, b);
    }
    // ---- (B) this thread's 32 logits and their maximum
    const int k0 = j * BNK + half * HC;
    const bool need_mask = (j * BNK + BNK - 1) > q0;
    float sv[HC];
    float mx = -INFINITY;
    {
      uint32_t r[32];
      tmem_ld_x32(tS + sb * BNK + lane_off + half * HC, r);
      tmem_ld_wait();
      if (need_mask && (k0 + 31) > qi) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float s = __uint_as_float(r[i]);
          if ((k0 + i) > qi) s = -INFINITY;
          sv[i] = s;
          mx = fmaxf(mx, s);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          sv[i] = __uint_as_float(r[i]);
          mx = fmaxf(mx, sv[i]);
        }
      }
    }
    xch[half * 128 + rowi] = mx;
    tc_fence_before();
    __syncthreads();  // (1) maxima exchanged; every thread has read tS[sb]
    // ---- (C) next block's logits go to the other S buffer while this block is exponentiated
    if (tid == 0 && j + 1 < n_kv) {
      mbar_wait(bar_k, (j + 1) & 1);
      tc_fence_after();
      issue_s(sb ^ 1);
    }
    const float m_new = fmaxf(m_run, fmaxf(mx, xch[(half ^ 1) * 128 + rowi]));  // finite for in-range rows: key 0 visible
    // ---- (D)+(E) the previous P V must be finished before P / O are touched
    if (j > 0) {
      mbar_wait(bar_o, (j - 1) & 1);
      tc_fence_after();
    }
    if (tid == 0 && j + 1 < n_kv) {  // sV[sb ^ 1] held V_{j-1}: free now
      mbar_expect_tx(bar_v + 8 * (sb ^ 1), C::KV_BYTES);
      tma_load_tile<DH>(sV + (sb ^ 1) * C::KV_BYTES, BNK * 128, &tmKV, bar_v + 8 * (sb ^ 1), 2 * H + h, (j + 1) * BNK, b);
    }
    if (j == 0) {
      m_used = m_new;
    } else {
      const bool need = (m_new - m_used) * c1 > 8.f;  // identical for the two threads of a row
      if (__any_sync(0xffffffffu, need)) {             // tcgen05.ld / st are warp-collective
        const float alpha = need ? ex2((m_used - m_new) * c1) : 1.f;
#pragma unroll
        for (int c = 0; c < HD / 32; ++c) {
          uint32_t r[32];
          tmem_ld_x32(tO + lane_off + half * HD + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
          tmem_st_x32(tO + lane_off + half * HD + c * 32, r);
        }
        tmem_st_wait();
        if (need) {
          l_run *= alpha;
          m_used = m_new;
        }
      }
    }
    // ---- (F) P_j = 2^(c1 (s - m_used)) -> bf16 -> swizzled smem tile
    const float mc = m_used * c1;
    float lsum = 0.f;
    {
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const float p0 = ex2(fmaf(sv[i], c1, -mc)), p1 = ex2(fmaf(sv[i + 1], c1, -mc));
        lsum += p0 + p1;
        pk[i >> 1] = pack_bf16x2(p0, p1);
      }
      const int col = half * HC;  // column inside the [128][64] P tile (one 64-wide sub-tile)
#pragma unroll
      for (int g = 0; g < 4; ++g)
        st_shared_v4(sP + sw128_offset(rowi, col + g * 8), pk[g * 4], pk[g * 4 + 1], pk[g * 4 + 2], pk[g * 4 + 3]);
    }
    l_run += lsum;
    m_run = m_new;
    fence_proxy_async_smem();
    tc_fence_before();  // orders the tcgen05.st of a rescale before the MMA issued after the barrier
    __syncthreads();    // (2) P complete, O rescaled
    // ---- (G) O (+)= P_j V_j
    if (tid == 0) {
      mbar_wait(bar_v + 8 * sb, (j >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < BNK / 16; ++kk) {
        const uint64_t ad = umma_smem_desc_sw128(sP + (kk % 4) * 32, 0, 1024);
        const uint64_t bd = umma_smem_desc_sw128(sV + sb * C::KV_BYTES + kk * 2048, BNK * 128, 1024);
        umma_bf16_ss(tO, ad, bd, idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
      }
      umma_commit(bar_o);
    }
  }

  // ---- epilogue: O / l
  mbar_wait(bar_o, (n_kv - 1) & 1);
  tc_fence_after();
  xch[half * 128 + rowi] = l_run;
  __syncthreads();
  const float l_tot = l_run + xch[(half ^ 1) * 128 + rowi];
  const float inv = 1.f / l_tot;
  bf16* op = out + (((long long)b * S + qi) * H + h) * DH + half * HD;
#pragma unroll
  for (int c = 0; c < HD / 32; ++c) {
    uint32_t r[32];
    tmem_ld_x32(tO + lane_off + half * HD + c * 32, r);
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
  if (qi < S && half == 0) lse_out[((long long)b * H + h) * S + qi] = m_used * scale + logf(l_tot);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// backward: delta = rowsum(dO * O)
// ------------------------------------------------------------------------------------------------------------------
__global__ void attn_delta_kernel(const bf16* __restrict__ o, const bf16* __restrict__ dout,
                                  float* __restrict__ delta, int B, int S, int H, int dh) {
  // one warp per (b, s, h) row of dh elements
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long rows = (long long)B * S * H;
  if (row >= rows) return;
  const bf16* op = o + row * dh;
  const bf16* dp = dout + row * dh;
  float acc = 0.f;
  for (int e = lane * 2; e < dh; e += 64) {
    const float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(op + e));
    const float2 d = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dp + e));
    acc += a.x * d.x + a.y * d.y;
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    const int hh = (int)(row % H);
    const long long bs = row / H;
    const int s = (int)(bs % S);
    const int b = (int)(bs / S);
    delta[((long long)b * H + hh) * S + s] = acc;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// backward shared piece: from S and dP in TMEM build P and dS (bf16, swizzled [128 q rows][128 keys]) in smem.
//   p  = exp(scale*s - lse)            (0 where key > query, or the query row is out of range)
//   ds = p * (dp - delta) * scale
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bwd_make_p_ds(uint32_t tS, uint32_t tdP, uint32_t lane_off, uint32_t sP, uint32_t sdS,
                                              bool write_p, int rowi, int half, int qi, int k0, bool need_mask,
                                              bool row_ok, float lse_l2, float delta, float c1, float scale) {
  // two threads share a query row: `half` selects key columns [64*half, 64*half + 64) = sub-tile `half`
#pragma unroll 1
  for (int c = 0; c < 2; ++c) {
    const int col0 = half * 64 + c * 32;
    uint32_t rs[32], rd[32];
    tmem_ld_x32(tS + lane_off + col0, rs);
    tmem_ld_x32(tdP + lane_off + col0, rd);
    tmem_ld_wait();
    uint32_t pk[16], dk[16];
    if (!row_ok) {  // out-of-range query row: contributes nothing
#pragma unroll
      for (int i = 0; i < 16; ++i) pk[i] = dk[i] = 0u;
    } else if (need_mask && (k0 + col0 + 31) > qi) {  // chunk crosses the diagonal: per-element causal mask
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        float p0 = ex2(fmaf(__uint_as_float(rs[i]), c1, -lse_l2));
        float p1 = ex2(fmaf(__uint_as_float(rs[i + 1]), c1, -lse_l2));
        if ((k0 + col0 + i) > qi) p0 = 0.f;
        if ((k0 + col0 + i + 1) > qi) p1 = 0.f;
        pk[i >> 1] = pack_bf16x2(p0, p1);
        dk[i >> 1] = pack_bf16x2((p0 * scale) * (__uint_as_float(rd[i]) - delta),
                                 (p1 * scale) * (__uint_as_float(rd[i + 1]) - delta));
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const float p0 = ex2(fmaf(__uint_as_float(rs[i]), c1, -lse_l2));
        const float p1 = ex2(fmaf(__uint_as_float(rs[i + 1]), c1, -lse_l2));
        pk[i >> 1] = pack_bf16x2(p0, p1);
        dk[i >> 1] = pack_bf16x2((p0 * scale) * (__uint_as_float(rd[i]) - delta),
                                 (p1 * scale) * (__uint_as_float(rd[i + 1]) - delta));
      }
    }
    const uint32_t sub_off = half * (128 * 128);
    const int cc = c * 32;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const uint32_t off = sub_off + sw128_offset(rowi, cc + g * 8);
      if (write_p) st_shared_v4(sP + off, pk[g * 4], pk[g * 4 + 1], pk[g * 4 + 2], pk[g * 4 + 3]);
      st_shared_v4(sdS + off, dk[g * 4], dk[g * 4 + 1], dk[g * 4 + 2], dk[g * 4 + 3]);
    }
  }
}

constexpr uint32_t T128 = 128 * 128;  // bytes of one [128 rows][64] sub-tile

// ------------------------------------------------------------------------------------------------------------------
// backward dK / dV
// ------------------------------------------------------------------------------------------------------------------
template <int DH>
struct BwdCfg {
  static constexpr uint32_t TILE = 128 * DH * 2;   // a [128][DH] operand tile
  static constexpr uint32_t PT = 128 * 128 * 2;    // P / dS tile
  static constexpr size_t SMEM_DKDV = 1024 + 4 * TILE + 2 * PT + 128;
  static constexpr size_t SMEM_DQ = 1024 + 4 * TILE + PT + 128;
};

template <int DH>
__global__ void __launch_bounds__(256, 1)
attn_bwd_dkdv_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                     const float* __restrict__ lse, const float* __restrict__ delta, bf16* __restrict__ dqkv, int S,
                     int H, float scale) {
  using C = BwdCfg<DH>;
  constexpr int NT = DH / 64;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t sK = base, sV = sK + C::TILE, sQ = sV + C::TILE, sdO = sQ + C::TILE, sP = sdO + C::TILE,
                 sdS = sP + C::PT;
  const uint32_t bars = sdS + C::PT;
  const uint32_t bar_kv = bars, bar_q = bars + 8, bar_do = bars + 16, bar_a = bars + 24, bar_b = bars + 32;
  const uint32_t tmem_slot = bars + 40;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  const int tid = threadIdx.x, warp = tid >> 5;
  const int rowi = tid & 127, half = tid >> 7;
  const int jb = blockIdx.x;  // kv block (block 0 has the most work and is scheduled first)
  const int h = blockIdx.y, b = blockIdx.z;
  const int k0 = jb * 128;
  const int n_q = (S + 127) / 128;

  if (tid == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    mbar_init(bar_kv, 1); mbar_init(bar_q, 1); mbar_init(bar_do, 1); mbar_init(bar_a, 1); mbar_init(bar_b, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const uint32_t tS = tmem, tdP = tmem + 128, tdV = tmem + 256, tdK = tmem + 256 + DH;
  const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;

  if (tid == 0) {
    mbar_expect_tx(bar_kv, 2 * C::TILE);
    tma_load_tile<DH>(sK, T128, &tmQKV, bar_kv, 1 * H + h, k0, b);
    tma_load_tile<DH>(sV, T128, &tmQKV, bar_kv, 2 * H + h, k0, b);
    mbar_expect_tx(bar_q, C::TILE);
    tma_load_tile<DH>(sQ, T128, &tmQKV, bar_q, 0 * H + h, jb * 128, b);
    mbar_expect_tx(bar_do, C::TILE);
    tma_load_tile<DH>(sdO, T128, &tmDO, bar_do, h, jb * 128, b);
  }
  constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);  // S = Q K^T, dP = dO V^T
  constexpr uint32_t idesc_g = umma_idesc_bf16(128, DH, 1, 1);   // dV = P^T dO, dK = dS^T Q  (both MN-major)
  const float c1 = scale * LOG2E;

  mbar_wait(bar_kv, 0);
  int it = 0;
  for (int ib = jb; ib < n_q; ++ib, ++it) {
    const uint32_t ph = it & 1;
    const int q0 = ib * 128;
    const int qi = q0 + rowi;
    const bool row_ok = qi < S;
    float lse_l2 = 0.f, dl = 0.f;
    if (row_ok) {
      lse_l2 = lse[((long long)b * H + h) * S + qi] * LOG2E;
      dl = delta[((long long)b * H + h) * S + qi];
    }
    if (tid == 0) {
      mbar_wait(bar_q, ph);
      mbar_wait(bar_do, ph);
      tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < DH / 16; ++kk) {
        const uint32_t o = (kk / 4) * T128 + (kk % 4) * 32;
        umma_bf16_ss(tS, umma_smem_desc_sw128(sQ + o, 0, 1024), umma_smem_desc_sw128(sK + o, 0, 1024), idesc_s,
                     kk > 0);
      }
#pragma unroll
      for (int kk = 0; kk < DH / 16; ++kk) {
        const uint32_t o = (kk / 4) * T128 + (kk % 4) * 32;
        umma_bf16_ss(tdP, umma_smem_desc_sw128(sdO + o, 0, 1024), umma_smem_desc_sw128(sV + o, 0, 1024), idesc_s,
                     kk > 0);
      }
      umma_commit(bar_a);
      if (ib + 1 < n_q) {  // the smem buffers are single: at least pull the next Q / dO tiles into L2 now
        tma_prefetch_tile<DH>(&tmQKV, 0 * H + h, (ib + 1) * 128, b);
        tma_prefetch_tile<DH>(&tmDO, h, (ib + 1) * 128, b);
      }
    }
    mbar_wait(bar_a, ph);
    tc_fence_after();
    bwd_make_p_ds(tS, tdP, lane_off, sP, sdS, true, rowi, half, qi, k0, /*need_mask=*/ib == jb, row_ok, lse_l2, dl,
                  c1, scale);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      // dV += P^T dO ; dK += dS^T Q      (K dimension = the 128 query rows of this block)
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        umma_bf16_ss(tdV, umma_smem_desc_sw128(sP + kk * 2048, T128, 1024),
                     umma_smem_desc_sw128(sdO + kk * 2048, T128, 1024), idesc_g, (it > 0 || kk > 0));
      }
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        umma_bf16_ss(tdK, umma_smem_desc_sw128(sdS + kk * 2048, T128, 1024),
                     umma_smem_desc_sw128(sQ + kk * 2048, T128, 1024), idesc_g, (it > 0 || kk > 0));
      }
      umma_commit(bar_b);
    }
    mbar_wait(bar_b, ph);
    tc_fence_after();
    if (tid == 0 && ib + 1 < n_q) {  // Q / dO / P / dS buffers are free again
      mbar_expect_tx(bar_q, C::TILE);
      tma_load_tile<DH>(sQ, T128, &tmQKV, bar_q, 0 * H + h, (ib + 1) * 128, b);
      mbar_expect_tx(bar_do, C::TILE);
      tma_load_tile<DH>(sdO, T128, &tmDO, bar_do, h, (ib + 1) * 128, b);
    }
    (void)NT;
  }
  // ---- write dK, dV rows (thread pair = key row; each thread writes DH/2 columns)
  const int ki = k0 + rowi;
#pragma unroll 1
  for (int which = 0; which < 2; ++which) {
    const uint32_t tsrc = which == 0 ? tdK : tdV;
    bf16* dst = dqkv + ((((long long)b * S + ki) * 3 + (which == 0 ? 1 : 2)) * H + h) * DH + half * (DH / 2);
#pragma unroll 1
    for (int c = 0; c < DH / 64; ++c) {
      uint32_t r[32];
      tmem_ld_x32(tsrc + lane_off + half * (DH / 2) + c * 32, r);
      tmem_ld_wait();
      if (ki < S) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 q;
          q.x = pack_bf16x2(__uint_as_float(r[g * 8 + 0]), __uint_as_float(r[g * 8 + 1]));
          q.y = pack_bf16x2(__uint_as_float(r[g * 8 + 2]), __uint_as_float(r[g * 8 + 3]));
          q.z = pack_bf16x2(__uint_as_float(r[g * 8 + 4]), __uint_as_float(r[g * 8 + 5]));
          q.w = pack_bf16x2(__uint_as_float(r[g * 8 + 6]), __uint_as_float(r[g * 8 + 7]));
          *reinterpret_cast<uint4*>(dst + c * 32 + g * 8) = q;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// backward dQ
// ------------------------------------------------------------------------------------------------------------------
template <int DH>
__global__ void __launch_bounds__(256, 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                   const float* __restrict__ lse, const float* __restrict__ delta, bf16* __restrict__ dqkv, int S,
                   int H, float scale) {
  using C = BwdCfg<DH>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t sQ = base, sdO = sQ + C::TILE, sK = sdO + C::TILE, sV = sK + C::TILE, sdS = sV + C::TILE;
  const uint32_t bars = sdS + C::PT;
  const uint32_t bar_q = bars, bar_k = bars + 8, bar_v = bars + 16, bar_a = bars + 24, bar_b = bars + 32;
  const uint32_t tmem_slot = bars + 40;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  const int tid = threadIdx.x, warp = tid >> 5;
  const int rowi = tid & 127, half = tid >> 7;
  const int ib = gridDim.x - 1 - blockIdx.x;  // heavy query blocks first
  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = ib * 128;

  if (tid == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    mbar_init(bar_q, 1); mbar_init(bar_k, 1); mbar_init(bar_v, 1); mbar_init(bar_a, 1); mbar_init(bar_b, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const uint32_t tS = tmem, tdP = tmem + 128, tdQ = tmem + 256;
  const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;

  if (tid == 0) {
    mbar_expect_tx(bar_q, 2 * C::TILE);
    tma_load_tile<DH>(sQ, T128, &tmQKV, bar_q, 0 * H + h, q0, b);
    tma_load_tile<DH>(sdO, T128, &tmDO, bar_q, h, q0, b);
    mbar_expect_tx(bar_k, C::TILE);
    tma_load_tile<DH>(sK, T128, &tmQKV, bar_k, 1 * H + h, 0, b);
    mbar_expect_tx(bar_v, C::TILE);
    tma_load_tile<DH>(sV, T128, &tmQKV, bar_v, 2 * H + h, 0, b);
  }
  constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
  constexpr uint32_t idesc_q = umma_idesc_bf16(128, DH, 0, 1);  // dQ = dS K : A K-major (K = keys), B MN-major
  const float c1 = scale * LOG2E;
  const int qi = q0 + rowi;
  const bool row_ok = qi < S;
  float lse_l2 = 0.f, dl = 0.f;
  if (row_ok) {
    lse_l2 = lse[((long long)b * H + h) * S + qi] * LOG2E;
    dl = delta[((long long)b * H + h) * S + qi];
  }
  mbar_wait(bar_q, 0);
  for (int jb = 0; jb <= ib; ++jb) {
    const uint32_t ph = jb & 1;
    if (tid == 0) {
      mbar_wait(bar_k, ph);
      mbar_wait(bar_v, ph);
      tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < DH / 16; ++kk) {
        const uint32_t o = (kk / 4) * T128 + (kk % 4) * 32;
        umma_bf16_ss(tS, umma_smem_desc_sw128(sQ + o, 0, 1024), umma_smem_desc_sw128(sK + o, 0, 1024), idesc_s,
                     kk > 0);
      }
#pragma unroll
      for (int kk = 0; kk < DH / 16; ++kk) {
        const uint32_t o = (kk / 4) * T128 + (kk % 4) * 32;
        umma_bf16_ss(tdP, umma_smem_desc_sw128(sdO + o, 0, 1024), umma_smem_desc_sw128(sV + o, 0, 1024), idesc_s,
                     kk > 0);
      }
      umma_commit(bar_a);
    }
    mbar_wait(bar_a, ph);
    tc_fence_after();
    if (tid == 0 && jb + 1 <= ib) {  // V is no longer needed by this iteration
      mbar_expect_tx(bar_v, C::TILE);
      tma_load_tile<DH>(sV, T128, &tmQKV, bar_v, 2 * H + h, (jb + 1) * 128, b);
    }
    bwd_make_p_ds(tS, tdP, lane_off, 0, sdS, false, rowi, half, qi, jb * 128, /*need_mask=*/jb == ib, row_ok, lse_l2,
                  dl, c1, scale);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {  // K dimension = the 128 keys of this block
        umma_bf16_ss(tdQ, umma_smem_desc_sw128(sdS + (kk / 4) * T128 + (kk % 4) * 32, 0, 1024),
                     umma_smem_desc_sw128(sK + kk * 2048, T128, 1024), idesc_q, (jb > 0 || kk > 0));
      }
      umma_commit(bar_b);
    }
    mbar_wait(bar_b, ph);
    tc_fence_after();
    if (tid == 0 && jb + 1 <= ib) {
      mbar_expect_tx(bar_k, C::TILE);
      tma_load_tile<DH>(sK, T128, &tmQKV, bar_k, 1 * H + h, (jb + 1) * 128, b);
    }
  }
  bf16* dst = dqkv + ((((long long)b * S + qi) * 3 + 0) * H + h) * DH + half * (DH / 2);
#pragma unroll 1
  for (int c = 0; c < DH / 64; ++c) {
    uint32_t r[32];
    tmem_ld_x32(tdQ + lane_off + half * (DH / 2) + c * 32, r);
    tmem_ld_wait();
    if (row_ok) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 q;
        q.x = pack_bf16x2(__uint_as_float(r[g * 8 + 0]), __uint_as_float(r[g * 8 + 1]));
        q.y = pack_bf16x2(__uint_as_float(r[g * 8 + 2]), __uint_as_float(r[g * 8 + 3]));
        q.z = pack_bf16x2(__uint_as_float(r[g * 8 + 4]), __uint_as_float(r[g * 8 + 5]));
        q.w = pack_bf16x2(__uint_as_float(r[g * 8 + 6]), __uint_as_float(r[g * 8 + 7]));
        *reinterpret_cast<uint4*>(dst + c * 32 + g * 8) = q;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// backward, software-pipelined variants (EXPERIMENTAL: DB200_ATTN_V2 bit 1, not yet validated on hardware).
// Same two-kernel scheme (no atomics), but each inner block is 64 wide, S / dP are double-buffered in TMEM so the next
// block's two logit MMAs run while the CTA computes P / dS of the current one, and the operand tiles that are needed
// both early (logit MMAs) and late (gradient MMAs) live in 3-deep shared-memory rings fed by TMA two blocks ahead.
//   dQ    : CTA = 128 queries; loop over 64-key blocks.  S = Q K^T, dP = dO V^T (128 x 64), dS -> smem, dQ += dS K.
//   dK/dV : CTA = 128 keys; loop over 64-query blocks, in the TRANSPOSED orientation S^T = K Q^T, dP^T = V dO^T
//           (TMEM lanes = keys), so P^T and dS^T are written K-major ([keys][64 queries]) and dV += P^T dO,
//           dK += dS^T Q need no MN-major A operand; lse / delta of the 64 queries are staged in shared memory.
// ------------------------------------------------------------------------------------------------------------------
template <int DH>
struct Bwd2Cfg {
  static constexpr uint32_t TILE = 128 * DH * 2;  // [128][DH]
  static constexpr uint32_t HT = 64 * DH * 2;     // [64][DH]
  static constexpr uint32_t ST = 128 * 64 * 2;    // [128][64] P^T / dS^T / dS tile
  static constexpr size_t SMEM_DQ = 1024 + 2 * TILE + 5 * HT + ST + 256;
  static constexpr size_t SMEM_DKDV = 1024 + 2 * TILE + 6 * HT + 2 * ST + 256 + 2 * 2 * 64 * 4;
};
constexpr uint32_t T64 = 64 * 128;  // bytes of one [64 rows][64] sub-tile

template <int DH>
__global__ void __launch_bounds__(256, 1)
attn_bwd2_dq_kernel(const __grid_constant__ CUtensorMap tmQ128, const __grid_constant__ CUtensorMap tmKV64,
                    const __grid_constant__ CUtensorMap tmDO128, const float* __restrict__ lse,
                    const float* __restrict__ delta, bf16* __restrict__ dqkv, int S, int H, float scale) {
  using C = Bwd2Cfg<DH>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t sQ = base, sdO = sQ + C::TILE, sK = sdO + C::TILE /*3*/, sV = sK + 3 * C::HT /*2*/,
                 sdS = sV + 2 * C::HT;
  const uint32_t bars = sdS + C::ST;
  const uint32_t bar_q = bars, bar_k = bars + 8 /*3*/, bar_v = bars + 32 /*2*/, bar_a = bars + 48 /*2*/,
                 bar_b = bars + 64;
  const uint32_t tmem_slot = bars + 72;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  const int tid = threadIdx.x, warp = tid >> 5;
  const int rowi = tid & 127, half = tid >> 7;
  const int ib = gridDim.x - 1 - blockIdx.x;  // heavy query blocks first
  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = ib * 128;
  const int n_kv = (min(S, q0 + 128) + 63) / 64;

  if (tid == 0) {
    tma_prefetch_desc(&tmQ128);
    tma_prefetch_desc(&tmKV64);
    tma_prefetch_desc(&tmDO128);
    mbar_init(bar_q, 1);
    for (int i = 0; i < 3; ++i) mbar_init(bar_k + 8 * i, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(bar_v + 8 * i, 1); mbar_init(bar_a + 8 * i, 1); }
    mbar_init(bar_b, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const uint32_t tS = tmem /*2 x 64*/, tdP = tmem + 128 /*2 x 64*/, tdQ = tmem + 256;
  const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;

  constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, 0, 0);
  constexpr uint32_t idesc_q = umma_idesc_bf16(128, DH, 0, 1);
  auto load_k = [&](int j) {
    const uint32_t bk = bar_k + 8 * (j % 3);
    mbar_expect_tx(bk, C::HT);
    tma_load_tile<DH>(sK + (j % 3) * C::HT, T64, &tmKV64, bk, 1 * H + h, j * 64, b);
  };
  auto load_v = [&](int j) {
    const uint32_t bv = bar_v + 8 * (j & 1);
    mbar_expect_tx(bv, C::HT);
    tma_load_tile<DH>(sV + (j & 1) * C::HT, T64, &tmKV64, bv, 2 * H + h, j * 64, b);
  };
  auto issue_logits = [&](int j) {  // S_j, dP_j into TMEM buffers j & 1
    const uint32_t kb = sK + (j % 3) * C::HT, vb = sV + (j & 1) * C::HT;
#pragma unroll
    for (int kk = 0; kk < DH / 16; ++kk)
      umma_bf16_ss(tS + (j & 1) * 64, umma_smem_desc_sw128(sQ + (kk / 4) * T128 + (kk % 4) * 32, 0, 1024),
                   umma_smem_desc_sw128(kb + (kk / 4) * T64 + (kk % 4) * 32, 0, 1024), idesc_s, kk > 0);
#pragma unroll
    for (int kk = 0; kk < DH / 16; ++kk)
      umma_bf16_ss(tdP + (j & 1) * 64, umma_smem_desc_sw128(sdO + (kk / 4) * T128 + (kk % 4) * 32, 0, 1024),
                   umma_smem_desc_sw128(vb + (kk / 4) * T64 + (kk % 4) * 32, 0, 1024), idesc_s, kk > 0);
    umma_commit(bar_a + 8 * (j & 1));
  };

  if (tid == 0) {
    mbar_expect_tx(bar_q, 2 * C::TILE);
    tma_load_tile<DH>(sQ, T128, &tmQ128, bar_q, 0 * H + h, q0, b);
    tma_load_tile<DH>(sdO, T128, &tmDO128, bar_q, h, q0, b);
    load_k(0); load_v(0);
    if (n_kv > 1) { load_k(1); load_v(1); }
    mbar_wait(bar_q, 0);
    mbar_wait(bar_k, 0);
    mbar_wait(bar_v, 0);
    tc_fence_after();
    issue_logits(0);
  }
  const float c1 = scale * LOG2E;
  const int qi = q0 + rowi;
  const bool row_ok = qi < S;
  float lse_l2 = 0.f, dl = 0.f;
  if (row_ok) {
    lse_l2 = lse[((long long)b * H + h) * S + qi] * LOG2E;
    dl = delta[((long long)b * H + h) * S + qi];
  }

  for (int j = 0; j < n_kv; ++j) {
    const int sb = j & 1;
    mbar_wait(bar_a + 8 * sb, (j >> 1) & 1);  // (A) S_j, dP_j ready; V buffer sb is free
    tc_fence_after();
    if (tid == 0 && j + 2 < n_kv) load_v(j + 2);
    // (B) this thread's 32 columns -> dS (bf16 pairs in registers)
    uint32_t dk[16];
    {
      const int col0 = half * 32, k0 = j * 64 + col0;
      uint32_t rs[32], rd[32];
      tmem_ld_x32(tS + sb * 64 + lane_off + col0, rs);
      tmem_ld_x32(tdP + sb * 64 + lane_off + col0, rd);
      tmem_ld_wait();
      const bool diag = (j * 64 + 63) > q0;  // block reaches past the first query of the tile: mask per element
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        float p0 = ex2(fmaf(__uint_as_float(rs[i]), c1, -lse_l2));
        float p1 = ex2(fmaf(__uint_as_float(rs[i + 1]), c1, -lse_l2));
        if (!row_ok || (diag && (k0 + i) > qi)) p0 = 0.f;
        if (!row_ok || (diag && (k0 + i + 1) > qi)) p1 = 0.f;
        dk[i >> 1] = pack_bf16x2((p0 * scale) * (__uint_as_float(rd[i]) - dl),
                                 (p1 * scale) * (__uint_as_float(rd[i + 1]) - dl));
      }
    }
    tc_fence_before();
    __syncthreads();  // (1) every thread has read tS / tdP [sb]
    if (tid == 0 && j + 1 < n_kv) {  // (C) next block's logits under this block's epilogue math
      mbar_wait(bar_k + 8 * ((j + 1) % 3), ((j + 1) / 3) & 1);
      mbar_wait(bar_v + 8 * (sb ^ 1), ((j + 1) >> 1) & 1);
      tc_fence_after();
      issue_logits(j + 1);
    }
    if (j > 0) {  // (E) dQ MMAs of block j-1 done: sdS and K ring slot (j-1) % 3 are free
      mbar_wait(bar_b, (j - 1) & 1);
      tc_fence_after();
    }
    if (tid == 0 && j + 2 < n_kv) load_k(j + 2);
    {  // (F) dS -> swizzled [128][64] tile
      const int col = half * 32;
#pragma unroll
      for (int g = 0; g < 4; ++g)
        st_shared_v4(sdS + sw128_offset(rowi, col + g * 8), dk[g * 4], dk[g * 4 + 1], dk[g * 4 + 2], dk[g * 4 + 3]);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();  // (2)
    if (tid == 0) {   // (G) dQ += dS K_j
      tc_fence_after();
      const uint32_t kb = sK + (j % 3) * C::HT;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        umma_bf16_ss(tdQ, umma_smem_desc_sw128(sdS + kk * 32, 0, 1024), umma_smem_desc_sw128(kb + kk * 2048, T64, 1024),
                     idesc_q, (j > 0 || kk > 0));
      umma_commit(bar_b);
    }
  }
  mbar_wait(bar_b, (n_kv - 1) & 1);
  tc_fence_after();
  bf16* dst = dqkv + ((((long long)b * S + qi) * 3 + 0) * H + h) * DH + half * (DH / 2);
#pragma unroll 1
  for (int c = 0; c < DH / 64; ++c) {
    uint32_t r[32];
    tmem_ld_x32(tdQ + lane_off + half * (DH / 2) + c * 32, r);
    tmem_ld_wait();
    if (row_ok) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 q;
        q.x = pack_bf16x2(__uint_as_float(r[g * 8 + 0]), __uint_as_float(r[g * 8 + 1]));
        q.y = pack_bf16x2(__uint_as_float(r[g * 8 + 2]), __uint_as_float(r[g * 8 + 3]));
        q.z = pack_bf16x2(__uint_as_float(r[g * 8 + 4]), __uint_as_float(r[g * 8 + 5]));
        q.w = pack_bf16x2(__uint_as_float(r[g * 8 + 6]), __uint_as_float(r[g * 8 + 7]));
        *reinterpret_cast<uint4*>(dst + c * 32 + g * 8) = q;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <int DH>
__global__ void __launch_bounds__(256, 1)
attn_bwd2_dkdv_kernel(const __grid_constant__ CUtensorMap tmKV128, const __grid_constant__ CUtensorMap tmQ64,
                      const __grid_constant__ CUtensorMap tmDO64, const float* __restrict__ lse,
                      const float* __restrict__ delta, bf16* __restrict__ dqkv, int S, int H, float scale) {
  using C = Bwd2Cfg<DH>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t sK = base, sV = sK + C::TILE, sQ = sV + C::TILE /*3*/, sdO = sQ + 3 * C::HT /*3*/,
                 sPT = sdO + 3 * C::HT, sdST = sPT + C::ST;
  const uint32_t bars = sdST + C::ST;
  const uint32_t bar_kv = bars, bar_qd = bars + 8 /*3*/, bar_a = bars + 32 /*2*/, bar_b = bars + 48;
  const uint32_t tmem_slot = bars + 56;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  float* stat = reinterpret_cast<float*>(smem_raw + (bars + 256 - raw));  // [2 buffers][2: lse*log2e, delta][64]

  const int tid = threadIdx.x, warp = tid >> 5;
  const int rowi = tid & 127, half = tid >> 7;
  const int jb = blockIdx.x;  // kv block (block 0 has the most work and is scheduled first)
  const int h = blockIdx.y, b = blockIdx.z;
  const int k0 = jb * 128;
  const int i0 = 2 * jb;                 // first 64-query block that can see these keys
  const int n_it = (S + 63) / 64 - i0;   // >= 1 because k0 < S

  if (tid == 0) {
    tma_prefetch_desc(&tmKV128);
    tma_prefetch_desc(&tmQ64);
    tma_prefetch_desc(&tmDO64);
    mbar_init(bar_kv, 1);
    for (int i = 0; i < 3; ++i) mbar_init(bar_qd + 8 * i, 1);
    for (int i = 0; i < 2; ++i) mbar_init(bar_a + 8 * i, 1);
    mbar_init(bar_b, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const uint32_t tST = tmem /*2 x 64*/, tdPT = tmem + 128 /*2 x 64*/, tdV = tmem + 256, tdK = tmem + 256 + DH;
  const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;

  constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, 0, 0);   // S^T = K Q^T, dP^T = V dO^T
  constexpr uint32_t idesc_g = umma_idesc_bf16(128, DH, 0, 1);   // dV = P^T dO, dK = dS^T Q : A K-major, B MN-major
  auto load_qd = [&](int it) {  // Q and dO rows of 64-query block i0 + it into ring slot it % 3
    const uint32_t bq = bar_qd + 8 * (it % 3);
    mbar_expect_tx(bq, 2 * C::HT);
    tma_load_tile<DH>(sQ + (it % 3) * C::HT, T64, &tmQ64, bq, 0 * H + h, (i0 + it) * 64, b);
    tma_load_tile<DH>(sdO + (it % 3) * C::HT, T64, &tmDO64, bq, h, (i0 + it) * 64, b);
  };
  auto issue_logits = [&](int it) {
    const uint32_t qb = sQ + (it % 3) * C::HT, ob = sdO + (it % 3) * C::HT;
#pragma unroll
    for (int kk = 0; kk < DH / 16; ++kk)
      umma_bf16_ss(tST + (it & 1) * 64, umma_smem_desc_sw128(sK + (kk / 4) * T128 + (kk % 4) * 32, 0, 1024),
                   umma_smem_desc_sw128(qb + (kk / 4) * T64 + (kk % 4) * 32, 0, 1024), idesc_s, kk > 0);
#pragma unroll
    for (int kk = 0; kk < DH / 16; ++kk)
      umma_bf16_ss(tdPT + (it & 1) * 64, umma_smem_desc_sw128(sV + (kk / 4) * T128 + (kk % 4) * 32, 0, 1024),
                   umma_smem_desc_sw128(ob + (kk / 4) * T64 + (kk % 4) * 32, 0, 1024), idesc_s, kk > 0);
    umma_commit(bar_a + 8 * (it & 1));
  };
  auto stage_stats = [&](int it) {  // lse * log2(e) and delta of the block's 64 queries (+inf / 0 beyond S: p = 0)
    if (tid < 64) {
      const int q = (i0 + it) * 64 + tid;
      float l2 = INFINITY, d = 0.f;
      if (q < S) {
        l2 = lse[((long long)b * H + h) * S + q] * LOG2E;
        d = delta[((long long)b * H + h) * S + q];
      }
      stat[(it & 1) * 128 + tid] = l2;
      stat[(it & 1) * 128 + 64 + tid] = d;
    }
  };

  if (tid == 0) {
    mbar_expect_tx(bar_kv, 2 * C::TILE);
    tma_load_tile<DH>(sK, T128, &tmKV128, bar_kv, 1 * H + h, k0, b);
    tma_load_tile<DH>(sV, T128, &tmKV128, bar_kv, 2 * H + h, k0, b);
    load_qd(0);
    if (n_it > 1) load_qd(1);
    mbar_wait(bar_kv, 0);
    mbar_wait(bar_qd, 0);
    tc_fence_after();
    issue_logits(0);
  }
  stage_stats(0);
  __syncthreads();
  const float c1 = scale * LOG2E;
  const int ki = k0 + rowi;

  for (int it = 0; it < n_it; ++it) {
    const int sb = it & 1;
    const int qbase = (i0 + it) * 64;
    mbar_wait(bar_a + 8 * sb, (it >> 1) & 1);  // (A)
    tc_fence_after();
    // (B) P^T, dS^T for this key row and 32 of the block's 64 queries
    uint32_t pk[16], dk[16];
    {
      const int col0 = half * 32;
      uint32_t rs[32], rd[32];
      tmem_ld_x32(tST + sb * 64 + lane_off + col0, rs);
      tmem_ld_x32(tdPT + sb * 64 + lane_off + col0, rd);
      tmem_ld_wait();
      const float* l2 = stat + sb * 128 + col0;
      const float* dd = l2 + 64;
      const bool diag = (k0 + 127) > qbase;  // some (key, query) pairs of this block are in the future
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        float p0 = ex2(fmaf(__uint_as_float(rs[i]), c1, -l2[i]));
        float p1 = ex2(fmaf(__uint_as_float(rs[i + 1]), c1, -l2[i + 1]));
        if (diag && ki > (qbase + col0 + i)) p0 = 0.f;
        if (diag && ki > (qbase + col0 + i + 1)) p1 = 0.f;
        pk[i >> 1] = pack_bf16x2(p0, p1);
        dk[i >> 1] = pack_bf16x2((p0 * scale) * (__uint_as_float(rd[i]) - dd[i]),
                                 (p1 * scale) * (__uint_as_float(rd[i + 1]) - dd[i + 1]));
      }
    }
    tc_fence_before();
    __syncthreads();  // (1) TMEM logits and this block's stats consumed
    if (tid == 0 && it + 1 < n_it) {  // (C)
      mbar_wait(bar_qd + 8 * ((it + 1) % 3), ((it + 1) / 3) & 1);
      tc_fence_after();
      issue_logits(it + 1);
    }
    if (it + 1 < n_it) stage_stats(it + 1);  // other stats buffer: last read in iteration it - 1
    if (it > 0) {  // (E) gradient MMAs of block it-1 done: P^T / dS^T tiles and ring slot (it-1) % 3 are free
      mbar_wait(bar_b, (it - 1) & 1);
      tc_fence_after();
    }
    if (tid == 0 && it + 2 < n_it) load_qd(it + 2);
    {  // (F)
      const int col = half * 32;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint32_t off = sw128_offset(rowi, col + g * 8);
        st_shared_v4(sPT + off, pk[g * 4], pk[g * 4 + 1], pk[g * 4 + 2], pk[g * 4 + 3]);
        st_shared_v4(sdST + off, dk[g * 4], dk[g * 4 + 1], dk[g * 4 + 2], dk[g * 4 + 3]);
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();  // (2) tiles and next stats visible
    if (tid == 0) {   // (G) dV += P^T dO_i ; dK += dS^T Q_i   (K dimension = the 64 queries)
      tc_fence_after();
      const uint32_t qb = sQ + (it % 3) * C::HT, ob = sdO + (it % 3) * C::HT;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        umma_bf16_ss(tdV, umma_smem_desc_sw128(sPT + kk * 32, 0, 1024), umma_smem_desc_sw128(ob + kk * 2048, T64, 1024),
                     idesc_g, (it > 0 || kk > 0));
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        umma_bf16_ss(tdK, umma_smem_desc_sw128(sdST + kk * 32, 0, 1024), umma_smem_desc_sw128(qb + kk * 2048, T64, 1024),
                     idesc_g, (it > 0 || kk > 0));
      umma_commit(bar_b);
    }
  }
  mbar_wait(bar_b, (n_it - 1) & 1);
  tc_fence_after();
#pragma unroll 1
  for (int which = 0; which < 2; ++which) {
    const uint32_t tsrc = which == 0 ? tdK : tdV;
    bf16* dst = dqkv + ((((long long)b * S + ki) * 3 + (which == 0 ? 1 : 2)) * H + h) * DH + half * (DH / 2);
#pragma unroll 1
    for (int c = 0; c < DH / 64; ++c) {
      uint32_t r[32];
      tmem_ld_x32(tsrc + lane_off + half * (DH / 2) + c * 32, r);
      tmem_ld_wait();
      if (ki < S) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 q;
          q.x = pack_bf16x2(__uint_as_float(r[g * 8 + 0]), __uint_as_float(r[g * 8 + 1]));
          q.y = pack_bf16x2(__uint_as_float(r[g * 8 + 2]), __uint_as_float(r[g * 8 + 3]));
          q.z = pack_bf16x2(__uint_as_float(r[g * 8 + 4]), __uint_as_float(r[g * 8 + 5]));
          q.w = pack_bf16x2(__uint_as_float(r[g * 8 + 6]), __uint_as_float(r[g * 8 + 7]));
          *reinterpret_cast<uint4*>(dst + c * 32 + g * 8) = q;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
// DB200_ATTN_V2: bit 0 = pipelined forward, bit 1 = pipelined backward (experimental kernels; default 0)
static int attn_v2_bits() {
  static const int bits = [] { const char* e = getenv("DB200_ATTN_V2"); return e ? atoi(e) : 0; }();
  return bits;
}

static int make_qkv_map(CUtensorMap* tm, const void* qkv, int B, int S, int H, int dh, uint32_t box_rows) {
  uint64_t dims[4] = {(uint64_t)dh, (uint64_t)3 * H, (uint64_t)S, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)dh * 2, (uint64_t)3 * H * dh * 2, (uint64_t)S * 3 * H * dh * 2};
  uint32_t box[4] = {64, 1, box_rows, 1};
  return make_tmap_bf16(tm, qkv, 4, dims, strides, box);
}
static int make_o_map(CUtensorMap* tm, const void* o, int B, int S, int H, int dh, uint32_t box_rows) {
  uint64_t dims[4] = {(uint64_t)dh, (uint64_t)H, (uint64_t)S, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)dh * 2, (uint64_t)H * dh * 2, (uint64_t)S * H * dh * 2};
  uint32_t box[4] = {64, 1, box_rows, 1};
  return make_tmap_bf16(tm, o, 4, dims, strides, box);
}

template <int DH>
static int fwd_launch(cudaStream_t stream, const void* qkv, void* out, float* lse, int B, int S, int H, float scale) {
  using C = FwdCfg<DH>;
  CUtensorMap tmQ, tmKV;
  int rc = make_qkv_map(&tmQ, qkv, B, S, H, DH, 128);
  if (rc != DB200_OK) return rc;
  rc = make_qkv_map(&tmKV, qkv, B, S, H, DH, C::BNK);
  if (rc != DB200_OK) return rc;
  static bool attr = false;
  if (!attr) {
    DB200_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    attr = true;
  }
  dim3 grid((S + 127) / 128, H, B);
  if (attn_v2_bits() & 1) {  // experimental pipelined variant (see attn_fwd2_kernel); 64-key blocks for both head sizes
    using C2 = Fwd2Cfg<DH>;
    CUtensorMap tmKV2;
    rc = make_qkv_map(&tmKV2, qkv, B, S, H, DH, C2::BNK);
    if (rc != DB200_OK) return rc;
    static bool attr2 = false;
    if (!attr2) {
      DB200_CUDA(cudaFuncSetAttribute(attn_fwd2_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C2::SMEM));
      attr2 = true;
    }
    attn_fwd2_kernel<DH><<<grid, 256, C2::SMEM, stream>>>(tmQ, tmKV2, (bf16*)out, lse, S, H, scale);
    return check_launch("attn_fwd2_kernel");
  }
  attn_fwd_kernel<DH><<<grid, 256, C::SMEM, stream>>>(tmQ, tmKV, (bf16*)out, lse, S, H, scale);
  return check_launch("attn_fwd_kernel");
}

template <int DH>
static int bwd_launch(cudaStream_t stream, const void* qkv, const void* dout, const float* lse, const float* delta,
                      void* dqkv, int B, int S, int H, float scale) {
  using C = BwdCfg<DH>;
  CUtensorMap tmQKV, tmDO;
  int rc = make_qkv_map(&tmQKV, qkv, B, S, H, DH, 128);
  if (rc != DB200_OK) return rc;
  rc = make_o_map(&tmDO, dout, B, S, H, DH, 128);
  if (rc != DB200_OK) return rc;
  static bool attr = false;
  if (!attr) {
    DB200_CUDA(cudaFuncSetAttribute(attn_bwd_dkdv_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)C::SMEM_DKDV));
    DB200_CUDA(
        cudaFuncSetAttribute(attn_bwd_dq_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_DQ));
    attr = true;
  }
  dim3 grid((S + 127) / 128, H, B);
  if (attn_v2_bits() & 2) {  // experimental pipelined backward kernels
    using C2 = Bwd2Cfg<DH>;
    CUtensorMap tmQKV64, tmDO64;
    rc = make_qkv_map(&tmQKV64, qkv, B, S, H, DH, 64);
    if (rc != DB200_OK) return rc;
    rc = make_o_map(&tmDO64, dout, B, S, H, DH, 64);
    if (rc != DB200_OK) return rc;
    static bool attr2 = false;
    if (!attr2) {
      DB200_CUDA(cudaFuncSetAttribute(attn_bwd2_dkdv_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)C2::SMEM_DKDV));
      DB200_CUDA(cudaFuncSetAttribute(attn_bwd2_dq_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)C2::SMEM_DQ));
      attr2 = true;
    }
    attn_bwd2_dkdv_kernel<DH><<<grid, 256, C2::SMEM_DKDV, stream>>>(tmQKV, tmQKV64, tmDO64, lse, delta, (bf16*)dqkv, S, H,
                                                                  scale);
    rc = check_launch("attn_bwd2_dkdv_kernel");
    if (rc != DB200_OK) return rc;
    attn_bwd2_dq_kernel<DH><<<grid, 256, C2::SMEM_DQ, stream>>>(tmQKV, tmQKV64, tmDO, lse, delta, (bf16*)dqkv, S, H,
                                                              scale);
    return check_launch("attn_bwd2_dq_kernel");
  }
  attn_bwd_dkdv_kernel<DH><<<grid, 256, C::SMEM_DKDV, stream>>>(tmQKV, tmDO, lse, delta, (bf16*)dqkv, S, H, scale);
  rc = check_launch("attn_bwd_dkdv_kernel");
  if (rc != DB200_OK) return rc;
  attn_bwd_dq_kernel<DH><<<grid, 256, C::SMEM_DQ, stream>>>(tmQKV, tmDO, lse, delta, (bf16*)dqkv, S, H, scale);
  return check_launch("attn_bwd_dq_kernel");
}

}  // namespace db200

using namespace db200;

extern "C" int db200_attn_causal_fwd(db200_stream_t stream_, const void* qkv, void* out, float* lse, int B, int S,
                                     int H, int dh, float scale) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DB200_REQUIRE(B > 0 && S > 0 && H > 0, DB200_E_INVALID, "attn_fwd: B,S,H must be positive");
  DB200_REQUIRE(qkv && out && lse && aligned16(qkv) && aligned16(out), DB200_E_ALIGN,
                "attn_fwd: NULL or unaligned pointer");
  DB200_REQUIRE(scale > 0.f, DB200_E_INVALID, "attn_fwd: scale must be > 0");
  if (dh == 128) return fwd_launch<128>(stream, qkv, out, lse, B, S, H, scale);
  if (dh == 64) return fwd_launch<64>(stream, qkv, out, lse, B, S, H, scale);
  return set_error(DB200_E_UNSUPPORTED, "attn: head_dim %d not in {64,128}", dh);
}

extern "C" int db200_attn_causal_bwd(db200_stream_t stream_, const void* qkv, const void* out, const void* dout,
                                     const float* lse, float* dq_accum, float* delta, void* dqkv, int B, int S, int H,
                                     int dh, float scale) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  (void)dq_accum;  // kept in the ABI for an atomics-based variant; the two-kernel scheme does not need it
  DB200_REQUIRE(B > 0 && S > 0 && H > 0, DB200_E_INVALID, "attn_bwd: B,S,H must be positive");
  DB200_REQUIRE(qkv && out && dout && lse && delta && dqkv && aligned16(qkv) && aligned16(out) && aligned16(dout) &&
                    aligned16(dqkv),
                DB200_E_ALIGN, "attn_bwd: NULL or unaligned pointer");
  DB200_REQUIRE(scale > 0.f, DB200_E_INVALID, "attn_bwd: scale must be > 0");
  DB200_REQUIRE(dh == 64 || dh == 128, DB200_E_UNSUPPORTED, "attn: head_dim %d not in {64,128}", dh);
  const long long rows = (long long)B * S * H;
  attn_delta_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>((const bf16*)out, (const bf16*)dout, delta, B, S,
                                                                    H, dh);
  int rc = check_launch("attn_delta_kernel");
  if (rc != DB200_OK) return rc;
  if (dh == 128) return bwd_launch<128>(stream, qkv, dout, lse, delta, dqkv, B, S, H, scale);
  return bwd_launch<64>(stream, qkv, dout, lse, delta, dqkv, B, S, H, scale);
}
