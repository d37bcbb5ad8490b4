// K4 (warp-specialised) — causal flash attention, forward and backward, on tcgen05 / TMEM / TMA for sm_100a.
//
// Replaces mtf_transformer.attention.attention + the [S,S] additive -1e10 mask the reference materialises
// (src/dalle_mtf/models.py:221-227, 287-299) and their gradients (mtf.gradients, src/optimizers.py:34): logits fp32
// (TMEM accumulators), softmax fp32, P rounded to bf16 for the P.V product (what mtf does when it casts the weights to
// v's dtype), nothing of size S x S reaches HBM or even shared memory.  `scale` multiplies q.k (reference: 1.0 — mtf
// folds 1/sqrt(dh) into the q initialiser).
//
// Layout: qkv bf16 [B][S][3][H][dh] (the fused q|k|v projection output), out / dout bf16 [B][S][H][dh],
//         lse f32 [B][H][S] (natural log of sum exp(scale*s)), delta f32 [B][H][S] = rowsum(dO * O), dqkv like qkv.
//
// Common structure of the three kernels (one CTA per 128-row tile of one (batch, head), 1-D grid ordered heaviest tile
// first): warps 0-3 and 4-7 = two "math" warpgroups (thread r of either group owns TMEM lane r; group g owns columns
// [64g, 64g+64) of every 128-wide block), warp 8 = TMA producer (one lane), warp 9 = MMA issuer (one lane), warp 10 =
// lse / delta stager where needed.  All hand-offs are mbarriers: TMA -> MMA (complete_tx), MMA -> TMA and MMA -> math
// (tcgen05.commit), math -> MMA (one arrive per warp).  Probabilities / logit gradients are handed to the next product
// THROUGH TENSOR MEMORY: the math threads write them as packed bf16 with tcgen05.st and the product reads them as its A
// operand (tcgen05.mma TS form) — no shared-memory round trip, no proxy fence.  Each group writes its half of such an
// operand into the first 32 columns of ITS OWN 64-column slice, so the two groups never touch each other's columns and
// the operand's K-steps are addressed as slice 0 (columns +0..+31) then slice 1 (+64..+95).
//
//   forward : S double-buffered (2 x 128 columns); issue order  S0 S1 | PV0 S2 | PV1 S3 | ...  so the logits of block
//             j+1 are always ready when the softmax of block j ends; O accumulates in TMEM across blocks and a row is
//             rescaled (tcgen05.ld -> scale -> tcgen05.st) only when its running maximum moved by more than 2^8 since
//             the scale it uses (lazy rescaling); the two groups exchange their half-row maxima through shared memory.
//   dK / dV : TMEM lanes = keys, transposed orientation S^T = K Q_i^T, dP^T = V dO_i^T.  Two-phase pipeline per query
//             block:  phase A  P^T = exp(S^T - lse)  runs under the dP^T product, phase B  dS^T = P^T (dP^T - delta)
//             under the next block's S^T product:   S0 dP0 | [A0] dV0 S1 | [B0] dK0 dP1 | [A1] dV1 S2 | ...
//   dQ      : lanes = queries; the groups release S / dP as soon as they are in registers, so the next block's two logit
//             products run under the dS arithmetic; dS goes to its own double-buffered TMEM slot:  dQ += dS K_j.
#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"

namespace db200 {

namespace {

constexpr float LOG2E_F = 1.4426950408889634f;
constexpr uint32_t T128B = 128 * 128;  // bytes of one [128 rows][64] swizzled sub-tile

// Development builds only (make DEV=1): per-CTA event timeline.  Every TRACE_STRIDE-th CTA records (code, SM clock)
// pairs from lane 0 of four of its warps (role 0 = TMA, 1 = MMA issuer, 2 = first math warp, 3 = last math warp) into
// the buffer set with db200_dev_attn_trace(); tools/attn_trace.py prints them.  Compiled out of the shipped library.
#ifdef DB200_DEV_KNOBS
constexpr int TRACE_EV = 320, TRACE_STRIDE = 37;
__device__ unsigned long long* g_attn_trace = nullptr;
__device__ int g_attn_trace_slots = 0;
struct Tracer {
  unsigned long long* p;
  int n;
  __device__ __forceinline__ void init(int role, int lane) {
    p = nullptr; n = 0;
    const int slot = (int)blockIdx.x / TRACE_STRIDE;
    if (g_attn_trace && lane == 0 && role >= 0 && (int)blockIdx.x % TRACE_STRIDE == 0 && slot < g_attn_trace_slots) {
      p = g_attn_trace + ((size_t)slot * 4 + role) * TRACE_EV;
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      p[0] = ((unsigned long long)(0xF00000u | smid) << 40) | ((unsigned long long)blockIdx.x & 0xFFFFFFFFFFull);
      n = 1;
    }
  }
  __device__ __forceinline__ void ev(int type, int j) {
    if (p && n < TRACE_EV) p[n++] = ((unsigned long long)((type << 12) | (j & 0xFFF)) << 40) | ((unsigned long long)clock64() & 0xFFFFFFFFFFull);
  }
};
#else
struct Tracer {
  __device__ __forceinline__ void init(int, int) {}
  __device__ __forceinline__ void ev(int, int) {}
};
#endif

// MUFU.EX2 directly: arguments are <= 8 after the running-max subtraction; ex2.approx(-inf) = +0.
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// all threads of the NG math warpgroups (named barrier 1; barrier 0 is __syncthreads)
template <int NG>
__device__ __forceinline__ void math_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(128 * NG) : "memory"); }

template <int DH>
__device__ __forceinline__ void ws_load_tile(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int chan, int row0,
                                             int b) {
#pragma unroll
  for (int t = 0; t < DH / 64; ++t) tma_load_4d(dst + t * T128B, tm, bar, 64 * t, chan, row0, b);
}

// UMMA shared-memory descriptors of a [128][DH] tile: built ONCE per tile, K-steps are `base + constant` (the start
// address field holds address >> 4 and cannot overflow: shared memory ends below 256 KiB).
//   K-major  (K = dh):                  K-step kk = 16 elements  -> (kk / 4) sub-tiles of 16 KiB + (kk % 4) * 32 bytes
//   MN-major (K = the 128 rows, N = dh): K-step kk = 16 rows      -> kk * 2048 bytes; 64-wide N atoms are T128B apart (LBO)
__device__ __forceinline__ uint64_t desc_k_base(uint32_t tile) { return umma_smem_desc_sw128(tile, 0, 1024); }
__device__ __forceinline__ uint64_t desc_mn_base(uint32_t tile) { return umma_smem_desc_sw128(tile, T128B, 1024); }
__device__ __forceinline__ constexpr uint64_t kstep_k(int kk) { return (uint64_t)((kk / 4) * (T128B >> 4) + (kk % 4) * 2); }
__device__ __forceinline__ constexpr uint64_t kstep_mn(int kk) { return (uint64_t)(kk * 128); }
// column of K-step kk (16 bf16 = 8 packed columns) of a TMEM A operand whose 128 K-elements are stored group by group:
// group g's 128/NG elements sit packed at the start of its own 128/NG-column slice
template <int NG>
__device__ __forceinline__ uint32_t ts_split_col(int kk) {
  constexpr int CG = 128 / NG;            // elements (= fp32 columns) per group slice
  const int e = kk * 16;                  // first K element of this step
  return (e / CG) * CG + (e % CG) / 2;
}
// 32-bit TMEM loads / stores of N consecutive columns (N = 16 or 32)
template <int N>
__device__ __forceinline__ void tmem_ld_n(uint32_t taddr, uint32_t* r) {
  if constexpr (N == 32) tmem_ld_x32(taddr, r); else tmem_ld_x16(taddr, r);
}
template <int N>
__device__ __forceinline__ void tmem_st_n(uint32_t taddr, const uint32_t* r) {
  if constexpr (N == 32) tmem_st_x32(taddr, r); else tmem_st_x16(taddr, r);
}
template <int N>
__device__ __forceinline__ void store_cols_bf16(bf16* dst, const uint32_t* r, float mul) {
#pragma unroll
  for (int e = 0; e < N; e += 8) {
    uint4 q;
    q.x = pack_bf16x2(__uint_as_float(r[e]) * mul, __uint_as_float(r[e + 1]) * mul);
    q.y = pack_bf16x2(__uint_as_float(r[e + 2]) * mul, __uint_as_float(r[e + 3]) * mul);
    q.z = pack_bf16x2(__uint_as_float(r[e + 4]) * mul, __uint_as_float(r[e + 5]) * mul);
    q.w = pack_bf16x2(__uint_as_float(r[e + 6]) * mul, __uint_as_float(r[e + 7]) * mul);
    *reinterpret_cast<uint4*>(dst + e) = q;
  }
}

__device__ __forceinline__ void warp_arrive(uint32_t bar, int lane) {
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

// ------------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------------
template <int DH, int NG>
struct FwdWs {
  static constexpr int NK = (DH == 128) ? 3 : 4;   // K ring depth (128-key blocks): logits run two blocks ahead
  static constexpr int NV = (DH == 128) ? 2 : 4;   // V ring depth
  static constexpr uint32_t TILE = 128 * DH * 2;   // one [128][DH] bf16 operand tile
  static constexpr uint32_t XCH_BYTES = 2 * NG * 128 * 4;
  static constexpr uint32_t BAR_BYTES = 320;
  // dynamic shared memory is declared 1024-byte aligned (the 128-byte swizzle needs it): no alignment slack
  static constexpr size_t SMEM = TILE + (NK + NV) * TILE + XCH_BYTES + BAR_BYTES;
  static_assert(SMEM <= 232448, "forward attention: shared memory over the 227 KiB per-CTA limit");
  static constexpr int THREADS = (4 * NG + 2) * 32;  // NG math warpgroups + TMA warp + MMA warp
};

}  // namespace

// NG = number of softmax warpgroups = column groups of every 128-key block (2: 64 columns per thread, 4: 32).  More
// groups = more resident warps per scheduler: the per-warp instruction stream is a chain of dependent fp32 / MUFU ops,
// and with two math warps per scheduler it issues once every ~6 cycles (ncu, profiles/ncu_attn_r02.md).
//
// PERSISTENT: the grid is one CTA per SM; CTA c walks the work items c, c + G, c + 2G, ... of the list ordered by
// decreasing work (item w = (query tile n_qt-1 - w / (B H), head, batch)), so barrier initialisation, the TMEM
// allocation and — above all — the exposed latency of the first loads are paid once per SM instead of once per
// 128-query tile (1 280 tiles of 1..10 key blocks at the bench shape: the fixed cost per tile was as large as its work).
// Every barrier therefore runs on GLOBAL use counters that continue across items:
//   kc / vc   K / V ring uses         slot = c % N, parity (c / N) & 1
//   bc        key blocks processed    S / P buffer bc & 1, parity (bc >> 1) & 1;  o_done completes once per P.V: the
//                                     product of block bc has parity bc & 1
//   it        items of this CTA       Q buffer parity it & 1; O accumulator it & 1 (double-buffered: the epilogue of
//                                     item `it` overlaps the first products of item it + 1), o_free[it & 1]
// The TMA warp is a free-running stream of (Q, K_j, V_j) requests over all items: the loads of the next tile are in
// flight while the current one finishes.
template <int DH, int NG>
__global__ void __launch_bounds__((4 * NG + 2) * 32, 1)
attn_fwd_ws_kernel(const __grid_constant__ CUtensorMap tmQKV, bf16* __restrict__ out, float* __restrict__ lse_out,
                   int S, int H, int n_items, float scale, int xflags) {
#ifndef DB200_DEV_KNOBS
  xflags = 0;  // timing experiments exist in development builds only (make DEV=1); results are wrong under them
#endif
  using C = FwdWs<DH, NG>;
  constexpr int NK = C::NK, NV = C::NV;
  constexpr int CG = 128 / NG;       // key columns of a block per group (= per thread)
  constexpr int OG = DH / NG;        // output columns per group
  constexpr int TMA_WARP = 4 * NG, MMA_WARP = 4 * NG + 1;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = raw;                             // 1024-byte aligned (checked below)
  const uint32_t sQ = base, sK = sQ + C::TILE, sV = sK + NK * C::TILE;
  const uint32_t sX = sV + NV * C::TILE;                 // [2 parities][NG groups][128 rows] f32 maxima / sums
  const uint32_t bars = sX + C::XCH_BYTES;
  const uint32_t q_full = bars;
  const uint32_t q_empty = bars + 8;               // every logits product of the item has retired: Q may be replaced
  const uint32_t k_full = bars + 16;               // [NK]
  const uint32_t k_empty = k_full + 8 * NK;        // [NK]
  const uint32_t v_full = k_empty + 8 * NK;        // [NV]
  const uint32_t v_empty = v_full + 8 * NV;        // [NV]
  const uint32_t s_ready = v_empty + 8 * NV;       // [2]  S buffer b holds the block with (bc & 1) == b
  const uint32_t p_ready = s_ready + 16;           // [2]
  const uint32_t o_done = p_ready + 16;            // a P.V product has landed in O
  const uint32_t o_free = o_done + 8;              // [2]  the epilogue has read O accumulator b
  const uint32_t tmem_slot = o_free + 16;
  static_assert(8 * (2 + 2 * NK + 2 * NV + 2 + 2 + 1 + 2) + 8 <= C::BAR_BYTES, "barrier area");
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  float* xch = reinterpret_cast<float*>(smem_raw + (sX - raw));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_qt = (S + 127) >> 7;
  const int n_bh = n_items / n_qt;
  const int G = (int)gridDim.x;

  if (tid == 0) {
    if (raw & 1023u) __trap();  // the 128-byte swizzle needs 1024-byte aligned tiles
    tma_prefetch_desc(&tmQKV);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int i = 0; i < NK; ++i) { mbar_init(k_full + 8 * i, 1); mbar_init(k_empty + 8 * i, 1); }
    for (int i = 0; i < NV; ++i) { mbar_init(v_full + 8 * i, 1); mbar_init(v_empty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(s_ready + 8 * i, 1);
      mbar_init(p_ready + 8 * i, 4 * NG);
      mbar_init(o_free + 8 * i, 4 * NG);
    }
    mbar_init(o_done, 1);
    fence_mbar_init();
  }
  if (warp == TMA_WARP) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;

  Tracer tr;
  tr.init(warp == TMA_WARP ? 0 : warp == MMA_WARP ? 1 : warp == 0 ? 2 : warp == 4 * NG - 1 ? 3 : -1, lane);
  if (warp == TMA_WARP) {
    // ------------------------------------------------------------------------------------------- TMA producer
    if (lane == 0) {
      // three independent request streams over this CTA's items (Q per item, K_j and V_j per key block); each advances
      // when its slot is free — polled, never blocking on one
      int wq = (int)blockIdx.x, itq = 0;
      int wk = (int)blockIdx.x, jk = 0, kc = 0;
      int wv = (int)blockIdx.x, jv = 0, vc = 0;
      while (wq < n_items || wk < n_items || wv < n_items) {
        if (wq < n_items && mbar_try_wait(q_empty, ((uint32_t)itq & 1u) ^ 1u)) {
          const int qt = n_qt - 1 - wq / n_bh, bh = wq % n_bh;
          mbar_expect_tx(q_full, C::TILE);
          ws_load_tile<DH>(sQ, &tmQKV, q_full, 0 * H + bh % H, qt * 128, bh / H);
          tr.ev(1, qt + 1);
          wq += G; ++itq;
        }
        if (wk < n_items && mbar_try_wait(k_empty + 8 * (kc % NK), ((uint32_t)(kc / NK) & 1u) ^ 1u)) {
          const int qt = n_qt - 1 - wk / n_bh, bh = wk % n_bh, sk = kc % NK;
          mbar_expect_tx(k_full + 8 * sk, C::TILE);
          ws_load_tile<DH>(sK + sk * C::TILE, &tmQKV, k_full + 8 * sk, 1 * H + bh % H, jk * 128, bh / H);
          tr.ev(2, jk);
          ++kc;
          if (++jk > qt) { jk = 0; wk += G; }
        }
        if (wv < n_items && mbar_try_wait(v_empty + 8 * (vc % NV), ((uint32_t)(vc / NV) & 1u) ^ 1u)) {
          const int qt = n_qt - 1 - wv / n_bh, bh = wv % n_bh, sv = vc % NV;
          mbar_expect_tx(v_full + 8 * sv, C::TILE);
          ws_load_tile<DH>(sV + sv * C::TILE, &tmQKV, v_full + 8 * sv, 2 * H + bh % H, jv * 128, bh / H);
          tr.ev(3, jv);
          ++vc;
          if (++jv > qt) { jv = 0; wv += G; }
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ------------------------------------------------------------------------------------------- MMA issuer
    // The whole warp runs this loop convergently (waits, descriptor arithmetic in uniform registers); only the tensor
    // instructions are predicated on the leader lane.
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);  // S = Q K^T : both K-major (K = dh)
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, DH, 0, 1);   // O = P V   : A from TMEM, B MN-major (K = keys)
    const uint64_t dq = desc_k_base(sQ);
    uint32_t kc = 0, vc = 0, sc = 0, pc = 0;  // K uses, V uses, logits products issued, P.V products issued
    int it = 0;
    for (int w = (int)blockIdx.x; w < n_items; w += G, ++it) {
      const int n_kv = n_qt - w / n_bh;       // key blocks 0 .. qt
      const uint32_t tO = tmem + 256 + (it & 1) * DH;
      auto issue_s = [&](int j) {
        const uint32_t sk = kc % NK;
        tr.ev(9, j);    // about to wait for K_j
        mbar_wait(k_full + 8 * sk, (kc / NK) & 1u);
        tc_fence_after();
        const uint64_t dk = desc_k_base(sK + sk * C::TILE);
        const uint32_t tS = tmem + (sc & 1u) * 128;
        if (elect_one_sync()) {
#pragma unroll
          for (int kk = 0; kk < DH / 16; ++kk)
            umma_bf16_ss(tS, dq + kstep_k(kk), dk + kstep_k(kk), idesc_s, kk > 0 ? 1u : 0u);
          umma_commit(s_ready + 8 * (sc & 1u));
          umma_commit(k_empty + 8 * sk);
          if (j == n_kv - 1) umma_commit(q_empty);   // the item's last logits product: Q is free when it retires
        }
        __syncwarp();
        ++kc; ++sc;
        tr.ev(4, j);    // S_j issued
      };
      mbar_wait(q_full, (uint32_t)it & 1u);
      tr.ev(1, n_kv);
      issue_s(0);
      if (n_kv > 1) issue_s(1);
      for (int j = 0; j < n_kv; ++j) {
        const uint32_t sv = vc % NV;
        mbar_wait(v_full + 8 * sv, (vc / NV) & 1u);
        tr.ev(11, j);   // V_j has landed
        mbar_wait(p_ready + 8 * (pc & 1u), (pc >> 1) & 1u);
        tr.ev(12, j);   // P_j is in TMEM
        if (j == 0) mbar_wait(o_free + 8 * (it & 1), ((uint32_t)(it >> 1) & 1u) ^ 1u);  // the epilogue of item it-2 is done
        tc_fence_after();
        const uint64_t dv = desc_mn_base(sV + sv * C::TILE);
        const uint32_t tP = tmem + (pc & 1u) * 128;
        if (elect_one_sync()) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_bf16_ts(tO, tP + ts_split_col<NG>(kk), dv + kstep_mn(kk), idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
          umma_commit(o_done);
          umma_commit(v_empty + 8 * sv);
        }
        __syncwarp();
        ++vc; ++pc;
        tr.ev(5, j);    // P.V_j issued
        if (j + 2 < n_kv) issue_s(j + 2);  // into the buffer whose P the product above has just been queued to consume
      }
    }
  } else {
    // ------------------------------------------------------------------------------------------- softmax warpgroups
    const int g = warp >> 2;                       // column group of every key block / of the output
    const int row = tid & 127;                     // query row inside the tile = TMEM lane
    const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;
    const float c1 = scale * LOG2E_F;
    uint32_t bc = 0, xc = 0;   // key blocks processed (S / P buffer + o_done parity), exchange-buffer uses
    int it = 0;
    for (int w = (int)blockIdx.x; w < n_items; w += G, ++it) {
      const int qt = n_qt - 1 - w / n_bh, bh = w % n_bh;
      const int h = bh % H, b = bh / H;
      const int n_kv = qt + 1;
      const int qi = qt * 128 + row;
      const uint32_t tO = tmem + 256 + (it & 1) * DH;
      float m_used = -INFINITY;  // the maximum the accumulated P / O / l are scaled by (identical in all groups)
      float l_run = 0.f;         // partial row sum over this group's key columns
      tr.ev(1, n_kv);
      for (int j = 0; j < n_kv; ++j, ++bc) {
        const uint32_t sb = bc & 1u;
        const uint32_t tS = tmem + sb * 128 + CG * g + lane_off;
        mbar_wait(s_ready + 8 * sb, (bc >> 1) & 1u);
        tr.ev(6, j);      // S_j complete (seen by this math warp)
        tc_fence_after();
        if (xflags & 2) {  // experiment: pure hand-off chain, no softmax work at all
          warp_arrive(p_ready + 8 * sb, lane);
          continue;
        }
        uint32_t sv[CG];
#pragma unroll
        for (int c = 0; c < CG / 32; ++c) tmem_ld_x32(tS + c * 32, sv + c * 32);
        tmem_ld_wait();
        tr.ev(20, j);     // logits in registers
        if (j == qt) {  // diagonal block: keys after the query are masked (src/dalle_mtf/models.py:221-227)
          const int lim = row - CG * g;  // columns c > lim of this slice are in the future
#pragma unroll
          for (int c = 0; c < CG; ++c)
            if (c > lim) sv[c] = 0xff800000u;  // -inf
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int c = 0; c < CG; c += 4) {
          mx0 = fmaxf(mx0, __uint_as_float(sv[c]));
          mx1 = fmaxf(mx1, __uint_as_float(sv[c + 1]));
          mx2 = fmaxf(mx2, __uint_as_float(sv[c + 2]));
          mx3 = fmaxf(mx3, __uint_as_float(sv[c + 3]));
        }
        float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        float* xp = xch + (xc & 1u) * (NG * 128);
        ++xc;
        if (!(xflags & 4)) {  // (experiment bit 2: no cross-group exchange)
          xp[g * 128 + row] = mx;
          math_bar_sync<NG>();  // all groups' maxima visible; every group holds its S values in registers
#pragma unroll
          for (int o = 1; o < NG; ++o) mx = fmaxf(mx, xp[((g + o) % NG) * 128 + row]);  // finite: key 0 is always visible
        }
        tr.ev(21, j);     // row maximum known (after the cross-group exchange)
        if (j == 0) {
          m_used = mx;
        } else {
          const bool need = (mx - m_used) * c1 > 8.f;  // same decision in every thread of a row
          if (__any_sync(0xffffffffu, need)) {         // tcgen05.ld / st are warp-collective
            mbar_wait(o_done, (bc - 1u) & 1u);          // the previous P.V has landed in O
            tc_fence_after();
            const float alpha = need ? ex2f((m_used - mx) * c1) : 1.f;
            constexpr int W = OG >= 32 ? 32 : 16;  // this group's share of the output columns
#pragma unroll
            for (int c = 0; c < OG / W; ++c) {
              uint32_t r[W];
              const uint32_t ta = tO + lane_off + g * OG + c * W;
              tmem_ld_n<W>(ta, r);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < W; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
              tmem_st_n<W>(ta, r);
            }
            if (need) {
              l_run *= alpha;
              m_used = mx;
            }
          }
        }
        // P_j = 2^(c1 (s - m_used)) -> bf16 pairs -> the first CG/2 columns of this group's slice of the S buffer
        const float mc = m_used * c1;
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int c = 0; c < CG / 32; ++c) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float p0 = fmaf(__uint_as_float(sv[c * 32 + i]), c1, -mc);
            float p1 = fmaf(__uint_as_float(sv[c * 32 + i + 1]), c1, -mc);
            if (!(xflags & 1)) { p0 = ex2f(p0); p1 = ex2f(p1); }  // (experiment bit 0: no MUFU)
            l0 += p0;
            l1 += p1;
            pk[i >> 1] = pack_bf16x2(p0, p1);
          }
          tmem_st_x16(tS + c * 16, pk);
        }
        l_run += l0 + l1;
        tr.ev(22, j);     // exponentials computed, P stores issued
        tmem_st_wait();
        warp_arrive(p_ready + 8 * sb, lane);
        tr.ev(7, j);      // this warp's share of P_j written
      }
      // ---- epilogue: O / l -> bf16 (this group's share of the columns), lse.  The exchange slot is the one the last
      // block did not use; the next block (of the next item) takes the other one, so every reuse of a slot is separated
      // from its last readers by a math_bar_sync.
      float* xp = xch + (xc & 1u) * (NG * 128);
      ++xc;
      if (xflags & 2) m_used = 0.f;
      xp[g * 128 + row] = l_run;
      mbar_wait(o_done, (bc - 1u) & 1u);   // the item's last P.V
      tr.ev(8, 0);        // last P.V retired: epilogue starts
      tc_fence_after();
      math_bar_sync<NG>();
      float l_tot = 0.f;
#pragma unroll
      for (int o = 0; o < NG; ++o) l_tot += xp[o * 128 + row];  // same order in every group: identical 1 / l
      const float inv = 1.f / l_tot;
      bf16* op = out + (((long long)b * S + qi) * H + h) * DH + g * OG;
      constexpr int W = OG >= 32 ? 32 : 16;
#pragma unroll
      for (int c = 0; c < OG / W; ++c) {
        uint32_t r[W];
        tmem_ld_n<W>(tO + lane_off + g * OG + c * W, r);
        tmem_ld_wait();
        if (qi < S) store_cols_bf16<W>(op + c * W, r, inv);
      }
      warp_arrive(o_free + 8 * (it & 1), lane);   // this O accumulator may be overwritten (item it + 2)
      if (qi < S && g == 0) lse_out[((long long)b * H + h) * S + qi] = m_used * scale + logf(l_tot);
      tr.ev(10, 0);       // item done
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
  tr.ev(13, 0);         // CTA about to exit
}

// ------------------------------------------------------------------------------------------------------------------
// forward, two query tiles per work item — DEVELOPMENT builds only (make DEV=1, DB200_ATTN_FWD2=1): numerically validated
// (same 65 diagnostic cases), not faster than the kernel above on B200 (96 vs 91 us at (32,1280,4,128)): with one
// warpgroup per tile each scheduler holds a single math warp of the tile that is in its exponentials, and the chain of
// dependent FFMA -> MUFU -> FADD per row is latency-bound; the shipped library does not contain it.
// ------------------------------------------------------------------------------------------------------------------
#ifdef DB200_DEV_KNOBS
// The event timeline of the kernel above shows the softmax chain as the critical path: per 128-key block the math warps
// spend ~300 cycles waiting for the logits, ~500 reading them and exchanging row maxima across the column groups and
// ~1 400 in the exponentials (MUFU runs at 16 / cycle / SM, i.e. 1 024 cycles for a 128 x 128 block) — every warp in
// the same phase at the same time, the tensor pipe (1 024 cycles of products per block) waiting for P.
// Here a work item is TWO consecutive 128-query tiles (A, B) of one (batch, head) that share every K / V tile, and each
// softmax warpgroup owns one tile: thread r of warpgroup x owns row r of tile x with all its 128 key columns, so the
// row maximum and the row sum are thread-local (no shared-memory exchange, no named barrier) and the two warpgroups are
// naturally out of phase — while A's rows are in the exponentials, B's logits / P.V products run, and vice versa:
//     S_A0 S_B0 | PV_A0 S_A1 | PV_B0 S_B1 | PV_A1 S_A2 | ...
// TMEM: S_A [0,128) S_B [128,256) O_A [256,256+dh) O_B [384,384+dh); P overwrites the first 64 columns of its own S
// buffer (the logits are read twice: a max pass and an exp pass of 32 columns at a time, the bf16 P chunk trailing the
// read position).  Persistent: CTA c walks the items c, c + G, ... (item w = tile pair n_pairs-1 - w / (B H), heaviest
// first); counters: kc / vc ring uses (one K and one V tile per key block, shared by the two tiles), cx key blocks of
// tile x (parity of s_ready[x], p_ready[x], o_done[x]), ix items in which tile x existed (o_free[x]), it items (Q pair).
namespace {
template <int DH>
struct Fwd2Ws {
  static constexpr int NK = (DH == 128) ? 3 : 4;
  static constexpr int NV = (DH == 128) ? 2 : 4;
  static constexpr uint32_t TILE = 128 * DH * 2;
  static constexpr uint32_t BAR_BYTES = 256;
  static constexpr size_t SMEM = 2 * TILE + (NK + NV) * TILE + BAR_BYTES;
  static_assert(SMEM <= 232448, "forward attention: shared memory over the 227 KiB per-CTA limit");
  static constexpr int THREADS = 320;   // 2 softmax warpgroups + TMA warp + MMA warp
};
}  // namespace

template <int DH>
__global__ void __launch_bounds__(320, 1)
attn_fwd2_ws_kernel(const __grid_constant__ CUtensorMap tmQKV, bf16* __restrict__ out, float* __restrict__ lse_out,
                    int S, int H, int n_items, float scale) {
  using C = Fwd2Ws<DH>;
  constexpr int NK = C::NK, NV = C::NV;
  constexpr int TMA_WARP = 8, MMA_WARP = 9;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sQ = raw;                       // [2] Q_A, Q_B
  const uint32_t sK = sQ + 2 * C::TILE, sV = sK + NK * C::TILE;
  const uint32_t bars = sV + NV * C::TILE;
  const uint32_t q_full = bars, q_empty = bars + 8;
  const uint32_t k_full = bars + 16;               // [NK]
  const uint32_t k_empty = k_full + 8 * NK;        // [NK]
  const uint32_t v_full = k_empty + 8 * NK;        // [NV]
  const uint32_t v_empty = v_full + 8 * NV;        // [NV]
  const uint32_t s_ready = v_empty + 8 * NV;       // [2] per tile
  const uint32_t p_ready = s_ready + 16;           // [2]
  const uint32_t o_done = p_ready + 16;            // [2]
  const uint32_t o_free = o_done + 16;             // [2]
  const uint32_t tmem_slot = o_free + 16;
  static_assert(8 * (2 + 2 * NK + 2 * NV + 8) + 8 <= C::BAR_BYTES, "barrier area");
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_qt = (S + 127) >> 7;
  const int n_pairs = (n_qt + 1) >> 1;
  const int n_bh = n_items / n_pairs;
  const int G = (int)gridDim.x;

  if (tid == 0) {
    if (raw & 1023u) __trap();  // the 128-byte swizzle needs 1024-byte aligned tiles
    tma_prefetch_desc(&tmQKV);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int i = 0; i < NK; ++i) { mbar_init(k_full + 8 * i, 1); mbar_init(k_empty + 8 * i, 1); }
    for (int i = 0; i < NV; ++i) { mbar_init(v_full + 8 * i, 1); mbar_init(v_empty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(s_ready + 8 * i, 1);
      mbar_init(p_ready + 8 * i, 4);
      mbar_init(o_done + 8 * i, 1);
      mbar_init(o_free + 8 * i, 4);
    }
    fence_mbar_init();
  }
  if (warp == TMA_WARP) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;

  if (warp == TMA_WARP) {
    // ------------------------------------------------------------------------------------------- TMA producer
    if (lane == 0) {
      int wq = (int)blockIdx.x, itq = 0;
      int wk = (int)blockIdx.x, jk = 0, kc = 0;
      int wv = (int)blockIdx.x, jv = 0, vc = 0;
      while (wq < n_items || wk < n_items || wv < n_items) {
        if (wq < n_items && mbar_try_wait(q_empty, ((uint32_t)itq & 1u) ^ 1u)) {
          const int pt = n_pairs - 1 - wq / n_bh, bh = wq % n_bh;
          const bool has_b = 2 * pt + 1 < n_qt;
          mbar_expect_tx(q_full, has_b ? 2 * C::TILE : C::TILE);
          ws_load_tile<DH>(sQ, &tmQKV, q_full, 0 * H + bh % H, (2 * pt) * 128, bh / H);
          if (has_b) ws_load_tile<DH>(sQ + C::TILE, &tmQKV, q_full, 0 * H + bh % H, (2 * pt + 1) * 128, bh / H);
          wq += G; ++itq;
        }
        if (wk < n_items && mbar_try_wait(k_empty + 8 * (kc % NK), ((uint32_t)(kc / NK) & 1u) ^ 1u)) {
          const int pt = n_pairs - 1 - wk / n_bh, bh = wk % n_bh, sk = kc % NK;
          const int nmax = (2 * pt + 1 < n_qt) ? 2 * pt + 2 : 2 * pt + 1;   // key blocks of the pair
          mbar_expect_tx(k_full + 8 * sk, C::TILE);
          ws_load_tile<DH>(sK + sk * C::TILE, &tmQKV, k_full + 8 * sk, 1 * H + bh % H, jk * 128, bh / H);
          ++kc;
          if (++jk >= nmax) { jk = 0; wk += G; }
        }
        if (wv < n_items && mbar_try_wait(v_empty + 8 * (vc % NV), ((uint32_t)(vc / NV) & 1u) ^ 1u)) {
          const int pt = n_pairs - 1 - wv / n_bh, bh = wv % n_bh, sv = vc % NV;
          const int nmax = (2 * pt + 1 < n_qt) ? 2 * pt + 2 : 2 * pt + 1;
          mbar_expect_tx(v_full + 8 * sv, C::TILE);
          ws_load_tile<DH>(sV + sv * C::TILE, &tmQKV, v_full + 8 * sv, 2 * H + bh % H, jv * 128, bh / H);
          ++vc;
          if (++jv >= nmax) { jv = 0; wv += G; }
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ------------------------------------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);  // S = Q K^T : both K-major (K = dh)
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, DH, 0, 1);   // O = P V   : A from TMEM, B MN-major (K = keys)
    uint32_t kc = 0, vc = 0;          // ring uses at the start of the current item
    uint32_t cx[2] = {0, 0};          // P.V products issued per tile
    uint32_t ix[2] = {0, 0};          // items in which the tile existed
    int it = 0;
    for (int w = (int)blockIdx.x; w < n_items; w += G, ++it) {
      const int pt = n_pairs - 1 - w / n_bh;
      const bool has_b = 2 * pt + 1 < n_qt;
      const int nx[2] = {2 * pt + 1, has_b ? 2 * pt + 2 : 0};   // key blocks per tile
      const int nmax = has_b ? nx[1] : nx[0];
      const int last_tile_of = has_b ? 1 : 0;                   // for every block j < nmax the last active tile is this one
      // logits of block j for tile x; releases K_j when x is the last tile that reads it
      auto issue_s = [&](int x, int j) {
        const uint32_t kuse = kc + (uint32_t)j, sk = kuse % NK;
        mbar_wait(k_full + 8 * sk, (kuse / NK) & 1u);
        tc_fence_after();
        const uint64_t dq = desc_k_base(sQ + x * C::TILE), dk = desc_k_base(sK + sk * C::TILE);
        const uint32_t tS = tmem + x * 128;
        if (elect_one_sync()) {
#pragma unroll
          for (int kk = 0; kk < DH / 16; ++kk)
            umma_bf16_ss(tS, dq + kstep_k(kk), dk + kstep_k(kk), idesc_s, kk > 0 ? 1u : 0u);
          umma_commit(s_ready + 8 * x);
          if (x == last_tile_of) {
            umma_commit(k_empty + 8 * sk);
            if (j == nmax - 1) umma_commit(q_empty);   // the item's last logits product: the Q pair may be replaced
          }
        }
        __syncwarp();
      };
      mbar_wait(q_full, (uint32_t)it & 1u);
      issue_s(0, 0);
      if (has_b) issue_s(1, 0);
      for (int j = 0; j < nmax; ++j) {
        const uint32_t vuse = vc + (uint32_t)j, sv = vuse % NV;
#pragma unroll
        for (int x = 0; x < 2; ++x) {
          if (j >= nx[x]) continue;    // tile A has one block less than tile B
          mbar_wait(v_full + 8 * sv, (vuse / NV) & 1u);
          mbar_wait(p_ready + 8 * x, cx[x] & 1u);
          if (j == 0) mbar_wait(o_free + 8 * x, (ix[x] & 1u) ^ 1u);   // the previous epilogue of this tile slot has read O
          tc_fence_after();
          const uint64_t dv = desc_mn_base(sV + sv * C::TILE);
          const uint32_t tP = tmem + x * 128, tO = tmem + 256 + x * 128;
          if (elect_one_sync()) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_bf16_ts(tO, tP + kk * 8, dv + kstep_mn(kk), idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
            umma_commit(o_done + 8 * x);
            if (x == last_tile_of) umma_commit(v_empty + 8 * sv);
          }
          __syncwarp();
          ++cx[x];
          if (j + 1 < nx[x]) issue_s(x, j + 1);   // overwrites P_j of this tile, which the product above has consumed
        }
      }
      kc += (uint32_t)nmax;
      vc += (uint32_t)nmax;
      ++ix[0];
      if (has_b) ++ix[1];
    }
  } else {
    // ------------------------------------------------------------------------------------------- softmax warpgroups
    const int x = warp >> 2;                       // tile A (warps 0-3) or B (warps 4-7)
    const int row = tid & 127;                     // query row inside the tile = TMEM lane
    const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;
    const uint32_t tS = tmem + x * 128 + lane_off, tO = tmem + 256 + x * 128 + lane_off;
    const float c1 = scale * LOG2E_F;
    uint32_t cb = 0;   // key blocks of this tile slot processed
    for (int w = (int)blockIdx.x; w < n_items; w += G) {
      const int pt = n_pairs - 1 - w / n_bh, bh = w % n_bh;
      const int qt = 2 * pt + x;
      if (qt >= n_qt) continue;                    // odd number of tiles: the last pair has no tile B
      const int h = bh % H, b = bh / H;
      const int n_kv = qt + 1;
      const int qi = qt * 128 + row;
      float m_used = -INFINITY, l_run = 0.f;
      for (int j = 0; j < n_kv; ++j, ++cb) {
        mbar_wait(s_ready + 8 * x, cb & 1u);
        tc_fence_after();
        const int lim = (j == qt) ? row : 128;     // diagonal block: columns c > lim are keys after this query
        // ---- pass 1: row maximum
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t rs[32];
          tmem_ld_x32(tS + c * 32, rs);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const int cc = c * 32 + i;
            mx0 = fmaxf(mx0, cc <= lim ? __uint_as_float(rs[i]) : -INFINITY);
            mx1 = fmaxf(mx1, cc + 1 <= lim ? __uint_as_float(rs[i + 1]) : -INFINITY);
            mx2 = fmaxf(mx2, cc + 2 <= lim ? __uint_as_float(rs[i + 2]) : -INFINITY);
            mx3 = fmaxf(mx3, cc + 3 <= lim ? __uint_as_float(rs[i + 3]) : -INFINITY);
          }
        }
        const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));   // finite: column 0 is always visible
        if (j == 0) {
          m_used = mx;
        } else {
          const bool need = (mx - m_used) * c1 > 8.f;
          if (__any_sync(0xffffffffu, need)) {          // tcgen05.ld / st are warp-collective
            mbar_wait(o_done + 8 * x, (cb - 1u) & 1u);  // the previous P.V of this tile has landed in O
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
        // ---- pass 2: P = 2^(c1 (s - m_used)) -> bf16 pairs, 16 packed columns per 32 logits, trailing the read position
        const float mc = m_used * c1;
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t rs[32], pk[16];
          tmem_ld_x32(tS + c * 32, rs);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const int cc = c * 32 + i;
            float p0 = ex2f(fmaf(__uint_as_float(rs[i]), c1, -mc));
            float p1 = ex2f(fmaf(__uint_as_float(rs[i + 1]), c1, -mc));
            if (cc > lim) p0 = 0.f;
            if (cc + 1 > lim) p1 = 0.f;
            l0 += p0;
            l1 += p1;
            pk[i >> 1] = pack_bf16x2(p0, p1);
          }
          tmem_st_x16(tS + c * 16, pk);
        }
        l_run += l0 + l1;
        tmem_st_wait();
        warp_arrive(p_ready + 8 * x, lane);
      }
      // ---- epilogue: O / l -> bf16, lse (both thread-local)
      mbar_wait(o_done + 8 * x, (cb - 1u) & 1u);   // the tile's last P.V
      tc_fence_after();
      const float inv = 1.f / l_run;
      bf16* op = out + (((long long)b * S + qi) * H + h) * DH;
#pragma unroll
      for (int c = 0; c < DH / 32; ++c) {
        uint32_t r[32];
        tmem_ld_x32(tO + c * 32, r);
        tmem_ld_wait();
        if (qi < S) store_cols_bf16<32>(op + c * 32, r, inv);
      }
      warp_arrive(o_free + 8 * x, lane);   // the O accumulator may be overwritten by the next item's first P.V
      if (qi < S) lse_out[((long long)b * H + h) * S + qi] = m_used * scale + logf(l_run);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}
#endif  // DB200_DEV_KNOBS

// ------------------------------------------------------------------------------------------------------------------
// backward
//   p  = exp(scale*s - lse)   (0 where key > query or the query is out of range)
//   ds = p * (dp - delta) * scale
// ------------------------------------------------------------------------------------------------------------------
namespace {
template <int DH, int NG>
struct BwdWs {
  // two streamed [128][DH] tiles per block live in separate rings: the one that is needed until the LAST product of its
  // block (Q_i for dK, K_j for dQ) three deep, the one that is released early (dO_i after dV, V_j after dP) two deep
  static constexpr int NA = 3, NB = 2;
  static constexpr uint32_t TILE = 128 * DH * 2;
  static constexpr uint32_t STAT_BYTES = 2 * 256 * 4;
  static constexpr uint32_t BAR_BYTES = 256;
  // dynamic shared memory is declared 1024-byte aligned (the 128-byte swizzle needs it): no alignment slack
  static constexpr size_t SMEM = 2 * TILE + (NA + NB) * TILE + STAT_BYTES + BAR_BYTES;
  static constexpr int THREADS = (4 * NG + 4) * 32;  // NG math warpgroups + TMA, MMA, stager (+ one idle) warps
};
}  // namespace

// dK / dV: work item = 128 keys (TMEM lanes) of one (batch, head), loop over the 128-query blocks i >= its own.
// TMEM: dV [0,dh) dK [dh,2dh) S^T [256,384) dP^T [384,512).
// Persistent like the forward: CTA c walks items c, c + G, ... of the list ordered by decreasing work (item w = key
// block w / (B H), heaviest = block 0).  Global use counters: ac / bc2 (Q_i / dO_i ring uses), gb (query blocks processed:
// S^T / dP^T / P^T / dS^T hand-offs have parity gb & 1, the lse / delta staging slot is gb & 1), itm (items: resident
// K | V pair and the accumulators, parity itm & 1).  The accumulators are single-buffered (TMEM is full): the first dV
// product of an item waits for the previous epilogue (acc_free); the item's logits products and phase A run before that.
template <int DH, int NG>
__global__ void __launch_bounds__((4 * NG + 4) * 32, 1)
attn_bwd_dkdv_ws_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                        const float* __restrict__ lse, const float* __restrict__ delta, bf16* __restrict__ dqkv,
                        int S, int H, int n_items, float scale) {
  using C = BwdWs<DH, NG>;
  constexpr int NA = C::NA, NB = C::NB;
  constexpr int CG = 128 / NG;       // query columns of a block per group (= per thread)
  constexpr int TMA_WARP = 4 * NG, MMA_WARP = 4 * NG + 1, STAT_WARP = 4 * NG + 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = raw;                    // 1024-byte aligned (checked below)
  const uint32_t sK = base, sV = sK + C::TILE;
  const uint32_t sQr = sV + C::TILE;            // Q_i ring (NA deep): needed until dK of its block
  const uint32_t sOr = sQr + NA * C::TILE;      // dO_i ring (NB deep): released after dV of its block
  const uint32_t sStat = sOr + NB * C::TILE;    // [2][lse2 128 | delta 128]
  const uint32_t bars = sStat + C::STAT_BYTES;
  const uint32_t x_full = bars;                 // this item's K | V pair has landed
  const uint32_t x_empty = bars + 8;            // every logits product that reads it has retired
  const uint32_t a_full = bars + 16;            // [NA]
  const uint32_t a_empty = a_full + 8 * NA;     // [NA]
  const uint32_t b_full = a_empty + 8 * NA;     // [NB]
  const uint32_t b_empty = b_full + 8 * NB;     // [NB]
  const uint32_t stat_full = b_empty + 8 * NB;  // [2]
  const uint32_t sa_ready = stat_full + 16;     // S^T of the block in TMEM
  const uint32_t sb_ready = sa_ready + 8;       // dP^T
  const uint32_t pa_ready = sb_ready + 8;       // P^T written (all math warps)
  const uint32_t pb_ready = pa_ready + 8;       // dS^T written
  const uint32_t acc_done = pb_ready + 8;       // the item's last dK product has retired
  const uint32_t acc_free = acc_done + 8;       // the epilogue has read dV | dK (all math warps)
  const uint32_t tmem_slot = acc_free + 8;
  static_assert(8 * (2 + 2 * NA + 2 * NB + 2 + 4 + 2) + 8 <= C::BAR_BYTES, "barrier area");
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  const float* stat = reinterpret_cast<const float*>(smem_raw + (sStat - raw));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_blk = (S + 127) >> 7;
  const int n_bh = n_items / n_blk;
  const int G = (int)gridDim.x;

  if (tid == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    if (raw & 1023u) __trap();  // the 128-byte swizzle needs 1024-byte aligned tiles
    mbar_init(x_full, 1);
    mbar_init(x_empty, 1);
    for (int i = 0; i < NA; ++i) { mbar_init(a_full + 8 * i, 1); mbar_init(a_empty + 8 * i, 1); }
    for (int i = 0; i < NB; ++i) { mbar_init(b_full + 8 * i, 1); mbar_init(b_empty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) mbar_init(stat_full + 8 * i, 1);
    mbar_init(sa_ready, 1); mbar_init(sb_ready, 1);
    mbar_init(pa_ready, 4 * NG); mbar_init(pb_ready, 4 * NG);
    mbar_init(acc_done, 1);
    mbar_init(acc_free, 4 * NG);
    fence_mbar_init();
  }
  if (warp == TMA_WARP) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const uint32_t tST = tmem + 256, tdPT = tmem + 384;

  Tracer tr;
  tr.init(warp == TMA_WARP ? 0 : warp == MMA_WARP ? 1 : warp == 0 ? 2 : warp == 4 * NG - 1 ? 3 : -1, lane);
  if (warp == TMA_WARP) {
    // ------------------------------------------------------------------------------------------- TMA producer
    if (lane == 0) {
      // three request streams over this CTA's items: the resident K | V pair, Q_i, dO_i — polled, never blocking
      int wx = (int)blockIdx.x, itx = 0;
      int wa = (int)blockIdx.x, ia = 0, ac = 0;
      int wb = (int)blockIdx.x, ib = 0, bc2 = 0;
      while (wx < n_items || wa < n_items || wb < n_items) {
        if (wx < n_items && mbar_try_wait(x_empty, ((uint32_t)itx & 1u) ^ 1u)) {
          const int jb = wx / n_bh, bh = wx % n_bh;
          mbar_expect_tx(x_full, 2 * C::TILE);
          ws_load_tile<DH>(sK, &tmQKV, x_full, 1 * H + bh % H, jb * 128, bh / H);
          ws_load_tile<DH>(sV, &tmQKV, x_full, 2 * H + bh % H, jb * 128, bh / H);
          tr.ev(1, n_blk - jb);
          wx += G; ++itx;
        }
        if (wa < n_items && mbar_try_wait(a_empty + 8 * (ac % NA), ((uint32_t)(ac / NA) & 1u) ^ 1u)) {
          const int jb = wa / n_bh, bh = wa % n_bh, st = ac % NA;
          mbar_expect_tx(a_full + 8 * st, C::TILE);
          ws_load_tile<DH>(sQr + st * C::TILE, &tmQKV, a_full + 8 * st, 0 * H + bh % H, (jb + ia) * 128, bh / H);   // Q_i
          tr.ev(2, ia);
          ++ac;
          if (++ia >= n_blk - jb) { ia = 0; wa += G; }
        }
        if (wb < n_items && mbar_try_wait(b_empty + 8 * (bc2 % NB), ((uint32_t)(bc2 / NB) & 1u) ^ 1u)) {
          const int jb = wb / n_bh, bh = wb % n_bh, st = bc2 % NB;
          mbar_expect_tx(b_full + 8 * st, C::TILE);
          ws_load_tile<DH>(sOr + st * C::TILE, &tmDO, b_full + 8 * st, bh % H, (jb + ib) * 128, bh / H);            // dO_i
          tr.ev(3, ib);
          ++bc2;
          if (++ib >= n_blk - jb) { ib = 0; wb += G; }
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ------------------------------------------------------------------------------------------- MMA issuer
    // convergent warp: waits and descriptor arithmetic by all lanes (uniform registers), tensor instructions predicated
    constexpr uint32_t idesc_l = umma_idesc_bf16(128, 128, 0, 0);  // logits: both operands K-major (K = dh)
    constexpr uint32_t idesc_g = umma_idesc_bf16(128, DH, 0, 1);   // gradients: A from TMEM, B MN-major (K = queries)
    const uint64_t dkk = desc_k_base(sK), dvk = desc_k_base(sV);
    uint32_t sa = 0, sb = 0;   // S^T / dP^T products issued (= Q_i / dO_i ring uses consumed by them)
    uint32_t gb = 0;           // query blocks whose gradient products were issued
    int itm = 0;
    for (int w = (int)blockIdx.x; w < n_items; w += G, ++itm) {
      const int n_it = n_blk - w / n_bh;
      auto issue_s = [&](int it) {   // S^T = K Q_i^T
        const uint32_t st = sa % NA;
        mbar_wait(a_full + 8 * st, (sa / NA) & 1u);
        tc_fence_after();
        const uint64_t q = desc_k_base(sQr + st * C::TILE);
        if (elect_one_sync()) {
#pragma unroll
          for (int kk = 0; kk < DH / 16; ++kk)
            umma_bf16_ss(tST, dkk + kstep_k(kk), q + kstep_k(kk), idesc_l, kk > 0 ? 1u : 0u);
          umma_commit(sa_ready);
        }
        __syncwarp();
        ++sa;
        tr.ev(4, it);    // S^T issued
      };
      auto issue_dp = [&](int it) {  // dP^T = V dO_i^T
        const uint32_t st = sb % NB;
        mbar_wait(b_full + 8 * st, (sb / NB) & 1u);
        tc_fence_after();
        const uint64_t o = desc_k_base(sOr + st * C::TILE);
        if (elect_one_sync()) {
#pragma unroll
          for (int kk = 0; kk < DH / 16; ++kk)
            umma_bf16_ss(tdPT, dvk + kstep_k(kk), o + kstep_k(kk), idesc_l, kk > 0 ? 1u : 0u);
          umma_commit(sb_ready);
          if (it == n_it - 1) umma_commit(x_empty);   // the item's last logits product: K | V may be replaced
        }
        __syncwarp();
        ++sb;
        tr.ev(14, it);   // dP^T issued
      };
      mbar_wait(x_full, (uint32_t)itm & 1u);
      tr.ev(1, n_it);
      issue_s(0);
      issue_dp(0);
      for (int it = 0; it < n_it; ++it, ++gb) {
        const uint32_t sta = gb % NA, stb = gb % NB;   // ring slots of this block's Q_i / dO_i (use counter = block counter)
        const uint64_t q = desc_mn_base(sQr + sta * C::TILE), o = desc_mn_base(sOr + stb * C::TILE);
        const uint32_t acc = it > 0 ? 1u : 0u;
        mbar_wait(pa_ready, gb & 1u);
        tr.ev(12, it);   // P^T is in TMEM
        if (it == 0) mbar_wait(acc_free, ((uint32_t)itm & 1u) ^ 1u);   // the previous item's epilogue has read dV | dK
        tc_fence_after();
        if (elect_one_sync()) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)  // dV += P^T dO_i
            umma_bf16_ts(tmem, tST + ts_split_col<NG>(kk), o + kstep_mn(kk), idesc_g, kk > 0 ? 1u : acc);
          umma_commit(b_empty + 8 * stb);  // dO_i is free
        }
        __syncwarp();
        tr.ev(5, it);    // dV issued
        if (it + 1 < n_it) issue_s(it + 1);      // runs under phase B of this block
        mbar_wait(pb_ready, gb & 1u);
        tr.ev(15, it);   // dS^T is in TMEM
        tc_fence_after();
        if (elect_one_sync()) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)  // dK += dS^T Q_i
            umma_bf16_ts(tmem + DH, tdPT + ts_split_col<NG>(kk), q + kstep_mn(kk), idesc_g, kk > 0 ? 1u : acc);
          umma_commit(a_empty + 8 * sta);  // Q_i is free
          if (it == n_it - 1) umma_commit(acc_done);
        }
        __syncwarp();
        tr.ev(16, it);   // dK issued
        if (it + 1 < n_it) issue_dp(it + 1);     // runs under phase A of the next block
      }
    }
  } else if (warp == STAT_WARP) {
    // ------------------------------------------------------------------------------------------- lse / delta stager
    float* stat_w = reinterpret_cast<float*>(smem_raw + (sStat - raw));
    uint32_t gb = 0;
    for (int w = (int)blockIdx.x; w < n_items; w += G) {
      const int jb = w / n_bh, n_it = n_blk - jb;
      const long long bh = w % n_bh;   // = b * H + h
      for (int it = 0; it < n_it; ++it, ++gb) {
        const uint32_t st = gb & 1u;
        // slot `st` was last read by phase B of block gb-2, which precedes the dK product that frees that block's Q
        if (gb >= 2) mbar_wait(a_empty + 8 * ((gb - 2) % NA), ((gb - 2) / NA) & 1u);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = lane + 32 * u, q = (jb + it) * 128 + c;
          float l2 = INFINITY, d = 0.f;  // out-of-range query: p = 2^(-inf) = 0
          if (q < S) {
            l2 = lse[bh * S + q] * LOG2E_F;
            d = delta[bh * S + q];
          }
          stat_w[st * 256 + c] = l2;
          stat_w[st * 256 + 128 + c] = d;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(stat_full + 8 * st);
      }
    }
  } else if (warp < 4 * NG) {
    // ------------------------------------------------------------------------------------------- gradient warpgroups
    const int g = warp >> 2;            // query columns [CG g, CG g + CG) of every block
    const int row = tid & 127;          // key row = TMEM lane
    const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;
    const uint32_t tS = tST + CG * g + lane_off, tdP = tdPT + CG * g + lane_off;
    const float c1 = scale * LOG2E_F;
    uint32_t gb = 0;
    int itm = 0;
    for (int w = (int)blockIdx.x; w < n_items; w += G, ++itm) {
      const int jb = w / n_bh, n_it = n_blk - jb;
      const int h = (w % n_bh) % H, b = (w % n_bh) / H;
      const int ki = jb * 128 + row;
      tr.ev(1, n_it);
      for (int it = 0; it < n_it; ++it, ++gb) {
        const uint32_t st = gb & 1u;
        const float4* sl = reinterpret_cast<const float4*>(stat + st * 256 + CG * g);
        const float4* sd = reinterpret_cast<const float4*>(stat + st * 256 + 128 + CG * g);
        // ---- phase A: P^T = 2^(c1 s - lse2[query]); pair dropped where key > query (only the first block is diagonal)
        mbar_wait(stat_full + 8 * st, (gb >> 1) & 1u);
        mbar_wait(sa_ready, gb & 1u);
        tr.ev(6, it);
        tc_fence_after();
        uint32_t pk[CG / 2];
        const int lim = (it == 0) ? row - CG * g : -1;  // columns c < lim are queries before this key
#pragma unroll
        for (int c = 0; c < CG / 32; ++c) {
          uint32_t rs[32];
          tmem_ld_x32(tS + c * 32, rs);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 a = sl[(c * 32 + i) >> 2];
            float p0 = ex2f(fmaf(__uint_as_float(rs[i]), c1, -a.x)), p1 = ex2f(fmaf(__uint_as_float(rs[i + 1]), c1, -a.y));
            float p2 = ex2f(fmaf(__uint_as_float(rs[i + 2]), c1, -a.z)), p3 = ex2f(fmaf(__uint_as_float(rs[i + 3]), c1, -a.w));
            const int cc = c * 32 + i;
            if (cc < lim) p0 = 0.f;
            if (cc + 1 < lim) p1 = 0.f;
            if (cc + 2 < lim) p2 = 0.f;
            if (cc + 3 < lim) p3 = 0.f;
            pk[(cc >> 1)] = pack_bf16x2(p0, p1);
            pk[(cc >> 1) + 1] = pack_bf16x2(p2, p3);
          }
        }
        tmem_st_n<CG / 2>(tS, pk);
        tmem_st_wait();
        warp_arrive(pa_ready, lane);
        tr.ev(7, it);
        // ---- phase B: dS^T = (P^T * scale) (dP^T - delta[query])
        mbar_wait(sb_ready, gb & 1u);
        tr.ev(17, it);
        tc_fence_after();
        uint32_t dk[CG / 2];
#pragma unroll
        for (int c = 0; c < CG / 32; ++c) {
          uint32_t rd[32];
          tmem_ld_x32(tdP + c * 32, rd);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 d4 = sd[(c * 32 + i) >> 2];
            const int cc = c * 32 + i;
            const float2 pa = unpack_bf16x2(pk[cc >> 1]), pb = unpack_bf16x2(pk[(cc >> 1) + 1]);
            dk[cc >> 1] = pack_bf16x2((pa.x * scale) * (__uint_as_float(rd[i]) - d4.x),
                                      (pa.y * scale) * (__uint_as_float(rd[i + 1]) - d4.y));
            dk[(cc >> 1) + 1] = pack_bf16x2((pb.x * scale) * (__uint_as_float(rd[i + 2]) - d4.z),
                                            (pb.y * scale) * (__uint_as_float(rd[i + 3]) - d4.w));
          }
        }
        tmem_st_n<CG / 2>(tdP, dk);
        tmem_st_wait();
        warp_arrive(pb_ready, lane);
        tr.ev(18, it);
      }
      // ---- epilogue: the 2 dh accumulator columns (dV | dK) are split evenly over the groups
      mbar_wait(acc_done, (uint32_t)itm & 1u);
      tr.ev(8, 0);
      tc_fence_after();
      constexpr int EG = 2 * DH / NG;                 // columns per group
      const int acc_i = (g * EG) / DH;                // 0: dV, 1: dK
      const int col0 = (g * EG) % DH;
      bf16* dst = dqkv + ((((long long)b * S + ki) * 3 + (acc_i == 0 ? 2 : 1)) * H + h) * DH + col0;
#pragma unroll 1
      for (int c = 0; c < EG / 32; ++c) {
        uint32_t r[32];
        tmem_ld_x32(tmem + lane_off + g * EG + c * 32, r);
        tmem_ld_wait();
        if (ki < S) store_cols_bf16<32>(dst + c * 32, r, 1.f);
      }
      warp_arrive(acc_free, lane);   // dV | dK may be overwritten by the next item
      tr.ev(10, 0);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
  tr.ev(13, 0);
}

// dQ: work item = 128 queries (TMEM lanes) of one (batch, head), loop over the 128-key blocks j <= its own.
// TMEM: dQ [0,dh) S [128,256) dP [256,384) dS [384,448) [448,512).  Persistent (see the forward): item w = query tile
// n_blk-1 - w / (B H); global counters la (logits pairs issued = K_j / V_j ring uses), gb (key blocks whose dQ product
// was issued / processed: S | dP hand-offs have parity gb & 1, the dS slot is gb & 1), itm (items: resident Q | dO pair
// and the dQ accumulator).
template <int DH, int NG>
__global__ void __launch_bounds__((4 * NG + 4) * 32, 1)
attn_bwd_dq_ws_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                      const float* __restrict__ lse, const float* __restrict__ delta, bf16* __restrict__ dqkv, int S,
                      int H, int n_items, float scale) {
  using C = BwdWs<DH, NG>;
  constexpr int NA = C::NA, NB = C::NB;
  constexpr int CG = 128 / NG;       // key columns of a block per group (= per thread)
  constexpr int TMA_WARP = 4 * NG, MMA_WARP = 4 * NG + 1;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = raw;                     // 1024-byte aligned (checked below)
  const uint32_t sQ = base, sdO = sQ + C::TILE;
  const uint32_t sKr = sdO + C::TILE;            // K_j ring (NA deep): needed until dQ of its block
  const uint32_t sVr = sKr + NA * C::TILE;       // V_j ring (NB deep): released right after dP of its block
  const uint32_t bars = sVr + NB * C::TILE + C::STAT_BYTES;
  const uint32_t x_full = bars;                  // this item's Q | dO pair has landed
  const uint32_t x_empty = bars + 8;             // every logits product that reads it has retired
  const uint32_t a_full = bars + 16;             // [NA]
  const uint32_t a_empty = a_full + 8 * NA;      // [NA]
  const uint32_t b_full = a_empty + 8 * NA;      // [NB]
  const uint32_t b_empty = b_full + 8 * NB;      // [NB]
  const uint32_t sd_ready = b_empty + 8 * NB;    // S, dP of the block in TMEM
  const uint32_t sd_loaded = sd_ready + 8;       // every group holds them in registers (all math warps)
  const uint32_t ds_ready = sd_loaded + 8;       // [2] dS of the block written (all math warps)
  const uint32_t ds_free = ds_ready + 16;        // [2] the dQ product has consumed that dS slot
  const uint32_t acc_done = ds_free + 16;        // the item's last dQ product has retired
  const uint32_t acc_free = acc_done + 8;        // the epilogue has read dQ (all math warps)
  const uint32_t tmem_slot = acc_free + 8;
  static_assert(8 * (2 + 2 * NA + 2 * NB + 2 + 4 + 2) + 8 <= C::BAR_BYTES, "barrier area");
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_blk = (S + 127) >> 7;
  const int n_bh = n_items / n_blk;
  const int G = (int)gridDim.x;

  if (tid == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    if (raw & 1023u) __trap();  // the 128-byte swizzle needs 1024-byte aligned tiles
    mbar_init(x_full, 1);
    mbar_init(x_empty, 1);
    for (int i = 0; i < NA; ++i) { mbar_init(a_full + 8 * i, 1); mbar_init(a_empty + 8 * i, 1); }
    for (int i = 0; i < NB; ++i) { mbar_init(b_full + 8 * i, 1); mbar_init(b_empty + 8 * i, 1); }
    mbar_init(sd_ready, 1);
    mbar_init(sd_loaded, 4 * NG);
    for (int i = 0; i < 2; ++i) { mbar_init(ds_ready + 8 * i, 4 * NG); mbar_init(ds_free + 8 * i, 1); }
    mbar_init(acc_done, 1);
    mbar_init(acc_free, 4 * NG);
    fence_mbar_init();
  }
  if (warp == TMA_WARP) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const uint32_t tSb = tmem + 128, tdPb = tmem + 256, tdSb = tmem + 384;

  if (warp == TMA_WARP) {
    // ------------------------------------------------------------------------------------------- TMA producer
    if (lane == 0) {
      // three request streams over this CTA's items: the resident Q | dO pair, K_j, V_j — polled, never blocking
      int wx = (int)blockIdx.x, itx = 0;
      int wa = (int)blockIdx.x, ia = 0, ac = 0;
      int wb = (int)blockIdx.x, ib2 = 0, bc2 = 0;
      while (wx < n_items || wa < n_items || wb < n_items) {
        if (wx < n_items && mbar_try_wait(x_empty, ((uint32_t)itx & 1u) ^ 1u)) {
          const int ib = n_blk - 1 - wx / n_bh, bh = wx % n_bh;
          mbar_expect_tx(x_full, 2 * C::TILE);
          ws_load_tile<DH>(sQ, &tmQKV, x_full, 0 * H + bh % H, ib * 128, bh / H);
          ws_load_tile<DH>(sdO, &tmDO, x_full, bh % H, ib * 128, bh / H);
          wx += G; ++itx;
        }
        if (wa < n_items && mbar_try_wait(a_empty + 8 * (ac % NA), ((uint32_t)(ac / NA) & 1u) ^ 1u)) {
          const int ib = n_blk - 1 - wa / n_bh, bh = wa % n_bh, st = ac % NA;
          mbar_expect_tx(a_full + 8 * st, C::TILE);
          ws_load_tile<DH>(sKr + st * C::TILE, &tmQKV, a_full + 8 * st, 1 * H + bh % H, ia * 128, bh / H);   // K_j
          ++ac;
          if (++ia > ib) { ia = 0; wa += G; }
        }
        if (wb < n_items && mbar_try_wait(b_empty + 8 * (bc2 % NB), ((uint32_t)(bc2 / NB) & 1u) ^ 1u)) {
          const int ib = n_blk - 1 - wb / n_bh, bh = wb % n_bh, st = bc2 % NB;
          mbar_expect_tx(b_full + 8 * st, C::TILE);
          ws_load_tile<DH>(sVr + st * C::TILE, &tmQKV, b_full + 8 * st, 2 * H + bh % H, ib2 * 128, bh / H);  // V_j
          ++bc2;
          if (++ib2 > ib) { ib2 = 0; wb += G; }
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ------------------------------------------------------------------------------------------- MMA issuer
    // convergent warp: waits and descriptor arithmetic by all lanes (uniform registers), tensor instructions predicated
    constexpr uint32_t idesc_l = umma_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idesc_g = umma_idesc_bf16(128, DH, 0, 1);   // dQ = dS K : A from TMEM, B MN-major (K = keys)
    const uint64_t dqk = desc_k_base(sQ), dok = desc_k_base(sdO);
    uint32_t la = 0, gb = 0;
    int itm = 0;
    for (int w = (int)blockIdx.x; w < n_items; w += G, ++itm) {
      const int n_it = n_blk - w / n_bh;
      auto issue_l = [&](int j) {  // S = Q K_j^T, dP = dO V_j^T
        const uint32_t sa = la % NA, sb = la % NB;
        mbar_wait(a_full + 8 * sa, (la / NA) & 1u);
        mbar_wait(b_full + 8 * sb, (la / NB) & 1u);
        tc_fence_after();
        const uint64_t k = desc_k_base(sKr + sa * C::TILE), v = desc_k_base(sVr + sb * C::TILE);
        if (elect_one_sync()) {
#pragma unroll
          for (int kk = 0; kk < DH / 16; ++kk)
            umma_bf16_ss(tSb, dqk + kstep_k(kk), k + kstep_k(kk), idesc_l, kk > 0 ? 1u : 0u);
#pragma unroll
          for (int kk = 0; kk < DH / 16; ++kk)
            umma_bf16_ss(tdPb, dok + kstep_k(kk), v + kstep_k(kk), idesc_l, kk > 0 ? 1u : 0u);
          umma_commit(sd_ready);
          umma_commit(b_empty + 8 * sb);  // V_j is free as soon as dP has consumed it
          if (j == n_it - 1) umma_commit(x_empty);   // the item's last logits pair: Q | dO may be replaced
        }
        __syncwarp();
        ++la;
      };
      mbar_wait(x_full, (uint32_t)itm & 1u);
      issue_l(0);
      for (int j = 0; j < n_it; ++j, ++gb) {
        if (j + 1 < n_it) {  // the groups hold block j in registers: the next logits run under their arithmetic
          mbar_wait(sd_loaded, gb & 1u);
          issue_l(j + 1);
        }
        mbar_wait(ds_ready + 8 * (gb & 1u), (gb >> 1) & 1u);
        if (j == 0) mbar_wait(acc_free, ((uint32_t)itm & 1u) ^ 1u);   // the previous item's epilogue has read dQ
        tc_fence_after();
        const uint32_t sa = gb % NA;
        const uint64_t k = desc_mn_base(sKr + sa * C::TILE);
        const uint32_t tdS = tdSb + 64 * (gb & 1u);
        if (elect_one_sync()) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_bf16_ts(tmem, tdS + kk * 8, k + kstep_mn(kk), idesc_g, (j > 0 || kk > 0) ? 1u : 0u);
          umma_commit(a_empty + 8 * sa);  // K_j is free
          umma_commit(ds_free + 8 * (gb & 1u));
          if (j == n_it - 1) umma_commit(acc_done);
        }
        __syncwarp();
      }
    }
  } else if (warp < 4 * NG) {
    // ------------------------------------------------------------------------------------------- gradient warpgroups
    const int g = warp >> 2;            // key columns [CG g, CG g + CG) of every block
    const int row = tid & 127;          // query row = TMEM lane
    const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;
    const uint32_t tS = tSb + CG * g + lane_off, tdP = tdPb + CG * g + lane_off;
    const float c1 = scale * LOG2E_F;
    uint32_t gb = 0;
    int itm = 0;
    for (int w = (int)blockIdx.x; w < n_items; w += G, ++itm) {
      const int ib = n_blk - 1 - w / n_bh, n_it = ib + 1;
      const int h = (w % n_bh) % H, b = (w % n_bh) / H;
      const long long bh = w % n_bh;
      const int qi = ib * 128 + row;
      float lse2 = INFINITY, dl = 0.f;    // out-of-range query row: p = 0
      if (qi < S) {
        lse2 = lse[bh * S + qi] * LOG2E_F;
        dl = delta[bh * S + qi];
      }
      for (int j = 0; j < n_it; ++j, ++gb) {
        mbar_wait(sd_ready, gb & 1u);
        tc_fence_after();
        uint32_t rs[CG], rd[CG];
#pragma unroll
        for (int c = 0; c < CG / 32; ++c) {
          tmem_ld_x32(tS + c * 32, rs + c * 32);
          tmem_ld_x32(tdP + c * 32, rd + c * 32);
        }
        tmem_ld_wait();
        warp_arrive(sd_loaded, lane);      // S / dP may be overwritten by the next block's logits
        const int lim = (j == ib) ? row - CG * g : CG;  // columns c > lim are keys after this query (diagonal block)
        if (gb >= 2) {                     // the dQ product of block gb-2 has consumed this dS slot
          mbar_wait(ds_free + 8 * (gb & 1u), ((gb - 2) >> 1) & 1u);
          tc_fence_after();
        }
        const uint32_t tdS = tdSb + 64 * (gb & 1u) + (CG / 2) * g + lane_off;
#pragma unroll
        for (int c = 0; c < CG / 32; ++c) {
          uint32_t dk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const int cc = c * 32 + i;
            float p0 = ex2f(fmaf(__uint_as_float(rs[cc]), c1, -lse2)), p1 = ex2f(fmaf(__uint_as_float(rs[cc + 1]), c1, -lse2));
            if (cc > lim) p0 = 0.f;
            if (cc + 1 > lim) p1 = 0.f;
            dk[i >> 1] = pack_bf16x2((p0 * scale) * (__uint_as_float(rd[cc]) - dl),
                                     (p1 * scale) * (__uint_as_float(rd[cc + 1]) - dl));
          }
          tmem_st_x16(tdS + c * 16, dk);
        }
        tmem_st_wait();
        warp_arrive(ds_ready + 8 * (gb & 1u), lane);
      }
      // ---- epilogue: each group writes its share of the dQ columns
      mbar_wait(acc_done, (uint32_t)itm & 1u);
      tc_fence_after();
      constexpr int EG = DH / NG, W = EG >= 32 ? 32 : 16;
      bf16* dst = dqkv + ((((long long)b * S + qi) * 3 + 0) * H + h) * DH + g * EG;
#pragma unroll 1
      for (int c = 0; c < EG / W; ++c) {
        uint32_t r[W];
        tmem_ld_n<W>(tmem + lane_off + g * EG + c * W, r);
        tmem_ld_wait();
        if (qi < S) store_cols_bf16<W>(dst + c * W, r, 1.f);
      }
      warp_arrive(acc_free, lane);   // dQ may be overwritten by the next item
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
static int ws_qkv_map(CUtensorMap* tm, const void* qkv, int B, int S, int H, int dh) {
  uint64_t dims[4] = {(uint64_t)dh, (uint64_t)3 * H, (uint64_t)S, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)dh * 2, (uint64_t)3 * H * dh * 2, (uint64_t)S * 3 * H * dh * 2};
  uint32_t box[4] = {64, 1, 128, 1};
  return make_tmap_bf16(tm, qkv, 4, dims, strides, box);
}
static int ws_o_map(CUtensorMap* tm, const void* o, int B, int S, int H, int dh) {
  uint64_t dims[4] = {(uint64_t)dh, (uint64_t)H, (uint64_t)S, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)dh * 2, (uint64_t)H * dh * 2, (uint64_t)S * H * dh * 2};
  uint32_t box[4] = {64, 1, 128, 1};
  return make_tmap_bf16(tm, o, 4, dims, strides, box);
}

// Development A/B switches (make DEV=1 only; the shipped library runs NG = 2, persistent kernels):
//   DB200_ATTN_NG       number of math warpgroups per CTA, 2 or 4
//   DB200_ATTN_PERSIST  bit 0 = forward, bit 1 = backward as persistent kernels (default 3); 0 = one CTA per work item
//   DB200_ATTN_FWD2     1 = the two-tile forward kernel
#ifdef DB200_DEV_KNOBS
static int attn_ng() {
  static const int ng = [] { const char* e = getenv("DB200_ATTN_NG"); return (e && atoi(e) == 4) ? 4 : 2; }();
  return ng;
}
static int attn_persist_bits() {
  static const int v = [] { const char* e = getenv("DB200_ATTN_PERSIST"); return e ? atoi(e) : 3; }();
  return v;
}
static int attn_fwd2_on() {
  static const int v = [] { const char* e = getenv("DB200_ATTN_FWD2"); return e ? atoi(e) : 0; }();
  return v;
}
#else
static int attn_ng() { return 2; }
static int attn_persist_bits() { return 3; }
#endif

template <int DH, int NG>
static int fwd_ws_launch_t(cudaStream_t stream, const void* qkv, void* out, float* lse, int B, int S, int H,
                           float scale) {
  using C = FwdWs<DH, NG>;
  CUtensorMap tm;
  int rc = ws_qkv_map(&tm, qkv, B, S, H, DH);
  if (rc != DB200_OK) return rc;
  static const cudaError_t attr = cudaFuncSetAttribute(attn_fwd_ws_kernel<DH, NG>,
                                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
  DB200_CUDA(attr);
  const int n_items = ((S + 127) / 128) * H * B;
  // persistent: one CTA per SM walks the items in order of decreasing work (DB200_ATTN_PERSIST=0, development A/B
  // switch: one CTA per item as before)
  dim3 grid((attn_persist_bits() & 1) && n_items > sm_count() ? sm_count() : n_items);
#ifdef DB200_DEV_KNOBS
  static const int xflags = [] { const char* e = getenv("DB200_ATTN_EXP"); return e ? atoi(e) : 0; }();
#else
  const int xflags = 0;
#endif
  attn_fwd_ws_kernel<DH, NG><<<grid, C::THREADS, C::SMEM, stream>>>(tm, (bf16*)out, lse, S, H, n_items, scale, xflags);
  return check_launch("attn_fwd_ws_kernel");
}

#ifdef DB200_DEV_KNOBS
template <int DH>
static int fwd2_ws_launch_t(cudaStream_t stream, const void* qkv, void* out, float* lse, int B, int S, int H,
                            float scale) {
  using C = Fwd2Ws<DH>;
  CUtensorMap tm;
  int rc = ws_qkv_map(&tm, qkv, B, S, H, DH);
  if (rc != DB200_OK) return rc;
  static const cudaError_t attr = cudaFuncSetAttribute(attn_fwd2_ws_kernel<DH>,
                                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
  DB200_CUDA(attr);
  const int n_qt = (S + 127) / 128;
  const int n_items = ((n_qt + 1) / 2) * H * B;
  dim3 grid(n_items > sm_count() ? sm_count() : n_items);
  attn_fwd2_ws_kernel<DH><<<grid, C::THREADS, C::SMEM, stream>>>(tm, (bf16*)out, lse, S, H, n_items, scale);
  return check_launch("attn_fwd2_ws_kernel");
}
#endif  // DB200_DEV_KNOBS

int attn_fwd_ws_launch(cudaStream_t stream, const void* qkv, void* out, float* lse, int B, int S, int H, int dh,
                       float scale) {
#ifdef DB200_DEV_KNOBS
  if (attn_fwd2_on()) {
    if (dh == 128) return fwd2_ws_launch_t<128>(stream, qkv, out, lse, B, S, H, scale);
    return fwd2_ws_launch_t<64>(stream, qkv, out, lse, B, S, H, scale);
  }
#endif
  if (attn_ng() == 2) {
    if (dh == 128) return fwd_ws_launch_t<128, 2>(stream, qkv, out, lse, B, S, H, scale);
    return fwd_ws_launch_t<64, 2>(stream, qkv, out, lse, B, S, H, scale);
  }
#ifdef DB200_DEV_KNOBS
  if (dh == 128) return fwd_ws_launch_t<128, 4>(stream, qkv, out, lse, B, S, H, scale);
  return fwd_ws_launch_t<64, 4>(stream, qkv, out, lse, B, S, H, scale);
#else
  return set_error(DB200_E_UNSUPPORTED, "attn_fwd: unreachable");
#endif
}

template <int DH, int NG>
static int bwd_ws_launch_t(cudaStream_t stream, const void* qkv, const void* dout, const float* lse,
                           const float* delta, void* dqkv, int B, int S, int H, float scale) {
  using C = BwdWs<DH, NG>;
  CUtensorMap tq, to;
  int rc = ws_qkv_map(&tq, qkv, B, S, H, DH);
  if (rc == DB200_OK) rc = ws_o_map(&to, dout, B, S, H, DH);
  if (rc != DB200_OK) return rc;
  static const cudaError_t a0 = cudaFuncSetAttribute(attn_bwd_dkdv_ws_kernel<DH, NG>,
                                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
  static const cudaError_t a1 = cudaFuncSetAttribute(attn_bwd_dq_ws_kernel<DH, NG>,
                                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
  DB200_CUDA(a0);
  DB200_CUDA(a1);
  const int n_items = ((S + 127) / 128) * H * B;
  dim3 grid((attn_persist_bits() & 2) && n_items > sm_count() ? sm_count() : n_items);
  attn_bwd_dkdv_ws_kernel<DH, NG><<<grid, C::THREADS, C::SMEM, stream>>>(tq, to, lse, delta, (bf16*)dqkv, S, H, n_items,
                                                                         scale);
  rc = check_launch("attn_bwd_dkdv_ws_kernel");
  if (rc != DB200_OK) return rc;
  attn_bwd_dq_ws_kernel<DH, NG><<<grid, C::THREADS, C::SMEM, stream>>>(tq, to, lse, delta, (bf16*)dqkv, S, H, n_items,
                                                                       scale);
  return check_launch("attn_bwd_dq_ws_kernel");
}

#ifdef DB200_DEV_KNOBS
int attn_trace_set(void* buf, int slots) {
  unsigned long long* p = reinterpret_cast<unsigned long long*>(buf);
  DB200_CUDA(cudaMemcpyToSymbol(g_attn_trace, &p, sizeof(p)));
  DB200_CUDA(cudaMemcpyToSymbol(g_attn_trace_slots, &slots, sizeof(slots)));
  return DB200_OK;
}
#endif

int attn_bwd_ws_launch(cudaStream_t stream, const void* qkv, const void* dout, const float* lse, const float* delta,
                       void* dqkv, int B, int S, int H, int dh, float scale) {
  if (attn_ng() == 2) {
    if (dh == 128) return bwd_ws_launch_t<128, 2>(stream, qkv, dout, lse, delta, dqkv, B, S, H, scale);
    return bwd_ws_launch_t<64, 2>(stream, qkv, dout, lse, delta, dqkv, B, S, H, scale);
  }
#ifdef DB200_DEV_KNOBS
  if (dh == 128) return bwd_ws_launch_t<128, 4>(stream, qkv, dout, lse, delta, dqkv, B, S, H, scale);
  return bwd_ws_launch_t<64, 4>(stream, qkv, dout, lse, delta, dqkv, B, S, H, scale);
#else
  return set_error(DB200_E_UNSUPPORTED, "attn_bwd: unreachable");
#endif
}

}  // namespace db200

#ifdef DB200_DEV_KNOBS
// development library only: device buffer of slots x 4 roles x 160 events (u64) for the attention timeline
extern "C" int db200_dev_attn_trace(void* buf, int slots) { return db200::attn_trace_set(buf, slots); }
#endif
