// 2-CTA variant of the tcgen05 GEMM: a cluster of two CTAs (two SMs of one TPC) computes a 256 x 256 output tile with
// tcgen05.mma.cta_group::2 (UMMA M = 256).  CTA r of the pair stages A rows [128r, 128r+128) and B columns
// [128r, 128r+128) of the tile, so per k-block each SM pulls 16 KiB + 16 KiB instead of 16 KiB + 32 KiB: a third less
// L2->SMEM traffic per FLOP and six pipeline stages instead of four in the same shared memory.
//
// Protocol (per pair; "leader" = CTA rank 0):
//   * both CTAs' producers issue their TMA loads with .cta_group::2 and point them at the LEADER's full barrier
//     (same smem offset, peer bit cleared); the leader's producer arms it with the bytes of BOTH CTAs;
//   * only the leader issues tcgen05.mma.cta_group::2; tcgen05.commit ... .multicast::cluster arrives on the SAME-offset
//     barrier in both CTAs: "smem stage free" for both producers, "accumulator ready" for both epilogues;
//   * each CTA's 8 epilogue warps drain their own 128 TMEM lanes (all 256 columns) through the shared epilogue code;
//     all 16 warps of the pair arrive on the leader's "accumulator free" barrier (remote arrive for the peer);
//   * TMEM (2 x 256 columns per CTA) is allocated / freed with the cta_group::2 forms by the same warp in both CTAs.
#include <cstdlib>

#include "gemm_common.cuh"

namespace db200 {

// Six 32 KiB pipeline stages + one 4 KiB epilogue staging buffer per epilogue warp.
// -DDB200_G2_PF (A/B builds only): five stages + TWO staging buffers, the residual / ReLU-mask operand of the next
// 32 x 64 block arriving by cp.async while the current block is processed.  Measured in-step on B200: the epilogues with
// such an operand gain 7-8 % (0.248 vs 0.270 ms, 0.708 vs 0.761 ms per step), every other GEMM loses 1-5 % to the
// shallower pipeline — a wash on the step (17.49 vs 17.47 ms), so the deeper pipeline stays the default.
#ifdef DB200_G2_PF
constexpr int G2_STAGES = 5;
constexpr int G2_STG_BUFS = 2;
#else
constexpr int G2_STAGES = 6;
constexpr int G2_STG_BUFS = 1;
#endif
constexpr bool G2_PF = G2_STG_BUFS == 2;
constexpr uint32_t G2_A_BYTES = 128 * 64 * 2;  // this CTA's 128 rows of A
constexpr uint32_t G2_B_BYTES = 128 * 64 * 2;  // this CTA's 128 columns of B
constexpr uint32_t G2_STAGE = G2_A_BYTES + G2_B_BYTES;
constexpr size_t G2_SMEM = 1024 + size_t(G2_STAGES) * G2_STAGE + 256 + 8 * G2_STG_BUFS * 4096;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* m, uint32_t leader_bar, int c0,
                                                int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(dst),
      "l"(m), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst, const CUtensorMap* m, uint32_t leader_bar, int c0,
                                                int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
      "%5}], [%2];" ::"r"(dst),
      "l"(m), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_ss_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm_mc(uint32_t bar) {  // arrive on `bar` in BOTH CTAs of the pair
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar),
      "r"(cta)
      : "memory");
}

struct Tile2 {
  int m_blk, n_blk, kb0, kb1;  // m_blk in units of 256 rows
};
__device__ __forceinline__ Tile2 decode_tile2(const GemmParams& p, int m_tiles2, int tile) {
  const int mn_tiles = m_tiles2 * p.n_tiles;
  const int split = tile / mn_tiles;
  const int mn = tile - split * mn_tiles;
  const int group_sz = GROUP_M * p.n_tiles;
  const int group = mn / group_sz;
  const int first_m = group * GROUP_M;
  const int gm = min(GROUP_M, m_tiles2 - first_m);
  const int in_group = mn - group * group_sz;
  Tile2 t;
  t.m_blk = first_m + in_group % gm;
  t.n_blk = in_group / gm;
  const int per = (p.kb_total + p.splits - 1) / p.splits;
  t.kb0 = split * per;
  t.kb1 = min(p.kb_total, t.kb0 + per);
  return t;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p,
                const int m_tiles2) {
  constexpr int BN = 256;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t sA = base;
  const uint32_t sB = base + G2_STAGES * G2_A_BYTES;
  const uint32_t bars = base + G2_STAGES * G2_STAGE;
  const uint32_t full_bar = bars;
  const uint32_t empty_bar = bars + 8 * G2_STAGES;
  const uint32_t tfull_bar = bars + 16 * G2_STAGES;
  const uint32_t tempty_bar = tfull_bar + 16;
  const uint32_t tmem_slot = tempty_bar + 16;
  const uint32_t stg_base = bars + 256;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const uint32_t PEER_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> leader's copy

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < G2_STAGES; ++i) {
      mbar_init(full_bar + 8 * i, 1);   // leader: one arrive.expect_tx per use (peer's copy is unused)
      mbar_init(empty_bar + 8 * i, 1);  // one multicast commit per use
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull_bar + 8 * i, 1);    // one multicast commit per tile
      mbar_init(tempty_bar + 8 * i, 16);  // leader: 8 local + 8 remote epilogue warps (peer's copy is unused)
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();  // barrier inits and the TMEM allocation of BOTH CTAs are visible before anyone signals
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int n_clusters = gridDim.x >> 1;
  const int cid = blockIdx.x >> 1;
  const int total_tiles = m_tiles2 * p.n_tiles * p.splits;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    // The warp stays converged (every lane waits on the barrier); the TMA instructions are issued by the lane
    // elect.sync picks — ptxas then emits them as plain uniform-datapath instructions instead of a lane-by-lane
    // "waterfall" loop around each one (what `if (lane == 0)` around the whole loop compiled to).
    {
      uint32_t stage = 0, phase = 0;
      for (int tile = cid; tile < total_tiles; tile += n_clusters) {
        const Tile2 t = decode_tile2(p, m_tiles2, tile);
        const int m0 = t.m_blk * 256 + rank * 128, n0 = t.n_blk * BN + rank * 128;
        for (int kb = t.kb0; kb < t.kb1; ++kb) {
          mbar_wait(empty_bar + 8 * stage, phase ^ 1);
          const uint32_t fb_leader = (full_bar + 8 * stage) & PEER_MASK;
          const int k0 = kb * BK;
          const uint32_t a_dst = sA + stage * G2_A_BYTES;
          const uint32_t b_dst = sB + stage * G2_B_BYTES;
          if (elect_one_sync()) {
          if (leader) mbar_expect_tx(full_bar + 8 * stage, 2 * G2_STAGE);  // bytes of both CTAs land on this barrier
          if (p.a_3d) {
            tma_load_3d_2sm(a_dst, &tmA, fb_leader, 0, k0, m0 >> 6);
          } else if (p.a_mn) {
#pragma unroll
            for (int s = 0; s < 2; ++s) tma_load_2d_2sm(a_dst + s * SLAB_BYTES, &tmA, fb_leader, m0 + 64 * s, k0);
          } else {
            tma_load_2d_2sm(a_dst, &tmA, fb_leader, k0, m0);
          }
          if (p.b_3d) {
            tma_load_3d_2sm(b_dst, &tmB, fb_leader, 0, k0, n0 >> 6);
          } else if (p.b_mn) {
#pragma unroll
            for (int s = 0; s < 2; ++s) tma_load_2d_2sm(b_dst + s * SLAB_BYTES, &tmB, fb_leader, n0 + 64 * s, k0);
          } else {
            tma_load_2d_2sm(b_dst, &tmB, fb_leader, k0, n0);
          }
          }  // elect
          __syncwarp();
          if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader) {
      const uint32_t idesc = umma_idesc_bf16(256, BN, p.a_mn, p.b_mn);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int tile = cid; tile < total_tiles; tile += n_clusters) {
        const Tile2 t = decode_tile2(p, m_tiles2, tile);
        mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = t.kb0; kb < t.kb1; ++kb) {
          mbar_wait(full_bar + 8 * stage, phase);
          tc_fence_after();
          // descriptors = one base per operand tile + a constant per K-step (16 elements: 32 bytes inside the swizzle
          // row for K-major, 16 k-rows = 2 KiB for MN-major), all in uniform registers; elect.sync picks the issuer
          const uint32_t a_base = sA + stage * G2_A_BYTES;
          const uint32_t b_base = sB + stage * G2_B_BYTES;
          const uint64_t ad0 = p.a_mn ? umma_smem_desc_sw128(a_base, SLAB_BYTES, 1024) : umma_smem_desc_sw128(a_base, 0, 1024);
          const uint64_t bd0 = p.b_mn ? umma_smem_desc_sw128(b_base, SLAB_BYTES, 1024) : umma_smem_desc_sw128(b_base, 0, 1024);
          const uint64_t astep = p.a_mn ? 128u : 2u, bstep = p.b_mn ? 128u : 2u;
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16_ss_2sm(d_tmem, ad0 + k * astep, bd0 + k * bstep, idesc, (kb > t.kb0 || k > 0) ? 1u : 0u);
            umma_commit_2sm_mc(empty_bar + 8 * stage);
            if (kb == t.kb1 - 1) umma_commit_2sm_mc(tfull_bar + 8 * acc);
          }
          __syncwarp();
          if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (both CTAs, own 128 rows)
    const int ew = warp - 4;
    const int wq = ew & 3;
    const int half = ew >> 2;
    constexpr int CH = BN / 64;
    const uint32_t stg = stg_base + ew * (G2_STG_BUFS * STG_BYTES);
    const uint32_t stg2 = stg + STG_BYTES;   // only touched when G2_PF
    // per-element operand with the output's layout: residual (STORE) or saved activation (RELU_BWD)
    const bf16* pf_src = p.mode == DB200_EPI_RELU_BWD ? p.aux : (p.mode == DB200_EPI_STORE ? p.residual : nullptr);
    const long long pf_ld = p.mode == DB200_EPI_RELU_BWD ? p.ldaux : p.ldr;
    uint32_t acc = 0, acc_phase = 0;
    for (int tile = cid; tile < total_tiles; tile += n_clusters) {
      const Tile2 t = decode_tile2(p, m_tiles2, tile);
      const int row0 = t.m_blk * 256 + rank * 128 + wq * 32;
      const int row = row0 + lane;
      const bool row_ok = row < p.M;
      const int cbase = t.n_blk * BN + half * (BN / 2);
      // pull this half-tile's bias into L1 while the accumulator is still being produced: the epilogue's first use of
      // it sat on an L2 round trip per 32-column chunk (ncu source page: top stall of the CE epilogues)
      if (p.bias && lane < BN / 64 && cbase + lane * 32 < p.N)
        asm volatile("prefetch.global.L1 [%0];" ::"l"(p.bias + cbase + lane * 32));
      // same for the per-element operands the epilogue will read (ReLU mask / residual rows): DRAM -> L2 now
      if (row_ok && cbase < p.N) {
        if (p.aux) {
          const bf16* a = p.aux + (long long)row * p.ldaux + cbase;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
          if (BN == 256 && cbase + 64 < p.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + 64));
        }
        if (p.residual) {
          const bf16* a = p.residual + (long long)row * p.ldr + cbase;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
          if (BN == 256 && cbase + 64 < p.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + 64));
        }
      }
      if (G2_PF && pf_src && cbase < p.N) {  // block 0 of this half tile: in flight under the accumulator wait
        stage_prefetch_bf16(stg, pf_src, pf_ld, row0, cbase, p.M, p.N, lane);
        cp_async_commit();
      }
      mbar_wait(tfull_bar + 8 * acc, acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + acc * BN + half * (BN / 2) + (uint32_t(wq * 32) << 16);
      switch (p.mode) {
        case DB200_EPI_STORE:    epi_store<CH, G2_PF>(p, t_addr, row, row_ok, cbase, stg, row0, lane, stg2); break;
        case DB200_EPI_ATOMIC:   epi_atomic<CH>(p, t_addr, row, row_ok, cbase); break;
        case DB200_EPI_RELU_BWD: epi_relu_bwd<CH, G2_PF>(p, t_addr, row, row_ok, cbase, stg, row0, lane, stg2); break;
        case DB200_EPI_CE_STATS: epi_ce_stats<CH>(p, t_addr, row, row_ok, cbase, t.n_blk * 2 + half); break;
        case DB200_EPI_CE_GRAD:  epi_ce_grad<CH>(p, t_addr, row, row_ok, cbase, stg, row0, lane); break;
        default: break;  // mode -1 (DB200_GEMM_NOEPI=1, timing experiments only): drain nothing
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(tempty_bar + 8 * acc);
        else        mbar_arrive_remote(tempty_bar + 8 * acc, 0);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  cluster_sync_all();  // nobody exits (or frees TMEM) while the pair may still touch its smem / barriers / TMEM
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

int launch_gemm_2cta(cudaStream_t stream, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p) {
  static const cudaError_t attr_rc = cudaFuncSetAttribute(gemm_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G2_SMEM);  // once, thread-safe (magic static)
  DB200_CUDA(attr_rc);
  const int m_tiles2 = (p.M + 255) / 256;
  const int total = m_tiles2 * p.n_tiles * p.splits;
  int clusters = sm_count() / 2;
  if (total < clusters) clusters = total;
  gemm_tc2_kernel<<<clusters * 2, GEMM_THREADS, G2_SMEM, stream>>>(tmA, tmB, p, m_tiles2);
  return check_launch("gemm_tc2_kernel");
}

}  // namespace db200
