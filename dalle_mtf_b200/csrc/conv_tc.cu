// K11 (tensor-core path) — NHWC bf16 convolution forward as an im2col-free implicit GEMM on tcgen05.
//
//   y[n,oy,ox,co] = epi( sum_{tap,ci} x[n, src(oy,tap), src(ox,tap), ci] * w[tap][ci][co] )
//
// The GEMM view: M = output pixels (tile = TN x TH x TW = 128 pixels), N = Cout, K = taps x Cin.  One k-block is
// (one filter tap, 64 input channels): its A operand is ONE 4-D TMA box {64 ch, TW, TH, TN} of the activation tensor
// shifted by the tap offset — out-of-image coordinates are zero-filled by TMA, which IS the SAME padding — landing in
// shared memory as a [128 pixels][64 ch] 128B-swizzled K-major tile; its B operand is a [64 ci][BN co] slab of the
// HWIO kernel (MN-major).  Nothing resembling an im2col buffer exists in HBM or smem.
// Stride-2 convolutions use four parity views of the input (row/col parity), so that every tap is again a dense box.
// Same warp-specialised pipeline as gemm.cu (TMA producer / MMA issuer / TMEM double buffer / 8 epilogue warps).
//
// Reference: tf.layers.conv2d call sites src/vae_tf/models.py:95-109 (encoder; the DALL-E tokenizer path
// src/model_fns.py:72-77 runs exactly these).
#include "common.cuh"
#include "ptx.cuh"

namespace db200 {

constexpr int CT_THREADS = 384;
constexpr uint32_t CT_A_BYTES = 128 * 64 * 2;
constexpr uint32_t CT_SLAB = 64 * 128;

template <int BN>
struct ConvTcCfg {
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr uint32_t B_BYTES = BN * 64 * 2;
  static constexpr uint32_t STAGE_BYTES = CT_A_BYTES + B_BYTES;
  static constexpr uint32_t TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
  static constexpr size_t SMEM_BYTES = 1024 + size_t(STAGES) * STAGE_BYTES + 256;
};

struct ConvTcTap {
  int dy, dx, map;  // coordinate offsets inside parity view `map`
};

struct ConvTcParams {
  int NB, Ho, Wo, Cin, Cout;
  int TW, TH, TN;
  int tiles_w, tiles_h, tiles_n, tiles_c;
  int ntaps, kchunks;
  ConvTcTap taps[16];
  int relu;
  const float* bias;
  const bf16* residual;
  bf16* y;
};

struct ConvTile {
  int n0, oy0, ox0, c_blk;
};
__device__ __forceinline__ ConvTile conv_decode(const ConvTcParams& p, int tile) {
  ConvTile t;
  t.c_blk = tile % p.tiles_c; tile /= p.tiles_c;   // channel tiles innermost: neighbours share the A box (L2)
  t.ox0 = (tile % p.tiles_w) * p.TW; tile /= p.tiles_w;
  t.oy0 = (tile % p.tiles_h) * p.TH; tile /= p.tiles_h;
  t.n0 = tile * p.TN;
  return t;
}

template <int BN>
__global__ void __launch_bounds__(CT_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmA3,
               const __grid_constant__ CUtensorMap tmB, const ConvTcParams p) {
  using Cfg = ConvTcCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t sA = base;
  const uint32_t sB = base + STAGES * CT_A_BYTES;
  const uint32_t bars = base + STAGES * Cfg::STAGE_BYTES;
  const uint32_t full_bar = bars, empty_bar = bars + 8 * STAGES, tfull_bar = bars + 16 * STAGES,
                 tempty_bar = tfull_bar + 16, tmem_slot = tempty_bar + 16;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(full_bar + 8 * i, 1);
      mbar_init(empty_bar + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull_bar + 8 * i, 1);
      mbar_init(tempty_bar + 8 * i, 8);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int total_tiles = p.tiles_n * p.tiles_h * p.tiles_w * p.tiles_c;
  const int nkb = p.ntaps * p.kchunks;

  if (warp == 0) {
    if (lane == 0) {
      const CUtensorMap* maps[4] = {&tmA0, &tmA1, &tmA2, &tmA3};
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const ConvTile t = conv_decode(p, tile);
        for (int tap = 0; tap < p.ntaps; ++tap) {
          const ConvTcTap tp = p.taps[tap];
          for (int kc = 0; kc < p.kchunks; ++kc) {
            mbar_wait(empty_bar + 8 * stage, phase ^ 1);
            const uint32_t fb = full_bar + 8 * stage;
            mbar_expect_tx(fb, Cfg::STAGE_BYTES);
            tma_load_4d(sA + stage * CT_A_BYTES, maps[tp.map], fb, kc * 64, t.ox0 + tp.dx, t.oy0 + tp.dy, t.n0);
            const uint32_t b_dst = sB + stage * Cfg::B_BYTES;
            const int krow = (tap * p.kchunks + kc) * 64;
#pragma unroll
            for (int s = 0; s < BN / 64; ++s) tma_load_2d(b_dst + s * CT_SLAB, &tmB, fb, t.c_blk * BN + 64 * s, krow);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, BN, 0, 1);  // A K-major (K = channels), B MN-major
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(full_bar + 8 * stage, phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_base = sA + stage * CT_A_BYTES, b_base = sB + stage * Cfg::B_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ss(d_tmem, umma_smem_desc_sw128(a_base + k * 32, 0, 1024),
                         umma_smem_desc_sw128(b_base + k * 2048, CT_SLAB, 1024), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(empty_bar + 8 * stage);
          if (kb == nkb - 1) umma_commit(tfull_bar + 8 * acc);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    const int ew = warp - 4, wq = ew & 3, half = ew >> 2;
    constexpr int HALF = (BN >= 64) ? BN / 2 : BN;
    uint32_t acc = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const ConvTile t = conv_decode(p, tile);
      const int r = wq * 32 + lane;  // pixel index inside the tile: ((tn*TH + th)*TW + tw)
      const int tw = r % p.TW, th = (r / p.TW) % p.TH, tn = r / (p.TW * p.TH);
      const int n = t.n0 + tn, oy = t.oy0 + th, ox = t.ox0 + tw;
      const bool ok = n < p.NB && oy < p.Ho && ox < p.Wo;
      const long long o = (((long long)n * p.Ho + oy) * p.Wo + ox) * p.Cout;
      mbar_wait(tfull_bar + 8 * acc, acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + acc * BN + half * HALF + (uint32_t(wq * 32) << 16);
      const int cbase = t.c_blk * BN + half * HALF;
#pragma unroll 1
      for (int c = 0; c < HALF / 32; ++c) {
        const int col0 = cbase + c * 32;
        if (col0 >= p.Cout) break;
        uint32_t rr[32];
        tmem_ld_x32(t_addr + c * 32, rr);
        tmem_ld_wait();
        if (!ok) continue;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int col = col0 + g * 8;
          if (col + 8 > p.Cout) break;
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(rr[g * 8 + j]);
          if (p.bias) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col)),
                         b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col) + 1);
            v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
            v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if (p.residual) {
            const uint4 q = *reinterpret_cast<const uint4*>(p.residual + o + col);
            const float2 r0 = unpack_bf16x2(q.x), r1 = unpack_bf16x2(q.y), r2 = unpack_bf16x2(q.z),
                         r3 = unpack_bf16x2(q.w);
            v[0] += r0.x; v[1] += r0.y; v[2] += r1.x; v[3] += r1.y;
            v[4] += r2.x; v[5] += r2.y; v[6] += r3.x; v[7] += r3.y;
          }
          uint4 q;
          q.x = pack_bf16x2(v[0], v[1]); q.y = pack_bf16x2(v[2], v[3]);
          q.z = pack_bf16x2(v[4], v[5]); q.w = pack_bf16x2(v[6], v[7]);
          *reinterpret_cast<uint4*>(p.y + o + col) = q;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar + 8 * acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN>
static int conv_tc_launch(cudaStream_t stream, const CUtensorMap* tmA, const CUtensorMap& tmB, const ConvTcParams& p) {
  using Cfg = ConvTcCfg<BN>;
  static bool attr = false;
  if (!attr) {
    DB200_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)Cfg::SMEM_BYTES));
    attr = true;
  }
  const int total = p.tiles_n * p.tiles_h * p.tiles_w * p.tiles_c;
  const int grid = total < sm_count() ? total : sm_count();
  conv_tc_kernel<BN><<<grid, CT_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA[0], tmA[1], tmA[2], tmA[3], tmB, p);
  return check_launch("conv_tc_kernel");
}

static int pow2_floor(int x) {
  int p = 1;
  while (p * 2 <= x) p *= 2;
  return p;
}

}  // namespace db200

using namespace db200;

// y (bf16 NHWC) = [relu](conv(x bf16 NHWC, w bf16 HWIO) + bias f32) [+ residual bf16].  SAME padding.
// Supported: KH x KW <= 16 taps, stride 1, or stride 2 with even H, W;  Cin % 64 == 0;  Cout % 8 == 0.
extern "C" int db200_conv2d_fwd_tc(db200_stream_t stream_, const db200_conv_desc* c, const void* x_bf16,
                                   const void* w_bf16, const float* bias, const void* residual_bf16, void* y_bf16) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DB200_REQUIRE(c && x_bf16 && w_bf16 && y_bf16, DB200_E_INVALID, "conv2d_fwd_tc: NULL argument");
  DB200_REQUIRE(!c->transposed && !c->act_f32, DB200_E_UNSUPPORTED,
                "conv2d_fwd_tc: bf16 forward convolution only (no transposed / fp32 activations)");
  DB200_REQUIRE(c->Cin % 64 == 0 && c->Cout % 8 == 0, DB200_E_UNSUPPORTED,
                "conv2d_fwd_tc: needs Cin %% 64 == 0 and Cout %% 8 == 0 (got %d, %d)", c->Cin, c->Cout);
  DB200_REQUIRE(c->KH * c->KW <= 16 && (c->stride == 1 || c->stride == 2), DB200_E_UNSUPPORTED,
                "conv2d_fwd_tc: unsupported kernel/stride");
  DB200_REQUIRE(c->Ho == (c->H + c->stride - 1) / c->stride && c->Wo == (c->W + c->stride - 1) / c->stride,
                DB200_E_INVALID, "conv2d_fwd_tc: Ho/Wo do not match SAME padding");
  DB200_REQUIRE(c->stride == 1 || (c->H % 2 == 0 && c->W % 2 == 0), DB200_E_UNSUPPORTED,
                "conv2d_fwd_tc: stride 2 needs even H and W");
  DB200_REQUIRE(aligned16(bias) && aligned16(residual_bf16) && aligned16(y_bf16), DB200_E_ALIGN,
                "conv2d_fwd_tc: unaligned pointer");
  ConvTcParams p{};
  p.NB = c->N; p.Ho = c->Ho; p.Wo = c->Wo; p.Cin = c->Cin; p.Cout = c->Cout;
  p.TW = pow2_floor(c->Wo < 16 ? c->Wo : 16);
  int th = 128 / p.TW;
  p.TH = pow2_floor(c->Ho < th ? c->Ho : th);
  p.TN = 128 / (p.TW * p.TH);
  p.tiles_w = (c->Wo + p.TW - 1) / p.TW;
  p.tiles_h = (c->Ho + p.TH - 1) / p.TH;
  p.tiles_n = (c->N + p.TN - 1) / p.TN;
  p.ntaps = c->KH * c->KW;
  p.kchunks = c->Cin / 64;
  p.relu = c->relu; p.bias = bias;
  p.residual = reinterpret_cast<const bf16*>(residual_bf16);
  p.y = reinterpret_cast<bf16*>(y_bf16);
  const int s = c->stride;
  int total_h = (c->Ho - 1) * s + c->KH - c->H; if (total_h < 0) total_h = 0;
  int total_w = (c->Wo - 1) * s + c->KW - c->W; if (total_w < 0) total_w = 0;
  const int pt = total_h / 2, pl = total_w / 2;

  CUtensorMap tmA[4];
  const uint32_t box[4] = {64, (uint32_t)p.TW, (uint32_t)p.TH, (uint32_t)p.TN};
  const bf16* xb = reinterpret_cast<const bf16*>(x_bf16);
  int rc;
  if (s == 1) {
    uint64_t dims[4] = {(uint64_t)c->Cin, (uint64_t)c->W, (uint64_t)c->H, (uint64_t)c->N};
    uint64_t str[3] = {(uint64_t)c->Cin * 2, (uint64_t)c->W * c->Cin * 2, (uint64_t)c->H * c->W * c->Cin * 2};
    rc = make_tmap_bf16(&tmA[0], xb, 4, dims, str, box);
    if (rc != DB200_OK) return rc;
    tmA[1] = tmA[2] = tmA[3] = tmA[0];
    for (int kh = 0; kh < c->KH; ++kh)
      for (int kw = 0; kw < c->KW; ++kw) p.taps[kh * c->KW + kw] = ConvTcTap{kh - pt, kw - pl, 0};
  } else {
    // parity views: view(ph,pw)[n][h2][w2][c] = x[n][2*h2+ph][2*w2+pw][c]
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw) {
        uint64_t dims[4] = {(uint64_t)c->Cin, (uint64_t)c->W / 2, (uint64_t)c->H / 2, (uint64_t)c->N};
        uint64_t str[3] = {(uint64_t)2 * c->Cin * 2, (uint64_t)2 * c->W * c->Cin * 2,
                           (uint64_t)c->H * c->W * c->Cin * 2};
        rc = make_tmap_bf16(&tmA[ph * 2 + pw], xb + ((long long)ph * c->W + pw) * c->Cin, 4, dims, str, box);
        if (rc != DB200_OK) return rc;
      }
    for (int kh = 0; kh < c->KH; ++kh)
      for (int kw = 0; kw < c->KW; ++kw) {
        // input row = 2*oy + (kh - pt) = 2*(oy + dh) + ph
        const int eh = kh - pt, ew = kw - pl;
        const int ph = ((eh % 2) + 2) % 2, pw = ((ew % 2) + 2) % 2;
        p.taps[kh * c->KW + kw] = ConvTcTap{(eh - ph) / 2, (ew - pw) / 2, ph * 2 + pw};
      }
  }
  int bn = c->Cout >= 256 ? 256 : (c->Cout >= 128 ? 128 : 64);
  p.tiles_c = (c->Cout + bn - 1) / bn;
  CUtensorMap tmB;
  rc = make_tmap_2d(&tmB, w_bf16, (uint64_t)c->Cout, (uint64_t)p.ntaps * c->Cin, (uint64_t)c->Cout, 64, 64);
  if (rc != DB200_OK) return rc;
  if (bn == 256) return conv_tc_launch<256>(stream, tmA, tmB, p);
  if (bn == 128) return conv_tc_launch<128>(stream, tmA, tmB, p);
  return conv_tc_launch<64>(stream, tmA, tmB, p);
}
