// K3/K5/K6/K7/K8 — persistent, warp-specialised bf16 GEMM on tcgen05 for sm_100a.
//
//   D[M,N] = A[M,K] * B[K,N],  fp32 accumulation in TMEM, operands staged in shared memory by TMA (128B swizzle).
//
//   warp 0   : TMA producer (one elected lane)            smem ring: STAGES x {A 128x64, B BNx64} bf16
//   warp 1   : MMA issuer  (one elected lane, tcgen05.mma cta_group::1, UMMA 128 x BN x 16)
//   warp 2   : TMEM allocator / deallocator (2 accumulator stages of BN columns -> epilogue overlaps next tile)
//   warp 3   : idle
//   warps 4-11: epilogue (tcgen05.ld; each accumulator row is split between two threads) -> bias / ReLU /
//               residual / split-K reduction / cross-entropy statistics and gradient
//
// Either operand may be K-major or MN-major in global memory; the tensor maps and the UMMA descriptors absorb the
// difference, so forward (x*W), dgrad (dy*W^T) and wgrad (x^T*dy) are the same kernel.
//
// Reference call sites replaced: every mtf einsum / mtf.layers.dense of src/dalle_mtf/models.py (235-244, 303-311,
// 317-324, 361-371, 391-395) and their gradients produced by mtf.gradients (src/optimizers.py:34).
#include "common.cuh"
#include "ptx.cuh"

namespace db200 {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_THREADS = 384;  // 4 control warps + 8 epilogue warps
constexpr int GROUP_M = 8;
constexpr uint32_t SLAB_BYTES = BK * 128;  // one MN-major slab: [BK k-rows][64 bf16] = 8 KiB
constexpr uint32_t A_BYTES = BM * BK * 2;  // 16 KiB

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr uint32_t B_BYTES = BN * BK * 2;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr uint32_t TMEM_COLS = 2 * BN;  // 512 or 256: power of two >= 32
  static constexpr size_t SMEM_BYTES =
      1024 /*align slack*/ + size_t(STAGES) * STAGE_BYTES + 256 /*barriers*/ + 8 * 4096 /*epilogue staging*/;
};

struct GemmParams {
  int M, N, K;
  int a_mn, b_mn;
  int m_tiles, n_tiles, splits, kb_total;
  int mode, out_f32, relu;
  float alpha;
  void* D;
  long long ldd;
  const float* bias;
  const bf16* residual;
  long long ldr;
  const bf16* aux;
  long long ldaux;
  const int* labels;
  float* part_max;
  float* part_sum;
  float* label_logit;
  const float* lse;
  int n_valid;
  int n_parts;  // CE_STATS: partials per row = 2 * n_tiles (one per half tile)
  float* colsum;  // CE_GRAD / RELU_BWD: colsum[n] += sum_m D[m,n] (bias gradient), or NULL
};

struct TileCoord {
  int m_blk, n_blk, kb0, kb1;
};

__device__ __forceinline__ TileCoord decode_tile(const GemmParams& p, int tile) {
  const int mn_tiles = p.m_tiles * p.n_tiles;
  const int split = tile / mn_tiles;
  const int mn = tile - split * mn_tiles;
  const int group_sz = GROUP_M * p.n_tiles;
  const int group = mn / group_sz;
  const int first_m = group * GROUP_M;
  const int gm = min(GROUP_M, p.m_tiles - first_m);
  const int in_group = mn - group * group_sz;
  TileCoord t;
  t.m_blk = first_m + in_group % gm;
  t.n_blk = in_group / gm;
  const int per = (p.kb_total + p.splits - 1) / p.splits;
  t.kb0 = split * per;
  t.kb1 = min(p.kb_total, t.kb0 + per);
  return t;
}


// ------------------------------------------------------------------------------------------------------------------
// epilogues.  One thread = one accumulator row x (BN/2) columns, processed in 32-column chunks straight out of TMEM.
// All tcgen05.ld are executed by the whole warp (they are .sync.aligned); predicates only guard the global accesses.
// ------------------------------------------------------------------------------------------------------------------
constexpr float kLog2e = 1.4426950408889634f;
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void load8(const float* p, float* o) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}
__device__ __forceinline__ void store8_bf16(bf16* dp, const float* o) {
  uint4 q;
  q.x = pack_bf16x2(o[0], o[1]); q.y = pack_bf16x2(o[2], o[3]);
  q.z = pack_bf16x2(o[4], o[5]); q.w = pack_bf16x2(o[6], o[7]);
  *reinterpret_cast<uint4*>(dp) = q;
}
__device__ __forceinline__ void load8_bf16(const bf16* sp, float* o) {
  const uint4 rr = *reinterpret_cast<const uint4*>(sp);
  const float2 r0 = unpack_bf16x2(rr.x), r1 = unpack_bf16x2(rr.y), r2 = unpack_bf16x2(rr.z), r3 = unpack_bf16x2(rr.w);
  o[0] = r0.x; o[1] = r0.y; o[2] = r1.x; o[3] = r1.y; o[4] = r2.x; o[5] = r2.y; o[6] = r3.x; o[7] = r3.y;
}

// Coalesced epilogue I/O.  A thread owns one accumulator ROW, so touching global memory directly would make every
// warp-level access hit 32 different rows with 16 bytes each (partial sectors -> L2 read-modify-write: ncu showed
// 1.5-2 GB of DRAM reads for a 4.1 GB write, even with 64-byte segments).  Each epilogue warp therefore owns a 4 KiB
// smem buffer holding a [32 rows][128 B] block (16-byte slots XOR-swizzled by row: conflict-free both ways); global
// traffic is 8 instructions of 4 rows x 128 contiguous bytes = whole cache lines, for outputs AND for the
// residual / ReLU-mask operands.
constexpr uint32_t STG_BYTES = 32 * 128;  // per epilogue warp

__device__ __forceinline__ uint32_t stg_addr(uint32_t stg, int row, int slot) {
  return stg + row * 128 + (((slot ^ row) & 7) << 4);
}
__device__ __forceinline__ void st_shared_v4u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4u(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr)
               : "memory");
  return v;
}
// this lane's row: 32 bf16 values -> slots [4h, 4h+4)
__device__ __forceinline__ void stage_put_bf16(uint32_t stg, int lane, int h, const float* o) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const float* v = o + g * 8;
    st_shared_v4u(stg_addr(stg, lane, h * 4 + g), pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                  pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
  }
}
__device__ __forceinline__ void stage_get_bf16(uint32_t stg, int lane, int h, float* o) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const uint4 q = ld_shared_v4u(stg_addr(stg, lane, h * 4 + g));
    const float2 a = unpack_bf16x2(q.x), b = unpack_bf16x2(q.y), c = unpack_bf16x2(q.z), d = unpack_bf16x2(q.w);
    float* v = o + g * 8;
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
  }
}
// this lane's row: 32 fp32 values -> all 8 slots
__device__ __forceinline__ void stage_put_f32(uint32_t stg, int lane, const float* o) {
#pragma unroll
  for (int g = 0; g < 8; ++g)
    st_shared_v4u(stg_addr(stg, lane, g), __float_as_uint(o[g * 4]), __float_as_uint(o[g * 4 + 1]),
                  __float_as_uint(o[g * 4 + 2]), __float_as_uint(o[g * 4 + 3]));
}
// smem block -> global rows [row0, row0+32) x 128 bytes starting at element column col0 (ESZ bytes per element)
template <int ESZ>
__device__ __forceinline__ void stage_flush(uint32_t stg, void* D, long long ldd, int row0, int col0, int M, int N,
                                            int lane) {
  constexpr int EPS = 16 / ESZ;  // elements per 16-byte slot
  __syncwarp();
  const int slot = lane & 7, rsub = lane >> 3;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = i * 4 + rsub;
    const uint4 v = ld_shared_v4u(stg_addr(stg, r, slot));
    const int row = row0 + r, col = col0 + slot * EPS;
    if (row < M && col + EPS <= N)
      *reinterpret_cast<uint4*>(reinterpret_cast<char*>(D) + ((long long)row * ldd + col) * ESZ) = v;
  }
  __syncwarp();
}
// global bf16 rows -> smem block (zero where out of range)
__device__ __forceinline__ void stage_fetch_bf16(uint32_t stg, const bf16* src, long long ld, int row0, int col0, int M,
                                                 int N, int lane) {
  const int slot = lane & 7, rsub = lane >> 3;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = i * 4 + rsub;
    const int row = row0 + r, col = col0 + slot * 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (row < M && col + 8 <= N) v = __ldg(reinterpret_cast<const uint4*>(src + (long long)row * ld + col));
    st_shared_v4u(stg_addr(stg, r, slot), v.x, v.y, v.z, v.w);
  }
  __syncwarp();
}

// column sums of the staged [32 rows][64 bf16] block (the values exactly as they are written to D), accumulated into
// colsum[col0 .. col0+64): lane l owns columns 2l, 2l+1.  Rows >= M hold zeros (their `o` was zeroed).
__device__ __forceinline__ void stage_colsum_bf16(uint32_t stg, float* colsum, int col0, int N, int lane) {
  __syncwarp();
  float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
  for (int r = 0; r < 32; ++r) {
    uint32_t w;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"(stg_addr(stg, r, lane >> 2) + (lane & 3) * 4) : "memory");
    const float2 f = unpack_bf16x2(w);
    s0 += f.x;
    s1 += f.y;
  }
  const int col = col0 + 2 * lane;
  if (col < N) atomicAdd(colsum + col, s0);
  if (col + 1 < N) atomicAdd(colsum + col + 1, s1);
}

template <int CH>
__device__ __forceinline__ void epi_store(const GemmParams& p, uint32_t t_addr, int row, bool row_ok, int cbase,
                                          uint32_t stg, int row0, int lane) {
  static_assert(CH % 2 == 0, "epilogue works on pairs of 32-column chunks");
#pragma unroll 1
  for (int pc = 0; pc < CH / 2; ++pc) {
    const int colp = cbase + pc * 64;
    if (colp >= p.N) break;  // warp-uniform
    if (p.residual) stage_fetch_bf16(stg, p.residual, p.ldr, row0, colp, p.M, p.N, lane);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int col0 = colp + h * 32;
      uint32_t r[32];
      tmem_ld_x32(t_addr + (pc * 2 + h) * 32, r);
      tmem_ld_wait();
      float o[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(r[j]) * p.alpha;
      if (p.bias && col0 + 32 <= p.N) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float b[8];
          load8(p.bias + col0 + g * 8, b);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[g * 8 + j] += b[j];
        }
      } else if (p.bias) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col0 + j < p.N) o[j] += __ldg(p.bias + col0 + j);
      }
      if (p.relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] = fmaxf(o[j], 0.f);
      }
      if (p.residual) {
        float rr[32];
        stage_get_bf16(stg, lane, h, rr);
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] += rr[j];
      }
      if (p.out_f32) {  // 32 fp32 columns are a whole 128-byte line: flush per chunk
        if (p.residual) __syncwarp();
        // (fp32 output with a residual shares the buffer: the residual of chunk h was consumed above)
        if (p.residual && h == 0) {  // keep chunk 1's residual: spill it to registers before overwriting
          float keep[32];
          stage_get_bf16(stg, lane, 1, keep);
          stage_put_f32(stg, lane, o);
          stage_flush<4>(stg, p.D, p.ldd, row0, col0, p.M, p.N, lane);
          stage_put_bf16(stg, lane, 1, keep);
          __syncwarp();
        } else {
          stage_put_f32(stg, lane, o);
          stage_flush<4>(stg, p.D, p.ldd, row0, col0, p.M, p.N, lane);
        }
      } else {
        stage_put_bf16(stg, lane, h, o);
      }
    }
    if (!p.out_f32) stage_flush<2>(stg, p.D, p.ldd, row0, colp, p.M, p.N, lane);
  }
  (void)row; (void)row_ok;
}

template <int CH>
__device__ __forceinline__ void epi_atomic(const GemmParams& p, uint32_t t_addr, int row, bool row_ok, int cbase) {
#pragma unroll 1
  for (int c = 0; c < CH; ++c) {
    const int col0 = cbase + c * 32;
    if (col0 >= p.N) break;
    uint32_t r[32];
    tmem_ld_x32(t_addr + c * 32, r);
    tmem_ld_wait();
    if (!row_ok) continue;
    float* dp = reinterpret_cast<float*>(p.D) + (long long)row * p.ldd + col0;
    if (col0 + 32 <= p.N && (p.ldd & 3) == 0) {
#pragma unroll
      for (int j = 0; j < 32; j += 4)  // REDG.E.ADD.F32x4: one vector reduction per 16 bytes
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dp + j),
                     "f"(__uint_as_float(r[j]) * p.alpha), "f"(__uint_as_float(r[j + 1]) * p.alpha),
                     "f"(__uint_as_float(r[j + 2]) * p.alpha), "f"(__uint_as_float(r[j + 3]) * p.alpha)
                     : "memory");
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < p.N) atomicAdd(dp + j, __uint_as_float(r[j]) * p.alpha);
    }
  }
}

template <int CH>
__device__ __forceinline__ void epi_relu_bwd(const GemmParams& p, uint32_t t_addr, int row, bool row_ok, int cbase,
                                             uint32_t stg, int row0, int lane) {
#pragma unroll 1
  for (int pc = 0; pc < CH / 2; ++pc) {
    const int colp = cbase + pc * 64;
    if (colp >= p.N) break;
    stage_fetch_bf16(stg, p.aux, p.ldaux, row0, colp, p.M, p.N, lane);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint32_t r[32];
      tmem_ld_x32(t_addr + (pc * 2 + h) * 32, r);
      tmem_ld_wait();
      float am[32], o[32];
      stage_get_bf16(stg, lane, h, am);
#pragma unroll
      for (int j = 0; j < 32; ++j) o[j] = am[j] > 0.f ? __uint_as_float(r[j]) * p.alpha : 0.f;
      stage_put_bf16(stg, lane, h, o);  // in place: this lane overwrites the slots it just read
    }
    if (p.colsum) stage_colsum_bf16(stg, p.colsum, colp, p.N, lane);
    stage_flush<2>(stg, p.D, p.ldd, row0, colp, p.M, p.N, lane);
  }
  (void)row; (void)row_ok;
}

// per (row, half-tile): running max and sum exp of (acc + bias) over the valid vocabulary columns; label logit
template <int CH>
__device__ __forceinline__ void epi_ce_stats(const GemmParams& p, uint32_t t_addr, int row, bool row_ok, int cbase,
                                             int part_idx) {
  float run_max = -INFINITY, run_sum = 0.f;
  const int label = row_ok ? p.labels[row] : -1;
#pragma unroll 1
  for (int c = 0; c < CH; ++c) {
    const int col0 = cbase + c * 32;
    if (col0 >= p.N) break;
    uint32_t r[32];
    tmem_ld_x32(t_addr + c * 32, r);
    tmem_ld_wait();
    if (!row_ok || col0 >= p.n_valid) continue;
    float v[32];
    if (col0 + 32 <= p.n_valid) {  // interior chunk (warp-uniform): no per-element masking
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float b[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (p.bias) load8(p.bias + col0 + g * 8, b);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[g * 8 + j] = __uint_as_float(r[g * 8 + j]) + b[j];
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int col = col0 + j;
        v[j] = col < p.n_valid ? __uint_as_float(r[j]) + (p.bias ? __ldg(p.bias + col) : 0.f) : -INFINITY;
      }
    }
    if (label >= col0 && label < col0 + 32) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j == label) p.label_logit[row] = v[j];
    }
    float cmax = v[0];
#pragma unroll
    for (int j = 1; j < 32; ++j) cmax = fmaxf(cmax, v[j]);
    const float new_max = fmaxf(run_max, cmax);  // finite: at least one valid column in this chunk
    const float m2 = new_max * kLog2e;
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      s0 += fast_exp2(fmaf(v[j], kLog2e, -m2));
      s1 += fast_exp2(fmaf(v[j + 1], kLog2e, -m2));
    }
    run_sum = run_sum * fast_exp2((run_max - new_max) * kLog2e) + (s0 + s1);
    run_max = new_max;
  }
  if (row_ok) {
    p.part_max[(long long)row * p.n_parts + part_idx] = run_max;
    p.part_sum[(long long)row * p.n_parts + part_idx] = run_sum;
  }
}

// dlogits = alpha * (softmax - onehot), zero in the padded vocabulary columns
template <int CH>
__device__ __forceinline__ void epi_ce_grad(const GemmParams& p, uint32_t t_addr, int row, bool row_ok, int cbase,
                                            uint32_t stg, int row0, int lane) {
  const int label = row_ok ? p.labels[row] : -1;
  const float l2 = row_ok ? p.lse[row] * kLog2e : 0.f;
#pragma unroll 1
  for (int pc = 0; pc < CH / 2; ++pc) {
    const int colp = cbase + pc * 64;
    if (colp >= p.N) break;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int col0 = colp + h * 32;
      uint32_t r[32];
      tmem_ld_x32(t_addr + (pc * 2 + h) * 32, r);
      tmem_ld_wait();
      float o[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) o[j] = 0.f;
      if (row_ok && col0 < p.n_valid) {
        if (col0 + 32 <= p.n_valid) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float b[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            if (p.bias) load8(p.bias + col0 + g * 8, b);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              o[g * 8 + j] = fast_exp2(fmaf(__uint_as_float(r[g * 8 + j]) + b[j], kLog2e, -l2)) * p.alpha;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int cc = col0 + j;
            if (cc < p.n_valid)
              o[j] = fast_exp2(fmaf(__uint_as_float(r[j]) + (p.bias ? __ldg(p.bias + cc) : 0.f), kLog2e, -l2)) *
                     p.alpha;
          }
        }
        if (label >= col0 && label < col0 + 32) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j == label) o[j] -= p.alpha;
        }
      }
      stage_put_bf16(stg, lane, h, o);
    }
    if (p.colsum) stage_colsum_bf16(stg, p.colsum, colp, p.N, lane);
    stage_flush<2>(stg, p.D, p.ldd, row0, colp, p.M, p.N, lane);
  }
}

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
  const uint32_t sA = base;
  const uint32_t sB = base + STAGES * A_BYTES;
  const uint32_t bars = base + STAGES * Cfg::STAGE_BYTES;
  const uint32_t full_bar = bars;                   // STAGES x 8 B
  const uint32_t empty_bar = bars + 8 * STAGES;     // STAGES x 8 B
  const uint32_t tfull_bar = bars + 16 * STAGES;    // 2 x 8 B
  const uint32_t tempty_bar = tfull_bar + 16;       // 2 x 8 B
  const uint32_t tmem_slot = tempty_bar + 16;       // 4 B
  const uint32_t stg_base = bars + 256;             // 8 x STG_BYTES, 16-byte aligned
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(full_bar + 8 * i, 1);
      mbar_init(empty_bar + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull_bar + 8 * i, 1);
      mbar_init(tempty_bar + 8 * i, 8);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int total_tiles = p.m_tiles * p.n_tiles * p.splits;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(p, tile);
        const int m0 = t.m_blk * BM, n0 = t.n_blk * BN;
        for (int kb = t.kb0; kb < t.kb1; ++kb) {
          mbar_wait(empty_bar + 8 * stage, phase ^ 1);
          const uint32_t fb = full_bar + 8 * stage;
          mbar_expect_tx(fb, Cfg::STAGE_BYTES);
          const int k0 = kb * BK;
          const uint32_t a_dst = sA + stage * A_BYTES;
          const uint32_t b_dst = sB + stage * Cfg::B_BYTES;
          if (p.a_mn) {
#pragma unroll
            for (int s = 0; s < BM / 64; ++s) tma_load_2d(a_dst + s * SLAB_BYTES, &tmA, fb, m0 + 64 * s, k0);
          } else {
            tma_load_2d(a_dst, &tmA, fb, k0, m0);
          }
          if (p.b_mn) {
#pragma unroll
            for (int s = 0; s < BN / 64; ++s) tma_load_2d(b_dst + s * SLAB_BYTES, &tmB, fb, n0 + 64 * s, k0);
          } else {
            tma_load_2d(b_dst, &tmB, fb, k0, n0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = umma_idesc_bf16(BM, BN, p.a_mn, p.b_mn);
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(p, tile);
      mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = t.kb0; kb < t.kb1; ++kb) {
        mbar_wait(full_bar + 8 * stage, phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_base = sA + stage * A_BYTES;
          const uint32_t b_base = sB + stage * Cfg::B_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // K-major: advance 16 elements (32 B) inside the 128 B swizzle row.
            // MN-major: advance 16 k-rows (2 KiB); 64-wide MN atoms are SLAB_BYTES apart (LBO).
            const uint64_t adesc = p.a_mn ? umma_smem_desc_sw128(a_base + k * 2048, SLAB_BYTES, 1024)
                                          : umma_smem_desc_sw128(a_base + k * 32, 0, 1024);
            const uint64_t bdesc = p.b_mn ? umma_smem_desc_sw128(b_base + k * 2048, SLAB_BYTES, 1024)
                                          : umma_smem_desc_sw128(b_base + k * 32, 0, 1024);
            umma_bf16_ss(d_tmem, adesc, bdesc, idesc, (kb > t.kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar + 8 * stage);                  // frees the smem slot when these MMAs retire
          if (kb == t.kb1 - 1) umma_commit(tfull_bar + 8 * acc);  // accumulator complete -> epilogue
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue: 8 warps
    // warp w may only touch TMEM lanes 32*(w%4)..+31, so warps 4-7 take the left half of the tile's columns and
    // warps 8-11 the right half: every accumulator row is finished by two threads.
    const int ew = warp - 4;
    const int wq = ew & 3;
    const int half = ew >> 2;
    constexpr int CH = BN / 64;  // 32-column chunks per half
    const uint32_t stg = stg_base + ew * STG_BYTES;
    uint32_t acc = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(p, tile);
      const int row0 = t.m_blk * BM + wq * 32;
      const int row = row0 + lane;
      const bool row_ok = row < p.M;
      const int cbase = t.n_blk * BN + half * (BN / 2);
      mbar_wait(tfull_bar + 8 * acc, acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + acc * BN + half * (BN / 2) + (uint32_t(wq * 32) << 16);
      switch (p.mode) {
        case DB200_EPI_STORE:    epi_store<CH>(p, t_addr, row, row_ok, cbase, stg, row0, lane); break;
        case DB200_EPI_ATOMIC:   epi_atomic<CH>(p, t_addr, row, row_ok, cbase); break;
        case DB200_EPI_RELU_BWD: epi_relu_bwd<CH>(p, t_addr, row, row_ok, cbase, stg, row0, lane); break;
        case DB200_EPI_CE_STATS: epi_ce_stats<CH>(p, t_addr, row, row_ok, cbase, t.n_blk * 2 + half); break;
        default:                 epi_ce_grad<CH>(p, t_addr, row, row_ok, cbase, stg, row0, lane); break;
      }
      // release the accumulator stage back to the MMA warp
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
static int launch_gemm(cudaStream_t stream, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    DB200_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int total = p.m_tiles * p.n_tiles * p.splits;
  const int grid = total < sm_count() ? total : sm_count();
  gemm_tc_kernel<BN><<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, p);
  return check_launch("gemm_tc_kernel");
}

}  // namespace db200

using namespace db200;

extern "C" int db200_gemm_ce_tiles(int N) { return 2 * ((N + 255) / 256); }

extern "C" int db200_gemm_bf16(db200_stream_t stream_, const void* A, int a_mn_major, int64_t lda, const void* B,
                               int b_mn_major, int64_t ldb, void* D, int64_t ldd, int M, int N, int K,
                               const db200_gemm_epilogue* epi) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DB200_REQUIRE(epi != nullptr, DB200_E_INVALID, "gemm: epilogue descriptor is NULL");
  DB200_REQUIRE(M > 0 && N > 0 && K > 0, DB200_E_INVALID, "gemm: M,N,K must be positive (got %d,%d,%d)", M, N, K);
  DB200_REQUIRE(A && B, DB200_E_INVALID, "gemm: NULL operand");
  DB200_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, DB200_E_ALIGN,
                "gemm: lda/ldb must be multiples of 8 elements (16 B) for TMA (got %lld, %lld)", (long long)lda,
                (long long)ldb);
  DB200_REQUIRE(lda >= (a_mn_major ? M : K) && ldb >= (b_mn_major ? N : K), DB200_E_INVALID,
                "gemm: leading dimension smaller than the contiguous extent");
  const int mode = epi->mode;
  DB200_REQUIRE(mode >= DB200_EPI_STORE && mode <= DB200_EPI_CE_GRAD, DB200_E_INVALID, "gemm: unknown epilogue %d",
                mode);
  if (mode != DB200_EPI_CE_STATS) {
    DB200_REQUIRE(D != nullptr && aligned16(D), DB200_E_ALIGN, "gemm: D must be non-NULL and 16-byte aligned");
    DB200_REQUIRE(ldd >= N, DB200_E_INVALID, "gemm: ldd < N");
  }
  if (mode == DB200_EPI_STORE || mode == DB200_EPI_RELU_BWD || mode == DB200_EPI_CE_GRAD) {
    DB200_REQUIRE(N % 8 == 0 && ldd % 8 == 0, DB200_E_ALIGN,
                  "gemm: vectorised store epilogues need N and ldd to be multiples of 8 (got N=%d ldd=%lld)", N,
                  (long long)ldd);
  }
  DB200_REQUIRE(aligned16(epi->bias), DB200_E_ALIGN, "gemm: bias must be 16-byte aligned");
  if (mode == DB200_EPI_STORE && epi->residual)
    DB200_REQUIRE(aligned16(epi->residual) && epi->ldr % 8 == 0 && epi->ldr >= N, DB200_E_ALIGN,
                  "gemm: residual must be 16-byte aligned with ldr %% 8 == 0");
  if (mode == DB200_EPI_RELU_BWD)
    DB200_REQUIRE(epi->aux && aligned16(epi->aux) && epi->ldaux % 8 == 0 && epi->ldaux >= N, DB200_E_ALIGN,
                  "gemm: RELU_BWD needs an aligned aux tensor");
  if (mode == DB200_EPI_CE_STATS)
    DB200_REQUIRE(epi->labels && epi->part_max && epi->part_sum && epi->label_logit && epi->n_valid > 0 &&
                      epi->n_valid <= N,
                  DB200_E_INVALID, "gemm: CE_STATS needs labels/part_max/part_sum/label_logit and 0 < n_valid <= N");
  if (mode == DB200_EPI_CE_GRAD)
    DB200_REQUIRE(epi->labels && epi->lse && epi->n_valid > 0 && epi->n_valid <= N, DB200_E_INVALID,
                  "gemm: CE_GRAD needs labels/lse and 0 < n_valid <= N");
  int splits = 1;
  const int kb_total = (K + BK - 1) / BK;
  if (mode == DB200_EPI_ATOMIC) {
    splits = epi->split_k;
    if (splits <= 0) {  // auto: fill the machine, keep >= 4 k-blocks per split
      const int tiles = ((M + BM - 1) / BM) * ((N + 255) / 256);
      splits = sm_count() / (tiles > 0 ? tiles : 1);
      if (splits > kb_total / 4) splits = kb_total / 4;
      if (splits < 1) splits = 1;
    }
    if (splits > kb_total) splits = kb_total;
    // every split must own at least one k-block
    const int per = (kb_total + splits - 1) / splits;
    splits = (kb_total + per - 1) / per;
  } else {
    DB200_REQUIRE(epi->split_k <= 1, DB200_E_INVALID, "gemm: split_k > 1 is only valid with DB200_EPI_ATOMIC");
  }

  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.a_mn = a_mn_major ? 1 : 0;
  p.b_mn = b_mn_major ? 1 : 0;
  p.m_tiles = (M + BM - 1) / BM;
  p.splits = splits;
  p.kb_total = kb_total;
  p.mode = mode;
  p.out_f32 = epi->out_f32;
  p.relu = epi->relu;
  p.alpha = epi->alpha;
  p.D = D;
  p.ldd = ldd;
  p.bias = epi->bias;
  p.residual = reinterpret_cast<const bf16*>(epi->residual);
  p.ldr = epi->ldr;
  p.aux = reinterpret_cast<const bf16*>(epi->aux);
  p.ldaux = epi->ldaux;
  p.labels = epi->labels;
  p.part_max = epi->part_max;
  p.part_sum = epi->part_sum;
  p.label_logit = epi->label_logit;
  p.lse = epi->lse;
  p.n_valid = epi->n_valid;
  p.colsum = (mode == DB200_EPI_CE_GRAD || mode == DB200_EPI_RELU_BWD) ? epi->colsum : nullptr;

  // tile width: 256 unless that leaves most SMs idle (CE epilogues are defined on 256-wide tiles)
  int bn = 256;
  if (mode != DB200_EPI_CE_STATS && mode != DB200_EPI_CE_GRAD) {
    const int tiles256 = p.m_tiles * ((N + 255) / 256) * splits;
    if (N <= 128 || tiles256 * 2 <= sm_count()) bn = 128;
  }
  p.n_tiles = (N + bn - 1) / bn;
  p.n_parts = 2 * p.n_tiles;

  CUtensorMap tmA, tmB;
  int rc;
  if (p.a_mn) rc = make_tmap_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, BK);
  else        rc = make_tmap_2d(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, BM);
  if (rc != DB200_OK) return rc;
  if (p.b_mn) rc = make_tmap_2d(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, BK);
  else        rc = make_tmap_2d(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BK, (uint32_t)bn);
  if (rc != DB200_OK) return rc;

  if (bn == 256) return launch_gemm<256>(stream, tmA, tmB, p);
  return launch_gemm<128>(stream, tmA, tmB, p);
}
