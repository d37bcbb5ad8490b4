// N4 — incremental (KV-cache) decoding kernels: the "is_incremental_inference" path the reference sketches in
// src/dalle_mtf/models.py:246-254, 281-285 (new k/v blended into the cached states at context.position - 1, the query of
// the current position attends to keys <= position) but never wires up (PREDICT raises NotImplementedError,
// src/model_fns.py:135-136).  All of it is HBM / latency bound work on B rows per step.
//
//   embed_fwd_at   x[b,:] = wte[ids[b],:] + wpe[pos,:]                     (models.py:186-219 at one position)
//   attn_decode    append this step's k,v to the cache at `pos`, then out[b,h,:] = softmax_j<=pos(scale q.k_j) v_j
//                  fp32 scores / softmax, P rounded to bf16 for the PV product (same policy as attn.cu)
//   sample_rows    idx = lo + argmax_{lo<=c<hi}(logits[c] * inv_temp + gumbel(u[c-lo]))   (first maximum; u NULL = greedy)
//   onehot_rows    y[r, idx[r]-offset] = 1, 0 elsewhere (codebook lookup input for the VAE decoder, vae_tf/models.py:127)
#include "common.cuh"
#include "ptx.cuh"

namespace db200 {

// pos_dev != NULL: the position is read from device memory (CUDA-graph replay: one captured step serves every position)
// and the token of row b is ids[b * ld_ids + pos] (the [B][S] token matrix itself); else ids[b * ld_ids] at `pos`.
__global__ void embed_at_kernel(const int* __restrict__ ids, const bf16* __restrict__ wte, const bf16* __restrict__ wpe,
                                bf16* __restrict__ out, int d, int V, int pos, const int* __restrict__ pos_dev,
                                long long ld_ids) {
  const int b = blockIdx.x;
  if (pos_dev) pos = *pos_dev;
  int id = ids[b * ld_ids + (pos_dev ? pos : 0)];
  id = id < 0 ? 0 : (id >= V ? V - 1 : id);
  const bf16* w = wte + (long long)id * d;
  const bf16* p = wpe + (long long)pos * d;
  for (int i = threadIdx.x; i < d; i += blockDim.x)
    out[(long long)b * d + i] = __float2bfloat16(__bfloat162float(w[i]) + __bfloat162float(p[i]));
}

// one CTA per (head, batch); 8 warps.  Keys are split over the warps in BOTH passes (scores, then P.V), a warp reads a
// whole cached row per instruction (dh * 2 contiguous bytes) and keeps four rows in flight: the per-thread serial walk
// over `pos` keys of the first version made generation latency-bound (~100 us per call at pos ~ 1000).
// Scores live in dynamic shared memory (pos + 1 floats, or S when the position is read on the device).
template <int DH>
__global__ void __launch_bounds__(256)
attn_decode_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ kc, bf16* __restrict__ vc, bf16* __restrict__ out,
                   int S, int H, int pos, float scale, const int* __restrict__ pos_dev) {
  extern __shared__ float sc[];
  __shared__ float red[8];
  __shared__ float part[8][DH];
  if (pos_dev) pos = *pos_dev;
  constexpr int EPL = DH / 32;  // elements per lane (4 or 2): lane l owns [l * EPL, (l + 1) * EPL) of a row
  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bf16* q = qkv + ((long long)b * 3 + 0) * H * DH + (long long)h * DH;
  const bf16* k = qkv + ((long long)b * 3 + 1) * H * DH + (long long)h * DH;
  const bf16* v = qkv + ((long long)b * 3 + 2) * H * DH + (long long)h * DH;
  bf16* kcb = kc + ((long long)b * S * H + h) * DH;  // + j * H * DH
  bf16* vcb = vc + ((long long)b * S * H + h) * DH;
  const long long rs = (long long)H * DH;            // row stride of the caches
  for (int t = tid; t < DH; t += 256) {
    kcb[(long long)pos * rs + t] = k[t];
    vcb[(long long)pos * rs + t] = v[t];
  }
  float qr[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) qr[e] = __bfloat162float(q[lane * EPL + e]);
  __syncthreads();  // this block's own global writes are visible to it after the barrier
  const int n = pos + 1;
  auto load_row = [&](const bf16* base, int j, float* o) {
    if (EPL == 4) {
      const uint2 u = *reinterpret_cast<const uint2*>(base + (long long)j * rs + lane * 4);
      const float2 a = unpack_bf16x2(u.x), c = unpack_bf16x2(u.y);
      o[0] = a.x; o[1] = a.y; o[2] = c.x; o[3] = c.y;
    } else {
      const float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(base + (long long)j * rs + lane * 2));
      o[0] = a.x; o[1] = a.y;
    }
  };
  // ---- scores: warp w takes keys w, w + 8, ... four at a time
  float mx = -INFINITY;
  for (int j0 = warp; j0 < n; j0 += 32) {
    float kr[4][EPL], s[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (j0 + 8 * u < n) load_row(kcb, j0 + 8 * u, kr[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      s[u] = 0.f;
      if (j0 + 8 * u < n) {
#pragma unroll
        for (int e = 0; e < EPL; ++e) s[u] += qr[e] * kr[u][e];
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float t = warp_sum(s[u]) * scale;
      if (j0 + 8 * u < n) {
        if (lane == 0) sc[j0 + 8 * u] = t;
        mx = fmaxf(mx, t);
      }
    }
  }
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float sum = 0.f;
  for (int j = tid; j < n; j += 256) {
    const float p = __expf(sc[j] - mx);
    sum += p;
    sc[j] = __bfloat162float(__float2bfloat16(p));  // P goes through bf16 for the PV product, the sum does not
  }
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) sum += red[w];
  // ---- P.V: warp w accumulates its keys' contributions to all dh outputs, then the 8 partials are summed
  float acc[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) acc[e] = 0.f;
  for (int j0 = warp; j0 < n; j0 += 32) {
    float vr[4][EPL];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (j0 + 8 * u < n) load_row(vcb, j0 + 8 * u, vr[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (j0 + 8 * u < n) {
        const float p = sc[j0 + 8 * u];
#pragma unroll
        for (int e = 0; e < EPL; ++e) acc[e] += p * vr[u][e];
      }
  }
#pragma unroll
  for (int e = 0; e < EPL; ++e) part[warp][lane * EPL + e] = acc[e];
  __syncthreads();
  const float inv = 1.f / sum;
  for (int t = tid; t < DH; t += 256) {
    float o = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) o += part[w][t];
    out[((long long)b * H + h) * DH + t] = __float2bfloat16(o * inv);
  }
}

__global__ void __launch_bounds__(256)
sample_rows_kernel(const float* __restrict__ logits, const float* __restrict__ u, int* __restrict__ idx, long long ld,
                   int lo, int hi, float inv_temp, const int* __restrict__ pos_dev, long long ld_idx) {
  __shared__ float bv[8];
  __shared__ int bi[8];
  const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* row = logits + (long long)r * ld;
  const int n = hi - lo;
  float best = -INFINITY;
  int arg = 0x7fffffff;
  for (int c = tid; c < n; c += 256) {
    float x = row[lo + c] * inv_temp;
    if (u) x += -logf(-logf(u[(long long)r * n + c]));
    if (x > best || (x == best && c < arg)) { best = x; arg = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, arg, o);
    if (ov > best || (ov == best && oi < arg)) { best = ov; arg = oi; }
  }
  if (lane == 0) { bv[warp] = best; bi[warp] = arg; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 8; ++w)
      if (bv[w] > best || (bv[w] == best && bi[w] < arg)) { best = bv[w]; arg = bi[w]; }
    // pos_dev != NULL (graph replay): the sample is the token of position *pos_dev + 1 of the [rows][ld_idx] matrix
    int* dst = pos_dev ? idx + r * ld_idx + (*pos_dev + 1) : idx + r;
    *dst = lo + (arg == 0x7fffffff ? 0 : arg);  // all -inf / NaN row: first allowed id
  }
}

__global__ void incr_i32_kernel(int* p, int delta) { *p += delta; }

__global__ void onehot_rows_kernel(const int* __restrict__ idx, float* __restrict__ y, int K, int offset) {
  const int r = blockIdx.x;
  const int hot = idx[r] - offset;
  for (int c = threadIdx.x; c < K; c += blockDim.x) y[(long long)r * K + c] = (c == hot) ? 1.f : 0.f;
}

}  // namespace db200

using namespace db200;

extern "C" int db200_embed_fwd_at(db200_stream_t stream_, const int32_t* ids, const void* wte, const void* wpe,
                                  void* out, int B, int d, int V, int n_pos, int pos) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DB200_REQUIRE(ids && wte && wpe && out, DB200_E_INVALID, "embed_fwd_at: null pointer");
  DB200_REQUIRE(B > 0 && d > 0 && V > 0 && pos >= 0 && pos < n_pos, DB200_E_INVALID,
                "embed_fwd_at: B %d d %d V %d pos %d of %d", B, d, V, pos, n_pos);
  embed_at_kernel<<<B, 256, 0, stream>>>(ids, static_cast<const bf16*>(wte), static_cast<const bf16*>(wpe),
                                         static_cast<bf16*>(out), d, V, pos, nullptr, 1);
  return check_launch("embed_at_kernel");
}

// ---- device-side position variants (CUDA-graph replay of the per-position step): `pos_dev` points at one int32 in
// device memory that the caller advances with db200_incr_i32 between replays; `tokens` is the [B][ld] token matrix.
extern "C" int db200_embed_fwd_at_dev(db200_stream_t stream_, const int32_t* tokens, long long ld, const void* wte,
                                      const void* wpe, void* out, int B, int d, int V, const int32_t* pos_dev) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DB200_REQUIRE(tokens && wte && wpe && out && pos_dev && B > 0 && d > 0 && V > 0 && ld > 0, DB200_E_INVALID,
                "embed_fwd_at_dev: bad argument");
  embed_at_kernel<<<B, 256, 0, stream>>>(tokens, static_cast<const bf16*>(wte), static_cast<const bf16*>(wpe),
                                         static_cast<bf16*>(out), d, V, 0, pos_dev, ld);
  return check_launch("embed_at_kernel");
}

extern "C" int db200_incr_i32(db200_stream_t stream_, int32_t* p, int delta) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DB200_REQUIRE(p, DB200_E_INVALID, "incr_i32: null pointer");
  incr_i32_kernel<<<1, 1, 0, stream>>>(p, delta);
  return check_launch("incr_i32_kernel");
}

// pos_dev == NULL: position `pos` (scores need pos + 1 floats of shared memory); else the position is read on the
// device and shared memory is sized for all S keys.
static int attn_decode_launch(cudaStream_t stream, const void* qkv_step, void* k_cache, void* v_cache, void* out, int B,
                              int S, int H, int dh, int pos, float scale, const int32_t* pos_dev) {
  DB200_REQUIRE(qkv_step && k_cache && v_cache && out, DB200_E_INVALID, "attn_decode: null pointer");
  DB200_REQUIRE(B > 0 && H > 0 && S > 0 && pos >= 0 && pos < S, DB200_E_INVALID, "attn_decode: B %d H %d pos %d of %d",
                B, H, pos, S);
  DB200_REQUIRE(dh == 64 || dh == 128, DB200_E_UNSUPPORTED, "attn_decode: head_dim %d not in {64,128}", dh);
  const size_t keys = pos_dev ? (size_t)S : (size_t)(pos + 1);
  DB200_REQUIRE(keys * 4 <= 200 * 1024, DB200_E_UNSUPPORTED, "attn_decode: %zu keys exceed shared memory", keys);
  const size_t smem = keys * sizeof(float);
  dim3 grid(H, B);
  static const cudaError_t a128 = cudaFuncSetAttribute(attn_decode_kernel<128>,
                                                       cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  static const cudaError_t a64 = cudaFuncSetAttribute(attn_decode_kernel<64>,
                                                      cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  DB200_CUDA(a128);
  DB200_CUDA(a64);
  if (dh == 128)
    attn_decode_kernel<128><<<grid, 256, smem, stream>>>(static_cast<const bf16*>(qkv_step),
                                                         static_cast<bf16*>(k_cache), static_cast<bf16*>(v_cache),
                                                         static_cast<bf16*>(out), S, H, pos, scale, pos_dev);
  else
    attn_decode_kernel<64><<<grid, 256, smem, stream>>>(static_cast<const bf16*>(qkv_step), static_cast<bf16*>(k_cache),
                                                        static_cast<bf16*>(v_cache), static_cast<bf16*>(out), S, H, pos,
                                                        scale, pos_dev);
  return check_launch("attn_decode_kernel");
}

extern "C" int db200_attn_decode(db200_stream_t stream_, const void* qkv_step, void* k_cache, void* v_cache, void* out,
                                 int B, int S, int H, int dh, int pos, float scale) {
  return attn_decode_launch(static_cast<cudaStream_t>(stream_), qkv_step, k_cache, v_cache, out, B, S, H, dh, pos, scale,
                            nullptr);
}

extern "C" int db200_attn_decode_dev(db200_stream_t stream_, const void* qkv_step, void* k_cache, void* v_cache,
                                     void* out, int B, int S, int H, int dh, const int32_t* pos_dev, float scale) {
  DB200_REQUIRE(pos_dev, DB200_E_INVALID, "attn_decode_dev: null position pointer");
  return attn_decode_launch(static_cast<cudaStream_t>(stream_), qkv_step, k_cache, v_cache, out, B, S, H, dh, 0, scale,
                            pos_dev);
}

extern "C" int db200_sample_rows(db200_stream_t stream_, const float* logits, const float* u_or_null, int32_t* idx,
                                 int rows, long long ld, int lo, int hi, float inv_temp) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DB200_REQUIRE(logits && idx, DB200_E_INVALID, "sample_rows: null pointer");
  DB200_REQUIRE(rows > 0 && lo >= 0 && hi > lo && hi <= ld, DB200_E_INVALID, "sample_rows: rows %d range [%d,%d) ld %lld",
                rows, lo, hi, ld);
  DB200_REQUIRE(inv_temp > 0.f, DB200_E_INVALID, "sample_rows: inv_temp must be positive");
  sample_rows_kernel<<<rows, 256, 0, stream>>>(logits, u_or_null, idx, ld, lo, hi, inv_temp, nullptr, 1);
  return check_launch("sample_rows_kernel");
}

// graph-replay variant: the sample of row r becomes tokens[r * ld_tokens + *pos_dev + 1]
extern "C" int db200_sample_rows_at(db200_stream_t stream_, const float* logits, const float* u_or_null,
                                    int32_t* tokens, long long ld_tokens, int rows, long long ld, int lo, int hi,
                                    float inv_temp, const int32_t* pos_dev) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DB200_REQUIRE(logits && tokens && pos_dev, DB200_E_INVALID, "sample_rows_at: null pointer");
  DB200_REQUIRE(rows > 0 && lo >= 0 && hi > lo && hi <= ld && inv_temp > 0.f, DB200_E_INVALID,
                "sample_rows_at: rows %d range [%d,%d) ld %lld", rows, lo, hi, ld);
  sample_rows_kernel<<<rows, 256, 0, stream>>>(logits, u_or_null, tokens, ld, lo, hi, inv_temp, pos_dev, ld_tokens);
  return check_launch("sample_rows_kernel");
}

extern "C" int db200_onehot_rows_f32(db200_stream_t stream_, const int32_t* idx, float* y, int rows, int K,
                                     int offset) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DB200_REQUIRE(idx && y && rows > 0 && K > 0, DB200_E_INVALID, "onehot_rows: bad argument");
  onehot_rows_kernel<<<rows, 256, 0, stream>>>(idx, y, K, offset);
  return check_launch("onehot_rows_kernel");
}
