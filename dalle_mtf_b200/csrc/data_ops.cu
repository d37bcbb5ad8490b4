// N2/N3 — the data formats either side of the training step.
//
//   * CRC-32C (Castagnoli) and TFRecord framing: host code (the records live in host memory), slice-by-8 tables.
//     Replaces tf.io.TFRecordWriter / tf.data.TFRecordDataset as used by src/data/create_tfrecords.py:153-178 and
//     src/input_fns.py:80,116.
//   * centre-crop + bilinear resize + normalisation of a batch of decoded uint8 images: one CUDA kernel.  Replaces
//     tf.image.crop_and_resize + the (x - 127.5) / 127.5 scaling of src/input_fns.py:4-21 (crop_center_and_resize,
//     decode_img).  HBM-bound byte work: every output pixel reads 4 neighbours x C bytes and writes C floats.
#include <cmath>
#include <cstring>

#include "common.cuh"

namespace db200 {

// ------------------------------------------------------------------------------------------------ CRC-32C
struct Crc32cTables {
  uint32_t t[8][256];
  Crc32cTables() {
    const uint32_t poly = 0x82F63B78u;  // reflected 0x1EDC6F41
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ poly : (c >> 1);
      t[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xFF];
  }
};

static uint32_t crc32c_update(uint32_t crc, const uint8_t* p, size_t n) {
  static const Crc32cTables T;
  crc = ~crc;
  while (n && (reinterpret_cast<uintptr_t>(p) & 7)) {
    crc = (crc >> 8) ^ T.t[0][(crc ^ *p++) & 0xFF];
    --n;
  }
  while (n >= 8) {
    uint64_t w;
    memcpy(&w, p, 8);
    w ^= crc;  // little-endian host (x86-64 / aarch64)
    crc = T.t[7][w & 0xFF] ^ T.t[6][(w >> 8) & 0xFF] ^ T.t[5][(w >> 16) & 0xFF] ^ T.t[4][(w >> 24) & 0xFF] ^
          T.t[3][(w >> 32) & 0xFF] ^ T.t[2][(w >> 40) & 0xFF] ^ T.t[1][(w >> 48) & 0xFF] ^ T.t[0][(w >> 56) & 0xFF];
    p += 8;
    n -= 8;
  }
  while (n--) crc = (crc >> 8) ^ T.t[0][(crc ^ *p++) & 0xFF];
  return ~crc;
}

static inline uint32_t tfrecord_mask(uint32_t crc) { return ((crc >> 15) | (crc << 17)) + 0xa282ead8u; }

// ------------------------------------------------------------------------------------------------ crop + resize
// TensorFlow's CropAndResize (bilinear, extrapolation_value 0) for one crop box per image, then (v - 127.5) / 127.5.
//   in_y = y1 (H-1) + y * (y2 - y1)(H-1)/(S-1)        (S > 1; S == 1: the box centre)
//   outside [0, H-1] (or [0, W-1]) the output is the extrapolation value 0 -> normalised -1
//   top = floor, bottom = ceil, lerp in fp32: top + (bottom - top) * frac, x first then y
// `box` holds {y1, x1, y2, x2} per image (normalised coordinates, computed by the host exactly as the reference does).
template <int C>
__global__ void __launch_bounds__(256)
crop_resize_kernel(const uint8_t* __restrict__ packed, const long long* __restrict__ offsets,
                   const int* __restrict__ heights, const int* __restrict__ widths, const float* __restrict__ box,
                   float* __restrict__ out, int S) {
  const int b = blockIdx.y;
  const int H = heights[b], W = widths[b];
  const uint8_t* img = packed + offsets[b];
  const float y1 = box[4 * b], x1 = box[4 * b + 1], y2 = box[4 * b + 2], x2 = box[4 * b + 3];
  // every product / sum is a separately rounded fp32 operation (the __f*_rn intrinsics are never contracted into FMAs):
  // the TensorFlow CPU kernel this replaces is plain scalar float code, and the oracle is numpy float32.
  const float Hm1 = float(H - 1), Wm1 = float(W - 1), Sm1 = float(S - 1);
  const float hs = S > 1 ? __fdiv_rn(__fmul_rn(__fsub_rn(y2, y1), Hm1), Sm1) : 0.f;
  const float ws = S > 1 ? __fdiv_rn(__fmul_rn(__fsub_rn(x2, x1), Wm1), Sm1) : 0.f;
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < S * S; pix += gridDim.x * blockDim.x) {
    const int y = pix / S, x = pix - y * S;
    const float in_y = S > 1 ? __fadd_rn(__fmul_rn(y1, Hm1), __fmul_rn(float(y), hs))
                             : __fmul_rn(__fmul_rn(0.5f, __fadd_rn(y1, y2)), Hm1);
    const float in_x = S > 1 ? __fadd_rn(__fmul_rn(x1, Wm1), __fmul_rn(float(x), ws))
                             : __fmul_rn(__fmul_rn(0.5f, __fadd_rn(x1, x2)), Wm1);
    float v[C];
    if (in_y < 0.f || in_y > Hm1 || in_x < 0.f || in_x > Wm1) {
#pragma unroll
      for (int c = 0; c < C; ++c) v[c] = 0.f;
    } else {
      const int top = (int)floorf(in_y), bot = (int)ceilf(in_y);
      const int lft = (int)floorf(in_x), rgt = (int)ceilf(in_x);
      const float fy = __fsub_rn(in_y, float(top)), fx = __fsub_rn(in_x, float(lft));
      const uint8_t* tl = img + ((long long)top * W + lft) * C;
      const uint8_t* tr = img + ((long long)top * W + rgt) * C;
      const uint8_t* bl = img + ((long long)bot * W + lft) * C;
      const uint8_t* br = img + ((long long)bot * W + rgt) * C;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float a = tl[c], bb = tr[c], cc = bl[c], d = br[c];
        const float t = __fadd_rn(a, __fmul_rn(__fsub_rn(bb, a), fx));
        const float u = __fadd_rn(cc, __fmul_rn(__fsub_rn(d, cc), fx));
        v[c] = __fadd_rn(t, __fmul_rn(__fsub_rn(u, t), fy));
      }
    }
    float* o = out + (((long long)b * S + y) * S + x) * C;
#pragma unroll
    for (int c = 0; c < C; ++c) o[c] = __fdiv_rn(__fsub_rn(v[c], 127.5f), 127.5f);
  }
}

}  // namespace db200

using namespace db200;

extern "C" int db200_crc32c(const void* data, uint64_t n, uint32_t* crc_out) {
  DB200_REQUIRE((data || n == 0) && crc_out, DB200_E_INVALID, "crc32c: null pointer");
  *crc_out = crc32c_update(0, static_cast<const uint8_t*>(data), (size_t)n);
  return DB200_OK;
}

extern "C" int db200_tfrecord_masked_crc(const void* data, uint64_t n, uint32_t* crc_out) {
  DB200_REQUIRE((data || n == 0) && crc_out, DB200_E_INVALID, "masked_crc: null pointer");
  *crc_out = tfrecord_mask(crc32c_update(0, static_cast<const uint8_t*>(data), (size_t)n));
  return DB200_OK;
}

// Frame one record: out must hold n + 16 bytes.  Layout: u64 length | u32 masked crc(length) | data | u32 masked crc(data)
extern "C" int db200_tfrecord_frame(const void* data, uint64_t n, void* out) {
  DB200_REQUIRE((data || n == 0) && out, DB200_E_INVALID, "tfrecord_frame: null pointer");
  uint8_t* o = static_cast<uint8_t*>(out);
  memcpy(o, &n, 8);
  const uint32_t lc = tfrecord_mask(crc32c_update(0, o, 8));
  memcpy(o + 8, &lc, 4);
  if (n) memcpy(o + 12, data, n);
  const uint32_t dc = tfrecord_mask(crc32c_update(0, static_cast<const uint8_t*>(data), (size_t)n));
  memcpy(o + 12 + n, &dc, 4);
  return DB200_OK;
}

// Scan a TFRecord file image: payload offsets / lengths of up to max_records records.  A truncated or corrupt record
// is an error (TF raises DataLossError), never silently skipped.
extern "C" int db200_tfrecord_index(const void* buf_, uint64_t n, int verify_crc, uint64_t* offsets, uint64_t* lengths,
                                    uint64_t max_records, uint64_t* n_records) {
  DB200_REQUIRE((buf_ || n == 0) && n_records && (max_records == 0 || (offsets && lengths)), DB200_E_INVALID,
                "tfrecord_index: null pointer");
  const uint8_t* buf = static_cast<const uint8_t*>(buf_);
  uint64_t pos = 0, cnt = 0;
  while (pos < n) {
    DB200_REQUIRE(n - pos >= 12, DB200_E_INVALID, "tfrecord_index: truncated record header at byte %llu",
                  (unsigned long long)pos);
    uint64_t len;
    uint32_t lc;
    memcpy(&len, buf + pos, 8);
    memcpy(&lc, buf + pos + 8, 4);
    if (verify_crc)
      DB200_REQUIRE(tfrecord_mask(crc32c_update(0, buf + pos, 8)) == lc, DB200_E_INVALID,
                    "tfrecord_index: corrupted record length at byte %llu", (unsigned long long)pos);
    DB200_REQUIRE(len <= n - pos - 12 && n - pos - 12 - len >= 4, DB200_E_INVALID,
                  "tfrecord_index: truncated record at byte %llu (length %llu)", (unsigned long long)pos,
                  (unsigned long long)len);
    if (verify_crc) {
      uint32_t dc;
      memcpy(&dc, buf + pos + 12 + len, 4);
      DB200_REQUIRE(tfrecord_mask(crc32c_update(0, buf + pos + 12, (size_t)len)) == dc, DB200_E_INVALID,
                    "tfrecord_index: corrupted record data at byte %llu", (unsigned long long)pos);
    }
    if (cnt < max_records) {
      offsets[cnt] = pos + 12;
      lengths[cnt] = len;
    }
    ++cnt;
    pos += 12 + len + 4;
  }
  *n_records = cnt;
  return DB200_OK;
}

extern "C" int db200_image_crop_resize_normalize(db200_stream_t stream_, const uint8_t* packed,
                                                 const int64_t* offsets, const int32_t* heights,
                                                 const int32_t* widths, const float* boxes, float* out, int batch,
                                                 int channels, int out_size) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DB200_REQUIRE(packed && offsets && heights && widths && boxes && out, DB200_E_INVALID, "crop_resize: null pointer");
  DB200_REQUIRE(batch > 0 && out_size > 0 && (channels == 1 || channels == 3), DB200_E_INVALID,
                "crop_resize: batch %d out_size %d channels %d (need channels in {1,3})", batch, out_size, channels);
  const int pix = out_size * out_size;
  int gx = (pix + 255) / 256;
  if (gx > 64) gx = 64;
  dim3 grid(gx, batch);
  static_assert(sizeof(long long) == sizeof(int64_t), "offset type");
  if (channels == 3)
    crop_resize_kernel<3><<<grid, 256, 0, stream>>>(packed, reinterpret_cast<const long long*>(offsets), heights,
                                                    widths, boxes, out, out_size);
  else
    crop_resize_kernel<1><<<grid, 256, 0, stream>>>(packed, reinterpret_cast<const long long*>(offsets), heights,
                                                    widths, boxes, out, out_size);
  return check_launch("crop_resize_kernel");
}
