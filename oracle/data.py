"""ORACLE (test infrastructure only — never imported by the product path): CPU restatement of the reference's data
formats and image preprocessing for the "next" rows N2/N3.  PARITY UNPINNED for the TensorFlow pieces: tf 2.4 cannot
be installed here and the reference ships no fixtures; pinned instead by public known-answer vectors (CRC-32C: RFC 3720
B.4) and by the protobuf runtime (google.protobuf) for the tf.train.Example encoding.

  crc32c / masked_crc / tfrecord_frame   TFRecord wire format used by tf.io.TFRecordWriter
                                         (src/data/create_tfrecords.py:153-178; src/input_fns.py:80,116)
  crop_and_resize_bilinear               tf.image.crop_and_resize(img[None], [box], [0], [size, size]) as executed by
                                         TensorFlow's CPU kernel (scalar float32 arithmetic), src/input_fns.py:9-12
  decode_img                             crop_center_and_resize + (x - 127.5) / 127.5, src/input_fns.py:4-21
"""
import struct

import numpy as np


def crc32c(data: bytes) -> int:
    """Bit-at-a-time CRC-32C (Castagnoli, reflected polynomial 0x82F63B78)."""
    crc = 0xFFFFFFFF
    for byte in data:
        crc ^= byte
        for _ in range(8):
            crc = (crc >> 1) ^ 0x82F63B78 if crc & 1 else crc >> 1
    return crc ^ 0xFFFFFFFF


def masked_crc(data: bytes) -> int:
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def tfrecord_frame(payload: bytes) -> bytes:
    ln = struct.pack("<Q", len(payload))
    return ln + struct.pack("<I", masked_crc(ln)) + payload + struct.pack("<I", masked_crc(payload))


def reference_crop_box(height, width):
    """src/input_fns.py:4-10 as written (w, h = shape[0], shape[1]; float64 division; cast to float32 by the op)."""
    w, h = float(height), float(width)
    c = max(w, h)
    wn, hn = h / c, w / c
    return np.array([(1 - wn) / 2, (1 - hn) / 2, wn, hn], dtype=np.float64).astype(np.float32)


def crop_and_resize_bilinear(img_u8, box, size, extrapolation_value=0.0):
    """float32 [size, size, C]; every operation rounded to float32 like the scalar C++ kernel."""
    f = np.float32
    H, W, C = img_u8.shape
    y1, x1, y2, x2 = (f(v) for v in box)
    out = np.empty((size, size, C), dtype=np.float32)
    hs = f(f(f(y2 - y1) * f(H - 1)) / f(size - 1)) if size > 1 else f(0)
    ws = f(f(f(x2 - x1) * f(W - 1)) / f(size - 1)) if size > 1 else f(0)
    img = img_u8.astype(np.float32)
    for y in range(size):
        in_y = f(f(y1 * f(H - 1)) + f(f(y) * hs)) if size > 1 else f(f(f(0.5) * f(y1 + y2)) * f(H - 1))
        if in_y < 0 or in_y > H - 1:
            out[y] = extrapolation_value
            continue
        top, bot = int(np.floor(in_y)), int(np.ceil(in_y))
        fy = f(in_y - f(top))
        for x in range(size):
            in_x = f(f(x1 * f(W - 1)) + f(f(x) * ws)) if size > 1 else f(f(f(0.5) * f(x1 + x2)) * f(W - 1))
            if in_x < 0 or in_x > W - 1:
                out[y, x] = extrapolation_value
                continue
            lft, rgt = int(np.floor(in_x)), int(np.ceil(in_x))
            fx = f(in_x - f(lft))
            tl, tr, bl, br = img[top, lft], img[top, rgt], img[bot, lft], img[bot, rgt]
            t = (tl + ((tr - tl) * fx).astype(np.float32)).astype(np.float32)
            u = (bl + ((br - bl) * fx).astype(np.float32)).astype(np.float32)
            out[y, x] = (t + ((u - t) * fy).astype(np.float32)).astype(np.float32)
    return out


def decode_img(img_u8, size):
    """decode_img after the JPEG decode: crop_center_and_resize then (x - 127.5) / 127.5 in float32."""
    box = reference_crop_box(img_u8.shape[0], img_u8.shape[1])
    r = crop_and_resize_bilinear(img_u8, box, size)
    return ((r - np.float32(127.5)) / np.float32(127.5)).astype(np.float32)


def truncate_or_pad_label(label, text_seq_len, padding_id):
    """src/input_fns.py:32-38."""
    padded = list(label) + [padding_id] * text_seq_len
    return np.asarray(padded[:text_seq_len], dtype=np.int32)
