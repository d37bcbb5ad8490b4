"""TFRecord files and tf.train.Example messages without TensorFlow ("next" rows N2/N3 of SURVEY.md §8f).

The reference writes its datasets with tf.io.TFRecordWriter + tf.train.Example (src/data/create_tfrecords.py:37-56,
153-178) and reads them back with tf.data.TFRecordDataset + tf.parse_single_example (src/input_fns.py:41-66).  Both
formats are public and small, so they are restated here:

  record  = u64le length | u32le masked_crc32c(length bytes) | payload | u32le masked_crc32c(payload)
  Example = protobuf  { 1: Features { 1: map<string, Feature> } }
  Feature = oneof     { 1: BytesList{1: repeated bytes}, 2: FloatList{1: packed float}, 3: Int64List{1: packed varint} }

CRC-32C and the record framing / scanning are native (libdalle_b200.so, csrc/data_ops.cu: host code, no GPU needed).
"""
import ctypes
import struct

from . import lib as L


# ------------------------------------------------------------------------------------------------------ records
def crc32c(data: bytes) -> int:
    out = ctypes.c_uint32()
    L.check(L.load().db200_crc32c(data, len(data), ctypes.byref(out)), "db200_crc32c")
    return out.value


def masked_crc32c(data: bytes) -> int:
    out = ctypes.c_uint32()
    L.check(L.load().db200_tfrecord_masked_crc(data, len(data), ctypes.byref(out)), "db200_tfrecord_masked_crc")
    return out.value


def frame_record(payload: bytes) -> bytes:
    buf = ctypes.create_string_buffer(len(payload) + 16)
    L.check(L.load().db200_tfrecord_frame(payload, len(payload), buf), "db200_tfrecord_frame")
    return buf.raw


class TFRecordWriter:
    """tf.io.TFRecordWriter(path) with .write(bytes) / .close() / context manager (uncompressed files only)."""

    def __init__(self, path):
        self._f = open(path, "wb")

    def write(self, record: bytes):
        self._f.write(frame_record(bytes(record)))

    def flush(self):
        self._f.flush()

    def close(self):
        if self._f is not None:
            self._f.close()
            self._f = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def index_records(buf: bytes, verify_crc=True):
    """[(offset, length)] of every record payload in a TFRecord file image; raises on truncation / corruption."""
    lib = L.load()
    n = ctypes.c_uint64()
    cap = max(16, len(buf) // 4096)
    while True:
        offs = (ctypes.c_uint64 * cap)()
        lens = (ctypes.c_uint64 * cap)()
        L.check(lib.db200_tfrecord_index(buf, len(buf), 1 if verify_crc else 0, offs, lens, cap, ctypes.byref(n)),
                "db200_tfrecord_index")
        if n.value <= cap:
            return [(offs[i], lens[i]) for i in range(n.value)]
        cap = n.value


def tfrecord_iterator(path, verify_crc=True):
    """Payloads of one TFRecord file, in file order (tf.data.TFRecordDataset(path))."""
    with open(path, "rb") as f:
        buf = f.read()
    for off, ln in index_records(buf, verify_crc):
        yield buf[off:off + ln]


# ------------------------------------------------------------------------------------------------------ protobuf
def _varint(v: int) -> bytes:
    if v < 0:
        v += 1 << 64  # int64 two's complement: ten bytes on the wire
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _read_varint(buf, pos):
    shift = result = 0
    while True:
        if pos >= len(buf):
            raise ValueError("truncated varint")
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 63:
            raise ValueError("varint too long")


def _ld(field: int, payload: bytes) -> bytes:  # length-delimited field
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def bytes_feature(values):
    if isinstance(values, (bytes, bytearray)):
        values = [values]
    return ("bytes", [bytes(v) for v in values])


def int64_feature(values):
    return ("int64", [int(v) for v in values])


def float_feature(values):
    return ("float", [float(v) for v in values])


def _encode_feature(kind, values) -> bytes:
    if kind == "bytes":
        return _ld(1, b"".join(_ld(1, v) for v in values))
    if kind == "float":
        inner = _ld(1, struct.pack(f"<{len(values)}f", *values)) if values else b""
        return _ld(2, inner)
    if kind == "int64":
        inner = _ld(1, b"".join(_varint(v) for v in values)) if values else b""  # packed, as protobuf writes it
        return _ld(3, inner)
    raise ValueError(f"unknown feature kind {kind!r}")


def encode_example(features: dict) -> bytes:
    """tf.train.Example(features=tf.train.Features(feature=features)).SerializeToString().
    `features`: name -> ("bytes"|"int64"|"float", [values]).  Map entries are written in sorted key order (for the
    reference's {"caption", "image"} records that is byte-identical to protobuf's deterministic serialisation; in
    general the order of map entries carries no meaning and every protobuf parser accepts any order)."""
    entries = b""
    for name in sorted(features):
        kind, values = features[name]
        entry = _ld(1, name.encode("utf-8")) + _ld(2, _encode_feature(kind, values))
        entries += _ld(1, entry)
    return _ld(1, entries)


def _fields(buf):
    pos = 0
    while pos < len(buf):
        key, pos = _read_varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _read_varint(buf, pos)
        elif wt == 1:
            v, pos = buf[pos:pos + 8], pos + 8
        elif wt == 2:
            ln, pos = _read_varint(buf, pos)
            if pos + ln > len(buf):
                raise ValueError("truncated length-delimited field")
            v, pos = buf[pos:pos + ln], pos + ln
        elif wt == 5:
            v, pos = buf[pos:pos + 4], pos + 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        yield field, wt, v


def _decode_feature(buf):
    for field, wt, v in _fields(buf):
        if field == 1 and wt == 2:
            return "bytes", [bytes(x) for f, w, x in _fields(v) if f == 1 and w == 2]
        if field == 2 and wt == 2:
            vals = []
            for f, w, x in _fields(v):
                if f == 1 and w == 2:
                    vals += list(struct.unpack(f"<{len(x) // 4}f", x))
                elif f == 1 and w == 5:
                    vals.append(struct.unpack("<f", x)[0])
            return "float", vals
        if field == 3 and wt == 2:
            vals = []
            for f, w, x in _fields(v):
                if f == 1 and w == 2:  # packed
                    p = 0
                    while p < len(x):
                        q, p = _read_varint(x, p)
                        vals.append(q - (1 << 64) if q >= 1 << 63 else q)
                elif f == 1 and w == 0:  # unpacked (older writers)
                    vals.append(x - (1 << 64) if x >= 1 << 63 else x)
            return "int64", vals
    return "bytes", []  # empty Feature: kind unset


def decode_example(buf: bytes) -> dict:
    """Inverse of encode_example: name -> (kind, [values]).  Accepts packed and unpacked repeated scalars."""
    out = {}
    for field, wt, feats in _fields(buf):
        if field != 1 or wt != 2:
            continue
        for f, w, entry in _fields(feats):
            if f != 1 or w != 2:
                continue
            name, feat = None, b""
            for ef, ew, ev in _fields(entry):
                if ef == 1 and ew == 2:
                    name = bytes(ev).decode("utf-8")
                elif ef == 2 and ew == 2:
                    feat = ev
            if name is not None:
                out[name] = _decode_feature(feat)
    return out
