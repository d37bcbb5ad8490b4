"""CPU tests of the "next" rows N2/N3: TFRecord / tf.train.Example formats, the tf.data-equivalent record pipeline, the
dataset writer, and the data oracle itself (pinned by public known-answer vectors and the protobuf runtime)."""
import io
import os
import struct

import numpy as np
import pytest

from dalle_mtf_b200 import data_pipeline as dp
from dalle_mtf_b200 import dataset_tools, tfrecord
from oracle import data as odata

RFC3720 = [  # RFC 3720 appendix B.4 + the classic check value
    (b"123456789", 0xE3069283),
    (bytes(32), 0x8A9136AA),
    (bytes([0xFF] * 32), 0x62A8AB43),
    (bytes(range(32)), 0x46DD794E),
    (bytes(range(31, -1, -1)), 0x113FDB5C),
    (b"", 0x00000000),
]


def test_crc32c_known_answers_native_and_oracle():
    for data, want in RFC3720:
        assert tfrecord.crc32c(data) == want
        assert odata.crc32c(data) == want


def test_crc32c_native_equals_oracle_on_ragged_inputs():
    rng = np.random.default_rng(0)
    blob = rng.integers(0, 256, 5000, dtype=np.uint8).tobytes()
    for start in range(0, 9):                       # every alignment of the slice-by-8 loop
        for n in (0, 1, 7, 8, 9, 63, 64, 65, 1000, 4097):
            chunk = blob[start:start + n]
            assert tfrecord.crc32c(chunk) == odata.crc32c(chunk)
            assert tfrecord.masked_crc32c(chunk) == odata.masked_crc(chunk)


def test_tfrecord_frame_native_equals_oracle_and_known_layout():
    for payload in (b"", b"x", b"hello world", bytes(range(256)) * 9):
        fr = tfrecord.frame_record(payload)
        assert fr == odata.tfrecord_frame(payload)
        assert struct.unpack("<Q", fr[:8])[0] == len(payload) and len(fr) == len(payload) + 16
    # masked crc of the 8 length bytes of an empty record (value fixed by the format)
    assert struct.unpack("<I", tfrecord.frame_record(b"")[8:12])[0] == odata.masked_crc(bytes(8))


def test_tfrecord_roundtrip_and_corruption_is_an_error(tmp_path):
    path = tmp_path / "a.tfrecords"
    recs = [b"", b"one", os.urandom(70000), b"last"]
    with tfrecord.TFRecordWriter(path) as w:
        for r in recs:
            w.write(r)
    assert list(tfrecord.tfrecord_iterator(path)) == recs
    raw = bytearray(path.read_bytes())
    from dalle_mtf_b200.lib import DB200Error
    bad = bytearray(raw); bad[100] ^= 1                         # inside the payload of the third record
    (tmp_path / "bad.tfrecords").write_bytes(bad)
    with pytest.raises(DB200Error, match="corrupt"):
        list(tfrecord.tfrecord_iterator(tmp_path / "bad.tfrecords"))
    assert len(list(tfrecord.tfrecord_iterator(tmp_path / "bad.tfrecords", verify_crc=False))) == 4
    (tmp_path / "cut.tfrecords").write_bytes(raw[:-3])          # truncated tail
    with pytest.raises(DB200Error, match="truncated"):
        list(tfrecord.tfrecord_iterator(tmp_path / "cut.tfrecords"))
    (tmp_path / "empty.tfrecords").write_bytes(b"")
    assert list(tfrecord.tfrecord_iterator(tmp_path / "empty.tfrecords")) == []


def _protobuf_example_class():
    """tf.train.Example rebuilt from its public .proto with the protobuf runtime (no TensorFlow)."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fd = descriptor_pb2.FileDescriptorProto(name="example_for_test.proto", package="tftest", syntax="proto3")

    def msg(name):
        m = fd.message_type.add()
        m.name = name
        return m

    def field(m, name, number, ftype, label=1, type_name=None, packed=None, oneof=None):
        f = m.field.add()
        f.name, f.number, f.type, f.label = name, number, ftype, label
        if type_name:
            f.type_name = type_name
        if packed is not None:
            f.options.packed = packed
        if oneof is not None:
            f.oneof_index = oneof
    field(msg("BytesList"), "value", 1, 12, label=3)
    field(msg("FloatList"), "value", 1, 2, label=3, packed=True)
    field(msg("Int64List"), "value", 1, 3, label=3, packed=True)
    ft = msg("Feature")
    ft.oneof_decl.add().name = "kind"
    for i, (n, t) in enumerate([("bytes_list", "BytesList"), ("float_list", "FloatList"), ("int64_list", "Int64List")]):
        field(ft, n, i + 1, 11, type_name=".tftest." + t, oneof=0)
    fs = msg("Features")
    e = fs.nested_type.add()
    e.name = "FeatureEntry"
    e.options.map_entry = True
    field(e, "key", 1, 9)
    field(e, "value", 2, 11, type_name=".tftest.Feature")
    field(fs, "feature", 1, 11, label=3, type_name=".tftest.Features.FeatureEntry")
    field(msg("Example"), "features", 1, 11, type_name=".tftest.Features")
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("tftest.Example"))


def test_example_codec_matches_the_protobuf_runtime():
    Example = _protobuf_example_class()
    img = os.urandom(300)
    cases = [
        {"image": tfrecord.bytes_feature(img), "caption": tfrecord.int64_feature([5, 300, 50257, 0, 2 ** 40])},
        {"image": tfrecord.bytes_feature(b""), "caption": tfrecord.int64_feature([])},
        {"caption": tfrecord.int64_feature([-1, -2 ** 63, 2 ** 63 - 1]), "w": tfrecord.float_feature([0.5, -3.25]),
         "image": tfrecord.bytes_feature([b"a", b"bc"])},
    ]
    for feats in cases:
        m = Example()
        for k, (kind, vals) in feats.items():
            getattr(m.features.feature[k], kind + "_list").value.extend(vals)
        ours = tfrecord.encode_example(feats)
        assert ours == m.SerializeToString(deterministic=True)
        back = Example()
        back.ParseFromString(ours)
        assert back == m
        assert tfrecord.decode_example(m.SerializeToString()) == {k: (kind, list(v)) for k, (kind, v) in feats.items()}
    # unpacked repeated int64 (older writers) must parse too: field 1, wire type 0, twice
    unpacked_feature = bytes([0x1A, 0x04, 0x08, 0x07, 0x08, 0x09])
    entry = bytes([0x0A, 0x01]) + b"c" + bytes([0x12, len(unpacked_feature)]) + unpacked_feature
    ex = bytes([0x0A, len(entry) + 2, 0x0A, len(entry)]) + entry
    assert tfrecord.decode_example(ex) == {"c": ("int64", [7, 9])}


def _write_shards(tmp_path, counts, prefix="s"):
    paths = []
    for i, n in enumerate(counts):
        p = tmp_path / f"{prefix}_{i}.tfrecords"
        with tfrecord.TFRecordWriter(p) as w:
            for j in range(n):
                w.write(f"{i}:{j}".encode())
        paths.append(str(p))
    return paths


def test_interleave_is_the_deterministic_cycle_of_four(tmp_path):
    files = _write_shards(tmp_path, [3, 1, 0, 2, 2, 4])
    got = [r.decode() for r in dp.interleave_records(files, cycle_length=4)]
    # slots start as files 0,1,3,4 (file 2 is empty and skipped); file 1 runs out after one record and file 5 takes its slot
    assert got == ["0:0", "1:0", "3:0", "4:0", "0:1", "5:0", "3:1", "4:1", "0:2", "5:1", "5:2", "5:3"]
    assert sorted(got) == sorted(f"{i}:{j}" for i, n in enumerate([3, 1, 0, 2, 2, 4]) for j in range(n))


def test_shuffle_batch_repeat_and_rank_slices(tmp_path):
    _write_shards(tmp_path, [7, 6, 9])
    pattern = str(tmp_path / "s_*.tfrecords")
    rng = np.random.default_rng(3)
    out = list(dp.shuffle_buffer(iter(range(50)), 10, rng))
    assert sorted(out) == list(range(50)) and out != list(range(50))
    assert out[0] < 10 + 1                                        # first emission comes from the first 10 (+1) inputs
    assert list(dp.batched(iter(range(7)), 3)) == [[0, 1, 2], [3, 4, 5]]   # drop_remainder
    a = dp.record_batches(pattern, 4, True, seed=11)
    b = dp.record_batches(pattern, 4, True, seed=11)
    first = [next(a) for _ in range(12)]
    assert first == [next(b) for _ in range(12)]                  # same seed, same stream (every rank sees the same)
    epoch = [r for batch in first[:5] for r in batch]             # 22 records -> 5 full batches per epoch, 2 dropped
    assert len(set(epoch)) == 20
    assert set(r for batch in first[5:10] for r in batch) <= set(x for bt in first for x in bt)
    ev = dp.record_batches(pattern, 4, False, seed=0)             # eval: file order, interleaved, no shuffle
    assert [r.decode() for r in next(ev)] == ["0:0", "1:0", "2:0", "0:1"]
    with pytest.raises(FileNotFoundError):
        next(dp.record_batches(str(tmp_path / "nothing_*.tfrecords"), 4, True, 0))
    with pytest.raises(ValueError, match="fewer than batch_size"):
        next(dp.record_batches(pattern, 64, True, 0))


def test_label_and_box_helpers_follow_the_reference():
    for cap, n in (([], 4), ([1, 2], 4), ([1, 2, 3, 4], 4), ([1, 2, 3, 4, 5, 6], 4)):
        want = odata.truncate_or_pad_label(cap, n, 50257)
        got = dp.truncate_or_pad_label(cap, n, 50257)
        assert got.dtype == np.int32 and (got == want).all()
    assert (dp.reference_crop_box(32, 32) == np.array([0, 0, 1, 1], np.float32)).all()
    for h, w in ((32, 32), (480, 640), (640, 480), (1, 7)):
        assert (dp.reference_crop_box(h, w) == odata.reference_crop_box(h, w)).all()
    # as written in the reference: y1 = (1 - W/c)/2, x1 = (1 - H/c)/2, y2 = W/c, x2 = H/c
    assert np.allclose(dp.reference_crop_box(480, 640), [0.0, 0.125, 1.0, 0.75])


def test_oracle_crop_and_resize_closed_forms():
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (9, 9, 3), dtype=np.uint8)
    same = odata.crop_and_resize_bilinear(img, [0, 0, 1, 1], 9)
    assert (same == img.astype(np.float32)).all()                 # identity grid: exact pixels
    ramp = np.tile(np.arange(0, 90, 10, dtype=np.uint8)[None, :, None], (9, 1, 1))
    half = odata.crop_and_resize_bilinear(ramp, [0, 0, 1, 1], 5)   # samples at x = 0,2,4,6,8 -> 0,20,40,60,80
    assert (half[:, :, 0] == np.array([0, 20, 40, 60, 80], np.float32)).all()
    up = odata.crop_and_resize_bilinear(ramp, [0, 0, 1, 1], 17)    # midpoints interpolate linearly
    assert np.allclose(up[0, :, 0], np.arange(17) * 5.0)
    out = odata.crop_and_resize_bilinear(img, [0.0, 0.25, 1.5, 0.75], 4)
    assert (out[-1] == 0).all() and (out[0] != 0).any()            # rows sampled beyond H-1: extrapolation value
    one = odata.crop_and_resize_bilinear(img, [0, 0, 1, 1], 1)     # size 1: the box centre
    assert (one[0, 0] == img[4, 4]).all()
    norm = odata.decode_img(img, 9)
    assert np.array_equal(norm, (img.astype(np.float32) - np.float32(127.5)) / np.float32(127.5))


class _WordTokenizer:
    def encode(self, text):
        return [len(w) for w in text.split()]


def _jpeg(rng, h, w):
    from PIL import Image
    buf = io.BytesIO()
    Image.fromarray(rng.integers(0, 256, (h, w, 3), dtype=np.uint8)).save(buf, format="JPEG", quality=90)
    return buf.getvalue()


def test_dataset_writer_shards_and_pipeline_host_stage(tmp_path):
    rng = np.random.default_rng(5)
    src = tmp_path / "imgs"
    src.mkdir()
    for i in range(7):
        (src / f"{i}.jpg").write_bytes(_jpeg(rng, 20 + i, 24))
    n = dataset_tools.create_random_dataset(str(src / "*.jpg"), tmp_path / "paired", max_images_per_folder=3,
                                            words_per_caption=5, seed=0)
    assert n == 7 and sorted(os.listdir(tmp_path / "paired")) == ["0", "1", "2", "captions_data.jsonl"]
    count, shards = dataset_tools.create_paired_dataset(tmp_path / "paired" / "captions_data.jsonl", "T",
                                                        tmp_path / "rec", examples_per_tfrecord=3,
                                                        tokenizer=_WordTokenizer())
    assert count == 7 and [os.path.basename(s) for s in shards] == ["T_0.tfrecords", "T_1.tfrecords", "T_2.tfrecords"]
    per_shard = [len(list(tfrecord.tfrecord_iterator(s))) for s in shards]
    assert per_shard == [3, 3, 1]                                   # nothing overwritten (reference defect C.7 not reproduced)
    items = load = dataset_tools.load_jsonl(tmp_path / "paired" / "captions_data.jsonl")
    ex = tfrecord.decode_example(next(tfrecord.tfrecord_iterator(shards[0])))
    assert ex["image"][1][0] == (tmp_path / "paired" / load[0]["image_path"]).read_bytes()
    assert ex["caption"] == ("int64", [len(w) for w in items[0]["caption"].split()])
    # literal reference behaviour on request: only the first character of a string caption is encoded
    _, q = dataset_tools.create_paired_dataset(tmp_path / "paired" / "captions_data.jsonl", "Q", tmp_path / "recq",
                                               examples_per_tfrecord=100, tokenizer=_WordTokenizer(),
                                               first_char_quirk=True)
    assert tfrecord.decode_example(next(tfrecord.tfrecord_iterator(q[0])))["caption"] == ("int64", [1])

    # host stage of the input pipeline: parse + JPEG decode + packing + labels (no GPU needed up to here)
    from concurrent.futures import ThreadPoolExecutor
    from PIL import Image
    params = {"n_channels": 3, "text_seq_len": 6, "padding_id": 50257, "dataset": {"image_size": 16}}
    recs = next(dp.record_batches(str(tmp_path / "rec" / "T_*.tfrecords"), 4, False, 0))
    hb = dp.host_stage(recs, params, labeled=True, tfrecords=True, pool=ThreadPoolExecutor(2))
    assert hb.labels.shape == (4, 6) and hb.labels.dtype.is_floating_point is False
    for k, rec in enumerate(recs):
        e = tfrecord.decode_example(rec)
        want = np.asarray(Image.open(io.BytesIO(e["image"][1][0])).convert("RGB"))
        h, w = int(hb.heights[k]), int(hb.widths[k])
        got = hb.packed.numpy()[int(hb.offsets[k]):int(hb.offsets[k]) + h * w * 3].reshape(h, w, 3)
        assert (h, w) == want.shape[:2] and (got == want).all()
        assert (hb.labels[k].numpy() == odata.truncate_or_pad_label(e["caption"][1], 6, 50257)).all()
        assert (hb.boxes[k].numpy() == odata.reference_crop_box(h, w)).all()


def test_input_fns_switch_between_synthetic_and_real(tmp_path):
    from dalle_mtf_b200 import input_fns
    params = {"dataset": {"train_path": "synthetic", "eval_path": str(tmp_path / "none_*.tfrecords"), "image_size": 8},
              "train_batch_size": 2, "eval_batch_size": 2, "text_seq_len": 4, "text_vocab_size": 50258, "n_channels": 3}
    assert input_fns._real_data(params, eval=False) is False
    assert input_fns._real_data(params, eval=True) is False          # warns, falls back to the synthetic stand-in
    _write_shards(tmp_path, [2], prefix="none")
    assert input_fns._real_data(params, eval=True) is True


def test_real_input_fn_streams_batches_and_stops_its_threads(tmp_path, monkeypatch):
    """Host side of real_input_fn (record stream -> thread-pool decode -> prefetch queue) with the device stage stubbed:
    batches keep coming across epochs, labels follow the records, and closing the generator stops the producer."""
    import threading
    import time
    import torch
    rng = np.random.default_rng(0)
    with tfrecord.TFRecordWriter(tmp_path / "a_0.tfrecords") as w:
        for i in range(10):
            w.write(tfrecord.encode_example({"image": tfrecord.bytes_feature(_jpeg(rng, 16, 16)),
                                             "caption": tfrecord.int64_feature([i, i + 1])}))
    seen = []
    monkeypatch.setattr(dp, "device_stage", lambda hb, params, device: (seen.append(hb), torch.zeros(hb.heights.numel(), 8, 8, 3))[1])
    params = {"dataset": {"train_path": str(tmp_path / "a_*.tfrecords"), "eval_path": str(tmp_path / "a_*.tfrecords"),
                          "image_size": 8}, "train_batch_size": 4, "eval_batch_size": 4, "text_seq_len": 3, "n_channels": 3}
    before = threading.active_count()
    it = dp.real_input_fn(params, True, True, device="cpu")
    firsts = []
    for _ in range(5):                                   # 10 examples, batch 4 -> 2 batches per epoch: crosses epochs
        img, cap = next(it)
        assert img.shape == (4, 8, 8, 3) and cap.shape == (4, 3)
        firsts.append(cap[:, 0].tolist())
        assert (cap[:, 1] == cap[:, 0] + 1).all() and (cap[:, 2] == 50257).all()
    assert firsts[0] == [0, 1, 2, 3] and firsts[1] == [4, 5, 6, 7] and firsts[2] == [0, 1, 2, 3]   # eval: file order, repeat
    assert threading.active_count() > before
    it.close()
    deadline = time.time() + 5
    while threading.active_count() > before and time.time() < deadline:
        time.sleep(0.05)
    assert threading.active_count() == before
    # rank 1 of 2 decodes only its half of every global batch
    monkeypatch.setenv("RANK", "1")
    monkeypatch.setenv("WORLD_SIZE", "2")
    it = dp.real_input_fn(params, True, True, device="cpu")
    _, cap = next(it)
    assert cap[:, 0].tolist() == [2, 3]
    it.close()


def test_example_codec_property_roundtrip_against_protobuf():
    """Randomised feature maps: protobuf parses our bytes to the same message, we parse protobuf's bytes, and the bytes
    are identical whenever the map order is unambiguous (the runtime orders keys of different lengths its own way)."""
    from hypothesis import given, settings, strategies as st
    Example = _protobuf_example_class()
    i64 = st.integers(min_value=-2 ** 63, max_value=2 ** 63 - 1)
    f32 = st.floats(width=32, allow_nan=False, allow_infinity=False)
    feature = st.one_of(st.lists(st.binary(max_size=40), max_size=4).map(lambda v: ("bytes", v)),
                        st.lists(i64, max_size=12).map(lambda v: ("int64", v)),
                        st.lists(f32, max_size=6).map(lambda v: ("float", v)))
    names = st.text(alphabet="abcdefghij_/0123", min_size=1, max_size=8)

    @settings(max_examples=150, deadline=None)
    @given(st.dictionaries(names, feature, max_size=5))
    def check(feats):
        m = Example()
        m.features.SetInParent()                   # tf.train.Example(features=tf.train.Features(...)): always present
        for k, (kind, vals) in feats.items():
            lst = getattr(m.features.feature[k], kind + "_list")
            lst.SetInParent()                      # an empty list still selects the oneof member
            lst.value.extend(vals)
        ours = tfrecord.encode_example(feats)
        parsed = Example()
        parsed.ParseFromString(ours)
        assert parsed == m                         # map-entry order carries no meaning on the wire ...
        if len({len(k) for k in feats}) <= 1:      # ... and is only comparable byte-for-byte for same-length keys
            assert ours == m.SerializeToString(deterministic=True)
        assert tfrecord.decode_example(m.SerializeToString()) == {k: (kind, list(v)) for k, (kind, v) in feats.items()}
        back = tfrecord.decode_example(ours)
        assert back == {k: (kind, list(v)) for k, (kind, v) in feats.items()}
        assert tfrecord.frame_record(ours) == odata.tfrecord_frame(ours)

    check()
