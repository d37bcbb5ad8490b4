"""GPU parity of the input-pipeline kernel ("next" row N2): db200_image_crop_resize_normalize vs oracle/data.py
(bit-exact: both execute the same sequence of individually rounded float32 operations), and the full TFRecord -> batch
path through dalle_input_fn / vae_input_fn."""
import io
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from dalle_mtf_b200 import data_pipeline as dp  # noqa: E402
from dalle_mtf_b200 import dataset_tools, ops  # noqa: E402
from oracle import data as odata  # noqa: E402

DEV = "cuda"


def _run_kernel(images, boxes, size, channels):
    offs, cur = [], 0
    for im in images:
        offs.append(cur)
        cur += im.size
    packed = torch.from_numpy(np.concatenate([im.reshape(-1) for im in images])).to(DEV)
    out = torch.empty(len(images), size, size, channels, device=DEV)
    ops.image_crop_resize_normalize(packed, torch.tensor(offs, dtype=torch.int64, device=DEV),
                                    torch.tensor([im.shape[0] for im in images], dtype=torch.int32, device=DEV),
                                    torch.tensor([im.shape[1] for im in images], dtype=torch.int32, device=DEV),
                                    torch.from_numpy(np.stack(boxes)).to(DEV), out, channels, size)
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("channels", [3, 1])
def test_crop_resize_kernel_is_bit_exact_against_the_oracle(channels):
    rng = np.random.default_rng(7 + channels)
    shapes = [(32, 32), (17, 40), (40, 17), (1, 1), (2, 9), (64, 48), (33, 33)]
    images = [rng.integers(0, 256, (h, w, channels), dtype=np.uint8) for h, w in shapes]
    for size in (1, 8, 32, 45):
        boxes = [odata.reference_crop_box(h, w) for h, w in shapes]
        got = _run_kernel(images, boxes, size, channels)
        for k, im in enumerate(images):
            want = odata.decode_img(im, size)
            assert np.array_equal(got[k], want), (size, shapes[k])
    # arbitrary boxes, including ones that leave the image (extrapolation -> 0 -> -1 after normalisation)
    boxes = [np.array(b, np.float32) for b in ([0.1, 0.2, 0.9, 0.7], [-0.2, 0.0, 1.2, 1.0], [0.5, 0.5, 0.5, 0.5],
                                               [0, 0, 1, 1], [0.9, 0.1, 0.1, 0.9], [0.0, 0.3, 2.0, 0.31], [0, 0, 1, 1])]
    got = _run_kernel(images, boxes, 16, channels)
    for k, im in enumerate(images):
        want = (odata.crop_and_resize_bilinear(im, boxes[k], 16) - np.float32(127.5)) / np.float32(127.5)
        assert np.array_equal(got[k], want.astype(np.float32)), k


def test_identity_resize_at_full_size_is_the_reference_normalisation():
    """BASELINE image size (256x256 -> 256x256): crop_and_resize on the identity grid returns the pixels themselves."""
    g = torch.Generator().manual_seed(0)
    u8 = torch.randint(0, 256, (16, 256, 256, 3), generator=g, dtype=torch.uint8)
    images = [u8[i].numpy() for i in range(16)]
    got = _run_kernel(images, [odata.reference_crop_box(256, 256)] * 16, 256, 3)
    want = (u8.to(torch.float32) - 127.5) / 127.5
    assert np.array_equal(got, want.numpy())
    down = _run_kernel(images, [odata.reference_crop_box(256, 256)] * 16, 128, 3)   # different size: still in [-1, 1]
    assert down.min() >= -1.0 and down.max() <= 1.0


def _jpeg(arr, quality=92):
    from PIL import Image
    buf = io.BytesIO()
    Image.fromarray(arr).save(buf, format="JPEG", quality=quality)
    return buf.getvalue()


class _Tok:
    def encode(self, text):
        return [ord(c) for c in text][:300]


def test_tfrecords_to_device_batches_through_the_input_fns(tmp_path):
    from PIL import Image
    from dalle_mtf_b200.input_fns import dalle_input_fn, vae_input_fn
    from dalle_mtf_b200 import tfrecord
    rng = np.random.default_rng(2)
    (tmp_path / "imgs").mkdir()
    lines = []
    for i in range(24):
        h, w = (48, 48) if i % 3 else (40, 56)
        (tmp_path / "imgs" / f"{i}.jpg").write_bytes(_jpeg(rng.integers(0, 256, (h, w, 3), dtype=np.uint8)))
        lines.append({"image_path": f"imgs/{i}.jpg", "caption": "caption number %d" % i})
    dataset_tools.dump_jsonl(lines, tmp_path / "c.jsonl")
    n, shards = dataset_tools.create_paired_dataset(tmp_path / "c.jsonl", "D", tmp_path / "rec",
                                                    examples_per_tfrecord=10, tokenizer=_Tok())
    assert n == 24 and len(shards) == 3
    params = {"dataset": {"train_path": str(tmp_path / "rec" / "D_*.tfrecords"),
                          "eval_path": str(tmp_path / "rec" / "D_*.tfrecords"), "image_size": 32, "tfrecords": True},
              "train_batch_size": 8, "eval_batch_size": 8, "text_seq_len": 12, "padding_id": 50257,
              "text_vocab_size": 50258, "n_channels": 3}
    it = iter(dalle_input_fn(params, eval=True))
    img, cap = next(it)
    assert img.is_cuda and img.shape == (8, 32, 32, 3) and img.dtype == torch.float32
    assert cap.shape == (8, 12) and cap.dtype == torch.int32
    # eval order is deterministic: interleave of the three shards
    recs = next(dp.record_batches(params["dataset"]["eval_path"], 8, False, 0))
    for k, rec in enumerate(recs):
        ex = tfrecord.decode_example(rec)
        arr = np.asarray(Image.open(io.BytesIO(ex["image"][1][0])).convert("RGB"))
        assert np.array_equal(img[k].cpu().numpy(), odata.decode_img(arr, 32))
        assert (cap[k].numpy() == odata.truncate_or_pad_label(ex["caption"][1], 12, 50257)).all()
    # training stream: shuffled, repeats forever, every batch full
    tr = iter(dalle_input_fn(params, eval=False))
    seen = [next(tr)[0].shape for _ in range(7)]                    # 24 examples -> 3 batches per epoch: crosses epochs
    assert all(s == (8, 32, 32, 3) for s in seen)
    a, b = next(iter(vae_input_fn(params, eval=True)))
    assert a is b and torch.equal(a, img)                           # (image, image), same deterministic first batch
