"""The reference's tf.data input pipelines without TensorFlow ("next" row N2): src/input_fns.py:4-120.

What the reference builds                                        here
  tf.io.gfile.glob(path)                                          sorted(glob.glob(local_path(path), recursive=True))
  Dataset.from_tensor_slices(files).shuffle(n, reshuffle=False)   one seeded permutation of the file list (train only)
  parallel_interleave(TFRecordDataset, cycle_length=4,            deterministic round-robin over 4 open files, one
                      sloppy=False)                                record per turn, exhausted files replaced in order
  map(parse_single_example + decode_img + truncate_or_pad)        parse (tfrecord.decode_example), JPEG decode on a
                                                                  host thread pool (PIL / libjpeg-turbo), then ONE CUDA
                                                                  kernel per batch: centre-crop box + bilinear
                                                                  crop_and_resize + (x - 127.5) / 127.5
  shuffle(batch_size * 5)  [train]                                the same streaming shuffle buffer, seeded
  batch(batch_size, drop_remainder=True).prefetch().repeat()      same; a background thread keeps 2 batches in flight

Ordering, shuffling and batching operate on the SERIALISED records (cheap), so every data-parallel rank runs the same
record stream and decodes only its own rows [r*B/N, (r+1)*B/N) of each global batch (the reference decodes the full
batch on every core: BROADCAST input, train_dalle.py:69).

Decoded images go to the GPU as uint8 (3 bytes per source pixel instead of 12 per output pixel) and the crop / resize /
normalise kernel (csrc/data_ops.cu) writes the float32 NHWC batch the model functions consume; there is no CPU
implementation of that step in the product (oracle/data.py holds the checker).
"""
import glob
import io
import os
import queue
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import tfrecord
from .utils import local_path

CYCLE_LENGTH = 4  # parallel_interleave(..., cycle_length=4)                              input_fns.py:80,116
SHUFFLE_BATCHES = 5  # ds.shuffle(buffer_size=params["batch_size"] * 5)                   input_fns.py:26


def list_files(pattern):
    """tf.io.gfile.glob: sorted matches of a (possibly gs://) glob; ** recurses."""
    return sorted(glob.glob(local_path(pattern), recursive=True))


def reference_crop_box(height, width):
    """crop_center_and_resize's box, exactly as written (input_fns.py:4-13): s = shape(img) = (H, W, C) but the code
    names w, h = s[0], s[1]; c = max(w, h); wn, hn = h / c, w / c (float64 true division);
    box = [(1 - wn) / 2, (1 - hn) / 2, wn, hn] interpreted by crop_and_resize as [y1, x1, y2, x2] and cast to float32.
    For square images this is [0, 0, 1, 1] (a plain resize); for others it is NOT a centred crop (y2/x2 are extents,
    not end points, and the axes are swapped) — reproduced as is, the reference's datasets are square."""
    w, h = float(height), float(width)
    c = max(w, h)
    wn, hn = h / c, w / c
    return np.array([(1 - wn) / 2, (1 - hn) / 2, wn, hn], dtype=np.float64).astype(np.float32)


def truncate_or_pad_label(label, text_seq_len, padding_id):
    """input_fns.py:32-38: pad with text_seq_len padding ids, keep the first text_seq_len -> int32 [text_seq_len]."""
    out = np.full((text_seq_len,), padding_id, dtype=np.int32)
    n = min(len(label), text_seq_len)
    out[:n] = np.asarray(label[:n], dtype=np.int64).astype(np.int32)
    return out


def decode_jpeg(data, channels=3):
    """tf.image.decode_jpeg(img, channels): uint8 [H, W, channels] (libjpeg, default DCT, fancy upsampling)."""
    from PIL import Image
    im = Image.open(io.BytesIO(data))
    im = im.convert("RGB" if channels == 3 else "L")
    a = np.asarray(im, dtype=np.uint8)
    return a if a.ndim == 3 else a[:, :, None]


# ------------------------------------------------------------------------------------------------ record streams
def interleave_records(files, cycle_length=CYCLE_LENGTH):
    """Deterministic parallel_interleave(TFRecordDataset, cycle_length, block_length=1, sloppy=False): one record per
    turn from each of `cycle_length` open files; an exhausted file's slot is taken over by the next unopened file."""
    pending = iter(files)

    def open_next():
        for f in pending:
            it = tfrecord.tfrecord_iterator(f)
            first = next(it, None)
            if first is not None:  # empty files contribute nothing
                return [first, it]
        return None

    slots = [open_next() for _ in range(cycle_length)]
    while any(s is not None for s in slots):
        for i, s in enumerate(slots):
            if s is None:
                continue
            yield s[0]
            s[0] = next(s[1], None)
            if s[0] is None:
                slots[i] = open_next()


def shuffle_buffer(items, buffer_size, rng):
    """tf.data shuffle: fill a buffer, then emit a uniformly chosen slot and refill it from the input."""
    buf = []
    for x in items:
        if len(buf) < buffer_size:
            buf.append(x)
            continue
        j = int(rng.integers(len(buf)))
        yield buf[j]
        buf[j] = x
    while buf:
        j = int(rng.integers(len(buf)))
        yield buf[j]
        buf[j] = buf[-1]
        buf.pop()


def batched(items, batch_size):
    """batch(batch_size, drop_remainder=True)."""
    cur = []
    for x in items:
        cur.append(x)
        if len(cur) == batch_size:
            yield cur
            cur = []


def record_batches(pattern, batch_size, train, seed, tfrecords=True):
    """Global batches of raw items, forever (.repeat() after .batch(), as in the reference).
    tfrecords=True: items are serialised tf.train.Example records; False: items are image file paths (the
    list_files branch of vae_input_fn, input_fns.py:87-104)."""
    files = list_files(pattern)
    if not files:
        raise FileNotFoundError(f"no input files match {pattern!r} (resolved to {local_path(pattern)!r})")
    rng = np.random.default_rng(seed)
    if train:
        files = [files[i] for i in rng.permutation(len(files))]  # reshuffle_each_iteration=False: once
    while True:
        items = interleave_records(files) if tfrecords else iter(files)
        if train:
            items = shuffle_buffer(items, batch_size * SHUFFLE_BATCHES, rng)
        n = 0
        for b in batched(items, batch_size):
            n += 1
            yield b
        if n == 0:
            raise ValueError(f"{pattern!r} holds fewer than batch_size={batch_size} examples (drop_remainder=True)")


# ------------------------------------------------------------------------------------------------ decode + GPU stage
class HostBatch:
    """One rank-local batch after the host stage: uint8 pixels packed back to back + per-image geometry (+ labels)."""
    __slots__ = ("packed", "offsets", "heights", "widths", "boxes", "labels")


def _parse(item, labeled, tfrecords, channels):
    if tfrecords:
        ex = tfrecord.decode_example(item)
        kind, vals = ex["image"]
        if kind != "bytes" or len(vals) != 1:
            raise ValueError('feature "image" must be a single bytes value')
        img = decode_jpeg(vals[0], channels)
        cap = ex["caption"][1] if labeled else None
        if labeled and ex["caption"][0] != "int64":
            raise ValueError('feature "caption" must be an int64 list')
        return img, cap
    with open(item, "rb") as f:
        return decode_jpeg(f.read(), channels), None


def host_stage(items, params, labeled, tfrecords, pool):
    channels = params.get("n_channels") or 3
    decoded = list(pool.map(lambda it: _parse(it, labeled, tfrecords, channels), items))
    hb = HostBatch()
    sizes = [im.size for im, _ in decoded]
    offs = np.zeros(len(decoded), dtype=np.int64)
    offs[1:] = np.cumsum(sizes)[:-1]
    total = int(sum(sizes))
    packed = torch.empty(total, dtype=torch.uint8)
    if torch.cuda.is_available():
        packed = packed.pin_memory()
    pk = packed.numpy()
    for (im, _), o in zip(decoded, offs):
        pk[o:o + im.size] = im.reshape(-1)
    hb.packed = packed
    hb.offsets = torch.from_numpy(offs)
    hb.heights = torch.tensor([im.shape[0] for im, _ in decoded], dtype=torch.int32)
    hb.widths = torch.tensor([im.shape[1] for im, _ in decoded], dtype=torch.int32)
    hb.boxes = torch.from_numpy(np.stack([reference_crop_box(im.shape[0], im.shape[1]) for im, _ in decoded]))
    if labeled:
        pad = params.get("padding_id")
        pad = 50257 if pad is None else pad
        hb.labels = torch.from_numpy(np.stack([truncate_or_pad_label(c, params["text_seq_len"], pad)
                                               for _, c in decoded]))
        if torch.cuda.is_available():
            hb.labels = hb.labels.pin_memory()
    else:
        hb.labels = None
    return hb


def device_stage(hb, params, device):
    """uint8 batch -> float32 NHWC [B, size, size, C] on `device` through db200_image_crop_resize_normalize."""
    from . import ops
    size = params["dataset"]["image_size"]
    channels = params.get("n_channels") or 3
    B = hb.heights.numel()
    out = torch.empty(B, size, size, channels, dtype=torch.float32, device=device)
    ops.image_crop_resize_normalize(hb.packed.to(device, non_blocking=True), hb.offsets.to(device, non_blocking=True),
                                    hb.heights.to(device, non_blocking=True), hb.widths.to(device, non_blocking=True),
                                    hb.boxes.to(device, non_blocking=True), out, channels, size)
    return out


def real_input_fn(params, eval, labeled, seed=1234, device=None, workers=None, prefetch=2):
    """Generator of (image, label) / (image, image) batches for this rank; images are float32 NHWC on `device`."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    gb = params["eval_batch_size" if eval else "train_batch_size"]
    if gb % world != 0:
        raise ValueError(f"batch size {gb} not divisible by the data-parallel size {world}")
    lb = gb // world
    ds = params["dataset"]
    tfrecords = labeled or bool(ds.get("tfrecords"))
    pattern = ds["eval_path" if eval else "train_path"]
    if device is None:
        device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    pool = ThreadPoolExecutor(max_workers=workers or min(16, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity")
                                                          else (os.cpu_count() or 4)))
    q = queue.Queue(maxsize=prefetch)
    stop = threading.Event()

    def put(x):
        while not stop.is_set():
            try:
                q.put(x, timeout=0.2)
                return True
            except queue.Full:
                continue
        return False

    def produce():
        try:
            for items in record_batches(pattern, gb, not eval, seed, tfrecords):
                if not put(host_stage(items[rank * lb:(rank + 1) * lb], params, labeled, tfrecords, pool)):
                    return
        except BaseException as e:  # surfaced in the consumer
            put(e)

    t = threading.Thread(target=produce, daemon=True, name="db200-input")
    t.start()
    try:
        while True:
            hb = q.get()
            if isinstance(hb, BaseException):
                raise hb
            img = device_stage(hb, params, device)
            yield (img, hb.labels) if labeled else (img, img)
    finally:  # generator closed / garbage-collected: stop the producer and release its threads
        stop.set()
        pool.shutdown(wait=False, cancel_futures=True)
