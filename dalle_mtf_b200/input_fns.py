"""Input contract of src/input_fns.py.

`dataset.train_path` / `eval_path` == "synthetic" (the *_b200 configs), or a pattern that matches no local file (the
reference configs name gs:// buckets that cannot be reached here): synthetic data, generated below.  A pattern that
matches files: the real pipeline of data_pipeline.py (TFRecord / JPEG files -> host decode -> CUDA crop + resize +
normalise), which yields the image batch already on the GPU.

What the step functions rely on is the OUTPUT contract, reproduced exactly in both cases:
  * images: float32 NHWC [B, size, size, n_channels], (uint8 - 127.5) / 127.5               (input_fns.py:15-21)
  * captions: int32 [B, text_seq_len], truncated / right-padded with params["padding_id"]   (input_fns.py:32-38)
  * vae_input_fn yields (image, image); dalle_input_fn yields (image, caption)              (input_fns.py:64,41-52)
Each data-parallel rank generates only ITS shard of the global batch (seed 1234 + rank); the reference instead
broadcasts the full batch to every core (train_dalle.py:69).
Batches are produced in pinned host memory so the step's host->device copy is asynchronous.
"""
import os

import torch

PADDING_ID = 50257  # <|padding|> appended to the GPT-2 vocabulary (src/data/tokenizer_utils.py:7-14, train_dalle.py:49)


def _local_batch(params, eval):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    gb = params["eval_batch_size" if eval else "train_batch_size"]
    if gb % world != 0:
        raise ValueError(f"batch size {gb} not divisible by the data-parallel size {world}")
    return gb // world


def _pin(t):
    return t.pin_memory() if torch.cuda.is_available() else t


def synthetic_images(batch, size, channels, generator):
    """CIFAR-10-shaped stand-in: uint8 uniform[0,255] then the reference normalisation (input_fns.py:20)."""
    u8 = torch.randint(0, 256, (batch, size, size, channels), generator=generator, dtype=torch.uint8)
    return (u8.to(torch.float32) - 127.5) / 127.5


def synthetic_captions(batch, text_seq_len, padding_id, generator, vocab=50257, min_len=5, max_len=64):
    """Random-caption stand-in for create_random_dataset (src/data/create_tfrecords.py:59-97) + truncate_or_pad_label."""
    ids = torch.full((batch, text_seq_len), padding_id, dtype=torch.int32)
    lens = torch.randint(min_len, max_len + 1, (batch,), generator=generator)
    for b in range(batch):
        n = min(int(lens[b]), text_seq_len)
        ids[b, :n] = torch.randint(0, vocab, (n,), generator=generator, dtype=torch.int32)
    return ids


def _seed(params, eval):
    """1234 + rank for training data, a disjoint stream for eval; `_data_seed_offset` (set by the
    Estimator to the step a run resumes from) keeps a resumed run from replaying the batches it already saw."""
    rank = int(os.environ.get("RANK", "0"))
    return 1234 + rank + (10_000 if eval else 0) + 7919 * int(params.get("_data_seed_offset") or 0)


def _real_data(params, eval):
    """True when the configured path matches local files; warns (once per path) when a non-synthetic path matches none."""
    path = params["dataset"].get("eval_path" if eval else "train_path")
    if not path or path == "synthetic":
        return False
    from .data_pipeline import list_files
    if list_files(path):
        return True
    if path not in _WARNED:
        _WARNED.add(path)
        import logging
        logging.getLogger("dalle_b200").warning("no local files match %r: using SYNTHETIC data of the same shape", path)
    return False


_WARNED = set()


def vae_input_fn(params, eval=False):
    """src/input_fns.py:69-104."""
    if _real_data(params, eval):
        from .data_pipeline import real_input_fn
        yield from real_input_fn(params, eval, labeled=False, seed=1234 + 7919 * int(params.get("_data_seed_offset") or 0))
        return
    g = torch.Generator().manual_seed(_seed(params, eval))
    B = _local_batch(params, eval)
    size = params["dataset"]["image_size"]
    ch = params.get("n_channels") or 3
    while True:
        img = _pin(synthetic_images(B, size, ch, g))
        yield img, img


def dalle_input_fn(params, eval=False):
    """src/input_fns.py:106-120."""
    if _real_data(params, eval):
        from .data_pipeline import real_input_fn
        yield from real_input_fn(params, eval, labeled=True, seed=1234 + 7919 * int(params.get("_data_seed_offset") or 0))
        return
    g = torch.Generator().manual_seed(_seed(params, eval))
    B = _local_batch(params, eval)
    size = params["dataset"]["image_size"]
    ch = params.get("n_channels") or 3
    pad = params.get("padding_id")
    pad = PADDING_ID if pad is None else pad
    while True:
        img = _pin(synthetic_images(B, size, ch, g))
        cap = _pin(synthetic_captions(B, params["text_seq_len"], pad, g, vocab=min(50257, params["text_vocab_size"])))
        yield img, cap
