"""Throughput of the real input pipeline (N2) on one GPU box: writes a synthetic TFRecord dataset of 256x256 JPEGs,
then times (a) the host stage alone (TFRecord parse + JPEG decode + packing), (b) the device stage alone (H2D of the
uint8 pixels + db200_image_crop_resize_normalize, CUDA events), (c) the whole dalle_input_fn stream.
Prints one JSON line."""
import io
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dalle_mtf_b200 import data_pipeline as dp  # noqa: E402
from dalle_mtf_b200 import tfrecord  # noqa: E402


def main(n_images=512, size=256, batch=32):
    from PIL import Image
    from concurrent.futures import ThreadPoolExecutor
    rng = np.random.default_rng(0)
    tmp = tempfile.mkdtemp(prefix="db200_pipe_")
    # smooth images (random low-frequency fields) so the JPEGs have a realistic ~20-30 KB size
    per = n_images // 4
    jpeg_bytes = 0
    for s in range(4):
        with tfrecord.TFRecordWriter(os.path.join(tmp, f"P_{s}.tfrecords")) as w:
            for i in range(per):
                small = rng.integers(0, 256, (16, 16, 3), dtype=np.uint8)
                im = Image.fromarray(small).resize((size, size), Image.BICUBIC)
                buf = io.BytesIO()
                im.save(buf, format="JPEG", quality=90)
                jpeg_bytes += buf.tell()
                cap = rng.integers(0, 50257, int(rng.integers(5, 64))).tolist()
                w.write(tfrecord.encode_example({"image": tfrecord.bytes_feature(buf.getvalue()),
                                                 "caption": tfrecord.int64_feature(cap)}))
    params = {"dataset": {"train_path": os.path.join(tmp, "P_*.tfrecords"), "eval_path": "synthetic",
                          "image_size": size, "tfrecords": True},
              "train_batch_size": batch, "eval_batch_size": batch, "text_seq_len": 256, "padding_id": 50257,
              "text_vocab_size": 50258, "n_channels": 3}
    workers = min(16, len(os.sched_getaffinity(0)))
    pool = ThreadPoolExecutor(workers)
    rb = dp.record_batches(params["dataset"]["train_path"], batch, True, 0)
    for _ in range(2):  # warm-up: thread pool, PIL plugins, the native library
        dp.host_stage(next(rb), params, True, True, pool)
    t0 = time.perf_counter()
    hbs = [dp.host_stage(next(rb), params, True, True, pool) for _ in range(12)]
    host_s = (time.perf_counter() - t0) / 12
    dev = torch.device("cuda", 0)
    for hb in hbs[:3]:
        dp.device_stage(hb, params, dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for hb in hbs:
        out = dp.device_stage(hb, params, dev)
    e1.record()
    torch.cuda.synchronize()
    dev_ms = e0.elapsed_time(e1) / len(hbs)
    from dalle_mtf_b200.input_fns import dalle_input_fn
    it = iter(dalle_input_fn(params))
    for _ in range(3):
        next(it)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        img, cap = next(it)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / 20
    out_bytes = batch * size * size * 3 * 4
    in_bytes = batch * size * size * 3
    print(json.dumps({"metric": "input_pipeline_imgs_per_sec", "batch": batch, "image_size": size,
                      "host_threads": workers, "avg_jpeg_bytes": jpeg_bytes / (4 * per),
                      "host_stage_imgs_per_s": batch / host_s, "device_stage_ms": dev_ms,
                      "device_stage_GBps": (out_bytes + in_bytes) / (dev_ms * 1e-3) / 1e9,
                      "stream_imgs_per_s": batch / e2e_s}))


if __name__ == "__main__":
    main()
