"""Dataset writer of the reference ("next" row N3): src/data/create_tfrecords.py, without TensorFlow / OpenCV.

Same functions, arguments and on-disk format (TFRecord files of tf.train.Example{image: bytes (JPEG), caption: int64
token ids}).  Two defects of the reference are NOT reproduced and one is kept behind a flag:
  * shard numbering: the reference opens `{name}_0` twice (the counter is bumped after the second open), so the first
    `examples_per_tfrecord` examples are overwritten; here shards are numbered 0, 1, 2, ... and nothing is lost;
  * `tokenizer.encode(item["caption"][0])` encodes only the FIRST CHARACTER when the caption is a string (it is written
    for COCO-style lists of captions).  Here a list contributes its first caption and a string is encoded whole;
    `first_char_quirk=True` restores the literal behaviour.
"""
import glob
import io
import json
import os
import random
import shutil
from pathlib import Path, PurePath

from .tfrecord import TFRecordWriter, bytes_feature, encode_example, int64_feature


def dump_jsonl(data, output_path, append=False):
    """create_tfrecords.py:15-23."""
    with open(output_path, "a+" if append else "w", encoding="utf-8") as f:
        for line in data:
            f.write(json.dumps(line, ensure_ascii=False) + "\n")


def load_jsonl(input_path):
    """create_tfrecords.py:26-34."""
    with open(input_path, "r", encoding="utf-8") as f:
        return [json.loads(line.rstrip("\n|\r")) for line in f]


def serialize_example(image, caption):
    """create_tfrecords.py:50-56: {'image': bytes_list[image], 'caption': int64_list(caption)}."""
    return encode_example({"image": bytes_feature(image), "caption": int64_feature(caption)})


def create_random_dataset(path_to_images, out_dir, max_images_per_folder=1000, words_per_caption=50, words=None,
                          seed=None):
    """create_tfrecords.py:59-97: copies images into numbered sub-folders and writes captions_data.jsonl with random
    captions.  The reference downloads a word list; there is no network here, so `words` (a list of str) may be given
    and defaults to a small built-in list."""
    words = words or ["a", "photo", "of", "the", "red", "blue", "green", "small", "large", "cat", "dog", "bird", "car",
                      "tree", "house", "on", "in", "with", "near", "sky", "water", "street", "table", "person"]
    rnd = random.Random(seed)
    out_dir = Path(out_dir)
    jsonl_path = out_dir / "captions_data.jsonl"
    os.makedirs(out_dir, exist_ok=True)
    images = sorted(glob.glob(str(path_to_images), recursive=True))
    folder_count = 0
    sub_folder = None
    for i, image in enumerate(images):
        if i % max_images_per_folder == 0:
            sub_folder = out_dir / str(folder_count)
            os.makedirs(sub_folder, exist_ok=True)
            folder_count += 1
        image = Path(image)
        data = {"caption": " ".join(rnd.choice(words) for _ in range(words_per_caption)),
                "image_path": str(sub_folder.relative_to(out_dir) / image.name)}
        shutil.copy(image, sub_folder)
        dump_jsonl([data], jsonl_path, append=True)
    return len(images)


def _reencode_jpeg(path, quality=94):
    """cv2.imencode('.jpg', cv2.imread(path), quality 94) stand-in (PIL / libjpeg)."""
    from PIL import Image
    buf = io.BytesIO()
    Image.open(path).convert("RGB").save(buf, format="JPEG", quality=quality)
    return buf.getvalue()


def create_paired_dataset(path_to_jsonl, name, out_dir, examples_per_tfrecord=1000, tokenizer=None, reencode=False,
                          first_char_quirk=False):
    """create_tfrecords.py:100-178.  Returns (examples written, shard paths)."""
    if tokenizer is None:
        from .tokenizer import get_tokenizer
        tokenizer = get_tokenizer()
    out_dir = Path(out_dir)
    os.makedirs(out_dir, exist_ok=True)
    if isinstance(path_to_jsonl, (PurePath, str)):
        path_to_jsonl = [path_to_jsonl]
    if not isinstance(path_to_jsonl, list):
        raise TypeError("path_to_jsonl type not recognized, should be str, path, or list")
    shards = []
    writer = None
    example_count = 0
    try:
        for path in path_to_jsonl:
            path = Path(path)
            for item in load_jsonl(path):
                if example_count % examples_per_tfrecord == 0:
                    if writer is not None:
                        writer.close()
                    shards.append(str(out_dir / f"{name}_{len(shards)}.tfrecords"))
                    writer = TFRecordWriter(shards[-1])
                image_path = path.parent / item["image_path"]
                img = _reencode_jpeg(image_path) if reencode else open(image_path, "rb").read()
                cap = item["caption"]
                text = cap[0] if (first_char_quirk or isinstance(cap, (list, tuple))) else cap
                writer.write(serialize_example(img, tokenizer.encode(text)))
                example_count += 1
    finally:
        if writer is not None:
            writer.close()
    return example_count, shards
