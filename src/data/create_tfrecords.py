"""src/data/create_tfrecords.py of the reference, served by the TensorFlow-free writer (dalle_mtf_b200.dataset_tools)."""
from dalle_mtf_b200.dataset_tools import (create_paired_dataset, create_random_dataset, dump_jsonl,  # noqa: F401
                                          load_jsonl, serialize_example)
