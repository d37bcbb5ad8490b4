"""Host-side mirror of src/utils/utils.py (config loading, logging, --new handling) plus checkpoint helpers.

Only behaviour that the hot path's callers rely on is kept: JSON -> defaultdict(None) semantics, `--model` name or
path resolution, the y/n prompt before deleting a model dir, parameter-count printing, `logs/{config}.log`.
The TPU-only plumbing (simd_mesh_setup, host_call summaries) is out of scope.
"""
import glob
import json
import logging
import os
import re
import sys
from collections import defaultdict
from shutil import rmtree
from urllib.parse import urlparse

import torch

REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fetch_model_params(model):
    """src/utils/utils.py:13-17: `model` is a config name (-> ./configs/{name}.json) or a path to a .json file.
    Missing keys read as None (defaultdict), `.get(k, default)` still returns `default`."""
    model_path = model if model.endswith(".json") else f"./configs/{model}.json"
    if not os.path.exists(model_path) and not os.path.isabs(model_path):
        alt = os.path.join(REPO_ROOT, model_path)
        if os.path.exists(alt):
            model_path = alt
    with open(model_path) as f:
        params = json.load(f)
    return defaultdict(lambda: None, params)


def yes_or_no(question):
    while True:
        reply = str(input(question + " (y/n): ")).lower().strip()
        if reply[:1] == "y":
            return True
        if reply[:1] == "n":
            return False


def local_path(path):
    """The reference configs point at gs:// buckets.  There is no GCS here: gs://bucket/x maps to
    $DB200_GS_ROOT/bucket/x (default ./gs_local)."""
    if path is None:
        return None
    u = urlparse(path)
    if u.scheme == "gs":
        root = os.environ.get("DB200_GS_ROOT", os.path.join(REPO_ROOT, "gs_local"))
        return os.path.join(root, u.netloc, u.path.lstrip("/"))
    return path


def remove_gs_or_filepath(path):
    p = local_path(path)
    if os.path.isdir(p):
        rmtree(p)


def maybe_remove_gs_or_filepath(path):
    """src/utils/utils.py:48-52 (interactive confirmation before starting afresh)."""
    if yes_or_no(f"Are you sure you want to remove '{path}' to start afresh?"):
        remove_gs_or_filepath(path)
    else:
        sys.exit()


def setup_logging(args, logdir="logs"):
    """src/utils/utils.py:184-195: logs/{config name}.log + stdout."""
    os.makedirs(logdir, exist_ok=True)
    name = os.path.splitext(os.path.basename(args.model))[0]
    logger = logging.getLogger("dalle_b200")
    logger.setLevel(logging.INFO)
    logger.propagate = False
    logger.handlers = [logging.FileHandler(f"{logdir}/{name}.log"), logging.StreamHandler(sys.stdout)]
    return logger


def print_n_params(n):
    """src/utils/utils.py:55-70."""
    print(f"\n\nN PARAMS:\n{n:,}\n\n")


def parse_mesh(mesh_shape, layout):
    """`mesh_shape` "data:16,model:2" + `layout` "batch_dim:data" (src/model_fns.py:81-82).  Only data parallelism
    exists in the reference (nothing maps to `model`); any other layout rule is rejected."""
    mesh = {}
    for part in (mesh_shape or "data:1").replace(" ", "").split(","):
        if not part:
            continue
        name, size = part.split(":")
        mesh[name] = int(size)
    rules = {}
    for part in (layout or "").replace(" ", "").split(","):
        if not part:
            continue
        dim, axis = part.split(":")
        rules[dim] = axis
    for dim, axis in rules.items():
        if dim != "batch_dim" or axis != "data":
            raise ValueError(f"layout rule '{dim}:{axis}' needs tensor/model parallelism, which the reference never "
                             "exercises and this engine does not implement (only 'batch_dim:data')")
    return mesh, rules


# ---------------------------------------------------------------------------------------------- checkpoints (N1)
_CKPT_RE = re.compile(r"model\.ckpt-(\d+)\.pt$")


def list_checkpoints(model_dir):
    d = local_path(model_dir)
    out = []
    for f in glob.glob(os.path.join(d, "model.ckpt-*.pt")):
        m = _CKPT_RE.search(f)
        if m:
            out.append((int(m.group(1)), f))
    return sorted(out)


def latest_checkpoint(model_dir):
    """tf.train.latest_checkpoint equivalent (src/model_fns.py:39)."""
    c = list_checkpoints(model_dir)
    return c[-1][1] if c else None


def load_global_step_from_checkpoint_dir(model_dir):
    """estimator_lib._load_global_step_from_checkpoint_dir (train_dalle.py:39): 0 when there is no checkpoint."""
    c = list_checkpoints(model_dir)
    return c[-1][0] if c else 0


def save_checkpoint(model_dir, step, state, max_to_keep=5):
    """`state` is a dict of reference-named tensors (+ 'global_step'); format is ours (torch.save)."""
    d = local_path(model_dir)
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, f"model.ckpt-{int(step)}.pt")
    tmp = path + ".tmp"
    torch.save(state, tmp)
    os.replace(tmp, path)
    ckpts = list_checkpoints(model_dir)
    for _, f in ckpts[:-max_to_keep] if max_to_keep else []:
        try:
            os.remove(f)
        except OSError:
            pass
    return path


def load_checkpoint(path):
    return torch.load(path, map_location="cpu", weights_only=False)
