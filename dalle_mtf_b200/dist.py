"""Data parallelism: one process per GPU, NCCL over NVLink 5 / NVSwitch, bucketed gradient all-reduce that overlaps
with the rest of backward.

The reference's only parallelism is the layout rule "batch_dim:data" (configs/dalle_example.json:20-21): mtf lowers
every weight gradient to an all-reduce SUM across the `data` mesh axis (src/optimizers.py:34, src/model_fns.py:189),
one collective per variable and no overlap control.  Here the gradients live in ONE flat fp32 buffer in forward order;
backward completes it from the tail, and each finished contiguous range ("bucket": the vocabulary projection, then
one transformer layer at a time, then the embeddings) is all-reduced asynchronously on NCCL's own stream while the
earlier layers are still back-propagating.  The loss scalar rides in the first bucket.  The optimiser waits on the
handles; clip-by-global-norm is computed after the reduce, identically on every rank (src/optimizers.py:101-102), so
there is no second collective.
"""
import os

import torch
import torch.distributed as dist


class DataParallel:
    def __init__(self):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.handles = []
        self.enabled = self.world > 1

    def init(self, backend=None):
        if torch.cuda.is_available():
            torch.cuda.set_device(self.local_rank)
        if self.enabled and not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
            kw = {}
            if backend == "nccl":
                kw["device_id"] = torch.device("cuda", self.local_rank)
            dist.init_process_group(backend=backend, rank=self.rank, world_size=self.world, **kw)
        return self

    def shard(self, global_batch):
        """Rows [rank*B/N, (rank+1)*B/N) of the global batch."""
        if global_batch % self.world != 0:
            raise ValueError(f"global batch {global_batch} is not divisible by the data-parallel size {self.world}")
        per = global_batch // self.world
        return self.rank * per, per

    # --- gradient buckets ------------------------------------------------------------------------------
    def make_bucket_hook(self, flat):
        """Returns on_bucket_ready(start, end) for DalleEngine.backward: async all-reduce(SUM) of flat[start:end]."""
        if not self.enabled:
            return None

        def hook(start, end):
            end = min(end, flat.numel())
            self.handles.append(dist.all_reduce(flat[start:end], op=dist.ReduceOp.SUM, async_op=True))

        return hook

    def all_reduce_now(self, t):
        if self.enabled:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t

    def wait(self):
        """Make the current stream wait for every outstanding bucket (no host block with NCCL)."""
        for h in self.handles:
            h.wait()
        self.handles = []

    def barrier(self):
        if self.enabled:
            dist.barrier()

    def max_over_ranks(self, value):
        if not self.enabled:
            return value
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        t = torch.tensor([value], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def shutdown(self):
        if self.enabled and dist.is_initialized():
            dist.destroy_process_group()
