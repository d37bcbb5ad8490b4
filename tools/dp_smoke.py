"""Two-or-more-rank smoke test of the data-parallel path (both NCCL communicators): a few dalle steps at a tiny shape.
Run under torch.distributed.run; prints 'DP SMOKE OK' on rank 0."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dalle_mtf_b200.dist import DataParallel  # noqa: E402


def main():
    dp = DataParallel().init()
    dev = torch.device("cuda", torch.cuda.current_device())
    flat = torch.full((8 * 1024 * 1024,), float(dp.rank + 1), device=dev)
    hook = dp.make_bucket_hook(flat)
    n = flat.numel()
    for _ in range(3):
        flat.fill_(float(dp.rank + 1))
        for k in range(4, 0, -1):                     # tail-first buckets, the last one starts at offset 0
            hook((k - 1) * n // 4, k * n // 4)
        dp.wait()
        torch.cuda.synchronize()
    want = dp.world * (dp.world + 1) / 2
    ok = bool((flat == want).all().item())
    buf = torch.zeros(dp.world * 1024, dtype=torch.bfloat16, device=dev)
    buf[dp.rank * 1024:(dp.rank + 1) * 1024] = dp.rank + 1
    dp.all_gather_inplace(buf, 1024)
    torch.cuda.synchronize()
    ok = ok and all(bool((buf[r * 1024:(r + 1) * 1024] == r + 1).all().item()) for r in range(dp.world))
    dp.barrier()
    if dp.rank == 0:
        print("DP SMOKE OK" if ok else "DP SMOKE WRONG RESULT", flush=True)
    dp.shutdown()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
