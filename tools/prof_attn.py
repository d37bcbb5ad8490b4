"""One forward and one backward attention launch per shape (after a warm-up) — the target of `ncu -k regex:attn_`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dalle_mtf_b200 import ops  # noqa: E402

for (B, S, H, dh) in ((32, 1280, 4, 128), (16, 1280, 16, 64)):
    qkv = (torch.randn(B, S, 3, H, dh, device="cuda") * 0.3).to(torch.bfloat16)
    dout = torch.randn(B, S, H, dh, device="cuda").to(torch.bfloat16)
    out = torch.zeros(B, S, H, dh, dtype=torch.bfloat16, device="cuda")
    lse = torch.zeros(B, H, S, device="cuda")
    dqkv = torch.zeros_like(qkv)
    delta = torch.zeros(B, H, S, device="cuda")
    acc = torch.zeros(1, device="cuda")
    for _ in range(1):
        ops.attn_fwd(qkv, out, lse, B, S, H, dh, 1.0)
        ops.attn_bwd(qkv, out, dout, lse, acc, delta, dqkv, B, S, H, dh, 1.0)
    torch.cuda.synchronize()
