"""Timing decomposition of the warp-specialised attention forward (DEVELOPMENT library only: DB200_LIB=..._dev.so).
DB200_ATTN_EXP bits: 1 = no MUFU (ex2 skipped), 2 = no softmax work at all (pure hand-off chain), 4 = no cross-group
exchange.  Results are WRONG under any of them; only the time is of interest."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, torch
sys.path.insert(0, %r)
from dalle_mtf_b200 import ops
for (B, S, H, dh) in ((32, 1280, 4, 128), (16, 1280, 16, 64)):
    qkv = (torch.randn(B, S, 3, H, dh, device="cuda") * 0.3).to(torch.bfloat16)
    out = torch.zeros(B, S, H, dh, dtype=torch.bfloat16, device="cuda")
    lse = torch.zeros(B, H, S, device="cuda")
    for _ in range(3):
        ops.attn_fwd(qkv, out, lse, B, S, H, dh, 1.0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.attn_fwd(qkv, out, lse, B, S, H, dh, 1.0)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"  dh={dh}: fwd {ms*1e3:7.1f} us  {4.0*S*S*dh*B*H/2/ms/1e9:7.1f} TFLOP/s", flush=True)
    if os.environ.get("DB200_ATTN_EXP", "0") == "0":
        dout = torch.randn(B, S, H, dh, device="cuda").to(torch.bfloat16)
        dqkv = torch.zeros_like(qkv); delta = torch.zeros(B, H, S, device="cuda"); acc = torch.zeros(1, device="cuda")
        for _ in range(3):
            ops.attn_bwd(qkv, out, dout, lse, acc, delta, dqkv, B, S, H, dh, 1.0)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            ops.attn_bwd(qkv, out, dout, lse, acc, delta, dqkv, B, S, H, dh, 1.0)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"  dh={dh}: bwd {ms*1e3:7.1f} us  {2.5*4.0*S*S*dh*B*H/2/ms/1e9:7.1f} TFLOP/s (incl. delta pre-pass)", flush=True)
''' % ROOT

for ng, persist in (("2", "3"), ("4", "3"), ("2", "0")):
    for exp in ("0", "1"):
        env = dict(os.environ, DB200_LIB=os.path.join(ROOT, "dalle_mtf_b200", "libdalle_b200_dev.so"),
                   DB200_ATTN_NG=ng, DB200_ATTN_EXP=exp, DB200_ATTN_PERSIST=persist)
        print(f"NG={ng} PERSIST={persist} EXP={exp}", flush=True)
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=300)
        print(r.stdout + r.stderr[-600:], flush=True)
