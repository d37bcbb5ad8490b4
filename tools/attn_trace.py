"""Per-CTA event timeline of the attention kernels (DEVELOPMENT library: DB200_LIB=.../libdalle_b200_dev.so).
Prints, for a few traced CTAs, every recorded event as (role, type, block index, cycles since the CTA's first event).
Event types — forward: 1 roles start, 2 K_j load issued, 3 V_j load issued, 9 MMA warp starts waiting for K_j,
4 S_j issued, 11 V_j landed, 12 P_j in TMEM, 5 P.V_j issued, 6 math warp saw S_j, 7 math warp wrote its P_j share,
8 epilogue starts, 10 role done, 13 CTA exit; math detail: 20 logits in registers, 21 row maximum known, 22 exponentials done.  dK/dV: 2 Q_i / 3 dO_i load issued, 4 S^T issued, 14 dP^T issued,
12 P^T in TMEM, 5 dV issued, 15 dS^T in TMEM, 16 dK issued, 6/7 phase A begin/end, 17/18 phase B begin/end."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("DB200_LIB", os.path.join(ROOT, "dalle_mtf_b200", "libdalle_b200_dev.so"))
from dalle_mtf_b200 import lib as L, ops  # noqa: E402

EV, STRIDE, SLOTS = 320, 37, 40
ROLES = ["tma", "mma", "math0", "mathN"]


def dump(buf, title, want_cta):
    t = buf.cpu().view(SLOTS, 4, EV)
    print(f"==== {title}")
    per_sm = {}
    for slot in range(SLOTS):
        hdr = int(t[slot, 1, 0].item()) & ((1 << 64) - 1)
        if hdr == 0:
            continue
        sm = (hdr >> 40) & 0xFFFF
        cta = hdr & ((1 << 40) - 1)
        evs = []
        for role in range(4):
            for i in range(1, EV):
                v = int(t[slot, role, i].item()) & ((1 << 64) - 1)
                if v == 0:
                    break
                evs.append((v & ((1 << 40) - 1), role, (v >> 52) & 0xFFF, (v >> 40) & 0xFFF))
        if not evs:
            continue
        evs.sort()
        t0 = evs[0][0]
        dur = evs[-1][0] - t0
        per_sm.setdefault(sm, []).append((t0, evs[-1][0], cta))
        n_blocks = max(e[3] for e in evs if e[2] == 1)
        print(f"-- CTA {cta} on SM {sm}: {n_blocks} blocks, {dur} cycles from first to last event "
              f"({dur / max(n_blocks, 1):.0f} per block)")
        if cta in want_cta or slot < 2:
            for c, role, typ, j in evs:
                print(f"   {c - t0:8d}  {ROLES[role]:6s} type={typ:2d} j={j}")
    return per_sm


def main():
    lib = ctypes.CDLL(os.environ["DB200_LIB"])
    B, S, H, dh = 32, 1280, 4, 128
    qkv = (torch.randn(B, S, 3, H, dh, device="cuda") * 0.3).to(torch.bfloat16)
    dout = torch.randn(B, S, H, dh, device="cuda").to(torch.bfloat16)
    out = torch.zeros(B, S, H, dh, dtype=torch.bfloat16, device="cuda")
    lse = torch.zeros(B, H, S, device="cuda")
    dqkv = torch.zeros_like(qkv)
    delta = torch.zeros(B, H, S, device="cuda")
    acc = torch.zeros(1, device="cuda")
    for _ in range(2):
        ops.attn_fwd(qkv, out, lse, B, S, H, dh, 1.0)
        ops.attn_bwd(qkv, out, dout, lse, acc, delta, dqkv, B, S, H, dh, 1.0)
    torch.cuda.synchronize()
    buf = torch.zeros(SLOTS * 4 * EV, dtype=torch.int64, device="cuda")
    assert lib.db200_dev_attn_trace(ctypes.c_void_p(buf.data_ptr()), SLOTS) == 0
    ops.attn_fwd(qkv, out, lse, B, S, H, dh, 1.0)
    torch.cuda.synchronize()
    # CTA 0 = heaviest tile of the first wave (cold start); 370 / 740 = later waves (steady state)
    dump(buf, "forward (32,1280,4,128)", {0, 370, 740, 1110})
    buf.zero_()
    ops.attn_bwd(qkv, out, dout, lse, acc, delta, dqkv, B, S, H, dh, 1.0)
    torch.cuda.synchronize()
    dump(buf, "backward dK/dV (32,1280,4,128)", {0, 370, 740})
    assert lib.db200_dev_attn_trace(ctypes.c_void_p(0), 0) == 0


if __name__ == "__main__":
    main()
