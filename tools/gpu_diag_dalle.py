"""End-to-end DALL-E engine diagnostics vs the CPU oracle (loss, logits, every gradient, one optimiser step)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dalle_mtf_b200.dalle_engine import DalleEngine  # noqa: E402
from oracle import dalle as O  # noqa: E402
from oracle import optim as OO  # noqa: E402


def relerr(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def run_case(d, L, H, tv, iv, ts, isl, B, recompute=False, seed=0):
    cfg = O.DalleConfig(d, L, H, tv, iv, ts, isl)
    params = O.init_params(cfg, seed)
    # make biases / LN params non-trivial so their gradients and use are exercised
    g = torch.Generator().manual_seed(seed + 1)
    for k in params:
        if k.endswith("/b") or k.endswith("bias") or k.endswith("o_b"):
            params[k] = torch.randn(params[k].shape, generator=g) * 0.02
        if k.endswith("/g"):
            params[k] = 1 + torch.randn(params[k].shape, generator=g) * 0.05
    tokens = torch.randint(0, cfg.total_tokens - 1, (B, cfg.seq_len), generator=g)
    eng = DalleEngine(d, L, H, tv, iv, ts, isl, recompute_grad=recompute)
    eng.load_params(params)
    tok_dev = tokens.to(torch.int32).cuda()
    T = B * cfg.seq_len
    tag = f"d={d} L={L} H={H} V={cfg.total_tokens} S={cfg.seq_len} B={B} recompute={recompute}"
    # oracle (fp32, and with the reference's bf16 cast points)
    t0 = time.time()
    loss32, lb32, logits32, g32 = O.loss_and_grads(params, tokens, cfg, bf16=False)
    loss16, lb16, logits16, g16 = O.loss_and_grads(params, tokens, cfg, bf16=True)
    print(f"[{tag}] oracle {time.time() - t0:.1f}s loss fp32={loss32.item():.6f} bf16-emulated={loss16.item():.6f}")
    # engine
    eng.zero_grads()
    loss_acc = eng.forward(tok_dev)
    eng.backward(1.0 / T)
    torch.cuda.synchronize()
    loss = loss_acc.item() / T
    ok = True
    e32, e16 = abs(loss - loss32.item()) / loss32.item(), abs(loss - loss16.item()) / loss16.item()
    print(f"  loss engine={loss:.6f} rel vs fp32 oracle={e32:.2e} vs bf16 oracle={e16:.2e}")
    ok &= e32 < 2e-3
    lg = eng.logits(tok_dev).float().cpu()
    el32, el16 = relerr(lg, logits32), relerr(lg, logits16)
    print(f"  logits rel-fro vs fp32={el32:.2e} vs bf16 oracle={el16:.2e}")
    ok &= el32 < 2e-2
    grads = eng.export_params(eng.grads)
    worst = 0
    for k in sorted(g32):
        e_a, e_b = relerr(grads[k], g32[k]), relerr(grads[k], g16[k])
        base = relerr(g16[k], g32[k])
        worst = max(worst, e_a)
        flag = "" if e_a < 5e-2 else "  <-- BAD"
        if flag or k.startswith("layer_0") or "layer" not in k:
            print(f"    grad {k:48s} vs fp32 {e_a:.2e} vs bf16-oracle {e_b:.2e} (oracle bf16-vs-fp32 {base:.2e}){flag}")
        ok &= e_a < 5e-2
    print(f"  worst grad rel-fro vs fp32 oracle: {worst:.2e}")
    # one optimiser step (lr schedule value given) vs oracle applied to the ENGINE's gradients (isolates Adam/clip)
    hp = {"lr": 1e-3, "train_steps": 1000, "warmup_steps": 10}
    step = 5
    lr = OO.learning_rate(step, hp)
    m0 = {k: torch.zeros_like(v) for k, v in params.items()}
    newp, newm, newv, _, gn = OO.dalle_train_step(params, m0, m0, grads, step, hp)
    eng.optimizer_step(lr)
    torch.cuda.synchronize()
    after = eng.export_params()
    wp = max(relerr(after[k] - params[k], newp[k] - params[k]) for k in params if (newp[k] - params[k]).norm() > 0)
    print(f"  optimiser: gnorm oracle={gn.item():.5f} engine={eng.gnorm_sq.sqrt().item():.5f}; worst update rel err {wp:.2e}")
    ok &= wp < 1e-3
    print(f"[{'OK ' if ok else 'BAD'}] dalle e2e {tag}")
    return ok


def main():
    oks = [
        run_case(256, 2, 2, 1000, 100, 40, 24, 3),                 # dh=128, S=64 (single ragged tile)
        run_case(256, 2, 4, 1000, 100, 100, 60, 2),                # dh=64, S=160
        run_case(512, 2, 4, 3000, 512, 200, 100, 2, recompute=True),   # dh=128, S=300, recompute path
    ]
    print("SUMMARY dalle:", "all ok" if all(oks) else "FAILURES")
    return 0 if all(oks) else 1


if __name__ == "__main__":
    sys.exit(main())
