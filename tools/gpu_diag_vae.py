"""VAE kernel + engine diagnostics vs the CPU oracle."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dalle_mtf_b200 import ops  # noqa: E402
from dalle_mtf_b200.vae_engine import VaeEngine  # noqa: E402
from oracle import vae as OV  # noqa: E402

DEV = "cuda"
OKS = []


def relerr(a, b):
    a, b = a.detach().double().flatten().cpu(), b.detach().double().flatten().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def check(name, got, ref, tol):
    e = relerr(got, ref)
    ok = e < tol and bool(torch.isfinite(got.float()).all())
    OKS.append(ok)
    print(f"[{'OK ' if ok else 'BAD'}] {name}: rel-fro {e:.3e} (tol {tol})", flush=True)


def conv_case(N, H, Cin, Cout, k, stride, transposed, seed, bf16=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, H, H, Cin, generator=g)
    if transposed:
        w = torch.randn(k, k, Cout, Cin, generator=g) * 0.1
    else:
        w = torch.randn(k, k, Cin, Cout, generator=g) * 0.1
    bias = torch.randn(Cout, generator=g) * 0.1
    if bf16:
        x = x.to(torch.bfloat16).float()
    xr = x.clone().requires_grad_(True); wr = w.clone().requires_grad_(True); br = bias.clone().requires_grad_(True)
    if transposed:
        y = OV.conv2d_transpose_same(xr, wr, br)
    else:
        y = OV.conv2d_same(xr, wr, br, stride)
    dy = torch.randn(y.shape, generator=g)
    if bf16:
        dy = dy.to(torch.bfloat16).float()
    y.backward(dy)
    act = torch.bfloat16 if bf16 else torch.float32
    c = ops.conv_desc(N, H, H, Cin, Cout, k, k, stride, transposed=transposed, act_f32=not bf16)
    xd, wd, bd, dyd = x.to(DEV, act), w.to(DEV), bias.to(DEV), dy.to(DEV, act)
    yd = torch.empty(y.shape, dtype=act, device=DEV)
    ops.conv2d_fwd(c, xd, wd, bd, None, yd)
    dxd = torch.empty(x.shape, dtype=act, device=DEV)
    ops.conv2d_dgrad(c, dyd, wd, None, None, dxd)
    dwd = torch.zeros(w.shape, device=DEV); dbd = torch.zeros(Cout, device=DEV)
    ops.conv2d_wgrad(c, xd, dyd, dwd, dbd)
    torch.cuda.synchronize()
    tag = f"conv N={N} H={H} Cin={Cin} Cout={Cout} k={k} s={stride} T={int(transposed)} bf16={int(bf16)}"
    tol = 1e-2 if bf16 else 1e-5
    check(tag + " fwd", yd, y, tol)
    check(tag + " dgrad", dxd, xr.grad, tol)
    check(tag + " wgrad", dwd, wr.grad, 1e-4 if bf16 else 1e-5)
    check(tag + " dbias", dbd, br.grad, 1e-4 if bf16 else 1e-5)


def conv_tc_case(N, H, Cin, Cout, k, stride, relu, with_res, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, H, H, Cin, generator=g).to(torch.bfloat16)
    w = (torch.randn(k, k, Cin, Cout, generator=g) * (1.0 / (k * k * Cin) ** 0.5)).to(torch.bfloat16)
    bias = torch.randn(Cout, generator=g) * 0.1
    y = OV.conv2d_same(x.float(), w.float(), bias, stride)
    if relu:
        y = torch.relu(y)
    res = None
    if with_res:
        res = torch.randn(y.shape, generator=g).to(torch.bfloat16)
        y = y + res.float()
    c = ops.conv_desc(N, H, H, Cin, Cout, k, k, stride, act_f32=False, relu=relu)
    yd = torch.zeros(y.shape, dtype=torch.bfloat16, device=DEV)
    ops.conv2d_fwd_tc(c, x.to(DEV), w.to(DEV).view(-1, Cout), bias.to(DEV), None if res is None else res.to(DEV), yd)
    torch.cuda.synchronize()
    check(f"conv_tc N={N} H={H} Cin={Cin} Cout={Cout} k={k} s={stride} relu={int(relu)} res={int(with_res)}", yd, y, 1e-2)


def tokenizer_case(size, B, seed):
    """bf16 tensor-core tokenizer vs the fp32 oracle: logits error and token match rate (near-ties may flip)."""
    cb = [[3, 64], [3, 128], [3, 256]]
    p = OV.init_params(cb, 512, seed=seed)
    g = torch.Generator().manual_seed(seed)
    img = (torch.randint(0, 256, (B, size, size, 3), generator=g).float() - 127.5) / 127.5
    logits = OV.encoder(p, img, cb)
    tok_ref = logits.argmax(-1).reshape(B, -1)
    eng = VaeEngine(512, size, cb, use_bf16=True)
    eng.load_params(p)
    tok = eng.encode_tokens(img.to(DEV)).cpu().long()
    torch.cuda.synchronize()
    check(f"tokenizer(bf16 tc) size={size} B={B} logits", eng._b["logits"], logits.reshape(-1, 512), 3e-2)
    match = (tok == tok_ref).float().mean().item()
    top2 = logits.reshape(-1, 512).topk(2, -1).values
    margin = (top2[:, 0] - top2[:, 1])
    mism = (tok.flatten() != tok_ref.flatten())
    print(f"      token match rate vs fp32 oracle: {match:.4f}; median top1-top2 margin all={margin.median():.4f} "
          f"mismatched={margin[mism].median().item() if mism.any() else float('nan'):.5f}")
    OKS.append(match > 0.9)


def engine_case(convblocks, K, size, B, hard, tau, use_bf16, seed):
    g = torch.Generator().manual_seed(seed)
    p = OV.init_params(convblocks, K, seed=seed)
    for k in p:
        if k.endswith("/bias"):
            p[k] = torch.randn(p[k].shape, generator=g) * 0.05
    img = torch.rand(B, size, size, 3, generator=g) * 2 - 1
    hw = size // (2 ** len(convblocks))
    u = torch.rand(B * hw * hw, K, generator=g).clamp_(1e-9, 1.0)
    loss, out, logits, grads = OV.loss_and_grads(p, img, u.view(B, hw, hw, K), convblocks, tau, hard, bf16=False)
    eng = VaeEngine(K, size, convblocks, use_bf16=use_bf16)
    eng.load_params(p)
    eng.zero_grads()
    imgd, ud = img.to(DEV), u.to(DEV)
    acc = torch.zeros(1, device=DEV)
    recon = eng.forward(imgd, ud, tau, hard, loss_accum=acc)
    eng.backward()
    torch.cuda.synchronize()
    tag = f"vae blocks={convblocks} K={K} size={size} B={B} hard={hard} tau={tau} bf16={use_bf16}"
    tol = 3e-2 if use_bf16 else 1e-4
    check(tag + " logits", eng._b["logits"], logits.reshape(-1, K), tol)
    idx_ref = OV.gumbel_softmax(logits, u.view(B, hw, hw, K), tau, True).argmax(-1).flatten()
    match = (eng._b["idx"].cpu().long() == idx_ref).float().mean().item()
    print(f"      sampled code match rate vs oracle: {match:.4f}")
    tok = eng.encode_tokens(imgd).cpu().long()
    tmatch = (tok == OV.encode_tokens(p, img, convblocks)).float().mean().item()
    print(f"      argmax token match rate vs oracle: {tmatch:.4f}")
    if not use_bf16:
        OKS.append(match == 1.0 and tmatch == 1.0)
        check(tag + " recon", recon, out, tol)
        check(tag + " loss", acc, loss.reshape(1), 1e-4)
        eg = eng.export_params(eng.grads)
        worst = max(relerr(eg[k], grads[k]) for k in grads)
        for k in grads:
            e = relerr(eg[k], grads[k])
            if e > 1e-3:
                print(f"      grad {k}: {e:.3e}")
        ok = worst < 1e-3
        OKS.append(ok)
        print(f"[{'OK ' if ok else 'BAD'}] {tag} worst grad rel-fro {worst:.3e}")
    else:
        check(tag + " loss (bf16 act)", acc, loss.reshape(1), 5e-2)


def main():
    conv_case(2, 8, 16, 24, 3, 1, False, 0)
    conv_case(2, 9, 5, 70, 3, 1, False, 1)       # odd sizes, channel tails
    conv_case(2, 8, 3, 32, 4, 2, False, 2)       # first layer: Cin = 3
    conv_case(3, 12, 20, 36, 4, 2, False, 3)
    conv_case(2, 6, 24, 16, 4, 2, True, 4)       # conv-transpose
    conv_case(1, 5, 70, 9, 4, 2, True, 5)
    conv_case(2, 8, 16, 3, 1, 1, False, 6)       # 1x1 output conv
    conv_case(2, 8, 32, 32, 3, 1, False, 7, bf16=True)
    conv_case(2, 8, 32, 16, 4, 2, True, 8, bf16=True)
    conv_tc_case(2, 16, 64, 64, 3, 1, True, False, 20)
    conv_tc_case(2, 16, 128, 128, 3, 1, False, True, 21)
    conv_tc_case(3, 12, 64, 128, 3, 1, False, False, 22)     # ragged: 12 = 8 + 4 (TMA OOB + epilogue bounds)
    conv_tc_case(2, 16, 64, 128, 4, 2, False, False, 23)     # stride 2 via parity views
    conv_tc_case(5, 8, 128, 256, 4, 2, False, False, 24)     # 4x4 output: tile spans images
    conv_tc_case(2, 4, 256, 256, 3, 1, True, True, 25)
    conv_tc_case(1, 32, 256, 64, 3, 1, False, False, 26)
    tokenizer_case(64, 4, 30)
    engine_case([[2, 32], [2, 64]], 64, 16, 4, True, 1.0, False, 10)
    engine_case([[2, 32], [2, 64]], 64, 16, 4, False, 0.5, False, 11)
    engine_case([[3, 64], [3, 128], [3, 256]], 512, 32, 8, True, 1.0, False, 12)   # vae_example geometry
    engine_case([[2, 32], [2, 64]], 64, 16, 4, False, 1.0, True, 13)
    print("SUMMARY vae:", "all ok" if all(OKS) else f"{OKS.count(False)} FAILURES")
    return 0 if all(OKS) else 1


if __name__ == "__main__":
    sys.exit(main())
