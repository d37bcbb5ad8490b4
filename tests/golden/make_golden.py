"""Generates the golden fixtures that freeze the oracle's outputs (run from the repo root: python tests/golden/make_golden.py).

The reference itself cannot be imported here (tensorflow 2.4 / mesh_tensorflow 0.1.18 are not installable: Python 3.12,
no network), and it ships no golden vectors, so these fixtures pin the ORACLE (regression guard), not the reference.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import dalle as O  # noqa: E402
from oracle import vae as OV  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    torch.manual_seed(0)
    cfgd = dict(n_embd=32, n_layers=2, n_heads=2, text_vocab_size=40, image_vocab_size=16, text_seq_len=6,
                image_seq_len=6)
    cfg = O.DalleConfig(**cfgd)
    tokens = torch.randint(0, cfg.total_tokens - 1, (2, cfg.seq_len), generator=torch.Generator().manual_seed(7))
    loss, lb, logits, grads = O.loss_and_grads(O.init_params(cfg, 11), tokens, cfg)
    torch.save({"cfg": cfgd, "seed": 11, "tokens": tokens, "loss": loss, "logits_slice": logits[0, :, :8].clone(),
                "grad_norms": {k: grads[k].norm().item() for k in
                               ("embedding/wte", "layer_0/attn/q", "layer_1/mlp/mlp_linear_2/kernel",
                                "to_logits/linear_out/kernel")}},
               os.path.join(HERE, "dalle_tiny.pt"))
    cb = [[2, 8], [2, 16]]
    g = torch.Generator().manual_seed(5)
    img = torch.rand(2, 8, 8, 3, generator=g) * 2 - 1
    u = torch.rand(2, 2, 2, 12, generator=g).clamp_(1e-9, 1)
    p = OV.init_params(cb, 12, seed=13)
    loss, out, logits = OV.forward(p, img, u, cb, 1.0, True)
    torch.save({"convblocks": cb, "K": 12, "seed": 13, "img": img, "u": u, "loss": loss,
                "tokens": logits.argmax(-1), "out_px": out[0, 0, 0].clone()}, os.path.join(HERE, "vae_tiny.pt"))
    print("wrote fixtures to", HERE)


if __name__ == "__main__":
    main()
