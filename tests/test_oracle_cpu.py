"""CPU tests of the oracle itself.  The reference ships no tests / golden vectors (parity unpinned), so the oracle is
pinned by (a) self-consistency properties of the restated math, (b) closed-form checks against the formulas the
reference states in-tree (src/optimizers.py:128-178), (c) committed golden fixtures that freeze its current outputs.
"""
import math
import os

import pytest
import torch

from oracle import dalle as O
from oracle import optim as OO
from oracle import vae as OV

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def small_cfg():
    return O.DalleConfig(n_embd=32, n_layers=2, n_heads=2, text_vocab_size=50, image_vocab_size=16, text_seq_len=6,
                         image_seq_len=4)


def test_label_shift_and_eos():
    cfg = small_cfg()
    assert cfg.total_tokens == 50 + 16 + 1 and cfg.eos_token_id == cfg.total_tokens - 1   # models.py:157-158
    t = torch.arange(20).view(2, 10)
    lab = O.shift_labels(t, cfg.eos_token_id)
    assert torch.equal(lab[:, :-1], t[:, 1:]) and (lab[:, -1] == cfg.eos_token_id).all()


def test_token_assembly_offsets_image_ids():
    text = torch.tensor([[1, 2, 3]])
    img = torch.tensor([[0, 5]])
    assert O.assemble_tokens(text, img, 50).tolist() == [[1, 2, 3, 50, 55]]              # model_fns.py:117-122


def test_causality_future_tokens_do_not_change_past_logits():
    cfg = small_cfg()
    p = O.init_params(cfg, 0)
    g = torch.Generator().manual_seed(0)
    t = torch.randint(0, cfg.total_tokens - 1, (1, cfg.seq_len), generator=g)
    _, _, l0 = O.forward(p, t, cfg)
    t2 = t.clone()
    t2[0, 7] = (t2[0, 7] + 1) % (cfg.total_tokens - 1)
    _, _, l1 = O.forward(p, t2, cfg)
    assert torch.allclose(l0[0, :7], l1[0, :7], atol=1e-6)
    assert not torch.allclose(l0[0, 7:], l1[0, 7:], atol=1e-6)


def test_faithful_graph_equals_algorithmic_graph():
    cfg = small_cfg()
    p = O.init_params(cfg, 1)
    t = torch.randint(0, cfg.total_tokens - 1, (2, cfg.seq_len), generator=torch.Generator().manual_seed(1))
    a = O.forward(p, t, cfg, faithful=False)
    b = O.forward(p, t, cfg, faithful=True)
    assert torch.allclose(a[0], b[0], rtol=1e-5) and torch.allclose(a[2], b[2], atol=1e-5)


def test_loss_is_mean_over_all_positions_including_padding():
    cfg = small_cfg()
    p = O.init_params(cfg, 2)
    t = torch.randint(0, cfg.total_tokens - 1, (3, cfg.seq_len), generator=torch.Generator().manual_seed(2))
    loss, loss_batch, _ = O.forward(p, t, cfg)
    assert loss_batch.shape == (3, cfg.seq_len)
    assert torch.allclose(loss, loss_batch.mean())                                         # models.py:353-354


def test_autograd_matches_finite_differences():
    cfg = O.DalleConfig(16, 1, 2, 10, 5, 3, 2)
    p = {k: v.double() for k, v in O.init_params(cfg, 3).items()}
    t = torch.randint(0, cfg.total_tokens - 1, (2, cfg.seq_len), generator=torch.Generator().manual_seed(3))
    leaves = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    O.forward(leaves, t, cfg)[0].backward()
    for name in ("layer_0/attn/q", "layer_0/mlp/mlp_linear_2/kernel", "to_logits/linear_out/bias", "embedding/wte"):
        w = p[name]
        idx = tuple(int(i) for i in torch.tensor(w.shape) // 2)
        if name == "embedding/wte":
            idx = (int(t[0, 0]), 1)
        eps = 1e-5
        wp = {k: v.clone() for k, v in p.items()}
        wm = {k: v.clone() for k, v in p.items()}
        wp[name][idx] += eps
        wm[name][idx] -= eps
        fd = (O.forward(wp, t, cfg)[0] - O.forward(wm, t, cfg)[0]) / (2 * eps)
        assert abs(fd.item() - leaves[name].grad[idx].item()) < 1e-6 * max(1, abs(fd.item())), name


def test_lr_schedule_table():
    hp = {"lr": 1e-3, "train_steps": 100000}
    lr = lambda s: OO.learning_rate(s, hp)
    assert lr(0) == 0.0                                                   # warm-up multiplies by step/3000
    assert math.isclose(lr(1), 1e-3 * (0.9 * 0.5 * (1 + math.cos(math.pi * 1 / 1e5)) + 0.1) / 3000, rel_tol=1e-12)
    assert lr(2999) < lr(3000)
    assert math.isclose(lr(3000), 1e-3 * (0.9 * 0.5 * (1 + math.cos(math.pi * 0.03)) + 0.1), rel_tol=1e-12)
    assert math.isclose(lr(50000), 1e-3 * 0.55, rel_tol=1e-12)
    assert math.isclose(lr(100000), 1e-4, rel_tol=1e-12) and math.isclose(lr(200000), 1e-4, rel_tol=1e-12)
    assert math.isclose(OO.learning_rate(50000, dict(hp, lr_decay="linear")), 1e-3 * 0.55, rel_tol=1e-12)


def test_adam_mtf_closed_form_first_step():
    """After one step from m=v=0: p -= lr * 0.1 g / (sqrt(0.001) |g| + 1e-6)   (src/optimizers.py:155-172)."""
    p, g = torch.tensor([1.0, -2.0]), torch.tensor([0.5, -0.25])
    z = torch.zeros(2)
    p2, m2, v2 = OO.adam_mtf_step(p, z, z, g, lr=0.1)
    expect = p - 0.1 * (0.1 * g) / (torch.sqrt(0.001 * g * g) + 1e-6)
    assert torch.allclose(p2, expect) and torch.allclose(m2, 0.1 * g) and torch.allclose(v2, 0.001 * g * g)


def test_adam_tf_first_step_is_lr_sized():
    p, g = torch.tensor([1.0]), torch.tensor([0.3])
    p2, _, _ = OO.adam_tf_step(p, torch.zeros(1), torch.zeros(1), g, lr=1e-3, t=1)
    assert abs((p - p2).item() - 1e-3) < 1e-6          # bias-corrected first step has magnitude ~lr


def test_clip_by_global_norm():
    g = {"a": torch.tensor([3.0]), "b": torch.tensor([4.0])}
    c, gn = OO.clip_by_global_norm(g, 1.0)
    assert math.isclose(gn.item(), 5.0, rel_tol=1e-6) and math.isclose(c["a"].item(), 0.6, rel_tol=1e-6)
    c2, _ = OO.clip_by_global_norm({"a": torch.tensor([0.3])}, 1.0)
    assert math.isclose(c2["a"].item(), 0.3, rel_tol=1e-6)   # below the threshold: unchanged


# ----------------------------------------------------------------------------------------------------------- VAE
def test_conv_transpose_is_adjoint_of_same_conv():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 8, 8, 3, generator=g)
    w = torch.randn(4, 4, 3, 5, generator=g)
    y = OV.conv2d_same(x, w, None, 2)
    dy = torch.randn(y.shape, generator=g)
    # <conv(x), dy> == <x, convT(dy)> with the transposed-conv kernel layout [kh,kw,out,in] = w viewed as is
    lhs = (y * dy).sum()
    rhs = (x * OV.conv2d_transpose_same(dy, w, None)).sum()
    assert torch.allclose(lhs, rhs, rtol=1e-4)


def test_same_padding_stride2_matches_explicit_loop():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 6, 6, 2, generator=g)
    w = torch.randn(4, 4, 2, 3, generator=g)
    y = OV.conv2d_same(x, w, None, 2)
    ref = torch.zeros(1, 3, 3, 3)
    for oy in range(3):
        for ox in range(3):
            for kh in range(4):
                for kw in range(4):
                    iy, ix = 2 * oy + kh - 1, 2 * ox + kw - 1
                    if 0 <= iy < 6 and 0 <= ix < 6:
                        ref[0, oy, ox] += x[0, iy, ix] @ w[kh, kw]
    assert torch.allclose(y, ref, atol=1e-5)


def test_gumbel_hard_is_one_hot_with_soft_gradient():
    g = torch.Generator().manual_seed(2)
    logits = torch.randn(5, 7, generator=g, requires_grad=True)
    u = torch.rand(5, 7, generator=g).clamp_(1e-9, 1)
    y = OV.gumbel_softmax(logits, u, 0.7, hard=True)
    assert torch.allclose(y.sum(-1), torch.ones(5)) and ((y == 0) | (y == 1)).all()
    w = torch.randn(5, 7, generator=g)
    (y * w).sum().backward()
    l2 = logits.detach().clone().requires_grad_(True)
    (OV.gumbel_softmax(l2, u, 0.7, hard=False) * w).sum().backward()
    assert torch.allclose(logits.grad, l2.grad, atol=1e-6)   # straight-through: gradient of the soft sample


def test_argmax_first_max_tie_rule():
    x = torch.tensor([[1.0, 3.0, 3.0, 2.0]])
    assert x.argmax(-1).item() == 1


def test_image_seq_len_and_temperature():
    assert OV.image_seq_len(32, [[3, 64], [3, 128], [3, 256]]) == 16        # SURVEY fact 6 (32 px -> 16 tokens)
    assert OV.image_seq_len(256, [[3, 64], [3, 128], [3, 256]]) == 1024
    hp = {"temp_start": 1.0, "temp": 0.05, "temp_anneal_steps": 25000}
    assert OV.temperature(0, hp) == 1.0 and math.isclose(OV.temperature(12500, hp), 0.525)
    assert math.isclose(OV.temperature(10 ** 6, hp), 0.05) and OV.temperature(5, {}) == 1.0


# ----------------------------------------------------------------------------------------------------------- golden
@pytest.mark.parametrize("name", ["dalle_tiny", "vae_tiny"])
def test_golden_fixtures_freeze_the_oracle(name):
    path = os.path.join(GOLDEN, name + ".pt")
    fx = torch.load(path, weights_only=False)
    if name == "dalle_tiny":
        cfg = O.DalleConfig(**fx["cfg"])
        loss, lb, logits, grads = O.loss_and_grads(O.init_params(cfg, fx["seed"]), fx["tokens"], cfg)
        assert torch.allclose(loss, fx["loss"], rtol=1e-5)
        assert torch.allclose(logits[0, :, :8], fx["logits_slice"], atol=1e-5)
        for k, v in fx["grad_norms"].items():
            assert math.isclose(grads[k].norm().item(), v, rel_tol=1e-4), k
    else:
        p = OV.init_params(fx["convblocks"], fx["K"], seed=fx["seed"])
        loss, out, logits = OV.forward(p, fx["img"], fx["u"], fx["convblocks"], 1.0, True)
        assert torch.allclose(loss, fx["loss"], rtol=1e-5)
        assert torch.equal(logits.argmax(-1), fx["tokens"])
        assert torch.allclose(out[0, 0, 0], fx["out_px"], atol=1e-5)


@pytest.mark.parametrize("flip", ["attn_scale", "mask_value", "ln_eps"])
def test_golden_fixture_detects_a_flipped_quirk(flip):
    """Mutation check of the ‡ switches: the frozen golden outputs must NOT be reproduced when one recalled
    mesh-tensorflow behaviour is flipped — otherwise the fixture would not pin that behaviour at all."""
    fx = torch.load(os.path.join(GOLDEN, "dalle_tiny.pt"), weights_only=False)
    cfg = O.DalleConfig(**fx["cfg"])
    q = {"attn_scale": O.Quirks(attn_scale=cfg.head_dim ** -0.5), "mask_value": O.Quirks(mask_value=0.0),
         "ln_eps": O.Quirks(ln_eps=1e-2)}[flip]
    _, _, logits = O.forward(O.init_params(cfg, fx["seed"]), fx["tokens"], cfg, quirks=q)
    assert not torch.allclose(logits[0, :, :8], fx["logits_slice"], atol=1e-5)
