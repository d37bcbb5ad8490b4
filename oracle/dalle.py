"""ORACLE (test infrastructure, NOT product code) — CPU restatement of the reference DALL-E transformer.

PARITY UNPINNED: the reference (EleutherAI/DALLE-mtf) ships no tests, golden vectors or fixtures, and its arithmetic
lives in mesh_tensorflow 0.1.18 / tensorflow 2.4.0, neither of which is vendored nor installable here (Python 3.12,
no network).  This file restates the algorithm line by line from the reference's own Python and from the published
behaviour of those two packages (items marked ‡ are recalled mtf/TF semantics that cannot be re-verified offline; each
sits behind a named switch in `Quirks` so a later correction is a one-line change).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.

Every function cites the reference file:line it follows (paths relative to the reference root).
"""
import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass
class Quirks:
    """‡ switches: recalled mesh-tensorflow behaviour."""
    attn_scale: float = 1.0          # ‡ AttentionParams folds 1/sqrt(dh) into the q initialiser -> no runtime scale
    mask_value: float = -1e10        # src/dalle_mtf/models.py:226
    ln_eps: float = 1e-5             # src/dalle_mtf/models.py:373
    loss_mean_over_all: bool = True  # src/dalle_mtf/models.py:353-354 (no padding mask)


@dataclass
class DalleConfig:
    n_embd: int
    n_layers: int
    n_heads: int
    text_vocab_size: int
    image_vocab_size: int
    text_seq_len: int
    image_seq_len: int

    @property
    def total_tokens(self):  # src/dalle_mtf/models.py:157
        return self.text_vocab_size + self.image_vocab_size + 1

    @property
    def eos_token_id(self):  # src/dalle_mtf/models.py:158
        return self.total_tokens - 1

    @property
    def seq_len(self):  # src/dalle_mtf/models.py:153
        return self.text_seq_len + self.image_seq_len

    @property
    def head_dim(self):  # src/dalle_mtf/models.py:167
        return self.n_embd // self.n_heads


def init_params(cfg: DalleConfig, seed: int = 0):
    """Reference initialisers.  Names are the reference's variable names."""
    g = torch.Generator().manual_seed(seed)
    d, V, S, H, dh, L = cfg.n_embd, cfg.total_tokens, cfg.seq_len, cfg.n_heads, cfg.head_dim, cfg.n_layers

    def normal(shape, std):
        return torch.randn(*shape, generator=g) * std

    p = {}
    p["embedding/wte"] = normal((V, d), 0.02)                 # src/dalle_mtf/models.py:189-192
    p["positional_embedding/wpe"] = normal((S, d), 0.01)      # src/dalle_mtf/models.py:206-208
    for i in range(L):
        pre = f"layer_{i}/"
        p[pre + "norm_1/g"] = torch.ones(d)                   # src/dalle_mtf/models.py:378-385
        p[pre + "norm_1/b"] = torch.zeros(d)
        # ‡ mtf attention_params_simple: q ~ N(0, d^-.5 * dh^-.5), k,v ~ N(0, d^-.5), o ~ N(0, (H*dh)^-.5)
        p[pre + "attn/q"] = normal((d, H * dh), d ** -0.5 * dh ** -0.5)
        p[pre + "attn/k"] = normal((d, H * dh), d ** -0.5)
        p[pre + "attn/v"] = normal((d, H * dh), d ** -0.5)
        p[pre + "attn/o"] = normal((H * dh, d), (H * dh) ** -0.5)
        p[pre + "attn/compute_output_bias/o_b"] = torch.zeros(d)  # src/dalle_mtf/models.py:306-310
        p[pre + "norm_2/g"] = torch.ones(d)
        p[pre + "norm_2/b"] = torch.zeros(d)
        p[pre + "mlp/mlp_linear_1/kernel"] = normal((d, 4 * d), 0.02)  # src/dalle_mtf/models.py:320,361-371
        p[pre + "mlp/mlp_linear_1/bias"] = torch.zeros(4 * d)
        p[pre + "mlp/mlp_linear_2/kernel"] = normal((4 * d, d), 0.02 / math.sqrt(L))  # :321, :364-366
        p[pre + "mlp/mlp_linear_2/bias"] = torch.zeros(d)
    p["to_logits/layer_norm/g"] = torch.ones(d)               # src/dalle_mtf/models.py:392-393
    p["to_logits/layer_norm/b"] = torch.zeros(d)
    p["to_logits/linear_out/kernel"] = normal((d, V), 0.02)
    p["to_logits/linear_out/bias"] = torch.zeros(V)
    return p


def _r(x, bf16):
    """Round an activation to bf16 at the reference's cast points (activation_dtype, src/dalle_mtf/ops.py:76-82)."""
    return x.to(torch.bfloat16).to(torch.float32) if bf16 else x


def layer_norm(x, g, b, eps):
    """src/dalle_mtf/models.py:373-389 + src/dalle_mtf/layers.py:30-33 (biased variance, eps inside rsqrt)."""
    x = x - x.mean(-1, keepdim=True)
    s = (x * x).mean(-1, keepdim=True)
    return x * torch.rsqrt(s + eps) * g + b


def shift_labels(tokens, eos_id):
    """src/dalle_mtf/models.py:407-410: pad with EOS on the right, drop the first token."""
    return torch.cat([tokens[:, 1:], torch.full_like(tokens[:, :1], eos_id)], dim=1)


def assemble_tokens(text_ids, image_ids, text_vocab_size):
    """src/model_fns.py:117-122: tokens = concat(text, image_ids + text_vocab_size)."""
    return torch.cat([text_ids, image_ids + text_vocab_size], dim=1)


def forward(params, tokens, cfg: DalleConfig, bf16=False, quirks: Quirks = None, faithful=False,
            return_hidden=False):
    """DALLE.forward (src/dalle_mtf/models.py:397-416).  tokens: int64 [B,S].  Returns (loss, loss_batch, logits).

    bf16=True rounds weights and every op output to bf16 like activation_dtype=bf16 does; attention logits, the final
    logits and the loss stay fp32 (‡ mtf attention computes logits in fp32; models.py:395, :358).
    faithful=True executes the embedding and the CE targets as one-hot contractions and materialises the [S,S] mask —
    what mesh-tensorflow actually runs (‡ mtf.gather / integer targets lower to one_hot x einsum); used only for the
    timed "reference CPU path".  The numbers are identical to faithful=False.
    """
    q_ = quirks or Quirks()
    B, S = tokens.shape
    d, H, dh, V = cfg.n_embd, cfg.n_heads, cfg.head_dim, cfg.total_tokens
    W = (lambda n: _r(params[n], bf16))
    # --- embedding (models.py:186-201) + positional embedding (models.py:203-219)
    if faithful:
        x = F.one_hot(tokens, V).to(torch.float32) @ W("embedding/wte")
    else:
        x = W("embedding/wte")[tokens]
    x = _r(x + W("positional_embedding/wpe")[:S], bf16)
    # --- mask (models.py:221-227): -1e10 where query index < key index
    i = torch.arange(S)[:, None]
    j = torch.arange(S)[None, :]
    mask = (i < j).to(torch.float32) * q_.mask_value
    hidden = []
    for l in range(cfg.n_layers):
        pre = f"layer_{l}/"
        # attention (models.py:229-315)
        h = _r(layer_norm(x, W(pre + "norm_1/g"), W(pre + "norm_1/b"), q_.ln_eps), bf16)
        q = _r(h @ W(pre + "attn/q"), bf16).view(B, S, H, dh)
        k = _r(h @ W(pre + "attn/k"), bf16).view(B, S, H, dh)
        v = _r(h @ W(pre + "attn/v"), bf16).view(B, S, H, dh)
        logits = torch.einsum("bihe,bjhe->bhij", q, k) * q_.attn_scale + mask  # fp32 ‡
        p = torch.softmax(logits, dim=-1)
        p = _r(p, bf16)  # ‡ weights cast to v's dtype before the PV einsum
        a = _r(torch.einsum("bhij,bjhe->bihe", p, v).reshape(B, S, H * dh), bf16)
        a = _r(a @ W(pre + "attn/o") + W(pre + "attn/compute_output_bias/o_b"), bf16)  # models.py:303-311
        x = _r(x + a, bf16)                                                              # models.py:330
        # mlp (models.py:317-324): relu(x W1 + b1) W2 + b2
        h = _r(layer_norm(x, W(pre + "norm_2/g"), W(pre + "norm_2/b"), q_.ln_eps), bf16)
        h1 = _r(torch.relu(h @ W(pre + "mlp/mlp_linear_1/kernel") + W(pre + "mlp/mlp_linear_1/bias")), bf16)
        h2 = _r(h1 @ W(pre + "mlp/mlp_linear_2/kernel") + W(pre + "mlp/mlp_linear_2/bias"), bf16)
        x = _r(x + h2, bf16)                                                             # models.py:333
        if return_hidden:
            hidden.append(x)
    # --- to_logits (models.py:391-395): LN -> dense -> cast fp32
    hf = _r(layer_norm(x, W("to_logits/layer_norm/g"), W("to_logits/layer_norm/b"), q_.ln_eps), bf16)
    logits = hf @ W("to_logits/linear_out/kernel") + W("to_logits/linear_out/bias")
    logits = logits.to(torch.promote_types(logits.dtype, torch.float32))  # models.py:395 (fp64 kept for FD tests)
    # --- loss (models.py:407-411, 348-359)
    labels = shift_labels(tokens, cfg.eos_token_id)
    if faithful:
        logp = torch.log_softmax(logits, -1)
        loss_batch = -(F.one_hot(labels, V).to(torch.float32) * logp).sum(-1)
    else:
        loss_batch = torch.logsumexp(logits, -1) - logits.gather(-1, labels[..., None]).squeeze(-1)
    loss = loss_batch.mean()
    if return_hidden:
        return loss, loss_batch, logits, hidden
    return loss, loss_batch, logits


def loss_and_grads(params, tokens, cfg, bf16=False, quirks=None):
    """mtf.gradients([loss], trainable_variables) (src/optimizers.py:34) via autograd on the restated graph."""
    leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    loss, loss_batch, logits = forward(leaves, tokens, cfg, bf16=bf16, quirks=quirks)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    return loss.detach(), loss_batch.detach(), logits.detach(), grads
