"""ORACLE (test infrastructure, NOT product code) — learning-rate schedule, global-norm clip, Adam variants.

PARITY UNPINNED (see oracle/dalle.py header).  The mtf Adam math is corroborated in-tree by the reference's own
commented restatement at src/optimizers.py:128-178.
"""
import math

import torch


def learning_rate(step, params):
    """src/optimizers.py:24-26, 46-76.  `step` is the global step BEFORE the update (src/model_fns.py:201)."""
    lr0 = params["lr"]
    end_step = params.get("lr_decay_end") or params["train_steps"]      # optimizers.py:24
    lr_decay = params.get("lr_decay") or "cosine"                        # optimizers.py:25
    warmup_steps = params.get("warmup_steps")
    warmup_steps = 3000 if warmup_steps is None else warmup_steps        # optimizers.py:26
    s = min(step, end_step)
    if lr_decay == "linear":   # tf.train.polynomial_decay(power=1, end = 0.1*lr)  optimizers.py:46-53
        lr = (lr0 - 0.1 * lr0) * (1 - s / end_step) + 0.1 * lr0
    elif lr_decay == "cosine":  # tf.train.cosine_decay(alpha=0.1)                   optimizers.py:54-60
        cosine = 0.5 * (1 + math.cos(math.pi * s / end_step))
        lr = lr0 * ((1 - 0.1) * cosine + 0.1)
    else:
        lr = lr0
    if warmup_steps > 0 and step < warmup_steps:                         # optimizers.py:62-76
        lr = lr * (step / warmup_steps)
    return lr


def clip_by_global_norm(grads, clip_norm):
    """src/optimizers.py:11-16: g *= clip / max(||g||_2, clip)."""
    gn = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).float()
    mult = clip_norm / torch.maximum(gn, torch.tensor(float(clip_norm)))
    return {k: g * mult for k, g in grads.items()}, gn


def adam_mtf_step(p, m, v, g, lr, beta1=0.9, beta2=0.999, eps=1e-6, weight_decay=0.0, name=""):
    """mtf AdamWeightDecayOptimizer.apply_grad (restated at src/optimizers.py:155-172): no bias correction."""
    m2 = beta1 * m + (1 - beta1) * g
    v2 = beta2 * v + (1 - beta2) * g * g
    upd = m2 / (torch.sqrt(v2) + eps)
    if weight_decay and not any(t in name for t in ("norm", "bias")):    # optimizers.py:84,88 / :180-188
        upd = upd + weight_decay * p
    return p - lr * upd, m2, v2


def adam_tf_step(p, m, v, g, lr, t, beta1=0.9, beta2=0.999, eps=1e-8):
    """tf.train.AdamOptimizer (src/model_fns_tf.py:58-60) ‡: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); p -= lr_t*m/(sqrt(v)+eps)."""
    m2 = beta1 * m + (1 - beta1) * g
    v2 = beta2 * v + (1 - beta2) * g * g
    lr_t = lr * math.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
    return p - lr_t * m2 / (torch.sqrt(v2) + eps), m2, v2


def dalle_train_step(params, m, v, grads, step, hp):
    """get_optimizer (src/optimizers.py:19-104): schedule -> clip (default 1.0) -> mtf Adam.  Returns new state."""
    lr = learning_rate(step, hp)
    clip = hp["gradient_clipping"] if "gradient_clipping" in hp else 1.0  # optimizers.py:27 (explicit null: no clip)
    gn = None
    if clip is not None:
        grads, gn = clip_by_global_norm(grads, clip)                      # optimizers.py:101-102
    new_p, new_m, new_v = {}, {}, {}
    for k in params:
        new_p[k], new_m[k], new_v[k] = adam_mtf_step(
            params[k], m[k], v[k], grads[k], lr, hp.get("beta_1") or 0.9, hp.get("beta_2") or 0.999,
            hp.get("epsilon") or 1e-6, hp.get("weight_decay") or 0.0, name=k)
    return new_p, new_m, new_v, lr, gn
