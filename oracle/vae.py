"""ORACLE (test infrastructure, NOT product code) — CPU restatement of the plain-TF discrete VAE (src/vae_tf/).

PARITY UNPINNED (see oracle/dalle.py header): tf.layers.conv2d / conv2d_transpose / tf.nn.softmax / tf.argmax are
restated from their published semantics (‡), anchored on the reference call sites cited below.
Tensors follow the reference's layouts: activations NHWC, conv kernels HWIO ([kh,kw,cin,cout]), transposed-conv
kernels [kh,kw,cout,cin], codebook [n_hid, K].  Random numbers (Gumbel u, initial weights) are always INPUTS.
"""
import math

import torch
import torch.nn.functional as F


def image_seq_len(image_size, convblocks, stack_factor=1):
    """src/model_fns.py:68."""
    return (image_size // (2 ** len(convblocks))) ** 2 // (stack_factor ** 2)


def _glorot_uniform(shape, fan_in, fan_out, g):
    limit = math.sqrt(6.0 / (fan_in + fan_out))  # ‡ tf glorot_uniform_initializer (default of tf.layers / get_variable)
    return (torch.rand(*shape, generator=g) * 2 - 1) * limit


def init_params(convblocks, num_tokens, input_channels=3, stack_factor=1, seed=0):
    """Variables of src/vae_tf/models.py:81-163 under scope `vae/`."""
    g = torch.Generator().manual_seed(seed)
    p = {}
    cin = input_channels * stack_factor ** 2
    for b, (stack, ch) in enumerate(convblocks):                        # encoder, models.py:88-109
        for i in range(stack):
            pre = f"encoder/block_{b}/layer_{i}/"
            if i == 0:
                p[pre + "conv_downsample/kernel"] = _glorot_uniform((4, 4, cin, ch), 16 * cin, 16 * ch, g)
                p[pre + "conv_downsample/bias"] = torch.zeros(ch)
            else:
                for nm in ("conv_in", "conv_out"):
                    p[pre + nm + "/kernel"] = _glorot_uniform((3, 3, ch, ch), 9 * ch, 9 * ch, g)
                    p[pre + nm + "/bias"] = torch.zeros(ch)
        cin = ch
    n_hid = cin
    p["codebook/codebook"] = _glorot_uniform((n_hid, num_tokens), n_hid, num_tokens, g)   # models.py:111-113
    for b, (stack, ch) in enumerate(reversed(convblocks)):               # decoder, models.py:132-149
        for i in range(stack):
            pre = f"decoder/block_{b}/layer_{i}/"
            if i == 0:
                # tf conv2d_transpose kernel layout [kh, kw, out, in]
                p[pre + "conv_upsample/kernel"] = _glorot_uniform((4, 4, ch, cin), 16 * ch, 16 * cin, g)
                p[pre + "conv_upsample/bias"] = torch.zeros(ch)
            else:
                for nm in ("conv_in", "conv_out"):
                    p[pre + nm + "/kernel"] = _glorot_uniform((3, 3, ch, ch), 9 * ch, 9 * ch, g)
                    p[pre + nm + "/bias"] = torch.zeros(ch)
        cin = ch
    cout = input_channels * stack_factor ** 2
    p["decoder/conv2d/kernel"] = _glorot_uniform((1, 1, cin, cout), cin, cout, g)         # models.py:155
    p["decoder/conv2d/bias"] = torch.zeros(cout)
    return p


def _r(x, bf16):
    return x.to(torch.bfloat16).to(torch.float32) if bf16 else x


def conv2d_same(x, w, b, stride):
    """tf.layers.conv2d(padding="SAME") ‡, x NHWC, w HWIO.  SAME: total pad = max((ceil(H/s)-1)*s + k - H, 0),
    pad_before = total // 2 (the extra pixel, if any, goes after)."""
    N, H, W_, C = x.shape
    kh, kw = w.shape[0], w.shape[1]
    oh, ow = -(-H // stride), -(-W_ // stride)
    ph = max((oh - 1) * stride + kh - H, 0)
    pw = max((ow - 1) * stride + kw - W_, 0)
    xp = F.pad(x.permute(0, 3, 1, 2), (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))
    y = F.conv2d(xp, w.permute(3, 2, 0, 1), b, stride=stride)
    return y.permute(0, 2, 3, 1)


def conv2d_transpose_same(x, w, b, stride=2):
    """tf.layers.conv2d_transpose(k=4, s=2, padding="SAME") ‡ == gradient of the SAME conv w.r.t. its input:
    y[2*i - 1 + kh] += x[i] * w[kh, kw, out, in]   (== torch ConvTranspose2d(k=4, s=2, p=1), no kernel flip)."""
    assert w.shape[0] == 4 and w.shape[1] == 4 and stride == 2
    y = F.conv_transpose2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), b, stride=2, padding=1)
    return y.permute(0, 2, 3, 1)


def space_to_depth(x, s):
    """tf.space_to_depth ‡: channel index = (dy*s + dx)*C + c."""
    N, H, W_, C = x.shape
    x = x.view(N, H // s, s, W_ // s, s, C).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(N, H // s, W_ // s, s * s * C)


def depth_to_space(x, s):
    N, H, W_, C = x.shape
    c = C // (s * s)
    x = x.view(N, H, W_, s, s, c).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(N, H * s, W_ * s, c)


def encoder(p, img, convblocks, bf16=False, stack_factor=1):
    """DiscreteVAE.encoder (src/vae_tf/models.py:81-120) -> fp32 logits [N,h,w,K]."""
    x = _r(img, bf16)                                                    # models.py:82-83
    if stack_factor > 1:
        x = space_to_depth(x, stack_factor)                              # models.py:85-86
    W = (lambda n: _r(p[n], bf16))
    for b, (stack, ch) in enumerate(convblocks):
        for i in range(stack):
            pre = f"encoder/block_{b}/layer_{i}/"
            if i == 0:                                                   # models.py:95 (no activation after it)
                x = _r(conv2d_same(x, W(pre + "conv_downsample/kernel"), W(pre + "conv_downsample/bias"), 2), bf16)
            else:                                                        # models.py:99-109
                out = _r(conv2d_same(x, W(pre + "conv_in/kernel"), W(pre + "conv_in/bias"), 1), bf16)
                out = torch.relu(out)
                out = _r(conv2d_same(out, W(pre + "conv_out/kernel"), W(pre + "conv_out/bias"), 1), bf16)
                x = _r(x + out, bf16)
    return x.to(torch.float32) @ p["codebook/codebook"]                  # models.py:115-118 (fp32 matmul)


def gumbel_softmax(logits, u, temperature=1.0, hard=True):
    """src/vae_tf/layers.py:4-21 with the uniform noise u ~ U[1e-9, 1) passed in."""
    g = -torch.log(-torch.log(u))
    y = torch.softmax((logits + g) / temperature, dim=-1)
    if hard:
        idx = y.argmax(-1)  # ‡ tf.argmax returns the smallest index among ties; torch.argmax on CPU does too
        y_hard = F.one_hot(idx, y.shape[-1]).to(y.dtype)
        y = (y_hard - y).detach() + y                                    # straight-through, layers.py:17-19
    return y


def decoder(p, y, convblocks, bf16=False, stack_factor=1):
    """DiscreteVAE.decoder (src/vae_tf/models.py:123-163) -> fp32 image [N,H,W,C]."""
    x = y @ p["codebook/codebook"].t()                                   # models.py:127 (tied codebook, fp32)
    x = _r(x, bf16)                                                      # models.py:129-130
    W = (lambda n: _r(p[n], bf16))
    for b, (stack, ch) in enumerate(reversed(convblocks)):
        for i in range(stack):
            pre = f"decoder/block_{b}/layer_{i}/"
            if i == 0:                                                   # models.py:139
                x = _r(conv2d_transpose_same(x, W(pre + "conv_upsample/kernel"), W(pre + "conv_upsample/bias")), bf16)
            else:                                                        # models.py:143-153
                out = _r(conv2d_same(x, W(pre + "conv_in/kernel"), W(pre + "conv_in/bias"), 1), bf16)
                out = torch.relu(out)
                out = _r(conv2d_same(out, W(pre + "conv_out/kernel"), W(pre + "conv_out/bias"), 1), bf16)
                x = _r(x + out, bf16)
    x = _r(conv2d_same(x, W("decoder/conv2d/kernel"), W("decoder/conv2d/bias"), 1), bf16)   # models.py:155
    x = x.to(torch.float32)
    if stack_factor > 1:
        x = depth_to_space(x, stack_factor)                              # models.py:160-161
    return x


def forward(p, img, u, convblocks, temperature=1.0, hard=True, bf16=False, stack_factor=1):
    """DiscreteVAE.forward(return_recon_loss=True) (src/vae_tf/models.py:165-184): returns (loss, recon, logits)."""
    logits = encoder(p, img, convblocks, bf16, stack_factor)
    y = gumbel_softmax(logits, u, temperature, hard)
    out = decoder(p, y, convblocks, bf16, stack_factor)
    loss = ((img - out) ** 2).mean()                                     # src/vae_tf/layers.py:24-25
    return loss, out, logits


def encode_tokens(p, img, convblocks, bf16=False, stack_factor=1):
    """src/model_fns.py:72-77: argmax over codes (lowest index on ties ‡), reshaped [B, image_seq_len] row-major."""
    logits = encoder(p, img, convblocks, bf16, stack_factor)
    return logits.argmax(-1).reshape(img.shape[0], -1)


def temperature(step, params):
    """src/model_fns_tf.py:40-45."""
    if params.get("temp_anneal_steps"):
        frac = min(step / params["temp_anneal_steps"], 1.0)
        return params["temp_start"] - frac * (params["temp_start"] - params["temp"])
    t = params.get("temp")
    return 1.0 if t is None else t


def loss_and_grads(p, img, u, convblocks, temperature=1.0, hard=True, bf16=False):
    leaves = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    loss, out, logits = forward(leaves, img, u, convblocks, temperature, hard, bf16)
    loss.backward()
    return loss.detach(), out.detach(), logits.detach(), {k: v.grad for k, v in leaves.items()}
