"""The lazy-rescaling block algorithm of the experimental attention forward kernel (oracle/attention_blocks.py) equals
plain causal softmax attention, also when the logits are large enough to force rescales."""
import numpy as np
import pytest
import torch

from oracle.attention_blocks import lazy_rescale_attention


@pytest.mark.parametrize("S,dh,mag,scale", [(200, 64, 0.5, 1.0), (300, 64, 8.0, 1.0), (130, 128, 3.0, 0.125)])
def test_lazy_rescaling_matches_softmax_attention(S, dh, mag, scale):
    g = torch.Generator().manual_seed(S + dh)
    bf = lambda x: x.to(torch.bfloat16).float()
    q, k, v = bf(torch.randn(S, dh, generator=g) * mag), bf(torch.randn(S, dh, generator=g)), bf(torch.randn(S, dh, generator=g))
    out, lse, rescales = lazy_rescale_attention(q.numpy(), k.numpy(), v.numpy(), scale)
    s = (q @ k.t()) * scale
    s = s.masked_fill(torch.triu(torch.ones(S, S, dtype=torch.bool), 1), float("-inf"))
    ref, lref = torch.softmax(s, -1) @ v, torch.logsumexp(s, -1)
    assert (torch.from_numpy(out) - ref).abs().max() / ref.abs().max() < 1e-2      # bf16 rounding of P
    assert (torch.from_numpy(lse) - lref).abs().max() < 1e-3
    if mag >= 8.0:
        assert rescales > 0                                                          # the lazy path was exercised
