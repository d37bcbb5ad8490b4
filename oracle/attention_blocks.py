"""ORACLE (test infrastructure only): block-wise restatement of the arithmetic of the experimental pipelined attention
forward kernel (csrc/attn.cu: attn_fwd2_kernel) — 64-key blocks, P rounded to bf16 for the P V product, the output
accumulated across blocks and rescaled LAZILY: a row is only rescaled when its running maximum has moved by more than
2^threshold since the scale it is using.  It validates the algorithm (not the kernel's synchronisation) against plain
softmax attention (src/dalle_mtf/models.py:287-299 with the causal mask of :221-227)."""
import numpy as np
import torch

LOG2E = 1.4426950408889634


def lazy_rescale_attention(q, k, v, scale, block=64, threshold=8.0):
    """q, k, v: float32 numpy [S, dh] (already bf16-representable).  Returns (out [S, dh], lse [S], n_rescales)."""
    S, dh = q.shape
    c1 = np.float32(scale * LOG2E)
    out = np.zeros((S, dh), np.float32)
    lse = np.zeros(S, np.float32)
    rescales = 0
    logits = (q @ k.T).astype(np.float32)
    for qi in range(S):
        q0 = (qi // 128) * 128
        n_kv = (min(S, q0 + 128) + block - 1) // block
        m_run, m_used, l = -np.inf, -np.inf, np.float32(0)
        acc = np.zeros(dh, np.float32)
        for j in range(n_kv):
            ks = np.arange(j * block, min((j + 1) * block, S))
            s = np.where(ks <= qi, logits[qi, ks], -np.inf).astype(np.float32)
            m_new = max(m_run, float(s.max()))
            if j == 0:
                m_used = m_new
            elif (m_new - m_used) * c1 > threshold:
                alpha = np.float32(2.0 ** ((m_used - m_new) * c1))
                acc *= alpha
                l *= alpha
                m_used = m_new
                rescales += 1
            p = np.exp2(s * c1 - np.float32(m_used) * c1).astype(np.float32)
            l += p.sum(dtype=np.float32)
            pb = torch.from_numpy(p).to(torch.bfloat16).float().numpy()
            acc += pb @ v[ks]
            m_run = m_new
        out[qi] = acc / l
        lse[qi] = m_used * scale + np.log(l)
    return out, lse, rescales
