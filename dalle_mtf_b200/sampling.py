"""Autoregressive sampling with a K/V cache ("next" row N4).

The reference only sketches this path: `DALLE(is_incremental_inference=True, context=...)` blends the current
position's k, v into the cached states at `context.position - 1` and lets that position's query attend to keys up to
itself (src/dalle_mtf/models.py:246-254, 281-285); PREDICT mode raises NotImplementedError (src/model_fns.py:135-136)
and no sampling loop exists.  What is implemented here is that sketch, driven one position at a time over the SAME
parameters and the same kernels' numerics as training:

  per position p:   x = wte[token] + wpe[p]                                     (embed_fwd_at)
  per layer:        q|k|v = LN1(x) Wqkv;  cache[p] = k, v;  a = softmax_{j<=p}(q.k_j) v_j   (attn_decode)
                    x = x + a Wo + o_b;  x = x + relu(LN2(x) W1 + b1) W2 + b2   (the training GEMM, M = batch rows)
  logits = LN_f(x) Wout + bout (fp32); next token = argmax over the allowed id range of logits / T + Gumbel noise

Prompt (text) positions are teacher-forced through the same step function, so there is a single code path; image
positions are restricted to the image-token id range [text_vocab_size, text_vocab_size + image_vocab_size), matching
the ids training builds (src/model_fns.py:117-122).  `temperature == 0` (or None) is greedy decoding.
"""
import torch

from . import ops

BF16, F32, I32 = torch.bfloat16, torch.float32, torch.int32


class DalleSampler:
    def __init__(self, engine):
        self.e = engine
        self._B = None

    def _alloc(self, B):
        if self._B == B:
            return
        e = self.e
        dev, d = e.device, e.d
        mk = lambda *shape, dtype=BF16: torch.empty(*shape, dtype=dtype, device=dev)
        self.kc = [mk(B, e.S, e.H, e.dh) for _ in range(e.L)]
        self.vc = [mk(B, e.S, e.H, e.dh) for _ in range(e.L)]
        self.x = [mk(B, d), mk(B, d)]
        self.ln = mk(B, d)
        self.qkv = mk(B, 3 * d)
        self.att = mk(B, d)
        self.xmid = mk(B, d)
        self.h1 = mk(B, 4 * d)
        self.mean, self.rstd = mk(B, dtype=F32), mk(B, dtype=F32)
        self.logit = mk(B, e.Vpad, dtype=F32)
        self.tok = mk(B, dtype=I32)
        self._B = B

    def step(self, tokens_b, pos):
        """One position for every row: tokens_b int32 [B] -> fp32 logits [B, Vpad] of the NEXT position."""
        e = self.e
        ops.embed_fwd_at(tokens_b, e.W("wte"), e.W("wpe"), self.x[0], pos)
        x_in, x_out = self.x
        for i in range(e.L):
            p = f"l{i}."
            ops.layernorm_fwd(x_in, e.P(p + "ln1_g"), e.P(p + "ln1_b"), self.ln, self.mean, self.rstd, e.ln_eps)
            ops.linear_fwd(self.ln, e.W(p + "wqkv"), self.qkv)
            ops.attn_decode(self.qkv, self.kc[i], self.vc[i], self.att, pos, e.attn_scale)
            ops.linear_fwd(self.att, e.W(p + "wo"), self.xmid, bias=e.P(p + "o_b"), residual=x_in)
            ops.layernorm_fwd(self.xmid, e.P(p + "ln2_g"), e.P(p + "ln2_b"), self.ln, self.mean, self.rstd, e.ln_eps)
            ops.linear_fwd(self.ln, e.W(p + "w1"), self.h1, bias=e.P(p + "b1"), relu=True)
            ops.linear_fwd(self.h1, e.W(p + "w2"), x_out, bias=e.P(p + "b2"), residual=self.xmid)
            x_in, x_out = x_out, x_in
        ops.layernorm_fwd(x_in, e.P("lnf_g"), e.P("lnf_b"), self.ln, self.mean, self.rstd, e.ln_eps)
        B = tokens_b.shape[0]
        ops.gemm(self.ln, e.W("wout"), self.logit, B, e.Vpad, e.d, a_mn=False, b_mn=True, bias=e.P("bout"))
        return self.logit

    # -------------------------------------------------------------------------------------------- CUDA-graph path
    def _step_dev(self, tokens):
        """self.step with the position (and the input token column) read on the device from self.pos_dev."""
        e = self.e
        ops.embed_fwd_at_dev(tokens, e.W("wte"), e.W("wpe"), self.x[0], self.pos_dev)
        x_in, x_out = self.x
        for i in range(e.L):
            p = f"l{i}."
            ops.layernorm_fwd(x_in, e.P(p + "ln1_g"), e.P(p + "ln1_b"), self.ln, self.mean, self.rstd, e.ln_eps)
            ops.linear_fwd(self.ln, e.W(p + "wqkv"), self.qkv)
            ops.attn_decode_dev(self.qkv, self.kc[i], self.vc[i], self.att, self.pos_dev, e.attn_scale)
            ops.linear_fwd(self.att, e.W(p + "wo"), self.xmid, bias=e.P(p + "o_b"), residual=x_in)
            ops.layernorm_fwd(self.xmid, e.P(p + "ln2_g"), e.P(p + "ln2_b"), self.ln, self.mean, self.rstd, e.ln_eps)
            ops.linear_fwd(self.ln, e.W(p + "w1"), self.h1, bias=e.P(p + "b1"), relu=True)
            ops.linear_fwd(self.h1, e.W(p + "w2"), x_out, bias=e.P(p + "b2"), residual=self.xmid)
            x_in, x_out = x_out, x_in
        ops.layernorm_fwd(x_in, e.P("lnf_g"), e.P("lnf_b"), self.ln, self.mean, self.rstd, e.ln_eps)
        ops.gemm(self.ln, e.W("wout"), self.logit, tokens.shape[0], e.Vpad, e.d, a_mn=False, b_mn=True, bias=e.P("bout"))
        return self.logit

    def _capture(self, B, greedy, inv_t):
        """Two graphs over static buffers: the prompt step (teacher forcing: the next token is already in the token
        matrix) and the sampling step (step -> noise -> Gumbel-max into column pos + 1); both end by advancing pos_dev."""
        e = self.e
        key = (B, greedy, inv_t)
        if getattr(self, "_graph_key", None) == key:
            return
        dev = e.device
        self.tokens = torch.zeros(B, e.S, dtype=I32, device=dev)
        self.pos_dev = torch.zeros(1, dtype=I32, device=dev)
        lo, hi = e.text_vocab_size, e.text_vocab_size + e.image_vocab_size
        self.u = None if greedy else torch.empty(B, hi - lo, dtype=F32, device=dev)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):     # eager warm-up: function attributes, tensor maps, lazy allocations
            self._step_dev(self.tokens)
            if self.u is not None:
                self.u.uniform_(1e-9, 1.0)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.g_prompt, self.g_sample = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_prompt):
            self._step_dev(self.tokens)
            ops.incr_i32(self.pos_dev, 1)
        with torch.cuda.graph(self.g_sample):
            logits = self._step_dev(self.tokens)
            if self.u is not None:
                self.u.uniform_(1e-9, 1.0)          # default CUDA generator: graph-safe Philox offsets
            ops.sample_rows_at(logits, self.u, self.tokens, lo, hi, inv_t, self.pos_dev)
            ops.incr_i32(self.pos_dev, 1)
        self._graph_key = key

    @torch.no_grad()
    def generate_graphed(self, text_ids, temperature=1.0):
        """generate() with the per-position step replayed from CUDA graphs (position kept on the device): removes the
        ~45 us of Python / ctypes / launch overhead per kernel that bounds eager generation.  Noise comes from torch's
        default CUDA generator (seed it with torch.manual_seed)."""
        e = self.e
        B, TL = text_ids.shape
        assert TL == e.text_seq_len, f"expected {e.text_seq_len} text positions, got {TL}"
        self._alloc(B)
        greedy = temperature is None or temperature <= 0
        inv_t = 1.0 if greedy else 1.0 / float(temperature)
        self._capture(B, greedy, inv_t)
        self.tokens.zero_()
        self.tokens[:, :TL] = text_ids
        self.pos_dev.zero_()
        for pos in range(e.S - 1):
            (self.g_prompt if pos + 1 < TL else self.g_sample).replay()
        return self.tokens.clone()

    @torch.no_grad()
    def generate(self, text_ids, temperature=1.0, generator=None, return_logits=False):
        """text_ids int32 [B, text_seq_len] (device) -> int32 [B, text_seq_len + image_seq_len]: the prompt followed by
        image_seq_len sampled image-token ids (already offset by text_vocab_size, as in training)."""
        e = self.e
        B, TL = text_ids.shape
        assert TL == e.text_seq_len, f"expected {e.text_seq_len} text positions, got {TL}"
        self._alloc(B)
        out = torch.empty(B, e.S, dtype=I32, device=e.device)
        out[:, :TL] = text_ids
        lo, hi = e.text_vocab_size, e.text_vocab_size + e.image_vocab_size
        greedy = temperature is None or temperature <= 0
        inv_t = 1.0 if greedy else 1.0 / float(temperature)
        kept = [] if return_logits else None
        for pos in range(e.S - 1):
            cur = out[:, pos].contiguous()
            logits = self.step(cur, pos)
            if return_logits:
                kept.append(logits[:, :e.V].clone())
            if pos + 1 < TL:
                continue  # next position is part of the prompt: teacher forcing
            u = None
            if not greedy:
                u = torch.empty(B, hi - lo, dtype=F32, device=e.device).uniform_(1e-9, 1.0, generator=generator)
            ops.sample_rows(logits, u, self.tok, lo, hi, inv_t)
            out[:, pos + 1] = self.tok
        return (out, kept) if return_logits else out
