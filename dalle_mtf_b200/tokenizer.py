"""Tokenizer surface of src/data/tokenizer_utils.py.  The reference loads GPT-2 BPE from the HF hub and appends a
`<|padding|>` token (len 50258, pad id 50257); there is no network here, so unless a local HF cache has the files the
training entry points get a stand-in object with the same two facts the hot path uses: len() and pad_token_id.
Real tokenisation is dataset tooling ("next" row N3), not part of the step."""


class _PaddingOnlyTokenizer:
    pad_token = "<|padding|>"
    pad_token_id = 50257

    def __len__(self):
        return 50258

    def encode(self, text):
        if text == self.pad_token:
            return [self.pad_token_id]
        raise NotImplementedError("GPT-2 BPE files are not available offline; captions must be pre-tokenised")


def get_tokenizer(tokenizer_type=None, from_pretrained=True, add_padding_token=True):
    """src/data/tokenizer_utils.py:4-16."""
    if tokenizer_type is None or (tokenizer_type.lower() in ("hf_gpt2tokenizerfast", "hf_gp2tokenizer") and from_pretrained):
        try:
            from transformers import GPT2TokenizerFast
            tok = GPT2TokenizerFast.from_pretrained("gpt2", local_files_only=True)
            if add_padding_token:
                tok.add_special_tokens({"pad_token": "<|padding|>"})
            if len(tok) != 50258:   # an empty / partial local cache is not the GPT-2 vocabulary
                return _PaddingOnlyTokenizer()
            return tok
        except Exception:  # no local cache / no network
            return _PaddingOnlyTokenizer()
    raise NotImplementedError("TODO: add custom tokenizers")
