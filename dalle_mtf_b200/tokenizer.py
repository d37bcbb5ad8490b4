"""Tokenizer surface of src/data/tokenizer_utils.py: GPT-2 byte-level BPE + an appended `<|padding|>` token
(len 50258, pad id 50257).

The reference downloads the vocabulary from the HF hub (`from_pretrained('gpt2')`); there is no network here, so the
files come from, in this order:
  1. `vocab_dir` argument / `$DB200_GPT2_DIR`: a directory holding GPT-2's `vocab.json` and `merges.txt` (what
     `GPT2TokenizerFast.save_pretrained` or the original OpenAI release ships) — a complete offline tokenizer;
  2. a local HF cache that already has `gpt2`;
  3. neither: a stand-in object with the two facts the training step uses (len() and pad_token_id); `encode` raises,
     captions must then be pre-tokenised (the TFRecords carry int64 ids anyway, src/input_fns.py:41-52).
Tokenisation is dataset tooling ("next" row N3), not part of the step."""
import os


class _PaddingOnlyTokenizer:
    pad_token = "<|padding|>"
    pad_token_id = 50257

    def __len__(self):
        return 50258

    def encode(self, text):
        if text == self.pad_token:
            return [self.pad_token_id]
        raise NotImplementedError("GPT-2 BPE files are not available: point DB200_GPT2_DIR at a directory with "
                                  "vocab.json and merges.txt, or pre-tokenise the captions")


def _from_files(vocab_dir, fast, add_padding_token):
    from transformers import GPT2Tokenizer, GPT2TokenizerFast
    vocab, merges = os.path.join(vocab_dir, "vocab.json"), os.path.join(vocab_dir, "merges.txt")
    if not (os.path.exists(vocab) and os.path.exists(merges)):
        raise FileNotFoundError(f"{vocab_dir} must contain vocab.json and merges.txt")
    cls = GPT2TokenizerFast if fast else GPT2Tokenizer
    # from_pretrained on a local directory works across transformers 4.x / 5.x (the constructor's file arguments do not)
    tok = cls.from_pretrained(vocab_dir, local_files_only=True)
    if add_padding_token:
        tok.add_special_tokens({"pad_token": "<|padding|>"})       # src/data/tokenizer_utils.py:7-8
    return tok


def get_tokenizer(tokenizer_type=None, from_pretrained=True, add_padding_token=True, vocab_dir=None):
    """src/data/tokenizer_utils.py:4-16 (both spellings of the slow tokenizer's name are accepted: the reference has the
    typo "hf_gp2tokenizer")."""
    name = None if tokenizer_type is None else tokenizer_type.lower()
    if name is not None and not (name in ("hf_gpt2tokenizerfast", "hf_gp2tokenizer", "hf_gpt2tokenizer")
                                 and from_pretrained):
        raise NotImplementedError("TODO: add custom tokenizers")
    fast = name in (None, "hf_gpt2tokenizerfast")
    vocab_dir = vocab_dir or os.environ.get("DB200_GPT2_DIR")
    if vocab_dir:
        return _from_files(vocab_dir, fast, add_padding_token)       # explicit request: errors are raised, not hidden
    try:
        from transformers import GPT2Tokenizer, GPT2TokenizerFast
        tok = (GPT2TokenizerFast if fast else GPT2Tokenizer).from_pretrained("gpt2", local_files_only=True)
        if add_padding_token:
            tok.add_special_tokens({"pad_token": "<|padding|>"})
        if len(tok) != 50258:   # an empty / partial local cache is not the GPT-2 vocabulary
            return _PaddingOnlyTokenizer()
        return tok
    except Exception:  # no local cache / no network
        return _PaddingOnlyTokenizer()
