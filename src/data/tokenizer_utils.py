"""src/data/tokenizer_utils.py of the reference."""
from dalle_mtf_b200.tokenizer import get_tokenizer  # noqa: F401
