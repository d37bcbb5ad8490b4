from dalle_mtf_b200.models import DiscreteVAE  # noqa: F401
