from dalle_mtf_b200.models import DALLE  # noqa: F401
